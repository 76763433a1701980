"""Drop-in host classes for the multi_input_vocoder generator forward.

``MelCodeGenerator`` mirrors the class of the same name in the reference's
``multi_input_vocoder/models_multi_input.py:26-97`` (constructor from ``h``,
weight-normed ``state_dict`` key set, ``eval()``, ``remove_weight_norm()``,
``forward(**kwargs)``), and ``CodeGenerator`` the unit-only parent
(``speech-resynthesis/models.py:125-229``), so the reference's ``inference.py`` /
``inference_server.py`` keep working when this module shadows theirs on
``sys.path``.  All arithmetic happens in the C-ABI CUDA library
(``include/l2s_vocoder.h``); torch is used for parameter storage, device memory
and the current stream only.  There is no CPU path: CPU inputs raise.

Not supported (raises ``NotImplementedError``): autograd / ``.train()`` forward,
``h.text_supervision``, ResBlock2, the F0 / VQ branches, extra conditioning kwargs.
"""
import ctypes as C
import math
import os
import threading

import torch
import torch.nn as nn

try:                                   # package member (lip2speech_unit_b200.models_multi_input)
    from . import _cabi
except ImportError:                    # top-level module: this directory is on sys.path and shadows the reference's
    import _cabi                       # models_multi_input.py (inference.py:28 `from models_multi_input import ...`)

LRELU_SLOPE = 0.1


class AttrDict(dict):
    """dict with attribute access, like the ``h`` the reference builds from its JSON config
    (speech-resynthesis/utils.py:77-80)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


class _WeightNormed(nn.Module):
    """Parameter holder with the key names torch's weight_norm hook produces
    (``weight_g`` / ``weight_v`` / ``bias``); ``fold()`` leaves ``weight`` / ``bias``
    like ``remove_weight_norm`` (speech-resynthesis/models.py:43-47,116-122)."""

    def __init__(self, shape, bias_len, std=None):
        super().__init__()
        v = torch.empty(*shape)
        fan_in = shape[1] * shape[2]
        if std is None:
            bound = 1.0 / math.sqrt(fan_in)
            v.uniform_(-bound, bound)
        else:
            v.normal_(0.0, std)        # init_weights, speech-resynthesis/utils.py:32-35
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1).clone())
        self.weight_v = nn.Parameter(v)
        b = 1.0 / math.sqrt(fan_in)
        self.bias = nn.Parameter(torch.empty(bias_len).uniform_(-b, b))

    @property
    def folded(self):
        return "weight_v" not in self._parameters

    def plain_weight(self):
        if self.folded:
            return self.weight.detach()
        v = self.weight_v.detach()
        g = self.weight_g.detach()
        norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        return v * (g / norm)

    def fold(self):
        if self.folded:
            raise ValueError("weight_norm of this layer was already removed")
        w = self.plain_weight()
        bias = self.bias
        del self._parameters["weight_g"], self._parameters["weight_v"], self._parameters["bias"]
        self.weight = nn.Parameter(w)      # same key order as torch: weight before bias is not required
        self.bias = bias


class _Plain(nn.Module):
    def __init__(self, weight_shape, bias_len=None, normal=False):
        super().__init__()
        w = torch.empty(*weight_shape)
        if normal:
            w.normal_()
        else:
            fan_in = weight_shape[1] * (weight_shape[2] if len(weight_shape) > 2 else 1)
            bound = 1.0 / math.sqrt(fan_in)
            w.uniform_(-bound, bound)
        self.weight = nn.Parameter(w)
        if bias_len:
            self.bias = nn.Parameter(torch.empty(bias_len).uniform_(-0.05, 0.05))


class _ResBlockParams(nn.Module):
    def __init__(self, channels, kernel, dilations):
        super().__init__()
        self.convs1 = nn.ModuleList([_WeightNormed((channels, channels, kernel), channels, 0.01) for _ in dilations])
        self.convs2 = nn.ModuleList([_WeightNormed((channels, channels, kernel), channels, 0.01) for _ in dilations])


class _Engine:
    """One finalized C handle (weights resident on one device in one precision)."""

    def __init__(self, lib, handle, device_index, precision):
        self.lib = lib
        self.handle = handle
        self.device_index = device_index
        self.precision = precision
        self.hop = lib.l2s_hop(handle)
        self.workspaces = {}

    def close(self):
        if self.handle is not None:
            self.lib.l2s_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_PRECISIONS = {"fp32": _cabi.PREC_FP32, "bf16": _cabi.PREC_BF16, "tf32": _cabi.PREC_TF32}


class _GeneratorBase(nn.Module):
    _variant = None

    def __init__(self, h):
        super().__init__()
        self.h = h
        if str(h.resblock) != "1":
            raise NotImplementedError("only ResBlock1 configs are on the accelerated path")
        for key in ("lambda_commit", "lambda_commit_code", "f0_quantizer_path", "f0"):
            if h.get(key, None):
                raise NotImplementedError(f"h.{key}: the F0 / VQ branches are outside the accelerated path")
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        c0 = h.upsample_initial_channel
        self._in_dim = getattr(h, "model_in_dim", 128)
        self.conv_pre = _WeightNormed((c0, self._in_dim, 7), c0)
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        ch = c0
        for u, k in zip(h.upsample_rates, h.upsample_kernel_sizes):
            self.ups.append(_WeightNormed((ch, ch // 2, k), ch // 2, 0.01))
            ch //= 2
            for rk, dil in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(_ResBlockParams(ch, rk, dil))
        self.conv_post = _WeightNormed((1, ch, 7), 1, 0.01)
        self.dict = _Plain((h.num_embeddings, h.embedding_dim), normal=True)
        self.multispkr = h.get("multispkr", None)
        # bf16 tensor-core mode is the default (the path this library exists for: ~46 dB SNR / <= 2e-2 max-abs against the
        # fp32 reference).  "tf32" matches what the reference itself computes on an Ampere-or-later
        # GPU (cuDNN TF32 convolutions are on by default in torch), "fp32" is the CUDA-core reference mode.
        self.precision = os.environ.get("L2S_PRECISION", h.get("precision", "bf16"))
        # Out-of-range ids: always detected (sticky device flag).  Default: the NEXT forward (or check_index_errors())
        # raises IndexError -- late, like the asynchronous device assert of the CUDA reference, and without a
        # synchronisation; strict: synchronise after every forward and raise immediately, like the CPU reference.
        self.strict_index_check = bool(int(os.environ.get("L2S_STRICT_INDEX", "0")))
        self._engines = {}
        self._lock = threading.Lock()          # guards the engine table only
        self._dev_locks = {}                   # one lock per device: forwards on different GPUs run concurrently
        self._folded = False

    # ------------------------------------------------------------ reference surface
    def remove_weight_norm(self):
        """Fold w = g * v / ||v|| into plain weights (Generator.remove_weight_norm,
        speech-resynthesis/models.py:116-122).  This is also the repack point: the
        next forward uploads polyphase / tap-major bf16 copies to the device."""
        for m in self._weight_normed():
            m.fold()
        self._folded = True
        self._drop_engines()

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._drop_engines()
        return out

    def set_precision(self, precision: str):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.precision = precision

    def forward(self, **kwargs):
        raise NotImplementedError

    # ------------------------------------------------------------------ internals
    def _weight_normed(self):
        yield self.conv_pre
        yield self.conv_post
        for u in self.ups:
            yield u
        for rb in self.resblocks:
            for c in rb.convs1:
                yield c
            for c in rb.convs2:
                yield c

    def _drop_engines(self):
        with self._lock:
            for e in self._engines.values():
                e.close()
            self._engines = {}

    def _plain_weights(self):
        w = {}
        w["conv_pre.weight"], w["conv_pre.bias"] = self.conv_pre.plain_weight(), self.conv_pre.bias.detach()
        w["conv_post.weight"], w["conv_post.bias"] = self.conv_post.plain_weight(), self.conv_post.bias.detach()
        for i, u in enumerate(self.ups):
            w[f"ups.{i}.weight"], w[f"ups.{i}.bias"] = u.plain_weight(), u.bias.detach()
        for n, rb in enumerate(self.resblocks):
            for grp, convs in (("convs1", rb.convs1), ("convs2", rb.convs2)):
                for m, c in enumerate(convs):
                    w[f"resblocks.{n}.{grp}.{m}.weight"] = c.plain_weight()
                    w[f"resblocks.{n}.{grp}.{m}.bias"] = c.bias.detach()
        w["dict.weight"] = self.dict.weight.detach()
        return w

    def _config(self) -> _cabi.Config:
        h = self.h
        cfg = _cabi.Config()
        cfg.variant = self._variant
        cfg.precision = _PRECISIONS[self.precision]
        cfg.n_ups = len(h.upsample_rates)
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cfg.up_rates[i], cfg.up_ksizes[i] = int(u), int(k)
        cfg.up_init_ch = int(h.upsample_initial_channel)
        cfg.n_rk = len(h.resblock_kernel_sizes)
        n_dil = {len(d) for d in h.resblock_dilation_sizes}
        if len(n_dil) != 1:
            raise NotImplementedError("every resblock needs the same number of dilations")
        cfg.n_dil = n_dil.pop()
        for j, (rk, dil) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            cfg.rk_sizes[j] = int(rk)
            for m, d in enumerate(dil):
                cfg.rk_dils[j][m] = int(d)
        cfg.num_embeddings = int(h.num_embeddings)
        cfg.embedding_dim = int(h.embedding_dim)
        cfg.multispkr = 1 if self.multispkr else 0
        cfg.model_in_dim = int(self._in_dim)
        self._fill_variant(cfg)
        return cfg

    def _fill_variant(self, cfg):
        raise NotImplementedError

    def _engine(self, device: torch.device) -> _Engine:
        key = (device.index if device.index is not None else torch.cuda.current_device(), self.precision)
        with self._lock:
            eng = self._engines.get(key)
        if eng is not None:
            return eng
        lib = _cabi.load()
        cfg = self._config()
        handle = C.c_void_p()
        st = lib.l2s_create(C.byref(cfg), C.byref(handle))
        if st != _cabi.OK:
            msg = _cabi.last_error(lib, handle)
            lib.l2s_destroy(handle)
            if st == _cabi.ERR_UNSUPPORTED:
                raise NotImplementedError(msg)
            raise RuntimeError(msg)
        try:
            keep = []
            for name, t in self._plain_weights().items():
                t = t.detach().to("cpu", torch.float32).contiguous()
                keep.append(t)
                _cabi.raise_for(lib, handle, lib.l2s_set_weight(handle, name.encode(), t.data_ptr(), t.numel()))
            _cabi.raise_for(lib, handle, lib.l2s_finalize(handle, key[0]))
        except Exception:
            lib.l2s_destroy(handle)
            raise
        eng = _Engine(lib, handle, key[0], self.precision)
        with self._lock:
            self._engines[key] = eng
        return eng

    def _check_device(self, t, name):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if not t.is_cuda:
            raise RuntimeError(f"{name} is on {t.device}: this generator runs on a B200 only (no CPU fallback)")

    def _run(self, code, mel, spkr, frames, want_i16=False, out=None, out16=None):
        if self.training:
            raise NotImplementedError("the accelerated generator is inference-only: call .eval() first")
        device = code.device
        batch, units = code.shape
        with self._lock_for(device):
            eng = self._engine(device)
            lib = eng.lib
            stream = torch.cuda.current_stream(device)
            need = lib.l2s_workspace_bytes(eng.handle, batch, frames)
            if need < 0:
                raise RuntimeError("l2s_workspace_bytes failed")
            wkey = stream.cuda_stream
            ws = eng.workspaces.get(wkey)
            if ws is None or ws.numel() < need:
                eng.workspaces[wkey] = None
                ws = torch.empty(int(need), dtype=torch.uint8, device=device)
                eng.workspaces[wkey] = ws
            if out16 is not None:
                if (out16.dtype != torch.int16 or not out16.is_cuda or not out16.is_contiguous()
                        or out16.numel() != batch * eng.hop * frames):
                    raise RuntimeError(f"out16 must be a contiguous int16 CUDA tensor of {batch * eng.hop * frames} samples")
            elif out is None:
                out = torch.empty((batch, 1, eng.hop * frames), dtype=torch.float32, device=device)
            elif (out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous()
                  or tuple(out.shape) != (batch, 1, eng.hop * frames)):
                raise RuntimeError(f"out must be a contiguous float32 CUDA tensor of shape {(batch, 1, eng.hop * frames)}")
            mel_ptr, mel_tag = None, _cabi.F32
            if mel is not None:
                mel_tag = {torch.float32: _cabi.F32, torch.float16: _cabi.F16, torch.bfloat16: _cabi.BF16}[mel.dtype]
                mel_ptr = mel.data_ptr()
            spk_ptr = spkr.data_ptr() if spkr is not None else None
            with torch.cuda.device(device):
                if out16 is not None:              # int16 only, into the caller's buffer (no fp32 copy is written)
                    st = lib.l2s_forward_i16(eng.handle, stream.cuda_stream, code.data_ptr(), mel_ptr, mel_tag, spk_ptr,
                                             batch, units, frames, None, out16.data_ptr(), ws.data_ptr(), ws.numel())
                    _cabi.raise_for(lib, eng.handle, st)
                    if self.strict_index_check:
                        stream.synchronize()
                        _cabi.raise_for(lib, eng.handle, lib.l2s_poll_index_error(eng.handle))
                    return out16
                if want_i16:
                    out16 = torch.empty((batch, eng.hop * frames), dtype=torch.int16, device=device)
                    st = lib.l2s_forward_i16(eng.handle, stream.cuda_stream, code.data_ptr(), mel_ptr, mel_tag, spk_ptr,
                                             batch, units, frames, out.data_ptr(), out16.data_ptr(), ws.data_ptr(),
                                             ws.numel())
                else:
                    out16 = None
                    st = lib.l2s_forward(eng.handle, stream.cuda_stream, code.data_ptr(), mel_ptr, mel_tag, spk_ptr,
                                         batch, units, frames, out.data_ptr(), ws.data_ptr(), ws.numel())
            _cabi.raise_for(lib, eng.handle, st)
            if self.strict_index_check:
                stream.synchronize()
                _cabi.raise_for(lib, eng.handle, lib.l2s_poll_index_error(eng.handle))
        return (out, out16) if want_i16 else out

    def _lock_for(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        with self._lock:
            lk = self._dev_locks.get(idx)
            if lk is None:
                lk = self._dev_locks[idx] = threading.Lock()
        return lk

    def check_index_errors(self, device=None):
        """Synchronise the current stream of `device` and raise IndexError if a forward issued on it saw an out-of-range
        unit / speaker id (what the reference raises on the CPU; on CUDA it device-asserts).  Without this call the
        error surfaces at the next forward on the same device.  (Stream, not device: a device-wide synchronize could
        collide with another host thread capturing a CUDA graph.)"""
        device = torch.device(device if device is not None else "cuda")
        torch.cuda.current_stream(device).synchronize()
        with self._lock:
            engines = [e for (idx, _), e in self._engines.items()
                       if idx == (device.index if device.index is not None else torch.cuda.current_device())]
        for eng in engines:
            _cabi.raise_for(eng.lib, eng.handle, eng.lib.l2s_poll_index_error(eng.handle))

    def launch_count(self, batch, frames, device=None):
        device = torch.device(device if device is not None else "cuda")
        eng = self._engine(device)
        return eng.lib.l2s_launch_count(eng.handle, batch, frames)

    def debug_tap(self, name, shape, device=None):
        device = torch.device(device if device is not None else "cuda")
        eng = self._engine(device)
        out = torch.empty(shape, dtype=torch.float32)
        _cabi.raise_for(eng.lib, eng.handle, eng.lib.l2s_debug_tap(eng.handle, name.encode(), out.data_ptr(), out.numel()))
        return out


class MelCodeGenerator(_GeneratorBase):
    """units (B,U) int64 + mel (B,80,T=2U) + speaker embedding (B,256) -> waveform (B,1,160*T)."""

    _variant = _cabi.VARIANT_MULTI_INPUT

    def __init__(self, h):
        text_supervision = h.text_supervision      # AttributeError when absent, like models_multi_input.py:31
        super().__init__(h)
        self.text_supervision = text_supervision
        if self.text_supervision:
            raise NotImplementedError("h.text_supervision is outside the accelerated path")
        e = h.embedding_dim
        embedder_dim = h.get("embedder_dim", None)
        if self.multispkr and not embedder_dim:
            raise NotImplementedError("multi-input conditioning with a speaker-id table is outside the accelerated path")
        if embedder_dim:
            self.spkr = _Plain((e, embedder_dim), e)
        self._embedder_dim = embedder_dim
        self.layer = nn.ModuleList([_Plain((e, e, 4), e)])     # keys layer.0.weight / layer.0.bias
        self.fc = _Plain((e, e), e)
        self.num_mels = int(h.get("num_mels", 80))

    def _fill_variant(self, cfg):
        cfg.num_mels = self.num_mels
        cfg.spk_dim = int(self._embedder_dim or 0)
        cfg.num_speakers = 0

    def _plain_weights(self):
        w = super()._plain_weights()
        w["layer.0.weight"], w["layer.0.bias"] = self.layer[0].weight.detach(), self.layer[0].bias.detach()
        w["fc.weight"], w["fc.bias"] = self.fc.weight.detach(), self.fc.bias.detach()
        if self.multispkr:
            w["spkr.weight"], w["spkr.bias"] = self.spkr.weight.detach(), self.spkr.bias.detach()
        return w

    def _prepare(self, kwargs):
        mel = kwargs["mel"]                       # KeyError when absent, like models_multi_input.py:65
        code = kwargs["code"]
        spkr = kwargs["spkr"] if self.multispkr else None
        for k in kwargs:
            if k not in ("spkr", "code", "mel", "t_label"):
                raise NotImplementedError(f"extra conditioning '{k}' is outside the accelerated path")
        self._check_device(code, "code")
        self._check_device(mel, "mel")
        if code.dtype != torch.int64:
            raise TypeError("code must be int64 unit ids")
        if code.dim() != 2 or mel.dim() != 3:
            raise RuntimeError("expected code (B,U) and mel (B,num_mels,T)")
        if mel.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise TypeError("mel must be float32 / float16 / bfloat16")
        if mel.shape[0] != code.shape[0] or mel.shape[1] != self.num_mels:
            raise RuntimeError(f"Sizes of tensors must match: mel {tuple(mel.shape)} vs code {tuple(code.shape)}")
        frames = mel.shape[2]
        if frames != 2 * code.shape[1]:
            raise RuntimeError(
                f"Sizes of tensors must match except in dimension 1. Expected size {frames} but got size "
                f"{2 * code.shape[1]} for tensor number 1 in the list.")
        if spkr is not None:
            self._check_device(spkr, "spkr")
            if spkr.dim() != 2 or spkr.shape != (code.shape[0], self._embedder_dim):
                raise RuntimeError(f"spkr must be (B,{self._embedder_dim}), got {tuple(spkr.shape)}")
            spkr = spkr.to(torch.float32).contiguous()
        return code.contiguous(), mel.contiguous(), spkr, frames

    @torch.no_grad()
    def forward(self, **kwargs):
        code, mel, spkr, frames = self._prepare(kwargs)
        return self._run(code, mel, spkr, frames)

    @torch.no_grad()
    def forward_into(self, out, **kwargs):
        """Same forward, written into a caller-owned (B,1,L) float32 device tensor (no allocation per call: what a
        pipelined caller wants, see dispatch.HostPipeline).  Returns out."""
        code, mel, spkr, frames = self._prepare(kwargs)
        return self._run(code, mel, spkr, frames, out=out)

    @torch.no_grad()
    def forward_int16_into(self, out16, **kwargs):
        """int16 waveform only (inference.py:79-81 on the device), written into a caller-owned int16 device tensor of
        B * hop * T samples: half the device->host bytes of the float waveform, no allocation per call."""
        code, mel, spkr, frames = self._prepare(kwargs)
        return self._run(code, mel, spkr, frames, out16=out16)

    @torch.no_grad()
    def forward_int16(self, **kwargs):
        """Same forward plus the int16 waveform callers derive on the host
        (inference.py:79-81), produced on the device.  Returns (float (B,1,L), int16 (B,L))."""
        code, mel, spkr, frames = self._prepare(kwargs)
        return self._run(code, mel, spkr, frames, want_i16=True)


class CodeGenerator(_GeneratorBase):
    """Unit-only parent: units (B,U) int64 + speaker id (B,1) int64 -> waveform (B,1,hop*U)."""

    _variant = _cabi.VARIANT_UNIT_ONLY

    def __init__(self, h):
        super().__init__(h)
        if self.multispkr:
            self.spkr = _Plain((200, h.embedding_dim), normal=True)   # Embedding(200, E), models.py:132-133

    def _fill_variant(self, cfg):
        cfg.num_mels = 0
        cfg.spk_dim = 0
        cfg.num_speakers = 200 if self.multispkr else 0

    def _plain_weights(self):
        w = super()._plain_weights()
        if self.multispkr:
            w["spkr.weight"] = self.spkr.weight.detach()
        return w

    @torch.no_grad()
    def forward(self, **kwargs):
        code = kwargs["code"]
        spkr = kwargs["spkr"] if self.multispkr else None
        for k in kwargs:
            if k not in ("spkr", "code"):
                raise NotImplementedError(f"extra conditioning '{k}' is outside the accelerated path")
        self._check_device(code, "code")
        if code.dtype != torch.int64 or code.dim() != 2:
            raise TypeError("code must be int64 (B,U)")
        if spkr is not None:
            self._check_device(spkr, "spkr")
            if spkr.dtype != torch.int64 or spkr.numel() != code.shape[0]:
                raise RuntimeError("spkr must be int64 speaker ids of shape (B,1)")
            spkr = spkr.reshape(-1).contiguous()
        return self._run(code.contiguous(), None, spkr, code.shape[1])
