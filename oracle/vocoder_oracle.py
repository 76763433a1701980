"""CPU oracle for the multi_input_vocoder generator forward pass.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lip2speech-unit_b200/`` imports this
file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may.  The product path has no CPU
fallback.

This is a functional (state-dict driven, module-free) restatement of the
reference arithmetic.  The reference's math lives in un-vendored
``torch==1.13.1`` (requirements.txt:82) operator calls; the restatement calls the
same ATen operators (``conv1d`` / ``conv_transpose1d`` / ``linear`` /
``embedding``) from the torch in this image, so it follows the reference op by
op:

  * MelCodeGenerator.forward      multi_input_vocoder/models_multi_input.py:60-97
  * Generator.forward             speech-resynthesis/models.py:98-114
  * ResBlock1.forward             speech-resynthesis/models.py:34-41
  * CodeGenerator._upsample       speech-resynthesis/models.py:158-177
  * CodeGenerator.forward (a15)   speech-resynthesis/models.py:179-229
  * remove_weight_norm            speech-resynthesis/models.py:43-47,116-122
  * get_padding                   speech-resynthesis/utils.py:44-45

Parity pin: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the pin is the reference classes themselves, run in
the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference`` unmodified, loads the state dict produced by
``init_state_dict`` below with ``strict=True`` and stores fp64 outputs under
``tests/golden/``).  ``tests/test_oracle.py`` checks this restatement against
those vectors.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # speech-resynthesis/models.py:13


class HParams(dict):
    """dict with attribute access, like the reference's AttrDict (utils.py:77-80)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


# The generator hyper-parameters of configs/lrs3/multi_input{,_aug}.json:9-27.
SHIPPED_CONFIG = dict(
    resblock="1",
    upsample_rates=[5, 4, 2, 2, 2],
    upsample_kernel_sizes=[11, 8, 4, 4, 4],
    upsample_initial_channel=512,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    num_embeddings=200,
    embedding_dim=128,
    model_in_dim=336,
    embedder_dim=256,
    multispkr="_",
    num_mels=80,
    sampling_rate=16000,
    code_hop_size=320,
    mel_hop_size=160,
    seed=1234,
    text_supervision=False,
)

# Upstream speech-resynthesis unit-only variant.
UNIT_ONLY_CONFIG = dict(
    resblock="1",
    upsample_rates=[5, 4, 4, 2, 2],
    upsample_kernel_sizes=[11, 8, 8, 4, 4],
    upsample_initial_channel=512,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    num_embeddings=200,
    embedding_dim=128,
    model_in_dim=256,
    multispkr="_",
    sampling_rate=16000,
    code_hop_size=320,
    seed=1234,
    text_supervision=False,
)


def shipped_config(**over) -> HParams:
    h = HParams(SHIPPED_CONFIG)
    h.update(over)
    return h


def unit_only_config(**over) -> HParams:
    h = HParams(UNIT_ONLY_CONFIG)
    h.update(over)
    return h


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    # speech-resynthesis/utils.py:44-45
    return int((kernel_size * dilation - dilation) / 2)


# --------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------

def _weight_normed_keys(h) -> list:
    """Names of the modules the reference wraps in weight_norm (models.py:78-94,20-31)."""
    names = ["conv_pre", "conv_post"]
    names += [f"ups.{i}" for i in range(len(h["upsample_rates"]))]
    nk = len(h["resblock_kernel_sizes"])
    for i in range(len(h["upsample_rates"])):
        for j in range(nk):
            n = i * nk + j
            for m in range(len(h["resblock_dilation_sizes"][j])):
                names.append(f"resblocks.{n}.convs1.{m}")
                names.append(f"resblocks.{n}.convs2.{m}")
    return names


def init_state_dict(h, seed: int = 1234, style: str = "ref",
                    unit_only: bool = False) -> Dict[str, torch.Tensor]:
    """Deterministic random weights with the reference's key set and shapes.

    style "ref"     : the reference's init scale -- N(0, 0.01) on ups / resblocks /
                      conv_post (utils.py:32-35), PyTorch-default-like uniform
                      elsewhere.  The output is bias dominated.
    style "trained" : variances scaled so activations stay O(1) through the
                      stack and the tanh output has real dynamic range; weight_g is
                      perturbed away from ||v|| so weight-norm folding is exercised.

    The key set is the weight-normed checkpoint format (``*.weight_g`` /
    ``*.weight_v``) that ``load_state_dict`` receives in inference_server.py:113-118.
    The values are not those torch.manual_seed(1234) + the reference constructor
    would give (that consumes the RNG in module-construction order); the same
    dict is loaded into the reference classes when goldens are made.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    trained = style == "trained"

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def uniform(*shape, bound=1.0):
        return (torch.rand(*shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    def add_wn(name, shape, fan_in, std_ref, gain=1.0):
        # weight_norm over all dims but 0 (Conv1d: per out-channel; ConvTranspose1d:
        # dim 0 is the *input* channel).
        if trained:
            v = randn(*shape, std=gain / math.sqrt(fan_in))
        elif std_ref is None:
            v = uniform(*shape, bound=1.0 / math.sqrt(fan_in))
        else:
            v = randn(*shape, std=std_ref)
        norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (len(shape) - 1)))
        gg = norm.clone()
        if trained:
            gg = gg * (0.75 + 0.5 * torch.rand(gg.shape, generator=g))
        sd[name + ".weight_g"] = gg
        sd[name + ".weight_v"] = v
        nb = shape[1] if name.startswith("ups.") else shape[0]
        sd[name + ".bias"] = uniform(nb, bound=(0.05 if trained else 1.0 / math.sqrt(fan_in)))

    c0 = h["upsample_initial_channel"]
    cin = h.get("model_in_dim", 128)
    add_wn("conv_pre", (c0, cin, 7), cin * 7, None, gain=0.35)
    ch = c0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        # ConvTranspose1d weight (C_in, C_out, k); each output sample sees about k/u taps
        add_wn(f"ups.{i}", (ch, ch // 2, k), ch * k / u, 0.01, gain=1.4)
        ch //= 2
        for j, (rk, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            n = i * len(h["resblock_kernel_sizes"]) + j
            for m in range(len(dil)):
                add_wn(f"resblocks.{n}.convs1.{m}", (ch, ch, rk), ch * rk, 0.01, gain=1.2)
                add_wn(f"resblocks.{n}.convs2.{m}", (ch, ch, rk), ch * rk, 0.01, gain=0.6)
    add_wn("conv_post", (1, ch, 7), ch * 7, 0.01, gain=0.25)

    e = h["embedding_dim"]
    sd["dict.weight"] = randn(h["num_embeddings"], e)
    if unit_only:
        sd["spkr.weight"] = randn(200, e)  # CodeGenerator: Embedding(200, e), models.py:132-133
        return sd
    ed = h.get("embedder_dim", None)
    if ed:
        sd["spkr.weight"] = uniform(e, ed, bound=(2.0 if trained else 1.0) / math.sqrt(ed))
        sd["spkr.bias"] = uniform(e, bound=1.0 / math.sqrt(ed))
    elif h.get("multispkr", None):
        sd["spkr.weight"] = randn(h.get("num_speakers", 200), e)
    sd["layer.0.weight"] = uniform(e, e, 4, bound=(2.0 if trained else 1.0) / math.sqrt(e * 2))
    sd["layer.0.bias"] = uniform(e, bound=1.0 / math.sqrt(e * 2))
    sd["fc.weight"] = uniform(e, e, bound=(2.0 if trained else 1.0) / math.sqrt(e))
    sd["fc.bias"] = uniform(e, bound=1.0 / math.sqrt(e))
    return sd


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """w = v * (g / ||v||), norm over every dim but 0 -- what remove_weight_norm
    leaves behind (models.py:43-47,116-122).  Computed in the dtype of v (fp32 in
    a checkpoint), as torch._weight_norm does when the reference folds."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            continue
        if k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            gg = sd[base + ".weight_g"]
            norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
            out[base + ".weight"] = v * (gg / norm)
        else:
            out[k] = v
    return out


# --------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------

def _broadcast_over_time(signal: torch.Tensor, frames: int) -> torch.Tensor:
    """CodeGenerator._upsample (models.py:158-177): repeat each condition step
    frames // cond_length times; refuse lengths that do not divide."""
    if signal.dim() == 2:
        signal = signal.unsqueeze(2)
    elif signal.dim() != 3:
        signal = signal.reshape(-1, 1, 1)
    b, c, n = signal.shape
    rep = frames // n
    if rep == 0 or (frames - n * rep) // rep > 0:
        raise NotImplementedError("Padding condition signal - misalignment between condition features.")
    return signal.unsqueeze(3).expand(b, c, n, rep).reshape(b, c, n * rep)


def _tap(taps: Optional[dict], name: str, x: torch.Tensor):
    if taps is not None:
        taps[name] = x.detach().clone()


def hifigan_stack(w: Dict[str, torch.Tensor], h, x: torch.Tensor,
                  taps: Optional[dict] = None) -> torch.Tensor:
    """Generator.forward (models.py:98-114) with ResBlock1 (models.py:34-41) inlined."""
    if str(h["resblock"]) != "1":
        raise NotImplementedError("only ResBlock1 configs are covered")
    nk = len(h["resblock_kernel_sizes"])
    x = F.conv1d(x, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    _tap(taps, "conv_pre", x)
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, w[f"ups.{i}.weight"], w[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        _tap(taps, f"ups.{i}", x)
        total = None
        for j, (rk, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            n = i * nk + j
            y = x
            for m, d in enumerate(dil):
                t = F.leaky_relu(y, LRELU_SLOPE)
                t = F.conv1d(t, w[f"resblocks.{n}.convs1.{m}.weight"], w[f"resblocks.{n}.convs1.{m}.bias"],
                             dilation=d, padding=get_padding(rk, d))
                t = F.leaky_relu(t, LRELU_SLOPE)
                t = F.conv1d(t, w[f"resblocks.{n}.convs2.{m}.weight"], w[f"resblocks.{n}.convs2.{m}.bias"],
                             dilation=1, padding=get_padding(rk, 1))
                y = t + y
            _tap(taps, f"resblocks.{n}", y)
            total = y if total is None else total + y
        x = total / nk
        _tap(taps, f"mrf.{i}", x)
    x = F.leaky_relu(x)  # default slope 0.01, models.py:110
    x = F.conv1d(x, w["conv_post.weight"], w["conv_post.bias"], padding=3)
    return torch.tanh(x)


def mel_code_generator_forward(w: Dict[str, torch.Tensor], h, code: torch.Tensor, mel: torch.Tensor,
                               spkr: torch.Tensor, dtype=torch.float64,
                               taps: Optional[dict] = None) -> torch.Tensor:
    """MelCodeGenerator.forward in eval mode (models_multi_input.py:60-97).

    ``w`` holds folded plain weights (``fold_weight_norm``).  code int64 (B,U),
    mel float (B,80,T), spkr float (B,256).  Returns (B,1,prod(rates)*T) in ``dtype``.
    """
    w = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in w.items()}
    if code.dtype not in (torch.int64, torch.int32):
        raise TypeError("code must be an integer tensor")
    emb = F.embedding(code.long(), w["dict.weight"])                     # :67 (B,U,E)
    _tap(taps, "embed", emb)
    y = F.conv_transpose1d(emb.transpose(1, 2), w["layer.0.weight"], w["layer.0.bias"],
                           stride=2, padding=1)                           # :39-42,68
    y = F.gelu(y)                                                         # exact erf GELU
    y = F.linear(y.transpose(1, 2), w["fc.weight"], w["fc.bias"])         # :69-70 (dropout = identity)
    y = y.transpose(1, 2)                                                 # :71 (B,E,T)
    _tap(taps, "code_feats", y)
    x = torch.cat([mel.to(dtype), y], dim=1)                              # :73 raises if 2U != T
    if h.get("multispkr", None):
        if h.get("embedder_dim", None):
            s = F.linear(spkr.to(dtype), w["spkr.weight"], w["spkr.bias"])  # :37,80
        else:
            s = F.embedding(spkr.long(), w["spkr.weight"])
            if s.dim() == 3:
                s = s.transpose(1, 2)
        s = _broadcast_over_time(s, x.shape[-1])                          # :81
        x = torch.cat([x, s], dim=1)                                      # :82
    _tap(taps, "cond", x)
    return hifigan_stack(w, h, x, taps)


def code_generator_forward(w: Dict[str, torch.Tensor], h, code: torch.Tensor, spkr: torch.Tensor,
                           dtype=torch.float64, taps: Optional[dict] = None) -> torch.Tensor:
    """Unit-only parent CodeGenerator.forward (models.py:179-229; no f0 / VQ).
    code int64 (B,U), spkr int64 (B,1) speaker ids."""
    w = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in w.items()}
    x = F.embedding(code.long(), w["dict.weight"]).transpose(1, 2)        # :188
    if h.get("multispkr", None):
        s = F.embedding(spkr.long(), w["spkr.weight"]).transpose(1, 2)    # :214
        s = _broadcast_over_time(s, x.shape[-1])
        x = torch.cat([x, s], dim=1)
    _tap(taps, "cond", x)
    return hifigan_stack(w, h, x, taps)


# --------------------------------------------------------------------------
# synthetic inputs and metrics
# --------------------------------------------------------------------------

def synthetic_inputs(batch: int, frames: int, seed: int = 52, n_units: int = 200, n_mels: int = 80,
                     spk_dim: int = 256):
    """code ~ U[0,200) int64 (B,T/2); mel ~ N(-5.5,1.8^2) clamped to [-11.52,0.5];
    spkr = |N(0,1)| L2-normalised.  Seed 52 is the inference scripts' seed
    (inference_server.py:127)."""
    assert frames % 2 == 0
    g = torch.Generator().manual_seed(seed)
    code = torch.randint(0, n_units, (batch, frames // 2), generator=g, dtype=torch.int64)
    mel = (torch.randn(batch, n_mels, frames, generator=g) * 1.8 - 5.5).clamp_(-11.52, 0.5)
    spkr = torch.randn(batch, spk_dim, generator=g).abs_()
    spkr = spkr / spkr.norm(dim=1, keepdim=True)
    return code, mel, spkr


def snr_db(ref: torch.Tensor, test: torch.Tensor) -> float:
    ref = ref.double().flatten()
    err = test.double().flatten() - ref
    den = float((err * err).sum())
    num = float((ref * ref).sum())
    if den == 0.0:
        return float("inf")
    return 10.0 * math.log10(num / den)


def max_abs(ref: torch.Tensor, test: torch.Tensor) -> float:
    return float((test.double().flatten() - ref.double().flatten()).abs().max())


def algorithmic_flops_per_frame(h, unit_only: bool = False) -> float:
    """2*MACs of the mathematical convolutions per input frame."""
    c0 = h["upsample_initial_channel"]
    fl = 2.0 * h.get("model_in_dim", 128) * c0 * 7
    rate, ch = 1, c0
    for u, k in zip(h["upsample_rates"], h["upsample_kernel_sizes"]):
        fl += 2.0 * ch * (ch // 2) * k * rate
        rate *= u
        ch //= 2
        for rk, dil in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
            fl += 2.0 * ch * ch * rk * 2 * len(dil) * rate
    fl += 2.0 * ch * 7 * rate
    if not unit_only:
        e = h["embedding_dim"]
        fl += 2.0 * e * e * 4 / 2 + 2.0 * e * e  # unit ConvT (per output frame: 2 taps) + fc
    return fl
