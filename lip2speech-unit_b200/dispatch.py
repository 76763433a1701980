"""Host-side work partitioning around the generator forward: utterance sharding
across GPUs (no collective: utterances are independent) and the
halo-exact chunk plan for long streams."""
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# An output sample depends on conditioning within +-21.6 mel frames;
# 24 is the next even count, so chunk boundaries stay on unit (2-frame) boundaries.
HALO_FRAMES = 24


def shard_utterances(lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Indices of the utterances rank `rank` vocodes.  Longest-first round-robin
    dealing keeps per-rank work within one utterance of balanced; equal lengths
    degenerate to a strided split.  Every index is owned by exactly one rank."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return sorted(order[rank::world_size])


def chunk_plan(frames: int, core: int, halo: int = HALO_FRAMES) -> List[Tuple[int, int, int, int]]:
    """Split `frames` conditioning frames into chunks of `core` frames plus `halo`
    frames of context per side.  Returns (lo, hi, keep_lo, keep_hi): vocode
    frames [lo,hi) and keep the samples of frames [keep_lo,keep_hi).  All bounds
    are even so unit boundaries (2 frames) are respected."""
    if frames < 1 or core < 2 or core % 2 or halo % 2 or halo < 0:
        raise ValueError("frames >= 1, core even >= 2, halo even >= 0 required")
    out = []
    k = 0
    while k < frames:
        k2 = min(k + core, frames)
        out.append((max(0, k - halo), min(frames, k2 + halo), k, k2))
        k = k2
    return out


@torch.no_grad()
def vocode_long(generator, code: torch.Tensor, mel: torch.Tensor, spkr: torch.Tensor, core: int = 1000,
                halo: int = HALO_FRAMES, hop: int = 160) -> torch.Tensor:
    """Vocode one long stream (B=1) in overlapping chunks; chunks of equal length are
    batched into one forward.  With halo >= 22 frames the result equals the
    unchunked forward (every layer's zero padding only matters within the halo)."""
    if code.shape[0] != 1:
        raise ValueError("vocode_long takes a single stream (B=1)")
    frames = mel.shape[2]
    plan = chunk_plan(frames, core, halo)
    out = torch.empty((1, 1, frames * hop), dtype=torch.float32, device=mel.device)
    by_len = {}
    for c in plan:
        by_len.setdefault(c[1] - c[0], []).append(c)
    for n, chunks in by_len.items():
        mel_b = torch.stack([mel[0, :, lo:hi] for lo, hi, _, _ in chunks])
        code_b = torch.stack([code[0, lo // 2:hi // 2] for lo, hi, _, _ in chunks])
        spk_b = spkr.expand(len(chunks), -1).contiguous()
        y = generator(code=code_b, mel=mel_b, spkr=spk_b)
        for i, (lo, hi, klo, khi) in enumerate(chunks):
            out[0, 0, klo * hop:khi * hop] = y[i, 0, (klo - lo) * hop:(khi - lo) * hop]
    return out


class HostPipeline:
    """Back-to-back batches from pinned host memory: while batch i runs, the inputs of batch i+1 are already on
    their way to the device and the waveform of batch i-1 on its way back (two copy streams, double-buffered device
    inputs).  A caller that vocodes many batches (a manifest, a service queue) then pays the PCIe time only once.

        pipe = HostPipeline(generator, "cuda:0")
        for (code_h, mel_h, spk_h), out_h in zip(batches, pinned_outputs):
            pipe.submit(code_h, mel_h, spk_h, out_h)        # asynchronous; out_h: pinned float32 (B,1,L) or int16 (B,L)
        pipe.finish()                                       # every out_h is complete

    An int16 `out_h` gets the device-side int16 waveform (what every reference caller derives on the host,
    inference.py:79-81): half the device->host bytes.  The forward itself runs on the stream that is current when
    submit() is called."""

    def __init__(self, generator, device="cuda"):
        self.g = generator
        self.device = torch.device(device)
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.slot = 0
        self.in_free = [None, None]      # event: the forward that read slot s's device inputs has finished
        self.out_free = [None, None]     # event: the device->host copy of slot s's waveform has finished
        self.bufs = [None, None]         # per slot: (shapes, code, mel, spk, out) device buffers, allocated once per shape
        self.mel_cm = [None, None]       # per slot: channel-major mel when the caller hands time-major frames

    def _allocate(self, s, shapes, code_h, mel_h, spk_h, out_h, main, time_major=False):
        # a new shape: retire the old buffers first.  Only this pipeline's own work is waited for -- a device-wide
        # synchronize here could collide with another host thread that is capturing a CUDA graph on the same device.
        for ev in (self.in_free[s], self.out_free[s]):
            if ev is not None:
                ev.synchronize()
        with torch.cuda.device(self.device):
            bufs = (torch.empty_like(code_h, device=self.device), torch.empty_like(mel_h, device=self.device),
                    torch.empty_like(spk_h, device=self.device), torch.empty(out_h.shape, dtype=out_h.dtype, device=self.device))
            # time-major mel (n, T, 80), the layout of the .npy hand-off files: transposed on the device, not by the host
            self.mel_cm[s] = torch.empty((mel_h.shape[0], mel_h.shape[2], mel_h.shape[1]), dtype=torch.float32,
                                         device=self.device) if time_major else None
        # The blocks come from the allocating (main) stream's pool but are written on s_in and read on s_out: tell the
        # caching allocator, and make both copy streams wait until main has reached the allocation point (a recycled
        # block may still be in use by kernels queued on main).
        born = torch.cuda.Event()
        born.record(main)
        self.s_in.wait_event(born)
        self.s_out.wait_event(born)
        for t in bufs[:3]:
            t.record_stream(self.s_in)
        bufs[3].record_stream(self.s_out)
        self.bufs[s] = (shapes,) + bufs
        self.in_free[s] = self.out_free[s] = None

    @torch.no_grad()
    def submit(self, code_h, mel_h, spk_h, out_h, mel_time_major: bool = False):
        """mel_time_major: mel_h is (B, T, num_mels) as the hand-off files store it; the (B, num_mels, T) the generator takes
        is made on the device (a 4 MB transpose kernel instead of a strided host copy per utterance)."""
        s = self.slot
        self.slot ^= 1
        main = torch.cuda.current_stream(self.device)
        if out_h.dtype not in (torch.float32, torch.int16):
            raise TypeError("out_h must be float32 (B,1,L) or int16 (B,L)")
        shapes = (tuple(code_h.shape), tuple(mel_h.shape), tuple(spk_h.shape), mel_h.dtype, tuple(out_h.shape), out_h.dtype, mel_time_major)
        if self.bufs[s] is None or self.bufs[s][0] != shapes:
            self._allocate(s, shapes, code_h, mel_h, spk_h, out_h, main, mel_time_major)
        _, code, mel, spk, out = self.bufs[s]
        with torch.cuda.stream(self.s_in):
            if self.in_free[s] is not None:
                self.s_in.wait_event(self.in_free[s])          # the previous user of this slot is done with the inputs
            code.copy_(code_h, non_blocking=True)
            mel.copy_(mel_h, non_blocking=True)
            spk.copy_(spk_h, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.s_in)
        main.wait_event(ready)
        if self.out_free[s] is not None:
            main.wait_event(self.out_free[s])                  # the slot's previous waveform has left the device
        if mel_time_major:
            self.mel_cm[s].copy_(mel.transpose(1, 2))
            mel = self.mel_cm[s]
        if out.dtype == torch.int16:
            self.g.forward_int16_into(out, code=code, mel=mel, spkr=spk)
        else:
            self.g.forward_into(out, code=code, mel=mel, spkr=spk)
        done = torch.cuda.Event()
        done.record(main)
        self.in_free[s] = done
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            out_h.copy_(out, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.s_out)
        self.out_free[s] = copied
        return copied                                          # completes when out_h holds this batch's waveform

    def finish(self):
        """Wait until every submitted waveform is in its pinned host buffer."""
        self.s_out.synchronize()
        self.s_in.synchronize()

    def __del__(self):
        try:
            self.finish()           # the device buffers must not return to the pool while a copy is in flight
        except Exception:
            pass


def _pinned_stack(arrays, dtype):
    """One pinned (n, ...) tensor from n equal-shaped host arrays: a single torch.stack into pinned memory (the copy runs
    without the GIL, so the per-GPU host threads do not serialise on it)."""
    ts = [a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)) for a in arrays]
    t = torch.empty((len(ts),) + tuple(ts[0].shape), dtype=dtype, pin_memory=True)
    if ts[0].dtype == dtype:
        torch.stack(ts, out=t)
    else:
        for i, a in enumerate(ts):
            t[i].copy_(a)
    return t


class MultiGpuVocoder:
    """One PROCESS, every GPU of the box ("one host thread or process per GPU", no collective):
    the utterance list is dealt longest-first across the devices (shard_utterances), each device has its own engine
    (weights replicated, 28 MB), its own stream, its own HostPipeline and a host thread that feeds it; utterances of
    equal length share a forward, int16 waveforms come back through pinned memory (half the PCIe bytes of fp32).
    Results are bit-identical to vocoding every utterance alone on one GPU (batch invariance of the kernels).

        mg = MultiGpuVocoder(generator)                       # all visible GPUs
        wavs = mg.vocode(feats)                               # feats[i] = {"code": (U,), "mel": (80, 2U), "spkr": (256,)} host arrays
                                                              # -> list of int16 numpy arrays (160 * 2U samples), input order
    A service process keeps one instance; vocode() may be called from any thread (calls are serialised per instance)."""

    def __init__(self, generator, devices: Optional[Sequence] = None, max_batch: int = 32):
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError("MultiGpuVocoder needs at least one CUDA device (there is no CPU fallback)")
        self.g = generator
        self.devices = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices]
        self.max_batch = max_batch
        self._streams = [torch.cuda.Stream(d) for d in self.devices]
        self._pipes = [None] * len(self.devices)
        self._pool = ThreadPoolExecutor(max_workers=len(self.devices), thread_name_prefix="l2s-gpu")
        self._call = threading.Lock()

    def _worker(self, k: int, feats, mine: List[int], out: list):
        dev = self.devices[k]
        torch.cuda.set_device(dev)
        by_len: Dict[int, List[int]] = {}
        for i in mine:
            by_len.setdefault(int(feats[i]["mel"].shape[1]), []).append(i)
        batches = [idxs[s:s + self.max_batch] for _, idxs in sorted(by_len.items(), reverse=True)
                   for s in range(0, len(idxs), self.max_batch)]
        with torch.no_grad(), torch.cuda.stream(self._streams[k]):
            if self._pipes[k] is None:
                self._pipes[k] = HostPipeline(self.g, dev)
            pipe, pending = self._pipes[k], []
            for grp in batches:
                code = _pinned_stack([feats[i]["code"] for i in grp], torch.int64)
                mel = _pinned_stack([feats[i]["mel"] for i in grp], torch.float32)
                spk = _pinned_stack([feats[i]["spkr"] for i in grp], torch.float32)
                wav = torch.empty((len(grp), mel.shape[2] * 160), dtype=torch.int16, pin_memory=True)
                pending.append((grp, wav, pipe.submit(code, mel, spk, wav), (code, mel, spk)))
            for grp, wav, done, _keep in pending:
                done.synchronize()
                w = wav.numpy()                      # views of the pinned result buffer (kept alive by the views)
                for j, i in enumerate(grp):
                    out[i] = w[j]
            pipe.finish()
            self.g.check_index_errors(dev)          # synchronises this worker's stream only

    def vocode(self, feats: Sequence[dict]) -> List[np.ndarray]:
        with self._call:
            lengths = [int(f["mel"].shape[1]) for f in feats]
            out: list = [None] * len(feats)
            n = len(self.devices)
            order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))     # shard_utterances for every rank at once
            futs = [self._pool.submit(self._worker, k, feats, sorted(order[k::n]), out) for k in range(n)]
            for f in futs:
                f.result()
            return out

    def _batch_worker(self, k: int, code, mel, spk, wav, lo: int, hi: int):
        dev = self.devices[k]
        torch.cuda.set_device(dev)
        with torch.no_grad(), torch.cuda.stream(self._streams[k]):
            if self._pipes[k] is None:
                self._pipes[k] = HostPipeline(self.g, dev)
            pipe = self._pipes[k]
            for s in range(lo, hi, self.max_batch):
                e = min(hi, s + self.max_batch)
                pipe.submit(code[s:e], mel[s:e], spk[s:e], wav[s:e])
            pipe.finish()
            self.g.check_index_errors(dev)

    def vocode_batch(self, code: torch.Tensor, mel: torch.Tensor, spkr: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Equal-length utterances already stacked on the host: code (B,U) int64, mel (B,80,T) float32, spkr (B,256)
        float32 (pinned memory makes the copies asynchronous).  Utterances [k B / n, (k + 1) B / n) go to device k as
        batches of <= max_batch, straight from slices of the caller's tensors: no host staging at all.  Returns the
        int16 waveforms (B, 160 T) in a pinned tensor (or `out`)."""
        with self._call:
            b = code.shape[0]
            n = len(self.devices)
            if out is None:
                out = torch.empty((b, mel.shape[2] * 160), dtype=torch.int16, pin_memory=True)
            bounds = [(b * k) // n for k in range(n + 1)]
            futs = [self._pool.submit(self._batch_worker, k, code, mel, spkr, out, bounds[k], bounds[k + 1])
                    for k in range(n) if bounds[k + 1] > bounds[k]]
            for f in futs:
                f.result()
            return out

    def close(self):
        self._pool.shutdown(wait=True)
        self._pipes = [None] * len(self.devices)
