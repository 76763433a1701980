"""Whole-ResBlock kernels (fuse_branch=1) against the step-by-step kernels (fuse_branch=0), stage by stage:
SNR of the MRF output tap after every stage and of the waveform, then the waveform against the CPU oracle.

    python tools/branch_check.py [batch frames] [k=v knobs ...]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402

args = [a for a in sys.argv[1:] if "=" not in a]
knobs = [a for a in sys.argv[1:] if "=" in a]
batch = int(args[0]) if len(args) > 0 else 2
frames = int(args[1]) if len(args) > 1 else 64
pkg = ge.load_package()
lib = pkg._cabi.load()
dev = torch.device("cuda:0")
h = vo.shipped_config()
sd = vo.init_state_dict(h, seed=7, style="trained")
g = pkg.MelCodeGenerator(pkg.AttrDict(h))
g.load_state_dict(sd, strict=True)
g.eval()
g.remove_weight_norm()
g = g.to(dev)
for kv in knobs:
    k, v = kv.split("=")
    assert lib.l2s_debug_set(k.encode(), int(v)) == 0, kv
code, mel, spkr = vo.synthetic_inputs(batch, frames, seed=11)
cd, md, sp = code.to(dev), mel.to(dev), spkr.to(dev)


def snr(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float(10 * torch.log10((a * a).sum() / ((a - b) ** 2).sum().clamp_min(1e-300)))


chans = [256, 128, 64, 32, 16]
lens = [frames * 5, frames * 20, frames * 40, frames * 80, frames * 160]
ok = True
for stage in range(5):
    taps = []
    for fb in (0, 1):
        lib.l2s_debug_set(b"fuse_branch", fb)
        lib.l2s_debug_set(b"stop_after_stage", stage)
        g(code=cd, mel=md, spkr=sp)
        torch.cuda.synchronize()
        taps.append(g.debug_tap("mrf", (batch, lens[stage], chans[stage]), device=dev))
    s = snr(taps[0], taps[1])
    d = (taps[0] - taps[1]).abs()
    bad = torch.nonzero(d > 0.05 * taps[0].abs().max())
    print(f"stage {stage} C={chans[stage]:3d}: whole vs steps {s:6.1f} dB  max|d| {float(d.max()):.3e} (max|x| {float(taps[0].abs().max()):.3e})"
          f"  nan {int(torch.isnan(taps[1]).sum())}  big diffs {len(bad)}" + (f" first at {bad[0].tolist()} last at {bad[-1].tolist()}" if len(bad) else ""))
    ok = ok and s > 40
lib.l2s_debug_set(b"stop_after_stage", -1)
outs = []
for fb in (0, 1):
    lib.l2s_debug_set(b"fuse_branch", fb)
    for _ in range(2):   # second call replays the graph
        y = g(code=cd, mel=md, spkr=sp)
    torch.cuda.synchronize()
    outs.append(y.float().cpu())
ref = vo.mel_code_generator_forward(vo.fold_weight_norm(sd), h, code, mel, spkr)
ref = torch.as_tensor(np.asarray(ref)).float()
print(f"waveform: whole vs steps {snr(outs[0], outs[1]):.1f} dB; steps vs oracle {snr(ref, outs[0]):.1f} dB; whole vs oracle {snr(ref, outs[1]):.1f} dB")
print("launches", g.launch_count(batch, frames, device=dev))
print("OK" if ok else "MISMATCH")
