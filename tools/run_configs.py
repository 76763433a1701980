"""Measures the other BASELINE.json configs once on one GPU (device-resident inputs, CUDA events):
cfg3 per-GPU shard (32 x 8 s = the 8-GPU share of 256 x 8 s), cfg4 (120 s stream, chunked with a 24-frame
halo), cfg5 (64 x 6 s: multi-input vs unit-only at its 8-GPU share of 8 utterances).  Prints JSON."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402  (weights + synthetic inputs)

pkg = ge.load_package()
dev = torch.device("cuda:0")


def gen(cls, h, sd):
    g = getattr(pkg, cls)(pkg.AttrDict(h))
    g.load_state_dict(sd, strict=True)
    g.eval(); g.remove_weight_norm()
    return g.to(dev)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


h = vo.shipped_config()
g = gen("MelCodeGenerator", h, vo.init_state_dict(h, seed=1234, style="ref"))
fl = vo.algorithmic_flops_per_frame(h)
out = {}
code, mel, spk = (t.to(dev) for t in vo.synthetic_inputs(32, 800, seed=52))
ms = timed(lambda: g(code=code, mel=mel, spkr=spk))
out["cfg3_shard_32x8s"] = {"ms": ms, "audio_s_per_s": 32 * 8 / ms * 1e3, "tensor_frac": fl * 32 * 800 / ms / 1e9 / 1370.3}
code, mel, spk = (t.to(dev) for t in vo.synthetic_inputs(1, 12000, seed=52))
for core in (1000, 2000):
    ms = timed(lambda: pkg.vocode_long(g, code, mel, spk, core=core), iters=5)
    out[f"cfg4_120s_core{core}"] = {"ms": ms, "audio_s_per_s": 120 / ms * 1e3, "chunks": len(pkg.chunk_plan(12000, core))}
code, mel, spk = (t.to(dev) for t in vo.synthetic_inputs(8, 600, seed=52))
ms = timed(lambda: g(code=code, mel=mel, spkr=spk))
out["cfg5_multi_input_8x6s"] = {"ms": ms, "audio_s_per_s": 48 / ms * 1e3}
hu = vo.unit_only_config()
gu = gen("CodeGenerator", hu, vo.init_state_dict(hu, seed=1234, style="ref", unit_only=True))
ucode = torch.randint(0, 200, (8, 300), device=dev)
uspk = torch.randint(0, 200, (8, 1), device=dev)
ms = timed(lambda: gu(code=ucode, spkr=uspk))
out["cfg5_unit_only_8x6s"] = {"ms": ms, "audio_s_per_s": 48 / ms * 1e3}
print(json.dumps(out, indent=1))
