"""Cycle accounting of one epilogue warp over whole cfg2 forwards stopped after a given stage."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402

pkg = ge.load_package()
lib = pkg._cabi.load()
dev = torch.device("cuda:0")
h = vo.shipped_config()
g = pkg.MelCodeGenerator(pkg.AttrDict(h))
g.load_state_dict(vo.init_state_dict(h, seed=1234, style="ref"), strict=True)
g.eval(); g.remove_weight_norm(); g = g.to(dev)
code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(16, 400, seed=52))
g(code=code, mel=mel, spkr=spkr)
torch.cuda.synchronize()
buf = (C.c_longlong * 8)()
prev = [0] * 8
for stage in range(5):
    lib.l2s_debug_set(b"stop_after_stage", stage)
    lib.l2s_debug_set(b"epi_prof", 1)
    lib.l2s_debug_epi_prof(buf)          # reset
    g(code=code, mel=mel, spkr=spkr)
    lib.l2s_debug_epi_prof(buf)
    cur = list(buf)
    d = [c - p for c, p in zip(cur, prev)]
    prev = cur
    n = max(d[3], 1)
    print(f"stages 0..{stage} cumulative -> this stage: chunks {d[3]}, per chunk cycles: tmem_ld+wait {d[0]/n:.0f}, "
          f"transpose {d[1]/n:.0f}, finish {d[2]/n:.0f}; per-item locate+res-issue {d[4]:.0f} total, barrier wait {d[5]:.0f} total")
lib.l2s_debug_set(b"stop_after_stage", -1)
lib.l2s_debug_set(b"epi_prof", 0)
