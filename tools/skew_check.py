"""Skewed whole-ResBlock schedule (resq_tc.cuh) against res_tc_kernel: bit-equality on small and cfg2 shapes, and the
in-kernel spans of the nine narrow-stage ResBlock launches of one cfg2 forward for a list of knob settings.

    python tools/skew_check.py            (GPU)
"""
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


def knobs(**kw):
    for k, v in kw.items():
        assert lib.l2s_debug_set(k.encode(), int(v)) == 0, k


DEFAULT = dict(pack=0, res_mode=0, res_skew=0, res_cg2=4, res_wide=1, res_msub=8, use_graph=0, res_ng=2)
only = sys.argv[1:]
for shape in ((3, 150), (1, 34), (16, 400)):
    code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(*shape, seed=33))
    knobs(**DEFAULT)
    knobs(res_mode=2)
    ref = g(code=code, mel=mel, spkr=spkr).clone()
    torch.cuda.synchronize()
    for kw in (dict(res_skew=1, res_cg2=0, res_wide=0), dict(res_skew=1, res_cg2=0, res_wide=1), dict(res_skew=1, res_cg2=4, res_wide=0),
               dict(res_skew=1, res_cg2=4, res_wide=1), dict(res_skew=1, res_msub=4), dict(res_skew=1, res_msub=2), dict(res_skew=1, res_mode=0), dict(res_skew=1, res_ng=4), dict(res_skew=1, res_ng=8),
               dict(res_skew=1, res_ng=1)):
        knobs(**DEFAULT)
        knobs(res_mode=2)
        knobs(**kw)
        out = g(code=code, mel=mel, spkr=spkr).clone()
        torch.cuda.synchronize()
        print(f"[skew] shape {shape} {kw}: equal={torch.equal(ref, out)} max diff {float((ref - out).abs().max()):.3e}", flush=True)

code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(16, 400, seed=52))
names = [f"C={c} k={k}" for c in (64, 32, 16) for k in (3, 7, 11)]
for kw in (dict(), dict(res_mode=2), dict(res_mode=2, res_skew=1), dict(res_mode=2, res_skew=1, res_wide=0), dict(res_mode=2, res_skew=1, res_cg2=0),
           dict(res_mode=2, res_skew=1, res_msub=4), dict(res_mode=2, res_skew=1, res_ng=4), dict(res_mode=2, res_skew=1, res_ng=1), dict(res_mode=2, res_skew=1, res_ng=4, res_wide=0)):
    knobs(**DEFAULT)
    knobs(**kw)
    for _ in range(3):
        g(code=code, mel=mel, spkr=spkr)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        span = torch.zeros(128, 2, dtype=torch.int64, device=dev)
        span[:, 0] = 2 ** 62
        knobs(span_ptr=span.data_ptr())
        g(code=code, mel=mel, spkr=spkr)
        torch.cuda.synchronize()
        knobs(span_ptr=0)
        rows = [(int(a), int(b)) for a, b in span.cpu().tolist() if b > 0]
        us = [(b - a) / 1e3 for a, b in rows]
        best = us if best is None else [min(x, y) for x, y in zip(best, us)]
    # launches: conv_pre, ups0, 9 steps, ups1, 9 steps, ups2, 3 res, ups3, 3 res, ups4, 3 res
    res = best[22:25] + best[26:29] + best[30:33]
    print(f"[skew-span] {kw}: " + "  ".join(f"{n} {u:6.1f}" for n, u in zip(names, res)) + f"   sum {sum(res):.1f} us ({len(best)} launches)", flush=True)
knobs(**DEFAULT)
knobs(pack=1, res_skew=1, use_graph=1)
