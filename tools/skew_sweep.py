"""In-kernel spans of the narrow-stage ResBlock launches of one cfg2 forward for knob settings given on the command line:
    python tools/skew_sweep.py "res_mode=2 res_skew=1 res_tb=1" "res_mode=2 res_skew=1 res_gmax=1" ..."""
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


DEFAULT = dict(pack=0, res_mode=0, res_skew=0, res_cg2=4, res_wide=1, res_msub=8, use_graph=0, res_ng=2, res_tb=0, res_gmax=0, res_skew_pct=100, res_skew_iss2=1, res_iss2=0, max_nt=256, max_msub=8, tc_cg2=1)
code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(16, 400, seed=52))
names = [f"C={c} k={k}" for c in (64, 32, 16) for k in (3, 7, 11)]
ref = None
for arg in sys.argv[1:]:
    kw = dict(kv.split("=") for kv in arg.split())
    knobs(**DEFAULT)
    knobs(**kw)
    for _ in range(3):
        out = g(code=code, mel=mel, spkr=spkr)
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    same = torch.equal(ref, out)
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
    if os.environ.get("SWEEP_CONV"):        # conv_pre and the five upsamplers instead
        cv = [best[0], best[1], best[11], best[21], best[25], best[29]]
        print(f"[sweep-conv] {arg}: " + "  ".join(f"{n} {u:6.1f}" for n, u in zip(["conv_pre", "ups0", "ups1", "ups2", "ups3", "ups4"], cv)) + f"   sum {sum(cv):.1f} us  equal_to_first={same}", flush=True)
        continue
    res = best[22:25] + best[26:29] + best[30:33]
    print(f"[sweep] {arg}: " + "  ".join(f"{n} {u:6.1f}" for n, u in zip(names, res)) + f"   sum {sum(res):.1f} us  equal_to_first={same}", flush=True)
knobs(**DEFAULT)
knobs(pack=1, use_graph=1)
