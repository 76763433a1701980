"""Runs N forwards of the cfg2 batch (16 x 4 s, bf16 mode) and nothing else: the
command the ncu launch list / full captures are taken from."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402  (weights + synthetic inputs only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pkg = ge.load_package()
for kv in sys.argv[2:]:                      # debug knobs k=v (tools only)
    k, v = kv.split("=")
    assert pkg._cabi.load().l2s_debug_set(k.encode(), int(v)) == 0, kv
dev = torch.device("cuda:0")
h = vo.shipped_config()
g = pkg.MelCodeGenerator(pkg.AttrDict(h))
g.load_state_dict(vo.init_state_dict(h, seed=1234, style="ref"), strict=True)
g.eval()
g.remove_weight_norm()
g = g.to(dev)
code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(16, 400, seed=52))
for _ in range(n):
    y = g(code=code, mel=mel, spkr=spkr)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.abs().max()))
