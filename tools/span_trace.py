"""In-kernel start (min over CTAs) and end (max over CTAs) of every tcgen05 launch of one cfg2 forward:
shows the spans and the GAPS between consecutive kernels."""
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
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    assert lib.l2s_debug_set(k.encode(), int(v)) == 0
h = vo.shipped_config()
g = pkg.MelCodeGenerator(pkg.AttrDict(h))
g.load_state_dict(vo.init_state_dict(h, seed=1234, style="ref"), strict=True)
g.eval(); g.remove_weight_norm(); g = g.to(dev)
code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(16, 400, seed=52))
for _ in range(3):
    g(code=code, mel=mel, spkr=spkr)
torch.cuda.synchronize()
n = 128
span = torch.zeros(n, 2, dtype=torch.int64, device=dev)
span[:, 0] = 2 ** 62
lib.l2s_debug_set(b"span_ptr", span.data_ptr())
g(code=code, mel=mel, spkr=spkr)
torch.cuda.synchronize()
lib.l2s_debug_set(b"span_ptr", 0)
s = span.cpu()
rows = [(int(a), int(b)) for a, b in s.tolist() if b > 0]
t0 = rows[0][0]
tot_span = sum(b - a for a, b in rows)
gaps = [rows[i + 1][0] - rows[i][1] for i in range(len(rows) - 1)]
print(f"{len(rows)} tcgen05 launches; first start -> last end {(rows[-1][1] - t0) / 1e3:.1f} us; sum of spans {tot_span / 1e3:.1f} us; "
      f"sum of gaps {sum(gaps) / 1e3:.1f} us (avg {sum(gaps) / len(gaps) / 1e3:.2f}, max {max(gaps) / 1e3:.1f})")
for i, (a, b) in enumerate(rows):
    gap = (rows[i][0] - rows[i - 1][1]) / 1e3 if i else 0.0
    print(f"  launch {i:2d}: start {(a - t0) / 1e3:8.1f}  span {(b - a) / 1e3:6.1f}  gap before {gap:5.1f}")
