"""Timeline of CTA 0 of chosen skewed whole-ResBlock launches (resq_tc.cuh) of one cfg2 forward.

    python tools/resq_trace.py [launch ...] [k=v knobs]     launch = 3 * (stage - 2) + branch   (pack=0 res_mode=2 res_skew=1 set here)
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
code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(16, 400, seed=52))
for k, v in (("pack", 0), ("res_mode", 2), ("res_skew", 1), ("use_graph", 0)):
    lib.l2s_debug_set(k.encode(), v)
for kv in [a for a in sys.argv[1:] if "=" in a]:
    k, v = kv.split("=")
    assert lib.l2s_debug_set(k.encode(), int(v)) == 0, kv
for _ in range(2):
    g(code=code, mel=mel, spkr=spkr)
torch.cuda.synchronize()
n_dil = 3
for launch in [int(a) for a in sys.argv[1:] if "=" not in a] or [2, 1, 5, 8]:
    tr = torch.zeros(640, dtype=torch.int64, device=dev)
    lib.l2s_debug_set(b"trace_ptr", tr.data_ptr()); lib.l2s_debug_set(b"trace_launch", launch)
    g(code=code, mel=mel, spkr=spkr)
    torch.cuda.synchronize()
    lib.l2s_debug_set(b"trace_ptr", 0); lib.l2s_debug_set(b"trace_launch", -1)
    t = tr.cpu()
    e = [int(x) for x in t[:128].tolist() if int(x)]
    m = [int(x) for x in t[128:256].tolist() if int(x)]
    if not e or not m:
        print("launch", launch, "no stamps"); continue
    t0 = min(e[0], m[0])
    print(f"skewed whole-ResBlock launch {launch} (stage {2 + launch // 3}, branch {launch % 3}); us since first stamp")
    # E stamps per item: per step: [A start, A got gr0, A done, (B start, B got gr0, B done)], last step: loadS start, loadS done, out start, out done, loadX done
    names = []
    for s in range(n_dil):
        names += [f"s{s}:A[", "got0", "]"]
        if s + 1 < n_dil:
            names += [f"s{s}:B[", "got0", "]"]
    names += ["loadS[", "]", "out[", "]", "loadX]"]
    per = len(names)
    for it in range(min(4, len(e) // per)):
        seg = e[it * per:(it + 1) * per]
        print(f"  E item {it}: " + " ".join(f"{n}{(x - t0) / 1e3:.2f}" for n, x in zip(names, seg)))
    print(f"  MMA warp cycles: waiting for weights {int(t[500])}, waiting for e_done {int(t[501])}, whole loop {int(t[502])}")
    per_m = 2 * n_dil * 2
    for it in range(min(4, len(m) // per_m)):
        seg = m[it * per_m:(it + 1) * per_m]
        print(f"  M item {it} [conv start, all issued]: " + " ".join(f"c{j // 2}[{(seg[j] - t0) / 1e3:.2f} {(seg[j + 1] - t0) / 1e3:.2f}]" for j in range(0, len(seg), 2)))
