"""Timeline of CTA 0 of chosen whole-ResBlock launches of one cfg2 forward (epilogue warp 2 and the MMA warp).

    python tools/res_trace.py [launch ...] [k=v knobs]     launch = 3 * (stage - 2) + branch
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
for kv in [a for a in sys.argv[1:] if "=" in a]:
    k, v = kv.split("=")
    assert lib.l2s_debug_set(k.encode(), int(v)) == 0, kv
lib.l2s_debug_set(b"use_graph", 0)
for _ in range(2):
    g(code=code, mel=mel, spkr=spkr)
torch.cuda.synchronize()
n_dil = 3
for launch in [int(a) for a in sys.argv[1:] if "=" not in a] or [0, 2, 6, 8]:
    tr = torch.zeros(640, dtype=torch.int64, device=dev)
    lib.l2s_debug_set(b"trace_ptr", tr.data_ptr()); lib.l2s_debug_set(b"trace_launch", launch)
    g(code=code, mel=mel, spkr=spkr)
    torch.cuda.synchronize()
    lib.l2s_debug_set(b"trace_ptr", 0); lib.l2s_debug_set(b"trace_launch", -1)
    t = tr.cpu()
    e = [int(x) for x in t[:128].tolist() if int(x)]
    m = [int(x) for x in t[128:256].tolist() if int(x)]
    fine = t[256:320].view(4, 16)
    stg = [int(x) for x in t[320:448].tolist() if int(x)]
    for ph in range(6):
        row = [int(x) for x in t[448 + 16 * ph:448 + 16 * ph + 16].tolist() if int(x)]
        if row:
            print(f"  item 1 phase {'AB'[ph & 1]}{ph // 2}, per unit [start, tmem ld done, bias added, stored] us: " + " ".join(f"{(x - row[0]) / 1e3:.2f}" for x in row))
    if not e:
        print("launch", launch, "no stamps"); continue
    t0 = min(e[0], m[0])
    print(f"  MMA warp cycles of CTA 0: waiting for weights {int(t[500])}, waiting for the epilogue warps {int(t[501])}, whole loop {int(t[502])}")
    print(f"whole-ResBlock launch {launch} (stage {2 + launch // 3}, branch {launch % 3}); us since first stamp")
    # epilogue stamps per item: start, x loaded, then per step: [wait A.., got d1, phase A done, (got x, phase B done)], final done
    per_item_e = 2 + n_dil * 3 + (n_dil - 1) * 2 + 1
    names = ["start", "x_loaded"]
    for s in range(n_dil):
        names += [f"s{s}:waitA", f"s{s}:gotD1", f"s{s}:A_done"]
        if s + 1 < n_dil:
            names += [f"s{s}:gotX", f"s{s}:B_done"]
    names += ["out_done"]
    for it in range(len(e) // per_item_e):
        seg = e[it * per_item_e:(it + 1) * per_item_e]
        print(f"  E item {it}: " + " ".join(f"{n}={(x - t0) / 1e3:.2f}" for n, x in zip(names, seg)))
    per_item_m = 2 * n_dil * 3
    for it in range(len(m) // per_item_m):
        seg = m[it * per_item_m:(it + 1) * per_item_m]
        print(f"  M item {it}: " + " ".join(f"c{j // 3}[{(seg[j] - t0) / 1e3:.2f} {(seg[j + 1] - t0) / 1e3:.2f} {(seg[j + 2] - t0) / 1e3:.2f}]" for j in range(0, len(seg), 3)))
    if stg:
        print("  M weight stages [wait start, weights landed, MMAs issued] us: " + " ".join(f"[{(stg[j] - t0) / 1e3:.2f} {(stg[j + 1] - t0) / 1e3:.2f} {(stg[j + 2] - t0) / 1e3:.2f}]" for j in range(0, len(stg) - 2, 3)))
    for it in range(3):
        row = [int(x) for x in fine[it].tolist() if int(x)]
        if row:
            print(f"  phase A (step 0) of item {it}, per unit [ld issued, ld done, stored]: " + " ".join(f"{(x - t0) / 1e3:.2f}" for x in row))
