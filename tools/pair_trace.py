"""Dumps the per-role timestamps of CTA 0 for chosen fused ResBlock-step launches of one cfg2 forward."""
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
for _ in range(2):
    g(code=code, mel=mel, spkr=spkr)
torch.cuda.synchronize()
for launch in [int(a) for a in sys.argv[1:]] or [18, 0, 9, 27, 36]:
    tr = torch.zeros(3 * 64 * 4 + 512 * 3 + 32 * 8, dtype=torch.int64, device=dev)
    lib.l2s_debug_set(b"trace_ptr", tr.data_ptr()); lib.l2s_debug_set(b"trace_launch", launch)
    g(code=code, mel=mel, spkr=spkr)
    torch.cuda.synchronize()
    lib.l2s_debug_set(b"trace_ptr", 0); lib.l2s_debug_set(b"trace_launch", -1)
    cta = tr.cpu()[768:768 + 1536].view(512, 3)
    live = cta[cta[:, 1] > 0]
    if live.numel():
        t_first = int(live[:, 1].min())
        starts = sorted(int(x) - t_first for x in live[:, 1].tolist())
        ends = sorted(int(x) - t_first for x in live[:, 2].tolist())
        n = len(starts)
        print(f"  CTAs {n} on {len(set(live[:,0].tolist()))} SMs; start times (us) min/med/max {starts[0]/1e3:.1f}/{starts[n//2]/1e3:.1f}/{starts[-1]/1e3:.1f}; "
              f"end times min/10%/med/90%/max {ends[0]/1e3:.1f}/{ends[n//10]/1e3:.1f}/{ends[n//2]/1e3:.1f}/{ends[9*n//10]/1e3:.1f}/{ends[-1]/1e3:.1f}")
    t = tr.cpu()[:768].view(3, 64, 4)
    nz = t[t > 0]
    if nz.numel() == 0:
        print("launch", launch, "no stamps"); continue
    t0 = int(nz.min())
    f = lambda v: int(v) - t0 if int(v) else -1
    print(f"fused-step launch {launch} (stage {launch // 9}, branch {(launch % 9) // 3}, step {launch % 3}): ns since first stamp")
    print("  item: P[wait_empty,issue] M[wait_A,got_A,c1_committed,c2_start] E[got_d1,t_full,got_d2,done]")
    for i in range(64):
        if int(t[2, i, 3]) == 0:
            break
        fine = tr.cpu()[2304 + 8 * i: 2304 + 8 * i + 8] if i < 32 else None
        fs = ("  P2[enter,d2,ld0,res0,st0,ld1,res1,st1]=" + str([f(x) for x in fine.tolist()])) if fine is not None and int(fine[0]) else ""
        print(f"  {i:2d}: P[{f(t[0,i,0])},{f(t[0,i,1])}] M[{f(t[1,i,0])},{f(t[1,i,1])},{f(t[1,i,2])},{f(t[1,i,3])}] "
              f"E[{f(t[2,i,0])},{f(t[2,i,1])},{f(t[2,i,2])},{f(t[2,i,3])}]" + fs)
