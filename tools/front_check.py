"""front_fuse (speaker projection inside the conditioning kernel) vs the separate launch: bit-equality and timing of the
front end and the head of one cfg2 forward (cudaEvent pairs around each launch)."""
import os
import sys
import ctypes as C

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
outs = {}
for shape in ((3, 150), (16, 400), (1, 8)):
    code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(*shape, seed=33))
    for ff in (0, 1):
        lib.l2s_debug_set(b"front_fuse", ff)
        outs[ff] = g(code=code, mel=mel, spkr=spkr).clone()
    torch.cuda.synchronize()
    print(f"[front] shape {shape}: fused == separate: {torch.equal(outs[0], outs[1])}")
code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(16, 400, seed=52))
for ff in (0, 1):
    lib.l2s_debug_set(b"front_fuse", ff)
    for _ in range(3):
        g(code=code, mel=mel, spkr=spkr)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(5):
        ev[0].record()
        for _ in range(20):
            g(code=code, mel=mel, spkr=spkr)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]) / 20)
    print(f"[front] front_fuse={ff}: forward {best * 1e3:.1f} us")
lib.l2s_debug_set(b"front_fuse", 1)
