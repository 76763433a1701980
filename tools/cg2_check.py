"""cta_group::2 fused steps (knobs cluster=1 cg2=1) against the default kernels: waveform must be bit-identical; per-stage times."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import vocoder_oracle as vo
pkg = ge.load_package(); lib = pkg._cabi.load(); dev = torch.device("cuda:0")
h = vo.shipped_config()
g = pkg.MelCodeGenerator(pkg.AttrDict(h)); g.load_state_dict(vo.init_state_dict(h, seed=7, style="trained"), strict=True)
g.eval(); g.remove_weight_norm(); g = g.to(dev)
b, fr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3, 200)
code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(b, fr, seed=11))
lib.l2s_debug_set(b"use_graph", 0)
STAGE = int(os.environ.get("STAGE", "0"))
lib.l2s_debug_set(b"stop_after_stage", STAGE)
outs = []
for cl, cg2 in ((0, 0), (1, 0), (1, 1)):
    lib.l2s_debug_set(b"cluster", cl); lib.l2s_debug_set(b"cg2", cg2)
    try:
        g(code=code, mel=mel, spkr=spkr); torch.cuda.synchronize()
        outs.append(g.debug_tap("mrf", (b, fr * (5 if STAGE == 0 else 20), 256 if STAGE == 0 else 128), device=dev))
        d = (outs[0] - outs[-1]).abs()
        print(f"cluster={cl} cg2={cg2}: MRF tap max|d| {float(d.max()):.3e} nan {int(torch.isnan(outs[-1]).sum())} equal {bool(torch.equal(outs[0], outs[-1]))}")
    except Exception as e:
        print(f"cluster={cl} cg2={cg2}: FAILED {str(e)[:200]}")
