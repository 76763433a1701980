"""Chained ResBlock steps (knob chain, pair_tc.cuh) against plain stream order: max difference of the waveform over four forwards
(eager, graph capture, replays) for several launch modes.  Found the graph-capture pitfall: EVERY incoming edge of a programmatically
launched node becomes programmatic, also the one from another stream's event."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import vocoder_oracle as vo
pkg = ge.load_package(); lib = pkg._cabi.load(); dev = torch.device("cuda:0")
h = vo.shipped_config()
g = pkg.MelCodeGenerator(pkg.AttrDict(h)); g.load_state_dict(vo.init_state_dict(h, seed=1234, style="trained"), strict=True)
g.eval(); g.remove_weight_norm(); g = g.to(dev)
def K(**kw):
    for k, v in kw.items(): assert lib.l2s_debug_set(k.encode(), int(v)) == 0, k
for shape in ((3, 150), (16, 400)):
    code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(*shape, seed=33))
    for base in (dict(use_graph=1, branch_par=1, fuse_branch=0), dict(use_graph=0, branch_par=1, fuse_branch=0), dict(use_graph=1, branch_par=0, fuse_branch=0),
                 dict(use_graph=1, branch_par=1, fuse_branch=0, stop_after_stage=1), dict(use_graph=0, branch_par=1, fuse_branch=0, stop_after_stage=2), dict(use_graph=0, branch_par=1, fuse_branch=0, stop_after_stage=3)):
        K(chain=0, **base)
        for _ in range(3): a = g(code=code, mel=mel, spkr=spkr).clone()
        K(chain=1)
        res = []
        for _ in range(4):
            b = g(code=code, mel=mel, spkr=spkr).clone()
            torch.cuda.synchronize()
            res.append(float((a - b).abs().max()))
        print(shape, base, "max diffs over 4 forwards:", ["%.2e" % r for r in res], flush=True)
        K(use_graph=1, branch_par=1, fuse_branch=1, cluster=1, dual=1, chain=0, stop_after_stage=-1)
