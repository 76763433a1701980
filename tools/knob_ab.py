"""Interleaved A/B timing of knob settings on the cfg2 forward (graph replay, 20 forwards per sample, best and median of
the samples; settings alternate so that the power-capped clock drift hits all of them alike), plus output equality.
BURST=1 in the environment: one second of idle before every sample (the 20 timed forwards then run at burst clocks like
bench.py's timed region; back to back the GPU sits at its power cap and any setting that fills idle time just lowers the clock).

    python tools/knob_ab.py "narrow_par=0" "narrow_par=1" "post_rows=0" ...      ("-" = defaults)
"""
import os
import statistics
import time
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
settings = [dict(kv.split("=") for kv in a.split()) if a != "-" else {} for a in sys.argv[1:]] or [{}]
all_keys = sorted({k for s in settings for k in s})
base = {k: None for k in all_keys}


def apply(s):
    for k, v in s.items():
        assert lib.l2s_debug_set(k.encode(), int(v)) == 0, k


def restore(s, defaults):
    for k in s:
        lib.l2s_debug_set(k.encode(), defaults[k])


DEFAULTS = dict(res_iss2=0, chain=0, fuse_branch=1, dual=1, cluster=1, narrow_par=0, post_rows=1, front_fuse=1, res_skew=0, pack=1, branch_par=1, use_graph=1, res_wide=1, res_cg2=4, pk_chan=32)
samples = [[] for _ in settings]
outs = [None] * len(settings)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rnd in range(6):
    for i, s in enumerate(settings):
        apply(s)
        if os.environ.get("BURST"):
            torch.cuda.synchronize()
            time.sleep(1.0)
        for _ in range(3):
            o = g(code=code, mel=mel, spkr=spkr)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(20):
            o = g(code=code, mel=mel, spkr=spkr)
        ev[1].record()
        torch.cuda.synchronize()
        samples[i].append(ev[0].elapsed_time(ev[1]) / 20 * 1e3)
        outs[i] = o.clone()
        restore(s, DEFAULTS)
for i, s in enumerate(settings):
    print(f"[ab] {sys.argv[1 + i] if len(sys.argv) > 1 else '-'}: best {min(samples[i]):.1f} us  median {statistics.median(samples[i]):.1f} us   equal to first: {torch.equal(outs[0], outs[i])}  "
          f"max diff {float((outs[0] - outs[i]).abs().max()):.2e}", flush=True)
