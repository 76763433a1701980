"""Experiment: one cfg2 batch (16 x 4 s) as ONE forward vs as TWO / FOUR concurrent forwards of 8 / 4 utterances on their own
streams (own handles and workspaces): do the kernels of one group fill the wave-quantisation tails and kernel boundaries of the other?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402

pkg = ge.load_package()
dev = torch.device("cuda:0")
h = vo.shipped_config()
sd = vo.init_state_dict(h, seed=1234, style="ref")


def make():
    g = pkg.MelCodeGenerator(pkg.AttrDict(h))
    g.load_state_dict(sd, strict=True)
    g.eval(); g.remove_weight_norm()
    return g.to(dev)


code, mel, spkr = (t.to(dev) for t in vo.synthetic_inputs(16, 400, seed=52))
gens = [make() for _ in range(4)]
streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
ref = gens[0](code=code, mel=mel, spkr=spkr).clone()


def run(parts):
    n = 16 // parts
    outs = [None] * parts
    main = torch.cuda.current_stream()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)

    def once():
        fork = torch.cuda.Event(); fork.record(main)
        for p in range(parts):
            s = streams[p] if parts > 1 else main
            if parts > 1:
                s.wait_event(fork)
            with torch.cuda.stream(s):
                outs[p] = gens[p](code=code[p * n:(p + 1) * n], mel=mel[p * n:(p + 1) * n], spkr=spkr[p * n:(p + 1) * n])
                if parts > 1:
                    j = torch.cuda.Event(); j.record(s); main.wait_event(j)
    for _ in range(4):
        once()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(6):
        ev0.record(main)
        for _ in range(10):
            once()
        ev1.record(main)
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) / 10)
    out = torch.cat(outs, 0)
    return best, torch.equal(out, ref)


for parts in (1, 2, 4, 1, 2):
    ms, same = run(parts)
    print(f"[two-stream] {parts} concurrent forward(s) of {16 // parts} utterances: {ms * 1e3:.1f} us per 16 utterances -> {64 / ms * 1e3:.0f} audio-s/s, bit-equal to the single forward: {same}", flush=True)
