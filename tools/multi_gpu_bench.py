"""cfg3 (BASELINE.json configs[2]: 256 x 8 s utterances sharded by utterance) through the IN-PROCESS multi-GPU
dispatcher (dispatch.MultiGpuVocoder: one process, one host thread + engine + HostPipeline per GPU, int16 back),
for 1, 2, 4, ... up to all visible GPUs.  Wall-clock of the whole call (host staging, H2D, forwards, D2H).  Prints JSON.

    python tools/multi_gpu_bench.py [n_utts] [frames]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402  (weights + synthetic inputs)

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 256
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 800
pkg = ge.load_package()
h = vo.shipped_config()
g = pkg.MelCodeGenerator(pkg.AttrDict(h))
g.load_state_dict(vo.init_state_dict(h, seed=1234, style="ref"), strict=True)
g.eval(); g.remove_weight_norm()
code, mel, spkr = vo.synthetic_inputs(n_utts, frames, seed=52)
feats = [{"code": code[i].numpy(), "mel": mel[i].numpy(), "spkr": spkr[i].numpy()} for i in range(n_utts)]
audio_s = n_utts * frames / 100.0
res = {"workload": f"{n_utts} x {frames / 100:.0f} s utterances, in-process MultiGpuVocoder, int16 back", "audio_s": audio_s, "runs": {}}
n_all = torch.cuda.device_count()
base, base_b, first, pinned, wav = None, None, None, None, None
n = 1
while n <= n_all:
    mg = pkg.MultiGpuVocoder(g, devices=list(range(n)), max_batch=32)
    out = mg.vocode(feats)                       # warm-up: engines, plans, graphs, pinned pools
    out = mg.vocode(feats)
    best = 1e9
    for _ in range(3):
        for d in range(n):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        out = mg.vocode(feats)
        best = min(best, time.perf_counter() - t0)
    mg.close()
    # the same job with the utterances already stacked in pinned host tensors (no host staging)
    if pinned is None:
        pinned = (code.pin_memory(), mel.pin_memory(), spkr.pin_memory())
        wav = torch.empty((n_utts, frames * 160), dtype=torch.int16, pin_memory=True)
    mg2 = pkg.MultiGpuVocoder(g, devices=list(range(n)), max_batch=32)
    mg2.vocode_batch(*pinned, out=wav); mg2.vocode_batch(*pinned, out=wav)
    best_b = 1e9
    for _ in range(5):
        for d in range(n):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        mg2.vocode_batch(*pinned, out=wav)
        best_b = min(best_b, time.perf_counter() - t0)
    mg2.close()
    if first is None:
        first = [o.copy() for o in out]
    else:
        assert all((a == b).all() for a, b in zip(first, out)), "multi-GPU result differs from the 1-GPU result"
    assert all((a == wav[i].numpy()).all() for i, a in enumerate(first)), "stacked-tensor call differs from the per-utterance call"
    v, vb = audio_s / best, audio_s / best_b
    base = base or v
    base_b = base_b or vb
    res["runs"][str(n)] = {"list_of_utterances": {"seconds": round(best, 4), "audio_s_per_s": round(v, 1), "speedup_vs_1": round(v / base, 3)},
                           "stacked_pinned_tensors": {"seconds": round(best_b, 4), "audio_s_per_s": round(vb, 1), "speedup_vs_1": round(vb / base_b, 3)}}
    n *= 2
print(json.dumps(res))
