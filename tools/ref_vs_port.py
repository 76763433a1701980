"""The reference's own MelCodeGenerator (imported unmodified from /root/reference, build container only) against the
oracle port on the same weights and inputs: output agreement and CPU time, batch 1, fp32.  This is why bench.py's
reference arm may time the port on the GPU box (where the reference tree cannot travel): same ATen operators, same speed.

    python tools/ref_vs_port.py > profiles/r02_reference_vs_port.json
"""
import json
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402

ref = bench.import_reference_classes()
assert ref is not None, "needs /root/reference"
cls, attr = ref
torch.set_num_threads(os.cpu_count())
h = vo.shipped_config()
sd = vo.init_state_dict(h, seed=1234, style="trained")
code, mel, spkr = vo.synthetic_inputs(8, 400, seed=52)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    g = cls(attr(dict(h)))
    g.load_state_dict(sd, strict=True)
    g.eval()
    g.remove_weight_norm()
w = vo.fold_weight_norm(sd)


def t_ref(i):
    return g(code=code[i:i + 1], mel=mel[i:i + 1], spkr=spkr[i:i + 1])


def t_port(i):
    return vo.mel_code_generator_forward(w, h, code[i:i + 1], mel[i:i + 1], spkr[i:i + 1], dtype=torch.float32)


out = {"cores": os.cpu_count(), "torch": torch.__version__, "workload": "8 utterances x 4 s (T=400), batch 1, fp32"}
with torch.no_grad():
    a, b = t_ref(0), t_port(0)
    out["max_abs_reference_vs_port"] = float((a - b).abs().max())
    for name, fn in (("reference_classes", t_ref), ("oracle_port", t_port), ("reference_classes_again", t_ref)):
        t0 = time.perf_counter()
        for rep in range(2):
            for i in range(8):
                fn(i)
        dt = time.perf_counter() - t0
        out[name] = {"seconds": round(dt, 2), "audio_s_per_s": round(2 * 8 * 4.0 / dt, 2)}
print(json.dumps(out, indent=1))
