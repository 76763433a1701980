"""Runs a grid of tap-offset convolution cases against every kernel variant, one
process per case (a faulting tcgen05 variant must not poison the rest), and
writes gpurun_out/conv_probe.json.  GPU box only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # name, kwargs
    ("c64_k3_d1", dict(cin=64, cout=64, k=3, dil=1, lin=300)),
    ("c64_k7_d3", dict(cin=64, cout=64, k=7, dil=3, lin=1000)),
    ("c128_k3_d1", dict(cin=128, cout=128, k=3, dil=1, lin=300)),
    ("c128_k11_d5", dict(cin=128, cout=128, k=11, dil=5, lin=700)),
    ("c256_k3_d1", dict(cin=256, cout=256, k=3, dil=1, lin=300)),
    ("c256_k11_d5", dict(cin=256, cout=256, k=11, dil=5, lin=500)),
    ("c32_k7_d3", dict(cin=32, cout=32, k=7, dil=3, lin=3000)),
    ("c16_k11_d5", dict(cin=16, cout=16, k=11, dil=5, lin=5000)),
    ("c16_k3_d1", dict(cin=16, cout=16, k=3, dil=1, lin=200)),
    ("pre_336_512_k7", dict(cin=336, cout=512, k=7, dil=1, lin=400, use_res=False, use_acc=False, div=1.0)),
    ("up0_512_256_k11_u5", dict(cin=512, cout=256, k=11, up=5, lin=200, use_res=False, use_acc=False, div=1.0)),
    ("up1_256_128_k8_u4", dict(cin=256, cout=128, k=8, up=4, lin=300, use_res=False, use_acc=False, div=1.0)),
    ("up4_32_16_k4_u2", dict(cin=32, cout=16, k=4, up=2, lin=1500, use_res=False, use_acc=False, div=1.0)),
    ("tiny_c16_l5", dict(cin=16, cout=16, k=3, dil=1, lin=5, batch=1)),
]


def main():
    impls = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]
    results = {}
    for impl in impls:
        batch = [[name, dict(kw, impl=impl)] for name, kw in CASES]
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "convcase.py"), json.dumps(batch)],
                               capture_output=True, text=True, timeout=600)
            out, tail = p.stdout, (p.stderr or "")[-600:]
        except subprocess.TimeoutExpired as e:
            out, tail = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "timeout"
        seen = set()
        for line in out.splitlines():
            if line.startswith("RESULT "):
                name, res = json.loads(line[7:])
                results[f"impl{impl}/{name}"] = res
                seen.add(name)
                print(f"impl{impl}/{name}: {res}", flush=True)
        for name, _ in CASES:
            if name not in seen:
                results[f"impl{impl}/{name}"] = {"ok": False, "error": "not reached: " + tail}
                print(f"impl{impl}/{name}: not reached ({tail[-200:]})", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "conv_probe.json"), "w") as f:
        json.dump(results, f, indent=1)
    bad = [k for k, v in results.items() if not v.get("ok")]
    print("FAILED:", bad)


if __name__ == "__main__":
    main()
