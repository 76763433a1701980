"""Times single tap-offset convolution launches (tcgen05 kernel, l2s_debug_conv) at
the cfg2 layer shapes, rotating over enough buffer sets that nothing stays in L2.
GPU box only.   python tools/conv_bench.py [--cases c64k3c1,...] [--iters 20] [--knob k=v ...]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

B = 16
# name: (cin, cout, k, dil, rows per utterance, kind)  kind: c1 = act only, c2 = residual + raw + act
SHAPES = {
    "c256k3c1": (256, 256, 3, 1, 2000, "c1"), "c256k3c2": (256, 256, 3, 1, 2000, "c2"),
    "c256k11c1": (256, 256, 11, 5, 2000, "c1"), "c256k11c2": (256, 256, 11, 1, 2000, "c2"),
    "c128k3c1": (128, 128, 3, 1, 8000, "c1"), "c128k3c2": (128, 128, 3, 1, 8000, "c2"),
    "c128k11c1": (128, 128, 11, 5, 8000, "c1"), "c128k11c2": (128, 128, 11, 1, 8000, "c2"),
    "c64k3c1": (64, 64, 3, 1, 16000, "c1"), "c64k3c2": (64, 64, 3, 1, 16000, "c2"),
    "c64k11c1": (64, 64, 11, 5, 16000, "c1"), "c64k11c2": (64, 64, 11, 1, 16000, "c2"),
    "c32k3c1": (32, 32, 3, 1, 32000, "c1"), "c32k3c2": (32, 32, 3, 1, 32000, "c2"),
    "c32k11c1": (32, 32, 11, 5, 32000, "c1"),
    "c16k3c1": (16, 16, 3, 1, 64000, "c1"), "c16k3c2": (16, 16, 3, 1, 64000, "c2"),
    "c16k11c1": (16, 16, 11, 5, 64000, "c1"), "c16k11c2": (16, 16, 11, 1, 64000, "c2"),
    "tiny": (16, 16, 3, 1, 64, "c1"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(SHAPES))
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--impl", type=int, default=1)
    ap.add_argument("--knob", action="append", default=[])
    ap.add_argument("--out", default="")
    ap.add_argument("--burst", type=int, default=0, help="also time N back-to-back launches inside one event pair")
    ap.add_argument("--trace", action="store_true", help="dump per-role timestamps of CTA 0 for the last launch")
    args = ap.parse_args()
    pkg = ge.load_package()
    cabi = pkg._cabi
    lib = cabi.load()
    for kv in args.knob:
        k, v = kv.split("=")
        assert lib.l2s_debug_set(k.encode(), int(v)) == 0, kv
    dev = torch.device("cuda:0")
    results = {}
    for name in args.cases.split(","):
        cin, cout, k, dil, L, kind = SHAPES[name]
        elems = B * L * cout
        per_set = elems * (2 + 2 + (10 if kind == "c2" else 0))
        nsets = max(2, min(8, int(400e6 // per_set) + 1))
        w = (torch.randn(k, cout, cin, device=dev) / (cin * k) ** 0.5).bfloat16().contiguous()
        bias = torch.randn(cout, device=dev)
        sets = []
        for _ in range(nsets):
            x = torch.randn(B, L, cin, device=dev).bfloat16()
            act = torch.empty(B, L, cout, device=dev, dtype=torch.bfloat16)
            raw = torch.empty(B, L, cout, device=dev) if kind == "c2" else None
            res = torch.randn(B, L, cout, device=dev) if kind == "c2" else None
            sets.append((x, act, raw, res))
        descs = []
        for x, act, raw, res in sets:
            d = cabi.ConvDesc()
            d.inp, d.w, d.bias = x.data_ptr(), w.data_ptr(), bias.data_ptr()
            d.out_act = act.data_ptr()
            d.out_raw = raw.data_ptr() if raw is not None else None
            d.res = res.data_ptr() if res is not None else None
            d.acc_in = None
            d.act_bf16 = 1
            d.batch, d.lin, d.cin_pad, d.ntaps, d.ntot, d.mrows = B, L, cin, k, cout, L
            for j in range(k):
                d.tap_off[j] = (j - (k - 1) // 2) * dil
            d.out_shift, d.out_valid, d.scale, d.slope = 0, L * cout, 1.0, 0.1
            descs.append(d)
        err = C.create_string_buffer(512)
        stream = torch.cuda.current_stream().cuda_stream
        lib.l2s_debug_set(b"plan_report", 1)
        for d in descs[:2]:
            st = lib.l2s_debug_conv(C.byref(d), args.impl, 0, stream, err, 512)
            assert st == 0, err.value
        plan = err.value.decode()
        lib.l2s_debug_set(b"plan_report", 0)
        torch.cuda.synchronize()
        times = []
        for it in range(args.iters):
            d = descs[it % nsets]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.l2s_debug_conv(C.byref(d), args.impl, 0, stream, err, 512)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        if args.trace:
            tr = torch.zeros(3 * 64 * 4 + 512 * 3, dtype=torch.int64, device=dev)
            lib.l2s_debug_set(b"trace_ptr", tr.data_ptr())
            lib.l2s_debug_conv(C.byref(descs[0]), args.impl, 0, stream, err, 512)
            torch.cuda.synchronize()
            lib.l2s_debug_set(b"trace_ptr", 0)
            cta = tr.cpu()[768:].view(512, 3)
            live = cta[cta[:, 1] > 0]
            by_sm = {}
            for smid, a, b_ in live.tolist():
                by_sm.setdefault(smid, []).append((a, b_))
            overlap = sum(1 for v in by_sm.values() if len(v) > 1 and any(x[0] < y[1] and y[0] < x[1] for i_, x in enumerate(v) for y in v[i_ + 1:]))
            print(f'CTAs {len(live)} on {len(by_sm)} SMs; SMs with time-overlapping CTAs: {overlap}; kernel span {(int(live[:,2].max()) - int(live[:,1].min()))/1e3:.1f} us')
            t = tr.cpu()[:768].view(3, 64, 4)
            t0 = int(t[t > 0].min())
            print("trace (ns since first stamp) item: prod[wait_empty,issue] mma[wait_acc,got_acc,got_A,committed] epi[wait,got,done]")
            for i in range(64):
                if int(t[1, i, 3]) == 0:
                    break
                f = lambda v: int(v) - t0 if int(v) else -1
                print(f"  {i:2d}: P[{f(t[0,i,0])},{f(t[0,i,1])}] M[{f(t[1,i,0])},{f(t[1,i,1])},{f(t[1,i,2])},{f(t[1,i,3])}] "
                      f"E[{f(t[2,i,0])},{f(t[2,i,1])},{f(t[2,i,2])}]")
        burst_us = None
        if args.burst:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(args.burst):
                lib.l2s_debug_conv(C.byref(descs[it % nsets]), args.impl, 0, stream, err, 512)
            e1.record()
            torch.cuda.synchronize()
            burst_us = e0.elapsed_time(e1) * 1e3 / args.burst
        times.sort()
        med = times[len(times) // 2]
        flops = 2.0 * cin * cout * k * B * L
        byts = B * L * (cin * 2 + cout * 2 + (cout * 8 if kind == "c2" else 0))
        results[name] = dict(us=round(med, 1), min_us=round(times[0], 1), tflops=round(flops / med / 1e6, 1),
                             gbs=round(byts / med / 1e3, 1))
        if burst_us is not None:
            results[name]['burst_us'] = round(burst_us, 1)
        print(name, results[name], plan, flush=True)
        del sets, descs
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
