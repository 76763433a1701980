"""Benchmark of the multi_input_vocoder generator forward on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is one forward of the hot path over one batch of synthetic utterances
(BASELINE.md section 4 distribution, seed 52) with random-init weights of the shipped
vocoder config (seed 1234).  Per-GPU workload = BASELINE.json configs[1]:
16 utterances x 4 s (T = 400 mel frames, U = 200 units), bf16 tensor-core mode.
Utterances are independent, so N GPUs run N such batches with no collective on
the data path (weak scaling); torch.distributed is used only for the barrier and
the max-over-ranks of the device time.

Prints ONE JSON line (rank 0).  `value` = audio-seconds generated per second with
inputs resident in HBM; `e2e` = the same through the public class with pinned host
inputs copied in and the waveform copied out every step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH, FRAMES = 16, 400          # configs[1]: 16 x 4 s
SR, HOP = 16000, 160
L2_MB = 126


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained") or d["bf16_tflops"]), hbm=float(d["hbm_gbs"]),
                    src="MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md sustained figure)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(self.NAMES, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic():
    """dram read+write bytes per launch of the dominant kernel, from the committed ncu metrics pass (or None)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["traffic_bytes_per_launch_avg"]), d["source"]
    except Exception:
        return None, None


def cpu_reference(seconds_budget=12.0):
    """The reference's CPU path (oracle port: the same ATen ops as the reference's
    torch modules, fp32, all host threads) on a bounded sample of the workload:
    4 s utterances one at a time (the reference's own inference loop is batch 1,
    which is also its fastest CPU shape), repeated until ~seconds_budget."""
    from oracle import vocoder_oracle as vo
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    h = vo.shipped_config()
    w = vo.fold_weight_norm(vo.init_state_dict(h, seed=1234, style="ref"))
    code, mel, spkr = vo.synthetic_inputs(BATCH, FRAMES, seed=52)
    with torch.no_grad():
        vo.mel_code_generator_forward(w, h, code[:1], mel[:1], spkr[:1], dtype=torch.float32)   # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            i = n % BATCH
            vo.mel_code_generator_forward(w, h, code[i:i + 1], mel[i:i + 1], spkr[i:i + 1], dtype=torch.float32)
            n += 1
            el = time.perf_counter() - t0
            if el >= seconds_budget or n >= 4 * BATCH:
                break
    audio = n * FRAMES * HOP / SR
    return {"value": audio / el, "unit": "audio-s/s", "cores": threads, "kind": "port",
            "sample": f"{n} utterances x 4 s (T=400), batch 1, fp32, torch {torch.__version__} CPU, {el:.1f} s"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    res, total_audio, total_t = None, 0.0, 0.0
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        res = cpu_reference(per_step_budget)
        if i >= args.warmup:
            n = int(res["sample"].split()[0])
            a = n * FRAMES * HOP / SR
            total_audio += a
            total_t += a / res["value"]
    value = total_audio / total_t
    res["value"] = value
    line = {"impl": "reference", "metric": "audio-sec generated/sec (16 kHz)", "value": value, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: 16 x 4 s utterances (T=400,U=200), multi_input_aug.json generator, "
                                   "random-init seed 1234; each step a bounded batch-1 sample of it on the host cores"},
            "cpu_baseline": res,
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--knob", action="append", default=[], help="debug knob k=v passed to l2s_debug_set")
    ap.add_argument("--layers", action="store_true", help="also print a per-launch time table to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    from oracle import vocoder_oracle as vo      # weights / synthetic inputs / cpu_baseline leg only
    import __graft_entry__ as ge

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ge.build()
    pkg = ge.load_package()
    lib = pkg._cabi.load()
    for kv in args.knob:
        k, v = kv.split("=")
        assert lib.l2s_debug_set(k.encode(), int(v)) == 0, kv

    h = vo.shipped_config()
    sd = vo.init_state_dict(h, seed=1234, style="ref")
    gen = pkg.MelCodeGenerator(pkg.AttrDict(h))
    gen.load_state_dict(sd, strict=True)
    gen.eval()
    gen.remove_weight_norm()
    gen.set_precision(args.precision)
    gen = gen.to(dev)

    # every rank vocodes its own 16 x 4 s shard (seed differs per rank)
    code_h, mel_h, spk_h = vo.synthetic_inputs(BATCH, FRAMES, seed=52 + rank)
    code_h, mel_h, spk_h = code_h.pin_memory(), mel_h.pin_memory(), spk_h.pin_memory()
    code, mel, spk = code_h.to(dev), mel_h.to(dev), spk_h.to(dev)
    audio_per_step = BATCH * FRAMES * HOP / SR
    flops_per_step = vo.algorithmic_flops_per_frame(h) * BATCH * FRAMES

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                     # keeps sampling through warm-up, the timed region and the e2e loop
    for _ in range(args.warmup):
        y = gen(code=code, mel=mel, spkr=spk)
    launches_per_step = gen.launch_count(BATCH, FRAMES, dev)
    ws_bytes = lib.l2s_workspace_bytes(gen._engine(dev).handle, BATCH, FRAMES)

    # ---- device-resident timed region
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        y = gen(code=code, mel=mel, spkr=spk)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    # host time to ISSUE one step (no synchronisation): if it approaches ms_per_step the path is launch bound
    t_h0 = time.perf_counter()
    for _ in range(args.steps):
        y = gen(code=code, mel=mel, spkr=spk)
    host_ms = (time.perf_counter() - t_h0) * 1e3 / args.steps
    torch.cuda.synchronize()

    # ---- end to end through the public API: every step copies its inputs from pinned host memory to the device and
    # its waveform back to pinned host memory.  HostPipeline (dispatch.py) is the call a user with many batches makes:
    # the copies of neighbouring steps overlap the forward (two copy streams, double-buffered device inputs).
    out_h = [torch.empty((BATCH, 1, FRAMES * HOP), dtype=torch.float32).pin_memory() for _ in range(2)]
    pipe = pkg.HostPipeline(gen, dev)
    for i in range(4):
        pipe.submit(code_h, mel_h, spk_h, out_h[i & 1])
    pipe.finish()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        pipe.submit(code_h, mel_h, spk_h, out_h[i & 1])
    pipe.s_out.synchronize()
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)
    pipe.finish()
    if not torch.equal(out_h[(args.steps - 1) & 1], y.cpu()):
        raise RuntimeError("end-to-end pipeline output differs from the device-resident forward")
    # short steps: keep the GPU under the same load until nvidia-smi has delivered a few samples
    t_end = time.time() + 4.0
    while len(sampler.rows) < 8 and time.time() < t_end:
        for _ in range(10):
            gen(code=code, mel=mel, spkr=spk)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    h2d = code_h.numel() * 8 + mel_h.numel() * 4 + spk_h.numel() * 4
    d2h = out_h[0].numel() * 4

    # ---- per-launch times of one more forward (event pair per launch), for the roofline object
    lib.l2s_debug_set(b"layer_events", 1)
    gen(code=code, mel=mel, spkr=spk)
    torch.cuda.synchronize()
    import ctypes as C
    eng = gen._engine(dev)
    rows, i = [], 0
    while True:
        t, fl, nm = C.c_float(), C.c_double(), C.create_string_buffer(64)
        if lib.l2s_debug_layer_time(eng.handle, i, C.byref(t), C.byref(fl), nm, 64) != 0:
            break
        rows.append((nm.value.decode(), t.value, fl.value))
        i += 1
    lib.l2s_debug_set(b"layer_events", 0)

    t_max = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_max[0]), float(t_max[1])

    if rank == 0:
        pk = peaks()
        step_s = ms / 1e3 / args.steps
        conv_rows = [r for r in rows if r[2] > 0 and r[0] != "conv_post"]
        conv_ms = sum(r[1] for r in conv_rows)
        conv_flops = sum(r[2] for r in conv_rows)
        all_ms = sum(r[1] for r in rows)
        achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        stages = {}
        for name, t, fl in rows:
            if name.startswith("resblocks."):
                key = "mrf%d" % (int(name.split(".")[1].split()[0]) // len(h["resblock_kernel_sizes"]))
            elif name.startswith("ups."):
                key = "ups"
            else:
                key = name
            a = stages.setdefault(key, [0.0, 0.0])
            a[0] += t
            a[1] += fl
        stage_tbl = {k: {"ms": round(v[0], 4), "tflops": round(v[1] / (v[0] / 1e3) / 1e12, 1) if v[0] > 0 else 0.0}
                     for k, v in stages.items()}
        if args.layers:
            for name, t, fl in rows:
                sys.stderr.write(f"{name:28s} {t * 1e3:9.1f} us  {fl / max(t, 1e-9) / 1e9:8.1f} TFLOP/s\n")
        kernel = {"bf16": "pair_tc_kernel (fused ResBlock step, C >= 128) + res_tc_kernel (whole ResBlock, C <= 64) + conv_tc_kernel (conv_pre, ups): tcgen05 tap-offset convolutions",
                  "tf32": "conv_tc_kernel (tcgen05 kind::tf32 tap-offset conv)", "fp32": "conv_simt_kernel"}[args.precision]
        traffic, traffic_src = ncu_traffic() if args.precision == "bf16" else (None, None)
        line = {
            "metric": "audio-sec generated/sec (16 kHz)",
            "value": world * audio_per_step / step_s,
            "unit": "audio-s/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": args.precision,
            "data": "synthetic",
            "config": {"workload": "cfg2 per GPU: 16 x 4 s utterances (T=400 mel frames, U=200 KM200 units, 256-d spk emb), "
                                   "multi_input_aug.json generator, random-init seed 1234",
                       "global_batch": world * BATCH, "audio_s_per_step": world * audio_per_step,
                       "parallelism": f"utterance-sharded x{world}, no collective",
                       "l2": f"no flush: per-step activation working set {ws_bytes / 2**20:.0f} MiB > {L2_MB} MB L2"},
            "tensor_frac_whole_step": flops_per_step / step_s / 1e12 / pk["tflops"],
            "roofline": {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": pk["src"],
                         "launches": len(conv_rows), "sum_launch_ms": round(conv_ms, 4), "all_launch_ms": round(all_ms, 4),
                         "algorithmic_gflop_per_step": conv_flops / 1e9, "per_stage": stage_tbl},
            "e2e": {"value": world * audio_per_step / (ms_e2e / 1e3 / args.steps), "unit": "audio-s/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "api": "HostPipeline.submit(code, mel, spkr, out) per step: pinned host inputs -> device, MelCodeGenerator forward, fp32 waveform -> pinned host; copies of neighbouring steps overlap the forward"},
            "gpu_launches": launches_per_step * args.steps,
            "host_issue_ms_per_step": host_ms,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(12.0)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
