"""Turns the raw ncu output of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01        (run in the build container: needs `ncu -i` for the .ncu-rep files)
"""
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = "gpurun_out", "profiles"


def metric_rows(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[h]
    out = {}
    for r in rows[h + 1:]:
        d = dict(zip(hdr, r))
        k = int(d["ID"])
        out.setdefault(k, {"kernel": d["Kernel Name"].split("(")[0].replace("void ", "")})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    return [out[k] for k in sorted(out)]


# 1. launch list of one forward
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches_one_forward_cfg2.csv"))

# 2. per-launch metrics of the fused kernels
rows = metric_rows(os.path.join(G, "fused_metrics.csv"))


def entry(r, **extra):
    t = r["gpu__time_duration.sum"] / 1e3
    e = {"kernel": r["kernel"], **extra, "us": round(t, 1), "dram_read_MB": round(r["dram__bytes_read.sum"] / 1e6, 1),
         "dram_write_MB": round(r["dram__bytes_write.sum"] / 1e6, 1), "l2_MB": round(r["lts__t_bytes.sum"] / 1e6, 1),
         "tensor_pipe_active_pct": round(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"], 1),
         "issue_active_pct": round(r.get("sm__issue_active.avg.pct_of_peak_sustained_elapsed", 0.0), 1),
         "regs": int(r["launch__registers_per_thread"]), "grid": int(r["launch__grid_size"])}
    if "launch__shared_mem_per_block_dynamic" in r:
        e["smem_KB"] = round(r["launch__shared_mem_per_block_dynamic"] / 1024, 1)
    return e


# launch order of the second forward: with branch-parallel streams the C >= 128 steps interleave across branches, so the
# (C, k) label comes from the kernel family + order within the family
per_launch, per_stage = [], {}
n_pair = 0
n_res = 0
for r in rows:
    kn = r["kernel"]
    if "pair_tc_kernel" in kn:
        c = 256 if n_pair < 9 else 128
        n_pair += 1
        e = entry(r, C=c, what="fused ResBlock step")
    elif "respk_tc_kernel" in kn:
        c = 32
        e = entry(r, C=c, what="time-packed stage kernel (3 ResBlocks)")
    else:
        c = 64 if n_res < 3 else 16
        k = [3, 7, 11][n_res % 3]
        n_res += 1
        e = entry(r, C=c, k=k, what="whole ResBlock")
    per_launch.append(e)
    s_ = per_stage.setdefault(f"C={c}", {"launches": 0, "sum_us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "l2_MB": 0.0, "tw": 0.0})
    s_["launches"] += 1; s_["sum_us"] += e["us"]; s_["dram_read_MB"] += e["dram_read_MB"]; s_["dram_write_MB"] += e["dram_write_MB"]
    s_["l2_MB"] += e["l2_MB"]; s_["tw"] += e["tensor_pipe_active_pct"] * e["us"]
for s_ in per_stage.values():
    s_["tensor_pipe_active_pct"] = round(s_.pop("tw") / s_["sum_us"], 1)
    s_["dram_TBps"] = round((s_["dram_read_MB"] + s_["dram_write_MB"]) / s_["sum_us"], 2)   # MB / us = TB/s
    for k in ("sum_us", "dram_read_MB", "dram_write_MB", "l2_MB"):
        s_[k] = round(s_[k], 1)
json.dump({"note": "ncu metrics pass over the 18 fused ResBlock-step launches (C = 256, 128), the 6 whole-ResBlock launches (C = 64, 16) and the "
                   "time-packed stage launch (C = 32) of one cfg2 forward (16 x 4 s, bf16); cold-cache serialized durations",
           "per_stage": per_stage, "per_launch": per_launch}, open(os.path.join(P, f"{tag}_fused_steps_metrics.json"), "w"), indent=1)
tot_r = sum(e["dram_read_MB"] for e in per_launch) * 1e6
tot_w = sum(e["dram_write_MB"] for e in per_launch) * 1e6
json.dump({"kernel": f"pair_tc_kernel + res_tc_kernel + respk_tc_kernel ({len(per_launch)} launches of one cfg2 forward)", "dram_bytes_read_sum": tot_r,
           "dram_bytes_write_sum": tot_w, "traffic_bytes_per_launch_avg": (tot_r + tot_w) / len(per_launch),
           "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the {len(per_launch)} fused launches (profiles/{tag}_fused_steps_metrics.json); "
                     f"--set full captures: profiles/{tag}_ncu_full_pair_stage1.json, profiles/{tag}_ncu_full_respk_stage3.json"},
          open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
if os.path.exists(os.path.join(G, "small_metrics.csv")):
    small = [entry(r) for r in metric_rows(os.path.join(G, "small_metrics.csv"))]
    for e in small:
        e["dram_GBps"] = round((e["dram_read_MB"] + e["dram_write_MB"]) / e["us"] * 1e3, 1)
    json.dump({"note": "front end (spk_project_kernel, cond_multi_kernel), conv_pre + 5 upsamplers (conv_tc_kernel) and the head (post_kernel) of one cfg2 forward",
               "per_launch": small}, open(os.path.join(P, f"{tag}_small_kernels_metrics.json"), "w"), indent=1)

# 3. the two full captures: headline metrics + hottest instructions
KEEP = ["sm__issue_active.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for rep, what in (("prof_pair_stage1", "a fused ResBlock step of stage 1 (C=128)"),
                  ("prof_respk_stage3", "time-packed stage kernel of stage 3 (C=32, three ResBlocks k=3/7/11, dilations 1/3/5)"),
                  ("prof_conv_ups1", "polyphase ConvTranspose1d ups.1 (256 -> 128, k=8, u=4) in conv_tc_kernel"),
                  ("prof_post", "post_rows_kernel: leaky-ReLU(0.01) -> conv_post -> tanh, one row per thread"),
                  ("prof_cond", "cond_multi_kernel: speaker projection + unit gather -> table-folded ConvT -> GELU -> fc -> concat (the whole front end in one launch)"),
                  ("prof_pk16k11", "time-packed kernel, one ResBlock per launch (pk_fuse=0, pk_chan=112): C=16, k=11"),
                  ("prof_res_stage2_k11", "res_tc_kernel: whole ResBlock C=64, k=11 (one CTA per SM, sixteen epilogue warps, CTA pairs), the default plan"),
                  ("prof_resq_stage2_k11", "resq_tc_kernel: the same ResBlock with the skewed schedule (pack=0 res_mode=2 res_skew=1; off by default)")):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rr[0], rr[1], rr[2]
    out = {"capture": f"ncu --set full --clock-control none, {what}, cfg2 forward", "Kernel Name": vals[hdr.index("Kernel Name")]}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            out[f"{k} [{units[i]}]" if units[i] else k] = vals[i]
    json.dump(out, open(os.path.join(P, f"{tag}_ncu_full_{rep[5:]}.json"), "w"), indent=1)
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/_src.csv", "w").write(src)
    top = subprocess.run([sys.executable, "tools/ncu_top.py", "/tmp/_src.csv", "0", "40"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_ncu_full_{rep[5:]}_top_stalls.txt"), "w").write(top)
for f, dst in (("parity.txt", f"{tag}_parity.txt"), ("bench_bf16.json", f"{tag}_bench_line.json"), ("bench_reference.json", f"{tag}_bench_reference_line.json")):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, dst))
print("profiles written")
