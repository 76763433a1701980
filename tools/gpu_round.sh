#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list, per-kernel metrics, two full captures.
# Everything lands in gpurun_out/ (keep it under 64 MiB: --set full for two launches only).
set +e
mkdir -p gpurun_out
python -m pytest tests -q -s -m gpu > gpurun_out/pytest_all.log 2>&1
grep "\[parity\]" gpurun_out/pytest_all.log > gpurun_out/parity.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 50 --warmup 5 --layers > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# launch list of one forward (36 launches per forward in bf16 mode; the third forward of the process is listed)
python tools/one_forward.py 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 72 -c 36 --csv --log-file gpurun_out/launches_final.csv python tools/one_forward.py 3 > gpurun_out/ncu_launch.log 2>&1
# metrics of the 18 fused-step + 9 whole-ResBlock launches of the second forward
python tools/one_forward.py 2 > gpurun_out/plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size --clock-control none -k "regex:pair_tc_kernel|res_tc_kernel" -s 27 -c 27 --csv --log-file gpurun_out/fused_metrics.csv python tools/one_forward.py 2 > gpurun_out/ncu_pairs.log 2>&1
# full captures: the heaviest fused step (stage 1, k = 11) and the heaviest whole-ResBlock launch (stage 2, k = 11)
ncu --set full --clock-control none --import-source on -k regex:pair_tc_kernel -s 33 -c 1 -o gpurun_out/prof_pair_stage1 python tools/one_forward.py 2 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:res_tc_kernel -s 11 -c 1 -o gpurun_out/prof_res_stage2 python tools/one_forward.py 2 > gpurun_out/ncu_full2.log 2>&1
du -sh gpurun_out
tail -n 3 gpurun_out/pytest_all.log; tail -n 2 gpurun_out/smoke.log
cat gpurun_out/bench_bf16.json
