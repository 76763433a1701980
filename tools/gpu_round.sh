#!/bin/bash
# One GPU-box session for the ncu evidence of a round: launch list of one forward, per-kernel metrics of the fused
# launches, full captures of the heaviest kernel of every family.  Everything lands in gpurun_out/ (< 64 MiB).
set +e
mkdir -p gpurun_out
python tools/one_forward.py 3 > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
# launch list of the third forward (33 launches per forward in bf16 mode)
ncu --metrics gpu__time_duration.sum --clock-control none -s 66 -c 33 --csv --log-file gpurun_out/launches.csv python tools/one_forward.py 3 > gpurun_out/ncu_launch.log 2>&1
# metrics of the 18 fused-step + 6 whole-ResBlock + 1 time-packed stage launches of the second forward
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size --clock-control none -k "regex:pair_tc_kernel|res_tc_kernel|respk_tc_kernel" -s 25 -c 25 --csv --log-file gpurun_out/fused_metrics.csv python tools/one_forward.py 2 > gpurun_out/ncu_fused.log 2>&1
# small kernels: front end, upsamplers, head
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size --clock-control none -k "regex:conv_tc_kernel|post_kernel|post_rows_kernel|cond_multi_kernel|spk_project_kernel" -s 8 -c 8 --csv --log-file gpurun_out/small_metrics.csv python tools/one_forward.py 2 > gpurun_out/ncu_small.log 2>&1
# full captures: heaviest fused step (stage 1, k = 11), the time-packed stage kernel (C = 32), an upsampler (ups.1), the head
ncu --set full --clock-control none --import-source on -k regex:pair_tc_kernel -s 33 -c 1 -o gpurun_out/prof_pair_stage1 python tools/one_forward.py 2 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:respk_tc_kernel -s 1 -c 1 -o gpurun_out/prof_respk_stage3 python tools/one_forward.py 2 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 8 -c 1 -o gpurun_out/prof_conv_ups1 python tools/one_forward.py 2 > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none -k "regex:post_rows_kernel|post_kernel" -s 1 -c 1 -o gpurun_out/prof_post python tools/one_forward.py 2 > gpurun_out/ncu_full4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cond_multi_kernel -s 1 -c 1 -o gpurun_out/prof_cond python tools/one_forward.py 2 > gpurun_out/ncu_full5.log 2>&1
du -sh gpurun_out; ls gpurun_out
