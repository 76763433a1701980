#!/bin/bash
# One GPU-box session: kernel probes, parity tests, bench.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python tools/conv_probe.py 0 1 2 3 > gpurun_out/probe.log 2>&1
python -m pytest tests -q -m gpu -k "fp32 or simt or error or int16 or lazy or fp16" > gpurun_out/pytest_fp32.log 2>&1
python -m pytest tests -q -m gpu > gpurun_out/pytest_all.log 2>&1
python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python bench.py --steps 10 --warmup 3 --layers > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
tail -3 gpurun_out/probe.log gpurun_out/pytest_fp32.log gpurun_out/pytest_all.log
cat gpurun_out/bench_bf16.json
