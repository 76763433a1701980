#!/bin/bash
# One GPU-box session: kernel probes, parity tests, bench.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
python tools/conv_probe.py 1 > gpurun_out/probe.log 2>&1
python -m pytest tests -q -s -m gpu > gpurun_out/pytest_all.log 2>&1
grep "\[parity\]" gpurun_out/pytest_all.log > gpurun_out/parity.txt
python bench.py --steps 20 --warmup 5 --layers > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
tail -n 3 gpurun_out/probe.log; tail -n 3 gpurun_out/pytest_all.log
cat gpurun_out/bench_bf16.json
