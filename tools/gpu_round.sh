#!/bin/bash
# One GPU-box session: parity tests, smoke, bench.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
python -m pytest tests -q -s -m gpu > gpurun_out/pytest_all.log 2>&1
grep "\[parity\]" gpurun_out/pytest_all.log > gpurun_out/parity.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 20 --warmup 5 --layers > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
tail -n 3 gpurun_out/pytest_all.log; tail -n 2 gpurun_out/smoke.log
cat gpurun_out/bench_bf16.json
