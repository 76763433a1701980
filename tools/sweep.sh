set +e
C=c64k3c1,c64k3c2,c16k3c1,c16k3c2,c128k3c2,c256k3c1
python tools/conv_bench.py --cases c64k3c1,c64k3c2,c16k3c1 --trace --iters 4 > gpurun_out/trace.log 2>&1
for k in "max_msub=1" "max_msub=2" "max_msub=4" "sa_min=4" "slab_cap=81920" "max_ctas=296"; do echo "== $k"; python tools/conv_bench.py --cases $C --knob $k --iters 8; done > gpurun_out/sweep.log 2>&1
