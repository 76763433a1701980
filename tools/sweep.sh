set +e
C=c64k3c1,c64k3c2,c64k11c2,c32k3c2,c16k3c1,c16k3c2,c16k11c2,c128k3c1,c128k3c2,c128k11c2,c256k3c1
for k in "dual=0" "dual=1"; do echo "== $k"; python tools/conv_bench.py --cases $C --knob $k --iters 8; done > gpurun_out/sweep.log 2>&1
python tools/conv_probe.py 1 > gpurun_out/probe.log 2>&1; tail -n 1 gpurun_out/probe.log
