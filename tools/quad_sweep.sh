#!/bin/bash
# per-launch times of the whole-ResBlock kernels for several values of a knob:  tools/quad_sweep.sh knob v1 v2 ...
k=$1; shift
for v in "$@"; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers --knob $k=$v 2>/tmp/l_q.txt | tail -1 > /tmp/ks.json
  python - "$k=$v" <<'PY'
import json, sys
d = json.load(open("/tmp/ks.json"))
print(sys.argv[1], round(d["ms_per_step"], 4), end="  ")
PY
  grep whole /tmp/l_q.txt | awk '{print $3}' | tr "\n" " "; echo
done
