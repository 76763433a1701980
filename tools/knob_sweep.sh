#!/bin/bash
# usage: tools/knob_sweep.sh knob v1 v2 ...   -> ms per step and per-stage ms of the cfg2 bench for every value
k=$1; shift
for v in "$@"; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers --knob $k=$v 2>/dev/null | tail -1 > /tmp/ks.json
  python - "$k=$v" <<'PY'
import json, sys
d = json.load(open("/tmp/ks.json"))
ps = d["roofline"]["per_stage"]
print(sys.argv[1], d["ms_per_step"], {k: v["ms"] for k, v in ps.items()})
PY
done
