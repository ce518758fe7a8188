#!/bin/bash
# usage: tools/gpu_sweep.sh VAR v1 v2 ...   -- one quick bench line per value of an environment variable (same box)
var=$1; shift
for v in "$@"; do
  env "$var=$v" python bench.py --steps 10 --warmup 3 --quick 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$var=$v images/s %.1f ms %.3f' % (d['value'], d['ms_per_step']))"
done
