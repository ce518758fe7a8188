#!/bin/bash
# In-box A/B of an environment switch (box-to-box spread is ~2 %, so compare within one call):
# usage: tools/gpu_ab.sh VAR=VALUE [rounds]   -- alternates "unset" and "VAR=VALUE", one quick bench line each
kv=$1; n=${2:-3}
for i in $(seq $n); do
  for mode in base "$kv"; do
    if [ "$mode" == "base" ]; then out=$(python bench.py --steps 10 --warmup 3 --quick 2>/dev/null | tail -n 1)
    else out=$(env "$kv" python bench.py --steps 10 --warmup 3 --quick 2>/dev/null | tail -n 1); fi
    echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode images/s %.1f ms %.3f' % (d['value'], d['ms_per_step']))"
  done
done
