#!/bin/bash
# A/B of programmatic dependent launch per kernel family (DLIMG_B200_PDL_MASK), one quick bench line each.
for m in 0 0x7f 0x07 0x01 0x03 0x0f 0x1f 0x3f 0x47 0 0x7f; do
  DLIMG_B200_PDL_MASK=$m python bench.py --steps 10 --warmup 3 --quick 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mask $m images/s %.1f ms %.3f' % (d['value'], d['ms_per_step']))"
done
