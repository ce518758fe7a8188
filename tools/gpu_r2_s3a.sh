#!/bin/bash
# Session-3 check (dev build): API + encoder suites, lazy-softmax A/B, latency leg with library images.
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_api.py tests/test_gpu_encoder.py tests/test_gpu_dropin.py
tools/gpu_ab.sh DLIMG_B200_WA_MODE=0 3
for m in "" 0; do
  if [ -z "$m" ]; then out=$(python bench.py --steps 10 --warmup 3 --only none --cpu-sample 0 2>/dev/null | tail -n 1); else out=$(DLIMG_B200_WA_MODE=$m python bench.py --steps 10 --warmup 3 --only none --cpu-sample 0 2>/dev/null | tail -n 1); fi
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode[$m]', {k: round(v['ms_per_step'],3) for k,v in d.get('kernels',{}).items()})"
done
timeout 300 python bench.py --steps 5 --warmup 3 --only latency --cpu-sample 0 > gpurun_out/bench_lat.json 2> gpurun_out/bench_lat.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_lat.json').read().splitlines() if l.startswith('{')][-1])
for k, v in d['latency'].items():
    if isinstance(v, dict) and 'median_ms' in v: print(k, round(v['median_ms'], 3))
    elif isinstance(v, float): print(k, round(v, 3))
PY
tail -n 3 gpurun_out/bench_lat.err
