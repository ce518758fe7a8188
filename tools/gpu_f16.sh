#!/bin/bash
# f16 embedding download check: API suite + three short bench runs (e2e / fp16 download / resident legs).
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_api.py
for i in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 3 --only none --cpu-sample 0 > gpurun_out/bench.json 2> gpurun_out/bench.err
  python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench.json').read().splitlines() if l.startswith('{')][-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'f16', round(d['e2e_f16_download']['value']), 'resident', round(d['e2e_resident']['value']))
PY
done
