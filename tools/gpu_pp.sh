#!/bin/bash
# pre/post kernels: parity tests, then the config-4 bench section at a few resize tile budgets
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_prepost.py tests/test_gpu_edge_cases.py
for kb in 100 64 48; do
  DLIMG_B200_RESIZE_SMEM_KB=$kb timeout 300 python bench.py --only prepost > gpurun_out/bench_pp_$kb.json 2> gpurun_out/bench_pp.err
  echo "budget $kb KB rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_pp_$kb.json") if l.startswith("{")][-1])["prepost"]
for k,v in d.items():
    if isinstance(v,dict): print("  ",k, round(v["ms"]*1000,1),"us", round(v["gbs"]),"GB/s", round(v["frac_of_hbm"],3))
PY
done
