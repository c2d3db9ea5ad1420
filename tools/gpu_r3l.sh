#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $O/r3l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3l_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mgpu --no-wnaf-e2e 2>$O/r3l_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k in ('pairing_product_small', 'single_pairing'): print(k, {kk: vv for kk, vv in d['secondary'][k].items() if kk not in ('config', 'cpu_baseline')})
"; tail -3 $O/r3l_bench.err
