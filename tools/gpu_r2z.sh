#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "pow or fixed_base_and_gt or cpp" > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2z_pytest.log
tail -3 $O/r2z_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mgpu --no-wnaf-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k in ('gt_pow', 'pairing_shared_q'): print(k, {kk: vv for kk, vv in d['secondary'][k].items() if kk not in ('config', 'cpu_baseline')})
" | tee $O/r2z_bench.log
