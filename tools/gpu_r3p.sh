#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "mul or wnaf or cpp or affine" > $O/r3p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3p_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mgpu --no-wnaf-e2e 2>$O/r3p_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k in ('g1_mul_assign', 'g1_wnaf_mul'): print(k, {kk: vv for kk, vv in d['secondary'][k].items() if kk not in ('config', 'cpu_baseline')})
"; tail -3 $O/r3p_bench.err
