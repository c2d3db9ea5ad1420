#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $O/r3j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3j_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-mgpu --no-secondary 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('device', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
