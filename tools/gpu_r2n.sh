#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -x -q --durations=4 > $O/r2n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2n_pytest.log
tail -9 $O/r2n_pytest.log
python bench.py --steps 5 --warmup 3 > $O/r2n_bench_n1.json 2> $O/r2n_bench_n1.err; echo "bench rc=$?"; tail -3 $O/r2n_bench_n1.err
python -c "
import json
d = json.loads(open('gpurun_out/r2n_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
for k, v in d['secondary'].items(): print('  ', k, {kk: vv for kk, vv in v.items() if kk not in ('config', 'cpu_baseline', 'e2e_c_abi', 'e2e')})
"
