#!/bin/bash
# final verification of the round: the driver's own sequence -- GPU tests, smoke(), default bench, reference arm
mkdir -p gpurun_out; O=gpurun_out
timeout 500 python -m pytest tests/ -x -q -m gpu > $O/r3f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r3f_pytest.log; tail -3 $O/r3f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r3f_smoke.log
( time timeout 900 python bench.py > $O/r3f_bench_default.json 2> $O/r3f_bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3f_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['steps'], d['warmup'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_executed'), d['gpu_launches'], d['clocks'])
print({k: round(v['value']) for k, v in d['secondary'].items()})
PY
