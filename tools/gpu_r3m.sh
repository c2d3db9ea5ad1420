#!/bin/bash
# final state of round 2: the driver's sequence (GPU tests, smoke, bench N=1 with every secondary, reference arm)
mkdir -p gpurun_out; O=gpurun_out
timeout 500 python -m pytest tests/ -x -q -m gpu > $O/r3m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3m_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r3m_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r3m_bench_n1.json 2> $O/r3m_bench_n1.err; echo "bench rc=$?"; tail -2 $O/r3m_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r3m_bench_ref.json 2> $O/r3m_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3m_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_executed'), d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for k, v in d['secondary'].items(): print('  ', k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ('config', 'cpu_baseline', 'e2e_c_abi', 'e2e', 'phase_ms')})
r = json.loads(open('gpurun_out/r3m_bench_ref.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['cpu_baseline']['cores'])
PY
