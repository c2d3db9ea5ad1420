#!/bin/bash
# final build of round 2: N=1 bench line, the same kernel back to back (no L2 flush), ncu of the fused kernel, launch list
mkdir -p gpurun_out; O=gpurun_out
export PROFILE_OUT_DIR=$O
timeout 400 python -m pytest tests -m gpu -x -q -k "pairing or miller or final or engine or cpp or wide" > $O/r2y_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2y_pytest.log
tail -3 $O/r2y_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2y_bench_n1.json 2> $O/r2y_bench_n1.err; echo "bench rc=$?"; tail -3 $O/r2y_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2y_bench_ref.json 2> $O/r2y_bench_ref.err; echo "ref rc=$?"
timeout 200 python tools/bench_paths.py --skip mm,g1,g2 --log2 16 2>&1 | grep "batch\|config" | tee $O/r2y_paths.log
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # tag kernel-regex skip n what
  timeout 300 $NCU -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/prof_pairing.py $4 $5 > $O/r2y_ncu_$1.log 2>&1 && python tools/summarize_profiles.py - /tmp/$1.ncu-rep r2y_$1 >> $O/r2y_ncu_$1.log 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap pair_miller 'k_pair_miller$' 0 65536 pairing
cap shared_q k_pair_miller_shared_q 0 65536 sharedq
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2y_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mgpu --no-wnaf-e2e > $O/r2y_ncu_bench.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2y_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for k, v in d['secondary'].items(): print('  ', k, {kk: vv for kk, vv in v.items() if kk not in ('config', 'cpu_baseline', 'e2e_c_abi', 'e2e')})
print(open('gpurun_out/r2y_bench_ref.json').read()[:400])
PY
