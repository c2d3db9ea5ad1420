#!/bin/bash
# round 2, final build of the session: full parity suite, N=1 bench line + reference arm, latency table, ncu captures of the
# kernels that changed (fused pairing, final exponentiation, warp-cooperative pairing) and the launch list of bench.py
mkdir -p gpurun_out; O=gpurun_out
export PROFILE_OUT_DIR=$O
timeout 400 python -m pytest tests -m gpu -x -q --durations=4 > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2w_pytest.log
tail -9 $O/r2w_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2w_bench_n1.json 2> $O/r2w_bench_n1.err; echo "bench rc=$?"; tail -3 $O/r2w_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2w_bench_ref.json 2> $O/r2w_bench_ref.err; echo "ref rc=$?"
timeout 200 bash tools/bench_variants.sh pairing 2>&1 | tee $O/r2w_pair_variants.log
timeout 200 python tools/bench_latency.py > $O/r2w_latency.log 2>&1; head -14 $O/r2w_latency.log
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # tag kernel-regex skip n what
  timeout 300 $NCU -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/prof_pairing.py $4 $5 > $O/r2w_ncu_$1.log 2>&1 && python tools/summarize_profiles.py - /tmp/$1.ncu-rep r2w_$1 >> $O/r2w_ncu_$1.log 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap pair_miller 'k_pair_miller$' 0 65536 pairing
cap final_exp k_pair_final_exp 0 65536 finalexp
cap wide_pairing k_wide_pairing 0 1000 pairing_wide
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2w_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mgpu --no-wnaf-e2e > $O/r2w_ncu_bench.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2w_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])
for k, v in d['secondary'].items(): print('  ', k, {kk: vv for kk, vv in v.items() if kk not in ('config', 'cpu_baseline', 'e2e_c_abi', 'e2e')})
PY
ls -la $O | grep r2w
