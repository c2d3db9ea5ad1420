#!/bin/bash
# round-2 GPU call D: parity of the reworked warp-cooperative engine + pool + pipelines, latency sweep, ncu captures
mkdir -p gpurun_out; O=gpurun_out
export PROFILE_OUT_DIR=$O
python -m pytest tests -m gpu -x -q --durations=5 > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2d_pytest.log
tail -12 $O/r2d_pytest.log
python tools/bench_latency.py > $O/r2d_latency.log 2>&1; tail -28 $O/r2d_latency.log
python bench.py --steps 5 --warmup 3 > $O/r2d_bench_n1.json 2> $O/r2d_bench_n1.err; echo "bench rc=$?"; tail -3 $O/r2d_bench_n1.err
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # tag kernel-regex skip n what
  $NCU -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/prof_pairing.py $4 $5 > $O/r2d_ncu_$1.log 2>&1 && python tools/summarize_profiles.py - /tmp/$1.ncu-rep r2_$1 >> $O/r2d_ncu_$1.log 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap pair_miller 'k_pair_miller$' 0 65536 pairing
cap wide_pairing k_wide_pairing 0 1000 pairing_wide
cap multi_miller_2_17 k_pair_multi_miller 0 131072 mm
cap product_tail k_pair_product_tail 0 131072 product
cap shared_q k_pair_miller_shared_q 0 65536 sharedq
cap fixed_base k_wnaf_fixed_base 0 1048576 fixed
cap fq12_pow k_pair_fq12_pow 0 16384 pow
cap final_exp k_pair_final_exp 0 65536 finalexp
cap g2_wnaf k_wnaf_mul_lazyk 2 262144 g2wnaf
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mgpu --no-wnaf-e2e > $O/r2d_ncu_bench.log 2>&1
ls -la $O | grep r2d
