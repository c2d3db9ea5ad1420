#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for rep in 1 2; do for so in libpairing_b200 exp_fp2dual; do echo "== $so"; PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so timeout 200 python tools/bench_paths.py --skip pairing,mm,g1 --log2 20 2>&1 | grep "config\|mismatch\|Error\|exact"; done; done | tee $O/r3g_g2.log
PAIRING_B200_LIB=$PWD/pairing_b200/lib/exp_fp2dual.so timeout 300 python -m pytest tests -m gpu -x -q -k "g2 or True or codec or point_ops or wnaf or fq2 or fq6 or fq12" > $O/r3g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r3g_pytest.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_pair_fq12_pow -s 0 -c 1 -o /tmp/fq12_pow python tools/prof_pairing.py 16384 pow > $O/r3g_ncu_pow.log 2>&1 && PROFILE_OUT_DIR=$O python tools/summarize_profiles.py - /tmp/fq12_pow.ncu-rep r3g_fq12_pow >> $O/r3g_ncu_pow.log 2>&1
