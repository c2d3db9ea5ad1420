#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_pair_fq12_pow -s 0 -c 1 -o /tmp/fq12_pow python tools/prof_pairing.py 16384 pow > $O/r3h_ncu_pow.log 2>&1 && PROFILE_OUT_DIR=$O python tools/summarize_profiles.py - /tmp/fq12_pow.ncu-rep r3h_fq12_pow >> $O/r3h_ncu_pow.log 2>&1
grep "time_duration\|fmaheavy_cycles_active.avg" $O/r3h_fq12_pow.md
