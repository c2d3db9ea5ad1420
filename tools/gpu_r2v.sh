#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for rep in 1 2; do timeout 300 bash tools/bench_variants.sh pairing 2>&1; done | tee $O/r2v_pair_variants.log
for so in libpairing_b200 exp_onetu; do echo "== $so"; PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so timeout 200 python tools/bench_paths.py --skip mm,g1,g2 --log2 16 2>&1 | grep "batch\|config" ; done | tee $O/r2v_paths.log
