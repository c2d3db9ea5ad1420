#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --durations=4 > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2s_pytest.log
tail -9 $O/r2s_pytest.log
timeout 300 bash tools/bench_variants.sh pairing 2>&1 | tee $O/r2s_pair_variants.log
for so in libpairing_b200 exp_base; do echo "== $so"; PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so timeout 200 python tools/bench_paths.py --skip mm,g1,g2 --log2 16 2>&1 | grep "batch\|config" ; PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so timeout 100 python tools/bench_latency.py 2>&1 | tail -12; done | tee $O/r2s_paths.log
