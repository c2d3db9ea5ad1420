#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "wnaf or curve or full_size or stated or cpp" > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2t_pytest.log
tail -5 $O/r2t_pytest.log
for so in libpairing_b200 exp_wv1 exp_wnosmem; do echo "== $so"; PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so timeout 200 python tools/bench_paths.py --skip pairing,mm --log2 20 2>&1 | grep "config\|mismatch\|Error" ; done | tee $O/r2t_paths.log
