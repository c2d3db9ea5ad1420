#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "division_steps" > $O/r3i_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r3i_pytest.log
for rep in 1 2; do timeout 300 bash tools/bench_variants.sh pairing 2>&1; done | tee $O/r3i_pair_variants.log
