#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "pairing or miller or engine or cpp or final" > $O/r2u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2u_pytest.log
tail -4 $O/r2u_pytest.log
for rep in 1 2; do timeout 300 bash tools/bench_variants.sh pairing 2>&1; done | tee $O/r2u_pair_variants.log
