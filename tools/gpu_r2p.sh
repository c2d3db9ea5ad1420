#!/bin/bash
# round 2, step p: division-step inversion + compressed exp_by_x -- parity suite, then A/B of the three builds
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -x -q --durations=4 > $O/r2p_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2p_pytest.log
tail -9 $O/r2p_pytest.log
bash tools/bench_variants.sh pairing 2>&1 | tee $O/r2p_pair_variants.log
BENCH_PATHS_ARGS="--log2 16" bash tools/bench_variants.sh mm,g1,g2 2>&1 | tee $O/r2p_paths.log
