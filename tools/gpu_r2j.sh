#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2j_pytest.log
tail -4 $O/r2j_pytest.log
python tools/bench_latency.py --mm-log2 17 2>&1 | tee $O/r2j_latency.log | head -20
