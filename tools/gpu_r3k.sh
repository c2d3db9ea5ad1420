#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_codec.py tests/test_cpp_host.py -m gpu -x -q > $O/r3k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3k_pytest.log
for so in libpairing_b200 exp_byorder; do echo "== $so"; PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mgpu --no-wnaf-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k in ('g1_decode_compressed_checked',): print(k, {kk: vv for kk, vv in d['secondary'][k].items() if kk not in ('config', 'cpu_baseline')})
"; done | tee $O/r3k_decode.log
