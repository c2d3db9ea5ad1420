#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "miller or product or multi" > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2f_pytest.log
tail -4 $O/r2f_pytest.log
for so in libpairing_b200 exp_mmnosplit exp_mmchunk16 exp_mmchunk4; do
  echo "== $so" | tee -a $O/r2f_mm.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_latency.py --only-mm 2>&1 | tee -a $O/r2f_mm.log | tail -4
done
