#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
export PROFILE_OUT_DIR=$O
python -m pytest tests -m gpu -x -q -k "miller or product or multi or smoke" > $O/r2o_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2o_pytest.log
tail -4 $O/r2o_pytest.log
for rep in 1 2 3; do for so in libpairing_b200 exp_mmnosmem; do
  echo "== $so run $rep" | tee -a $O/r2o_mm.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_latency.py --only-mm --mm-sweep 2>&1 | grep "trips\|2^20" | tee -a $O/r2o_mm.log
done; done
bash tools/bench_variants.sh pairing 2>&1 | tee $O/r2o_pair_variants.log
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_pair_multi_miller -c 1 -o /tmp/mms python tools/prof_pairing.py 1048576 mm > $O/r2o_ncu_mm.log 2>&1 && python tools/summarize_profiles.py - /tmp/mms.ncu-rep r2_multi_miller_smem >> $O/r2o_ncu_mm.log 2>&1
grep "time_duration\|dram\|fmaheavy_cycles_active.avg\|stall_\|shared" $O/r2_multi_miller_smem.md
