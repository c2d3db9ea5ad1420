#!/bin/bash
# round-2 GPU call A: parity after the device-guard / dedicated-squaring changes, A/B of the experimental builds,
# multi-Miller timing modes, ncu captures.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; O=gpurun_out
export PROFILE_OUT_DIR=$O
python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2a_pytest.log
python - > $O/r2a_peaks.log 2>&1 <<'PY'
import sys; sys.path[:0]=['.']
import pairing_b200._native as nat
ctx = nat.Context(0)
for v, name in ((0, "IMAD.WIDE chain"), (1, "fp_mul chain (300)"), (4, "fp_sqr chain (234)"), (2, "IMAD 32")):
    macs, ms = ctx.imad_peak(v, 4000)
    print("peak[%d] %-24s %8.3f T MAC/s  (%.2f ms)" % (v, name, macs / 1e12, ms))
PY
bash tools/bench_variants.sh pairing > $O/r2a_pair_variants.log 2>&1
for so in libpairing_b200 exp_ksqr0; do
  echo "== $so" >> $O/r2a_paths.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_paths.py --log2 22 --skip pairing,mm 2>&1 | grep -E "config|norm|mismatch|Error|exact" >> $O/r2a_paths.log
done
for rep in 1 2 3; do for so in libpairing_b200 exp_mmstream; do
  echo "== $so run $rep" >> $O/r2a_mm.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_paths.py --log2 20 --skip pairing,g1,g2 2>&1 | grep -E "config|mismatch|Error|checksum" >> $O/r2a_mm.log
done; done
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_pair_multi_miller -c 1 -o /tmp/mm python tools/prof_pairing.py 1048576 mm > $O/r2a_ncu_mm.log 2>&1 && python tools/summarize_profiles.py - /tmp/mm.ncu-rep r2_multi_miller >> $O/r2a_ncu_mm.log 2>&1
$NCU -k regex:k_wnaf_mul_lazyk -s 2 -c 1 -o /tmp/g1 python tools/prof_pairing.py 1048576 wnaf > $O/r2a_ncu_g1.log 2>&1 && python tools/summarize_profiles.py - /tmp/g1.ncu-rep r2_g1_wnaf >> $O/r2a_ncu_g1.log 2>&1
PAIRING_B200_LIB=$PWD/pairing_b200/lib/exp_smem2.so $NCU -k regex:k_pair_miller -c 1 -o /tmp/ps python tools/prof_pairing.py 65536 pairing > $O/r2a_ncu_smem2.log 2>&1 && python tools/summarize_profiles.py - /tmp/ps.ncu-rep r2_pair_miller_smem2 >> $O/r2a_ncu_smem2.log 2>&1
ls -la $O | tail -20
