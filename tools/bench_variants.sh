#!/bin/bash
# bench every experimental build pairing_b200/lib/exp_*.so (tuning helper): $1 = "pairing" (bench.py) or paths to skip in tools/bench_paths.py
mode=${1:-pairing}
for so in pairing_b200/lib/libpairing_b200.so pairing_b200/lib/exp_*.so; do
  echo "== $(basename $so)"
  if [ "$mode" = "pairing" ]; then
    PAIRING_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f pairings/s kernel %.2f ms frac %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
  else
    PAIRING_B200_LIB=$PWD/$so python tools/bench_paths.py --skip $mode $BENCH_PATHS_ARGS 2>&1 | grep -E "config|mismatch|Error"
  fi
done
