#!/bin/bash
# bench every experimental build pairing_b200/lib/exp_*.so (tuning helper)
for so in pairing_b200/lib/exp_*.so; do
  echo -n "$(basename $so): "
  PAIRING_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f pairings/s kernel %.2f ms frac %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
