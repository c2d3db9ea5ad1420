#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "pairing or miller or wnaf or mul or batch or cpp or stated" > $O/r3a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r3a_pytest.log
tail -3 $O/r3a_pytest.log
for rep in 1 2; do for so in libpairing_b200 exp_pipe1; do echo "== $so"; PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-mgpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('device', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'match', d['e2e']['matches_device_path'])
g = d['secondary']['g1_wnaf_mul']; print('g1 wnaf', g['value'], 'e2e', g.get('e2e'))
"; done; done | tee $O/r3a_e2e.log
