#!/bin/bash
# round-2 GPU call C (2 GPUs): full GPU suite incl. stated sizes and the 2-device entry points, bench at N=1 and N=2 (torchrun)
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/r2c_gpus.log
python -m pytest tests -m gpu -x -q --durations=8 > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2c_pytest.log
tail -15 $O/r2c_pytest.log
python bench.py --steps 5 --warmup 3 > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; echo "bench n1 rc=$?"; tail -3 $O/r2c_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c_bench_n2.json 2> $O/r2c_bench_n2.err; echo "bench n2 rc=$?"; tail -3 $O/r2c_bench_n2.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c_bench_ref.json 2>&1
for so in exp_g1b3 exp_g1b4 exp_g1b4k2 exp_g1b6; do
  echo "== $so" >> $O/r2c_paths.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_paths.py --log2 22 --skip pairing,mm,g2 2>&1 | grep -E "config|norm|mismatch|Error|exact" >> $O/r2c_paths.log
done
cat $O/r2c_paths.log
python - <<'PY'
import json
for f in ("gpurun_out/r2c_bench_n1.json", "gpurun_out/r2c_bench_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
        for k, v in d["secondary"].items():
            print("  ", k, {kk: vv for kk, vv in v.items() if kk not in ("config", "cpu_baseline")})
            if "cpu_baseline" in v: print("     cpu:", v["cpu_baseline"])
    except Exception as e:
        print(f, "unreadable", e)
PY
