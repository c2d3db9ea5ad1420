#!/bin/bash
# round-2 GPU call H (8 GPUs): the multi-device entry points on 2/4/8 devices, the 2^20-pair product through them, bench at N=8
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/r2h_gpus.log
python -m pytest tests -m gpu -x -q -k "multi_device or leave_the_callers or multi_miller_product_of_2_20 or cpp" --durations=5 > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2h_pytest.log
tail -12 $O/r2h_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2h_bench_n8.json 2> $O/r2h_bench_n8.err; echo "bench n8 rc=$?"; tail -3 $O/r2h_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-wnaf-e2e > $O/r2h_bench_n4.json 2> $O/r2h_bench_n4.err; echo "bench n4 rc=$?"; tail -3 $O/r2h_bench_n4.err
python - <<'PY'
import json
for f in ("gpurun_out/r2h_bench_n8.json", "gpurun_out/r2h_bench_n4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
        for k, v in d["secondary"].items():
            print("  ", k, {kk: vv for kk, vv in v.items() if kk not in ("config", "cpu_baseline")})
    except Exception as e:
        print(f, "unreadable", e)
PY
