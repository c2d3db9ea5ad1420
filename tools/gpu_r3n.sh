#!/bin/bash
# bench.py at N GPUs (torchrun), final build of round 2: $1 = N
N=$1; mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L | wc -l > $O/r3n_gpus_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-wnaf-e2e > $O/r3n_bench_n$N.json 2> $O/r3n_bench_n$N.err; echo "bench n$N rc=$?"; tail -3 $O/r3n_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r3n_bench_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
for k, v in d["secondary"].items():
    print("  ", k, {kk: vv for kk, vv in v.items() if kk not in ("config", "cpu_baseline")})
PY
