#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2i_pytest.log
tail -10 $O/r2i_pytest.log
python tools/bench_latency.py --mm-log2 17 2>&1 | tee $O/r2i_latency.log | head -20
python bench.py --steps 5 --warmup 3 > $O/r2i_bench_n1.json 2> $O/r2i_bench_n1.err; echo "bench rc=$?"; tail -3 $O/r2i_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2i_bench_ref.json 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2i_smoke.log 2>&1; cat $O/r2i_smoke.log
