#!/usr/bin/env bash
# round-2 GPU call A (1 GPU): the whole -m gpu suite, smoke, the full bench line, sender-recovery throughput
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== gpu tests =="; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -40 | tee $OUT/pytest_gpu_r02a.log
echo "== smoke ==";     timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_r02a.log
echo "== bench ==";     timeout 900 python bench.py > $OUT/bench_n1_r02a.json 2> $OUT/bench_n1_r02a.err; tail -5 $OUT/bench_n1_r02a.err; cut -c1-1500 $OUT/bench_n1_r02a.json
echo "== ecrecover =="; timeout 300 python tools/ecrecover_bench.py > $OUT/ecrecover_r02a.json 2> $OUT/ecrecover_r02a.err; tail -3 $OUT/ecrecover_r02a.err; cat $OUT/ecrecover_r02a.json
