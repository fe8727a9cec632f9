#!/usr/bin/env bash
# round-2 GPU call B (1 GPU): the whole -m gpu suite (all failures, not the first), smoke, the full bench line, sender
# recovery, walk residency sweep on C3, compute-sanitizer memcheck of the small-batch workload
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== gpu tests =="; timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -60 | tee $OUT/pytest_gpu_r02b.log
echo "== smoke ==";     timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_r02b.log
echo "== bench ==";     timeout 700 python bench.py > $OUT/bench_n1_r02b.json 2> $OUT/bench_n1_r02b.err; tail -5 $OUT/bench_n1_r02b.err; cut -c1-3000 $OUT/bench_n1_r02b.json
echo "== ecrecover =="; timeout 300 python tools/ecrecover_bench.py > $OUT/ecrecover_r02b.json 2> $OUT/ecrecover_r02b.err; tail -3 $OUT/ecrecover_r02b.err; cat $OUT/ecrecover_r02b.json
echo "== walk residency on C3 =="
for m in 8 6; do PHANT_WALK_MINB=$m timeout 300 python tools/kbench.py --which 3 --n 2000000 --iters 5 --variants staged 2>&1 | tail -1 | sed "s/^/minb=$m /"; done | tee $OUT/walk_minb_r02b.log
echo "== sanitizer =="; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > $OUT/sanitize_memcheck_r02.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/sanitize_memcheck_r02.log; tail -5 $OUT/sanitize_memcheck_r02.log
