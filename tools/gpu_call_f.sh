#!/usr/bin/env bash
# round-2 GPU call F (1 GPU, short): sender recovery with the inlined hot loop, walk with one offset load per node
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== tests =="; timeout 900 python -m pytest tests/test_gpu_ecrecover.py tests/test_gpu_verify.py tests/test_fuzz_walk.py tests/test_gpu_host_py.py -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_r02f.log
echo "== ecrecover =="; timeout 300 python tools/ecrecover_bench.py | tee $OUT/ecrecover_r02f.json
echo "== walk =="
for m in 8 6; do PHANT_WALK_MINB=$m timeout 300 python tools/kbench.py --which 3 --n 2000000 --iters 5 --variants staged 2>&1 | tail -1 | sed "s/^/minb=$m /"; done | tee $OUT/walk_minb_r02f.log
PHANT_WALK_MINB=8 timeout 300 python tools/kbench.py --which 2 --n 1000000 --iters 5 --variants staged 2>&1 | tail -1 | tee -a $OUT/walk_minb_r02f.log
