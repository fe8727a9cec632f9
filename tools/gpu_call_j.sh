#!/usr/bin/env bash
# round-2 GPU call J (1 GPU): what the driver runs at round end -- the whole -m gpu suite, smoke, both bench arms
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== gpu tests =="; timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -15 | tee $OUT/pytest_gpu_r02j.log
echo "== smoke ==";     timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_r02j.log
echo "== bench ==";     timeout 900 python bench.py > $OUT/bench_n1_r02j.json 2> $OUT/bench_n1_r02j.err; tail -3 $OUT/bench_n1_r02j.err; cut -c1-700 $OUT/bench_n1_r02j.json
echo "== reference arm =="; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_r02j.json 2> $OUT/bench_ref_r02j.err; cut -c1-500 $OUT/bench_ref_r02j.json
echo "== e2e vs chunk size =="
for mb in 48 128 256; do PHANT_GPU_CHUNK_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --skip-extras --no-cpu 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('chunk_mb=$mb', d['e2e'])"; done | tee $OUT/e2e_chunk_r02j.log
