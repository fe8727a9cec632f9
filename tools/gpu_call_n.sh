#!/usr/bin/env bash
# round-2 GPU call N (1 GPU): last check of bench.py / smoke() as committed
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench =="; timeout 700 python bench.py > $OUT/bench_n1_r02n.json 2> $OUT/bench_n1_r02n.err; tail -3 $OUT/bench_n1_r02n.err | cut -c1-300
grep '^{' $OUT/bench_n1_r02n.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','parity','kernel_ms')}); print(d['c5']['witness_build_s_on_device'], d['c5']['torch_first_use_s'], d['c5']['parity'], d['c5']['ms_per_batch'], d['c3']['ms_per_step'], d['c4_sparse']['ms_per_update'])"
echo "== quick tests =="; timeout 600 python -m pytest tests/test_gpu_keccak.py tests/test_gpu_verify.py -m gpu -q 2>&1 | tail -3
