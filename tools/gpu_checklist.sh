#!/usr/bin/env bash
# One gpurun call's worth of standard evidence, so that a round does not pay the per-call overhead six times:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_checklist.sh r02'
# then here:  python tools/ncu_summary.py r02   (writes profiles/*_r02.md and profiles/keccak_traffic.json)
# Everything lands in gpurun_out/ (scratch); numbers printed under ncu are never bench values.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
step() { echo "== $* =="; }

step "gpu tests";            python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest_gpu_$TAG.log
step "smoke";                python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
step "bench N=1";            python bench.py > $OUT/bench_n1_$TAG.json 2> $OUT/bench_n1_$TAG.err; cut -c1-400 $OUT/bench_n1_$TAG.json
step "reference arm";        python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; cut -c1-300 $OUT/bench_ref_$TAG.json
step "kbench C2 / C3 / uniform"
python tools/kbench.py --which 2 --n 1000000 --iters 5 --variants staged,direct > $OUT/kbench_c2_$TAG.jsonl 2>&1
python tools/kbench.py --which 3 --n 2000000 --iters 5 --variants staged > $OUT/kbench_c3_$TAG.jsonl 2>&1
for sz in 532 112 32; do python tools/kbench.py --uniform $sz --n 4000000 --iters 5; done > $OUT/kbench_uniform_$TAG.jsonl 2>&1
tail -2 $OUT/kbench_c3_$TAG.jsonl | cut -c1-220
step "builders";             python tools/builders_bench.py > $OUT/builders_$TAG.jsonl 2> $OUT/builders_$TAG.err; cut -c1-160 $OUT/builders_$TAG.jsonl
step "ecrecover";           python tools/ecrecover_bench.py > $OUT/ecrecover_$TAG.json 2> $OUT/ecrecover_$TAG.err; cut -c1-300 $OUT/ecrecover_$TAG.json
step "ncu launch list (shares only)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_b_$TAG.log 2>&1
# bench.py --steps 1 --warmup 3 launches the 1M-proof Keccak kernel 4 times (3 warm-up + 1 timed) before the e2e leg's
# small chunk launches: skip 2 of them so the capture is a warm, FULL-SIZE launch (skipping 4 lands on a 12k-proof chunk)
step "ncu --set full: keccak"
ncu --set full --clock-control none --import-source on -k regex:keccak256_staged -s 2 -c 1 -f -o $OUT/prof_keccak_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_k_$TAG.log 2>&1
step "ncu --set full: walk"
ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -f -o $OUT/prof_walk_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_w_$TAG.log 2>&1
ls -la $OUT/*_$TAG.* | awk '{print $5, $9}'
