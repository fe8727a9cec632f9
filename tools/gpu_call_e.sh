#!/usr/bin/env bash
# round-2 GPU call E (1 GPU): walk residency with the new load pattern, the bench line with the sparse resident trie,
# ncu evidence of what ships (launch list, --set full of the Keccak / walk / frontier / encode / recovery kernels), racecheck
set -u
OUT=gpurun_out
TAG=r02
mkdir -p $OUT
echo "== walk residency on C3 (new loads) =="
for m in 8 6 10; do PHANT_WALK_MINB=$m timeout 300 python tools/kbench.py --which 3 --n 2000000 --iters 5 --variants staged 2>&1 | tail -1 | sed "s/^/minb=$m /"; done | tee $OUT/walk_minb_r02e.log
for m in 8 6; do PHANT_WALK_MINB=$m timeout 300 python tools/kbench.py --which 2 --n 1000000 --iters 5 --variants staged 2>&1 | tail -1 | sed "s/^/c2 minb=$m /"; done | tee -a $OUT/walk_minb_r02e.log
echo "== gpu tests (walk + verify + trie) =="; timeout 900 python -m pytest tests/test_gpu_verify.py tests/test_fuzz_walk.py tests/test_gpu_host_py.py tests/test_gpu_ecrecover.py tests/test_gpu_host_cpp.py -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_walk_r02e.log
echo "== bench ==";     timeout 900 python bench.py > $OUT/bench_n1_r02e.json 2> $OUT/bench_n1_r02e.err; tail -3 $OUT/bench_n1_r02e.err; python -c "
import json;d=json.load(open('$OUT/bench_n1_r02e.json'));print({k:d[k] for k in ('value','ms_per_step','kernel_ms')});print(d.get('c3',{}).get('kernel_ms_mean_per_rank'), d.get('c3',{}).get('walk_share'));print(json.dumps(d.get('c4_sparse'))[:900])"
echo "== ecrecover =="; timeout 300 python tools/ecrecover_bench.py > $OUT/ecrecover_r02e.json 2> $OUT/ecrecover_r02e.err; tail -3 $OUT/ecrecover_r02e.err; cat $OUT/ecrecover_r02e.json
echo "== ncu launch list =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --skip-extras > $OUT/ncu_b_$TAG.log 2>&1
echo "== ncu --set full =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:keccak256_staged -s 2 -c 1 -f -o $OUT/prof_keccak_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu --skip-extras > $OUT/ncu_k_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -f -o $OUT/prof_walk_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu --skip-extras > $OUT/ncu_w_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -f -o $OUT/prof_walkc3_$TAG \
    python tools/kbench.py --which 3 --n 2000000 --iters 3 --variants staged > $OUT/ncu_wc3_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:keccak256_staged -s 3 -c 1 -f -o $OUT/prof_keccakc3_$TAG \
    python tools/kbench.py --which 3 --n 2000000 --iters 3 --variants staged > $OUT/ncu_kc3_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"frontier_branch|branch_encode|ecrecover_kernel|st_top_branch|keccak_regroup" -c 12 -f -o $OUT/prof_builders_$TAG \
    python tools/sanitize.py > $OUT/ncu_builders_$TAG.log 2>&1
echo "== sparse trie update: where the time goes =="
timeout 300 python tools/strie_bench.py --keys 4000000 --dirty 100000 --steps 5 | tee $OUT/strie_r02e.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_strie_$TAG.csv \
    python tools/strie_bench.py --keys 1000000 --dirty 100000 --steps 1 > $OUT/ncu_strie_$TAG.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_strie_r02.csv")) if len(r)>10]
h=rows[0]; ki,vi,ui=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Unit")
# the last update only: launches after the last-but-one st_classify_kernel
idx=[i for i,r in enumerate(rows) if "st_classify" in r[ki]]
last=rows[idx[-1]:] if idx else rows[1:]
agg=collections.Counter(); cnt=collections.Counter()
for r in last:
    v=float(r[vi].replace(",","")); v*= {"ns":1e-6,"us":1e-3,"ms":1.0}.get(r[ui],1e-6)
    n=r[ki].split("(")[0].replace("void ","")[:60]; agg[n]+=v; cnt[n]+=1
print("last update: %d launches, %.3f ms of kernels" % (sum(cnt.values()), sum(agg.values())))
for n,v in agg.most_common(14): print("  %-62s %3d  %.3f ms" % (n,cnt[n],v))
PY
echo "== racecheck =="; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize.py > $OUT/sanitize_racecheck_r02.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/sanitize_racecheck_r02.log; tail -4 $OUT/sanitize_racecheck_r02.log
ls -la $OUT/*_$TAG.* | awk '{print $5, $9}'
