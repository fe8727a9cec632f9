#!/usr/bin/env python3
"""bench.py -- the contract benchmark (one JSON line on rank 0).

Workload (BASELINE.json configs[1]): synthetic account proofs, depth 8, 532-byte branch nodes,
1,000,000 proofs PER GPU (weak scaling: rank r verifies proofs [r*1M, (r+1)*1M) of the same PRNG
stream).  A "step" is one pass of the hot path over that batch: hash all 8M nodes (batched Keccak
kernel), walk all proofs, and -- at N > 1 -- one NCCL all-reduce that assembles the global accept
bitmap on every rank.

  value        proofs/s, whole job, witnesses already resident in HBM (device-pointer ABI), CUDA events,
               max over ranks
  e2e          the same metric through the host-pointer C ABI call a phant maintainer binds
               (phant_gpu_verify_proofs): pinned host witness -> H2D -> hash -> walk -> verdicts D2H,
               every step
  roofline     dominant kernel = batched Keccak; algorithmic bytes = 3,900 B/proof
  cpu_baseline the CPU path on this box's host cores (oracle walk over the reference's compiled
               keccak.c when oracle/_ref is present), bounded sample
  --impl reference   the CPU arm alone, same metric / config
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PROOFS_PER_GPU = 1_000_000
DEPTH = 8
ALGO_BYTES_PER_PROOF = 3900  # 7*532 + 112 node bytes + 32 key + 32 root
PERMS_PER_PROOF = 29
METRIC = "mpt_proofs_verified_per_sec"
UNIT = "proofs/s"


def config(n_gpus):
    return {"workload": "synthetic account proofs, depth 8 (7 x 532-byte full branch + 112-byte leaf), "
                        f"{PROOFS_PER_GPU} proofs per GPU, 1 in 97 corrupted", "proofs_per_gpu": PROOFS_PER_GPU,
            "global_batch": PROOFS_PER_GPU * n_gpus, "depth": DEPTH, "bytes_per_proof": ALGO_BYTES_PER_PROOF,
            "keccak_f_per_proof": PERMS_PER_PROOF, "parallelism": f"proof-shard x{n_gpus}",
            "l2": "inputs (3.9 GB per GPU) larger than L2; no flush needed"}


# ----------------------------------------------------------------------------------------------
# CPU arm
# ----------------------------------------------------------------------------------------------
def host_threads():
    """threads the CPU arm can really use: affinity mask capped by the cgroup CPU quota (cpu.max)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return n


def cpu_arm(sample_proofs, target_cpu_seconds, threads):
    """Verify a bounded sample on the host cores; returns dict(value, cores, kind, sample)."""
    import numpy as np
    import oracle_lib
    o = oracle_lib.get()
    kind = "reference" if o.use_reference_keccak(True) else "port"
    nodes, node_off, first, keys, roots = o.synth_c2(sample_proofs, depth=DEPTH, threads=threads)
    # calibrate with one pass, then repeat to reach the target amount of CPU work
    bitmap, status, _, _ = o.verify_proofs(nodes, node_off, first, keys, roots, threads=threads)  # warm-up + check
    t0 = time.perf_counter()
    o.verify_proofs(nodes, node_off, first, keys, roots, threads=threads)
    t1 = time.perf_counter() - t0
    expect = np.where(np.arange(sample_proofs) % 97 == 0, 0, 1)
    assert (status == expect).all(), "CPU arm verdicts wrong"
    reps = max(1, min(200, int(target_cpu_seconds / max(t1 * threads, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps):
        o.verify_proofs(nodes, node_off, first, keys, roots, threads=threads)
    dt = time.perf_counter() - t0
    value = sample_proofs * reps / dt
    # phant's own path is single-threaded (SURVEY.md 8d asks for both figures): the same walk on ONE core, ~2 s of work
    one = min(sample_proofs, 16384)
    first1 = first[:one + 1]
    t0 = time.perf_counter()
    o.verify_proofs(nodes, node_off, first1, keys[:32 * one], roots[:32 * one], threads=1)
    t1core = time.perf_counter() - t0
    reps1 = max(1, min(50, int(2.0 / max(t1core, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps1):
        o.verify_proofs(nodes, node_off, first1, keys[:32 * one], roots[:32 * one], threads=1)
    one_core = one * reps1 / (time.perf_counter() - t0)
    o.use_reference_keccak(False)
    what = ("oracle proof walk (phant has no verifier) over the reference's own keccak.c compiled unchanged (oracle/_ref)"
            if kind == "reference" else "oracle C port (oracle/_ref absent)")
    return {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "one_core": one_core,
            "sample": f"{sample_proofs} proofs of the same workload x {reps} passes, {threads} threads, {dt:.2f} s wall; {what}"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    sample = int(os.environ.get("PHANT_BENCH_CPU_SAMPLE", "131072"))
    steps = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_arm(sample, target_cpu_seconds=max(2.0, float(os.environ.get("PHANT_BENCH_CPU_SECONDS", "20")) / max(1, args.steps)),
                       threads=threads)
        if i >= args.warmup:
            steps.append(base["value"])
    value = statistics.mean(steps)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config(args.gpus), "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # only the samples under load say anything about the timed region
        busy = [s for s, p in zip(sm, power) if p > 250] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(gpu_index):
    """Restrict this process to the CPUs NVML reports as local to the GPU (same socket as its PCIe root), so the
    pinned host buffers of the e2e leg are allocated on that socket's memory.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)[0], len(cpus)
    except Exception:  # noqa: BLE001 -- no NVML / no permission: keep the default placement
        pass
    return None


def run_gpu(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from phant_b200 import gpu

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_to_gpu_numa_node(local_rank)  # pinned staging buffers are then first-touched next to this GPU's PCIe root
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = gpu.Context(local_rank)
    from phant_b200 import shard
    n = PROOFS_PER_GPU
    first_index, hi = shard.shard_range(world * n, rank, world)  # contiguous, 64-aligned proof ranges (weak scaling)
    assert hi - first_index == n

    # ---- witnesses generated in HBM (setup, untimed) ----
    n_nodes, n_bytes = ctx.synth_sizes(2, n, depth=DEPTH, first=first_index)
    d_nodes = torch.empty(n_bytes + 64, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n_nodes + 1, dtype=torch.int64, device=dev)
    d_first = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_keys = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_roots = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    ctx.synth(2, n, d_nodes, d_off, d_first, d_keys, d_roots, depth=DEPTH, first=first_index)
    words = (n + 63) // 64
    g_bitmap = torch.zeros(world * words, dtype=torch.int64, device=dev)  # global accept bitmap, my slice is mine
    my_bitmap = g_bitmap[rank * words:(rank + 1) * words]
    d_status = torch.empty(n, dtype=torch.uint8, device=dev)

    # one side stream shared by the library's kernels, torch's ops and NCCL (torch's default stream has handle 0,
    # which phant_gpu_set_stream reads as "restore the private stream")
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    ctx.set_stream(side.cuda_stream)

    def step_device():
        if world > 1:
            g_bitmap.zero_()
        ctx.verify_proofs(n, d_nodes, d_off, d_first, d_keys, d_roots, n, my_bitmap, d_status, None, None,
                          n_nodes=n_nodes, nodes_bytes=n_bytes)
        if world > 1:
            dist.all_reduce(g_bitmap, op=dist.ReduceOp.SUM)  # one collective per step; disjoint words: SUM == OR

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident value ----
    ctx.set_flags(gpu.FLAG_DEVICE_PTRS)
    for _ in range(args.warmup):
        step_device()
    ctx.synchronize()
    barrier()
    ctx.reset_stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    torch.cuda.synchronize()
    dt_local = ev0.elapsed_time(ev1) * 1e-3  # device time of exactly K steps on the launching stream
    st = ctx.stats()  # per-kernel device time (CUDA events inside the library, same stream)
    barrier()
    t_step = torch.tensor([dt_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
    dt = float(t_step.item())
    value = world * n * args.steps / dt

    # verdict check (outside the timed region): reject iff global index % 97 == 0
    expect = np.where((np.arange(n) + first_index) % 97 == 0, 0, 1)
    status_ok = bool((d_status.cpu().numpy() == expect).all())
    bits = np.unpackbits(g_bitmap.cpu().numpy().view(np.uint8), bitorder="little")
    if world > 1:
        allexp = np.concatenate([np.pad(np.where((np.arange(n) + r * n) % 97 == 0, 0, 1), (0, words * 64 - n)) for r in range(world)])
    else:
        allexp = np.pad(expect, (0, words * 64 - n))
    bitmap_ok = bool((bits == allexp).all())

    # ---- end to end through the host-pointer ABI ----
    h_nodes = torch.empty(n_bytes + 64, dtype=torch.uint8, pin_memory=True)
    h_nodes.copy_(d_nodes)
    h_off = d_off.cpu().pin_memory()
    h_first = d_first.cpu().pin_memory()
    h_keys = d_keys.cpu().pin_memory()
    h_roots = d_roots.cpu().pin_memory()
    h_bitmap = torch.zeros(words, dtype=torch.int64).pin_memory()
    h_status = torch.zeros(n, dtype=torch.uint8).pin_memory()
    ctx.set_flags(0)

    def step_e2e():
        ctx.verify_proofs(n, h_nodes, h_off, h_first, h_keys, h_roots, n, h_bitmap, h_status, None, None)
        if world > 1:
            my_bitmap.copy_(h_bitmap, non_blocking=True)
            dist.all_reduce(g_bitmap, op=dist.ReduceOp.SUM)

    e2e_steps = max(3, min(args.steps, 10))
    step_e2e()
    barrier()
    ctx.reset_stats()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    dt_e = time.perf_counter() - t0
    st_e = ctx.stats()
    t_e = torch.tensor([dt_e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t_e.item())
    e2e_ok = bool((h_status.numpy() == expect).all())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        keccak_ms = st["keccak_ms"] / args.steps
        walk_ms = st["walk_ms"] / args.steps
        achieved = ALGO_BYTES_PER_PROOF * n / (keccak_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "keccak_traffic.json")))["dram_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            pass
        sm_mhz = clocks.get("sm_mhz") or 1900.0
        # integer-issue ceiling of the permutation: 24 rounds x 180 ALU-pipe instructions (SASS count), 64 lanes/clk/SM;
        # the last permutation of a message needs only the digest, which prunes its 24th round from 180 to 58
        perms_per_step = st["keccak_perms"] / args.steps
        instr_per_perm = 24 * 180 - 122 * n_nodes / perms_per_step
        alu_peak_perm_s = 148 * 64 * sm_mhz * 1e6 / instr_per_perm
        perm_s = perms_per_step / (keccak_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": config(world),
            "keccak_mh_s": world * n_nodes / (keccak_ms * 1e-3) / 1e6, "keccak_gperm_s": world * perm_s / 1e9,
            "kernel_ms": {"keccak": keccak_ms, "walk": walk_ms},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "keccak256_staged_kernel",
                         "note": "Keccak-f is integer-issue bound, not HBM bound: see alu",
                         "alu": {"achieved_gperm_s": perm_s / 1e9, "peak_gperm_s": alu_peak_perm_s / 1e9,
                                 "frac": perm_s / alu_peak_perm_s,
                                 "model": "148 SM x 64 INT lanes/clk x sm_mhz / (24 rounds x 180 ALU instr - 122 per message: digest-only last round)"}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st_e["h2d_bytes"] // e2e_steps,
                    "d2h_bytes_per_step": st_e["d2h_bytes"] // e2e_steps, "steps": e2e_steps,
                    "ms_per_step": 1e3 * float(t_e.item()) / e2e_steps},
            "gpu_launches": int(st["launches"]), "clocks": clocks, "host_numa_pin": numa,
            "parity": {"status_ok": status_ok, "bitmap_ok": bitmap_ok, "e2e_ok": e2e_ok},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_arm(131072, target_cpu_seconds=20.0, threads=host_threads())
        print(json.dumps(line), flush=True)
        if not (status_ok and bitmap_ok and e2e_ok):
            sys.exit("verdicts differ from the expected pattern")
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="phant_b200", choices=["phant_b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        sys.exit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N")
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
