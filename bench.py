#!/usr/bin/env python
"""psac-b200 benchmark: suffixes/s of SA (+ISA) construction on synthetic random DNA.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--log2n 30] [--lcp]

One "step" = one complete suffix-array construction (the hot path of BASELINE.json) over one synthetic text.
N = 1 workload = BASELINE.json configs[1]: 1 GiB random DNA (|Sigma| = 4), SA-only, 32-bit index, one B200.
  value     : suffixes/s with the text already resident in HBM and the outputs left in HBM (psacb200_construct_device)
  e2e       : suffixes/s through the reference-facing C-ABI call with HOST (pinned) buffers -- H2D of the text and D2H
              of SA and ISA inside the timed region (psacb200_construct)
  roofline  : dominant kernel = one radix digit pass of the first sort over the carried keys; algorithmic bytes per
              launch = n * 2 * (4-byte carried key + 4-byte suffix index), see DESIGN.md
  cpu_baseline / --impl reference : the UNMODIFIED reference (oracle/_ref, MPI shim, np = 1 -> one core) on a bounded
              prefix of the same text
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 2  # C2 = 2^30 uniform ACGT, seed 2


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def reference_run(text_np, index_bytes, want_lcp, steps, warmup):
    """Times the unmodified reference (oracle/_ref) -- or the plain-C port when _ref is absent -- on host cores."""
    from oracle import pyoracle as O
    kind = "reference" if O.have_ref() else "port"
    fn = (lambda: O.ref_construct(text_np, index_bytes, want_lcp)) if kind == "reference" else (lambda: O.construct(text_np, index_bytes * 8, 0, want_lcp))
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return kind, text_np.size / dt, dt * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=30, help="log2 of the characters per GPU")
    ap.add_argument("--lcp", action="store_true", help="also build the LCP array")
    ap.add_argument("--cpu-log2n", type=int, default=25, help="log2 of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = 1 << args.log2n
    ngpu = max(world, args.gpus)
    if ngpu > 1:
        # ONE text of ngpu * 2^log2n characters, block-distributed over the ranks (BASELINE configs[2] shape: 64-bit index)
        index_bytes = 8
        workload = "random DNA (|Sigma|=4), ONE text of %d x 2^%d chars sharded by block over %d GPUs, SA%s (+ISA), 64-bit index (BASELINE configs[2] shape)" % (
            ngpu, args.log2n, ngpu, "+LCP" if args.lcp else "-only")
    else:
        index_bytes = 4
        workload = "random DNA (|Sigma|=4) 2^%d chars per GPU, SA%s, %d-bit index (BASELINE configs[1] shape)" % (
            args.log2n, "+LCP" if args.lcp else "-only (+ISA)", index_bytes * 8)
    from psac_b200 import textgen as G

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        m = 1 << min(args.cpu_log2n, args.log2n)
        text = G.random_dna(m, SEED)
        steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
        kind, sps, ms = reference_run(text, index_bytes, args.lcp, steps, warmup)
        line = {"impl": "reference", "metric": "suffixes/sec SA build", "value": sps, "unit": "suffixes/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u%d" % (index_bytes * 8),
                "data": "synthetic", "config": {"workload": workload, "cpu_sample": "first 2^%d characters of the same text per step" % int(np.log2(m)), "n_per_gpu": n, "n_total": n * ngpu},
                "cpu_baseline": {"value": sps, "unit": "suffixes/s", "cores": 1, "kind": kind,
                                 "sample": "first 2^%d characters of the text; unmodified psac at np=1 under the MPI shim (single core: the box has no MPI)" % int(np.log2(m))},
                "e2e": {"value": sps, "unit": "suffixes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from psac_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- psac-b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sharded = world > 1
    flags = (api.LCP if args.lcp else 0) | api.FAST_RESOLVAL
    n_total = n * world
    if sharded:
        # SA / ISA / LCP come back block-distributed.  Rank r generates its own block of the text (seed differs per rank).
        from psac_b200.sharded import ShardedSuffixArray
        ssa = ShardedSuffixArray(index_bytes, args.lcp)
        eng = ssa.engine
    else:
        eng = api.Engine(local_rank)
    text = G.random_dna_torch(n, SEED + rank, dev)
    tdt = torch.int32 if index_bytes == 4 else torch.int64  # raw storage for unsigned outputs
    d_sa = torch.empty(n, dtype=tdt, device=dev)
    d_isa = torch.empty(n, dtype=tdt, device=dev)
    d_lcp = torch.empty(n, dtype=tdt, device=dev) if args.lcp else None
    if not sharded:
        eng.reserve(n, index_bytes, flags)
    torch.cuda.synchronize()

    def step_device():
        if sharded:
            eng.construct_sharded_ptr(text.data_ptr(), n, n_total, index_bytes, flags, 0, d_sa.data_ptr(), d_isa.data_ptr(),
                                      d_lcp.data_ptr() if d_lcp is not None else None)
        else:
            eng.construct_ptr(text.data_ptr(), n, index_bytes, flags, 0, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr() if d_lcp is not None else None,
                              device=True)

    ext = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ms, scatter_ms, phase = [], [], {}
    ev0.record(ext)
    for _ in range(args.steps):
        step_device()
        s = eng.stats()
        pass_ms.append(s["ms_sort_pass_avg"])
        scatter_ms.append(s.get("ms_scatter_avg", 0.0))
        for k, v in s.items():
            if k.startswith("ms_"):
                phase[k] = phase.get(k, 0.0) + v / args.steps
    ev1.record(ext)
    barrier()
    ms_dev = ev0.elapsed_time(ev1) / args.steps
    launches = eng.launches - launches0
    clocks = sampler.summary()
    stats = eng.stats()

    # size-independent certificate of the last result, on the device
    if sharded:
        # SA and ISA are permutations of 0..n-1: their sums over all ranks must be n(n-1)/2 (mod 2^64, int64 wrap-around)
        sums = torch.stack([d_sa.sum(), d_isa.sum()])
        dist.all_reduce(sums)
        want = (n_total * (n_total - 1) // 2) % (1 << 64)
        ok = all((int(v.item()) % (1 << 64)) == want for v in sums)
        verified = "sum(SA) == sum(ISA) == n(n-1)/2 over all ranks (parity itself: tests/test_gpu_sharded.py)"
    else:
        ar = torch.arange(n, dtype=torch.int64, device=dev)
        ok = bool((d_isa.to(torch.int64).bitwise_and(0xFFFFFFFF)[d_sa.to(torch.int64).bitwise_and(0xFFFFFFFF)] == ar).all().item())
        del ar
        verified = "ISA[SA[i]]==i on device"
    if not ok:
        raise SystemExit("bench.py: result certificate failed -- result invalid")

    # ------------------------------------------------------------------ end to end: pinned host buffers through the public API
    e2e = None
    if not args.no_e2e:
        h_text = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_text.copy_(text)
        h_sa = torch.empty(n, dtype=tdt).pin_memory()
        h_isa = torch.empty(n, dtype=tdt).pin_memory()
        h_lcp = torch.empty(n, dtype=tdt).pin_memory() if args.lcp else None
        torch.cuda.synchronize()

        def step_host():
            if sharded:
                # the sharded C ABI takes device blocks: the host block goes up and the result blocks come down on the
                # engine's stream inside the timed region
                with torch.cuda.stream(ext):
                    text.copy_(h_text, non_blocking=True)
                step_device()
                with torch.cuda.stream(ext):
                    h_sa.copy_(d_sa, non_blocking=True)
                    h_isa.copy_(d_isa, non_blocking=True)
                    if h_lcp is not None:
                        h_lcp.copy_(d_lcp, non_blocking=True)
                ext.synchronize()
            else:
                eng.construct_ptr(h_text.data_ptr(), n, index_bytes, flags, 0, h_sa.data_ptr(), h_isa.data_ptr(), h_lcp.data_ptr() if h_lcp is not None else None)

        step_host()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_e2e = max(1, min(args.steps, 3))
        e0.record(ext)
        for _ in range(k_e2e):
            step_host()
        e1.record(ext)
        barrier()
        ms_e2e = e0.elapsed_time(e1) / k_e2e
        outs = 2 + (1 if args.lcp else 0)
        e2e = {"ms": ms_e2e, "h2d": n * world, "d2h": outs * n * index_bytes * world, "steps": k_e2e}

    # ------------------------------------------------------------------ reduce over ranks (max time)
    t = torch.tensor([ms_dev, e2e["ms"] if e2e else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # dominant kernel: one 8-bit digit pass over the carried keys (passes 2..P of the first sort); algorithmic bytes
    # per launch = read + write of one carried key and one suffix index (+ one auxiliary byte with 32-bit carried keys)
    # per suffix sorted on this GPU (DESIGN.md section "Kernels")
    if sharded:
        key_bytes, val_bytes, aux_bytes = 8, 8, 0
    else:
        key_bytes = 4 if stats["key_chars"] * stats["pack_bits"] - 8 <= 32 else 8  # 32-bit carried keys: the top digit is implied by the segment
        val_bytes, aux_bytes = stats["internal_index_bytes"], 0
    pass_bytes = float(n) * 2 * (key_bytes + val_bytes + aux_bytes)
    pass_avg_ms = float(np.mean(pass_ms))
    pass_achieved = pass_bytes / (pass_avg_ms * 1e-3) / 1e9 if pass_avg_ms > 0 else 0.0
    # The dominant kernel is the scatter kernel of a digit pass; the engine brackets each of its launches with CUDA events
    # on its stream (psacb200_stats.ms_scatter_avg).  Where that is not available (sharded / 64-bit keys) the whole pass
    # (histogram + scans + scatter) is reported as one unit.
    sc_ms = float(np.mean(scatter_ms)) if scatter_ms and min(scatter_ms) > 0 else 0.0
    if sc_ms > 0:
        kernel_name = "radix_scatter_seg_kernel<ArraySrc<u%d,u%d>,512,16> (scatter kernel of one 8-bit digit pass inside the segments; %d launches per step)" % (
            key_bytes * 8, val_bytes * 8, stats["sort_passes"] - 1)
        kern_ms, achieved = sc_ms, pass_bytes / (sc_ms * 1e-3) / 1e9
    else:
        kernel_name = "one 8-bit digit pass over the carried keys = tile histogram + scan kernels + radix_scatter kernel <ArraySrc<u%d,u%d>> (timed as one unit; rank 0)" % (
            key_bytes * 8, val_bytes * 8)
        kern_ms, achieved = pass_avg_ms, pass_achieved
    traffic = None  # DRAM bytes per launch from the committed ncu capture (same kernel, n = 2^30), scaled to this n
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
        if not sharded and key_bytes == 4:
            per = tj["dram_bytes_scatter_kernel"] if sc_ms > 0 else tj["dram_bytes_per_pass"]
            traffic = float(per) * n / float(tj["n"])
    except Exception:
        traffic = None
    line = {
        "metric": "suffixes/sec SA build", "value": n_total / (ms_dev * 1e-3), "unit": "suffixes/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u%d" % (index_bytes * 8), "data": "synthetic",
        "config": {"workload": workload, "n_per_gpu": n, "n_total": n_total,
                   "parallelism": ("block-sharded text/SA/ISA over %d GPUs, NCCL all-to-all-v" % world) if sharded else "single GPU", "seed": SEED,
                   "l2": "inputs_exceed_l2 (every pass streams >= 8 GiB)", "key_chars": stats["key_chars"], "sort_passes": stats["sort_passes"],
                   "rounds": stats["rounds"], "unresolved_after_first": stats["unresolved_after_first"], "verified": verified,
                   "exchange": ("peer stores over NVLink fused into the owner partition kernel" if stats.get("peer_exchange") else "NCCL all-to-all-v") if sharded else None},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": "profiles/roofline_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None,
                     "peak_source": peak_src, "bytes_per_launch": pass_bytes, "ms_per_launch": kern_ms,
                     "whole_pass": {"what": "histogram pre-pass + scans + scatter of one digit pass", "ms": pass_avg_ms, "achieved": pass_achieved,
                                    "frac": pass_achieved / peak}},
        "phases_ms": {k: round(v, 3) for k, v in sorted(phase.items())},
    }
    if e2e:
        line["e2e"] = {"value": n_total / (ms_e2e * 1e-3), "unit": "suffixes/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "ms_per_step": ms_e2e, "steps": e2e["steps"]}
    if not args.no_cpu_baseline and world == 1:
        m = 1 << min(args.cpu_log2n, args.log2n)
        kind, sps, ms = reference_run(G.random_dna(m, SEED), index_bytes, args.lcp, 1, 0)
        line["cpu_baseline"] = {"value": sps, "unit": "suffixes/s", "cores": 1, "kind": kind, "ms": ms,
                                "sample": "first 2^%d characters of the same text, one run; unmodified psac at np=1 under the MPI shim" % int(np.log2(m))}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
