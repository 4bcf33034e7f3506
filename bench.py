#!/usr/bin/env python
"""psac-b200 benchmark: suffixes/s of SA+LCP construction on synthetic text (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|1|3|5|6] [--log2n L]

One "step" = one complete construction (the hot path of BASELINE.json) over one synthetic text.

  --config 2 (default): the per-GPU slice of BASELINE configs[2] -- random DNA, 2^30 characters per GPU, SA+LCP (+ISA),
             64-bit index.  N = 1: one text on one GPU; N > 1: ONE text of N x 2^30 characters sharded by block over the N
             GPUs (N = 8 is configs[2] itself: 8 GiB).  The 1 -> 8 curve is like for like (same index width, same outputs).
  --config 1: configs[1], 2^30 DNA, SA only, 32-bit index, one GPU (also reported as the extra key "configs1" of the default line)
  --config 3: configs[3], 2^32 random bytes (|Sigma| = 256, last byte != 0xFF), SA only, 64-bit index, one GPU
  --config 5: configs[4], DNA, SA+LCP + ANSV + suffix-tree child table, 2^29 characters per GPU
  --config 6: generalized SA+LCP of a string set (reference construct_ss): 2^28 bytes of '$'-separated DNA reads, one GPU (bench_gsa.py)

  value     : suffixes/s, text resident in HBM, outputs left in HBM (psacb200_construct_device / _construct_sharded)
  e2e       : suffixes/s through the reference-facing C-ABI call with HOST (pinned) buffers: H2D of the text and D2H of
              SA, ISA and LCP inside the timed region
  roofline  : dominant kernel = the scatter kernel of one radix digit pass of the first sort; algorithmic bytes per launch
              = elements sorted on this GPU x 2 x (carried key + suffix index) bytes
  verified  : the LAST timed result is certified on the device by psacb200_check_device / _check_sharded (reference
              d_check_sa conditions + LCP by direct comparison); N > 1 additionally runs small sharded parity cases against
              the CPU oracle on rank 0 before timing
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

SEEDS = {1: 2, 2: 3, 3: 4, 5: 5}
METRIC = {1: "suffixes/sec SA build", 2: "suffixes/sec SA+LCP build", 3: "suffixes/sec SA build", 5: "suffixes/sec SA+LCP+suffix-tree build"}


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


def make_text_np(config, m, seed, start=0):
    from psac_b200 import textgen as G
    if config == 3:
        t = G.random_bytes(m, seed, start)
        if m and t[-1] == 0xFF:
            t[-1] = 0xFE
        return t
    return G.random_dna(m, seed, start)


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


def sharded_parity_cases(dev):
    """Small sharded constructions compared element-wise with the CPU oracle on rank 0 (the oracle as the checker only)."""
    import torch
    import torch.distributed as dist
    from psac_b200 import api, textgen as G
    from psac_b200.sharded import ShardedSuffixArray
    rank, p = dist.get_rank(), dist.get_world_size()
    cases = [("random DNA aligned", G.random_dna(p << 17, 101), 8, True, 0),
             ("random DNA ragged, 2 rounds", G.random_dna((p << 17) + 11, 102), 8, True, 6),
             ("random bytes (|Sigma|=256 quirk)", G.random_bytes_config4(p << 16, 103), 8, False, 0)]
    report = []
    for name, text, ib, lcp, k in cases:
        n = text.size
        start, size = api.blk_dist(n, p, rank)
        ssa = ShardedSuffixArray(ib, lcp)
        ssa.construct(text[start:start + size], k=k)
        chk = ssa.check()
        sizes = [api.blk_dist(n, p, r)[1] for r in range(p)]
        mx = max(sizes)

        def gather(t):
            pad = torch.zeros(mx, dtype=t.dtype, device=dev)
            pad[: t.numel()] = t
            out = [torch.zeros(mx, dtype=t.dtype, device=dev) for _ in range(p)]
            dist.all_gather(out, pad)
            return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)]).view(np.uint64)

        got = {"sa": gather(ssa.local_SA), "isa": gather(ssa.local_B)}
        if lcp:
            got["lcp"] = gather(ssa.local_LCP)
        ok = 1
        if rank == 0:
            from oracle import pyoracle as O  # the checker
            exp = O.construct(text, 64, 0, lcp)
            ok = int(all((got[key] == exp[key]).all() for key in got) and chk["ok"])
        flag = torch.tensor([ok], device=dev)
        dist.broadcast(flag, 0)
        ssa.close()
        report.append("%s (n=%d): %s" % (name, n, "bit-exact vs oracle, device check ok" if flag.item() else "MISMATCH"))
        if not flag.item():
            raise SystemExit("bench.py: sharded parity case failed: " + report[-1])
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 5, 6])
    ap.add_argument("--log2n", type=int, default=0, help="log2 of the characters per GPU (default: the config's size)")
    ap.add_argument("--cpu-log2n", type=int, default=25, help="log2 of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the small sharded parity cases before timing")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the extra configs[1] measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.config
    if cfg == 5:
        from bench_tree import main_tree  # configs[4]: SA + LCP + ANSV + suffix tree
        return main_tree(args, rank, world, local_rank)
    if cfg == 6:
        from bench_gsa import main_gsa  # generalized suffix array of a string set
        return main_gsa(args, rank, world, local_rank)
    log2n = args.log2n or {1: 30, 2: 30, 3: 32}[cfg]
    n = 1 << log2n
    ngpu = max(world, args.gpus)
    want_lcp = cfg == 2
    index_bytes = 4 if cfg == 1 else 8
    seed = SEEDS[cfg]
    what = {1: "random DNA (|Sigma|=4)", 2: "random DNA (|Sigma|=4)", 3: "random bytes (|Sigma|=256, last byte != 0xFF)"}[cfg]
    if ngpu > 1:
        workload = "%s, ONE text of %d x 2^%d chars sharded by block over %d GPUs, SA%s (+ISA), %d-bit index (BASELINE configs[2]%s)" % (
            what, ngpu, log2n, ngpu, "+LCP" if want_lcp else "-only", index_bytes * 8, "" if (ngpu == 8 and log2n == 30) else " shape")
    else:
        workload = "%s 2^%d chars on one GPU, SA%s (+ISA), %d-bit index (%s)" % (
            what, log2n, "+LCP" if want_lcp else "-only", index_bytes * 8,
            {1: "BASELINE configs[1]", 2: "per-GPU slice of BASELINE configs[2]", 3: "BASELINE configs[3]"}[cfg])
    from psac_b200 import textgen as G

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        m = 1 << min(args.cpu_log2n, log2n)
        text = make_text_np(cfg, m, seed)
        steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
        kind, sps, ms = reference_run(text, index_bytes, want_lcp, steps, warmup)
        line = {"impl": "reference", "metric": METRIC[cfg], "value": sps, "unit": "suffixes/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u%d" % (index_bytes * 8),
                "data": "synthetic", "config": {"workload": workload, "cpu_sample": "first 2^%d characters of the same text per step" % int(np.log2(m)), "n_per_gpu": n, "n_total": n * ngpu},
                "cpu_baseline": {"value": sps, "unit": "suffixes/s", "cores": 1, "kind": kind,
                                 "sample": "first 2^%d characters of the text; unmodified psac at np=1 under the MPI shim (single core: the box has no MPI; the reference is single-threaded per rank)" % int(np.log2(m))},
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
    parity_report = None
    if sharded and not args.no_parity:
        parity_report = sharded_parity_cases(dev)
    flags = (api.LCP if want_lcp else 0) | api.FAST_RESOLVAL
    n_total = n * world
    if sharded:
        # SA / ISA / LCP come back block-distributed.  Rank r generates its own block of the text (seed differs per rank).
        from psac_b200.sharded import ShardedSuffixArray
        ssa = ShardedSuffixArray(index_bytes, want_lcp)
        eng = ssa.engine
    else:
        eng = api.Engine(local_rank)
    if cfg == 3:
        text = G.random_bytes_torch(n, seed, dev)  # uniform bytes; last byte forced != 0xFF
        if int(text[-1].item()) == 0xFF:
            text[-1] = 0xFE
    else:
        text = G.random_dna_torch(n, seed + rank, dev)
    tdt = torch.int32 if index_bytes == 4 else torch.int64  # raw storage for unsigned outputs
    d_sa = torch.empty(n, dtype=tdt, device=dev)
    d_isa = torch.empty(n, dtype=tdt, device=dev)
    d_lcp = torch.empty(n, dtype=tdt, device=dev) if want_lcp else None
    if not sharded and cfg != 3:
        eng.reserve(n, index_bytes, flags)  # (configs[3] lets the first warm-up call size the buffers: the caller's 64-bit outputs hold SA / ISA)
    torch.cuda.synchronize()

    def step_device():
        if sharded:
            eng.construct_sharded_ptr(text.data_ptr(), n, n_total, index_bytes, flags, 0, d_sa.data_ptr(), d_isa.data_ptr(),
                                      d_lcp.data_ptr() if d_lcp is not None else None)
        else:
            eng.construct_ptr(text.data_ptr(), n, index_bytes, flags, 0, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr() if d_lcp is not None else None,
                              device=True)

    ext = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    warm = max(args.warmup, 3)
    for _ in range(warm):
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
    trace = eng.trace()  # device timeline of the last timed step (this rank)
    clocks = sampler.summary()
    stats = eng.stats()

    # ------------------------------------------------------------------ certificate of the last timed result, on the device
    if sharded:
        chk = eng.check_sharded_ptr(text.data_ptr(), n, n_total, index_bytes, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr() if d_lcp is not None else None)
    else:
        chk = eng.check_device_ptr(text.data_ptr(), n, index_bytes, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr() if d_lcp is not None else None)
    if not chk["ok"]:
        raise SystemExit("bench.py: device-side check FAILED -- result invalid: %r" % (chk,))
    verified = "d_check_sa (SA permutation, ISA inverse, suffix order) %son device over all %d positions of the last timed result: 0 violations (%.1f ms)" % (
        "+ LCP by direct text comparison " if want_lcp else "", n_total, chk["ms"])

    # ------------------------------------------------------------------ end to end: pinned host buffers through the public API
    e2e = None
    if not args.no_e2e:
        h_text = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_text.copy_(text)
        h_sa = torch.empty(n, dtype=tdt).pin_memory()
        h_isa = torch.empty(n, dtype=tdt).pin_memory()
        h_lcp = torch.empty(n, dtype=tdt).pin_memory() if want_lcp else None
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
        outs = 2 + (1 if want_lcp else 0)
        e2e = {"ms": ms_e2e, "h2d": n * world, "d2h": outs * n * index_bytes * world, "steps": k_e2e}
        del h_text, h_sa, h_isa, h_lcp

    # ------------------------------------------------------------------ extra: BASELINE configs[1] on the same GPU (N = 1 only)
    extra = None
    if not sharded and cfg == 2 and not args.no_extra:
        del d_lcp
        t1 = G.random_dna_torch(n, SEEDS[1], dev)
        s32 = d_sa.view(torch.int32)[:n]
        i32 = d_isa.view(torch.int32)[:n]
        torch.cuda.synchronize()
        f1 = api.FAST_RESOLVAL
        for _ in range(2):
            eng.construct_ptr(t1.data_ptr(), n, 4, f1, 0, s32.data_ptr(), i32.data_ptr(), None, device=True)
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k1 = max(1, min(args.steps, 5))
        x0.record(ext)
        for _ in range(k1):
            eng.construct_ptr(t1.data_ptr(), n, 4, f1, 0, s32.data_ptr(), i32.data_ptr(), None, device=True)
        x1.record(ext)
        torch.cuda.synchronize()
        ms1 = x0.elapsed_time(x1) / k1
        c1 = eng.check_device_ptr(t1.data_ptr(), n, 4, s32.data_ptr(), i32.data_ptr(), None)
        if not c1["ok"]:
            raise SystemExit("bench.py: device-side check of the configs[1] result FAILED: %r" % (c1,))
        extra = {"workload": "BASELINE configs[1]: random DNA 2^%d chars, SA only (+ISA), 32-bit index, one GPU" % log2n, "value": n / (ms1 * 1e-3),
                 "unit": "suffixes/s", "ms_per_step": ms1, "steps": k1, "verified": "d_check_sa on device: 0 violations"}

    # ------------------------------------------------------------------ reduce over ranks (max time)
    t = torch.tensor([ms_dev, e2e["ms"] if e2e else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            ssa.close()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # dominant kernel: the scatter kernel of one 8-bit digit pass over the carried keys (passes 2..P of the first sort);
    # algorithmic bytes per launch = read + write of one carried key and one suffix index per suffix sorted on this GPU
    elt_bytes = int(stats.get("sort_elt_bytes", 0)) or (8 if not sharded and stats["key_chars"] * stats["pack_bits"] - 8 <= 32 and not (cfg == 3) else 16)
    pass_bytes = float(n) * 2 * elt_bytes
    pass_avg_ms = float(np.mean(pass_ms))
    pass_achieved = pass_bytes / (pass_avg_ms * 1e-3) / 1e9 if pass_avg_ms > 0 else 0.0
    # The engine brackets every launch of that kernel with CUDA events on its stream (psacb200_stats.ms_scatter_avg).
    # Where that is not available the whole pass (histogram + scans + scatter) is reported as one unit.
    sc_ms = float(np.mean(scatter_ms)) if scatter_ms and min(scatter_ms) > 0 else 0.0
    if sc_ms > 0:
        kernel_name = "radix_scatter_seg_kernel, %d-byte elements (scatter kernel of one 8-bit digit pass inside the segments; %d launches per step)" % (
            elt_bytes, stats["sort_passes"] - 1)
        kern_ms, achieved = sc_ms, pass_bytes / (sc_ms * 1e-3) / 1e9
    else:
        kernel_name = "one 8-bit digit pass over %d-byte elements = tile histogram + scan kernels + radix_scatter kernel (timed as one unit; rank 0)" % elt_bytes
        kern_ms, achieved = pass_avg_ms, pass_achieved
    traffic = None  # DRAM bytes per launch from the committed ncu capture (same kernel, n = 2^30), scaled to this n
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
        if elt_bytes == 8:
            per = tj["dram_bytes_scatter_kernel"] if sc_ms > 0 else tj["dram_bytes_per_pass"]
            traffic = float(per) * n / float(tj["n"])
    except Exception:
        traffic = None
    line = {
        "metric": METRIC[cfg], "value": n_total / (ms_dev * 1e-3), "unit": "suffixes/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u%d" % (index_bytes * 8), "data": "synthetic",
        "config": {"workload": workload, "n_per_gpu": n, "n_total": n_total,
                   "parallelism": ("block-sharded text/SA/ISA/LCP over %d GPUs, one process per GPU" % world) if sharded else "single GPU", "seed": seed,
                   "l2": "inputs_exceed_l2 (every pass streams >= 8 GiB)", "key_chars": stats["key_chars"], "sort_passes": stats["sort_passes"],
                   "rounds": stats["rounds"], "unresolved_after_first": stats["unresolved_after_first"], "verified": verified,
                   "internal_index_bytes": stats["internal_index_bytes"], "device_bytes": stats["device_bytes"],
                   "sharded_parity": parity_report,
                   "exchange": ("peer stores over NVLink fused into the partition kernels" if stats.get("peer_exchange") else "NCCL all-to-all-v") if sharded else None},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": "profiles/roofline_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None,
                     "peak_source": peak_src, "bytes_per_launch": pass_bytes, "ms_per_launch": kern_ms,
                     "whole_pass": {"what": "histogram pre-pass + scans + scatter of one digit pass", "ms": pass_avg_ms, "achieved": pass_achieved,
                                    "frac": pass_achieved / peak}},
        "phases_ms": {k: round(v, 3) for k, v in sorted(phase.items())},
        "trace_ms": [[k, round(v, 3)] for k, v in trace],
    }
    if e2e:
        line["e2e"] = {"value": n_total / (ms_e2e * 1e-3), "unit": "suffixes/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "ms_per_step": ms_e2e, "steps": e2e["steps"]}
    if extra:
        line["configs1"] = extra
    if not args.no_cpu_baseline and world == 1:
        m = 1 << min(args.cpu_log2n, log2n)
        kind, sps, ms = reference_run(make_text_np(cfg, m, seed), index_bytes, want_lcp, 1, 0)
        line["cpu_baseline"] = {"value": sps, "unit": "suffixes/s", "cores": 1, "kind": kind, "ms": ms,
                                "sample": "first 2^%d characters of the same text, one run; unmodified psac at np=1 under the MPI shim" % int(np.log2(m))}
    print(json.dumps(line), flush=True)
    if world > 1:
        ssa.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
