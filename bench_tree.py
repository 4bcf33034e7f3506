"""bench.py --config 5: BASELINE configs[4] -- DNA, SA + LCP, then ANSV + suffix-tree child table, device resident,
block-sharded over the GPUs (2^29 characters per GPU; 4 GPUs = the 2 GiB configuration).

One step = construct (SA + ISA + LCP, 64-bit index) followed by the suffix-tree construction from its blocks
(psacb200_suffix_tree_device / _suffix_tree_sharded: ANSV furthest_eq / nearest_sm searched on the fly over a min-tree,
child table rows filled in place, cross-rank edges through peer memory).  The standalone ANSV (psacb200_ansv_device /
_ansv_sharded, materialised left / right arrays) is timed separately and reported in `phases_ms`.
"""
import json
import os

import numpy as np


def main_tree(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, measured_peaks, reference_run, SEEDS, METRIC
    from psac_b200 import api, textgen as G

    log2n = args.log2n or 29
    n = 1 << log2n
    seed = SEEDS[5]
    if args.impl == "reference":
        if rank != 0:
            return
        m = 1 << min(args.cpu_log2n - 2, log2n)
        text = G.random_dna(m, seed)
        from oracle import pyoracle as O
        import time
        steps = max(1, args.steps)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.ref_suffix_tree(text)  # construct (SA + LCP) + construct_suffix_tree inside the reference driver
        dt = (time.perf_counter() - t0) / steps
        sps = m / dt
        line = {"impl": "reference", "metric": METRIC[5], "value": sps, "unit": "suffixes/s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": "random DNA, SA+LCP + suffix tree (BASELINE configs[4] shape)", "cpu_sample": "first 2^%d characters per step" % int(np.log2(m))},
                "cpu_baseline": {"value": sps, "unit": "suffixes/s", "cores": 1, "kind": "reference",
                                 "sample": "first 2^%d characters; unmodified psac construct + construct_suffix_tree at np=1 under the MPI shim" % int(np.log2(m))},
                "e2e": {"value": sps, "unit": "suffixes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- psac-b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sharded = world > 1
    if sharded:
        dist.init_process_group("nccl", device_id=dev)
        from psac_b200.sharded import ShardedSuffixArray
        ssa = ShardedSuffixArray(8, True)
        eng = ssa.engine
    else:
        eng = api.Engine(local_rank)
    n_total = n * world

    def barrier():
        if sharded:
            dist.barrier()
        torch.cuda.synchronize()

    text = G.random_dna_torch(n, seed + rank, dev)
    d_sa = torch.empty(n, dtype=torch.int64, device=dev)
    d_isa = torch.empty(n, dtype=torch.int64, device=dev)
    d_lcp = torch.empty(n, dtype=torch.int64, device=dev)
    width = 5  # sigma + 1
    d_nodes = torch.empty((n, width), dtype=torch.int64, device=dev)
    d_l = torch.empty(n, dtype=torch.int64, device=dev)
    d_r = torch.empty(n, dtype=torch.int64, device=dev)
    flags = api.LCP | api.FAST_RESOLVAL
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)

    def construct():
        if sharded:
            eng.construct_sharded_ptr(text.data_ptr(), n, n_total, 8, flags, 0, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr())
        else:
            eng.construct_ptr(text.data_ptr(), n, 8, flags, 0, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr(), device=True)

    def tree():
        if sharded:
            return eng.suffix_tree_sharded_ptr(text.data_ptr(), n, n_total, 8, d_sa.data_ptr(), d_lcp.data_ptr(), d_nodes.data_ptr(), d_nodes.numel())
        return eng.suffix_tree_device_ptr(text.data_ptr(), n, 8, d_sa.data_ptr(), d_lcp.data_ptr(), d_nodes.data_ptr(), d_nodes.numel())

    def ansv():
        if sharded:
            eng.ansv_sharded_ptr(d_lcp.data_ptr(), n, n_total, 8, 2, 0, 2 ** 63 - 1, d_l.data_ptr(), d_r.data_ptr())
        else:
            eng.ansv_device_ptr(d_lcp.data_ptr(), n, 8, 2, 0, 2 ** 63 - 1, d_l.data_ptr(), d_r.data_ptr())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        construct()
        tree()
    ansv()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    tr_tree = []
    ev[0].record(ext)
    for s in range(args.steps):
        construct()
        ev[2 * s + 1].record(ext)
        tree()
        tr_tree = eng.trace()
        ev[2 * s + 2].record(ext)
    barrier()
    ms_c = sum(ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps)) / args.steps
    ms_t = sum(ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps)) / args.steps
    ms_dev = ev[0].elapsed_time(ev[-1]) / args.steps
    launches = eng.launches - launches0
    clocks = sampler.summary()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(ext)
    ansv()
    a1.record(ext)
    barrier()
    ms_ansv = a0.elapsed_time(a1)
    tr_ansv = eng.trace()

    # ---- certificates on the device: SA / ISA / LCP by the checker; the tree by its invariants (every leaf n + i occurs exactly
    #      once in the table; internal node ids occur at most once; the ANSV arrays agree with the table's parents on a sample)
    if sharded:
        chk = eng.check_sharded_ptr(text.data_ptr(), n, n_total, 8, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr())
    else:
        chk = eng.check_device_ptr(text.data_ptr(), n, 8, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr())
    if not chk["ok"]:
        raise SystemExit("bench.py: device-side check FAILED: %r" % (chk,))
    flat = d_nodes.reshape(-1)
    leaves = flat[flat >= n_total] - n_total
    cnt = torch.tensor([leaves.numel(), int((flat > 0).sum().item())], dtype=torch.int64, device=dev)
    lsum = leaves.sum().reshape(1)
    if sharded:
        dist.all_reduce(cnt)
        dist.all_reduce(lsum)
    want_sum = (n_total * (n_total - 1) // 2) % (1 << 64)
    got_sum = int(lsum.item()) % (1 << 64)
    if int(cnt[0].item()) != n_total or got_sum != want_sum:
        raise SystemExit("bench.py: suffix tree certificate FAILED (leaves %d of %d)" % (int(cnt[0].item()), n_total))
    del flat, leaves

    t = torch.tensor([ms_dev, ms_c, ms_t, ms_ansv], dtype=torch.float64, device=dev)
    if sharded:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_c, ms_t, ms_ansv = [float(x) for x in t]
    if rank != 0:
        ssa.close()
        dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    stats = eng.stats()
    # tree fill: reads SA and LCP (8 + 8 bytes), writes every table row once (8 * (sigma + 1)); the ANSV is searched on the fly
    fill_bytes = float(n) * (16 + 8 * width)
    tree_kernel_ms = dict(tr_tree).get("tree", ms_t)
    line = {
        "metric": METRIC[5], "value": n_total / (ms_dev * 1e-3), "unit": "suffixes/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "random DNA (|Sigma|=4), ONE text of %d x 2^%d chars%s, SA+LCP (+ISA) then ANSV + suffix-tree child table, 64-bit index (BASELINE configs[4]%s)" % (
            world, log2n, " sharded by block over %d GPUs" % world if sharded else "", "" if (world == 4 and log2n == 29) else " shape"),
            "n_per_gpu": n, "n_total": n_total, "seed": seed, "l2": "inputs_exceed_l2",
            "verified": "d_check_sa + LCP on device: 0 violations; suffix tree: all %d leaves occur exactly once in the child table (count and checksum), %d cells occupied" % (
                n_total, int(cnt[1].item())),
            "cross_rank_edges_rank0": stats["unresolved_after_first"] if sharded else 0},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "suffix_tree_tile_kernel + tree_list_kernel (child table fill with on-the-fly ANSV; DNA: rows assembled in shared memory, no clearing pass)", "achieved": fill_bytes / (tree_kernel_ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": fill_bytes / (tree_kernel_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "bytes_per_launch": fill_bytes, "ms_per_launch": tree_kernel_ms},
        "phases_ms": {"construct_sa_isa_lcp": round(ms_c, 3), "suffix_tree": round(ms_t, 3), "ansv_standalone_left_furthest_eq_right_nearest_sm": round(ms_ansv, 3)},
        "trace_ms": {"suffix_tree": [[k, round(v, 3)] for k, v in tr_tree], "ansv": [[k, round(v, 3)] for k, v in tr_ansv]},
        "suffixes_per_s": {"construct": n_total / (ms_c * 1e-3), "suffix_tree": n_total / (ms_t * 1e-3), "ansv": n_total / (ms_ansv * 1e-3)},
    }
    if not args.no_cpu_baseline and world == 1:
        m = 1 << min(args.cpu_log2n - 2, log2n)
        from oracle import pyoracle as O
        import time
        tx = G.random_dna(m, seed)
        t0 = time.perf_counter()
        O.ref_suffix_tree(tx)  # construct (SA + LCP) + construct_suffix_tree inside the reference driver
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": m / dt, "unit": "suffixes/s", "cores": 1, "kind": "reference", "ms": dt * 1e3,
                                "sample": "first 2^%d characters, one run; unmodified psac construct + construct_suffix_tree at np=1" % int(np.log2(m))}
    print(json.dumps(line), flush=True)
    if sharded:
        ssa.close()
        dist.destroy_process_group()
