"""bench.py --config 6: generalized suffix array of a string set (reference construct_ss,
include/suffix_array.hpp:269-363) -- a set of DNA reads of 150 characters separated by '$', 2^28 characters in total,
SA + ISA + LCP, 64-bit index, one GPU.  Not one of BASELINE.json's configs (the reference publishes no number for
construct_ss); it is the measurement of the "next" row f2 at the bar of the others: device-resident `value`, `e2e`
through the host-buffer C ABI, `roofline` of the dominant kernel, the unmodified reference timed beside it.

One step = psacb200_construct_ss_device over the flat text.  Certificate of the last timed result: SA is a permutation and
ISA its inverse (all positions, on the device with torch), suffix order with ties by position and LCP by direct comparison
on 20 000 sampled neighbours (host).
"""
import json
import time

import numpy as np

READ_LEN = 150


def make_reads_torch(n_flat, seed, dev):
    """flat text of n_flat bytes: uniform DNA reads of READ_LEN characters, '$' after every read"""
    from psac_b200 import textgen as G
    t = G.random_dna_torch(n_flat, seed, dev)
    t[READ_LEN::READ_LEN + 1] = ord("$")
    return t


def make_reads_np(n_flat, seed):
    from psac_b200 import textgen as G
    t = G.random_dna(n_flat, seed)
    t[READ_LEN::READ_LEN + 1] = ord("$")
    return t


def sample_check(flat_np, sa, lcp, samples, seed=1):
    """order (ties by position) and LCP of sampled neighbours by direct comparison"""
    n = sa.size
    stride = READ_LEN + 1

    def suffix(c):  # c indexes the concatenation without separators
        r, o = divmod(int(c), READ_LEN)
        return flat_np[r * stride + o: r * stride + READ_LEN].tobytes()

    rng = np.random.default_rng(seed)
    for q in rng.integers(1, n, samples):
        a, b = int(sa[q - 1]), int(sa[q])
        x, y = suffix(a), suffix(b)
        if not (x < y or (x == y and a < b)):
            return False
        l = 0
        while l < len(x) and l < len(y) and x[l] == y[l]:
            l += 1
        if int(lcp[q]) != l:
            return False
    return int(lcp[0]) == 0


def main_gsa(args, rank, world, local_rank):
    import torch
    from bench import ClockSampler, measured_peaks
    from psac_b200 import api

    log2n = args.log2n or 28
    n_flat = 1 << log2n
    seed = 6
    metric = "suffixes/sec generalized SA+LCP build (string set)"
    workload = "DNA reads of %d characters ('$'-separated string set), 2^%d bytes of flat text on one GPU, generalized SA+LCP (+ISA), 64-bit index (reference construct_ss)" % (READ_LEN, log2n)
    if args.impl == "reference" or not args.no_cpu_baseline:
        from oracle import pyoracle as O
        m = 1 << min(args.cpu_log2n - 3, log2n)
        sample = make_reads_np(m, seed)
        kind = "reference" if O.have_ref() else "port"
        fn = (lambda: O.ref_construct_ss(sample, ord("$"), 8)) if kind == "reference" else (lambda: O.construct_ss(sample, ord("$"), 64, True))
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, args.steps)
        t0 = time.perf_counter()
        for _ in range(steps):
            r = fn()
        dt = (time.perf_counter() - t0) / steps
        sps = r["n"] / dt
        line = {"impl": "reference", "metric": metric, "value": sps, "unit": "suffixes/s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload, "cpu_sample": "first 2^%d bytes of the same flat text per step" % int(np.log2(m))},
                "cpu_baseline": {"value": sps, "unit": "suffixes/s", "cores": 1, "kind": kind,
                                 "sample": "first 2^%d bytes of the flat text; unmodified psac construct_ss at np=1 under the MPI shim" % int(np.log2(m))},
                "e2e": {"value": sps, "unit": "suffixes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": metric, "unavailable": "construct_ss runs on one GPU (string sets are not sharded in this round)"}), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- psac-b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = api.Engine(local_rank)
    flat = make_reads_torch(n_flat, seed, dev)
    d_sa = torch.empty(n_flat, dtype=torch.int64, device=dev)
    d_isa = torch.empty(n_flat, dtype=torch.int64, device=dev)
    d_lcp = torch.empty(n_flat, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    n = 0

    def step_device():
        return eng.construct_ss_ptr(flat.data_ptr(), n_flat, ord("$"), 8, api.LCP, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr(), device=True)

    ext = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        n = step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase, pass_ms = {}, []
    ev0.record(ext)
    for _ in range(args.steps):
        step_device()
        s = eng.stats()
        pass_ms.append(s["ms_sort_pass_avg"])
        for k, v in s.items():
            if k.startswith("ms_"):
                phase[k] = phase.get(k, 0.0) + v / args.steps
    ev1.record(ext)
    torch.cuda.synchronize()
    ms_dev = ev0.elapsed_time(ev1) / args.steps
    launches = eng.launches - launches0
    trace = eng.trace()
    clocks = sampler.summary()
    stats = eng.stats()

    # ---- certificate of the last timed result
    sa, isa, lcp = d_sa[:n], d_isa[:n], d_lcp[:n]
    ar = torch.arange(n, dtype=torch.int64, device=dev)
    perm_ok = bool((torch.sort(sa).values == ar).all().item()) and bool((isa[sa] == ar).all().item())
    del ar
    flat_np = flat.cpu().numpy()
    samples = 20000
    order_ok = sample_check(flat_np, sa.cpu().numpy(), lcp.cpu().numpy(), samples)
    if not (perm_ok and order_ok):
        raise SystemExit("bench.py: generalized SA check FAILED (permutation %s, sampled order / LCP %s)" % (perm_ok, order_ok))
    verified = "SA permutation + ISA inverse over all %d suffixes on the device; suffix order (ties by position) and LCP by direct comparison on %d sampled neighbours: 0 violations" % (n, samples)

    # ---- end to end: pinned host buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        h_flat = torch.empty(n_flat, dtype=torch.uint8).pin_memory()
        h_flat.copy_(flat)
        h_out = [torch.empty(n_flat, dtype=torch.int64).pin_memory() for _ in range(3)]
        torch.cuda.synchronize()

        def step_host():
            return eng.construct_ss_ptr(h_flat.data_ptr(), n_flat, ord("$"), 8, api.LCP, h_out[0].data_ptr(), h_out[1].data_ptr(), h_out[2].data_ptr(), device=False)

        step_host()
        k_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            step_host()  # (returns after synchronising the engine's stream)
        ms_e2e = (time.perf_counter() - t0) * 1e3 / k_e2e
        e2e = {"value": n / (ms_e2e * 1e-3), "unit": "suffixes/s", "h2d_bytes_per_step": n_flat, "d2h_bytes_per_step": 3 * n * 8, "ms_per_step": ms_e2e,
               "steps": k_e2e}
        del h_flat, h_out

    peak, peak_src = measured_peaks()
    elt = 8 + int(stats["internal_index_bytes"])  # 64-bit cut key + suffix index of the flat text
    pass_bytes = float(n_flat) * 2 * elt
    pass_avg = float(np.mean(pass_ms))
    achieved = pass_bytes / (pass_avg * 1e-3) / 1e9 if pass_avg > 0 else 0.0
    line = {"metric": metric, "value": n / (ms_dev * 1e-3), "unit": "suffixes/s", "n_gpus": 1, "steps": args.steps, "warmup": warm, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload, "n_flat": n_flat, "n_total": n, "strings": n_flat // (READ_LEN + 1) + 1, "seed": seed, "l2": "inputs_exceed_l2",
                       "key_chars": stats["key_chars"], "pack_bits": stats["pack_bits"], "sort_passes": stats["sort_passes"], "rounds": stats["rounds"],
                       "unresolved_after_first": stats["unresolved_after_first"], "verified": verified, "device_bytes": stats["device_bytes"]},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "one 8-bit digit pass over %d-byte (cut key, suffix) pairs = tile histogram + scans + radix_scatter_kernel, timed as one unit (%d passes per step)" % (elt, stats["sort_passes"]),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "bytes_per_launch": pass_bytes, "ms_per_launch": pass_avg},
            "phases_ms": {k: round(v, 3) for k, v in sorted(phase.items())}, "trace_ms": [[k, round(v, 3)] for k, v in trace]}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": r["n"] / dt, "unit": "suffixes/s", "cores": 1, "kind": kind, "ms": dt * 1e3,
                                "sample": "first 2^%d bytes of the same flat text, one run; unmodified psac construct_ss at np=1 under the MPI shim" % int(np.log2(m))}
    print(json.dumps(line), flush=True)
    eng.close()
