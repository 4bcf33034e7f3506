"""CPU checks of the HOST-side logic of the sharded construction (psac_b200/csrc/sharded.cuh): block distribution,
key-range splitters and the exchange plans, run at world_size 2 and 3 over torch.distributed's gloo backend.
The device kernels are not called here (no GPU): the test plays their part with numpy / the CPU oracle and checks that
the plans the product computes (psacb200_blk_dist, psacb200_choose_splitters) move every suffix to the right place,
i.e. that the block-distributed SA / ISA assembled through those plans equal the oracle's.
Reference behaviour mirrored: mxx::blk_dist (ext/mxx/include/mxx/partition.hpp:283-331), sample-sort splitters with
equal keys kept together (ext/mxx/include/mxx/samplesort.hpp:191-238), bulk_permute_inplace (include/bulk_permute.hpp:14-73)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from psac_b200 import api  # noqa: E402


def test_blk_dist_matches_reference_rule():
    # first n % p ranks hold ceil(n / p) (partition.hpp:283-331)
    for n, p in ((10, 3), (11, 4), (8, 8), (7, 8), (1 << 20, 13), (0, 2)):
        sizes = [api.blk_dist(n, p, r)[1] for r in range(p)]
        starts = [api.blk_dist(n, p, r)[0] for r in range(p)]
        assert sum(sizes) == n
        assert sizes == [n // p + (1 if r < n % p else 0) for r in range(p)]
        assert starts == [int(np.sum(sizes[:r])) for r in range(p)]


def test_splitters_keep_bins_whole_and_balance():
    rng = np.random.default_rng(5)
    for p in (2, 3, 8):
        hist = rng.integers(0, 50, size=1 << 10).astype(np.uint64)
        n = int(hist.sum())
        first, count = api.choose_splitters(hist, n, p)
        assert first[0] == 0 and first[p] == hist.size and (np.diff(first.astype(np.int64)) >= 0).all()
        assert int(count.sum()) == n
        for r in range(p):
            assert int(count[r]) == int(hist[int(first[r]):int(first[r + 1])].sum())
        assert int(count.max()) <= n // p + int(hist.max()) + 1  # a range overshoots its share by less than one bin
    # one huge bin cannot be split: everything lands on one rank, the others get nothing (the engine then falls back)
    hist = np.zeros(16, np.uint64)
    hist[3] = 1000
    first, count = api.choose_splitters(hist, 1000, 4)
    assert sorted(count.tolist()) == [0, 0, 0, 1000]


def _worker(rank, world, port, n, seed, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as O
        from psac_b200 import textgen as G
        text = G.random_dna(n, seed)
        start, size = api.blk_dist(n, world, rank)
        block = text[start:start + size]
        # (S1) alphabet: local histograms summed over the ranks
        h = torch.from_numpy(np.bincount(block, minlength=256).astype(np.int64))
        dist.all_reduce(h)
        assert (h.numpy() == np.bincount(text, minlength=256)).all()
        # (S2) replicated text: all-gather of the blocks (ragged: pad to the largest block)
        mx = max(api.blk_dist(n, world, r)[1] for r in range(world))
        pad = torch.zeros(mx, dtype=torch.uint8)
        pad[:size] = torch.from_numpy(block.copy())
        parts = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, pad)
        full = np.concatenate([parts[r].numpy()[:api.blk_dist(n, world, r)[1]] for r in range(world)])
        assert (full == text).all()
        # (S3) key-prefix histogram of the LOCAL block (7 characters = 14 bits), all-gathered
        C, pchars = 20, 7
        codes = np.searchsorted(np.unique(text), full).astype(np.uint64)  # dense codes 0..3
        padded = np.concatenate([codes, np.zeros(C, np.uint64)])
        def key_of(g, chars):
            k = np.zeros(g.shape, np.uint64)
            for c in range(chars):
                k = (k << np.uint64(2)) | padded[g + c]
            return k
        gl = np.arange(start, start + size, dtype=np.int64)
        hist_local = np.bincount(key_of(gl, pchars).astype(np.int64), minlength=1 << 14).astype(np.int64)
        rows = [torch.zeros(1 << 14, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(rows, torch.from_numpy(hist_local))
        h2d = np.stack([r.numpy() for r in rows]).astype(np.uint64)
        first, count = api.choose_splitters(h2d.sum(axis=0), n, world)
        # (S5/S6) every rank selects ITS key range out of the whole text and sorts it (numpy plays the device sort)
        allg = np.arange(n, dtype=np.int64)
        bins = key_of(allg, pchars)
        mine = allg[(bins >= first[rank]) & (bins < first[rank + 1])]
        assert mine.size == int(count[rank])
        exp = O.construct(text, 64, 0, False)
        off = int(count[:rank].sum())
        order = np.argsort(exp["isa"][mine], kind="stable")  # the oracle's order stands in for the device sort + rounds
        sa_range = mine[order]
        assert (sa_range == exp["sa"][off:off + mine.size].astype(np.int64)).all(), "key ranges are not contiguous SA ranges"
        # (S9) SA -> ISA exchange plan: counts from the 2-D histogram, pairs partitioned by owner, all-to-all
        send_counts = [int(h2d[b, int(first[rank]):int(first[rank + 1])].sum()) for b in range(world)]
        recv_counts = [int(h2d[rank, int(first[a]):int(first[a + 1])].sum()) for a in range(world)]
        assert sum(send_counts) == mine.size and sum(recv_counts) == size
        owner = np.array([next(r for r in range(world) if api.blk_dist(n, world, r)[0] + api.blk_dist(n, world, r)[1] > g) for g in sa_range[:2000]])
        bounds = np.cumsum([0] + [api.blk_dist(n, world, r)[1] for r in range(world)])
        owner_all = np.searchsorted(bounds, sa_range, side="right") - 1
        assert (owner == owner_all[:2000]).all()
        assert [int((owner_all == b).sum()) for b in range(world)] == send_counts
        pos = off + np.arange(mine.size, dtype=np.int64)
        outs = [torch.from_numpy(np.stack([sa_range[owner_all == b], pos[owner_all == b]], axis=1).copy()) for b in range(world)]
        ins = [torch.zeros((recv_counts[a], 2), dtype=torch.int64) for a in range(world)]
        dist.all_to_all(ins, outs) if dist.get_backend() != "gloo" else _a2a_gloo(ins, outs, rank, world)
        isa_local = np.full(size, -1, np.int64)
        for t in ins:
            a = t.numpy()
            isa_local[a[:, 0] - start] = a[:, 1]
        assert (isa_local == exp["isa"][start:start + size].astype(np.int64)).all()
        # (S10) re-balancing of SA from key-range ownership to exact blocks
        offs = np.concatenate([[0], np.cumsum(count.astype(np.int64))])
        sends = []
        for b in range(world):
            bs, bl = api.blk_dist(n, world, b)
            lo, hi = max(off, bs), min(off + mine.size, bs + bl)
            sends.append(torch.from_numpy(sa_range[lo - off:hi - off].copy() if hi > lo else np.zeros(0, np.int64)))
        recvs = []
        for a in range(world):
            lo, hi = max(int(offs[a]), start), min(int(offs[a + 1]), start + size)
            recvs.append(torch.zeros(max(hi - lo, 0), dtype=torch.int64))
        _a2a_gloo(recvs, sends, rank, world)
        sa_block = np.concatenate([r.numpy() for r in recvs])
        assert (sa_block == exp["sa"][start:start + size].astype(np.int64)).all()
        results[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        results[rank] = "FAIL: %r" % (e,)
    finally:
        dist.destroy_process_group()


def _a2a_gloo(ins, outs, rank, world):
    """all-to-all-v with point-to-point messages (gloo has no all_to_all)"""
    reqs = []
    for peer in range(world):
        if peer == rank:
            ins[peer].copy_(outs[peer])
            continue
        if outs[peer].numel():
            reqs.append(dist.isend(outs[peer].contiguous(), peer))
        if ins[peer].numel():
            reqs.append(dist.irecv(ins[peer], peer))
    for r in reqs:
        r.wait()


@pytest.mark.parametrize("world,n", [(2, 60000), (3, 50021), (4, 40009)])
def test_exchange_plans_over_gloo(world, n):
    port = 29650 + world
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, 100 + world, results), nprocs=world, join=True)
    assert [results.get(r) for r in range(world)] == ["ok"] * world


# ------------------------------------------------------------------------------------------------ v2: word exchange plan
def _v2_worker(rank, world, port, n, seed, results):
    """The exchange that is fused into digit pass 1 of the sharded sort (sharded.cuh construct_sharded_v2), with numpy playing
    the kernels: every rank partitions ITS text block by the top key digit and sends each digit's run to the offset the
    product's plan (psacb200_plan_word_exchange) assigns inside the owner's padded buffer."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as O
        from psac_b200 import textgen as G
        text = G.random_dna(n, seed)
        C, tb, tile = 12, 8, 64
        K = 2 * C
        cb = K - tb
        ib = int(n - 1).bit_length()
        start, size = api.blk_dist(n, world, rank)
        codes = np.searchsorted(np.unique(text), text).astype(np.uint64)
        padded = np.concatenate([codes, np.zeros(C, np.uint64)])
        T = C - 1
        # my elements in feeding order: on the last rank the suffixes that run past the end come first, shortest first
        if rank == world - 1:
            g = np.concatenate([n - 1 - np.arange(T, dtype=np.int64), np.arange(start, n - T, dtype=np.int64)])
        else:
            g = np.arange(start, start + size, dtype=np.int64)
        key = np.zeros(g.shape, np.uint64)
        for c in range(C):
            key = (key << np.uint64(2)) | padded[g + c]
        digit = (key >> np.uint64(cb)).astype(np.int64)
        word = ((key & np.uint64((1 << cb) - 1)) << np.uint64(ib)) | g.astype(np.uint64)
        cnt_local = np.bincount(digit, minlength=1 << tb).astype(np.int64)
        rows = [torch.zeros(1 << tb, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(rows, torch.from_numpy(cnt_local))
        cnt = np.stack([r.numpy() for r in rows]).astype(np.uint64)
        P = api.plan_word_exchange(cnt, n, tile)
        assert P["balanced"]
        assert (P["seg_pad"] % tile == 0).all() and int(P["cnt_key"].sum()) == n
        # send every digit's run (stable partition) to its owner at the planned offset
        order = np.argsort(digit, kind="stable")
        outs = [[] for _ in range(world)]
        pos = 0
        for d in range(1 << tb):
            c = int(cnt_local[d])
            if c:
                outs[int(P["owner"][d])].append((int(P["run_off"][rank, d]), word[order[pos:pos + c]]))
                pos += c
        bufsize = int(P["seg_pad"][:, 256].max())
        mine = np.full(bufsize, np.uint64(2**64 - 1))
        written = np.zeros(bufsize, np.int64)
        for peer in range(world):
            msg = outs[peer]
            offs = np.array([o for o, _ in msg], np.int64)
            lens = np.array([w.size for _, w in msg], np.int64)
            data = np.concatenate([w for _, w in msg]) if msg else np.zeros(0, np.uint64)
            objs = [None] * world
            dist.all_gather_object(objs, (peer, offs, lens, data.view(np.int64)))
            for (dst, o_, l_, d_) in objs:
                if dst != rank:
                    continue
                q = 0
                for o1, l1 in zip(o_, l_):
                    mine[o1:o1 + l1] = d_[q:q + l1].view(np.uint64)
                    written[o1:o1 + l1] += 1
                    q += l1
        # every slot of my dense segments written exactly once, nothing in the padding
        dense, pad = P["seg_dense"][rank].astype(np.int64), P["seg_pad"][rank].astype(np.int64)
        exp = O.construct(text, 64, 0, False)
        off = int(P["cnt_key"][:rank].sum())
        sa_mine = []
        for d in range(256):
            ln = int(dense[d + 1] - dense[d])
            seg = mine[pad[d]:pad[d] + ln]
            assert (written[pad[d]:pad[d] + ln] == 1).all() and (written[pad[d] + ln:pad[d + 1]] == 0).all()
            # LSD passes = a stable sort of the segment by the carried key
            o2 = np.argsort(seg >> np.uint64(ib), kind="stable")
            sa_mine.append((seg[o2] & np.uint64((1 << ib) - 1)).astype(np.int64))
        sa_mine = np.concatenate(sa_mine)
        assert sa_mine.size == int(P["cnt_key"][rank])
        want = exp["sa"][off:off + sa_mine.size].astype(np.int64)
        # equal keys form unresolved buckets (any order inside), except that suffixes running past the end are singletons in
        # their final place already: compare as sets per complete key, and exactly for the tails
        kk = np.zeros(want.shape, np.uint64)
        for c in range(C):
            kk = (kk << np.uint64(2)) | padded[want + c]
        km = np.zeros(sa_mine.shape, np.uint64)
        for c in range(C):
            km = (km << np.uint64(2)) | padded[sa_mine + c]
        assert (kk == km).all(), "sorted keys differ from the oracle's SA range"
        assert (np.sort(sa_mine) == np.sort(want)).all()
        tails = sa_mine > n - C
        assert (sa_mine[tails] == want[tails]).all(), "suffixes that run past the end of the text are not in their final place"
        results[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback
        results[rank] = "FAIL: %r %s" % (e, traceback.format_exc()[-800:])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 40000), (3, 30011)])
def test_word_exchange_plan_over_gloo(world, n):
    port = 29680 + world
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_v2_worker, args=(world, port, n, 200 + world, results), nprocs=world, join=True)
    assert [results.get(r) for r in range(world)] == ["ok"] * world


def test_word_exchange_plan_skew_is_reported():
    cnt = np.zeros((4, 256), np.uint64)
    cnt[:, 7] = 1000  # one digit holds everything: no digit-boundary splitters can balance it
    P = api.plan_word_exchange(cnt, 4000, 8192)
    assert not P["balanced"]
