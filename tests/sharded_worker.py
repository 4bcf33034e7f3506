"""Worker of tests/test_gpu_sharded.py: run under torchrun with one rank per GPU.  Every rank builds its blocks of
SA / ISA / LCP with ShardedSuffixArray; rank 0 gathers them and compares with the CPU oracle.  Exit code 0 = parity."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402  (the checker)
from psac_b200 import api, textgen as G  # noqa: E402
from psac_b200.sharded import ShardedSuffixArray  # noqa: E402


def gather_blocks(t, n, p, rank, dev, sizes=None):
    """all ranks' blocks of a block-distributed int tensor -> one numpy array on rank 0"""
    if sizes is None:
        sizes = [api.blk_dist(n, p, r)[1] for r in range(p)]
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=t.dtype, device=dev)
    pad[: t.numel()] = t
    out = [torch.zeros(mx, dtype=t.dtype, device=dev) for _ in range(p)]
    dist.all_gather(out, pad)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)])


def case(name, text, index_bytes, lcp, k=0, scheme=None, tree=False):
    rank, p = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    n = text.size
    start, size = api.blk_dist(n, p, rank)
    sa = ShardedSuffixArray(index_bytes, lcp).construct(text[start:start + size], k=k)
    udt = np.uint32 if index_bytes == 4 else np.uint64
    got_sa = gather_blocks(sa.local_SA, n, p, rank, dev).view(udt)
    got_isa = gather_blocks(sa.local_B, n, p, rank, dev).view(udt)
    got_lcp = gather_blocks(sa.local_LCP, n, p, rank, dev).view(udt) if lcp else None
    # collective device-side certificate (same report on every rank); it reads the ISA blocks through peer memory
    chk = sa.check() if not os.environ.get("PSACB200_NO_PEER") else {"ok": True}
    ok = True
    if rank == 0:
        exp = O.construct(text, 64, 0, lcp)
        bad = []
        for name_, got_, exp_ in (("SA", got_sa, exp["sa"]), ("ISA", got_isa, exp["isa"])) + ((("LCP", got_lcp, exp["lcp"]),) if lcp else ()):
            neq = np.nonzero(got_.astype(np.uint64) != exp_)[0]
            if neq.size:
                bad.append("%s: %d mismatches, first at %d (got %d, want %d)" % (name_, neq.size, neq[0], int(got_[neq[0]]), int(exp_[neq[0]])))
        if not chk["ok"]:
            bad.append("device check: %r" % (chk,))
        ok = not bad
        if bad:
            print("   " + "; ".join(bad), flush=True)
        st = sa.engine.stats()
        if scheme is not None and st["sharded_scheme"] != scheme:
            print("   expected sharded scheme %d, ran %d" % (scheme, st["sharded_scheme"]), flush=True)
            ok = False
        print("%-38s n=%9d ib=%d lcp=%d k=%d scheme=%d rounds=%d unresolved=%d %s" % (name, n, index_bytes, int(lcp), k, st["sharded_scheme"], st["rounds"],
                                                                                    st["unresolved_after_first"], "ok" if ok else "MISMATCH"), flush=True)
    if tree and lcp:
        # suffix tree + ANSV from the blocks, sharded (psacb200_suffix_tree_sharded / _ansv_sharded)
        nodes = sa.construct_suffix_tree()
        width = nodes.shape[1]
        got_nodes = gather_blocks(nodes.reshape(-1), n * width, p, rank, dev, sizes=[api.blk_dist(n, p, r)[1] * width for r in range(p)]).view(np.uint64)
        lc_ = sa.left_branching_chars()
        got_lc = gather_blocks(lc_, n, p, rank, dev)
        if rank == 0 and not (got_lc == O.lc_from_sa_lcp(text, exp["sa"], exp["lcp"])).all():
            print("   left-branching characters differ", flush=True)
            ok = False
        l_, r_ = sa.ansv(2, 0, 2 ** 63 - 1)
        got_l = gather_blocks(l_, n, p, rank, dev).view(np.uint64)
        got_r = gather_blocks(r_, n, p, rank, dev).view(np.uint64)
        if rank == 0:
            want = O.ref_suffix_tree(text) if (O.have_ref() and index_bytes == 8) else None
            if want is None:
                e1 = api.Engine(torch.cuda.current_device())
                want = e1.suffix_tree(text, exp["sa"].astype(udt), exp["lcp"].astype(udt))
                e1.close()
            if not (got_nodes.reshape(n, width) == want).all():
                print("   suffix tree differs in %d cells" % int((got_nodes.reshape(n, width) != want).sum()), flush=True)
                ok = False
            wl = O.ansv_sequential  # nearest smaller on the right is what the sequential oracle gives directly
            er = wl(exp["lcp"].astype(np.uint64), False, 2 ** 63 - 1)
            if not (got_r == er).all():
                print("   sharded ANSV (right, nearest_sm) differs in %d places" % int((got_r != er).sum()), flush=True)
                ok = False
            if n <= 300000:
                el, _ = O.ansv(exp["lcp"].astype(np.uint64), 2, 0, 2 ** 63 - 1)
                if not (got_l == el).all():
                    print("   sharded ANSV (left, furthest_eq) differs in %d places" % int((got_l != el).sum()), flush=True)
                    ok = False
    if name.startswith("random DNA, aligned") and lcp:
        # the certificate must also FAIL when it should: corrupt one LCP entry and one SA entry of the last rank's block
        sa.local_LCP[3] += 1
        c1 = sa.check()
        sa.local_LCP[3] -= 1
        a, b = int(sa.local_SA[5]), int(sa.local_SA[6])
        sa.local_SA[5], sa.local_SA[6] = b, a
        c2 = sa.check()
        if rank == 0 and not (c1["bad_lcp"] == p and c1["bad_order"] == 0 and c2["bad_inverse"] == 2 * p and not c2["ok"]):
            print("   device check did not flag the corruption: %r %r" % (c1, c2), flush=True)
            ok = False
    sa.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    return bool(flag.item())


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p = dist.get_world_size()
    ok = True
    # scheme 2: word exchange fused into digit pass 1, distributed later rounds (sharded.cuh construct_sharded_v2)
    ok &= case("random DNA, aligned blocks", G.random_dna(p << 20, 11), 8, True, scheme=2)
    ok &= case("random DNA, ragged blocks + tree", G.random_dna((p << 20) + 13, 12), 8, True, scheme=2, tree=True)
    ok &= case("random DNA, 32-bit index + tree", G.random_dna((p << 17) + 5, 13), 4, True, scheme=2, tree=True)
    ok &= case("random DNA, no LCP", G.random_dna((p << 19) + 3, 23), 8, False, scheme=2)
    ok &= case("random DNA, k=7: distributed rounds", G.random_dna(p << 20, 18), 8, True, k=7, scheme=2)
    ok &= case("random DNA, short first key (k=4)", G.random_dna(p << 18, 14), 8, True, k=4, scheme=2)
    ok &= case("random DNA, one-digit key (k=2)", G.random_dna(p << 17, 24), 8, True, k=2, scheme=2)
    ok &= case("protein-like alphabet + tree", (G.random_bytes(p << 18, 16) % 20 + 65).astype(np.uint8), 8, True, scheme=2, tree=True)
    ok &= case("repetitive text + tree", G.repeats_text(40000 * p, 3), 8, True, tree=True)
    ok &= case("periodic text (abc)^k + tree", G.periodic_text(b"abc", 60000 * p), 4, True, tree=True)
    ok &= case("two-symbol text", (G.random_bytes(p << 17, 25) % 2 + 97).astype(np.uint8), 8, True)
    # keys that end inside a character (at full size: 39 key bits for 8 x 2^30 DNA characters); forced here at test size
    for kb in ("39", "23", "9"):
        os.environ["PSACB200_V2_KEYBITS"] = kb
        ok &= case("random DNA, %s key bits" % kb, G.random_dna((p << 19) + 17, 30 + int(kb)), 8, True, scheme=2)
    os.environ["PSACB200_V2_KEYBITS"] = "13"
    ok &= case("16-symbol text, 13 key bits", (G.random_bytes(p << 18, 27) % 16 + 97).astype(np.uint8), 8, True, scheme=2)
    ok &= case("repetitive text, 13 key bits", G.repeats_text(30000 * p, 5), 8, True, scheme=2)
    del os.environ["PSACB200_V2_KEYBITS"]
    ok &= case("random bytes (sigma=256 quirk)", G.random_bytes_config4(p << 18, 15), 8, False, scheme=1)
    ok &= case("small input (replicated path) + tree", G.random_dna(1000 + p, 17), 8, True, scheme=0, tree=True)
    # scheme 1 (the fallback): key-range selection + replicated rounds, all four exchange variants
    os.environ["PSACB200_SHARDED_V1"] = "1"
    ok &= case("v1: random DNA", G.random_dna((p << 19) + 1, 26), 8, True, scheme=1)
    os.environ["PSACB200_NO_PACK"] = "1"  # the unpacked (suffix, bucket) exchange
    ok &= case("v1: unpacked exchange", G.random_dna((p << 19) + 7, 19), 8, True, scheme=1)
    os.environ["PSACB200_NO_PEER"] = "1"  # ... and over NCCL all-to-all-v instead of peer stores
    ok &= case("v1: unpacked over NCCL", G.random_dna((p << 19) + 9, 20), 8, True, scheme=1)
    del os.environ["PSACB200_NO_PACK"]
    ok &= case("v1: packed over NCCL, k=7", G.random_dna((p << 19) + 11, 21), 8, True, k=7, scheme=1)
    del os.environ["PSACB200_NO_PEER"]
    del os.environ["PSACB200_SHARDED_V1"]
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
