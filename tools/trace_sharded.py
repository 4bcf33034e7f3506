# prints both ranks' traces of one sharded construct
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from psac_b200 import api, textgen as G
from psac_b200.sharded import ShardedSuffixArray
lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rank, p = dist.get_rank(), dist.get_world_size()
n = 1 << 30
ssa = ShardedSuffixArray(8, True); eng = ssa.engine
text = G.random_dna_torch(n, 3 + rank, dev)
sa = torch.empty(n, dtype=torch.int64, device=dev); isa = torch.empty_like(sa); lcp = torch.empty_like(sa)
torch.cuda.synchronize()
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    eng.construct_sharded_ptr(text.data_ptr(), n, n * p, 8, api.LCP, 0, sa.data_ptr(), isa.data_ptr(), lcp.data_ptr())
    tr = eng.trace()
    if it >= 2:
        for r in range(p):
            dist.barrier()
            if r == rank:
                print("rank %d it %d total %.2f: " % (rank, it, eng.stats()["ms_total"]) + " ".join("%s=%.2f" % kv for kv in tr), flush=True)
ssa.close(); dist.destroy_process_group()
