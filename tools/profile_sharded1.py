"""One-rank run of the SHARDED code path (PSACB200_FORCE_SHARDED=1) so that ncu can capture its kernels on a single GPU:
   PSACB200_FORCE_SHARDED=1 ncu ... python tools/profile_sharded1.py [log2n]"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1"); os.environ.setdefault("LOCAL_RANK", "0")
os.environ["PSACB200_FORCE_SHARDED"] = "1"
from psac_b200 import api, textgen as G
from psac_b200.sharded import ShardedSuffixArray
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
dist.init_process_group("nccl", device_id=dev)
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
ssa = ShardedSuffixArray(8, True); eng = ssa.engine
text = G.random_dna_torch(n, 3, dev)
sa = torch.empty(n, dtype=torch.int64, device=dev); isa = torch.empty_like(sa); lcp = torch.empty_like(sa)
torch.cuda.synchronize()
for it in range(2):
    eng.construct_sharded_ptr(text.data_ptr(), n, n, 8, api.LCP, 0, sa.data_ptr(), isa.data_ptr(), lcp.data_ptr())
print("scheme", eng.stats()["sharded_scheme"], "total ms", eng.stats()["ms_total"], " ".join("%s=%.2f" % kv for kv in eng.trace()))
chk = eng.check_sharded_ptr(text.data_ptr(), n, n, 8, sa.data_ptr(), isa.data_ptr(), lcp.data_ptr())
print("check", chk["ok"])
ssa.close(); dist.destroy_process_group()
