"""Sharded mirror of the reference class: one process (rank) per GPU, launched with torchrun.

``ShardedSuffixArray`` mirrors ``suffix_array<char, index_t, _CONSTRUCT_LCP>(comm)`` of the reference at p = world ranks
(include/suffix_array.hpp:170-228, 469-486): every rank calls ``construct`` with ITS block of the text (block-decomposed
like mxx::blk_dist) and ends up with its blocks ``local_SA`` / ``local_B`` (= ISA) / ``local_LCP`` -- here as CUDA tensors.
torch.distributed is only the plumbing that hands the NCCL id to every rank; the construction itself, including its
exchange steps over NCCL, runs inside libpsacb200.so (psac_b200/csrc/sharded.cuh).  There is no CPU path.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api


class ShardedSuffixArray:
    def __init__(self, index_bytes=8, construct_lcp=False, device=None):
        if not dist.is_initialized():
            raise api.PsacError("ShardedSuffixArray needs an initialised torch.distributed process group (one rank per GPU)")
        self.rank, self.p = dist.get_rank(), dist.get_world_size()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.index_bytes = index_bytes
        self.construct_lcp = construct_lcp
        self.engine = api.Engine(self.device.index)
        # rank 0 creates the NCCL id; every rank receives it through the caller's process group
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            uid = torch.from_numpy(api.Engine.comm_unique_id().copy())
        if dist.get_backend() == "nccl":
            t = uid.to(self.device)
            dist.broadcast(t, 0)
            uid = t.cpu()
        else:
            dist.broadcast(uid, 0)
        self.engine.comm_init(uid.numpy(), self.rank, self.p)
        self.n = 0
        self.local_size = 0
        self.local_SA = self.local_B = self.local_LCP = None

    def init_size(self, local_size):
        """reference :217-228 -- allreduce of the local sizes, then the block-decomposition check"""
        t = torch.tensor([int(local_size)], dtype=torch.int64, device=self.device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t)
        self.n = int(t.item())
        self.local_size = int(local_size)
        if api.blk_dist(self.n, self.p, self.rank)[1] != self.local_size:
            raise api.PsacError("The input string must be equally block decomposed accross all MPI processes.")

    def construct(self, local_text, fast_resolval=True, k=0):
        """local_text: this rank's block (uint8 CUDA tensor, or anything np.asarray takes)."""
        if not isinstance(local_text, torch.Tensor):
            local_text = torch.from_numpy(np.ascontiguousarray(np.frombuffer(local_text, np.uint8) if isinstance(local_text, (bytes, bytearray))
                                                                 else local_text, np.uint8))
        text = local_text.to(self.device).contiguous()
        self.init_size(text.numel())
        dt = torch.int32 if self.index_bytes == 4 else torch.int64  # raw storage of unsigned indices
        m = self.local_size
        self.local_SA = torch.empty(m, dtype=dt, device=self.device)
        self.local_B = torch.empty(m, dtype=dt, device=self.device)
        self.local_LCP = torch.empty(m, dtype=dt, device=self.device) if self.construct_lcp else None
        flags = (api.LCP if self.construct_lcp else 0) | (api.FAST_RESOLVAL if fast_resolval else 0)
        self._text = text
        # the engine works on its own stream: everything torch queued for the inputs (the copy above, the caller's
        # producer kernels) must have completed before the engine reads them
        torch.cuda.current_stream(self.device).synchronize()
        self.engine.construct_sharded_ptr(text.data_ptr() if m else None, m, self.n, self.index_bytes, flags, k, self.local_SA.data_ptr() if m else None,
                                          self.local_B.data_ptr() if m else None, self.local_LCP.data_ptr() if (m and self.construct_lcp) else None)
        return self

    def check(self):
        """Collective device-side certificate of the last construct (reference d_check_sa + check_lcp); dict, ``ok`` = correct."""
        m = self.local_size
        torch.cuda.current_stream(self.device).synchronize()
        return self.engine.check_sharded_ptr(self._text.data_ptr() if m else None, m, self.n, self.index_bytes, self.local_SA.data_ptr() if m else None,
                                             self.local_B.data_ptr() if m else None,
                                             self.local_LCP.data_ptr() if (m and self.construct_lcp) else None)

    def construct_suffix_tree(self, sigma_plus_one=None):
        """Collective; reference construct_suffix_tree(sa, begin, end, comm) (include/suffix_tree.hpp:413-499) on the blocks of the last
        construct (needs construct_lcp): returns this rank's rows of the child table, an (local_size, sigma + 1) int64 CUDA tensor."""
        if not self.construct_lcp:
            raise api.PsacError("construct_suffix_tree needs the LCP array (construct_lcp=True)")
        m = self.local_size
        if sigma_plus_one is None:
            # distinct characters of the whole text
            h = torch.bincount(self._text.to(torch.int64), minlength=256)
            if dist.get_backend() == "nccl":
                dist.all_reduce(h)
            else:
                hc = h.cpu()
                dist.all_reduce(hc)
                h = hc
            sigma_plus_one = int((h > 0).sum().item()) + 1
        nodes = torch.empty((m, sigma_plus_one), dtype=torch.int64, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self.engine.suffix_tree_sharded_ptr(self._text.data_ptr() if m else None, m, self.n, self.index_bytes, self.local_SA.data_ptr() if m else None,
                                            self.local_LCP.data_ptr() if m else None, nodes.data_ptr() if m else None, nodes.numel())
        return nodes

    def ansv(self, left_type=0, right_type=0, nonsv=0):
        """Collective; reference ansv<index_t, left, right, global_indexing>(local_LCP, ...) (include/ansv.hpp:2042-2051): global indices."""
        m = self.local_size
        left = torch.empty(m, dtype=torch.int64, device=self.device)
        right = torch.empty(m, dtype=torch.int64, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self.engine.ansv_sharded_ptr(self.local_LCP.data_ptr() if m else None, m, self.n, self.index_bytes, left_type, right_type, nonsv,
                                     left.data_ptr() if m else None, right.data_ptr() if m else None)
        return left, right

    def left_branching_chars(self):
        """Collective; the reference's local_Lc (_CONSTRUCT_LC, include/suffix_array.hpp:212): this rank's block, a uint8 CUDA tensor."""
        m = self.local_size
        lc = torch.empty(m, dtype=torch.uint8, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self.engine.lc_sharded_ptr(self._text.data_ptr() if m else None, m, self.n, self.index_bytes, self.local_SA.data_ptr() if m else None,
                                   self.local_LCP.data_ptr() if m else None, lc.data_ptr() if m else None)
        return lc

    def close(self):
        """Collective: ordered release of the peer-visible memory, then the engine."""
        if self.engine is not None:
            self.engine.comm_finalize()
            self.engine.close()
            self.engine = None
