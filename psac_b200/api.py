"""ctypes binding of the psacb200 C ABI (include/psacb200.h) and a Python mirror of the reference class.

``SuffixArray`` mirrors ``suffix_array<char_t, index_t, _CONSTRUCT_LCP>`` of the reference
(include/suffix_array.hpp:170-213): same method names (``construct``, ``construct_arr``), same argument meaning
(``fast_resolval``, ``k``) and the same post-conditions (``n``, ``local_size``, ``local_SA``, ``local_B`` = ISA,
``local_LCP``), at p = 1.  There is no CPU fallback: if the CUDA library or a GPU is missing every call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpsacb200.so")

LCP = 1
FAST_RESOLVAL = 2

_lib = None


class PsacError(RuntimeError):
    """Mirrors the reference's std::runtime_error (suffix_array.hpp:226-227)."""


class Stats(C.Structure):
    _fields_ = [
        ("n", C.c_uint64),
        ("sigma", C.c_uint32),
        ("bits_per_char", C.c_uint32),
        ("pack_bits", C.c_uint32),
        ("key_chars", C.c_uint32),
        ("rounds", C.c_uint32),
        ("sort_passes", C.c_uint32),
        ("internal_index_bytes", C.c_uint32),
        ("peer_exchange", C.c_uint32),
        ("unresolved_after_first", C.c_uint64),
        ("device_bytes", C.c_uint64),
        ("ms_total", C.c_float),
        ("ms_h2d", C.c_float),
        ("ms_alphabet", C.c_float),
        ("ms_pack", C.c_float),
        ("ms_keygen", C.c_float),
        ("ms_hist", C.c_float),
        ("ms_sort", C.c_float),
        ("ms_resolve", C.c_float),
        ("ms_rounds", C.c_float),
        ("ms_output", C.c_float),
        ("ms_d2h", C.c_float),
        ("ms_sort_pass_avg", C.c_float),
        ("ms_isa", C.c_float),
        ("ms_sort_pass1", C.c_float),
        ("ms_scatter_avg", C.c_float),
        ("sort_elt_bytes", C.c_uint32),
        ("sharded_scheme", C.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


class CheckReport(C.Structure):
    """psacb200_check_report: verdict of the device-side certificate (reference d_check_sa + check_lcp)."""
    _fields_ = [("n", C.c_uint64), ("bad_range", C.c_uint64), ("bad_inverse", C.c_uint64), ("bad_order", C.c_uint64), ("bad_lcp", C.c_uint64),
                ("first_bad", C.c_uint64), ("ms", C.c_float), ("checked_lcp", C.c_uint32)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["ok"] = self.ok
        return d

    @property
    def ok(self):
        return self.bad_range == 0 and self.bad_inverse == 0 and self.bad_order == 0 and self.bad_lcp == 0


def lib():
    """Load libpsacb200.so (built in-tree by ``make`` / ``__graft_entry__.build``).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PsacError("psac_b200/libpsacb200.so is missing: run `make` (or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.psacb200_last_error.restype = C.c_char_p
        L.psacb200_launch_count.restype = C.c_uint64
        L.psacb200_launch_count.argtypes = [C.c_void_p]
        L.psacb200_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.psacb200_stream.restype = C.c_void_p
        L.psacb200_stream.argtypes = [C.c_void_p]
        L.psacb200_destroy.argtypes = [C.c_void_p]
        L.psacb200_destroy.restype = None
        L.psacb200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.psacb200_reserve.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint]
        L.psacb200_alphabet.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        cargs = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        L.psacb200_construct.argtypes = cargs
        L.psacb200_construct_device.argtypes = cargs
        L.psacb200_construct_alphabet.argtypes = cargs[:6] + [C.c_void_p] + cargs[6:]
        L.psacb200_sort_pairs.argtypes = [C.c_void_p] * 5 + [C.c_size_t] + [C.c_int] * 4
        wargs = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                 C.POINTER(C.c_uint32)]
        L.psacb200_construct_wide.argtypes = wargs
        L.psacb200_construct_wide_device.argtypes = wargs
        ssargs = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint8, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.psacb200_construct_ss.argtypes = ssargs
        L.psacb200_construct_ss_device.argtypes = ssargs
        L.psacb200_sort_pairs_host.argtypes = [C.c_void_p] * 3 + [C.c_size_t] + [C.c_int] * 4
        L.psacb200_comm_unique_id.argtypes = [C.c_void_p]
        L.psacb200_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.psacb200_construct_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                                 C.c_void_p]
        L.psacb200_ansv.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.psacb200_suffix_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.psacb200_comm_finalize.argtypes = [C.c_void_p]
        L.psacb200_trace.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.psacb200_rank_mode.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        chk = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(CheckReport)]
        L.psacb200_check_device.argtypes = chk
        L.psacb200_check.argtypes = chk
        L.psacb200_check_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(CheckReport)]
        L.psacb200_blk_dist.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.psacb200_blk_dist.restype = None
        L.psacb200_choose_splitters.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.psacb200_ansv_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.psacb200_ansv_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.psacb200_suffix_tree_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
        L.psacb200_suffix_tree_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                   C.POINTER(C.c_uint32)]
        L.psacb200_lc.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.psacb200_lc_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.psacb200_lc_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.psacb200_multi_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.psacb200_multi_destroy.argtypes = [C.c_void_p]
        L.psacb200_multi_destroy.restype = None
        L.psacb200_multi_gpus.argtypes = [C.c_void_p]
        L.psacb200_multi_engine.argtypes = [C.c_void_p, C.c_int]
        L.psacb200_multi_engine.restype = C.c_void_p
        L.psacb200_multi_construct.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        L.psacb200_multi_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.psacb200_plan_word_exchange.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64] + [C.c_void_p] * 6 + [C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise PsacError("psacb200 status %d: %s" % (rc, lib().psacb200_last_error().decode()))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class Engine:
    """One GPU, one stream, reusable device buffers (psacb200_create / psacb200_destroy)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().psacb200_create(int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().psacb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().psacb200_launch_count(self._h))

    @property
    def stream_ptr(self):
        """cudaStream_t of the engine (for torch.cuda.ExternalStream / event timing)."""
        return int(lib().psacb200_stream(self._h))

    def stats(self):
        s = Stats()
        _check(lib().psacb200_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def trace(self):
        """Device timeline of the last construct call: list of (label, ms) between consecutive marks."""
        buf = C.create_string_buffer(8192)
        _check(lib().psacb200_trace(self._h, buf, 8192))
        out = []
        for item in buf.value.decode().split(";"):
            if "=" in item:
                k, v = item.split("=")
                out.append((k, float(v)))
        return out

    def reserve(self, n, index_bytes=4, flags=0):
        _check(lib().psacb200_reserve(self._h, n, index_bytes, flags))

    def alphabet(self, text):
        t = _as_text(text)
        lut = np.zeros(256, np.uint8)
        sigma = C.c_uint32()
        bpc = C.c_uint32()
        _check(lib().psacb200_alphabet(self._h, _ptr(t), t.size, _ptr(lut), C.byref(sigma), C.byref(bpc)))
        return lut, sigma.value, bpc.value

    def construct(self, text, index_bytes=8, want_lcp=False, k=0, fast_resolval=True, want_isa=True, lut=None, out=None):
        """Host buffers in, numpy arrays out: dict(sa, isa, lcp)."""
        t = _as_text(text)
        n = t.size
        dt = np.uint32 if index_bytes == 4 else np.uint64
        if out is None:
            sa = np.empty(n, dt)
            isa = np.empty(n, dt) if want_isa else None
            lcp = np.empty(n, dt) if want_lcp else None
        else:
            sa, isa, lcp = out
        flags = (LCP if want_lcp else 0) | (FAST_RESOLVAL if fast_resolval else 0)
        if lut is None:
            _check(lib().psacb200_construct(self._h, _ptr(t), n, index_bytes, flags, k, _ptr(sa), _ptr(isa), _ptr(lcp)))
        else:
            lut = np.ascontiguousarray(lut, np.uint8)
            _check(lib().psacb200_construct_alphabet(self._h, _ptr(t), n, index_bytes, flags, k, _ptr(lut), _ptr(sa), _ptr(isa), _ptr(lcp)))
        return dict(sa=sa, isa=isa, lcp=lcp)

    def construct_wide(self, text, index_bytes=8, want_lcp=False, k=0):
        """Text over 2- or 4-byte characters (numpy int16 / uint16 / int32 / uint32), ordered by value -- reference
        suffix_array<int, ...> (int_alphabet).  Returns dict(sa, isa, lcp, distinct)."""
        t = np.ascontiguousarray(text)
        if t.dtype not in (np.int16, np.uint16, np.int32, np.uint32):
            raise PsacError("construct_wide: characters must be 16- or 32-bit integers")
        n = t.size
        dt = np.uint32 if index_bytes == 4 else np.uint64
        sa, isa = np.empty(n, dt), np.empty(n, dt)
        lcp = np.empty(n, dt) if want_lcp else None
        distinct = np.zeros(255, np.int64)
        nd = C.c_uint32(0)
        _check(lib().psacb200_construct_wide(self._h, _ptr(t), n, t.dtype.itemsize, 1 if t.dtype.kind == "i" else 0, index_bytes,
                                             (LCP if want_lcp else 0) | FAST_RESOLVAL, k, _ptr(sa), _ptr(isa), _ptr(lcp), _ptr(distinct), C.byref(nd)))
        return dict(sa=sa, isa=isa, lcp=lcp, distinct=distinct[: nd.value])

    def construct_ss(self, flat, sep=ord("$"), index_bytes=8, want_lcp=True, want_isa=True, lut=None):
        """Generalized suffix array of the `sep`-separated strings of `flat` (reference construct_ss, suffix_array.hpp:269-363).
        Host buffers in, numpy arrays out: dict(sa, isa, lcp, n); positions index the concatenation without separators."""
        t = _as_text(flat)
        dt = np.uint32 if index_bytes == 4 else np.uint64
        cap = max(t.size, 1)
        sa = np.empty(cap, dt)
        isa = np.empty(cap, dt) if want_isa else None
        lcp = np.empty(cap, dt) if want_lcp else None
        n = C.c_uint64(0)
        if lut is not None:
            lut = np.ascontiguousarray(lut, np.uint8)
        _check(lib().psacb200_construct_ss(self._h, _ptr(t), t.size, C.c_uint8(sep), index_bytes, LCP if want_lcp else 0, _ptr(lut), _ptr(sa),
                                           _ptr(isa), _ptr(lcp), C.byref(n)))
        m = n.value
        return dict(sa=sa[:m], isa=None if isa is None else isa[:m], lcp=None if lcp is None else lcp[:m], n=m)

    def construct_ss_ptr(self, flat_ptr, length, sep, index_bytes, flags, sa_ptr, isa_ptr, lcp_ptr, device=True):
        """Raw-pointer form of construct_ss; returns the number of non-separator characters."""
        n = C.c_uint64(0)
        f = lib().psacb200_construct_ss_device if device else lib().psacb200_construct_ss
        _check(f(self._h, _ptr(flat_ptr), length, C.c_uint8(sep), index_bytes, flags, None, _ptr(sa_ptr), _ptr(isa_ptr), _ptr(lcp_ptr), C.byref(n)))
        return n.value

    def construct_ptr(self, text_ptr, n, index_bytes, flags, k, sa_ptr, isa_ptr, lcp_ptr, device=False):
        """Raw-pointer form (pinned host buffers or device buffers owned by the caller, e.g. torch tensors)."""
        f = lib().psacb200_construct_device if device else lib().psacb200_construct
        _check(f(self._h, _ptr(text_ptr), n, index_bytes, flags, k, _ptr(sa_ptr), _ptr(isa_ptr), _ptr(lcp_ptr)))

    # ---- sharded construction (one rank per GPU)
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id, to be created on rank 0 and handed to every rank (psacb200_comm_unique_id)."""
        buf = np.zeros(128, np.uint8)
        _check(lib().psacb200_comm_unique_id(_ptr(buf)))
        return buf

    def comm_init(self, unique_id, rank, world):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        assert uid.size == 128
        _check(lib().psacb200_comm_init(self._h, _ptr(uid), int(rank), int(world)))

    def construct_sharded_ptr(self, text_ptr, n_local, n_global, index_bytes, flags, k, sa_ptr, isa_ptr, lcp_ptr):
        """Collective; all pointers are DEVICE buffers of this rank's blocks."""
        _check(lib().psacb200_construct_sharded(self._h, _ptr(text_ptr), n_local, n_global, index_bytes, flags, k, _ptr(sa_ptr), _ptr(isa_ptr),
                                                _ptr(lcp_ptr)))

    def comm_finalize(self):
        """Collective: ordered release of the peer-visible memory (call on every rank before close())."""
        _check(lib().psacb200_comm_finalize(self._h))

    # ---- device-side certificate (reference d_check_sa, check_suffix_array.hpp:206-267, + check_lcp)
    def check(self, text, sa, isa, lcp=None):
        """Host arrays in; returns the report as a dict (``ok`` = all conditions hold)."""
        t = _as_text(text)
        sa = np.ascontiguousarray(sa)
        assert sa.dtype in (np.uint32, np.uint64)
        isa = np.ascontiguousarray(isa, sa.dtype)
        lcp = None if lcp is None else np.ascontiguousarray(lcp, sa.dtype)
        rep = CheckReport()
        _check(lib().psacb200_check(self._h, _ptr(t), t.size, sa.dtype.itemsize, _ptr(sa), _ptr(isa), _ptr(lcp), C.byref(rep)))
        return rep.as_dict()

    def check_device_ptr(self, text_ptr, n, index_bytes, sa_ptr, isa_ptr, lcp_ptr=None):
        rep = CheckReport()
        _check(lib().psacb200_check_device(self._h, _ptr(text_ptr), n, index_bytes, _ptr(sa_ptr), _ptr(isa_ptr), _ptr(lcp_ptr), C.byref(rep)))
        return rep.as_dict()

    def check_sharded_ptr(self, text_ptr, n_local, n_global, index_bytes, sa_ptr, isa_ptr, lcp_ptr=None):
        """Collective; DEVICE blocks of this rank."""
        rep = CheckReport()
        _check(lib().psacb200_check_sharded(self._h, _ptr(text_ptr), n_local, n_global, index_bytes, _ptr(sa_ptr), _ptr(isa_ptr), _ptr(lcp_ptr), C.byref(rep)))
        return rep.as_dict()

    # ---- ANSV / suffix tree
    NEAREST_SM, NEAREST_EQ, FURTHEST_EQ = 0, 1, 2

    def ansv(self, vals, left_type=0, right_type=0, nonsv=0):
        """reference ansv<T, left_type, right_type, global_indexing>(in, left, right, comm, nonsv) at p = 1 (ansv.hpp:2042-2051)"""
        v = np.ascontiguousarray(vals)
        assert v.dtype in (np.uint32, np.uint64)
        left = np.empty(v.size, np.uint64)
        right = np.empty(v.size, np.uint64)
        _check(lib().psacb200_ansv(self._h, _ptr(v), v.size, v.dtype.itemsize, left_type, right_type, nonsv, _ptr(left), _ptr(right)))
        return left, right

    def suffix_tree(self, text, sa, lcp):
        """reference construct_suffix_tree(sa, begin, end, comm) at p = 1 (suffix_tree.hpp:413-499): (n, sigma+1) child table"""
        t = _as_text(text)
        sa = np.ascontiguousarray(sa)
        lcp = np.ascontiguousarray(lcp, sa.dtype)
        assert sa.dtype in (np.uint32, np.uint64) and sa.size == t.size == lcp.size
        sigma = len(np.unique(t))
        nodes = np.zeros((t.size, sigma + 1), np.uint64)
        _check(lib().psacb200_suffix_tree(self._h, _ptr(t), t.size, sa.dtype.itemsize, _ptr(sa), _ptr(lcp), _ptr(nodes), nodes.size))
        return nodes

    def lc(self, text, sa, lcp):
        """left-branching characters local_Lc (reference _CONSTRUCT_LC, suffix_array.hpp:212): Lc[i] = text[SA[i-1] + LCP[i]]"""
        t = _as_text(text)
        sa = np.ascontiguousarray(sa)
        lcp = np.ascontiguousarray(lcp, sa.dtype)
        out = np.zeros(t.size, np.uint8)
        _check(lib().psacb200_lc(self._h, _ptr(t), t.size, sa.dtype.itemsize, _ptr(sa), _ptr(lcp), _ptr(out)))
        return out

    def lc_sharded_ptr(self, text_ptr, n_local, n_global, index_bytes, sa_ptr, lcp_ptr, lc_ptr):
        _check(lib().psacb200_lc_sharded(self._h, _ptr(text_ptr), n_local, n_global, index_bytes, _ptr(sa_ptr), _ptr(lcp_ptr), _ptr(lc_ptr)))

    # ---- the same on device pointers (one GPU, or collective over the ranks of comm_init)
    def ansv_device_ptr(self, vals_ptr, n, val_bytes, left_type, right_type, nonsv, left_ptr, right_ptr):
        _check(lib().psacb200_ansv_device(self._h, _ptr(vals_ptr), n, val_bytes, left_type, right_type, nonsv, _ptr(left_ptr), _ptr(right_ptr)))

    def ansv_sharded_ptr(self, vals_ptr, n_local, n_global, val_bytes, left_type, right_type, nonsv, left_ptr, right_ptr):
        _check(lib().psacb200_ansv_sharded(self._h, _ptr(vals_ptr), n_local, n_global, val_bytes, left_type, right_type, nonsv, _ptr(left_ptr), _ptr(right_ptr)))

    def suffix_tree_device_ptr(self, text_ptr, n, index_bytes, sa_ptr, lcp_ptr, nodes_ptr, nodes_len):
        sigma = C.c_uint32()
        _check(lib().psacb200_suffix_tree_device(self._h, _ptr(text_ptr), n, index_bytes, _ptr(sa_ptr), _ptr(lcp_ptr), _ptr(nodes_ptr), nodes_len, C.byref(sigma)))
        return sigma.value

    def suffix_tree_sharded_ptr(self, text_ptr, n_local, n_global, index_bytes, sa_ptr, lcp_ptr, nodes_ptr, nodes_len):
        sigma = C.c_uint32()
        _check(lib().psacb200_suffix_tree_sharded(self._h, _ptr(text_ptr), n_local, n_global, index_bytes, _ptr(sa_ptr), _ptr(lcp_ptr), _ptr(nodes_ptr), nodes_len,
                                                  C.byref(sigma)))
        return sigma.value

    def sort_pairs_host(self, keys, vals, begin_bit, end_bit):
        """In-place stable radix sort of numpy keys (uint32/uint64) and optional values by key bits [begin_bit, end_bit)."""
        assert keys.flags.c_contiguous and keys.dtype in (np.uint32, np.uint64)
        vb = 0 if vals is None else vals.dtype.itemsize
        _check(lib().psacb200_sort_pairs_host(self._h, _ptr(keys), _ptr(vals), keys.size, keys.dtype.itemsize, vb, begin_bit, end_bit))


class MultiEngine:
    """Several GPUs of one box behind one host call (psacb200_multi_*): whole host arrays in and out, sharded internally."""

    def __init__(self, n_gpus, dev_ids=None):
        self._h = C.c_void_p()
        ids = None if dev_ids is None else np.ascontiguousarray(dev_ids, np.int32)
        _check(lib().psacb200_multi_create(int(n_gpus), _ptr(ids), C.byref(self._h)))
        self.n_gpus = int(n_gpus)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().psacb200_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        s = Stats()
        _check(lib().psacb200_multi_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def construct(self, text, index_bytes=8, want_lcp=False, k=0, out=None):
        t = _as_text(text)
        n = t.size
        dt = np.uint32 if index_bytes == 4 else np.uint64
        if out is None:
            sa, isa, lcp = np.empty(n, dt), np.empty(n, dt), (np.empty(n, dt) if want_lcp else None)
        else:
            sa, isa, lcp = out
        flags = (LCP if want_lcp else 0) | FAST_RESOLVAL
        _check(lib().psacb200_multi_construct(self._h, _ptr(t), n, index_bytes, flags, k, _ptr(sa), _ptr(isa), _ptr(lcp)))
        return dict(sa=sa, isa=isa, lcp=lcp)

    def construct_ptr(self, text_ptr, n, index_bytes, flags, k, sa_ptr, isa_ptr, lcp_ptr):
        """Raw (pinned) host pointers."""
        _check(lib().psacb200_multi_construct(self._h, _ptr(text_ptr), n, index_bytes, flags, k, _ptr(sa_ptr), _ptr(isa_ptr), _ptr(lcp_ptr)))


def _as_text(t):
    if isinstance(t, (bytes, bytearray)):
        t = np.frombuffer(bytes(t), dtype=np.uint8)
    return np.ascontiguousarray(t, dtype=np.uint8)


_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_engine


class SuffixArray:
    """Python mirror of ``suffix_array<char, index_t, _CONSTRUCT_LCP>`` at p = 1 (reference suffix_array.hpp:170-213)."""

    def __init__(self, index_bytes=8, construct_lcp=False, engine=None, construct_lc=False):
        self.index_bytes = index_bytes
        self.construct_lcp = construct_lcp or construct_lc
        self.construct_lc = construct_lc  # reference template parameter _CONSTRUCT_LC: also fill local_Lc
        self.local_Lc = None
        self.engine = engine
        self.n = 0
        self.local_size = 0
        self.p = 1
        self.local_SA = None
        self.local_B = None
        self.local_LCP = None
        self.alpha = None
        self._text = None

    def _eng(self):
        return self.engine if self.engine is not None else default_engine()

    def construct(self, text, fast_resolval=True, k=0):
        """reference: construct(begin, end, fast_resolval = true, k = 0), suffix_array.hpp:469-486"""
        t = _as_text(text)
        r = self._eng().construct(t, self.index_bytes, self.construct_lcp, k, fast_resolval)
        self._text = t
        self.n = self.local_size = t.size
        self.local_SA, self.local_B, self.local_LCP = r["sa"], r["isa"], r["lcp"]
        if self.construct_lc:
            self.local_Lc = self._eng().lc(t, self.local_SA, self.local_LCP)
        return self

    def construct_ss(self, flat, sep=ord("$"), lut=None):
        """reference: construct_ss(simple_dstringset&, alphabet), suffix_array.hpp:269-363 -- generalized SA (+ LCP when the
        object was created with construct_lcp) of the `sep`-separated strings; n = characters that are not separators"""
        t = _as_text(flat)
        r = self._eng().construct_ss(t, sep, self.index_bytes, self.construct_lcp, True, lut)
        self._text = t[t != sep]
        self.n = self.local_size = r["n"]
        self.local_SA, self.local_B, self.local_LCP = r["sa"], r["isa"], r["lcp"]
        return self

    def write(self, basename, text=None):
        """reference: write(basename), suffix_array.hpp:232-243 (.alpha needs the text or a previous construct)"""
        from . import fileio
        fileio.write_suffix_array(basename, self.local_SA, self.local_LCP if self.construct_lcp else None, text if text is not None else self._text)

    def read(self, basename):
        """reference: read(basename), suffix_array.hpp:245-265"""
        from . import fileio
        r = fileio.read_suffix_array(basename, self.index_bytes, with_lcp=self.construct_lcp)
        self.local_SA, self.local_LCP, self.alpha = r["sa"], r["lcp"], r["lut"]
        self.local_B = None
        self.n = self.local_size = int(r["n"])
        return self

    def construct_arr(self, text, L=2, fast_resolval=True):
        """reference: construct_arr<L>(begin, end, fast_resolval), suffix_array.hpp:490-641 -- SA/ISA only, no LCP (:555-567)"""
        if L < 2:
            raise PsacError("construct_arr: L must be >= 2")
        t = _as_text(text)
        r = self._eng().construct(t, self.index_bytes, False, 0, fast_resolval)
        self.n = self.local_size = t.size
        self.local_SA, self.local_B, self.local_LCP = r["sa"], r["isa"], None
        return self


# ---------------------------------------------------------------------------------------------- host-side plans (no GPU)
def blk_dist(n, p, r):
    """(start, size) of rank r's block: mxx::blk_dist (reference ext/mxx/include/mxx/partition.hpp:283-331)."""
    a, b = C.c_uint64(), C.c_uint64()
    lib().psacb200_blk_dist(int(n), int(p), int(r), C.byref(a), C.byref(b))
    return a.value, b.value


def choose_splitters(hist, n, p):
    """Key-range ownership over a key-prefix histogram: returns (first[p+1], count[p]); rank r sorts bins [first[r], first[r+1])."""
    h = np.ascontiguousarray(hist, np.uint64)
    first = np.zeros(p + 1, np.uint64)
    count = np.zeros(p, np.uint64)
    _check(lib().psacb200_choose_splitters(_ptr(h), h.size, int(n), int(p), _ptr(first), _ptr(count)))
    return first, count


def plan_word_exchange(cnt, n, pad_tile):
    """Host plan of the exchange fused into digit pass 1 of the sharded sort (psacb200_plan_word_exchange); cnt: (p, nb) counts."""
    c = np.ascontiguousarray(cnt, np.uint64)
    p, nb = c.shape
    first = np.zeros(p + 1, np.uint64)
    cnt_key = np.zeros(p, np.uint64)
    owner = np.zeros(nb, np.int32)
    seg_dense = np.zeros((p, 257), np.uint64)
    seg_pad = np.zeros((p, 257), np.uint64)
    run_off = np.zeros((p, nb), np.uint64)
    bal = C.c_int()
    _check(lib().psacb200_plan_word_exchange(_ptr(c), p, nb, int(n), int(pad_tile), _ptr(first), _ptr(cnt_key), _ptr(owner), _ptr(seg_dense), _ptr(seg_pad),
                                             _ptr(run_off), C.byref(bal)))
    return dict(first=first, cnt_key=cnt_key, owner=owner, seg_dense=seg_dense, seg_pad=seg_pad, run_off=run_off, balanced=bool(bal.value))
