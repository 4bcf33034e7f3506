/*
 * TEST INFRASTRUCTURE ONLY -- single-process stand-in for <mpi.h>.
 *
 * The reference (patflick/psac) is header-only C++ over MPI.  This container
 * and the GPU box have no MPI, so the oracle build compiles the UNMODIFIED
 * reference headers (from /root/reference, never copied) against this shim and
 * runs them at np=1.  Every collective degenerates to a memcpy of
 * count*type_size bytes; point-to-point calls abort (the reference guards all
 * of them with rank/size tests, so they are unreachable at np=1).
 *
 * Written from the MPI-3 standard's prototypes; nothing here is taken from an
 * MPI implementation or from the reference.
 */
#ifndef PSACB200_ORACLE_MPI_SHIM_H
#define PSACB200_ORACLE_MPI_SHIM_H

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <vector>

#define MPI_VERSION 3
#define MPI_SUBVERSION 0

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_Errhandler;
typedef int MPI_Win;
typedef int MPI_Message;
typedef FILE* MPI_File;
typedef std::ptrdiff_t MPI_Aint;
typedef long long MPI_Offset;
typedef long long MPI_Count;

struct MPI_Status {
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    long long shim_count;
};

typedef void(MPI_User_function)(void*, void*, int*, MPI_Datatype*);
typedef void(MPI_Comm_errhandler_function)(MPI_Comm*, int*, ...);
typedef MPI_Comm_errhandler_function MPI_Handler_function;
typedef int(MPI_Type_copy_attr_function)(MPI_Datatype, int, void*, void*, void*, int*);
typedef int(MPI_Type_delete_attr_function)(MPI_Datatype, int, void*, void*);

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_DATATYPE_NULL 0
#define MPI_OP_NULL 0
#define MPI_REQUEST_NULL 0
#define MPI_INFO_NULL 0
#define MPI_FILE_NULL ((MPI_File)0)
#define MPI_WIN_NULL 0
#define MPI_MESSAGE_NULL 0
#define MPI_ERRHANDLER_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)1)
#define MPI_BOTTOM ((void*)0)
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-2)
#define MPI_PROC_NULL (-3)
#define MPI_UNDEFINED (-32766)
#define MPI_MAX_PROCESSOR_NAME 128
#define MPI_MAX_ERROR_STRING 256
#define MPI_COMM_TYPE_SHARED 1
#define MPI_ERRORS_ARE_FATAL 1
#define MPI_ERRORS_RETURN 2
#define MPI_TYPE_NULL_COPY_FN ((MPI_Type_copy_attr_function*)0)
#define MPI_TYPE_NULL_DELETE_FN ((MPI_Type_delete_attr_function*)0)
#define MPI_TYPE_DUP_FN ((MPI_Type_copy_attr_function*)0)

#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_RDWR 4
#define MPI_MODE_CREATE 8
#define MPI_MODE_APPEND 16
#define MPI_MODE_NOSUCCEED 1
#define MPI_MODE_NOSTORE 2
#define MPI_MODE_NOPUT 4
#define MPI_MODE_NOPRECEDE 8

/* builtin datatypes: small fixed handles, sizes pre-filled in the type table */
enum {
    MPI_CHAR = 1, MPI_SIGNED_CHAR, MPI_UNSIGNED_CHAR, MPI_BYTE, MPI_SHORT, MPI_UNSIGNED_SHORT,
    MPI_INT, MPI_UNSIGNED, MPI_LONG, MPI_UNSIGNED_LONG, MPI_LONG_LONG, MPI_UNSIGNED_LONG_LONG,
    MPI_FLOAT, MPI_DOUBLE, MPI_LONG_DOUBLE, MPI_2INT, MPI_SHORT_INT, MPI_LONG_INT, MPI_FLOAT_INT,
    MPI_DOUBLE_INT, MPI_LONG_DOUBLE_INT, MPI_WCHAR, MPI_C_BOOL, MPI_INT8_T, MPI_INT16_T, MPI_INT32_T,
    MPI_INT64_T, MPI_UINT8_T, MPI_UINT16_T, MPI_UINT32_T, MPI_UINT64_T, MPI_AINT, MPI_OFFSET, MPI_COUNT,
    SHIM_FIRST_USER_TYPE
};
#define MPI_LONG_LONG_INT MPI_LONG_LONG

enum { MPI_SUM = 1, MPI_PROD, MPI_MAX, MPI_MIN, MPI_LAND, MPI_LOR, MPI_LXOR, MPI_BAND, MPI_BOR, MPI_BXOR,
       MPI_MAXLOC, MPI_MINLOC, MPI_REPLACE, MPI_NO_OP, SHIM_FIRST_USER_OP };

namespace mpishim {

struct type_rec {
    long long size = 0;        /* bytes of data */
    MPI_Aint lb = 0, extent = 0, true_lb = 0, true_extent = 0;
    std::map<int, void*> attrs;
    bool live = false;
};

struct keyval_rec { MPI_Type_delete_attr_function* del; void* extra; };

struct state {
    std::vector<type_rec> types;
    std::vector<keyval_rec> keyvals;
    int next_comm = 3, next_op = SHIM_FIRST_USER_OP, next_misc = 1;
    bool initialized = false, finalized = false;
    state() {
        types.resize(SHIM_FIRST_USER_TYPE);
        struct S1 { short a; int b; }; struct S2 { long a; int b; }; struct S3 { float a; int b; };
        struct S4 { double a; int b; }; struct S5 { long double a; int b; };
        auto set = [this](int h, size_t s) { type_rec& t = types[h]; t.size = s; t.extent = s; t.true_extent = s; t.live = true; };
        set(MPI_CHAR, 1); set(MPI_SIGNED_CHAR, 1); set(MPI_UNSIGNED_CHAR, 1); set(MPI_BYTE, 1);
        set(MPI_SHORT, sizeof(short)); set(MPI_UNSIGNED_SHORT, sizeof(short));
        set(MPI_INT, sizeof(int)); set(MPI_UNSIGNED, sizeof(unsigned));
        set(MPI_LONG, sizeof(long)); set(MPI_UNSIGNED_LONG, sizeof(long));
        set(MPI_LONG_LONG, sizeof(long long)); set(MPI_UNSIGNED_LONG_LONG, sizeof(long long));
        set(MPI_FLOAT, sizeof(float)); set(MPI_DOUBLE, sizeof(double)); set(MPI_LONG_DOUBLE, sizeof(long double));
        set(MPI_2INT, 2 * sizeof(int)); set(MPI_SHORT_INT, sizeof(S1)); set(MPI_LONG_INT, sizeof(S2));
        set(MPI_FLOAT_INT, sizeof(S3)); set(MPI_DOUBLE_INT, sizeof(S4)); set(MPI_LONG_DOUBLE_INT, sizeof(S5));
        set(MPI_WCHAR, sizeof(wchar_t)); set(MPI_C_BOOL, 1);
        set(MPI_INT8_T, 1); set(MPI_INT16_T, 2); set(MPI_INT32_T, 4); set(MPI_INT64_T, 8);
        set(MPI_UINT8_T, 1); set(MPI_UINT16_T, 2); set(MPI_UINT32_T, 4); set(MPI_UINT64_T, 8);
        set(MPI_AINT, sizeof(MPI_Aint)); set(MPI_OFFSET, sizeof(MPI_Offset)); set(MPI_COUNT, sizeof(MPI_Count));
    }
};

inline state& S() { static state s; return s; }

inline int new_type(const type_rec& r) {
    state& s = S();
    s.types.push_back(r);
    s.types.back().live = true;
    s.types.back().attrs.clear();
    return (int)s.types.size() - 1;
}
inline type_rec& T(MPI_Datatype t) {
    state& s = S();
    if (t <= 0 || (size_t)t >= s.types.size() || !s.types[t].live) {
        std::fprintf(stderr, "[mpi shim] invalid datatype handle %d\n", t);
        std::abort();
    }
    return s.types[t];
}
/* span of `count` elements laid out extent apart (what a collective moves at np=1) */
inline size_t span(long long count, MPI_Datatype t) {
    if (count <= 0) return 0;
    type_rec& r = T(t);
    return (size_t)((count - 1) * (long long)r.extent + (long long)r.true_lb + (long long)r.true_extent);
}
inline void copy(const void* src, void* dst, long long count, MPI_Datatype t) {
    if (src == MPI_IN_PLACE || src == dst || count <= 0) return;
    std::memmove(dst, src, span(count, t));
}
[[noreturn]] inline void unreachable(const char* what) {
    std::fprintf(stderr, "[mpi shim] %s is not available in the single-process shim\n", what);
    std::abort();
}
} // namespace mpishim

/* ---- environment ---- */
static inline int MPI_Init(int*, char***) { mpishim::S().initialized = true; return MPI_SUCCESS; }
static inline int MPI_Finalize() { mpishim::S().finalized = true; return MPI_SUCCESS; }
static inline int MPI_Initialized(int* f) { *f = mpishim::S().initialized; return MPI_SUCCESS; }
static inline int MPI_Finalized(int* f) { *f = mpishim::S().finalized; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm, int code) { std::fprintf(stderr, "[mpi shim] MPI_Abort(%d)\n", code); std::abort(); return 0; }
static inline double MPI_Wtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline int MPI_Get_processor_name(char* name, int* len) { std::strcpy(name, "localhost"); *len = 9; return MPI_SUCCESS; }
static inline int MPI_Error_class(int code, int* cls) { *cls = code; return MPI_SUCCESS; }
static inline int MPI_Error_string(int, char* s, int* len) { std::strcpy(s, "mpi shim error"); *len = (int)std::strlen(s); return MPI_SUCCESS; }
static inline int MPI_Comm_create_errhandler(MPI_Comm_errhandler_function*, MPI_Errhandler* e) { *e = 3; return MPI_SUCCESS; }
static inline int MPI_Comm_set_errhandler(MPI_Comm, MPI_Errhandler) { return MPI_SUCCESS; }
static inline int MPI_Errhandler_create(MPI_Handler_function*, MPI_Errhandler* e) { *e = 3; return MPI_SUCCESS; }
static inline int MPI_Errhandler_set(MPI_Comm, MPI_Errhandler) { return MPI_SUCCESS; }
static inline int MPI_Errhandler_free(MPI_Errhandler* e) { *e = MPI_ERRHANDLER_NULL; return MPI_SUCCESS; }
static inline int MPI_Info_create(MPI_Info* i) { *i = 1; return MPI_SUCCESS; }
static inline int MPI_Info_set(MPI_Info, const char*, const char*) { return MPI_SUCCESS; }
static inline int MPI_Info_free(MPI_Info* i) { *i = MPI_INFO_NULL; return MPI_SUCCESS; }

/* ---- communicators: every communicator has exactly one rank ---- */
static inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_dup(MPI_Comm, MPI_Comm* o) { *o = mpishim::S().next_comm++; return MPI_SUCCESS; }
static inline int MPI_Comm_split(MPI_Comm, int color, int, MPI_Comm* o) {
    *o = (color == MPI_UNDEFINED) ? MPI_COMM_NULL : mpishim::S().next_comm++; return MPI_SUCCESS;
}
static inline int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm* o) { *o = mpishim::S().next_comm++; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return MPI_SUCCESS; }
static inline int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* r) { *r = (a == b) ? 0 : 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }

/* ---- datatypes ---- */
static inline int MPI_Type_size(MPI_Datatype t, int* s) { *s = (int)mpishim::T(t).size; return MPI_SUCCESS; }
static inline int MPI_Type_get_extent(MPI_Datatype t, MPI_Aint* lb, MPI_Aint* ext) {
    *lb = mpishim::T(t).lb; *ext = mpishim::T(t).extent; return MPI_SUCCESS;
}
static inline int MPI_Type_get_true_extent(MPI_Datatype t, MPI_Aint* lb, MPI_Aint* ext) {
    *lb = mpishim::T(t).true_lb; *ext = mpishim::T(t).true_extent; return MPI_SUCCESS;
}
static inline int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype* nt) {
    mpishim::type_rec o = mpishim::T(old), r;
    r.size = o.size * count; r.lb = o.lb; r.extent = o.extent * count;
    r.true_lb = o.true_lb; r.true_extent = count > 0 ? (count - 1) * o.extent + o.true_extent : 0;
    *nt = mpishim::new_type(r); return MPI_SUCCESS;
}
static inline int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype old, MPI_Datatype* nt) {
    mpishim::type_rec o = mpishim::T(old), r;
    r.size = o.size * count * blocklen; r.lb = o.lb;
    long long elems = count > 0 ? (long long)(count - 1) * stride + blocklen : 0;
    r.extent = (MPI_Aint)(elems * o.extent); r.true_lb = o.true_lb;
    r.true_extent = elems > 0 ? (MPI_Aint)((elems - 1) * o.extent + o.true_extent) : 0;
    *nt = mpishim::new_type(r); return MPI_SUCCESS;
}
static inline int MPI_Type_create_struct(int count, const int* blens, const MPI_Aint* displs, const MPI_Datatype* types, MPI_Datatype* nt) {
    mpishim::type_rec r; bool first = true; MPI_Aint lo = 0, hi = 0, tlo = 0, thi = 0;
    for (int i = 0; i < count; ++i) {
        if (blens[i] <= 0) continue;
        mpishim::type_rec o = mpishim::T(types[i]);
        r.size += o.size * blens[i];
        MPI_Aint l = displs[i] + o.lb, h = displs[i] + o.lb + o.extent * blens[i];
        MPI_Aint tl = displs[i] + o.true_lb, th = displs[i] + (blens[i] - 1) * o.extent + o.true_lb + o.true_extent;
        if (first) { lo = l; hi = h; tlo = tl; thi = th; first = false; }
        else { if (l < lo) lo = l; if (h > hi) hi = h; if (tl < tlo) tlo = tl; if (th > thi) thi = th; }
    }
    r.lb = lo; r.extent = hi - lo; r.true_lb = tlo; r.true_extent = thi - tlo;
    *nt = mpishim::new_type(r); return MPI_SUCCESS;
}
static inline int MPI_Type_create_resized(MPI_Datatype old, MPI_Aint lb, MPI_Aint extent, MPI_Datatype* nt) {
    mpishim::type_rec r = mpishim::T(old); r.lb = lb; r.extent = extent;
    *nt = mpishim::new_type(r); return MPI_SUCCESS;
}
static inline int MPI_Type_dup(MPI_Datatype old, MPI_Datatype* nt) {
    mpishim::type_rec r = mpishim::T(old); *nt = mpishim::new_type(r); return MPI_SUCCESS; /* attributes not copied */
}
static inline int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype* t) {
    if (*t >= SHIM_FIRST_USER_TYPE) {
        mpishim::type_rec& r = mpishim::T(*t);
        std::map<int, void*> attrs; attrs.swap(r.attrs);
        for (auto& kv : attrs) {
            mpishim::keyval_rec& k = mpishim::S().keyvals[kv.first];
            if (k.del) k.del(*t, kv.first, kv.second, k.extra);
        }
        mpishim::S().types[*t].live = false;
    }
    *t = MPI_DATATYPE_NULL; return MPI_SUCCESS;
}
static inline int MPI_Get_address(const void* p, MPI_Aint* a) { *a = (MPI_Aint)(reinterpret_cast<std::uintptr_t>(p)); return MPI_SUCCESS; }
static inline int MPI_Type_create_keyval(MPI_Type_copy_attr_function*, MPI_Type_delete_attr_function* del, int* keyval, void* extra) {
    mpishim::S().keyvals.push_back({del, extra}); *keyval = (int)mpishim::S().keyvals.size() - 1; return MPI_SUCCESS;
}
static inline int MPI_Type_set_attr(MPI_Datatype t, int keyval, void* val) { mpishim::T(t).attrs[keyval] = val; return MPI_SUCCESS; }
static inline int MPI_Type_get_attr(MPI_Datatype t, int keyval, void* val_out, int* flag) {
    auto& m = mpishim::T(t).attrs; auto it = m.find(keyval);
    if (it == m.end()) { *flag = 0; } else { *flag = 1; *(void**)val_out = it->second; }
    return MPI_SUCCESS;
}
static inline int MPI_Get_count(const MPI_Status* s, MPI_Datatype, int* c) { *c = s ? (int)s->shim_count : 0; return MPI_SUCCESS; }
static inline int MPI_Get_elements_x(const MPI_Status* s, MPI_Datatype, MPI_Count* c) { *c = s ? s->shim_count : 0; return MPI_SUCCESS; }

/* ---- reduction ops: with one rank every reduction is the identity ---- */
static inline int MPI_Op_create(MPI_User_function*, int, MPI_Op* op) { *op = mpishim::S().next_op++; return MPI_SUCCESS; }
static inline int MPI_Op_free(MPI_Op* op) { *op = MPI_OP_NULL; return MPI_SUCCESS; }

/* ---- collectives ---- */
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { mpishim::copy(s, r, n, t); return MPI_SUCCESS; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { mpishim::copy(s, r, n, t); return MPI_SUCCESS; }
static inline int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { mpishim::copy(s, r, n, t); return MPI_SUCCESS; }
/* rank 0's receive buffer is undefined after an exclusive scan: leave it untouched */
static inline int MPI_Exscan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int, MPI_Datatype, int, MPI_Comm) { mpishim::copy(s, r, sn, st); return MPI_SUCCESS; }
static inline int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int, MPI_Datatype, MPI_Comm) { mpishim::copy(s, r, sn, st); return MPI_SUCCESS; }
static inline int MPI_Scatter(const void* s, int sn, MPI_Datatype st, void* r, int, MPI_Datatype, int, MPI_Comm) {
    if (r != MPI_IN_PLACE) mpishim::copy(s, r, sn, st); return MPI_SUCCESS;
}
static inline int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int*, const int* displs, MPI_Datatype rt, int, MPI_Comm) {
    mpishim::copy(s, (char*)r + (displs ? displs[0] : 0) * mpishim::T(rt).extent, sn, st); return MPI_SUCCESS;
}
static inline int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int*, const int* displs, MPI_Datatype rt, MPI_Comm) {
    mpishim::copy(s, (char*)r + (displs ? displs[0] : 0) * mpishim::T(rt).extent, sn, st); return MPI_SUCCESS;
}
static inline int MPI_Scatterv(const void* s, const int* sn, const int* displs, MPI_Datatype st, void* r, int, MPI_Datatype, int, MPI_Comm) {
    if (r != MPI_IN_PLACE) mpishim::copy((const char*)s + (displs ? displs[0] : 0) * mpishim::T(st).extent, r, sn[0], st); return MPI_SUCCESS;
}
static inline int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int, MPI_Datatype, MPI_Comm) { mpishim::copy(s, r, sn, st); return MPI_SUCCESS; }
static inline int MPI_Alltoallv(const void* s, const int* sn, const int* sd, MPI_Datatype st, void* r, const int*, const int* rd, MPI_Datatype rt, MPI_Comm) {
    mpishim::copy((const char*)s + sd[0] * mpishim::T(st).extent, (char*)r + rd[0] * mpishim::T(rt).extent, sn[0], st); return MPI_SUCCESS;
}
static inline int MPI_Alltoallw(const void* s, const int* sn, const int* sd, const MPI_Datatype* st, void* r, const int*, const int* rd, const MPI_Datatype*, MPI_Comm) {
    mpishim::copy((const char*)s + sd[0], (char*)r + rd[0], sn[0], st[0]); return MPI_SUCCESS;
}

/* ---- point-to-point: unreachable at np=1 ---- */
static inline int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { mpishim::unreachable("MPI_Send"); }
static inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { mpishim::unreachable("MPI_Recv"); }
static inline int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { mpishim::unreachable("MPI_Isend"); }
static inline int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { mpishim::unreachable("MPI_Irecv"); }
static inline int MPI_Sendrecv(const void*, int, MPI_Datatype, int, int, void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { mpishim::unreachable("MPI_Sendrecv"); }
static inline int MPI_Probe(int, int, MPI_Comm, MPI_Status*) { mpishim::unreachable("MPI_Probe"); }
static inline int MPI_Mprobe(int, int, MPI_Comm, MPI_Message*, MPI_Status*) { mpishim::unreachable("MPI_Mprobe"); }
static inline int MPI_Mrecv(void*, int, MPI_Datatype, MPI_Message*, MPI_Status*) { mpishim::unreachable("MPI_Mrecv"); }
static inline int MPI_Wait(MPI_Request* r, MPI_Status*) { if (r) *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status*) { for (int i = 0; i < n; ++i) r[i] = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Test(MPI_Request*, int* flag, MPI_Status*) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Testall(int, MPI_Request*, int* flag, MPI_Status*) { *flag = 1; return MPI_SUCCESS; }

/* ---- one-sided: stubs (the shm/RMA variants of bulk_rma are not on the np=1 path) ---- */
static inline int MPI_Win_create(const void*, MPI_Aint, int, MPI_Info, MPI_Comm, MPI_Win* w) { *w = mpishim::S().next_misc++; return MPI_SUCCESS; }
static inline int MPI_Win_allocate_shared(MPI_Aint size, int, MPI_Info, MPI_Comm, void* base, MPI_Win* w) {
    *(void**)base = std::malloc((size_t)(size > 0 ? size : 1)); *w = mpishim::S().next_misc++; return MPI_SUCCESS;
}
static inline int MPI_Win_shared_query(MPI_Win, int, MPI_Aint*, int*, void*) { mpishim::unreachable("MPI_Win_shared_query"); }
static inline int MPI_Win_fence(int, MPI_Win) { return MPI_SUCCESS; }
static inline int MPI_Win_sync(MPI_Win) { return MPI_SUCCESS; }
static inline int MPI_Win_free(MPI_Win* w) { *w = MPI_WIN_NULL; return MPI_SUCCESS; }
static inline int MPI_Get(void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win) { mpishim::unreachable("MPI_Get"); }

/* ---- file IO over stdio (one rank: "ordered" == sequential) ---- */
static inline int MPI_File_open(MPI_Comm, const char* name, int amode, MPI_Info, MPI_File* fh) {
    const char* mode = "rb";
    if (amode & MPI_MODE_WRONLY) mode = "wb";
    else if (amode & MPI_MODE_RDWR) mode = (amode & MPI_MODE_CREATE) ? "w+b" : "r+b";
    *fh = std::fopen(name, mode);
    return *fh ? MPI_SUCCESS : MPI_ERR_OTHER;
}
static inline int MPI_File_close(MPI_File* fh) { if (*fh) std::fclose(*fh); *fh = MPI_FILE_NULL; return MPI_SUCCESS; }
static inline int MPI_File_delete(const char* name, MPI_Info) { std::remove(name); return MPI_SUCCESS; }
static inline int MPI_File_get_size(MPI_File fh, MPI_Offset* size) {
    long cur = std::ftell(fh); std::fseek(fh, 0, SEEK_END); *size = std::ftell(fh); std::fseek(fh, cur, SEEK_SET); return MPI_SUCCESS;
}
static inline int MPI_File_set_size(MPI_File, MPI_Offset) { return MPI_SUCCESS; }
static inline int MPI_File_write_ordered(MPI_File fh, const void* buf, int n, MPI_Datatype t, MPI_Status*) {
    std::fwrite(buf, 1, mpishim::span(n, t), fh); return MPI_SUCCESS;
}
static inline int MPI_File_read_ordered(MPI_File fh, void* buf, int n, MPI_Datatype t, MPI_Status* st) {
    size_t got = std::fread(buf, 1, mpishim::span(n, t), fh); if (st) st->shim_count = (long long)(got / (size_t)mpishim::T(t).size); return MPI_SUCCESS;
}
static inline int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int n, MPI_Datatype t, MPI_Status*) {
    std::fseek(fh, (long)off, SEEK_SET); std::fwrite(buf, 1, mpishim::span(n, t), fh); return MPI_SUCCESS;
}
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int n, MPI_Datatype t, MPI_Status* st) {
    std::fseek(fh, (long)off, SEEK_SET); size_t got = std::fread(buf, 1, mpishim::span(n, t), fh);
    if (st) st->shim_count = (long long)(got / (size_t)mpishim::T(t).size); return MPI_SUCCESS;
}

#endif /* PSACB200_ORACLE_MPI_SHIM_H */
