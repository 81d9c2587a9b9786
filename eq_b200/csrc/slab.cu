// Row-slab decomposition plumbing: halo exchange and scalar all-reduce between the GPUs of one box
// (NVLink 5 / NVSwitch).  The reference's precedent is the PETSc DMDA star-stencil halo of
// diffuclass.cpp:364-370,659-669 and the KSP's internal dot-product MPI_Allreduce (diffuclass.cpp:410 [ext]).
// Two transports: peer memory (the default: own kernels over CUDA IPC mappings, second half of this file) and NCCL
// (set-up, long reductions, fallback).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy torch has already loaded in the
// calling process, or the system one), so libeqgpu.so itself has no link-time dependency on it and
// single-GPU use never touches it.
#include "eqgpu_internal.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };

static struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;   // optional (peer set-up)
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

static bool nccl_load(std::string &err)
{
    if (g_nccl.lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);  // already in the process (torch)?
        if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name)                                                     \
    *(void **)(&g_nccl.field) = dlsym(h, name);                              \
    if (!g_nccl.field) { err = std::string("NCCL symbol missing: ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    *(void **)(&g_nccl.AllGather) = dlsym(h, "ncclAllGather");
    g_nccl.lib = h;
    return true;
}

#define EQ_NCCL(call)                                                                       \
    do {                                                                                    \
        ncclResult_t r__ = (call);                                                          \
        if (r__ != 0) {                                                                     \
            s->set_error(std::string(#call) + ": " + g_nccl.GetErrorString(r__));           \
            return EQGPU_ECUDA;                                                             \
        }                                                                                   \
    } while (0)

int slab_unique_id(void *out128)
{
    std::string err;
    if (!nccl_load(err)) return EQGPU_ECUDA;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return EQGPU_ECUDA;
    memcpy(out128, &id, sizeof id);
    return 0;
}

int slab_init_comm(eqgpu_solver *s, const void *unique_id)
{
    std::string err;
    if (!nccl_load(err)) { s->set_error(err); return EQGPU_ECUDA; }
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    ncclComm_t comm = nullptr;
    EQ_NCCL(g_nccl.CommInitRank(&comm, s->slab_world, id, s->slab_rank));
    s->nccl_comm = comm;
    return 0;
}

void slab_destroy_comm(eqgpu_solver *s)
{
    if (s->nccl_comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)s->nccl_comm);
    s->nccl_comm = nullptr;
}

static bool peer_add(eqgpu_solver *s, const LevelDev &L, double *v, int depth);
static int peer_flush(eqgpu_solver *s);

// Refresh `depth` halo rows of a level vector (local view L): my first/last `depth` owned rows go to the
// neighbours below/above, theirs arrive in my halo rows.  Stream-ordered; every rank issues the same sequence.
int slab_exchange2(eqgpu_solver *s, const LevelDev &L, double *v1, double *v2, int depth)
{
    ncclComm_t comm = (ncclComm_t)s->nccl_comm;
    const bool below = L.own0 > 0, above = L.own1 < L.ny;
    if (!below && !above) return 0;
    if (s->peer_ok) {   // peer-memory exchange (a vector too long for the staging buffers goes through NCCL below)
        if (v1 && peer_add(s, L, v1, depth)) v1 = nullptr;
        if (v2 && peer_add(s, L, v2, depth)) v2 = nullptr;
        if (s->slab_group_depth == 0) {   // not inside a bracket: the exchange kernel goes out now
            int rc = peer_flush(s);
            if (rc) return rc;
        }
        if (!v1 && !v2) { if (s->slab_group_depth == 0) s->comm_exchange_groups++; return 0; }
    }
    const size_t nx = (size_t)L.nx, cnt = nx * depth;
    EQ_NCCL(g_nccl.GroupStart());
    // an error inside the group still closes it (an open group would swallow every later NCCL call of the process)
    ncclResult_t bad = (ncclResult_t)0;
    auto note = [&](ncclResult_t r) { if (r != 0 && bad == 0) bad = r; };
    double *vs[2] = {v1, v2};
    for (int q = 0; q < 2; ++q) {
        double *v = vs[q];
        if (!v) continue;
        if (below) {
            note(g_nccl.Send(v + (size_t)L.own0 * nx, cnt, ncclFloat64, s->slab_rank - 1, comm, s->stream));
            note(g_nccl.Recv(v + (size_t)(L.own0 - depth) * nx, cnt, ncclFloat64, s->slab_rank - 1, comm, s->stream));
            s->comm_halo_bytes += (long long)cnt * 8;
        }
        if (above) {
            note(g_nccl.Send(v + (size_t)(L.own1 - depth) * nx, cnt, ncclFloat64, s->slab_rank + 1, comm, s->stream));
            note(g_nccl.Recv(v + (size_t)L.own1 * nx, cnt, ncclFloat64, s->slab_rank + 1, comm, s->stream));
            s->comm_halo_bytes += (long long)cnt * 8;
        }
    }
    note(g_nccl.GroupEnd());
    if (s->slab_group_depth == 0) { s->comm_exchange_groups++; solver_trace_mark(s->stream, "nccl-xch"); }
    if (bad != 0) {
        s->set_error(std::string("halo exchange: ") + g_nccl.GetErrorString(bad));
        return EQGPU_ECUDA;
    }
    return 0;
}

int slab_exchange(eqgpu_solver *s, const LevelDev &L, double *v, int depth) { return slab_exchange2(s, L, v, nullptr, depth); }

// ---------------------------------------------------------------------------------------------------------------------
// Peer-memory halos and scalar all-reduce (the default transport between the GPUs of one box).
// Measured with NCCL (2 x B200, 16384 x 4096, events between all launches): a send/recv halo exchange 15-28 us, a
// one-double all-reduce 11-18 us, twelve plus three of them per PCG iteration.  Two peer versions came first and are
// gone: PULLING the neighbours' rows (25-35 us: three remote round trips in sequence -- poll, load, poll) and pushing
// rows, then a fence, then a flag (20-23 us: the fence waits for the acknowledgement of every posted store).  This one
// only ever WRITES to remote memory, never fences and only ever POLLS local memory:
//   * every rank maps the other ranks' flag blocks and staging buffers once (CUDA IPC over NVLink / NVSwitch);
//   * a halo exchange of up to PEER_MAX_JOBS vectors is ONE kernel on the solver's stream (k_halo_push):
//       1. store my boundary rows into the neighbours' staging buffers (parity q & 1) as 16-byte slots
//          (low word, q, high word, q) -- every aligned 8-byte half names the exchange it belongs to;
//       2. read MY staging buffers until every slot shows q in both halves and copy the values into my halo rows.
//     Staging parity q is free again when exchange q comes round: a rank that starts exchange q has completed q-1, so it
//     has received its neighbours' rows of q-1, which they sent after finishing their kernel q-2 -- the last reader of
//     that parity.  Nothing but my own stream ever writes my vectors, so there is no hazard on the halo rows themselves.
//   * an all-reduce of up to 8 doubles is one warp (k_peer_allreduce): lane r stores my partials into rank r's slots as
//     16-byte (value, tag) pairs, tag = q ^ bits(value) ^ salt -- a torn or stale pair fails the check and is read
//     again --, then every lane polls the LOCAL slot of one rank and the partials are added up in RANK ORDER, so all
//     ranks hold bit-identical sums (each derives the converged flag from them on its own).
// Both kernels carry the programmatic-serialization attribute (their launch overlaps the tail of the kernel before them).
// q counts exchanges (all-reduces) and is the same on every rank because all ranks issue the same sequence of calls.
// Every wait carries a time-out that raises an error flag in mapped host memory, checked after the step's stream
// synchronisation, instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------------------------------
enum { PEER_MAX_JOBS = 4, PEER_F_SLOTS = 8, PEER_AR_MAX = 8, PEER_MAX_WORLD = 16,
       PEER_FLAG_WORDS = PEER_F_SLOTS + PEER_MAX_WORLD * 2 * PEER_AR_MAX * 2, PEER_PUSH_BLOCKS = 148 };

struct PushJob {
    const double *snd_lo, *snd_hi;      // my first / last `depth` owned rows (go to rank-1 / rank+1); null: no neighbour there
    double *rcv_lo, *rcv_hi;            // my halo rows below / above my owned rows
    unsigned long long cnt, off;        // doubles per side; offset of this job inside a staging buffer (even)
};
struct PushBatch {
    int n;
    PushJob j[PEER_MAX_JOBS];
};

__device__ __forceinline__ unsigned long long peer_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool peer_wait(const volatile unsigned long long *flag, unsigned long long q, long long timeout_ns, int *err)
{
    if (*flag >= q) return true;
    const unsigned long long t0 = peer_now();
    for (unsigned spin = 0;; ++spin) {
        if (*flag >= q) return true;
        if ((spin & 63u) == 63u && (long long)(peer_now() - t0) > timeout_ns) break;
    }
    *(volatile int *)err = 1;
    return false;
}

// Staging holds 16 bytes per double: (low word, q, high word, q) as 32-bit values -- each 8-byte half carries its own
// copy of the exchange number, and an aligned 8-byte store is atomic, so the receiver needs no separate flag and the
// sender no fence: data that has arrived says so itself (the scheme of NCCL's LL protocol).
__device__ __forceinline__ void ll_store(uint4 *dst, double v, unsigned q32)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"((unsigned)b), "r"(q32), "r"((unsigned)(b >> 32)), "r"(q32)
                 : "memory");
}
__device__ __forceinline__ bool ll_try(const uint4 *src, unsigned q32, double &v)
{
    unsigned a, f1, c, f2;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(f1), "=r"(c), "=r"(f2) : "l"(src) : "memory");
    v = __longlong_as_double((long long)(((unsigned long long)c << 32) | a));
    return f1 == q32 && f2 == q32;
}

// my rows -> a neighbour's staging buffer; this block's share (part of parts) of cnt doubles
__device__ __forceinline__ void peer_send(uint4 *dst, const double *__restrict__ src, unsigned long long cnt, unsigned part, unsigned parts,
                                          unsigned q32)
{
    const unsigned tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long per = (cnt + parts - 1) / parts, e0 = per * part, e1 = e0 + per < cnt ? e0 + per : cnt;
    for (unsigned long long e = e0 + tid; e < e1; e += 4ull * nt) {
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e + (unsigned long long)k * nt < e1) v[k] = __ldcg(src + e + (unsigned long long)k * nt);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e + (unsigned long long)k * nt < e1) ll_store(dst + e + (unsigned long long)k * nt, v[k], q32);
    }
}
// my staging buffer -> my halo rows, each value as soon as it has landed; false: timed out
__device__ __forceinline__ bool peer_recv(double *__restrict__ dst, const uint4 *src, unsigned long long cnt, unsigned part, unsigned parts,
                                          unsigned q32, long long timeout_ns)
{
    const unsigned tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long per = (cnt + parts - 1) / parts, e0 = per * part, e1 = e0 + per < cnt ? e0 + per : cnt;
    unsigned long long t0 = 0;
    for (unsigned long long e = e0 + tid; e < e1; e += 4ull * nt) {
        double v[4];
        bool have[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) have[k] = !(e + (unsigned long long)k * nt < e1);
        for (unsigned spin = 0;; ++spin) {
            bool all = true;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (!have[k]) { have[k] = ll_try(src + e + (unsigned long long)k * nt, q32, v[k]); all = all && have[k]; }
            if (all) break;
            if ((spin & 63u) == 63u) {
                if (t0 == 0) t0 = peer_now();
                else if ((long long)(peer_now() - t0) > timeout_ns) return false;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e + (unsigned long long)k * nt < e1) dst[e + (unsigned long long)k * nt] = v[k];
    }
    return true;
}

// The (job, side) pairs of a batch are spread over the grid: with at least as many blocks as pairs every pair gets an equal
// share of the blocks, otherwise block b takes pairs b, b + grid, ...
template <class F>
__device__ __forceinline__ void peer_for_my_parts(const PushBatch &B, bool lo, bool hi, F &&f)
{
    const int sides = B.n * ((lo ? 1 : 0) + (hi ? 1 : 0));
    if (sides == 0) return;
    const unsigned per = gridDim.x / sides > 0 ? gridDim.x / sides : 1;
    int sidx = 0;
    for (int k = 0; k < B.n; ++k)
        for (int side = 0; side < 2; ++side) {
            if (!(side == 0 ? lo : hi)) continue;
            if (gridDim.x >= (unsigned)sides) {
                const unsigned b0 = (unsigned)sidx * per;
                if (blockIdx.x >= b0 && blockIdx.x < b0 + per) f(B.j[k], side, blockIdx.x - b0, per);
            } else if ((unsigned)sidx % gridDim.x == blockIdx.x) {
                f(B.j[k], side, 0u, 1u);
            }
            ++sidx;
        }
}

// stage_*: staging buffers, [parity][cap] 16-byte slots each.  my_from_lo / my_from_hi are mine (written by rank-1 / rank+1);
// lo_from_hi is rank-1's buffer for what its upper neighbour (me) sends, hi_from_lo rank+1's for its lower neighbour (me).
// Launched with the programmatic-serialization attribute: its launch overlaps the tail of the kernel before it.
__global__ void __launch_bounds__(256)
k_halo_push(PushBatch B, uint4 *lo_from_hi, uint4 *hi_from_lo, const uint4 *my_from_lo, const uint4 *my_from_hi,
            unsigned long long cap, unsigned long long q, long long timeout_ns, int *err)
{
    pdl_trigger();
    pdl_wait();   // the kernels that produced my boundary rows (and read my halo rows) have completed
    const bool has_lo = lo_from_hi != nullptr, has_hi = hi_from_lo != nullptr;
    const unsigned long long par = (q & 1ull) * cap;
    const unsigned q32 = (unsigned)q;
    // 1. my boundary rows into the neighbours' staging buffers (posted stores; nobody waits for them here)
    peer_for_my_parts(B, has_lo, has_hi, [&](const PushJob &J, int side, unsigned part, unsigned parts) {
        if (side == 0) peer_send(lo_from_hi + par + J.off, J.snd_lo, J.cnt, part, parts, q32);
        else peer_send(hi_from_lo + par + J.off, J.snd_hi, J.cnt, part, parts, q32);
    });
    // 2. the neighbours' rows out of my staging buffers as they land
    bool good = true;
    peer_for_my_parts(B, has_lo, has_hi, [&](const PushJob &J, int side, unsigned part, unsigned parts) {
        if (side == 0) good = peer_recv(J.rcv_lo, my_from_lo + par + J.off, J.cnt, part, parts, q32, timeout_ns) && good;
        else good = peer_recv(J.rcv_hi, my_from_hi + par + J.off, J.cnt, part, parts, q32, timeout_ns) && good;
    });
    if (!good) *(volatile int *)err = 1;
}

struct PeerFlagPtrs { unsigned long long *p[PEER_MAX_WORLD]; };

__device__ __forceinline__ unsigned long long peer_salt(int k) { return 0x9E3779B97F4A7C15ull * (unsigned long long)(k + 1); }

// dst[k] = sum over ranks of src[k], k < count <= PEER_AR_MAX; one warp.  Rank r's block holds slots[sender][parity][k] as
// (value bits, tag) pairs.  Parity q & 1 of my slots is free again at all-reduce q: whoever writes it has finished q-1, so
// it has seen my pair of q-1, which I stored after finishing q-2 -- my last read of that parity.
__global__ void __launch_bounds__(32)
k_peer_allreduce(const double *src, double *dst, int count, PeerFlagPtrs F, int rank, int world, unsigned long long q,
                 long long timeout_ns, int *err)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x;
    const unsigned long long par = q & 1ull;
    unsigned long long mybits[PEER_AR_MAX];
#pragma unroll
    for (int k = 0; k < PEER_AR_MAX; ++k) mybits[k] = k < count ? (unsigned long long)__double_as_longlong(src[k]) : 0ull;
    if (lane < world) {
        unsigned long long *slot = F.p[lane] + PEER_F_SLOTS + ((unsigned long long)rank * 2 + par) * PEER_AR_MAX * 2;
#pragma unroll
        for (int k = 0; k < PEER_AR_MAX; ++k)
            if (k < count) {
                const unsigned long long tag = q ^ mybits[k] ^ peer_salt(k);
                asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(slot + 2 * k), "l"(mybits[k]), "l"(tag) : "memory");
            }
    }
    double v[PEER_AR_MAX];
#pragma unroll
    for (int k = 0; k < PEER_AR_MAX; ++k) v[k] = 0.0;
    bool good = true;
    if (lane < world) {
        const unsigned long long *slot = F.p[rank] + PEER_F_SLOTS + ((unsigned long long)lane * 2 + par) * PEER_AR_MAX * 2;
        const unsigned long long t0 = peer_now();
#pragma unroll
        for (int k = 0; k < PEER_AR_MAX; ++k)
            if (k < count) {
                for (unsigned spin = 0;; ++spin) {
                    unsigned long long a, b;
                    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(slot + 2 * k) : "memory");
                    if ((a ^ b ^ peer_salt(k)) == q) { v[k] = __longlong_as_double((long long)a); break; }
                    if ((spin & 63u) == 63u && (long long)(peer_now() - t0) > timeout_ns) { good = false; *(volatile int *)err = 1; break; }
                }
            }
    }
    good = __all_sync(0xffffffffu, good);
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < PEER_AR_MAX; ++k)
        for (int r = 0; r < world; ++r) {
            const double t = __shfl_sync(0xffffffffu, v[k], r);
            if (lane == k) sum += t;
        }
    if (good && lane < count) dst[lane] = sum;
}

struct PeerEntry {          // what a rank tells the others about one of its allocations
    cudaIpcMemHandle_t handle;
    unsigned long long offset;
    int live, pad;
};

int slab_peer_setup(eqgpu_solver *s)
{
    s->peer_ok = false;
    const char *e = getenv("EQGPU_SLAB_PEER");
    if (!s->slab || s->slab_world < 2 || s->slab_world > PEER_MAX_WORLD || (e && atoi(e) == 0) || !g_nccl.AllGather) return 0;
    if (const char *t = getenv("EQGPU_PEER_TIMEOUT_MS")) s->peer_timeout_ns = std::max(1LL, atoll(t)) * 1000000LL;
    const int W = s->slab_world, R = s->slab_rank;
    // staging: room for PEER_MAX_JOBS vectors of 8 halo rows of the finest level, per side and parity
    s->peer_stage_cap = (unsigned long long)PEER_MAX_JOBS * (((unsigned long long)s->levels[0].dev.nx * 8 + 1) & ~1ull);
    EQ_CUDA(cudaMalloc(&s->peer_flags, sizeof(unsigned long long) * PEER_FLAG_WORDS));
    EQ_CUDA(cudaMemset(s->peer_flags, 0, sizeof(unsigned long long) * PEER_FLAG_WORDS));
    EQ_CUDA(cudaMalloc(&s->peer_stage, sizeof(uint4) * 4 * s->peer_stage_cap));   // [from_lo | from_hi][parity][cap] 16-byte slots
    EQ_CUDA(cudaMemset(s->peer_stage, 0, sizeof(uint4) * 4 * s->peer_stage_cap));    // exchange numbers start at 1
    EQ_CUDA(cudaHostAlloc(&s->peer_err, sizeof(int), cudaHostAllocMapped));
    *s->peer_err = 0;
    // (the driver API is reached through the runtime, as for the tensor maps: libeqgpu.so does not link libcuda)
    typedef CUresult (*AddressRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
    AddressRangeFn address_range = nullptr;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            address_range = (AddressRangeFn)fn;
        (void)cudaGetLastError();
    }
    auto describe = [&](void *ptr, PeerEntry &pe) -> bool {
        memset(&pe, 0, sizeof pe);
        CUdeviceptr base = 0;
        size_t size = 0;
        if (!ptr || !address_range || address_range(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) return false;
        if (cudaIpcGetMemHandle(&pe.handle, (void *)base) != cudaSuccess) return false;
        pe.offset = (unsigned long long)((CUdeviceptr)ptr - base);
        pe.live = 1;
        return true;
    };
    PeerEntry tab[2];
    bool good = describe(s->peer_flags, tab[0]);
    good = describe(s->peer_stage, tab[1]) && good;
    (void)cudaGetLastError();
    if (!good) tab[0].live = tab[1].live = 0;   // a rank that cannot export its memory says so; then nobody uses the peer path
    // every rank's two entries to every rank (NCCL, once)
    std::vector<PeerEntry> all((size_t)2 * W);
    char *d_tab = nullptr, *d_all = nullptr;
    EQ_CUDA(cudaMalloc(&d_tab, sizeof tab));
    EQ_CUDA(cudaMalloc(&d_all, sizeof tab * W));
    EQ_CUDA(cudaMemcpyAsync(d_tab, tab, sizeof tab, cudaMemcpyHostToDevice, s->stream));
    EQ_NCCL(g_nccl.AllGather(d_tab, d_all, sizeof tab, 0 /* ncclInt8 */, (ncclComm_t)s->nccl_comm, s->stream));
    EQ_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof tab * W, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(d_tab); cudaFree(d_all);
    double agree = good ? 0.0 : 1.0;
    auto map_entry = [&](const PeerEntry &pe) -> void * {
        if (!pe.live) { agree = 1.0; return nullptr; }
        void *base = nullptr;
        if (cudaIpcOpenMemHandle(&base, pe.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); agree = 1.0; return nullptr; }
        s->peer_opened.push_back(base);
        return (char *)base + pe.offset;
    };
    for (int r = 0; r < W; ++r)
        s->peer_flags_of[r] = r == R ? s->peer_flags : (unsigned long long *)map_entry(all[2 * r]);
    s->peer_stage_lo = R > 0 ? map_entry(all[2 * (R - 1) + 1]) : nullptr;
    s->peer_stage_hi = R + 1 < W ? map_entry(all[2 * (R + 1) + 1]) : nullptr;
    // everyone must agree to use the peer path
    double *d_agree = nullptr, total = 1.0;
    EQ_CUDA(cudaMalloc(&d_agree, sizeof(double) * 2));
    EQ_CUDA(cudaMemcpyAsync(d_agree, &agree, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    EQ_NCCL(g_nccl.AllReduce(d_agree, d_agree + 1, 1, ncclFloat64, ncclSum, (ncclComm_t)s->nccl_comm, s->stream));
    EQ_CUDA(cudaMemcpyAsync(&total, d_agree + 1, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(d_agree);
    s->peer_ok = total == 0.0;
    if (s->peer_ok) {   // no NCCL kernel between the solver's kernels any more: programmatic dependent launches as on one GPU
        s->pdl = true;
        if (const char *e2 = getenv("EQGPU_PDL")) s->pdl = atoi(e2) != 0;
    }
    s->peer_xseq = s->peer_arseq = 0;
    if (!s->peer_batch) s->peer_batch = new PushBatch();
    ((PushBatch *)s->peer_batch)->n = 0;
    s->peer_batch_fill = 0;
    return 0;
}

void slab_peer_teardown(eqgpu_solver *s)
{
    for (void *b : s->peer_opened) cudaIpcCloseMemHandle(b);
    s->peer_opened.clear();
    // my flag block and staging buffers are NOT freed: the other ranks still hold them mapped (they close their mappings in
    // their own teardown, whenever that runs), and freeing exported memory that is mapped elsewhere is undefined; they stay
    // allocated until the process ends (a few MB per slab solver)
    s->peer_flags = nullptr;
    s->peer_stage = s->peer_stage_lo = s->peer_stage_hi = nullptr;
    for (auto &f : s->peer_flags_of) f = nullptr;
    if (s->peer_err) { cudaFreeHost(s->peer_err); s->peer_err = nullptr; }
    delete (PushBatch *)s->peer_batch; s->peer_batch = nullptr;
    s->peer_ok = false;
    (void)cudaGetLastError();
}

int slab_peer_check(eqgpu_solver *s)
{
    if (!s->peer_err || *(volatile int *)s->peer_err == 0) return 0;
    *s->peer_err = 0;
    s->peer_ok = false;   // what follows goes through NCCL (every rank that timed out does the same; a solver in this state should be rebuilt)
    s->set_error("peer-memory halo / all-reduce: a wait for another rank's data timed out (EQGPU_PEER_TIMEOUT_MS)");
    return EQGPU_ECUDA;
}

static int peer_flush(eqgpu_solver *s)
{
    PushBatch &B = *(PushBatch *)s->peer_batch;
    if (B.n == 0) return 0;
    const int R = s->slab_rank, W = s->slab_world;
    const unsigned long long cap = s->peer_stage_cap;
    const int sides = B.n * ((R > 0 ? 1 : 0) + (R + 1 < W ? 1 : 0));
    static const int max_blocks = getenv("EQGPU_PEER_BLOCKS") ? std::max(1, atoi(getenv("EQGPU_PEER_BLOCKS"))) : (int)PEER_PUSH_BLOCKS;   // tuning knob
    // measured, 2 x B200, 16384 x 4096: 16 blocks of 16384 doubles 72.6 steps/s, 32 x 8192 77.8, 64 x 4096 80.9, 128 x 2048 83.0,
    // 148 x 1024 84.1, 148 x 256 .. 592 x 256 84.5-84.6 -- the exchange wants every SM's load/store slots, not few fat blocks
    static const int per_block = getenv("EQGPU_PEER_PER_BLOCK") ? std::max(256, atoi(getenv("EQGPU_PEER_PER_BLOCK"))) : 512;
    const int blocks = (int)std::min<unsigned long long>(max_blocks, std::max<unsigned long long>(sides, s->peer_batch_fill * (unsigned long long)sides / B.n / per_block));
    ++s->peer_xseq;
    // staging: mine is [from_lo: parity 0, 1][from_hi: parity 0, 1]; a neighbour's has the same layout
    uint4 *mine = (uint4 *)s->peer_stage, *lo = (uint4 *)s->peer_stage_lo, *hi = (uint4 *)s->peer_stage_hi;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    EQ_CUDA(cudaLaunchKernelEx(&cfg, k_halo_push, B, lo ? lo + 2 * cap : (uint4 *)nullptr, hi, (const uint4 *)mine,
                               (const uint4 *)(mine + 2 * cap), cap, s->peer_xseq, s->peer_timeout_ns, s->peer_err));
    B.n = 0;
    s->peer_batch_fill = 0;
    s->comm_peer_exchanges++;
    s->launches++;
    solver_trace_mark(s->stream, "peer-xch");
    EQ_CUDA(cudaGetLastError());
    return 0;
}

// one vector into the current batch; false: the peer path cannot take it (the caller falls back to NCCL).  The decision
// depends only on sizes that are the same on every rank, so all ranks take the same branch.
static bool peer_add(eqgpu_solver *s, const LevelDev &L, double *v, int depth)
{
    const bool below = L.own0 > 0, above = L.own1 < L.ny;
    const size_t nx = (size_t)L.nx, cnt = nx * depth, padded = (cnt + 1) & ~(size_t)1;
    if (padded > s->peer_stage_cap) return false;
    PushBatch &B = *(PushBatch *)s->peer_batch;
    if ((B.n == PEER_MAX_JOBS || s->peer_batch_fill + padded > s->peer_stage_cap) && peer_flush(s)) return false;
    PushJob &J = B.j[B.n++];
    J.snd_lo = below ? v + (size_t)L.own0 * nx : nullptr;
    J.snd_hi = above ? v + (size_t)(L.own1 - depth) * nx : nullptr;
    J.rcv_lo = below ? v + (size_t)(L.own0 - depth) * nx : nullptr;
    J.rcv_hi = above ? v + (size_t)L.own1 * nx : nullptr;
    J.cnt = cnt;
    J.off = s->peer_batch_fill;
    s->peer_batch_fill += padded;
    s->comm_halo_bytes += (long long)cnt * 8 * ((below ? 1 : 0) + (above ? 1 : 0));
    return true;
}

// Several exchanges (different levels, different vectors) that have no kernel between them travel as ONE NCCL group: NCCL
// groups nest, the sends and receives start at the outermost ncclGroupEnd -- one launch instead of one per exchange.
int slab_group_begin(eqgpu_solver *s)
{
    if (!s->slab || s->slab_world < 2) return 0;
    EQ_NCCL(g_nccl.GroupStart());
    s->slab_group_depth++;
    return 0;
}
int slab_group_end(eqgpu_solver *s)
{
    if (!s->slab || s->slab_world < 2 || s->slab_group_depth == 0) return 0;
    s->slab_group_depth--;
    EQ_NCCL(g_nccl.GroupEnd());
    if (s->slab_group_depth == 0 && !s->peer_ok) solver_trace_mark(s->stream, "nccl-xch-grp");
    if (s->slab_group_depth == 0 && s->peer_ok) {
        int rc = peer_flush(s);
        if (rc) return rc;
    }
    if (s->slab_group_depth == 0) s->comm_exchange_groups++;
    return 0;
}

int slab_allreduce(eqgpu_solver *s, const double *src, double *dst, int count)
{
    if (s->peer_ok && count <= PEER_AR_MAX) {   // one warp over peer memory instead of an NCCL kernel
        PeerFlagPtrs F;
        for (int r = 0; r < PEER_MAX_WORLD; ++r) F.p[r] = s->peer_flags_of[r];
        ++s->peer_arseq;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        EQ_CUDA(cudaLaunchKernelEx(&cfg, k_peer_allreduce, src, dst, count, F, s->slab_rank, s->slab_world, s->peer_arseq,
                                   s->peer_timeout_ns, s->peer_err));
        s->comm_allreduce_calls++;
        s->comm_allreduce_doubles += count;
        s->comm_peer_allreduces++;
        s->launches++;
        solver_trace_mark(s->stream, "peer-allred");
        EQ_CUDA(cudaGetLastError());
        return 0;
    }
    EQ_NCCL(g_nccl.AllReduce(src, dst, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)s->nccl_comm, s->stream));
    solver_trace_mark(s->stream, "nccl-allred");
    s->comm_allreduce_calls++;
    s->comm_allreduce_doubles += count;
    return 0;
}
