// Row-slab decomposition plumbing: one-row halo exchange and scalar all-reduce over NCCL
// (NVLink 5 / NVSwitch between the GPUs of one box).  The reference's precedent is the PETSc DMDA
// star-stencil halo of diffuclass.cpp:364-370,659-669 and the KSP's internal dot-product
// MPI_Allreduce (diffuclass.cpp:410 [ext]).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy torch has already loaded in the
// calling process, or the system one), so libeqgpu.so itself has no link-time dependency on it and
// single-GPU use never touches it.
#include "eqgpu_internal.cuh"
#include <dlfcn.h>
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

// Refresh `depth` halo rows of a level vector (local view L): my first/last `depth` owned rows go to the
// neighbours below/above, theirs arrive in my halo rows.  Stream-ordered; every rank issues the same sequence.
int slab_exchange2(eqgpu_solver *s, const LevelDev &L, double *v1, double *v2, int depth)
{
    ncclComm_t comm = (ncclComm_t)s->nccl_comm;
    const bool below = L.own0 > 0, above = L.own1 < L.ny;
    if (!below && !above) return 0;
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
    if (s->slab_group_depth == 0) s->comm_exchange_groups++;
    if (bad != 0) {
        s->set_error(std::string("halo exchange: ") + g_nccl.GetErrorString(bad));
        return EQGPU_ECUDA;
    }
    return 0;
}

int slab_exchange(eqgpu_solver *s, const LevelDev &L, double *v, int depth) { return slab_exchange2(s, L, v, nullptr, depth); }

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
    if (s->slab_group_depth == 0) s->comm_exchange_groups++;
    return 0;
}

int slab_allreduce(eqgpu_solver *s, const double *src, double *dst, int count)
{
    EQ_NCCL(g_nccl.AllReduce(src, dst, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)s->nccl_comm, s->stream));
    s->comm_allreduce_calls++;
    s->comm_allreduce_doubles += count;
    return 0;
}
