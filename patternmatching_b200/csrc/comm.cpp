// comm.cpp -- the multi-GPU gather of SURVEY 8(e) in the library: one process per GPU, contiguous shards with a halo,
// no exchange step in the scan; afterwards the per-rank position-sorted record lists are gathered to one rank with
// NCCL over NVLink -- an all-gather of the counts, then grouped ncclSend / ncclRecv of the variable-length lists (rank
// order is position order, so the concatenation is already sorted).
//
// The lists themselves do not go through NCCL when the GPUs can map each other's memory: the root owns a staging buffer
// that every rank maps with CUDA IPC, each rank pushes its list to its place in it with ONE peer-to-peer copy (copy
// engine over NVLink, no SM involved) and a small NCCL collective afterwards tells the root that everything has
// arrived.  Many-to-one grouped ncclSend / ncclRecv reached 260-457 GB/s into the root at N = 8 (1.3-2.3 ms for 598 MB,
// varying from run to run); that path stays as the fall-back (PM_COMM_NO_P2P=1, or when IPC mapping is not possible).
//
// NCCL is resolved at run time (dlopen of the process's libnccl.so.2 -- the one torch already loaded when the caller
// is a torch.distributed program, else the system library): libpm_b200.so itself has no link-time NCCL dependency and
// single-GPU users never touch it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pm_b200.h"

namespace {
thread_local std::string g_comm_err;
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process (torch)?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        auto sym = [&](const char* n) { return dlsym(h, n); };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.Send && a.Recv && a.GroupStart && a.GroupEnd;
    });
    return a;
}
int comm_fail(const std::string& m) { g_comm_err = m; return -1; }
int nccl_fail(ncclResult_t r, const char* what) {
    const NcclApi& a = api();
    return comm_fail(std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"));
}
#define NC(call)                                                   \
    do {                                                           \
        ncclResult_t r__ = (call);                                 \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);      \
    } while (0)
#define CUC(call)                                                                           \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) return comm_fail(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
}  // namespace

struct pm_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    unsigned long long* d_counts = nullptr;   // world + 1 entries: [0] = this rank's count, [1..] = everybody's
    unsigned long long* h_counts = nullptr;   // pinned, world entries
    // peer-to-peer path: the root's staging buffer (owned there, mapped by everybody else through CUDA IPC)
    int p2p_state = 0;                        // 0 = not tried, 1 = usable, -1 = not available (NCCL send / recv instead)
    int p2p_root = -1;
    unsigned char* stage = nullptr;           // root: cudaMalloc; others: cudaIpcOpenMemHandle
    size_t stage_bytes = 0;
    unsigned char* d_xchg = nullptr;          // (world + 1) x 128 bytes of exchange space for handles / flags
    unsigned char* h_xchg = nullptr;          // pinned copy
};

extern "C" {

const char* pm_comm_last_error(void) { return g_comm_err.c_str(); }

int pm_comm_unique_id(uint8_t id[128]) {
    NcclApi& a = api();
    if (!a.ok) return comm_fail("pm_comm_unique_id: libnccl.so.2 not found");
    ncclUniqueId u;
    NC(a.GetUniqueId(&u));
    static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return 0;
}

pm_comm* pm_comm_create(const uint8_t id[128], int rank, int world, int device) {
    NcclApi& a = api();
    if (!a.ok) { comm_fail("pm_comm_create: libnccl.so.2 not found"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { comm_fail("pm_comm_create: cudaSetDevice failed"); return nullptr; }
    pm_comm* c = new pm_comm();
    c->rank = rank; c->world = world; c->device = device;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclResult_t r = a.CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) { nccl_fail(r, "ncclCommInitRank"); delete c; return nullptr; }
    if (cudaMalloc(reinterpret_cast<void**>(&c->d_counts), size_t(world + 1) * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&c->h_counts), size_t(world) * sizeof(unsigned long long)) != cudaSuccess) {
        comm_fail("pm_comm_create: allocation failed");
        pm_comm_free(c);
        return nullptr;
    }
    return c;
}

void pm_comm_free(pm_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comm) api().CommDestroy(c->comm);
    if (c->stage) { if (c->rank == c->p2p_root) cudaFree(c->stage); else cudaIpcCloseMemHandle(c->stage); }
    if (c->d_xchg) cudaFree(c->d_xchg);
    if (c->h_xchg) cudaFreeHost(c->h_xchg);
    if (c->d_counts) cudaFree(c->d_counts);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    delete c;
}

}  // extern "C"

namespace {
constexpr size_t kXchg = 128;   // bytes per rank in the exchange buffers

// all ranks: out[r] = what rank r passed in `mine` (kXchg bytes each); synchronises the stream
int exchange(pm_comm* c, const void* mine, cudaStream_t st) {
    NcclApi& a = api();
    memcpy(c->h_xchg, mine, kXchg);
    CUC(cudaMemcpyAsync(c->d_xchg, c->h_xchg, kXchg, cudaMemcpyHostToDevice, st));
    NC(a.AllGather(c->d_xchg, c->d_xchg + kXchg, kXchg, ncclChar, c->comm, st));
    CUC(cudaMemcpyAsync(c->h_xchg, c->d_xchg + kXchg, kXchg * size_t(c->world), cudaMemcpyDeviceToHost, st));
    CUC(cudaStreamSynchronize(st));
    return 0;
}

// Collective: make the root's staging buffer at least `need` bytes and mapped on every rank.  Returns 1 when the
// peer-to-peer path is usable on ALL ranks, 0 when not (agreed among the ranks), -1 on a hard error.
int ensure_stage(pm_comm* c, int root, size_t need, cudaStream_t st) {
    if (c->p2p_state < 0) return 0;
    if (c->p2p_state == 1 && c->p2p_root == root && c->stage_bytes >= need) return 1;
    if (!c->d_xchg) {
        if (cudaMalloc(reinterpret_cast<void**>(&c->d_xchg), kXchg * size_t(c->world + 1)) != cudaSuccess ||
            cudaMallocHost(reinterpret_cast<void**>(&c->h_xchg), kXchg * size_t(c->world + 1)) != cudaSuccess)
            return comm_fail("pm_comm: allocation failed");
    }
    struct Msg { cudaIpcMemHandle_t handle; unsigned long long bytes; int ok; } msg;
    static_assert(sizeof(Msg) <= kXchg, "exchange record");
    memset(&msg, 0, sizeof(msg));
    // everybody lets go of the old mapping before the root frees it
    if (c->stage && c->rank != c->p2p_root) { cudaIpcCloseMemHandle(c->stage); c->stage = nullptr; }
    if (c->p2p_root >= 0) { unsigned char pad[kXchg] = {0}; if (exchange(c, pad, st)) return -1; }
    if (c->stage && c->rank == c->p2p_root) { cudaFree(c->stage); c->stage = nullptr; }
    c->stage_bytes = 0; c->p2p_root = root;
    const size_t want = need + need / 4 + (size_t(1) << 20);
    msg.ok = getenv("PM_COMM_NO_P2P") ? 0 : 1;
    if (c->rank == root && msg.ok) {
        if (cudaMalloc(reinterpret_cast<void**>(&c->stage), want) != cudaSuccess ||
            cudaIpcGetMemHandle(&msg.handle, c->stage) != cudaSuccess) {
            cudaGetLastError();
            if (c->stage) { cudaFree(c->stage); c->stage = nullptr; }
            msg.ok = 0;
        }
        msg.bytes = want;
    }
    unsigned char rec[kXchg] = {0};
    memcpy(rec, &msg, sizeof(msg));
    if (exchange(c, rec, st)) return -1;
    Msg from_root;
    memcpy(&from_root, c->h_xchg + kXchg * size_t(root), sizeof(from_root));
    int ok = from_root.ok;
    if (ok && c->rank != root) {
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, from_root.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        c->stage = static_cast<unsigned char*>(p);
    }
    // agree: one rank that cannot map the buffer sends everybody to the NCCL path
    memset(rec, 0, sizeof(rec));
    rec[0] = ok ? 1 : 0;
    if (exchange(c, rec, st)) return -1;
    for (int r = 0; r < c->world; ++r) ok = ok && c->h_xchg[kXchg * size_t(r)] == 1;
    if (!ok) {
        if (c->stage && c->rank != root) cudaIpcCloseMemHandle(c->stage);
        else if (c->stage) cudaFree(c->stage);
        c->stage = nullptr; c->p2p_state = -1;
        return 0;
    }
    c->stage_bytes = size_t(from_root.bytes);
    c->p2p_state = 1;
    return 1;
}
}  // namespace

extern "C" {
int pm_comm_gather_records(pm_comm* c, const uint64_t* d_local, uint64_t n_local, uint64_t* d_all, uint64_t cap,
                           uint64_t* counts, uint64_t* n_all, int root, void* cuda_stream) {
    NcclApi& a = api();
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    CUC(cudaSetDevice(c->device));
    const unsigned long long mine = n_local;
    CUC(cudaMemcpyAsync(c->d_counts, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    NC(a.AllGather(c->d_counts, c->d_counts + 1, 1, ncclUint64, c->comm, st));
    CUC(cudaMemcpyAsync(c->h_counts, c->d_counts + 1, size_t(c->world) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUC(cudaStreamSynchronize(st));   // the list lengths size the receives
    uint64_t total = 0;
    for (int r = 0; r < c->world; ++r) { if (counts) counts[r] = c->h_counts[r]; total += c->h_counts[r]; }
    if (n_all) *n_all = total;
    if (c->rank == root && total > cap) return comm_fail("pm_comm_gather_records: the gathered list does not fit d_all");
    // the lists: peer-to-peer into the root's staging buffer when every rank can map it, else grouped send / receive
    const int p2p = ensure_stage(c, root, size_t(total) * sizeof(uint64_t), st);
    if (p2p < 0) return -1;
    if (p2p == 1) {
        uint64_t off = 0;
        for (int r = 0; r < c->rank; ++r) off += c->h_counts[r];
        if (n_local)   // ONE copy per rank: copy engine over NVLink (the root's own part is a local copy)
            CUC(cudaMemcpyAsync(c->stage + off * sizeof(uint64_t), d_local, n_local * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        // everything has arrived once every rank's collective -- enqueued after its copy -- has run
        NC(a.AllGather(c->d_counts, c->d_counts + 1, 1, ncclUint64, c->comm, st));
        if (c->rank == root && total)
            CUC(cudaMemcpyAsync(d_all, c->stage, size_t(total) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    NC(a.GroupStart());
    if (c->rank == root) {
        uint64_t off = 0;
        for (int r = 0; r < c->world; ++r) {
            const uint64_t k = c->h_counts[r];
            if (r != root && k) NC(a.Recv(d_all + off, k, ncclUint64, r, c->comm, st));
            off += k;
        }
    } else if (n_local) {
        NC(a.Send(d_local, n_local, ncclUint64, root, c->comm, st));
    }
    NC(a.GroupEnd());
    if (c->rank == root && n_local) {   // the root's own list: a device-to-device copy into its place
        uint64_t off = 0;
        for (int r = 0; r < root; ++r) off += c->h_counts[r];
        CUC(cudaMemcpyAsync(d_all + off, d_local, n_local * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

}  // extern "C"
