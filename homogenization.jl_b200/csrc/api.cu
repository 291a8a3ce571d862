// C ABI of libhmg_b200 (see include/hmg.h): context, state vectors, orchestration of the V-cycle.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <dlfcn.h>
#include <nccl.h>

#include "../../include/hmg.h"
#include "hmg_host.hpp"
#include "kernels.cuh"

// NVTX ranges per V-cycle, level and phase (build flag HMG_NVTX, on in the Makefile; NVTX 3 is header-only and costs a
// null-pointer test per range unless a profiler is attached)
#ifdef HMG_NVTX
#include <nvtx3/nvToolsExt.h>
#endif

using namespace hmg;

namespace {

thread_local std::string g_err;

struct Range {
#ifdef HMG_NVTX
    explicit Range(const char* name, int level = -1) {
        if (level < 0) { nvtxRangePushA(name); return; }
        char buf[64];
        snprintf(buf, sizeof(buf), "%s L%d", name, level);
        nvtxRangePushA(buf);
    }
    ~Range() { nvtxRangePop(); }
#else
    explicit Range(const char*, int = -1) {}
#endif
    Range(const Range&) = delete;
    Range& operator=(const Range&) = delete;
};

#define CUDA_OK(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            throw Error(std::string("hmg: CUDA error: ") + cudaGetErrorString(e_) + " in " #call); \
    } while (0)
#define NCCL_OK(call)                                                                        \
    do {                                                                                     \
        ncclResult_t r_ = (call);                                                            \
        if (r_ != ncclSuccess)                                                               \
            throw Error(std::string("hmg: NCCL error: ") + nccl().GetErrorString(r_) + " in " #call); \
    } while (0)
#define CUSOLVER_OK(call)                                                                    \
    do {                                                                                     \
        cusolverStatus_t s_ = (call);                                                        \
        if (s_ != CUSOLVER_STATUS_SUCCESS)                                                   \
            throw Error(std::string("hmg: cuSOLVER error ") + std::to_string((int)s_) + " in " #call); \
    } while (0)

// NCCL is bound at run time (dlopen), never at link time: the process usually holds PyTorch's bundled
// libnccl already, and two NCCL builds in one process do not mix.  Order: the copy already loaded,
// $HMG_NCCL_LIB (the Python binding points it at PyTorch's copy), the system library.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
const NcclApi& nccl() {
    static NcclApi api;
    static bool loaded = false;
    if (loaded) return api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h)
        if (const char* path = getenv("HMG_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw Error("hmg: cannot load libnccl.so.2 (set HMG_NCCL_LIB): partitioned contexts need NCCL");
    auto sym = [&](const char* name) {
        void* f = dlsym(h, name);
        if (!f) throw Error(std::string("hmg: NCCL symbol missing: ") + name);
        return f;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    loaded = true;
    return api;
}

struct LevelDev {
    LevelView view{};
    ApplyConfig cfg{};
    ApplyConfig cfg_fused{};      // launch shape of the fused p-update + product (ring_rows <= 0: does not fit)
    ApplyConfig cfg_rhs{};        // launch shape of the residual / mul! variants
    double* p2 = nullptr;         // the other search-direction buffer of the fused p-update (lazy)
    std::vector<double> tab;      // StencilTab of the level (travels in the kernel parameter block)
    int32_t* hier2lat = nullptr;
    double* vec[HMG_NVEC] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

}  // namespace

namespace { void nccl_destroy(ncclComm_t comm) { nccl().CommDestroy(comm); } }

struct hmg_ctx {
    int dim = 0, nlevels = 0, device = 0;
    int rank = 0, nranks = 1;
    int64_t ne = 0, ne_global = 0, nn = 0;
    int W = 32, wshift = 5;              // elements per unit of the interleaved device layout (= warp size)
    int64_t nunits = 0;
    double lambda = 1.0;
    cudaStream_t stream = nullptr;
    RefElement ref;
    Topology topo;                       // of the WHOLE base mesh
    Partition part;                      // this rank's share (the whole mesh on one GPU)
    std::vector<int64_t> elems;          // local elements, 0-based global node ids, (dim+1) x ne
    std::vector<int64_t> elems_global;   // all elements (coarse operator on rank 0)
    std::vector<double> sigma_global;    // dim x ne_global
    std::vector<double> nodes;           // dim x nn
    std::vector<int64_t> local_to_global;
    // multi-GPU
    ncclComm_t comm = nullptr;
    CutView cutv[3] = {};
    int64_t cut_nglobal[3] = {0, 0, 0};
    // neighbour exchange: one message per rank that shares a cut cell, laid out per level as
    // [shared faces x npf][shared edges x npe][shared vertices]
    std::vector<int> neighbors;
    std::vector<std::vector<int64_t>> msg_off, msg_len;   // [level][neighbour]
    std::vector<int64_t*> kbase;         // per level, device: [nranks * 3] first entry of a kind's section
    double *p2p_send = nullptr, *p2p_recv = nullptr;
    // peer memory (NVLink, CUDA IPC): scalar all-reduces inside the reduction kernels and the cut exchange without a
    // collective call.  peer_on is agreed on by all ranks at creation (HMG_PEER=0 or a failed mapping: NCCL path).
    bool peer_on = false;
    void* peer_buf = nullptr;                    // this rank's communication buffer (exported)
    std::vector<void*> peer_mapped;              // the other ranks' buffers as mapped here (nullptr for the own rank)
    std::vector<CutPeer> cut_peer;               // per level
    uint8_t* node_contrib = nullptr;
    // driver functionals (finest level)
    double* dphi = nullptr;              // [nf][dim]
    double* flux = nullptr;              // [nunits][dim][W]  -|J| J^-1 (sigma .* xi)
    double* coef_mass1 = nullptr;        // element coefficients of the bare reference mass matrix (P = 0, last = 1)
    double* coef_massJ = nullptr;        // ... of |J| M (P = 0, last = |J|)
    int32_t* gidx = nullptr;             // [nunits * W] global element index of a column, -1 for padding
    std::vector<LevelDev> lv;
    TopoView tview{};
    double* elem_coef = nullptr;         // [nunits][CS][W]
    std::vector<double> elem_coef_host;  // [ne][CS] (local elements)
    uint16_t* cmask = nullptr;           // [nunits * W]
    uint8_t* mult = nullptr;             // [nunits][16][W] owners of the cell of every node class
    int32_t* belems = nullptr;           // elements touching the domain boundary
    int64_t nbelems = 0;
    int32_t* node_first = nullptr;
    int32_t* elems32 = nullptr;
    Reducer red{};
    std::vector<void*> allocs;
    // coarse level
    int64_t n_interior = 0;
    int64_t* interior_idx = nullptr;
    double* Ainv = nullptr;
    // lambda / sigma the coarse inverse was built for: op_gen counts the changes of the operator, coarse_gen is the
    // count the inverse belongs to; coarse_internal = the library assembled the matrix itself (it can do so again)
    int64_t op_gen = 0, coarse_gen = -1;
    bool coarse_internal = false;
    double *ubase = nullptr, *bint = nullptr, *xint = nullptr;
    double* symv_work = nullptr;         // per-tile partial sums of the half-traffic coarse mat-vec
    // staging
    double* staging[2] = {nullptr, nullptr};       // double-buffered: the copy of chunk i+1 overlaps the permutation of chunk i
    size_t staging_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_permuted[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    // CUDA graphs of whole V-cycles, one per (top level, steps, with norm): HMG_GRAPH=0 launches eagerly.  A graph is
    // captured at the second request of its kind (the first runs eagerly and performs every lazy allocation) and dropped
    // whenever lambda, sigma or the coarse matrix change (lambda travels in the kernel parameters).
    struct CycleGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0; bool warmed = false; };
    std::map<std::tuple<int, int, int>, CycleGraph> graphs;
    int graph_mode = 1;                  // 0 = off, 1 = on (partitioned contexts: with peer memory only), 2 = also on the NCCL path
    void drop_graphs() {
        for (auto& kv : graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        graphs.clear();
    }

    ~hmg_ctx() {
        // also runs when hmg_create fails half-way: nothing of the context outlives it
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        drop_graphs();
        if (peer_on) peer_teardown();
        if (comm) nccl_destroy(comm);
        for (void* p : allocs) cudaFree(p);
        for (int q = 0; q < 2; ++q) {
            if (ev_copied[q]) cudaEventDestroy(ev_copied[q]);
            if (ev_permuted[q]) cudaEventDestroy(ev_permuted[q]);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
    void peer_teardown();
    template <class T> T* dalloc(size_t n, bool zero = true) {
        void* p = nullptr;
        if (n == 0) n = 1;
        CUDA_OK(cudaMalloc(&p, n * sizeof(T)));
        allocs.push_back(p);
        if (zero) CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(T), stream));
        return reinterpret_cast<T*>(p);
    }
    template <class T> T* dupload(const std::vector<T>& h) {
        T* p = dalloc<T>(h.size(), false);
        if (!h.empty()) CUDA_OK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaStreamSynchronize(stream));   // h may be a temporary
        return p;
    }
    void dfree(void* p) {
        auto it = std::find(allocs.begin(), allocs.end(), p);
        if (it != allocs.end()) { cudaFree(p); allocs.erase(it); }
    }
    LevelDev& level(int l) {
        HMG_CHECK(l >= 1 && l <= nlevels, "level out of range");
        return lv[l - 1];
    }
    double* vecp(int l, int which) {
        HMG_CHECK(which >= 0 && which < HMG_NVEC, "unknown state vector id");
        LevelDev& L = level(l);
        if (!L.vec[which]) L.vec[which] = dalloc<double>((size_t)nstored(l));   // scratch vectors are lazy
        return L.vec[which];
    }
    // stored entries of a level vector including the zero columns that pad the last unit
    int64_t nstored(int l) { return (int64_t)level(l).view.nf * W * nunits; }
    // packed cut buffer of a level: [cut faces x npf][cut edges x npe][cut vertices]
    int64_t cut_base(int l, int kind) {
        const LevelView& V = level(l).view;
        const int64_t f = cut_nglobal[0] * V.npf, e = cut_nglobal[1] * V.npe;
        return kind == 0 ? 0 : (kind == 1 ? f : f + e);
    }
    int64_t cut_slots(int l) { return cut_base(l, 2) + cut_nglobal[2]; }
};

// collective: nobody unmaps or frees a buffer a peer may still touch
void hmg_ctx::peer_teardown() {
    int* d = nullptr;
    auto barrier = [&]() {
        if (!comm || !d) return;
        nccl().AllReduce(d, d, 1, ncclInt, ncclSum, comm, stream);
        cudaStreamSynchronize(stream);
    };
    if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) d = nullptr;
    if (d) cudaMemsetAsync(d, 0, sizeof(int), stream);
    barrier();
    for (void*& m : peer_mapped) if (m) { cudaIpcCloseMemHandle(m); m = nullptr; }
    barrier();
    if (peer_buf) { cudaFree(peer_buf); peer_buf = nullptr; }
    if (d) cudaFree(d);
    peer_on = false;
}

namespace {

void peer_setup(hmg_ctx* c, int64_t max_msg);

void check_launch(hmg_ctx* c, int n) {
    c->launches += n;
    CUDA_OK(cudaGetLastError());
}

void upload_operator(hmg_ctx* c, const double* sigma /* dim x ne_global */) {
    const int cs = c->dim == 3 ? 8 : 4, dim = c->dim;
    c->sigma_global.assign(sigma, sigma + (size_t)c->ne_global * dim);
    std::vector<double> sl((size_t)c->ne * dim);
    for (int64_t e = 0; e < c->ne; ++e)
        for (int d = 0; d < dim; ++d) sl[(size_t)e * dim + d] = sigma[(size_t)c->local_to_global[e] * dim + d];
    element_coefficients(c->dim, c->ne, c->nodes.data(), c->elems.data(), sl.data(), c->elem_coef_host, cs);
    const int W = c->W;
    std::vector<double> inter((size_t)c->nunits * cs * W, 0.0);     // [unit][component][lane]
    for (int64_t e = 0; e < c->ne; ++e)
        for (int q = 0; q < cs; ++q) inter[((size_t)(e / W) * cs + q) * W + e % W] = c->elem_coef_host[(size_t)e * cs + q];
    if (!c->elem_coef) c->elem_coef = c->dalloc<double>(inter.size(), false);
    CUDA_OK(cudaMemcpyAsync(c->elem_coef, inter.data(), inter.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
}

// Collective over the ranks of the context: agree on the size of the receive areas, allocate and export the own
// communication buffer, map everybody else's (CUDA IPC; the devices reach each other over NVLink), exchange the
// message layouts.  Any rank that cannot do it makes ALL ranks keep the NCCL path (the decision is all-reduced).
void peer_setup(hmg_ctx* c, int64_t max_msg) {
    const int nr = c->nranks, nl = c->nlevels;
    const char* env = getenv("HMG_PEER");
    int ok = !(env && atoi(env) == 0) && nr <= 32;
    // 1. sizes and layouts: max message length, and every rank's kbase table of every level
    std::vector<int64_t> mine((size_t)1 + (size_t)nl * nr * 3, 0), all((size_t)nr * mine.size(), 0);
    mine[0] = max_msg;
    for (int l = 0; l < nl; ++l)
        CUDA_OK(cudaMemcpyAsync(&mine[1 + (size_t)l * nr * 3], c->kbase[l], (size_t)nr * 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    int64_t* d_mine = c->dupload(mine);
    int64_t* d_all = c->dalloc<int64_t>(all.size());
    NCCL_OK(nccl().AllGather(d_mine, d_all, mine.size() * sizeof(int64_t), ncclChar, c->comm, c->stream));
    CUDA_OK(cudaMemcpyAsync(all.data(), d_all, all.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    int64_t stride = 1;
    for (int q = 0; q < nr; ++q) stride = std::max(stride, all[(size_t)q * mine.size()]);
    stride = (stride + 15) / 16 * 16;
    // 2. the own buffer: [mail PEER_SLOTS x nr][flags nr][ticket, counters][recv 2 x stride]
    const size_t mail_bytes = (size_t)PEER_SLOTS * nr * sizeof(PeerMail);
    const size_t flag_bytes = ((size_t)nr * sizeof(unsigned long long) + 127) / 128 * 128;
    const size_t ctr_bytes = 128;
    const size_t recv_off = (mail_bytes + flag_bytes + ctr_bytes + 255) / 256 * 256;
    const size_t total = recv_off + (size_t)2 * stride * sizeof(double);
    cudaIpcMemHandle_t handle;
    std::memset(&handle, 0, sizeof(handle));
    if (ok) {
        if (cudaMalloc(&c->peer_buf, total) != cudaSuccess) { cudaGetLastError(); c->peer_buf = nullptr; ok = 0; }
    }
    if (ok) {
        CUDA_OK(cudaMemsetAsync(c->peer_buf, 0, total, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        if (cudaIpcGetMemHandle(&handle, c->peer_buf) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    // 3. everybody's handle, then the mapping
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    std::vector<char> hmine(64 + 8, 0), hall((size_t)nr * hmine.size(), 0);
    std::memcpy(hmine.data(), &handle, 64);
    hmine[64] = (char)ok;
    char* d_hm = c->dupload(hmine);
    char* d_ha = c->dalloc<char>(hall.size());
    NCCL_OK(nccl().AllGather(d_hm, d_ha, hmine.size(), ncclChar, c->comm, c->stream));
    CUDA_OK(cudaMemcpyAsync(hall.data(), d_ha, hall.size(), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int q = 0; q < nr; ++q) ok = ok && hall[(size_t)q * hmine.size() + 64];
    c->peer_mapped.assign(nr, nullptr);
    if (ok) {
        for (int q = 0; q < nr && ok; ++q) {
            if (q == c->rank) continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, &hall[(size_t)q * hmine.size()], 64);
            if (cudaIpcOpenMemHandle(&c->peer_mapped[q], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                c->peer_mapped[q] = nullptr;
                ok = 0;
            }
        }
    }
    // 4. the decision is collective: one rank without the mapping and everybody stays on NCCL
    int* d_ok = c->dupload(std::vector<int>(1, ok));
    NCCL_OK(nccl().AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c->comm, c->stream));
    CUDA_OK(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->dfree(d_mine); c->dfree(d_all); c->dfree(d_hm); c->dfree(d_ha); c->dfree(d_ok);
    if (!ok) {
        for (void*& m : c->peer_mapped) if (m) { cudaIpcCloseMemHandle(m); m = nullptr; }
        if (c->peer_buf) { cudaFree(c->peer_buf); c->peer_buf = nullptr; }
        if (getenv("HMG_DEBUG_CFG")) fprintf(stderr, "hmg: rank %d: peer memory off, NCCL path\n", c->rank);
        return;
    }
    // 5. device views
    std::vector<PeerMail*> mail(nr);
    std::vector<unsigned long long*> flag(nr);
    std::vector<double*> recv(nr);
    for (int q = 0; q < nr; ++q) {
        char* base = static_cast<char*>(q == c->rank ? c->peer_buf : c->peer_mapped[q]);
        mail[q] = reinterpret_cast<PeerMail*>(base);
        flag[q] = reinterpret_cast<unsigned long long*>(base + mail_bytes);
        recv[q] = reinterpret_cast<double*>(base + recv_off);
    }
    char* own = static_cast<char*>(c->peer_buf);
    PeerView& P = c->red.peer;
    P.rank = c->rank; P.nranks = nr;
    P.mail = c->dupload(mail);
    P.flag = c->dupload(flag);
    P.recv = c->dupload(recv);
    P.recv_stride = stride;
    P.rseq = reinterpret_cast<unsigned long long*>(own + mail_bytes + flag_bytes);
    P.xseq = P.rseq + 1;
    P.xticket = reinterpret_cast<unsigned int*>(P.rseq + 2);
    std::vector<int32_t> nbr(c->neighbors.begin(), c->neighbors.end());
    const int32_t* d_nbr = c->dupload(nbr);
    c->cut_peer.resize(nl);
    for (int l = 0; l < nl; ++l) {
        // where section `kind` of MY message starts inside the receive area of rank q: q's own table entry for me
        std::vector<int64_t> rb((size_t)nr * 3, 0);
        for (int q = 0; q < nr; ++q)
            for (int kind = 0; kind < 3; ++kind)
                rb[(size_t)q * 3 + kind] = all[(size_t)q * mine.size() + 1 + ((size_t)l * nr + c->rank) * 3 + kind];
        c->cut_peer[l].rbase = c->dupload(rb);
        c->cut_peer[l].nbr = d_nbr;
        c->cut_peer[l].nnbr = (int)nbr.size();
    }
    c->peer_on = true;
    if (getenv("HMG_DEBUG_CFG")) fprintf(stderr, "hmg: rank %d: peer memory on (%d neighbours, receive areas 2 x %lld doubles)\n", c->rank, (int)nbr.size(), (long long)stride);
}

hmg_ctx* create_impl(int dim, int nlevels, int64_t ne_global, int64_t nn, const double* base_nodes,
                     const int64_t* base_elems, const double* sigma, double lambda, int device, int rank, int nranks,
                     const int32_t* owner_rank, const void* nccl_id) {
    HMG_CHECK(base_nodes && base_elems && sigma, "null input array");
    HMG_CHECK(ne_global > 0 && nn > 0, "empty base mesh");
    HMG_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank");
    HMG_CHECK(nranks == 1 || (owner_rank && nccl_id), "a partitioned context needs owner_rank and the NCCL id");
    int ndev = 0;
    cudaError_t de = cudaGetDeviceCount(&ndev);
    if (de != cudaSuccess || ndev == 0)
        throw Error("hmg: no CUDA device available -- this library has no CPU fallback");
    HMG_CHECK(device >= 0 && device < ndev, "device index out of range");
    CUDA_OK(cudaSetDevice(device));
    std::unique_ptr<hmg_ctx> c(new hmg_ctx);
    c->dim = dim;
    c->nlevels = nlevels;
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    c->ne_global = ne_global;
    c->nn = nn;
    c->lambda = lambda;
    if (const char* v = getenv("HMG_GRAPH")) c->graph_mode = atoi(v);
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreate(&c->ev0));
    CUDA_OK(cudaEventCreate(&c->ev1));

    const int nv = dim + 1;
    c->elems_global.resize((size_t)ne_global * nv);
    for (size_t q = 0; q < c->elems_global.size(); ++q) c->elems_global[q] = base_elems[q] - 1;   // Julia is 1-based
    c->nodes.assign(base_nodes, base_nodes + (size_t)nn * dim);

    c->ref = build_reference(dim, nlevels);
    c->topo = build_topology(dim, ne_global, nn, c->elems_global.data());
    c->part = build_partition(c->topo, nranks > 1 ? owner_rank : nullptr, rank, nranks);
    const Partition& P = c->part;
    c->local_to_global = P.local_to_global;
    const int64_t ne = (int64_t)P.local_to_global.size();
    HMG_CHECK(ne > 0, "this rank owns no coarse element");
    c->ne = ne;
    c->elems.resize((size_t)ne * nv);
    for (int64_t e = 0; e < ne; ++e)
        for (int a = 0; a < nv; ++a) c->elems[(size_t)e * nv + a] = c->elems_global[(size_t)P.local_to_global[e] * nv + a];
    if (nranks > 1) {
        ncclUniqueId id;
        static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id size");
        std::memcpy(&id, nccl_id, sizeof(id));
        NCCL_OK(nccl().CommInitRank(&c->comm, nranks, id, rank));
    }
    // interleave width: one warp lane per element of a unit
    c->W = 32;
    c->wshift = 5;
    c->nunits = (ne + c->W - 1) / c->W;

    // per-level tables and state vectors
    c->lv.resize(nlevels);
    for (int l = 1; l <= nlevels; ++l) {
        const RefLevel& R = c->ref.lv[l - 1];
        LevelDev& L = c->lv[l - 1];
        LevelView& V = L.view;
        V.m = R.m; V.nf = R.nf;
        V.W = c->W; V.wshift = c->wshift;
        V.n_boundary = (int)R.boundary.size();
        V.npf = dim == 3 ? (int)R.face_bary.size() : 0;
        V.npe = R.m - 1;
        V.nodeinfo = c->dupload(R.nodeinfo);
        V.boundary = c->dupload(R.boundary);
        V.G = c->dupload(R.G);
        V.iface_idx = c->dupload(R.iface_idx);
        V.interp_tab = c->dupload(R.interp_tab);
        V.restrict_tab = c->dupload(R.restrict_tab);
        for (int v = 0; v < 4; ++v) V.vpos[v] = v < nv ? R.hier2lat[v] : 0;
        L.hier2lat = c->dupload(R.hier2lat);
        L.cfg = make_apply_config(dim, R.m, R.nf, c->W);
        HMG_CHECK(L.cfg.ring_rows > 0, "a level of this hierarchy does not fit the shared-memory ring of the apply kernel");
        L.cfg_rhs = make_apply_config(dim, R.m, R.nf, c->W, false, true);
        if (L.cfg_rhs.ring_rows <= 0) L.cfg_rhs = L.cfg;
        L.cfg_fused = make_apply_config(dim, R.m, R.nf, c->W, true);
        if (getenv("HMG_FUSE_P") && atoi(getenv("HMG_FUSE_P")) == 0) L.cfg_fused.ring_rows = -1;     // tests: the unfused path
        if (getenv("HMG_DEBUG_CFG"))
            for (const ApplyConfig* q : {&L.cfg, &L.cfg_rhs, &L.cfg_fused})
                fprintf(stderr, "hmg: level %d (m=%d) %s: warps %d ring %d spill %d chunk %d run %d seg %d conv %d slots %d smem %zu\n", l, R.m,
                        q == &L.cfg ? "product " : (q == &L.cfg_rhs ? "residual" : "fused   "), q->nwarps, q->ring_rows, q->spill_rows,
                        1 << q->chunk_shift, q->run, q->seg, q->nconv, q->nconv ? 1 << q->slot_shift : 0, q->smem_bytes);
        L.tab = R.gi;
        L.tab.insert(L.tab.end(), R.gc.begin(), R.gc.end());
        L.tab.insert(L.tab.end(), R.ge.begin(), R.ge.end());
        for (int w = 0; w <= HMG_AP; ++w) L.vec[w] = c->dalloc<double>((size_t)V.nf * c->W * c->nunits);
    }
    // topology of the cells whose owners are all local
    {
        const CellMap& pairs = dim == 3 ? P.faces : P.edges;     // codimension-1 cells: exactly two owners
        std::vector<int32_t> partner((size_t)ne * 4, -1);
        for (int64_t q = 0; q < pairs.ncells(); ++q) {
            HMG_CHECK(pairs.offset[q + 1] - pairs.offset[q] == 2, "a face of the base mesh is shared by more than two elements");
            const int32_t a = pairs.owner[pairs.offset[q]], b = pairs.owner[pairs.offset[q] + 1];
            partner[(size_t)(a >> 3) * 4 + (a & 7)] = b;
            partner[(size_t)(b >> 3) * 4 + (b & 7)] = a;
        }
        c->tview.ne = ne;
        c->tview.partner = c->dupload(partner);
        static const std::vector<int64_t> empty_off(1, 0);
        static const std::vector<int32_t> empty_own;
        c->tview.nedges = dim == 3 ? P.edges.ncells() : 0;
        c->tview.edge_off = c->dupload(dim == 3 ? P.edges.offset : empty_off);
        c->tview.edge_own = c->dupload(dim == 3 ? P.edges.owner : empty_own);
        c->tview.nverts = P.verts.ncells();
        c->tview.vert_off = c->dupload(P.verts.offset);
        c->tview.vert_own = c->dupload(P.verts.owner);
    }
    // cut cells (owners on several ranks)
    if (nranks > 1) {
        for (int kind = 0; kind < 3; ++kind) {
            const CutCells& C = P.cut[kind];
            c->cut_nglobal[kind] = C.nglobal;
            c->cutv[kind].ncells = C.ncells();
            c->cutv[kind].off = c->dupload(C.offset);
            c->cutv[kind].own = c->dupload(C.owner);
            c->cutv[kind].first_local = c->dupload(C.first_local);
        }
        c->node_contrib = c->dupload(P.node_contrib);
        for (int kind = 0; kind < 3; ++kind) {
            const CutCells& C = P.cut[kind];
            c->cutv[kind].peer_off = c->dupload(C.peer_off);
            c->cutv[kind].peer_rank = c->dupload(C.peer_rank);
            c->cutv[kind].peer_idx = c->dupload(C.peer_idx);
            c->cutv[kind].my_pos = c->dupload(C.my_pos);
        }
        for (int q = 0; q < nranks; ++q)
            if (q != rank && P.shared_with[(size_t)q * 3] + P.shared_with[(size_t)q * 3 + 1] + P.shared_with[(size_t)q * 3 + 2] > 0)
                c->neighbors.push_back(q);
        c->msg_off.resize(nlevels); c->msg_len.resize(nlevels); c->kbase.resize(nlevels);
        int64_t max_msg = 1;
        for (int l = 1; l <= nlevels; ++l) {
            const LevelView& V = c->lv[l - 1].view;
            std::vector<int64_t> kb((size_t)nranks * 3, 0);
            int64_t at = 0;
            for (int q : c->neighbors) {
                c->msg_off[l - 1].push_back(at);
                const int64_t nper[3] = {V.npf, V.npe, 1};
                for (int kind = 0; kind < 3; ++kind) {
                    kb[(size_t)q * 3 + kind] = at;
                    at += P.shared_with[(size_t)q * 3 + kind] * nper[kind];
                }
                c->msg_len[l - 1].push_back(at - c->msg_off[l - 1].back());
            }
            c->kbase[l - 1] = c->dupload(kb);
            max_msg = std::max(max_msg, at);
        }
        c->p2p_send = c->dalloc<double>((size_t)max_msg);
        c->p2p_recv = c->dalloc<double>((size_t)max_msg);
        peer_setup(c.get(), max_msg);
    }
    {
        std::vector<uint16_t> cm((size_t)c->nunits * c->W, 0);
        std::vector<int32_t> be;
        for (int64_t e = 0; e < ne; ++e) {
            cm[e] = P.cmask[e];
            if (P.cmask[e]) be.push_back((int32_t)e);
        }
        c->cmask = c->dupload(cm);
        c->nbelems = (int64_t)be.size();
        c->belems = c->dupload(be);
        // owners (on all ranks) of the base-mesh cell behind every node class: weights of the fused dot products
        std::vector<uint8_t> mult((size_t)c->nunits * 16 * c->W, 1);
        for (int64_t e = 0; e < ne; ++e)
            for (int q = 0; q < 16; ++q)
                mult[((size_t)(e / c->W) * 16 + q) * c->W + e % c->W] = P.mult[(size_t)e * 16 + q];
        c->mult = c->dupload(mult);
    }
    c->node_first = c->dupload(P.node_first);
    std::vector<int32_t> e32(c->elems.begin(), c->elems.end());
    c->elems32 = c->dupload(e32);
    upload_operator(c.get(), sigma);
    // reductions
    c->red.max_blocks = 148 * 8;
    c->red.partials = c->dalloc<double>(c->red.max_blocks);
    c->red.scalars = c->dalloc<double>(S_COUNT);
    c->red.ticket = c->dalloc<unsigned int>(1);
    c->ubase = c->dalloc<double>(nn);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return c.release();
}

// ---- building blocks ---------------------------------------------------------------------
void do_apply(hmg_ctx* c, int l, int mode, double alpha, const double* x, double* y, const double* b, int dot_post = -1) {
    LevelDev& L = c->level(l);
    ApplyArgs a;
    a.L = L.view;
    a.cfg = L.cfg;
    a.cfg_fused = L.cfg_fused;
    a.cfg_rhs = L.cfg_rhs;
    a.nunits = c->nunits;
    a.tab = L.tab.data();
    a.coef = c->elem_coef;
    a.cmask = c->cmask;
    a.mult = c->mult;
    a.x = x; a.y = y; a.b = b;
    a.alpha = alpha; a.lambda = c->lambda;
    a.mode = mode;
    a.dot_post = dot_post;
    a.red = c->red;
    const int n = launch_apply(c->dim, a, c->stream);
    HMG_CHECK(n >= 0, "apply kernel refused the launch configuration");
    check_launch(c, n);
}
// cut cells: only interface partial sums move (NCCL over NVLink).  Every rank sends the partial sum of a cut node to
// the ranks that share it (grouped ncclSend / ncclRecv with its <= 7 neighbours in a block partition) and adds the
// partial sums in ascending rank order.  sq: also add owners x total^2 to S_TMP.
void do_cut_exchange_impl(hmg_ctx* c, int l, double* x, bool sq, int sq_post = POST_ADD) {
    const LevelView& V = c->level(l).view;
    if (c->peer_on) {
        // peer memory: the pack kernel stores into the neighbours' receive areas and raises their flags, the unpack
        // kernel waits for the neighbours' flags -- two launches, no collective call
        const CutPeer* cp = &c->cut_peer[l - 1];
        check_launch(c, launch_cut_p2p(c->dim, CUT_PACK, V, c->cutv, c->kbase[l - 1], x, nullptr, false, c->red, c->stream, cp));
        check_launch(c, launch_cut_p2p(c->dim, CUT_UNPACK, V, c->cutv, c->kbase[l - 1], x, nullptr, sq, c->red, c->stream, cp, sq_post));
        return;
    }
    const std::vector<int64_t>& off = c->msg_off[l - 1];
    const std::vector<int64_t>& len = c->msg_len[l - 1];
    check_launch(c, launch_cut_p2p(c->dim, CUT_PACK, V, c->cutv, c->kbase[l - 1], x, c->p2p_send, false, c->red, c->stream));
    NCCL_OK(nccl().GroupStart());
    for (size_t q = 0; q < c->neighbors.size(); ++q) {
        if (len[q] == 0) continue;
        NCCL_OK(nccl().Send(c->p2p_send + off[q], (size_t)len[q], ncclDouble, c->neighbors[q], c->comm, c->stream));
        NCCL_OK(nccl().Recv(c->p2p_recv + off[q], (size_t)len[q], ncclDouble, c->neighbors[q], c->comm, c->stream));
    }
    NCCL_OK(nccl().GroupEnd());
    check_launch(c, launch_cut_p2p(c->dim, CUT_UNPACK, V, c->cutv, c->kbase[l - 1], x, c->p2p_recv, sq, c->red, c->stream, nullptr, sq_post));
}
void do_cut_exchange(hmg_ctx* c, int l, double* x) {
    const int64_t slots = c->cut_slots(l);
    if (c->nranks == 1 || slots == 0) return;
    do_cut_exchange_impl(c, l, x, false);
}
// The two halves of the peer-memory exchange.  Cut cells and rank-local cells touch disjoint entries, so the pack can
// run BEFORE the local interface kernel: the messages cross NVLink (and the neighbours get to their own pack) while that
// kernel sums the local cells, and the unpack finds the flags raised.  Measured on 8 ranks (tools/comm_bench.py, C4):
// an exchange on its own costs 71 us at level 6 and 27 us on the small levels.
bool peer_exchange(hmg_ctx* c, int l) { return c->nranks > 1 && c->peer_on && c->cut_slots(l) > 0; }
void cut_pack_peer(hmg_ctx* c, int l, double* x) {
    check_launch(c, launch_cut_p2p(c->dim, CUT_PACK, c->level(l).view, c->cutv, c->kbase[l - 1], x, nullptr, false, c->red, c->stream,
                                   &c->cut_peer[l - 1]));
}
void cut_unpack_peer(hmg_ctx* c, int l, double* x, bool sq = false, int sq_post = POST_ADD) {
    check_launch(c, launch_cut_p2p(c->dim, CUT_UNPACK, c->level(l).view, c->cutv, c->kbase[l - 1], x, nullptr, sq, c->red, c->stream,
                                   &c->cut_peer[l - 1], sq_post));
}
// part: 3 = every shared cell; 1 / 2 = only the two-owner cells / the cells with more owners (measurement) -- the cut
// cells are exchanged either way
void do_broadcast(hmg_ctx* c, int l, double* x, int part = 3) {
    const bool px = peer_exchange(c, l);
    if (px) cut_pack_peer(c, l, x);
    check_launch(c, launch_interface_sum(c->dim, c->level(l).view, c->tview, x, c->stream, part));
    if (px) cut_unpack_peer(c, l, x);
    else do_cut_exchange(c, l, x);
}
void do_zero_all_but_one(hmg_ctx* c, int l, double* x) {
    check_launch(c, launch_zero_all_but_one(c->dim, c->level(l).view, c->tview, x, c->stream));
    if (c->nranks > 1) check_launch(c, launch_cut_zero_but_first(c->dim, c->level(l).view, c->cutv, x, c->stream));
}
// Several ranks: with peer memory the kernel itself sums over the ranks (POST_GLOBAL) and nothing is left to do; on the
// NCCL path the kernel stores the local sum in S_TMP, the sums are all-reduced and a one-thread kernel derives the scalars
int kernel_post(hmg_ctx* c, int post) {
    if (c->nranks == 1) return post;
    return c->peer_on ? (post | POST_GLOBAL) : (int)POST_STORE;
}
// in_kernel = false: the reduction kernel does not know about ranks (it stored its local sum in S_TMP)
void finish_reduction(hmg_ctx* c, int post, int slot, bool in_kernel = true) {
    if (c->nranks == 1) return;
    if (c->peer_on) {
        if (!in_kernel) check_launch(c, launch_scalar_post(c->red, post | POST_GLOBAL, slot, c->stream));
        return;
    }
    NCCL_OK(nccl().AllReduce(c->red.scalars + S_TMP, c->red.scalars + S_TMP, 1, ncclDouble, ncclSum, c->comm, c->stream));
    check_launch(c, launch_scalar_post(c->red, post, slot, c->stream));
}
void do_local_residual(hmg_ctx* c, int l) {
    do_apply(c, l, APPLY_RESIDUAL, 1.0, c->vecp(l, HMG_X), c->vecp(l, HMG_R), c->vecp(l, HMG_B));
}
// y = broadcast(constraint(A x)); with dot_post >= 0 the apply kernel also reduces
// sum_entries owners(entry) * x * y_local = dot(x, y) over all stored entries (x consistent across owners).
// dot_only: nothing but that reduction is wanted -- y is neither stored nor interface-summed.
void do_global_product(hmg_ctx* c, int l, const double* x, double* y, int dot_post = -1, bool dot_only = false, int part = 3) {
    LevelDev& L = c->level(l);
    ApplyArgs a;
    a.L = L.view; a.cfg = L.cfg; a.cfg_fused = L.cfg_fused; a.cfg_rhs = L.cfg_rhs;
    a.nunits = c->nunits; a.tab = L.tab.data();
    a.coef = c->elem_coef; a.cmask = c->cmask; a.mult = c->mult;
    a.x = x; a.y = y; a.b = nullptr;
    a.alpha = 1.0; a.lambda = c->lambda; a.mode = APPLY_AX;
    a.dot_post = dot_post >= 0 ? kernel_post(c, dot_post) : -1; a.red = c->red;
    a.store = !dot_only;
    const int n = launch_apply(c->dim, a, c->stream);                // y = constraint(A x), column-local
    HMG_CHECK(n >= 0, "apply kernel refused the launch configuration");
    check_launch(c, n);
    if (dot_post >= 0) finish_reduction(c, dot_post, S_TMP);
    if (!dot_only) do_broadcast(c, l, y, part);                      // interface sums
}
// r = broadcast(r) and rho = dot(r, r) over all stored entries without a pass over r: the residual apply
// left the interior part in S_TMP; the interface kernels add (owners x sum^2) of every shared node
void do_broadcast_rho(hmg_ctx* c, int l, double* r) {
    const LevelView& V = c->level(l).view;
    if (c->nranks == 1) {
        const int n = launch_interface_sum_sq(c->dim, V, c->tview, r, c->red, POST_RHO_ADD, c->stream);
        check_launch(c, n);
        if (n == 0) check_launch(c, launch_scalar_post(c->red, POST_RHO, 0, c->stream));     // no shared cell at all
        return;
    }
    const int64_t slots = c->cut_slots(l);
    if (c->peer_on) {
        // pack first (see do_broadcast); the last kernel of the chain adds its part to S_TMP, sums over the ranks and
        // sets rho
        if (slots > 0) cut_pack_peer(c, l, r);
        check_launch(c, launch_interface_sum_sq(c->dim, V, c->tview, r, c->red, POST_ADD, c->stream));
        if (slots > 0) cut_unpack_peer(c, l, r, true, POST_RHO_ADD | POST_GLOBAL);
        else check_launch(c, launch_scalar_post(c->red, POST_RHO | POST_GLOBAL, 0, c->stream));
        return;
    }
    check_launch(c, launch_interface_sum_sq(c->dim, V, c->tview, r, c->red, POST_ADD, c->stream));
    if (slots > 0) do_cut_exchange_impl(c, l, r, true);
    finish_reduction(c, POST_RHO, S_TMP);
}
// p' = r + beta p and Ap = broadcast(constraint(A p')) with p' applied straight out of shared memory: the new
// direction goes to the level's second p buffer (other CTAs still read the old one), then the buffers swap.
// dot_only: p' and p'.Ap only (Ap is neither stored nor summed).
void do_fused_direction_product(hmg_ctx* c, int l, bool dot_only = false, int part = 3) {
    LevelDev& L = c->level(l);
    if (!L.p2) L.p2 = c->dalloc<double>((size_t)c->nstored(l));
    ApplyArgs a;
    a.L = L.view; a.cfg = L.cfg; a.cfg_fused = L.cfg_fused; a.cfg_rhs = L.cfg_rhs;
    a.nunits = c->nunits; a.tab = L.tab.data();
    a.coef = c->elem_coef; a.cmask = c->cmask; a.mult = c->mult;
    a.x = c->vecp(l, HMG_P); a.r2 = c->vecp(l, HMG_R); a.pout = L.p2;
    a.y = c->vecp(l, HMG_AP); a.b = nullptr;
    a.alpha = 1.0; a.lambda = c->lambda; a.mode = APPLY_AX;
    a.dot_post = kernel_post(c, POST_PAP); a.red = c->red;
    a.store = !dot_only;
    const int n = launch_apply(c->dim, a, c->stream);
    HMG_CHECK(n >= 0, "apply kernel refused the fused launch configuration");
    check_launch(c, n);
    std::swap(L.vec[HMG_P], L.p2);
    finish_reduction(c, POST_PAP, S_TMP);
    if (!dot_only) do_broadcast(c, l, c->vecp(l, HMG_AP), part);
}
// smoothing_steps! (src/multigrid.jl:46-71).  need_r = false (inside a V-cycle, wherever nothing reads the residual of
// the last step: before the restriction -- local_residual! recomputes r -- and after the correction on every level
// below the top): the last step shrinks to alpha = rho / p.Ap and x += alpha p; its r-update, rho', the interface sum
// of Ap and Ap itself are dead in the reference too.  x is bit-identical either way.
void do_smoothing(hmg_ctx* c, int l, int steps, bool need_r = true) {
    Range range("hmg smoothing_steps", l);
    const int64_t n = c->nstored(l);
    double *x = c->vecp(l, HMG_X), *r = c->vecp(l, HMG_R), *p = c->vecp(l, HMG_P), *Ap = c->vecp(l, HMG_AP);
    // r = broadcast(constraint(b - A x)), rho = r.r (src/multigrid.jl:50-54); the copy p = r is folded into the
    // first update below, the first product reads r itself
    do_apply(c, l, APPLY_RESIDUAL, 1.0, x, r, c->vecp(l, HMG_B), POST_STORE);
    do_broadcast_rho(c, l, r);
    if (steps == 0) CUDA_OK(cudaMemcpyAsync(p, r, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    const bool fuse = c->level(l).cfg_fused.ring_rows > 0;
    for (int i = 0; i < steps; ++i) {
        const bool x_only = !need_r && i == steps - 1;
        // Ap = broadcast(constraint(A p)), alpha = rho / p.Ap; after the first step the direction update
        // p = r + beta p (src/multigrid.jl:68) happens inside the product
        const double* dir = p;
        if (i == 0) { do_global_product(c, l, r, Ap, POST_PAP, x_only); dir = r; }
        else if (fuse) { do_fused_direction_product(c, l, x_only); p = c->vecp(l, HMG_P); dir = p; }
        else {
            check_launch(c, launch_p_update(c->red, p, r, n, c->stream));
            do_global_product(c, l, p, Ap, POST_PAP, x_only);
        }
        if (x_only) {
            check_launch(c, launch_x_update(c->red, x, dir, n, c->stream));
            break;
        }
        check_launch(c, launch_cg_update(c->red, x, p, r, Ap, n, kernel_post(c, POST_RSQR), i == 0, c->stream));
        finish_reduction(c, POST_RSQR, S_TMP);
        // the reference also updates p after the last step, but that value is never used
        // (the next smoothing call starts from a fresh residual)
    }
}
void assemble_coarse_impl(hmg_ctx* c);
void ensure_coarse(hmg_ctx* c) {
    HMG_CHECK(c->n_interior > 0, "coarse matrix not set: call hmg_set_coarse_matrix or hmg_assemble_coarse first");
    if (c->coarse_gen != c->op_gen) {
        // hmg_set_lambda / hmg_set_sigma after the factorisation: the inverse belongs to another operator
        HMG_CHECK(c->coarse_internal, "lambda or sigma changed after hmg_set_coarse_matrix: hand over the new coarse matrix first");
        assemble_coarse_impl(c);
    }
}
void do_coarse_solve(hmg_ctx* c) {
    Range range("hmg coarse solve", 1);
    ensure_coarse(c);
    LevelDev& L1 = c->level(1);
    do_broadcast(c, 1, c->vecp(1, HMG_B));
    if (c->nranks == 1) {
        check_launch(c, launch_copy_to_base(L1.view, c->nn, c->node_first, c->vecp(1, HMG_B), c->ubase, c->stream));
    } else {
        // every base node is reported by one rank; the sum over ranks lands on rank 0, which solves
        check_launch(c, launch_masked_copy_to_base(L1.view, c->nn, c->node_first, c->node_contrib, c->vecp(1, HMG_B), c->ubase,
                                                   c->stream));
        NCCL_OK(nccl().Reduce(c->ubase, c->ubase, (size_t)c->nn, ncclDouble, ncclSum, 0, c->comm, c->stream));
    }
    if (c->rank == 0) {
        check_launch(c, launch_gather(c->interior_idx, c->n_interior, c->ubase, c->bint, c->stream));
        if (c->symv_work) check_launch(c, launch_symv_half(c->Ainv, c->n_interior, c->bint, c->xint, c->symv_work, c->stream));
        else check_launch(c, launch_symv_full(c->Ainv, c->n_interior, c->bint, c->xint, c->stream));
        check_launch(c, launch_fill(c->ubase, 0.0, c->nn, c->stream));
        check_launch(c, launch_scatter(c->interior_idx, c->n_interior, c->xint, c->ubase, c->stream));
    }
    if (c->nranks > 1) NCCL_OK(nccl().Broadcast(c->ubase, c->ubase, (size_t)c->nn, ncclDouble, 0, c->comm, c->stream));
    check_launch(c, launch_distribute(c->dim, L1.view, c->ne, c->elems32, c->ubase, c->vecp(1, HMG_X), c->stream));
}
void do_vcycle(hmg_ctx* c, int k, int steps, int top) {
    if (k == 1) { do_coarse_solve(c); return; }
    Range range("hmg vcycle", k);
    do_smoothing(c, k, steps, false);
    {
    Range transfer("hmg residual + restrict", k);
    do_local_residual(c, k);
    check_launch(c, launch_restrict(c->dim, c->level(k).view, c->level(k - 1).view, c->nunits, c->vecp(k, HMG_R),
                                    c->vecp(k - 1, HMG_B), c->stream));
    check_launch(c, launch_fill(c->vecp(k - 1, HMG_X), 0.0, c->nstored(k - 1), c->stream));
    }
    do_vcycle(c, k - 1, 2, top);   // the reference does not forward `steps` (src/multigrid.jl:109)
    {
    Range transfer("hmg interpolate", k);
    check_launch(c, launch_interp_add(c->dim, c->level(k).view, c->level(k - 1).view, c->nunits, c->vecp(k, HMG_X),
                                      c->vecp(k - 1, HMG_X), c->stream));
    }
    do_smoothing(c, k, steps, k == top);      // the residual of the top level is the caller's (logged norm)
}
double read_scalar(hmg_ctx* c, int slot) {
    double v = 0.0;
    CUDA_OK(cudaMemcpyAsync(&v, c->red.scalars + slot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return v;
}
void ensure_staging(hmg_ctx* c, size_t bytes) {
    if (!c->copy_stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int q = 0; q < 2; ++q) {
            CUDA_OK(cudaEventCreateWithFlags(&c->ev_copied[q], cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&c->ev_permuted[q], cudaEventDisableTiming));
        }
    }
    if (c->staging_bytes >= bytes) return;
    for (int q = 0; q < 2; ++q) {
        if (c->staging[q]) c->dfree(c->staging[q]);
        c->staging[q] = c->dalloc<double>(bytes / sizeof(double), false);
    }
    c->staging_bytes = bytes;
}
// columns per staged chunk: ~128 MB, whole units
int64_t staging_chunk(hmg_ctx* c, int nf) {
    int64_t chunk = std::max<int64_t>(1, (int64_t)(128 << 20) / ((int64_t)nf * 8));
    chunk = std::max<int64_t>(c->W, chunk / c->W * c->W);
    return std::min<int64_t>(chunk, (c->ne + c->W - 1) / c->W * c->W);
}

// all ranks of a partitioned context meet here (collective): used where one rank does seconds of work the others do
// not -- a kernel of theirs would otherwise spin on its peer mailbox for that long
void rank_barrier(hmg_ctx* c) {
    if (c->nranks == 1) return;
    int* d = c->dalloc<int>(1);
    NCCL_OK(nccl().AllReduce(d, d, 1, ncclInt, ncclSum, c->comm, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->dfree(d);
}
void set_coarse_dense(hmg_ctx* c, int64_t n, const std::vector<int64_t>& colptr, const std::vector<int64_t>& rowval,
                      const std::vector<double>& nzval, const std::vector<int64_t>& interior0) {
    // dense Cholesky inverse on the device: A^-1 is formed once, every V-cycle applies it with one
    // bandwidth-bound symmetric matrix-vector product (replaces CHOLMOD's F \ b, src/multigrid.jl:84)
    if (c->Ainv) { c->dfree(c->Ainv); c->Ainv = nullptr; }
    if (c->interior_idx) { c->dfree(c->interior_idx); c->interior_idx = nullptr; }
    if (c->bint) { c->dfree(c->bint); c->bint = nullptr; }
    if (c->xint) { c->dfree(c->xint); c->xint = nullptr; }
    if (c->symv_work) { c->dfree(c->symv_work); c->symv_work = nullptr; }
    c->n_interior = n;
    if (c->rank != 0) return;            // the coarsest-grid solve stays on rank 0
    {
        // the dense inverse is n^2 doubles (twice that while the 64-bit path forms it): refuse what cannot fit
        size_t free_b = 0, total_b = 0;
        CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
        const bool wide_ = n >= 46340 || getenv("HMG_COARSE_64BIT") != nullptr;
        const double need = (wide_ ? 2.0 : 1.0) * 8.0 * (double)n * (double)n + (double)(256u << 20);
        HMG_CHECK(need <= (double)free_b, "coarse problem too large for the dense coarse solver (the inverse does not fit the device memory left)");
    }
    c->interior_idx = c->dupload(interior0);
    c->bint = c->dalloc<double>(n);
    c->xint = c->dalloc<double>(n);
    c->Ainv = c->dalloc<double>((size_t)n * n);
    const int64_t half_min = getenv("HMG_SYMV_HALF_MIN") ? atoll(getenv("HMG_SYMV_HALF_MIN")) : 2048;
    if (n >= half_min)      // small problems: the one-kernel full mat-vec is faster
        c->symv_work = c->dalloc<double>((size_t)2 * ((n + 127) / 128) * n);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    // scatter the CSC entries column by column through a bounded host buffer
    {
        const int64_t chunk_cols = std::max<int64_t>(1, (int64_t)(64 << 20) / (n * 8));
        std::vector<double> buf;
        for (int64_t c0 = 0; c0 < n; c0 += chunk_cols) {
            const int64_t nc = std::min(chunk_cols, n - c0);
            buf.assign((size_t)nc * n, 0.0);
            for (int64_t j = 0; j < nc; ++j)
                for (int64_t q = colptr[c0 + j]; q < colptr[c0 + j + 1]; ++q) buf[(size_t)j * n + rowval[q]] += nzval[q];
            CUDA_OK(cudaMemcpy(c->Ainv + (size_t)c0 * n, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    // the handle (and the parameter block of the 64-bit interface) must not outlive a failed check
    struct Solver {
        cusolverDnHandle_t h = nullptr;
        cusolverDnParams_t params = nullptr;
        ~Solver() {
            if (params) cusolverDnDestroyParams(params);
            if (h) cusolverDnDestroy(h);
        }
    } sv;
    CUSOLVER_OK(cusolverDnCreate(&sv.h));
    CUSOLVER_OK(cusolverDnSetStream(sv.h, c->stream));
    int* info = c->dalloc<int>(1);
    int hinfo = 0;
    auto read_info = [&]() {
        CUDA_OK(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        return hinfo;
    };
    // potri indexes the matrix with 32-bit integers: from n = 46 340 (n^2 >= 2^31) on -- or on request, for the tests --
    // the inverse is formed through the 64-bit interface instead: A = L L', then L L' X = I.  It costs a second n x n
    // matrix during the setup (the factor is freed afterwards) and 2 n^3 instead of 2/3 n^3 operations.
    const bool wide = n >= 46340 || getenv("HMG_COARSE_64BIT") != nullptr;
    if (!wide) {
        int lwork1 = 0, lwork2 = 0;
        CUSOLVER_OK(cusolverDnDpotrf_bufferSize(sv.h, CUBLAS_FILL_MODE_LOWER, (int)n, c->Ainv, (int)n, &lwork1));
        CUSOLVER_OK(cusolverDnDpotri_bufferSize(sv.h, CUBLAS_FILL_MODE_LOWER, (int)n, c->Ainv, (int)n, &lwork2));
        const int lwork = std::max(lwork1, lwork2);
        double* work = c->dalloc<double>(lwork, false);
        CUSOLVER_OK(cusolverDnDpotrf(sv.h, CUBLAS_FILL_MODE_LOWER, (int)n, c->Ainv, (int)n, work, lwork, info));
        HMG_CHECK(read_info() == 0, "coarse matrix is not positive definite (potrf failed)");
        CUSOLVER_OK(cusolverDnDpotri(sv.h, CUBLAS_FILL_MODE_LOWER, (int)n, c->Ainv, (int)n, work, lwork, info));
        HMG_CHECK(read_info() == 0, "potri failed on the coarse matrix");
        c->dfree(work);
    } else {
        CUSOLVER_OK(cusolverDnCreateParams(&sv.params));
        size_t wdev = 0, whost = 0;
        CUSOLVER_OK(cusolverDnXpotrf_bufferSize(sv.h, sv.params, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, c->Ainv, n, CUDA_R_64F,
                                                &wdev, &whost));
        char* dwork = c->dalloc<char>(wdev, false);
        std::vector<char> hwork(whost + 1);
        CUSOLVER_OK(cusolverDnXpotrf(sv.h, sv.params, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, c->Ainv, n, CUDA_R_64F, dwork, wdev,
                                     hwork.data(), whost, info));
        HMG_CHECK(read_info() == 0, "coarse matrix is not positive definite (potrf failed)");
        c->dfree(dwork);
        double* inv = c->dalloc<double>((size_t)n * n);           // zeroed
        check_launch(c, launch_set_diagonal(inv, n, 1.0, c->stream));
        CUSOLVER_OK(cusolverDnXpotrs(sv.h, sv.params, CUBLAS_FILL_MODE_LOWER, n, n, CUDA_R_64F, c->Ainv, n, CUDA_R_64F, inv, n, info));
        HMG_CHECK(read_info() == 0, "potrs failed on the coarse matrix");
        c->dfree(c->Ainv);
        c->Ainv = inv;
    }
    // exactly symmetric from the lower triangle (the half-traffic mat-vec reads a tile as T and as T')
    check_launch(c, launch_symmetrize_lower(c->Ainv, n, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->dfree(info);
}

void assemble_coarse_impl(hmg_ctx* c) {
    // P1 assembly of lambda*M + K(sigma) on the base mesh, restricted to the interior nodes
    const int dim = c->dim, nv = dim + 1, cs = dim == 3 ? 8 : 4, nc = dim == 3 ? 7 : 4;
    std::vector<double> coef;            // of ALL elements: the coarse operator couples the whole base mesh
    element_coefficients(dim, c->ne_global, c->nodes.data(), c->elems_global.data(), c->sigma_global.data(), coef, cs);
    const std::vector<int64_t>& interior = c->topo.interior_nodes;
    const int64_t n = (int64_t)interior.size();
    HMG_CHECK(n > 0, "base mesh has no interior nodes");
    std::vector<int64_t> pos(c->nn, -1);
    for (int64_t q = 0; q < n; ++q) pos[interior[q]] = q;
    const double fact = dim == 3 ? 6.0 : 2.0;
    const double gref[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    std::vector<std::map<int64_t, double>> cols(n);
    c->coarse_internal = true;
    c->coarse_gen = c->op_gen;
    c->drop_graphs();
    if (c->rank != 0) { c->n_interior = n; rank_barrier(c); return; }      // wait for rank 0's factorisation
    for (int64_t e = 0; e < c->ne_global; ++e) {
        const double* ec = &coef[(size_t)e * cs];
        double P[3][3];
        int q = 0;
        for (int k = 0; k < dim; ++k)
            for (int l = k; l < dim; ++l, ++q) P[k][l] = P[l][k] = ec[q];
        const double detJ = ec[nc - 1];
        for (int a = 0; a < nv; ++a) {
            const int64_t ra = pos[c->elems_global[e * nv + a]];
            if (ra < 0) continue;
            for (int b = 0; b < nv; ++b) {
                const int64_t rb = pos[c->elems_global[e * nv + b]];
                if (rb < 0) continue;
                double s = 0.0;
                for (int k = 0; k < dim; ++k)
                    for (int l = 0; l < dim; ++l) s += gref[a][k] * P[k][l] * gref[b][l];
                const double mass = detJ * (a == b ? 2.0 : 1.0) / (fact * (dim + 1) * (dim + 2));
                cols[rb][ra] += s / fact + c->lambda * mass;
            }
        }
    }
    std::vector<int64_t> cp(n + 1, 0), rv;
    std::vector<double> nz;
    for (int64_t j = 0; j < n; ++j) {
        for (const auto& kv : cols[j]) { rv.push_back(kv.first); nz.push_back(kv.second); }
        cp[j + 1] = (int64_t)rv.size();
    }
    set_coarse_dense(c, n, cp, rv, nz, interior);
    rank_barrier(c);
}

}  // namespace

namespace hmg { void set_last_error(const std::string& msg) { g_err = msg; } }   // field.cu

#define HMG_API_BEGIN try {
#define HMG_API_END                                  \
    return 0;                                        \
    }                                                \
    catch (const std::exception& ex) {               \
        g_err = ex.what();                           \
        return 1;                                    \
    }                                                \
    catch (...) {                                    \
        g_err = "hmg: unknown error";                \
        return 1;                                    \
    }
#define NEED_CTX(c) HMG_CHECK((c) != nullptr, "null context")

extern "C" {

const char* hmg_last_error(void) { return g_err.c_str(); }
int hmg_version(void) { return 100; }

int hmg_create(int dim, int nlevels, int64_t ne, int64_t nn, const double* base_nodes, const int64_t* base_elems,
               const double* sigma, double lambda, int device, hmg_ctx** out) {
    HMG_API_BEGIN
    HMG_CHECK(out != nullptr, "null output pointer");
    *out = nullptr;
    *out = create_impl(dim, nlevels, ne, nn, base_nodes, base_elems, sigma, lambda, device, 0, 1, nullptr, nullptr);
    HMG_API_END
}

int hmg_create_partitioned(int dim, int nlevels, int64_t ne, int64_t nn, const double* base_nodes, const int64_t* base_elems,
                           const double* sigma, double lambda, int device, int rank, int nranks, const int32_t* owner_rank,
                           const void* nccl_id, hmg_ctx** out) {
    HMG_API_BEGIN
    HMG_CHECK(out != nullptr, "null output pointer");
    *out = nullptr;
    *out = create_impl(dim, nlevels, ne, nn, base_nodes, base_elems, sigma, lambda, device, rank, nranks, owner_rank, nccl_id);
    HMG_API_END
}
int hmg_nccl_unique_id(void* out128) {
    HMG_API_BEGIN
    HMG_CHECK(out128 != nullptr, "null output pointer");
    ncclUniqueId id;
    NCCL_OK(nccl().GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof(id));
    HMG_API_END
}

int hmg_destroy(hmg_ctx* c) {
    HMG_API_BEGIN
    if (!c) return 0;
    delete c;
    HMG_API_END
}

int64_t hmg_nf(const hmg_ctx* c, int level) {
    if (!c || level < 1 || level > c->nlevels) return -1;
    return c->lv[level - 1].view.nf;
}
int64_t hmg_ne_local(const hmg_ctx* c) { return c ? c->ne : -1; }
int64_t hmg_ld(const hmg_ctx* c, int level) {
    if (!c || level < 1 || level > c->nlevels) return -1;
    return c->lv[level - 1].view.nf;
}
int hmg_group_width(const hmg_ctx* c) { return c ? c->W : -1; }
int hmg_comm_mode(const hmg_ctx* c) { return !c ? -1 : (c->nranks == 1 ? 0 : (c->peer_on ? 2 : 1)); }
int hmg_local_elements(const hmg_ctx* c, int64_t* out) {
    HMG_API_BEGIN
    NEED_CTX(c);
    for (int64_t e = 0; e < c->ne; ++e) out[e] = c->local_to_global[e] + 1;
    HMG_API_END
}

int hmg_set_lambda(hmg_ctx* c, double lambda) {
    HMG_API_BEGIN
    NEED_CTX(c);
    if (lambda != c->lambda) { ++c->op_gen; c->drop_graphs(); }      // the coarse inverse (if any) belongs to the old operator
    c->lambda = lambda;
    HMG_API_END
}
int hmg_set_sigma(hmg_ctx* c, const double* sigma) {
    HMG_API_BEGIN
    NEED_CTX(c);
    HMG_CHECK(sigma != nullptr, "null sigma");
    CUDA_OK(cudaSetDevice(c->device));
    upload_operator(c, sigma);
    ++c->op_gen;
    c->drop_graphs();
    HMG_API_END
}

int hmg_upload(hmg_ctx* c, int level, int which, const double* host, int64_t ld_host) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    LevelDev& L = c->level(level);
    const int nf = L.view.nf;
    HMG_CHECK(host != nullptr && ld_host >= nf, "bad host matrix");
    double* dst = c->vecp(level, which);
    const int64_t chunk = staging_chunk(c, nf);
    ensure_staging(c, (size_t)chunk * nf * 8);
    // copy stream: host -> staging[i & 1]; main stream: staging -> device layout.  Events order the two.
    CUDA_OK(cudaStreamSynchronize(c->stream));
    int i = 0;
    for (int64_t c0 = 0; c0 < c->ne; c0 += chunk, ++i) {
        const int64_t nc = std::min(chunk, c->ne - c0);
        const int q = i & 1;
        if (i >= 2) CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_permuted[q], 0));
        CUDA_OK(cudaMemcpy2DAsync(c->staging[q], (size_t)nf * 8, host + c0 * ld_host, (size_t)ld_host * 8, (size_t)nf * 8,
                                  (size_t)nc, cudaMemcpyHostToDevice, c->copy_stream));
        CUDA_OK(cudaEventRecord(c->ev_copied[q], c->copy_stream));
        CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copied[q], 0));
        check_launch(c, launch_permute_in(L.view, L.hier2lat, c->staging[q], nf, dst, c0, nc, c->stream));
        CUDA_OK(cudaEventRecord(c->ev_permuted[q], c->stream));
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    HMG_API_END
}

int hmg_download(hmg_ctx* c, int level, int which, double* host, int64_t ld_host) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    LevelDev& L = c->level(level);
    const int nf = L.view.nf;
    HMG_CHECK(host != nullptr && ld_host >= nf, "bad host matrix");
    const double* src = c->vecp(level, which);
    const int64_t chunk = staging_chunk(c, nf);
    ensure_staging(c, (size_t)chunk * nf * 8);
    // main stream: device layout -> staging[i & 1]; copy stream: staging -> host
    int i = 0;
    for (int64_t c0 = 0; c0 < c->ne; c0 += chunk, ++i) {
        const int64_t nc = std::min(chunk, c->ne - c0);
        const int q = i & 1;
        if (i >= 2) CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copied[q], 0));
        check_launch(c, launch_permute_out(L.view, L.hier2lat, src, c->staging[q], nf, c0, nc, c->stream));
        CUDA_OK(cudaEventRecord(c->ev_permuted[q], c->stream));
        CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_permuted[q], 0));
        CUDA_OK(cudaMemcpy2DAsync(host + c0 * ld_host, (size_t)ld_host * 8, c->staging[q], (size_t)nf * 8, (size_t)nf * 8,
                                  (size_t)nc, cudaMemcpyDeviceToHost, c->copy_stream));
        CUDA_OK(cudaEventRecord(c->ev_copied[q], c->copy_stream));
    }
    CUDA_OK(cudaStreamSynchronize(c->copy_stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    HMG_API_END
}

int hmg_download_rows(hmg_ctx* c, int level, int which, int64_t nrows, double* host, int64_t ld_host) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    LevelDev& L = c->level(level);
    const int nf = L.view.nf;
    HMG_CHECK(nrows >= 0 && nrows <= nf, "download_rows: more rows than the level has nodes");
    if (nrows == 0 || c->ne == 0) return 0;
    HMG_CHECK(host != nullptr && ld_host >= nrows, "bad host matrix");
    const double* src = c->vecp(level, which);
    // chunks of ~128 MB of *output*; the staging buffers are shared with hmg_upload / hmg_download
    int64_t chunk = std::max<int64_t>(c->W, (int64_t)(128 << 20) / (nrows * 8) / c->W * c->W);
    chunk = std::min<int64_t>(chunk, (c->ne + c->W - 1) / c->W * c->W);
    ensure_staging(c, (size_t)chunk * nrows * 8);
    int i = 0;
    for (int64_t c0 = 0; c0 < c->ne; c0 += chunk, ++i) {
        const int64_t nc = std::min(chunk, c->ne - c0);
        const int q = i & 1;
        if (i >= 2) CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copied[q], 0));
        check_launch(c, launch_permute_rows_out(L.view, L.hier2lat, src, c->staging[q], nrows, (int)nrows, c0, nc, c->stream));
        CUDA_OK(cudaEventRecord(c->ev_permuted[q], c->stream));
        CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_permuted[q], 0));
        CUDA_OK(cudaMemcpy2DAsync(host + c0 * ld_host, (size_t)ld_host * 8, c->staging[q], (size_t)nrows * 8, (size_t)nrows * 8,
                                  (size_t)nc, cudaMemcpyDeviceToHost, c->copy_stream));
        CUDA_OK(cudaEventRecord(c->ev_copied[q], c->copy_stream));
    }
    CUDA_OK(cudaStreamSynchronize(c->copy_stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    HMG_API_END
}

int hmg_copy_columns_from(hmg_ctx* dst, int level, int dst_which, hmg_ctx* src, int src_which) {
    HMG_API_BEGIN
    NEED_CTX(dst);
    NEED_CTX(src);
    HMG_CHECK(dst != src, "copy_columns_from: source and destination are the same context");
    HMG_CHECK(dst->nranks == 1 && src->nranks == 1, "copy_columns_from: partitioned contexts keep different column sets");
    HMG_CHECK(dst->dim == src->dim && dst->device == src->device && dst->W == src->W, "copy_columns_from: contexts do not match");
    HMG_CHECK(dst->level(level).view.nf == src->level(level).view.nf, "copy_columns_from: the level has a different shape");
    HMG_CHECK(dst->ne <= src->ne, "copy_columns_from: the destination has more columns than the source");
    CUDA_OK(cudaSetDevice(dst->device));
    const double* from = src->vecp(level, src_which);
    double* to = dst->vecp(level, dst_which);
    CUDA_OK(cudaStreamSynchronize(src->stream));      // the source vector is complete before the other stream reads it
    check_launch(dst, launch_copy_columns(dst->level(level).view, dst->ne, to, from, dst->stream));
    CUDA_OK(cudaStreamSynchronize(dst->stream));      // the caller may destroy `src` right away
    HMG_API_END
}

int hmg_fill(hmg_ctx* c, int level, int which, double value) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    if (value == 0.0) {
        check_launch(c, launch_fill(c->vecp(level, which), 0.0, c->nstored(level), c->stream));
    } else {
        // the zero columns padding the last unit must stay zero
        check_launch(c, launch_fill_columns(c->level(level).view, c->ne, c->vecp(level, which), value, c->stream));
    }
    HMG_API_END
}
int hmg_copy(hmg_ctx* c, int level, int dst, int src) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaMemcpyAsync(c->vecp(level, dst), c->vecp(level, src), c->nstored(level) * sizeof(double),
                            cudaMemcpyDeviceToDevice, c->stream));
    HMG_API_END
}
int hmg_axpy(hmg_ctx* c, int level, double alpha, int x, int y) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    check_launch(c, launch_axpy(alpha, c->vecp(level, x), c->vecp(level, y), c->nstored(level), c->stream));
    HMG_API_END
}
int hmg_dot(hmg_ctx* c, int level, int a, int b, double* out) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    check_launch(c, launch_dot(c->red, c->vecp(level, a), c->vecp(level, b), c->nstored(level), kernel_post(c, POST_STORE), S_TMP, c->stream));
    finish_reduction(c, POST_STORE, S_TMP);
    *out = read_scalar(c, S_TMP);
    HMG_API_END
}

int hmg_mul(hmg_ctx* c, int level, double alpha, int x, int y) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(x != y, "mul!: x and y must be different vectors");
    do_apply(c, level, APPLY_MULADD, alpha, c->vecp(level, x), c->vecp(level, y), nullptr);
    HMG_API_END
}
int hmg_apply_global(hmg_ctx* c, int level, int x, int y) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(x != y, "apply: x and y must be different vectors");
    do_global_product(c, level, c->vecp(level, x), c->vecp(level, y));
    HMG_API_END
}
int hmg_apply_constraint(hmg_ctx* c, int level, int which) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    check_launch(c, launch_apply_constraint(c->dim, c->level(level).view, c->nbelems, c->belems, c->cmask,
                                            c->vecp(level, which), c->stream));
    HMG_API_END
}
int hmg_broadcast_interfaces(hmg_ctx* c, int level, int which) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    do_broadcast(c, level, c->vecp(level, which));
    HMG_API_END
}
int hmg_zero_out_all_but_one(hmg_ctx* c, int level, int which) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    do_zero_all_but_one(c, level, c->vecp(level, which));
    HMG_API_END
}
int hmg_local_residual(hmg_ctx* c, int level) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    do_local_residual(c, level);
    HMG_API_END
}
int hmg_restrict(hmg_ctx* c, int k) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(k >= 2 && k <= c->nlevels, "restrict: level must be in 2..nlevels");
    check_launch(c, launch_restrict(c->dim, c->level(k).view, c->level(k - 1).view, c->nunits, c->vecp(k, HMG_R),
                                    c->vecp(k - 1, HMG_B), c->stream));
    HMG_API_END
}
int hmg_interpolate_add(hmg_ctx* c, int k) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(k >= 2 && k <= c->nlevels, "interpolate: level must be in 2..nlevels");
    check_launch(c, launch_interp_add(c->dim, c->level(k).view, c->level(k - 1).view, c->nunits, c->vecp(k, HMG_X),
                                      c->vecp(k - 1, HMG_X), c->stream));
    HMG_API_END
}
int hmg_smoothing_steps(hmg_ctx* c, int level, int steps) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(steps >= 0, "negative number of smoothing steps");
    do_smoothing(c, level, steps);
    HMG_API_END
}

int hmg_set_coarse_matrix(hmg_ctx* c, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                          const int64_t* interior_nodes) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(n > 0 && colptr && rowval && nzval && interior_nodes, "bad coarse matrix");
    std::vector<int64_t> cp(colptr, colptr + n + 1), in0(interior_nodes, interior_nodes + n);
    for (auto& v : cp) v -= 1;
    for (auto& v : in0) { v -= 1; HMG_CHECK(v >= 0 && v < c->nn, "interior node out of range"); }
    std::vector<int64_t> rv(rowval, rowval + cp[n]);
    for (auto& v : rv) { v -= 1; HMG_CHECK(v >= 0 && v < n, "coarse matrix row out of range"); }
    std::vector<double> nz(nzval, nzval + cp[n]);
    set_coarse_dense(c, n, cp, rv, nz, in0);
    c->coarse_internal = false;
    c->coarse_gen = c->op_gen;
    c->drop_graphs();
    rank_barrier(c);                       // (collective on a partitioned context: rank 0 factorises, the others wait)
    HMG_API_END
}

int hmg_assemble_coarse(hmg_ctx* c) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    assemble_coarse_impl(c);
    HMG_API_END
}

int hmg_copy_to_base(hmg_ctx* c, int which, double* u_host) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    if (c->nranks == 1) {
        check_launch(c, launch_copy_to_base(c->level(1).view, c->nn, c->node_first, c->vecp(1, which), c->ubase, c->stream));
    } else {      // every rank receives the whole base vector
        check_launch(c, launch_masked_copy_to_base(c->level(1).view, c->nn, c->node_first, c->node_contrib, c->vecp(1, which),
                                                   c->ubase, c->stream));
        NCCL_OK(nccl().AllReduce(c->ubase, c->ubase, (size_t)c->nn, ncclDouble, ncclSum, c->comm, c->stream));
    }
    CUDA_OK(cudaMemcpyAsync(u_host, c->ubase, c->nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    HMG_API_END
}
int hmg_distribute(hmg_ctx* c, int which, const double* u_host) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaMemcpyAsync(c->ubase, u_host, c->nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    check_launch(c, launch_distribute(c->dim, c->level(1).view, c->ne, c->elems32, c->ubase, c->vecp(1, which), c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    HMG_API_END
}

static void vcycle_with_norm(hmg_ctx* c, int top, int steps, bool want_norm, int slot) {
    do_vcycle(c, top, steps, top);
    if (want_norm) {
        double* r = c->vecp(top, HMG_R);
        do_zero_all_but_one(c, top, r);
        check_launch(c, launch_dot(c->red, r, r, c->nstored(top), kernel_post(c, POST_STORE), S_TMP, c->stream));
        finish_reduction(c, POST_STORE, S_TMP);
        CUDA_OK(cudaMemcpyAsync(c->red.scalars + slot, c->red.scalars + S_TMP, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
}

// One V-cycle (+ the logged norm) through a CUDA graph of its ~150 launches
static void run_vcycle(hmg_ctx* c, int top, int steps, bool want_norm) {
    if (top >= 2) ensure_coarse(c);                 // never inside a capture: it allocates and synchronises
    // several ranks: with peer memory only the coarse-level ncclReduce / ncclBroadcast are NCCL calls inside the
    // capture (measured on 2 and 8 GPUs: 22.77 vs 22.96 ms per V-cycle of C4 on 8); the all-NCCL path is captured only
    // on request (HMG_GRAPH=2)
    const bool use_graph = top >= 2 && (c->nranks == 1 || c->peer_on ? c->graph_mode >= 1 : c->graph_mode >= 2);
    if (!use_graph) { vcycle_with_norm(c, top, steps, want_norm, S_NRM); return; }
    hmg_ctx::CycleGraph& G = c->graphs[std::make_tuple(top, steps, (int)want_norm)];
    if (!G.warmed) {                                // lazy buffers, function attributes: once, eagerly
        vcycle_with_norm(c, top, steps, want_norm, S_NRM);
        G.warmed = true;
        return;
    }
    if (!G.exec) {
        const int64_t before = c->launches;
        cudaGraph_t graph = nullptr;
        CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        try {
            vcycle_with_norm(c, top, steps, want_norm, S_NRM);
        } catch (...) {
            cudaStreamEndCapture(c->stream, &graph);
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            c->graph_mode = 0;
            throw;
        }
        CUDA_OK(cudaStreamEndCapture(c->stream, &graph));
        G.launches = c->launches - before;
        c->launches = before;
        const cudaError_t e = cudaGraphInstantiate(&G.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {                     // stay eager
            cudaGetLastError();
            G.exec = nullptr;
            c->graph_mode = 0;
            vcycle_with_norm(c, top, steps, want_norm, S_NRM);
            return;
        }
    }
    CUDA_OK(cudaGraphLaunch(G.exec, c->stream));
    c->launches += G.launches;
}

int hmg_vcycle(hmg_ctx* c, int top_level, int steps, double* out_resnorm) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(top_level >= 1 && top_level <= c->nlevels, "level out of range");
    run_vcycle(c, top_level, steps, out_resnorm != nullptr);
    if (out_resnorm) *out_resnorm = std::sqrt(read_scalar(c, S_NRM));
    HMG_API_END
}
int hmg_vcycles(hmg_ctx* c, int top_level, int steps, int ncycles, double* resnorms) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(top_level >= 1 && top_level <= c->nlevels, "level out of range");
    HMG_CHECK(ncycles >= 0, "negative cycle count");
    double* dn = nullptr;
    if (resnorms) dn = c->dalloc<double>(ncycles);
    for (int i = 0; i < ncycles; ++i) {
        run_vcycle(c, top_level, steps, resnorms != nullptr);
        if (resnorms)
            CUDA_OK(cudaMemcpyAsync(dn + i, c->red.scalars + S_NRM, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    if (resnorms) {
        CUDA_OK(cudaMemcpyAsync(resnorms, dn, ncycles * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < ncycles; ++i) resnorms[i] = std::sqrt(resnorms[i]);
        c->dfree(dn);
    }
    HMG_API_END
}

// ---- driver functionals on the finest level ---------------------------
namespace {
void ensure_functional_tables(hmg_ctx* c) {
    if (c->dphi) return;
    const int dim = c->dim, cs = dim == 3 ? 8 : 4, nc = dim == 3 ? 7 : 4, W = c->W;
    c->dphi = c->dupload(c->ref.lv[c->nlevels - 1].dphi);
    std::vector<double> m1((size_t)c->nunits * cs * W, 0.0), mJ((size_t)c->nunits * cs * W, 0.0);
    std::vector<int32_t> gi((size_t)c->nunits * W, -1);
    for (int64_t e = 0; e < c->ne; ++e) {
        const size_t at = ((size_t)(e / W) * cs + (nc - 1)) * W + e % W;
        m1[at] = 1.0;
        mJ[at] = c->elem_coef_host[(size_t)e * cs + nc - 1];
        gi[e] = (int32_t)c->local_to_global[e];
    }
    c->coef_mass1 = c->dupload(m1);
    c->coef_massJ = c->dupload(mJ);
    c->gidx = c->dupload(gi);
    c->flux = c->dalloc<double>((size_t)c->nunits * dim * W);
}
void upload_flux(hmg_ctx* c, const double* xi) {
    const int dim = c->dim, W = c->W;
    std::vector<double> sl((size_t)c->ne * dim), fl;
    for (int64_t e = 0; e < c->ne; ++e)
        for (int d = 0; d < dim; ++d) sl[(size_t)e * dim + d] = c->sigma_global[(size_t)c->local_to_global[e] * dim + d];
    element_flux_vectors(dim, c->ne, c->nodes.data(), c->elems.data(), sl.data(), xi, fl);
    std::vector<double> inter((size_t)c->nunits * dim * W, 0.0);
    for (int64_t e = 0; e < c->ne; ++e)
        for (int d = 0; d < dim; ++d) inter[((size_t)(e / W) * dim + d) * W + e % W] = fl[(size_t)e * dim + d];
    CUDA_OK(cudaMemcpyAsync(c->flux, inter.data(), inter.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
}
// y = (reference mass matrix or |J| * lambda * it) * x, column-local, through the apply kernel
void mass_apply(hmg_ctx* c, const double* coef, double lambda, const double* x, double* y) {
    const int l = c->nlevels;
    LevelDev& L = c->level(l);
    check_launch(c, launch_fill(y, 0.0, c->nstored(l), c->stream));
    ApplyArgs a;
    a.L = L.view; a.cfg = L.cfg; a.cfg_rhs = L.cfg_rhs; a.nunits = c->nunits; a.tab = L.tab.data();
    a.coef = coef; a.cmask = c->cmask; a.mult = c->mult;
    a.x = x; a.y = y; a.b = nullptr;
    a.alpha = 1.0; a.lambda = lambda; a.mode = APPLY_MULADD; a.dot_post = -1; a.red = c->red;
    const int n = launch_apply(c->dim, a, c->stream);
    HMG_CHECK(n >= 0, "apply kernel refused the launch configuration");
    check_launch(c, n);
}
double integrate(hmg_ctx* c, const double* v, const double* v2, const double* flux, int64_t nsubset) {
    const int l = c->nlevels;
    HMG_CHECK(nsubset >= 0 && nsubset <= c->ne_global, "element subset out of range");
    double* Mv = c->vecp(l, HMG_W);
    HMG_CHECK(v != Mv && v2 != Mv, "the work vector w cannot be an argument of an integral");
    mass_apply(c, c->coef_mass1, 1.0, v, Mv);
    check_launch(c, launch_integrate(c->dim, c->red, c->level(l).view, c->nunits, nsubset, c->gidx, c->elem_coef, c->dphi,
                                     flux, v, v2, Mv, c->stream));
    finish_reduction(c, POST_STORE, S_TMP, false);
    return read_scalar(c, S_TMP);
}
}  // namespace

int hmg_rhs_axi_grad(hmg_ctx* c, const double* xi, int which_b) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(xi != nullptr, "null xi");
    ensure_functional_tables(c);
    upload_flux(c, xi);
    const int l = c->nlevels;
    check_launch(c, launch_rhs_flux(c->dim, c->level(l).view, c->nunits, c->dphi, c->flux, c->vecp(l, which_b), c->stream));
    HMG_API_END
}
int hmg_integrate_first_term(hmg_ctx* c, int which_v, const double* xi, int64_t nsubset, double* out) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(xi != nullptr && out != nullptr, "null argument");
    ensure_functional_tables(c);
    upload_flux(c, xi);
    *out = integrate(c, c->vecp(c->nlevels, which_v), nullptr, c->flux, nsubset);
    HMG_API_END
}
int hmg_integrate_terms(hmg_ctx* c, int which_vk, int which_vkm1, int64_t nsubset, double* out) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(out != nullptr, "null argument");
    ensure_functional_tables(c);
    *out = integrate(c, c->vecp(c->nlevels, which_vk), c->vecp(c->nlevels, which_vkm1), nullptr, nsubset);
    HMG_API_END
}
int hmg_integrate_area(hmg_ctx* c, int64_t nsubset, double* out) {
    HMG_API_BEGIN
    NEED_CTX(c);
    HMG_CHECK(out != nullptr, "null argument");
    HMG_CHECK(nsubset >= 0 && nsubset <= c->ne_global, "element subset out of range");
    // 1' M 1 under the selected base cells: sum(mass) * |J_e|, summed in element order like the reference
    const int dim = c->dim, cs = dim == 3 ? 8 : 4, nc = dim == 3 ? 7 : 4;
    std::vector<double> coef;
    element_coefficients(dim, nsubset, c->nodes.data(), c->elems_global.data(), c->sigma_global.data(), coef, cs);
    const double mtot = c->ref.lv[c->nlevels - 1].mass_total;
    double area = 0.0;
    for (int64_t e = 0; e < nsubset; ++e) area += mtot * coef[(size_t)e * cs + nc - 1];
    *out = area;
    HMG_API_END
}
int hmg_next_rhs(hmg_ctx* c, int which_b, int which_x) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(which_b != which_x, "next_rhs!: b and x must be different vectors");
    ensure_functional_tables(c);
    mass_apply(c, c->coef_massJ, c->lambda, c->vecp(c->nlevels, which_x), c->vecp(c->nlevels, which_b));
    HMG_API_END
}

int hmg_synchronize(hmg_ctx* c) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    HMG_API_END
}

int hmg_time_op(hmg_ctx* c, int op, int level, int steps, int reps, float* ms_out) {
    HMG_API_BEGIN
    NEED_CTX(c);
    CUDA_OK(cudaSetDevice(c->device));
    HMG_CHECK(reps > 0 && ms_out, "bad arguments");
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i) {
        if (op == 0) {
            do_global_product(c, level, c->vecp(level, HMG_P), c->vecp(level, HMG_AP));
        } else if (op == 1) {
            run_vcycle(c, level, steps, false);
        } else if (op == 2) {
            do_apply(c, level, APPLY_MULADD, 1.0, c->vecp(level, HMG_P), c->vecp(level, HMG_AP), nullptr);
        } else if (op == 3) {
            do_apply(c, level, APPLY_AX, 1.0, c->vecp(level, HMG_P), c->vecp(level, HMG_AP), nullptr);
        } else if (op == 4) {
            do_broadcast(c, level, c->vecp(level, HMG_AP));
        } else if (op == 13 || op == 14) {
            check_launch(c, launch_interface_sum(c->dim, c->level(level).view, c->tview, c->vecp(level, HMG_AP), c->stream, op - 12));
        } else if (op == 5) {
            do_apply(c, level, APPLY_RESIDUAL, 1.0, c->vecp(level, HMG_X), c->vecp(level, HMG_R), c->vecp(level, HMG_B));
        } else if (op == 6) {
            check_launch(c, launch_cg_update(c->red, c->vecp(level, HMG_X), c->vecp(level, HMG_P), c->vecp(level, HMG_R),
                                             c->vecp(level, HMG_AP), c->nstored(level), POST_RSQR, false, c->stream));
        } else if (op == 7) {
            check_launch(c, launch_p_update(c->red, c->vecp(level, HMG_P), c->vecp(level, HMG_R), c->nstored(level), c->stream));
        } else if (op == 8) {
            check_launch(c, launch_copy_dot(c->red, c->vecp(level, HMG_R), c->vecp(level, HMG_P), c->nstored(level), POST_RHO, c->stream));
        } else if (op == 9) {
            check_launch(c, launch_restrict(c->dim, c->level(level).view, c->level(level - 1).view, c->nunits,
                                            c->vecp(level, HMG_R), c->vecp(level - 1, HMG_B), c->stream));
        } else if (op == 10) {
            check_launch(c, launch_interp_add(c->dim, c->level(level).view, c->level(level - 1).view, c->nunits,
                                              c->vecp(level, HMG_X), c->vecp(level - 1, HMG_X), c->stream));
        } else if (op == 12) {
            HMG_CHECK(c->level(level).cfg_fused.ring_rows > 0, "the fused p-update does not fit this level");
            do_fused_direction_product(c, level);
        } else if (op == 17) {
            // one scalar sum over all stored entries AND all ranks (the reduction pattern of every CG scalar)
            check_launch(c, launch_dot(c->red, c->vecp(level, HMG_P), c->vecp(level, HMG_AP), c->nstored(level),
                                       kernel_post(c, POST_STORE), S_TMP, c->stream));
            finish_reduction(c, POST_STORE, S_TMP);
        } else if (op == 18) {
            do_cut_exchange(c, level, c->vecp(level, HMG_AP));       // pack, exchange with the neighbours, unpack
        } else if (op == 16) {
            check_launch(c, launch_x_update(c->red, c->vecp(level, HMG_X), c->vecp(level, HMG_P), c->nstored(level), c->stream));
        } else if (op == 11) {
            do_apply(c, level, APPLY_AX, 1.0, c->vecp(level, HMG_P), c->vecp(level, HMG_AP), nullptr, POST_STORE);
        } else {
            throw Error("hmg: unknown op for hmg_time_op");
        }
    }
    CUDA_OK(cudaEventRecord(c->ev1, c->stream));
    CUDA_OK(cudaEventSynchronize(c->ev1));
    CUDA_OK(cudaEventElapsedTime(ms_out, c->ev0, c->ev1));
    HMG_API_END
}

int hmg_refined_mesh(const hmg_ctx* c, int level, double* nodes, int64_t* elems1, int64_t* nel) {
    HMG_API_BEGIN
    NEED_CTX(c);
    HMG_CHECK(level >= 1 && level <= c->nlevels, "level out of range");
    const RefLevel& L = c->ref.lv[level - 1];
    const int dim = c->dim;
    if (nel) *nel = (int64_t)L.cells.size() / (dim + 1);
    if (nodes)
        for (int n = 0; n < L.nf; ++n)
            for (int d = 0; d < dim; ++d) nodes[(size_t)n * dim + d] = (double)L.hier_coords[(size_t)n * 3 + d] / (double)L.m;
    if (elems1)
        for (size_t q = 0; q < L.cells.size(); ++q) elems1[q] = (int64_t)L.cells[q] + 1;
    HMG_API_END
}
int64_t hmg_launch_count(const hmg_ctx* c) { return c ? c->launches : -1; }
void* hmg_device_ptr(hmg_ctx* c, int level, int which) {
    try {
        if (!c) return nullptr;
        cudaSetDevice(c->device);
        return c->vecp(level, which);
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return nullptr;
    }
}
int hmg_hier_to_lattice(const hmg_ctx* c, int level, int32_t* out) {
    HMG_API_BEGIN
    NEED_CTX(c);
    HMG_CHECK(level >= 1 && level <= c->nlevels && out, "bad arguments");
    const auto& h = c->ref.lv[level - 1].hier2lat;
    std::copy(h.begin(), h.end(), out);
    HMG_API_END
}

}  // extern "C"
