// Device-side parameter blocks and launcher declarations of libhmg_b200 (sm_100a only).
//
// Device layout of every state vector ("element-interleaved"): W consecutive coarse elements form
// a unit; entry (element e, packed lattice node p) of a level with nf nodes lives at
//     ((e / W) * nf + p) * W + e % W .
// The last unit is padded with zero columns.  A warp lane therefore IS a coarse element: node
// class, neighbour offsets and transfer tables are warp-uniform and every access is one coalesced
// 64/128-byte line.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hmg {

// ---- per-level device view -------------------------------------------------------------
struct LevelView {
    int m, nf;
    int W, wshift;              // elements per unit, log2
    int n_boundary, npf, npe;
    const uint32_t* nodeinfo;   // [nf]   i | j<<8 | k<<16 | class<<24   (lattice order)
    const uint32_t* boundary;   // [n_boundary] p | class<<14   (sorted by class)
    const double* G;            // [ncls][ndir][nc]
    const uint16_t* iface_idx;  // paired-node packed indices: faces [4][npf], edges [6|3][npe], vertices
    const uint32_t* interp_tab; // [nf] coarse parents pa | pb<<16 (levels >= 2)
    const uint16_t* restrict_tab; // [nf(level-1)][ndir] fine indices, 0xFFFF = outside
    int vpos[4];                // packed lattice index of the reference vertices
};

struct TopoView {
    int64_t ne;
    int64_t nedges, nverts;                 // multi-owner cells handled cell by cell (3D: edges + vertices, 2D: vertices)
    const int64_t *edge_off, *vert_off;
    const int32_t *edge_own, *vert_own;     // element*8 + local id, ascending element
    const int32_t* partner;                 // [ne][4]: other owner (element*8 + local id) of local face (3D) / edge (2D), -1 none
};

// scalar slots on the device (no host round trip inside a V-cycle)
enum Scalar { S_RHO = 0, S_PAP = 1, S_ALPHA = 2, S_RSQR = 3, S_BETA = 4, S_TMP = 5, S_NRM = 6, S_COUNT = 16 };
enum PostOp { POST_STORE = 0, POST_RHO = 1, POST_PAP = 2, POST_RSQR = 3, POST_ADD = 4, POST_RHO_ADD = 5 };

// ---- multi-GPU: peer memory over NVLink (one process per GPU, buffers mapped with CUDA IPC) -------------------------
// A rank's communication buffer holds, in this order: scalar mailboxes [PEER_SLOTS][nranks] (value + sequence number),
// exchange flags [nranks], and two receive areas (even / odd exchange) for the messages of the cut-cell exchange.
// Every rank maps the buffers of all others; kernels store straight into a peer's buffer and spin on their own.
struct PeerMail { double value; unsigned long long seq; };
constexpr int PEER_SLOTS = 4;
struct PeerView {
    int rank = 0, nranks = 1;                 // nranks <= 1: off (single GPU, or the NCCL path)
    PeerMail* const* mail = nullptr;          // [nranks] device pointers (peer-mapped; [rank] = the own buffer)
    unsigned long long* const* flag = nullptr;   // [nranks] exchange flags of every rank (peer-mapped)
    double* const* recv = nullptr;            // [nranks] receive areas of every rank (peer-mapped)
    int64_t recv_stride = 0;                  // doubles per receive area (the same on every rank)
    unsigned long long* rseq = nullptr;       // reductions completed on this rank (device counter)
    unsigned long long* xseq = nullptr;       // cut-cell exchanges completed on this rank
    unsigned int* xticket = nullptr;          // last-block-done counter of the exchange kernels
};
// a reduction whose post-op carries POST_GLOBAL is summed over all ranks INSIDE the kernel: the last block of every rank
// stores its total into the mailbox of every peer, waits for the others' totals and adds them in rank order (the same
// bits everywhere) before the post-op runs -- no collective call, no extra launch
constexpr int POST_GLOBAL = 0x100;

struct Reducer {
    double* partials;     // [max_blocks]
    double* scalars;      // [S_COUNT]
    unsigned int* ticket; // last-block-done counter (self-resetting)
    int max_blocks;
    PeerView peer;
};

// cut cells of one kind (faces / edges / vertices) this rank takes part in (multi-GPU, hmg_host.hpp)
struct CutView {
    int64_t ncells;
    const int64_t* off;         // CSR over the local owners
    const int32_t* own;         // local element * 8 + local id
    const uint8_t* first_local; // the globally first owner is own[off[c]]
    // neighbour exchange: the other ranks sharing the cell (ascending) and the cell's ordinal in the message
    const int64_t* peer_off;
    const int32_t* peer_rank;
    const int32_t* peer_idx;
    const int32_t* my_pos;      // peers with a smaller rank: where the own partial sum enters the ordered total
};
// peer-memory form of the exchange: rbase[rank * 3 + kind] = first entry of the kind's section of MY message inside the
// receive area of `rank`; nbr[0 .. nnbr) = the ranks this rank exchanges with at all
struct CutPeer {
    const int64_t* rbase = nullptr;
    const int32_t* nbr = nullptr;
    int nnbr = 0;
};
enum CutOp { CUT_PACK = 0, CUT_UNPACK = 1 };

enum ApplyMode { APPLY_AX = 0, APPLY_RESIDUAL = 1, APPLY_MULADD = 2 };

// launch configuration of the streaming apply kernel for one level (chosen on the host, api.cu)
struct ApplyConfig {
    int nwarps;        // consumer warps per CTA (one more warp issues the TMA copies)
    int ring_rows;     // rows (W doubles each) of the shared-memory ring
    int spill_rows;    // rows mirrored behind the ring so that a line never wraps
    int chunk_shift;   // log2(rows per TMA chunk)
    int seg;           // 2D: log2(nodes per task) (a line is split into segments); 3D: unused
    int run;           // 3D: consecutive lines of a plane per task
    int stage_shift, nstage;   // fused p-update: log2(rows per staging slot), staging slots (0: not fused)
    int nconv, slot_shift;     // ... converter warps, log2(staging slots per converter warp)
    int ctas_per_sm;
    int oversub;       // CTAs per SM slot of the grid (0: chosen from the problem size)
    size_t smem_bytes;
};

struct ApplyArgs {
    LevelView L;
    ApplyConfig cfg;
    ApplyConfig cfg_fused;     // launch shape of the fused p-update variant (ring_rows <= 0: not available)
    ApplyConfig cfg_rhs;       // launch shape of the residual / mul! variants
    const double* r2 = nullptr;   // fused p-update: input is r2 + beta * x (x = old p), p' goes to pout
    double* pout = nullptr;
    int64_t nunits;
    const double* tab;         // host pointer: StencilTab<DIM> of the level (copied into the parameter block)
    const double* coef;        // [nunits][CS][W]  |J| P (upper triangle) and |J|, element-interleaved
    const uint16_t* cmask;     // [nunits * W]
    const uint8_t* mult;       // [nunits][16][W] owners of the cell of every node class (dot weights), or null
    const double* x;           // input
    double* y;                 // output
    const double* b;           // rhs for APPLY_RESIDUAL
    double alpha, lambda;
    int mode;
    // fused reduction  sum_entries owners(entry) * x * y  ( = dot(x, broadcast(y)) over all stored entries,
    // src/multigrid.jl:62 ) -> Reducer post-op; only with APPLY_AX
    int dot_post;              // -1: none
    Reducer red;
    bool store = true;         // false (APPLY_AX with dot_post >= 0 only): y is not written, only the reduction is wanted
};

// launchers (all asynchronous on `st`); return the number of kernels launched
int launch_apply(int dim, const ApplyArgs& a, cudaStream_t st);
ApplyConfig make_apply_config(int dim, int m, int nf, int W, bool fused = false, bool streaming_rhs = false);
// part (measurement only): 3 = all shared cells, 1 = two-owner cells only, 2 = cells with more owners only
int launch_interface_sum(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st, int part = 3);
int launch_interface_sum_sq(int dim, const LevelView& L, const TopoView& T, double* x, const Reducer& R, int post, cudaStream_t st);
int launch_zero_all_but_one(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st);
// zero every local copy of a cut node except the globally first owner's (all three kinds in one launch)
int launch_cut_zero_but_first(int dim, const LevelView& L, const CutView* C, double* x, cudaStream_t st);
// neighbour exchange of the cut cells: CUT_PACK writes the partial sum of every cut node into the message of every
// rank that shares it, CUT_UNPACK adds the partial sums of all sharing ranks in ascending rank order (every rank gets
// the same bits).  kbase[rank * 3 + kind] = first entry of the kind's section in the message to / from `rank`.
int launch_cut_p2p(int dim, int op, const LevelView& L, const CutView* C, const int64_t* kbase, double* x, double* msg, bool sq,
                   const Reducer& R, cudaStream_t st, const CutPeer* peer = nullptr, int sq_post = POST_ADD);
// derived CG scalars after a cross-rank all-reduce of the raw dot product in S_TMP
int launch_scalar_post(const Reducer& R, int post, int slot, cudaStream_t st);
int launch_masked_copy_to_base(const LevelView& L1, int64_t nn, const int32_t* node_first, const uint8_t* contrib, const double* v,
                               double* u, cudaStream_t st);
int launch_rhs_flux(int dim, const LevelView& L, int64_t nunits, const double* dphi, const double* flux, double* b, cudaStream_t st);
int launch_integrate(int dim, const Reducer& R, const LevelView& L, int64_t nunits, int64_t nsubset, const int32_t* gidx,
                     const double* coef, const double* dphi, const double* flux, const double* v, const double* v2,
                     const double* Mv, cudaStream_t st);
int launch_apply_constraint(int dim, const LevelView& L, int64_t nbelems, const int32_t* belems, const uint16_t* cmask,
                            double* x, cudaStream_t st);
int launch_restrict(int dim, const LevelView& Lf, const LevelView& Lc, int64_t nunits, const double* rf, double* bc, cudaStream_t st);
int launch_interp_add(int dim, const LevelView& Lf, const LevelView& Lc, int64_t nunits, double* xf, const double* xc, cudaStream_t st);
int launch_dot(const Reducer& R, const double* a, const double* b, int64_t n, int post, int slot, cudaStream_t st);
int launch_copy_dot(const Reducer& R, const double* r, double* p, int64_t n, int post, cudaStream_t st);
int launch_cg_update(const Reducer& R, double* x, double* p, double* r, const double* Ap, int64_t n, int post, bool first, cudaStream_t st);
int launch_p_update(const Reducer& R, double* p, const double* r, int64_t n, cudaStream_t st);
int launch_x_update(const Reducer& R, double* x, const double* p, int64_t n, cudaStream_t st);   // x += S_ALPHA * p
int launch_axpy(double alpha, const double* x, double* y, int64_t n, cudaStream_t st);
int launch_fill(double* x, double v, int64_t n, cudaStream_t st);
int launch_fill_columns(const LevelView& L, int64_t ne, double* x, double v, cudaStream_t st);
// host layout (hierarchical rows, Nf x ncols column-major) <-> device layout, columns e0 .. e0+ncols
int launch_permute_in(const LevelView& L, const int32_t* hier2lat, const double* staged, int64_t ld_staged, double* dst,
                      int64_t e0, int64_t ncols, cudaStream_t st);
int launch_permute_out(const LevelView& L, const int32_t* hier2lat, const double* src, double* staged, int64_t ld_staged,
                       int64_t e0, int64_t ncols, cudaStream_t st);
// the first `nrows` hierarchical rows only (export of a coarser level's nodes)
int launch_permute_rows_out(const LevelView& L, const int32_t* hier2lat, const double* src, double* staged, int64_t ld_staged,
                            int nrows, int64_t e0, int64_t ncols, cudaStream_t st);
// dst = the first `ne` columns of src (same level shape; the columns padding dst's last unit are zeroed)
int launch_copy_columns(const LevelView& L, int64_t ne, double* dst, const double* src, cudaStream_t st);
// level 1 <-> base vector
int launch_copy_to_base(const LevelView& L1, int64_t nn, const int32_t* node_first, const double* v, double* u, cudaStream_t st);
int launch_distribute(int dim, const LevelView& L1, int64_t ne, const int32_t* elems, const double* u, double* v, cudaStream_t st);
int launch_gather(const int64_t* idx, int64_t n, const double* src, double* dst, cudaStream_t st);
int launch_scatter(const int64_t* idx, int64_t n, const double* src, double* dst, cudaStream_t st);
int launch_symmetrize_lower(double* A, int64_t n, cudaStream_t st);
int launch_set_diagonal(double* A, int64_t n, double v, cudaStream_t st);
int launch_symv_full(const double* A, int64_t n, const double* x, double* y, cudaStream_t st);
int launch_symv_half(const double* A, int64_t n, const double* x, double* y, double* work, cudaStream_t st);

}  // namespace hmg
