// Hand-written sm_100a kernels of libhmg_b200 (fp64, HBM-bound; tensor cores are not used).
//
//   K1 apply_kernel        y = A x per coarse element: constant-coefficient lattice stencil on the
//                          refined reference simplex.  Lanes are coarse elements (element-interleaved
//                          layout), the node planes of a unit stream through a shared-memory ring
//                          filled by TMA bulk copies (cp.async.bulk + mbarrier, a producer warp and
//                          free-running consumer warps), each lane walks lattice lines with a
//                          register sliding window.  Replaces the dim^2+1 CSC scatter-SpMVs of
//                          src/apply_local_operators.jl:93-133.
//   K2 interface kernels   sum the owners' copies of every shared face/edge/vertex node and write the
//                          sum back (src/implicit_fine_grid.jl:209-328); gather form, no atomics.
//   K3 vector kernels      fused CG updates with device-resident scalars (src/multigrid.jl:50-69).
//   K4 transfer kernels    restriction / interpolation in lattice form (src/interpolation.jl:52-74),
//                          level-1 gather/scatter (src/implicit_fine_grid.jl:148-202).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <type_traits>

#include "kernels.cuh"
#include "lattice.hpp"
#include "apply_core.cuh"

// HMG_BOUNDS (build variant, `make checked`): the pool has no compute-sanitizer, so the checked library asserts every
// ring read, every output store of the apply kernel and every entry the interface / CG kernels touch against its bounds
// (tests run the parity suite through it with HMG_LIB=variants/libhmg_checked.so)
#ifdef HMG_BOUNDS
#include <cassert>
#define HMG_DEV_ASSERT(c) assert(c)
#else
#define HMG_DEV_ASSERT(c) ((void)0)
#endif

namespace hmg {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA 1-D bulk copy global -> shared (SASS: UBLKCP), completion signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// non-blocking probe of a phase: several of them issue back to back, their latencies overlap
__device__ __forceinline__ unsigned mbar_test(uint64_t* bar, uint32_t parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void scalar_post(double* S, int post, int slot, double tot) {
    if (post == POST_STORE) S[slot] = tot;
    else if (post == POST_RHO || post == POST_RHO_ADD) S[S_RHO] = tot;
    else if (post == POST_PAP) { S[S_PAP] = tot; S[S_ALPHA] = S[S_RHO] / tot; }
    else if (post == POST_RSQR) { S[S_RSQR] = tot; S[S_BETA] = tot / S[S_RHO]; S[S_RHO] = tot; }
    else if (post == POST_ADD) S[slot] += tot;
}
// what a reduction kernel contributes to the post-op: its own total, plus what earlier kernels of the chain left in S_TMP
__device__ __forceinline__ double scalar_pre(const double* S, int post, double tot) {
    return post == POST_RHO_ADD ? S[S_TMP] + tot : tot;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- peer memory (NVLink): system-scope stores / loads and the in-kernel all-reduce of one double --------------------
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// a peer that never answers (a rank that died) must not hang the GPU: ~60 s of spinning, then the kernel traps and the
// host sees a launch failure.  (Ranks legitimately wait for each other for seconds where one of them does more host-side
// work; the coarse factorisation on rank 0, the longest such wait, ends in a barrier of its own, api.cu.)
constexpr long long PEER_TIMEOUT_CYCLES = 120000000000ll;
__device__ __forceinline__ void peer_wait_ge(const unsigned long long* p, unsigned long long want) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < want) {
        if (clock64() - t0 > PEER_TIMEOUT_CYCLES) __trap();     // (no printf: a call in here costs every reduction kernel registers)
        __nanosleep(20);
    }
}
// Sum of `local` over all ranks, added in rank order (every rank gets the same bits).  Called by ALL threads of one
// block per rank (the last block of a reduction kernel), which needs at least nranks threads.  Thread q stores the
// rank's value into the mailbox of rank q (value, then the sequence number with release semantics) and waits for
// rank q's value in its own mailbox; mailboxes are a ring of PEER_SLOTS sequence numbers (a rank is never more than
// one reduction ahead of a peer: it needs the peer's value of reduction s to finish s).
__device__ __forceinline__ double peer_allreduce(const PeerView& P, double local) {
    __shared__ double vals[32];
    const unsigned long long s = *P.rseq + 1;
    const int slot = (int)(s & (PEER_SLOTS - 1));
    const int t = threadIdx.x;
    if (t < P.nranks) {
        PeerMail* dst = P.mail[t] + slot * P.nranks + P.rank;
        st_relaxed_sys(&dst->value, local);
        st_release_sys(&dst->seq, s);
        const PeerMail* src = P.mail[P.rank] + slot * P.nranks + t;
        peer_wait_ge(&src->seq, s);
        vals[t] = ld_relaxed_sys(&src->value);
    }
    __syncthreads();
    double g = 0.0;
    for (int q = 0; q < P.nranks; ++q) g += vals[q];
    __syncthreads();
    if (t == 0) *P.rseq = s;
    return g;
}

// deterministic: fixed grid, fixed per-block tree, partials summed in block order by the last block
__device__ __forceinline__ void block_reduce_finish(double v, const Reducer& R, int post, int slot) {
    __shared__ double wsum[32];
    __shared__ bool last;
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wsum[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += wsum[w];
        R.partials[blockIdx.x] = s;
        __threadfence();
        const unsigned t = atomicAdd(R.ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(R.partials + b);
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) wsum[wid] = s;
    __syncthreads();
    const int op = post & 0xff;
    if (post & POST_GLOBAL) {
        // the sum over all ranks, inside this kernel (peer memory)
        __shared__ double mine;
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += wsum[w];
            mine = scalar_pre(R.scalars, op, tot);
        }
        __syncthreads();
        const double g = peer_allreduce(R.peer, mine);
        if (threadIdx.x == 0) {
            scalar_post(R.scalars, op, slot, g);
            *R.ticket = 0u;
        }
        return;
    }
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += wsum[w];
        scalar_post(R.scalars, op, slot, scalar_pre(R.scalars, op, tot));
        *R.ticket = 0u;
    }
}


// ------------------------------------------------------------------------------------------
// K1: local operator apply
// ------------------------------------------------------------------------------------------
// Lanes are coarse elements (W = 32 interleaved columns), so node class, neighbour offsets and control flow are
// warp-uniform and every shared/global access of a warp is one conflict-free row of W doubles.
// A CTA owns a contiguous range of (element group, lattice plane) pairs, balanced by rows.  Its input
// rows stream ONCE from HBM through a shared-memory ring: one elected thread issues TMA bulk copies
// (cp.async.bulk, completion on mbarriers) of fixed-size chunks, consumer warps take the lines of the
// planes round-robin, wait for the chunk that holds the last row their line reads, and hand chunks
// back once their window has passed them.  The ring is addressed by (stream row mod R); the first
// SP rows are mirrored behind the ring so that a line never wraps.  Each lane walks its line with a
// register sliding window: 7 (3D) / 3 (2D) shared loads per node instead of 15 / 7.
constexpr int APPLY_Q = 64, APPLY_QS = 6;   // mbarrier slots (chunks in flight), log2
constexpr int APPLY_STAGES = 16;             // staging slots of the fused p-update variant: 4 per converter warp
constexpr int APPLY_MAXT = 16 * 32;    // 15 consumer warps + the producer warp: 128 registers per thread
#ifndef HMG_TD2
#define HMG_TD2 4
#endif
#ifndef HMG_TD3_DOT
#define HMG_TD3_DOT 3      // values of b in flight per lane in the 3D residual with the fused reduction (2: 24 instead of 48 bytes of spills, but 0.6 % slower V-cycles)
#endif
#ifndef HMG_MAXT3
#define HMG_MAXT3 512
#endif

template <int DIM> struct ApplyParams {
    StencilTab<DIM> T;
    const double* x;
    double* y;
    const double* t;           // b (RESIDUAL) or y (MULADD)
    const double* coef;
    const uint16_t* cmask;
    const uint8_t* mult;
    int64_t nunits;
    double sa, lambda;
    int m, nf, nwarps, R, SP, cs, seg_shift, run;
    int dot_post;
    Reducer red;
    // FUSEP: x = old p, r2 = r, pout = the other p buffer, staging of 2^sr_shift rows x nstage slots
    const double* r2;
    double* pout;
    int sr_shift, nstage, nconv, slot_shift;
};

template <int DIM> __device__ __forceinline__ int plane_off(int m, int t) { return DIM == 3 ? lat_off3(m, t) : lat_off2(m, t); }

// first (group, plane) boundary at or after global row r; returns group * (m + 1) + plane
template <int DIM> __device__ __forceinline__ int64_t plane_at_or_after(int64_t r, int m, int nf) {
    const int64_t u = r / nf;
    const int rem = (int)(r - u * nf);
    if (rem == 0) return u * (m + 1);
    for (int t = 1; t <= m; ++t)
        if (plane_off<DIM>(m, t) >= rem) return u * (m + 1) + t;
    return (u + 1) * (m + 1);
}

// STORE = false: only the fused reduction of the product is wanted (the last CG step of a smoothing call whose residual
// nobody reads: alpha = rho / p.Ap needs the dot, x += alpha p needs p -- the vector Ap itself is dead)
template <int DIM, int W, int MODE, bool DOT, bool STORE = true> struct OutDev {
    static constexpr int APPLY_W = W;
    static constexpr bool UNROLL4 = DIM == 2 || !DOT;      // interior loop of the line sweep: 4 nodes per trip
    double *ylo, *yhi;     // bounds of the output vector (read by the checked build only)
    double* yl;            // output row of node 0 of the current line (lane included)
    const double* tl;
    double sa;
    unsigned cm;
    double dsum;
    unsigned long long ml, mh;   // owners of every class, one byte each (converted where a boundary node needs it:
                                 // keeping the four face weights as doubles cost the dot variants 8 registers and spills)
    template <int CLS> __device__ __forceinline__ double weight() const {
        if (CLS == 0) return 1.0;
        return (double)(unsigned)((CLS < 8 ? (ml >> (8 * (CLS & 7))) : (mh >> (8 * (CLS & 7)))) & 255ull);
    }
    // b / y of the next TD nodes travel in registers (they come from L2, where the producer's bulk prefetch
    // put them): the load of node k + TD is issued before node k is finished, which keeps
    // warps x TD x 256 bytes in flight per SM
    static constexpr int TD = DIM == 2 ? HMG_TD2 : (DOT ? HMG_TD3_DOT : 3);
    double tq[TD];
    int klast;
    __device__ __forceinline__ void begin(int k0, int k1) {
        if (MODE == APPLY_AX) return;
        klast = k1 - 1;
#pragma unroll
        for (int q = 0; q < TD; ++q) tq[q] = __ldcs(tl + min(k0 + q, klast) * APPLY_W);
    }
    template <int CLS> __device__ __forceinline__ void put(int k, double acc, double x0) {
        // bit 0 of the class mask is never set (topology.cpp only sets face / edge / vertex classes): interior nodes need no
        // Dirichlet select
        const bool fixed = CLS != 0 && MODE != APPLY_MULADD && ((cm >> CLS) & 1u);
        double v;
        if (MODE == APPLY_AX) v = fixed ? 0.0 : acc;
        else {
            const double tv = tq[0];
#pragma unroll
            for (int q = 0; q + 1 < TD; ++q) tq[q] = tq[q + 1];
            tq[TD - 1] = __ldcs(tl + min(k + TD, klast) * APPLY_W);
            v = MODE == APPLY_RESIDUAL ? (fixed ? 0.0 : tv - acc) : fma(sa, acc, tv);
        }
        HMG_DEV_ASSERT(!STORE || (yl + k * APPLY_W >= ylo && yl + k * APPLY_W < yhi));
        if (STORE) yl[k * APPLY_W] = v;
        if (DOT) {
            if (MODE == APPLY_AX) dsum = fma(weight<CLS>() * x0, v, dsum);
            else if (CLS == 0) dsum = fma(v, v, dsum);      // interior part of dot(r, r); interfaces: K2
        }
    }
};

struct SmemLoad {
    const double* sm;
    int limit;             // doubles in the ring + mirror rows
    __device__ __forceinline__ double operator()(int addr) const {
        HMG_DEV_ASSERT(addr >= 0 && addr < limit);
        return sm[addr];
    }
    __device__ __forceinline__ const double* ptr(int addr) const { return sm + addr; }
    __device__ __forceinline__ double ld(const double* p) const {
        HMG_DEV_ASSERT(p >= sm && p < sm + limit);
        return *p;
    }
};

// FUSEP (with MODE = AX): the input is not a stored vector but the new search direction p' = r + beta p of
// src/multigrid.jl:68.  The producer warp bulk-copies chunks of r and p into a small staging area, converts them
// into the ring (p' is what the stencil reads) and stores p' to the OTHER p buffer (neighbouring CTAs still read
// the old p for their halo planes), so p' never makes a round trip through HBM before it is applied.
template <int DIM, int W, int MODE, bool DOT, bool FUSEP, bool STORE = true>
__global__ void __launch_bounds__(DIM == 3 ? HMG_MAXT3 : APPLY_MAXT, 1) apply_kernel(const __grid_constant__ ApplyParams<DIM> a) {
    using D = Dims<DIM>;
    constexpr int APPLY_W = W;
    static_assert(W == 32, "a unit is one warp wide");
    static_assert(STORE || (DOT && MODE == APPLY_AX), "without the store only the fused p.Ap is left");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[APPLY_Q], empty_bar[APPLY_Q], stage_bar[APPLY_STAGES];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const int m = a.m, nf = a.nf, NW = a.nwarps, R = a.R, SP = a.SP, CS = a.cs, CH = 1 << a.cs;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int el = lane;                                 // element of the unit
    const int NPL = m + 1;

    if (threadIdx.x == 0) {
        // FUSEP: a ring chunk is complete after all of its staging sub-chunks were converted
        for (int s = 0; s < APPLY_Q; ++s) { mbar_init(&full_bar[s], FUSEP ? (1u << (a.cs - a.sr_shift)) : 1u); mbar_init(&empty_bar[s], NW); }
        for (int s = 0; s < APPLY_STAGES; ++s) mbar_init(&stage_bar[s], 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();

    // this CTA's range of (group, plane) pairs and the rows it streams (one halo plane on each side)
    const int64_t total_rows = a.nunits * nf;
    const int64_t P0 = plane_at_or_after<DIM>(total_rows * blockIdx.x / gridDim.x, m, nf);
    const int64_t P1 = blockIdx.x + 1 == gridDim.x ? a.nunits * NPL
                                                   : plane_at_or_after<DIM>(total_rows * (blockIdx.x + 1) / gridDim.x, m, nf);
    const int64_t u0 = P0 / NPL, u1 = P1 / NPL;
    const int t0 = (int)(P0 - u0 * NPL), t1 = (int)(P1 - u1 * NPL);
    const int64_t g0 = u0 * nf + plane_off<DIM>(m, t0 > 0 ? t0 - 1 : 0);
    const int64_t gend = t1 == 0 ? u1 * nf : u1 * nf + plane_off<DIM>(m, t1 + 1);
    const int stotal = P1 > P0 ? (int)(gend - g0) : 0;
    const int nchunks = (stotal + CH - 1) >> CS;
    double dsum = 0.0;

    if (FUSEP && warp >= NW) {
        // ---------------- producers + converters: warp NW + c takes the staging chunks j = c (mod NCV) --------
        // Each converter warp owns 4 staging slots, issues its own bulk copies (r and p of a chunk of 16 rows
        // complete on one mbarrier), converts the chunk into the ring and arrives on the ring chunk's barrier.
        const int cw = warp - NW, NCV = a.nconv;
        constexpr int SR = 16, SRS = 4;
        const int SLS = a.slot_shift, SLOTS = 1 << SLS;      // staging slots per converter warp
        const int nsub = (stotal + SR - 1) >> SRS;
        if (nchunks > 0 && cw < NCV) {
            const uint32_t RB = APPLY_W * 8;
            double* stage = sm + (size_t)(R + SP) * APPLY_W + (size_t)cw * SLOTS * 2 * SR * APPLY_W;
            uint64_t* sbar = stage_bar + cw * SLOTS;
            const double* rsrc = a.r2 + g0 * APPLY_W;
            const double* psrc = a.x + g0 * APPLY_W;
            double* pdst = a.pout + g0 * APPLY_W + lane;
            const double beta = a.red.scalars[S_BETA];
            // rows of the stream this CTA owns (the halo planes belong to its neighbours, who store them)
            const int own0 = (int)(u0 * nf + plane_off<DIM>(m, t0) - g0);
            const int own1 = (int)(t1 == 0 ? u1 * nf - g0 : u1 * nf + plane_off<DIM>(m, t1) - g0);
            auto issue = [&](int j, int slot) {
                const int s0 = j << SRS, n = min(SR, stotal - s0);
                double* dst = stage + (size_t)slot * 2 * SR * APPLY_W;
                mbar_expect_tx(&sbar[slot], 2u * (uint32_t)n * RB);
                bulk_g2s(dst, rsrc + (int64_t)s0 * APPLY_W, (uint32_t)n * RB, &sbar[slot]);
                bulk_g2s(dst + SR * APPLY_W, psrc + (int64_t)s0 * APPLY_W, (uint32_t)n * RB, &sbar[slot]);
            };
            if (lane == 0)
                for (int q = 0; q < SLOTS; ++q)
                    if (cw + q * NCV < nsub) issue(cw + q * NCV, q);
            int confirmed = 0, it = 0;
            int phys = (cw << SRS) % R;                       // ring row of the chunk's first row
            const int pstep = (NCV << SRS) % R;
            for (int j = cw; j < nsub; j += NCV, ++it) {
                const int s0 = j << SRS, n = min(SR, stotal - s0), slot = it & (SLOTS - 1);
                mbar_wait(&sbar[slot], (unsigned)(it >> SLS) & 1u);
                const int old = s0 + n - 1 - R;
                if (old >= 0) {
                    const int c_old = old >> CS;
                    while (confirmed <= c_old) {
                        mbar_wait(&empty_bar[confirmed & (APPLY_Q - 1)], (confirmed >> APPLY_QS) & 1u);
                        ++confirmed;
                    }
                }
                const double* sr = stage + (size_t)slot * 2 * SR * APPLY_W + lane;
                const double* sp = sr + SR * APPLY_W;
                if (n == SR && phys >= SP && phys + SR <= R && s0 >= own0 && s0 + SR <= own1) {
                    // common case: a whole slot, no ring wrap, no mirror rows, every row owned -- fully unrolled
                    double* dr = sm + phys * APPLY_W + lane;
                    double* dg = pdst + (int64_t)s0 * APPLY_W;
                    double v[SR];
#pragma unroll
                    for (int q = 0; q < SR; ++q) v[q] = fma(beta, sp[q * APPLY_W], sr[q * APPLY_W]);
#pragma unroll
                    for (int q = 0; q < SR; ++q) { dr[q * APPLY_W] = v[q]; dg[q * APPLY_W] = v[q]; }
                } else {
                    for (int q = 0; q < n; ++q) {
                        const double v = fma(beta, sp[q * APPLY_W], sr[q * APPLY_W]);
                        int ph = phys + q;
                        if (ph >= R) ph -= R;
                        sm[ph * APPLY_W + lane] = v;
                        if (ph < SP) sm[(R + ph) * APPLY_W + lane] = v;
                        const int srow = s0 + q;
                        if (srow >= own0 && srow < own1) pdst[(int64_t)srow * APPLY_W] = v;
                    }
                }
                phys += pstep;
                if (phys >= R) phys -= R;
                __syncwarp();
                if (lane == 0) {
                    fence_proxy_async();                       // the slot's generic reads are done before TMA rewrites it
                    if (j + SLOTS * NCV < nsub) issue(j + SLOTS * NCV, slot);
                    uint64_t* bar = &full_bar[(s0 >> CS) & (APPLY_Q - 1)];
                    mbar_arrive(bar);
                    if (j == nsub - 1) {                       // the last ring chunk may hold fewer sub-chunks
                        const int have = ((stotal - ((s0 >> CS) << CS)) + SR - 1) >> SRS;
                        for (int q = have; q < (1 << (CS - SRS)); ++q) mbar_arrive(bar);
                    }
                }
            }
        }
    } else if (warp == NW) {
        // ---------------- producer ----------------
        if (lane == 0 && nchunks > 0) {
            const uint32_t RB = APPLY_W * 8;
            int phys = 0, confirmed = 0;
            const double* src = a.x + g0 * APPLY_W;
            const char* pre = MODE == APPLY_AX ? nullptr : reinterpret_cast<const char*>(a.t + g0 * APPLY_W);
            for (int q = 0; q < nchunks; ++q) {
                const int s0 = q << CS;
                const int n = min(CH, stotal - s0);
                const int old = s0 + n - 1 - R;
                if (old >= 0) {
                    const int c_old = old >> CS;
                    while (confirmed <= c_old) {
                        mbar_wait(&empty_bar[confirmed & (APPLY_Q - 1)], (confirmed >> APPLY_QS) & 1u);
                        ++confirmed;
                    }
                }
                uint64_t* bar = &full_bar[q & (APPLY_Q - 1)];
                const int n1 = min(n, R + SP - phys);
                const int wrapped = phys + n - R;
                const int dup = phys < SP ? min(n, SP - phys) : 0;
                mbar_expect_tx(bar, (uint32_t)(n1 + (wrapped > 0 ? wrapped : 0) + dup) * RB);
                const double* s = src + (int64_t)s0 * APPLY_W;
                bulk_g2s(sm + (size_t)phys * APPLY_W, s, (uint32_t)n1 * RB, bar);
                if (wrapped > 0) bulk_g2s(sm, s + (size_t)(R - phys) * APPLY_W, (uint32_t)wrapped * RB, bar);
                if (dup > 0) bulk_g2s(sm + (size_t)(R + phys) * APPLY_W, s, (uint32_t)dup * RB, bar);
                if (MODE != APPLY_AX)      // the rows of b / y the consumers will read directly: pull them into L2
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pre + (size_t)s0 * RB), "r"((uint32_t)n * RB) : "memory");
                phys += n;
                if (phys >= R) phys -= R;
            }
        }
    } else if (nchunks > 0) {
        // ---------------- consumers ----------------
        SmemLoad mem{sm, (R + SP) * APPLY_W};
        OutDev<DIM, W, MODE, DOT, STORE> out;
        out.ylo = a.y; out.yhi = a.y + a.nunits * (int64_t)nf * APPLY_W;
        out.sa = a.sa;
        out.dsum = 0.0;
        out.cm = 0; out.ml = out.mh = 0ull;
        // variants short of registers re-read |J| P for the rare special coefficients
        constexpr bool LEAN = DOT || (DIM == 3 && MODE != APPLY_AX);
        typename std::conditional<LEAN, LaneOpMem<DIM>, LaneOp<DIM>>::type op;
        const int RL = DIM == 3 ? a.run : 1;     // 3D: consecutive lines of a plane per task
        const int SEGS = a.seg_shift;           // 2D: log2(nodes per task)
        int64_t u = u0, ucur = -1;
        int t = t0, i = 0;                 // 3D: run i of plane t; 2D: segment i of line t
        int su = 0;                        // stream row of node 0 of the current group
        int waited = 0, released = 0;
        int sb = 0, pb = 0;                // stream row / ring row of the oldest row of the current window
        double* ybase = nullptr;
        const double* tbase = nullptr;
        int step = warp;                   // tasks to skip before the next one of this warp
        for (;;) {
            // advance the cursor by `step` tasks
            i += step;
            for (;;) {
                // 3D: runs of RL lines, or (SEGS < 30, RL = 1) every line longer than 2^SEGS nodes cut into two halves along
                // k, so that the warps of the CTA spread over fewer lines of the plane (a smaller window in the ring)
                const int cnt = DIM == 3 ? (SEGS < 30 ? (t + 1) << (m - t + 1 > (1 << SEGS) ? 1 : 0) : (t + RL) / RL)
                                         : ((m - t + (1 << SEGS)) >> SEGS);
                if (i < cnt) break;
                i -= cnt;
                if (++t > m) { t = 0; ++u; }
                if (u * NPL + t >= P1) break;
            }
            if (u * NPL + t >= P1) break;
            step = NW;
            if (u != ucur) {
                ucur = u;
                const int64_t e = u * APPLY_W + el;
                if constexpr (LEAN) {
                    op.ecp = a.coef + u * D::CS * APPLY_W + el;
                    op.stride = APPLY_W;
                    op.lambda = a.lambda;
                } else {
#pragma unroll
                    for (int q = 0; q < D::NC; ++q) op.ec[q] = __ldg(a.coef + (u * D::CS + q) * APPLY_W + el);
                    op.ec[D::NC - 1] *= a.lambda;
                }
                interior_coefficients(op, a.T);
                if (MODE != APPLY_MULADD) out.cm = (unsigned)__ldg(a.cmask + e);
                if (DOT && MODE == APPLY_AX) {
                    const uint8_t* mp = a.mult + u * 16 * APPLY_W + el;
                    unsigned long long lo = 0, hi = 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        lo |= (unsigned long long)__ldg(mp + c * APPLY_W) << (8 * c);
                        hi |= (unsigned long long)__ldg(mp + (c + 8) * APPLY_W) << (8 * c);
                    }
                    out.ml = lo; out.mh = hi;
                }
                su = (int)(u * nf - g0);
                ybase = a.y + u * (int64_t)nf * APPLY_W + el;
                tbase = MODE == APPLY_AX ? nullptr : a.t + u * (int64_t)nf * APPLY_W + el;
            }
            // rows of the task: first line, number of lines, row window [behind, need]
            LineGeo<DIM> g;
            int rc, rm[Sweep<DIM>::NP], rp[Sweep<DIM>::NP], need, behind, nl = 1, il = 0;
            if constexpr (DIM == 3) {
                int k0 = 0, k1 = m - t + 1;
                if (SEGS < 30) {
                    // (shifts only: the two integer divisions of a general segment count cost ~50 instructions per task)
                    const int Lt = m - t + 1, two = Lt > (1 << SEGS) ? 1 : 0;
                    il = i >> two;
                    const int len = (Lt + two) >> two;                  // balanced halves
                    k0 = (i & two) * len;
                    k1 = min(Lt, k0 + len);
                    nl = 1;
                } else {
                    il = i * RL;
                    nl = min(RL, t + 1 - il);
                }
                const LineRows3 r = line_rows3(m, t, il, lat_off3(m, t));
                g.L = r.L; g.k0 = k0; g.k1 = k1;
                rc = r.c; behind = r.behind;
                need = t < m ? r.need + (nl - 1) * (r.L - 1) : min(r.c + nl, plane_off<3>(m, m + 1) - 1);
#pragma unroll
                for (int q = 0; q < Sweep<DIM>::NP; ++q) { rm[q] = r.rm[q]; rp[q] = r.rp[q]; }
            } else {
                const int k0 = i << SEGS;
                const LineRows2 r = line_rows2(m, t, k0, min(m - t + 1, k0 + (1 << SEGS)));
                g.L = r.L; g.k0 = k0; g.k1 = min(r.L, k0 + (1 << SEGS));
                rc = r.c; need = r.need; behind = r.behind;
                rm[0] = r.rm[0]; rp[0] = r.rp[0];
            }
            // Observe every chunk up to the one holding the last row this task reads and hand back the chunks
            // below its window.  A slot is handed back only after its chunk was seen (an arrival on an empty
            // barrier must not overtake the phase of the previous user of the slot), and hand-backs are not
            // postponed behind a wait (the producer may need them to load what this warp waits for).
            {
                const int qn = (su + need) >> CS, qb = (su + behind) >> CS;
                const int r0 = min(qb, waited);
                if (released < r0) {
                    __syncwarp();
                    if (lane == 0)
                        for (int c = released; c < r0; ++c) mbar_arrive(&empty_bar[c & (APPLY_Q - 1)]);
                    released = r0;
                }
                // A warp sees every chunk (a line of its own every NW lines: ~8 chunks per task at 3D level 6).  Probing
                // them one blocking wait after the other costs the latency of a barrier probe per chunk even when the data
                // has long arrived (ncu, round 2: 38 % of the kernel's stall samples sat on these branches); four probes
                // issued back to back cost one.  Only when a probe fails does the warp block, on the oldest chunk.
#ifdef HMG_SERIAL_PROBES          // (A/B build of round 2: one blocking wait per chunk, as in round 1)
                while (waited <= qn) {
                    mbar_wait(&full_bar[waited & (APPLY_Q - 1)], (waited >> APPLY_QS) & 1u);
                    ++waited;
                    if (released < qb) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[released & (APPLY_Q - 1)]);
                        ++released;
                    }
                }
#else
                while (waited <= qn) {
                    const int n = min(qn - waited + 1, 4);
                    unsigned ok = 1u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int cq = waited + min(j, n - 1);
                        ok &= mbar_test(&full_bar[cq & (APPLY_Q - 1)], (cq >> APPLY_QS) & 1u);
                    }
                    if (ok) waited += n;
                    else {
                        mbar_wait(&full_bar[waited & (APPLY_Q - 1)], (waited >> APPLY_QS) & 1u);
                        ++waited;
                    }
                    const int r1 = min(qb, waited);
                    if (released < r1) {
                        __syncwarp();
                        if (lane == 0)
                            for (int c = released; c < r1; ++c) mbar_arrive(&empty_bar[c & (APPLY_Q - 1)]);
                        released = r1;
                    }
                }
#endif
            }
            // ring rows: the window start moves monotonically, every line starts less than R rows after it
            pb += su + behind - sb;
            sb = su + behind;
            while (pb >= R) pb -= R;
            auto ring = [&](int row) {
                int p = pb + (row - behind);
                if (p >= R) p -= R;
                return p;
            };
            int qc = 0, qm[Sweep<DIM>::NP], qp[Sweep<DIM>::NP];
            if constexpr (DIM == 3) {
                qc = ring(rc);
#pragma unroll
                for (int q = 0; q < Sweep<DIM>::NP; ++q) { qm[q] = ring(rm[q]); qp[q] = ring(rp[q]); }
            }
            out.yl = ybase + (int64_t)rc * APPLY_W;
            out.tl = MODE == APPLY_AX ? nullptr : tbase + (int64_t)rc * APPLY_W;
            if constexpr (DIM == 3) {
                const int L = g.L;
                auto adv = [&](int& p, int d) { p += d; if (p >= R) p -= R; };
                for (int li = 0; li < nl; ++li) {
                    g.bc = qc * APPLY_W + el;
                    g.bm[0] = qm[0] * APPLY_W + el; g.bm[1] = qm[1] * APPLY_W + el; g.bm[2] = qm[2] * APPLY_W + el;
                    g.bp[0] = qp[0] * APPLY_W + el; g.bp[1] = qp[1] * APPLY_W + el; g.bp[2] = qp[2] * APPLY_W + el;
                    if (g.k0 < g.k1) {                  // (a half can only be empty for lines of one node, which are never cut)
                        out.begin(g.k0, g.k1);
                        run_line3(op, a.T, mem, APPLY_W, g, t, il + li, out);
                    }
                    // the next line of the plane: every base moves by one line of its own plane
                    adv(qc, L); adv(qm[0], L - 1); adv(qm[1], L - 1); adv(qm[2], L);
                    adv(qp[0], L + 1); adv(qp[1], L + 1); adv(qp[2], L);
                    out.yl += L * APPLY_W;
                    if (MODE != APPLY_AX) out.tl += L * APPLY_W;
                }
            } else {
                // ring rows are taken at node kb = k0 - 1 of each line (the first node the segment reads), so that
                // a segment never runs more than SEG + 2 rows past its base: the mirror behind the ring stays small
                const int kb = g.k0 > 0 ? g.k0 - 1 : 0;
                g.bc = (ring(rc + kb) - kb) * APPLY_W + el;
                g.bm[0] = (ring(rm[0] + kb) - kb) * APPLY_W + el;
                g.bp[0] = (ring(rp[0] + kb) - kb) * APPLY_W + el;
                out.begin(g.k0, g.k1);
                run_line2(op, a.T, mem, APPLY_W, g, t, out);
            }
        }
        // hand every remaining chunk back (other warps may still need the slots they occupy)
        while (released < nchunks) {
            if (waited <= released) {
                mbar_wait(&full_bar[waited & (APPLY_Q - 1)], (waited >> APPLY_QS) & 1u);
                ++waited;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[released & (APPLY_Q - 1)]);
            ++released;
        }
        dsum = out.dsum;
    }
    if (DOT) block_reduce_finish(dsum, a.red, a.dot_post, S_TMP);
}

// ring size and launch shape of one level
static ApplyConfig make_apply_config_slots(int dim, int m, int nf, int W, bool fused, int slot_shift, bool streaming_rhs = false);
ApplyConfig make_apply_config(int dim, int m, int nf, int W, bool fused, bool streaming_rhs) {
    if (!fused) return make_apply_config_slots(dim, m, nf, W, false, 2, streaming_rhs);
    // fused p-update: prefer 4 staging slots per converter warp, fall back to 2 where the ring is tight (3D level 6)
    const char* v = getenv("HMG_APPLY_SLOT_SHIFT");
    for (int ss : {2, 1}) {
        if (v && atoi(v) != ss) continue;
        const ApplyConfig c = make_apply_config_slots(dim, m, nf, W, true, ss);
        if (c.ring_rows > 0) return c;
    }
    ApplyConfig none{};
    none.ring_rows = -1;
    return none;
}
static ApplyConfig make_apply_config_slots(int dim, int m, int nf, int W, bool fused, int slot_shift, bool streaming_rhs) {
    ApplyConfig c{};
    auto envi = [](const char* name, int dflt) {
        const char* v = getenv(name);
        return v ? atoi(v) : dflt;
    };
    const int maxw = (dim == 3 ? HMG_MAXT3 : APPLY_MAXT) / 32;      // warps of a CTA (consumers + producer / converters)
    c.nwarps = std::max(1, std::min(maxw - 1, envi("HMG_APPLY_WARPS", maxw - 1)));
    c.ctas_per_sm = 1;
    c.oversub = envi("HMG_APPLY_OVERSUB", 0);
    // 2D: log2(nodes per task); 3D: 30 = whole lines, else every line longer than 2^seg nodes is cut into two halves (one line per task)
    c.seg = dim == 2 ? std::max(2, std::min(8, envi("HMG_APPLY_SEG_SHIFT", 5))) : envi("HMG_APPLY_SEG3_SHIFT", 30);
    if (dim == 3 && (c.seg < 2 || c.seg > 8)) c.seg = 30;
    c.spill_rows = dim == 2 ? (1 << c.seg) + 3 : m + 3;       // rows a task may run past the base of a line
    const int rowb = W * 8;
    // fused p-update: staging slots of 16 rows of r and p each; enough of them to keep ~48 KB in flight
    c.stage_shift = 4;
    // fused p-update: converter warps with 2^slot_shift staging slots of 16 rows of r and p each
    c.nconv = fused ? std::max(1, std::min(4, envi("HMG_APPLY_CONVERTERS", 2))) : 0;
    c.slot_shift = slot_shift;
    c.nstage = c.nconv << c.slot_shift;
    const int stage_bytes = c.nstage * 2 * (1 << c.stage_shift) * rowb;
    const int max_rows = std::min((227 * 1024 - 2048 - stage_bytes) / rowb - c.spill_rows, envi("HMG_APPLY_RING_ROWS", 1 << 20));
    // the largest row window of a task, for `run` lines per task (3D)
    auto window = [&](int run) {
        int w = 1;
        if (dim == 3) {
            for (int t = 0; t <= m; ++t)
                for (int i = 0; i <= t; i += run) {
                    const LineRows3 r = line_rows3(m, t, i, lat_off3(m, t));
                    const int nl = std::min(run, t + 1 - i);
                    const int need = t < m ? r.need + (nl - 1) * (r.L - 1) : r.c + nl;
                    w = std::max(w, need - r.behind + 1);
                }
        } else {
            for (int i = 0; i <= m; ++i) {
                const LineRows2 r = line_rows2(m, i, 0, m - i + 1);
                w = std::max(w, r.need - r.behind + 1);
            }
        }
        return w;
    };
    // prefer long runs (per-task overhead) and large chunks (every warp observes every chunk), as long as
    // the ring keeps `slack` rows of prefetch beyond the window: bytes in flight hide the HBM latency
    // The variants that also stream b / y straight from L2 (residual, mul!) prefer large chunks -- the producer's
    // L2 prefetch runs one chunk ahead -- and short runs (measured on B200).
    const int slack = streaming_rhs ? 48 : 160;
    const int run_env = envi("HMG_APPLY_RUN", 0), cs_env = envi("HMG_APPLY_CHUNK_SHIFT", 0);
    c.run = 1; c.chunk_shift = 5;
    bool found = false;
    auto fits = [&](int run, int cs) {
        if (dim == 2 && run != 1) return false;
        if ((run_env && run != run_env) || (cs_env && cs != cs_env)) return false;
        // (with a tight ring, runs of two lines cost the residual variant 30 % at 3D level 6)
        return window(run) + 2 * (1 << cs) + slack + (streaming_rhs && run > 1 ? 200 : 0) <= max_rows;
    };
    if (streaming_rhs) {
        for (int cs : {7, 6, 5}) {
            for (int run : {2, 1})
                if (fits(run, cs)) { c.run = run; c.chunk_shift = cs; found = true; break; }
            if (found) break;
        }
    } else {
        for (int run : {4, 2, 1}) {
            for (int cs : {7, 6, 5})
                if (fits(run, cs)) { c.run = run; c.chunk_shift = cs; found = true; break; }
            if (found) break;
        }
    }
    if (!found) { c.run = run_env ? run_env : 1; c.chunk_shift = cs_env ? cs_env : 5; }
    // 3D hierarchies of 6 grids: the two-plane window leaves the ring no room for runs of lines, and the 15 warps of a
    // CTA spread over 15 lines of a plane; cutting the lines in two halves that spread (measured on C4: apply 4.86 ->
    // 4.72 ms, residual 7.05 -> 6.85 ms, profiles/r02e_launch_shapes_C4.jsonl)
    if (dim == 3 && c.seg >= 30 && m >= 32 && !run_env && !getenv("HMG_APPLY_SEG3_SHIFT")) c.seg = 4;
    if (dim == 3 && c.seg < 30) c.run = 1;
    const int CH = 1 << c.chunk_shift;
    const int min_rows = window(c.run) + 2 * CH;
    int R = std::min(max_rows, (APPLY_Q - 2) * CH);
    // the variants that read b / y straight through L1 leave the rest of the SM's memory to the cache
    if (streaming_rhs && !getenv("HMG_APPLY_RING_ROWS")) R = std::min(R, min_rows + 200);
    c.ring_rows = R >= min_rows ? R : -1;        // -1: the level does not fit (refused by the launcher)
    c.smem_bytes = (size_t)(std::max(R, 1) + c.spill_rows) * rowb + stage_bytes;
    if (fused) c.nwarps = std::min(c.nwarps, maxw - c.nconv);
    (void)nf;
    return c;
}

template <int DIM, int W, int MODE, bool DOT, bool FUSEP = false, bool STORE = true>
static int launch_apply_t(const ApplyArgs& a, cudaStream_t st) {
    // per device: the opt-in shared-memory size is a per-device function attribute.  One host thread per process drives
    // the library (include/hmg.h), so these caches need no lock.
    static size_t configured_dev[64] = {0};
    static int sms_dev[64] = {0};
    auto kern = apply_kernel<DIM, W, MODE, DOT, FUSEP, STORE>;
    const ApplyConfig& cfg = FUSEP ? a.cfg_fused : (MODE == APPLY_AX ? a.cfg : a.cfg_rhs);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return -1;
    size_t& configured = configured_dev[dev];
    int& sms = sms_dev[dev];
    if (cfg.smem_bytes > configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes) != cudaSuccess) return -1;
        configured = cfg.smem_bytes;
    }
    if (sms == 0) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ApplyParams<DIM> p;
    memcpy(&p.T, a.tab, sizeof(p.T));
    p.x = a.x; p.y = a.y;
    p.t = MODE == APPLY_RESIDUAL ? a.b : (MODE == APPLY_MULADD ? a.y : nullptr);
    p.coef = a.coef; p.cmask = a.cmask; p.mult = a.mult;
    p.nunits = a.nunits;
    p.sa = MODE == APPLY_MULADD ? a.alpha : 1.0;
    p.lambda = a.lambda;
    p.m = a.L.m; p.nf = a.L.nf;
    p.nwarps = cfg.nwarps; p.R = cfg.ring_rows; p.SP = cfg.spill_rows; p.cs = cfg.chunk_shift;
    p.seg_shift = cfg.seg;
    p.run = cfg.run;
    p.r2 = a.r2; p.pout = a.pout; p.sr_shift = cfg.stage_shift; p.nstage = cfg.nstage; p.nconv = cfg.nconv; p.slot_shift = cfg.slot_shift;
    p.dot_post = a.dot_post;
    p.red = a.red;
    // more CTAs than SMs when the problem is large: the hardware hands the next CTA to whichever SM finishes
    // first (SMs do not all see the same memory latency); every CTA still streams >= 8192 rows
    const int64_t planes = a.nunits * (a.L.m + 1);
    const int64_t rows_per_sm = a.nunits * a.L.nf / sms;
    // (measured: + 6 % in 2D; in 3D the extra halo planes and pipeline fills cost more than they gain)
    const int64_t over = cfg.oversub > 0 ? cfg.oversub : (DIM == 2 ? std::max<int64_t>(1, std::min<int64_t>(8, rows_per_sm / 8192)) : 1);
    int64_t grid = std::min<int64_t>((int64_t)sms * cfg.ctas_per_sm * over, planes);
    if (DOT) grid = std::min<int64_t>(grid, a.red.max_blocks);
    kern<<<(unsigned)grid, (cfg.nwarps + (FUSEP ? cfg.nconv : 1)) * 32, cfg.smem_bytes, st>>>(p);
    return 1;
}

template <int DIM, int W> static int launch_apply_d(const ApplyArgs& a, cudaStream_t st) {
    if (!a.store && (a.mode != APPLY_AX || a.dot_post < 0)) return -1;       // nothing would be left of the product
    if (a.mode == APPLY_AX && a.r2 != nullptr) {
        if (a.dot_post < 0 || a.cfg_fused.ring_rows <= 0) return -1;
        return a.store ? launch_apply_t<DIM, 32, APPLY_AX, true, true>(a, st) : launch_apply_t<DIM, 32, APPLY_AX, true, true, false>(a, st);
    }
    if (a.mode == APPLY_AX) {
        if (a.dot_post < 0) return launch_apply_t<DIM, W, APPLY_AX, false>(a, st);
        return a.store ? launch_apply_t<DIM, W, APPLY_AX, true>(a, st) : launch_apply_t<DIM, W, APPLY_AX, true, false, false>(a, st);
    }
    if (a.mode == APPLY_RESIDUAL) return a.dot_post >= 0 ? launch_apply_t<DIM, W, APPLY_RESIDUAL, true>(a, st) : launch_apply_t<DIM, W, APPLY_RESIDUAL, false>(a, st);
    return launch_apply_t<DIM, W, APPLY_MULADD, false>(a, st);
}

int launch_apply(int dim, const ApplyArgs& a, cudaStream_t st) {
    if (a.nunits == 0) return 0;
    if (a.cfg.ring_rows <= 0 || a.L.W != 32) return -1;
    return dim == 3 ? launch_apply_d<3, 32>(a, st) : launch_apply_d<2, 32>(a, st);
}

// ------------------------------------------------------------------------------------------
// K2: interface sums
// ------------------------------------------------------------------------------------------
// Codimension-1 cells (3D faces, 2D edges) have exactly two owners: lanes are the elements of a
// unit, the lower owner of a pair reads both copies, adds them in ascending owner order
// (src/implicit_fine_grid.jl:219-244) and writes both.  Cells with more owners (3D edges, vertices):
// a group of 8 threads per cell, one owner per thread, walking the cell's shared nodes four at a time -- the owner
// copies of a node are loaded at once, every thread of the group adds them in ascending owner order (the same bits
// as the serial loop of src/implicit_fine_grid.jl:262-283) and writes the sum to its own owner's copy.  OP 0: sum +
// broadcast; OP 1: zero all but the first owner (src/implicit_fine_grid.jl:334-386).
// SQ: additionally reduce sum over the touched entries of value^2 ( = owners * sum^2 per shared node) and add it
// to S_TMP -- the part of rho = dot(r, r) that lives on interfaces; the apply kernel reduces the interior part
// (src/multigrid.jl:54 without a pass over r).  With SQ the grid is bounded and blocks loop over virtual blocks.
constexpr int MULTI_G = 8;                      // threads per multi-owner cell (one owner each)
constexpr int MULTI_ITEMS = 256 / MULTI_G;      // cells per virtual block
template <int DIM, int OP, bool SQ>
__global__ void __launch_bounds__(256) interface_kernel(const LevelView L, const TopoView T, int64_t npair_blocks,
                                                        int64_t vb_begin, int64_t nvirtual, double* __restrict__ x, const Reducer R,
                                                        int post) {
    constexpr int NF = DIM == 3 ? 4 : 3;
    const int W = L.W, ws = L.wshift, nf = L.nf;
    double sq = 0.0;
    for (int64_t vb = vb_begin + blockIdx.x; vb < nvirtual; vb += gridDim.x) {
        if (vb < npair_blocks) {
            const int npc = DIM == 3 ? L.npf : L.npe;
            const uint16_t* tab = L.iface_idx;   // 3D: faces first; 2D: edges first
            const int l = threadIdx.x & (W - 1), ks = threadIdx.x >> ws, nk = blockDim.x >> ws;
            const int64_t u = vb / NF;
            const int f = (int)(vb - u * NF);
            const int64_t e = u * W + l;
            if (e >= T.ne) continue;
            const int32_t pr = T.partner[e * 4 + f];
            if (pr < 0 || (pr >> 3) < e) continue;       // the lower owner drives the pair
            const int64_t pe = pr >> 3;
            const int pf = pr & 7;
            double* A = x + u * (int64_t)nf * W + l;
            double* B = x + (pe >> ws) * (int64_t)nf * W + (pe & (W - 1));
            const uint16_t* ta = tab + f * npc;
            const uint16_t* tb = tab + pf * npc;
            // four nodes per step: all eight loads are in flight before the first store
            constexpr int U = 4;
            for (int k0 = ks; k0 < npc; k0 += nk * U) {
                int64_t oa[U], ob[U];
                double va[U], vb_[U];
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    const int k = min(k0 + q * nk, npc - 1);
                    oa[q] = (int64_t)__ldg(ta + k) * W;
                    ob[q] = (int64_t)__ldg(tb + k) * W;
                }
                if (OP == 0) {
#pragma unroll
                    for (int q = 0; q < U; ++q) {
                        HMG_DEV_ASSERT(oa[q] >= 0 && oa[q] < (int64_t)nf * W && ob[q] >= 0 && ob[q] < (int64_t)nf * W && pe < T.ne);
                        va[q] = A[oa[q]]; vb_[q] = B[ob[q]];
                    }
                }
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    if (k0 + q * nk >= npc) break;
                    if (OP == 0) {
                        const double sum = va[q] + vb_[q];
                        A[oa[q]] = sum;
                        B[ob[q]] = sum;
                        if (SQ) sq = fma(2.0 * sum, sum, sq);
                    } else {
                        B[ob[q]] = 0.0;
                    }
                }
            }
            continue;
        }
        // multi-owner cells: MULTI_ITEMS cells per virtual block, MULTI_G threads each (thread j <-> owner j); the
        // owner list is read once per cell, the shared nodes of the cell are walked four at a time
        const int nel = DIM == 3 ? 6 : 3, nfl = DIM == 3 ? 4 : 0;
        const int64_t nedge_cells = DIM == 3 && L.npe > 0 ? T.nedges : 0;
        const int64_t total = nedge_cells + T.nverts;
        const int sub = threadIdx.x & (MULTI_G - 1);
        const int lane0 = (threadIdx.x & 31) & ~(MULTI_G - 1);         // first lane of the group
        const unsigned gmask = ((1u << MULTI_G) - 1u) << lane0;
        const int64_t t = (vb - npair_blocks) * MULTI_ITEMS + (threadIdx.x / MULTI_G);
        if (t >= total) continue;                                      // whole groups leave together
        const int64_t* off;
        const int32_t* own;
        const uint16_t* tab;
        int64_t cell;
        int npc;
        if (t < nedge_cells) {
            cell = t;
            npc = L.npe;
            off = T.edge_off; own = T.edge_own;
            tab = L.iface_idx + nfl * L.npf;
        } else {
            cell = t - nedge_cells;
            npc = 1;
            off = T.vert_off; own = T.vert_own;
            tab = L.iface_idx + nfl * L.npf + nel * L.npe;
        }
        const int64_t b = off[cell], en = off[cell + 1];
        const int n = (int)(en - b);
        // column of owner o and its row table
        auto column = [&](int64_t o, const uint16_t*& rows) -> double* {
            const int32_t id = __ldg(own + o);
            const int64_t el = id >> 3;
            rows = tab + (id & 7) * npc;
            HMG_DEV_ASSERT(o >= b && o < en && el >= 0 && el < T.ne && (id & 7) < (t < nedge_cells ? nel : DIM + 1));
            return x + (el >> ws) * (int64_t)nf * W + (el & (W - 1));
        };
        if (OP == 1) {
            for (int64_t o = b + 1 + sub; o < en; o += MULTI_G) {
                const uint16_t* rows;
                double* col = column(o, rows);
                for (int k = 0; k < npc; ++k) col[(int64_t)__ldg(rows + k) * W] = 0.0;
            }
            continue;
        }
        if (n <= MULTI_G) {
            const bool have = sub < n;
            const uint16_t* rows;
            double* col = column(b + (have ? sub : 0), rows);
            constexpr int U = 4;
            for (int k0 = 0; k0 < npc; k0 += U) {
                int64_t at[U];
                double v[U];
#pragma unroll
                for (int q = 0; q < U; ++q) at[q] = (int64_t)__ldg(rows + min(k0 + q, npc - 1)) * W;
#pragma unroll
                for (int q = 0; q < U; ++q) v[q] = have ? col[at[q]] : 0.0;
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    double sum = 0.0;
                    for (int j = 0; j < n; ++j) sum += __shfl_sync(gmask, v[q], lane0 + j);      // ascending owner order
                    if (k0 + q < npc) {
                        if (have) col[at[q]] = sum;
                        if (SQ && sub == 0) sq = fma((double)n * sum, sum, sq);
                    }
                }
            }
            continue;
        }
        // more owners than threads in the group (vertices of dense meshes): rounds of MULTI_G owners per node
        for (int k = 0; k < npc; ++k) {
            double sum = 0.0;
            for (int64_t o0 = b; o0 < en; o0 += MULTI_G) {
                const bool have = o0 + sub < en;
                const uint16_t* rows;
                double* col = column(have ? o0 + sub : b, rows);
                const double v = have ? col[(int64_t)__ldg(rows + k) * W] : 0.0;
                const int cnt = (int)min((int64_t)MULTI_G, en - o0);
                for (int j = 0; j < cnt; ++j) sum += __shfl_sync(gmask, v, lane0 + j);
            }
            for (int64_t o = b + sub; o < en; o += MULTI_G) {
                const uint16_t* rows;
                double* col = column(o, rows);
                col[(int64_t)__ldg(rows + k) * W] = sum;
            }
            if (SQ && sub == 0) sq = fma((double)n * sum, sum, sq);
        }
    }
    if (SQ) block_reduce_finish(sq, R, post, S_TMP);
}

static unsigned grid_for(int64_t n, int block, int max_blocks = 148 * 16) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (unsigned)g;
}

// returns the number of launches (0: the level has no shared cell on this rank).  part: 3 = everything, 1 = only the
// two-owner cells (faces 3D / edges 2D), 2 = only the cells with more owners -- the partial runs exist for hmg_time_op.
template <int OP, bool SQ>
static int launch_interface(int dim, const LevelView& L, const TopoView& T, double* x, const Reducer& R, int post, cudaStream_t st,
                            int part = 3) {
    const int npc = dim == 3 ? L.npf : L.npe;
    const int64_t nunits = (T.ne + L.W - 1) / L.W;
    const int64_t npair_blocks = npc > 0 ? nunits * (dim == 3 ? 4 : 3) : 0;
    const int64_t multi = (dim == 3 && L.npe > 0 ? T.nedges : 0) + T.nverts;       // cells
    const int64_t nmulti_blocks = (multi + MULTI_ITEMS - 1) / MULTI_ITEMS;
    const int64_t vb_begin = (part & 1) ? 0 : npair_blocks;
    const int64_t nvirtual = (part & 2) ? npair_blocks + nmulti_blocks : npair_blocks;
    if (nvirtual <= vb_begin) return 0;
    const unsigned grid = (unsigned)(SQ ? std::min<int64_t>(nvirtual - vb_begin, R.max_blocks) : nvirtual - vb_begin);
    if (dim == 3) interface_kernel<3, OP, SQ><<<grid, 256, 0, st>>>(L, T, npair_blocks, vb_begin, nvirtual, x, R, post);
    else interface_kernel<2, OP, SQ><<<grid, 256, 0, st>>>(L, T, npair_blocks, vb_begin, nvirtual, x, R, post);
    return 1;
}
int launch_interface_sum(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st, int part) {
    return launch_interface<0, false>(dim, L, T, x, Reducer{}, 0, st, part);
}
int launch_interface_sum_sq(int dim, const LevelView& L, const TopoView& T, double* x, const Reducer& R, int post, cudaStream_t st) {
    return launch_interface<0, true>(dim, L, T, x, R, post, st);
}
int launch_zero_all_but_one(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st) {
    return launch_interface<1, false>(dim, L, T, x, Reducer{}, 0, st);
}

// Cut cells (owners on several ranks).  One thread per (cell, paired node); faces, edges and vertices in ONE launch.
struct CutAll {
    CutView kind[3];
    int npc[3];               // paired nodes per cell
    const uint16_t* tab[3];   // packed node index of the t-th paired node of every local cell
    int64_t items[4];         // prefix sums of ncells * npc
};
static CutAll make_cut_all(int dim, const LevelView& L, const CutView* C) {
    const int nfl = dim == 3 ? 4 : 0, nel = dim == 3 ? 6 : 3;
    CutAll A;
    A.items[0] = 0;
    for (int kd = 0; kd < 3; ++kd) {
        A.kind[kd] = C[kd];
        A.npc[kd] = kd == 0 ? L.npf : (kd == 1 ? L.npe : 1);
        A.tab[kd] = L.iface_idx + (kd == 0 ? 0 : (kd == 1 ? nfl * L.npf : nfl * L.npf + nel * L.npe));
        A.items[kd + 1] = A.items[kd] + C[kd].ncells * A.npc[kd];
        if (A.npc[kd] == 0) A.npc[kd] = 1;      // never divided by: the kind has no item
    }
    return A;
}
// zero_out_all_but_one! across ranks: every local copy of a cut node but the one of the GLOBALLY first owner
// (src/implicit_fine_grid.jl:334-386; first_local says whether that owner lives here)
__global__ void __launch_bounds__(256) cut_zero_kernel(const LevelView L, const CutAll A, double* __restrict__ x) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < A.items[3]; t += (int64_t)gridDim.x * blockDim.x) {
        const int kd = t < A.items[1] ? 0 : (t < A.items[2] ? 1 : 2);
        const CutView& C = A.kind[kd];
        const int npc = A.npc[kd];
        const int64_t q = t - A.items[kd];
        const int64_t cell = q / npc;
        const int k = (int)(q - cell * npc);
        const int64_t b = C.off[cell], en = C.off[cell + 1];
        for (int64_t o = b; o < en; ++o) {
            if (o == b && C.first_local[cell]) continue;
            const int32_t id = C.own[o];
            const int64_t el = id >> 3;
            x[((el >> L.wshift) * (int64_t)L.nf + __ldg(A.tab[kd] + (id & 7) * npc + k)) * L.W + (el & (L.W - 1))] = 0.0;
        }
    }
}
int launch_cut_zero_but_first(int dim, const LevelView& L, const CutView* C, double* x, cudaStream_t st) {
    const CutAll A = make_cut_all(dim, L, C);
    if (A.items[3] == 0) return 0;
    cut_zero_kernel<<<grid_for(A.items[3], 256), 256, 0, st>>>(L, A, x);
    return 1;
}

// PEER: the messages travel through peer memory instead of ncclSend / ncclRecv.  CUT_PACK stores every partial sum
// straight into the receive area (even / odd exchange) of the rank it is meant for; when the last block is done
// (system-scope fences, a ticket) it raises this rank's flag on every neighbour to the exchange number.  CUT_UNPACK
// first waits until every neighbour's flag has reached the exchange number, then reads its own receive area.  Two
// receive areas suffice: a rank packs exchange j + 2 only after it unpacked j + 1, which needed every neighbour's
// message j + 1, which the neighbour sent after it had unpacked j.
template <int OP, bool SQ, bool PEER>
__global__ void __launch_bounds__(256) cut_p2p_kernel(const LevelView L, const CutAll A, const int64_t* __restrict__ kbase,
                                                      double* __restrict__ x, double* __restrict__ msg, const Reducer R,
                                                      const CutPeer CP, int sq_post) {
    double sq = 0.0;
    const PeerView& P = R.peer;
    unsigned long long xs = 0;
    const double* inbox = msg;
    if (PEER) {
        xs = *P.xseq + 1;
        if (OP == CUT_UNPACK) {
            if ((int)threadIdx.x < CP.nnbr) peer_wait_ge(P.flag[P.rank] + CP.nbr[threadIdx.x], xs);
            __syncthreads();
            inbox = P.recv[P.rank] + (xs & 1ull) * P.recv_stride;
        }
    }
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < A.items[3]; t += (int64_t)gridDim.x * blockDim.x) {
        const int kd = t < A.items[1] ? 0 : (t < A.items[2] ? 1 : 2);
        const CutView& C = A.kind[kd];
        const int npc = A.npc[kd];
        const int64_t q = t - A.items[kd];
        const int64_t cell = q / npc;
        const int k = (int)(q - cell * npc);
        const int64_t b = C.off[cell], en = C.off[cell + 1];
        double part = 0.0;
        for (int64_t o = b; o < en; ++o) {
            const int32_t id = C.own[o];
            const int64_t el = id >> 3;
            part += x[((el >> L.wshift) * (int64_t)L.nf + __ldg(A.tab[kd] + (id & 7) * npc + k)) * L.W + (el & (L.W - 1))];
        }
        const int64_t pb = C.peer_off[cell], pe = C.peer_off[cell + 1];
        if (OP == CUT_PACK) {
            for (int64_t j = pb; j < pe; ++j) {
                const int pr = C.peer_rank[j];
                if (PEER) P.recv[pr][(xs & 1ull) * P.recv_stride + CP.rbase[pr * 3 + kd] + (int64_t)C.peer_idx[j] * npc + k] = part;
                else msg[kbase[pr * 3 + kd] + (int64_t)C.peer_idx[j] * npc + k] = part;
            }
            continue;
        }
        const int mine = C.my_pos[cell];
        double tot = 0.0;
        for (int64_t j = pb; j < pe; ++j) {
            if ((int)(j - pb) == mine) tot += part;
            const int64_t at = kbase[C.peer_rank[j] * 3 + kd] + (int64_t)C.peer_idx[j] * npc + k;
            tot += PEER ? __ldcg(inbox + at) : inbox[at];      // peers wrote it: never from a stale L1 line
        }
        if ((int)(pe - pb) == mine) tot += part;
        for (int64_t o = b; o < en; ++o) {
            const int32_t id = C.own[o];
            const int64_t el = id >> 3;
            x[((el >> L.wshift) * (int64_t)L.nf + __ldg(A.tab[kd] + (id & 7) * npc + k)) * L.W + (el & (L.W - 1))] = tot;
        }
        if (SQ) sq = fma((double)(en - b) * tot, tot, sq);
    }
    if (PEER) {
        // the last block to finish publishes: PACK -> the flags on the neighbours, UNPACK -> the exchange counter
        __shared__ bool last_x;
        if (OP == CUT_PACK) __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(P.xticket, 1u);
            last_x = (t == gridDim.x - 1);
            if (last_x) *P.xticket = 0u;
        }
        __syncthreads();
        if (last_x) {
            if (OP == CUT_PACK) {
                __threadfence_system();
                if ((int)threadIdx.x < CP.nnbr) st_release_sys(P.flag[CP.nbr[threadIdx.x]] + P.rank, xs);
            } else if (threadIdx.x == 0) {
                *P.xseq = xs;
            }
        }
    }
    if (SQ) block_reduce_finish(sq, R, sq_post, S_TMP);
}
int launch_cut_p2p(int dim, int op, const LevelView& L, const CutView* C, const int64_t* kbase, double* x, double* msg, bool sq,
                   const Reducer& R, cudaStream_t st, const CutPeer* peer, int sq_post) {
    const CutAll A = make_cut_all(dim, L, C);
    // (with peer memory the kernels also run for a rank without items: the exchange is collective)
    if (A.items[3] == 0 && !sq && !peer) return 0;
    const unsigned grid = grid_for(std::max<int64_t>(A.items[3], 1), 256, sq ? R.max_blocks : 148 * 16);
    const CutPeer cp = peer ? *peer : CutPeer{};
    if (peer) {
        if (op == CUT_PACK) cut_p2p_kernel<CUT_PACK, false, true><<<grid, 256, 0, st>>>(L, A, kbase, x, msg, R, cp, sq_post);
        else if (sq) cut_p2p_kernel<CUT_UNPACK, true, true><<<grid, 256, 0, st>>>(L, A, kbase, x, msg, R, cp, sq_post);
        else cut_p2p_kernel<CUT_UNPACK, false, true><<<grid, 256, 0, st>>>(L, A, kbase, x, msg, R, cp, sq_post);
        return 1;
    }
    if (op == CUT_PACK) cut_p2p_kernel<CUT_PACK, false, false><<<grid, 256, 0, st>>>(L, A, kbase, x, msg, R, cp, sq_post);
    else if (sq) cut_p2p_kernel<CUT_UNPACK, true, false><<<grid, 256, 0, st>>>(L, A, kbase, x, msg, R, cp, sq_post);
    else cut_p2p_kernel<CUT_UNPACK, false, false><<<grid, 256, 0, st>>>(L, A, kbase, x, msg, R, cp, sq_post);
    return 1;
}

// apply_constraint!: zero every stored node whose class is on the domain boundary; only elements
// that touch the domain boundary are visited (belems)
__global__ void __launch_bounds__(256) constraint_kernel(const LevelView L, int64_t nbelems, const int32_t* __restrict__ belems,
                                                         const uint16_t* __restrict__ cmask, double* __restrict__ x) {
    const int64_t total = nbelems * L.n_boundary;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t be = t / L.n_boundary;
        const int q = (int)(t - be * L.n_boundary);
        const int64_t e = belems[be];
        const uint32_t u = L.boundary[q];
        if ((cmask[e] >> (u >> 14)) & 1u) x[((e >> L.wshift) * (int64_t)L.nf + (u & 0x3fff)) * L.W + (e & (L.W - 1))] = 0.0;
    }
}
int launch_apply_constraint(int, const LevelView& L, int64_t nbelems, const int32_t* belems, const uint16_t* cmask,
                            double* x, cudaStream_t st) {
    if (nbelems * L.n_boundary == 0) return 0;
    constraint_kernel<<<grid_for(nbelems * L.n_boundary, 256), 256, 0, st>>>(L, nbelems, belems, cmask, x);
    return 1;
}

// ------------------------------------------------------------------------------------------
// K4: restriction / interpolation (column-local, lattice form; lanes are elements)
// ------------------------------------------------------------------------------------------
template <int NDIR>
__global__ void __launch_bounds__(256) restrict_kernel(const LevelView Lf, const LevelView Lc, int64_t nunits,
                                                       const double* __restrict__ rf, double* __restrict__ bc) {
    const int W = Lf.W, ws = Lf.wshift;
    const int l = threadIdx.x & (W - 1);
    const int64_t nodes = nunits * Lc.nf;       // (unit, coarse node) items
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> ws;
    for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> ws; it < nodes; it += stride) {
        const int64_t u = it / Lc.nf;
        const int pc = (int)(it - u * Lc.nf);
        const uint16_t* tab = Lf.restrict_tab + pc * NDIR;
        const double* r = rf + u * (int64_t)Lf.nf * W + l;
        double s = 0.0;
#pragma unroll
        for (int d = 1; d < NDIR; ++d) {
            const unsigned q = __ldg(tab + d);
            if (q != 0xFFFFu) s += r[(int64_t)q * W];
        }
        bc[(u * Lc.nf + pc) * W + l] = r[(int64_t)__ldg(tab) * W] + 0.5 * s;
    }
}

__global__ void __launch_bounds__(256) interp_kernel(const LevelView Lf, const LevelView Lc, int64_t nunits,
                                                     double* __restrict__ xf, const double* __restrict__ xc) {
    const int W = Lf.W, ws = Lf.wshift;
    const int l = threadIdx.x & (W - 1);
    const int64_t nodes = nunits * Lf.nf;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> ws;
    for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> ws; it < nodes; it += stride) {
        const int64_t u = it / Lf.nf;
        const int p = (int)(it - u * Lf.nf);
        const unsigned v = __ldg(Lf.interp_tab + p);
        const double* c = xc + u * (int64_t)Lc.nf * W + l;
        xf[it * W + l] += 0.5 * c[(int64_t)(v & 0xFFFFu) * W] + 0.5 * c[(int64_t)(v >> 16) * W];
    }
}

int launch_restrict(int dim, const LevelView& Lf, const LevelView& Lc, int64_t nunits, const double* rf, double* bc, cudaStream_t st) {
    if (nunits == 0) return 0;
    const unsigned grid = grid_for(nunits * Lc.nf * Lf.W, 256, 148 * 32);
    if (dim == 3) restrict_kernel<15><<<grid, 256, 0, st>>>(Lf, Lc, nunits, rf, bc);
    else restrict_kernel<7><<<grid, 256, 0, st>>>(Lf, Lc, nunits, rf, bc);
    return 1;
}
int launch_interp_add(int, const LevelView& Lf, const LevelView& Lc, int64_t nunits, double* xf, const double* xc, cudaStream_t st) {
    if (nunits == 0) return 0;
    interp_kernel<<<grid_for(nunits * Lf.nf * Lf.W, 256, 148 * 32), 256, 0, st>>>(Lf, Lc, nunits, xf, xc);
    return 1;
}

// ------------------------------------------------------------------------------------------
// K3: reductions and fused CG vector updates (scalars stay on the device)
// ------------------------------------------------------------------------------------------
// the post-op of a reduction whose total earlier kernels left in S_TMP; with POST_GLOBAL that total is first summed over
// all ranks (peer memory)
__global__ void __launch_bounds__(32) scalar_post_kernel(const Reducer R, int post, int slot) {
    double* S = R.scalars;
    double tot = S[S_TMP];
    if (post & POST_GLOBAL) tot = peer_allreduce(R.peer, tot);
    if (threadIdx.x == 0) scalar_post(S, post & 0xff, slot, tot);
}
int launch_scalar_post(const Reducer& R, int post, int slot, cudaStream_t st) {
    scalar_post_kernel<<<1, 32, 0, st>>>(R, post, slot);
    return 1;
}

__global__ void __launch_bounds__(256) dot_kernel(const Reducer R, const double* __restrict__ a, const double* __restrict__ b,
                                                  int64_t n, int post, int slot) {
    double s = 0.0;
    const int64_t n2 = n >> 1;
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* b2 = reinterpret_cast<const double2*>(b);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        const double2 u = a2[t], v = b2[t];
        s = fma(u.x, v.x, s);
        s = fma(u.y, v.y, s);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) s = fma(a[n - 1], b[n - 1], s);
    block_reduce_finish(s, R, post, slot);
}

// p = r ; rho = dot(r, r)   (src/multigrid.jl:53-54)
__global__ void __launch_bounds__(256) copy_dot_kernel(const Reducer R, const double* __restrict__ r, double* __restrict__ p, int64_t n, int post) {
    double s = 0.0;
    const int64_t n2 = n >> 1;
    const double2* r2 = reinterpret_cast<const double2*>(r);
    double2* p2 = reinterpret_cast<double2*>(p);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = r2[t];
        p2[t] = v;
        s = fma(v.x, v.x, s);
        s = fma(v.y, v.y, s);
    }
    block_reduce_finish(s, R, post, S_TMP);
}

// x += alpha p ; r -= alpha Ap ; rsqr = dot(r, r) -> beta, rho   (src/multigrid.jl:64-68)
// FIRST: the first step of a smoothing call, where the search direction IS the residual: p is written
// (p = r_old) instead of read, which replaces the p = r copy of src/multigrid.jl:53
template <bool FIRST>
__global__ void __launch_bounds__(256) cg_update_kernel(const Reducer R, double* __restrict__ x, double* __restrict__ p,
                                                        double* __restrict__ r, const double* __restrict__ Ap, int64_t n, int post) {
    const double alpha = R.scalars[S_ALPHA];
    double s = 0.0;
    const int64_t n2 = n >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    double2* p2 = reinterpret_cast<double2*>(p);
    const double2* q2 = reinterpret_cast<const double2*>(Ap);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        double2 xv = x2[t], rv = r2[t];
        const double2 qv = q2[t];
        double2 pv;
        if (FIRST) { pv = rv; p2[t] = rv; } else pv = p2[t];
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
        x2[t] = xv; r2[t] = rv;
        s = fma(rv.x, rv.x, s); s = fma(rv.y, rv.y, s);
    }
    block_reduce_finish(s, R, post, S_TMP);
}

// p = r + beta p   (src/multigrid.jl:68)
__global__ void __launch_bounds__(256) p_update_kernel(const double* __restrict__ scalars, double* __restrict__ p,
                                                       const double* __restrict__ r, int64_t n) {
    const double beta = scalars[S_BETA];
    const int64_t n2 = n >> 1;
    double2* p2 = reinterpret_cast<double2*>(p);
    const double2* r2 = reinterpret_cast<const double2*>(r);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        double2 pv = p2[t];
        const double2 rv = r2[t];
        pv.x = fma(beta, pv.x, rv.x); pv.y = fma(beta, pv.y, rv.y);
        p2[t] = pv;
    }
}

// x += alpha p with alpha on the device: all that is left of the last CG step of a smoothing call whose residual nobody
// reads (src/multigrid.jl:65; the r-update, rho' and the p-update of that step are dead)
__global__ void __launch_bounds__(256) x_update_kernel(const double* __restrict__ scalars, double* __restrict__ x,
                                                       const double* __restrict__ p, int64_t n) {
    const double alpha = scalars[S_ALPHA];
    const int64_t n2 = n >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    const double2* p2 = reinterpret_cast<const double2*>(p);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        double2 xv = x2[t];
        const double2 pv = p2[t];
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        x2[t] = xv;
    }
}

__global__ void __launch_bounds__(256) axpy_kernel(double alpha, const double* __restrict__ x, double* __restrict__ y, int64_t n) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        y[t] = fma(alpha, x[t], y[t]);
}
__global__ void __launch_bounds__(256) fill_kernel(double* __restrict__ x, double v, int64_t n) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) x[t] = v;
}

static const int kVecBlocks = 148 * 8;

int launch_dot(const Reducer& R, const double* a, const double* b, int64_t n, int post, int slot, cudaStream_t st) {
    dot_kernel<<<grid_for(n / 2 + 1, 256, R.max_blocks), 256, 0, st>>>(R, a, b, n, post, slot);
    return 1;
}
int launch_copy_dot(const Reducer& R, const double* r, double* p, int64_t n, int post, cudaStream_t st) {
    copy_dot_kernel<<<grid_for(n / 2 + 1, 256, R.max_blocks), 256, 0, st>>>(R, r, p, n, post);
    return 1;
}
int launch_cg_update(const Reducer& R, double* x, double* p, double* r, const double* Ap, int64_t n, int post, bool first, cudaStream_t st) {
    if (first) cg_update_kernel<true><<<grid_for(n / 2 + 1, 256, R.max_blocks), 256, 0, st>>>(R, x, p, r, Ap, n, post);
    else cg_update_kernel<false><<<grid_for(n / 2 + 1, 256, R.max_blocks), 256, 0, st>>>(R, x, p, r, Ap, n, post);
    return 1;
}
int launch_p_update(const Reducer& R, double* p, const double* r, int64_t n, cudaStream_t st) {
    p_update_kernel<<<grid_for(n / 2 + 1, 256, kVecBlocks), 256, 0, st>>>(R.scalars, p, r, n);
    return 1;
}
int launch_x_update(const Reducer& R, double* x, const double* p, int64_t n, cudaStream_t st) {
    x_update_kernel<<<grid_for(n / 2 + 1, 256, kVecBlocks), 256, 0, st>>>(R.scalars, x, p, n);
    return 1;
}
int launch_axpy(double alpha, const double* x, double* y, int64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    axpy_kernel<<<grid_for(n, 256, kVecBlocks), 256, 0, st>>>(alpha, x, y, n);
    return 1;
}
int launch_fill(double* x, double v, int64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    if (v == 0.0) { cudaMemsetAsync(x, 0, n * sizeof(double), st); return 1; }
    fill_kernel<<<grid_for(n, 256, kVecBlocks), 256, 0, st>>>(x, v, n);
    return 1;
}

// fill!(x, v) for v != 0: padded columns stay zero
__global__ void __launch_bounds__(256) fill_columns_kernel(const LevelView L, int64_t ne, double* __restrict__ x, double v) {
    const int64_t nunits = (ne + L.W - 1) >> L.wshift;
    const int64_t total = nunits * L.nf * L.W;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = t / ((int64_t)L.nf * L.W);
        const int64_t e = u * L.W + (t & (L.W - 1));
        x[t] = e < ne ? v : 0.0;
    }
}
int launch_fill_columns(const LevelView& L, int64_t ne, double* x, double v, cudaStream_t st) {
    if (ne == 0) return 0;
    const int64_t nunits = (ne + L.W - 1) >> L.wshift;
    fill_columns_kernel<<<grid_for(nunits * L.nf * L.W, 256, kVecBlocks), 256, 0, st>>>(L, ne, x, v);
    return 1;
}

// Column prefix of another context's vector (domain shrink, src/examples/homogenized_coefficients.jl:54-60: l.x[:, OneTo(n)]):
// both layouts agree on the units of the prefix, only the last unit gets zero columns where the prefix ends.
__global__ void __launch_bounds__(256) copy_columns_kernel(const LevelView L, int64_t ne, double* __restrict__ dst,
                                                           const double* __restrict__ src) {
    const int64_t nunits = (ne + L.W - 1) >> L.wshift;
    const int64_t per_unit = (int64_t)L.nf * L.W;
    const int64_t total = nunits * per_unit;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = (t / per_unit) * L.W + (t & (L.W - 1));
        dst[t] = e < ne ? src[t] : 0.0;
    }
}
int launch_copy_columns(const LevelView& L, int64_t ne, double* dst, const double* src, cudaStream_t st) {
    if (ne == 0) return 0;
    const int64_t nunits = (ne + L.W - 1) >> L.wshift;
    copy_columns_kernel<<<grid_for(nunits * L.nf * L.W, 256, kVecBlocks), 256, 0, st>>>(L, ne, dst, src);
    return 1;
}

// ------------------------------------------------------------------------------------------
// driver functionals on the finest level (src/examples/homogenized_coefficients.jl:449-474, 592-667)
// ------------------------------------------------------------------------------------------
// b[p, e] = dot(dphi[p], flux_e): rhs_a_xi_grad_v! (un-summed, local)
template <int DIM>
__global__ void __launch_bounds__(256) rhs_flux_kernel(const LevelView L, int64_t nunits, const double* __restrict__ dphi,
                                                       const double* __restrict__ flux, double* __restrict__ b) {
    const int W = L.W, ws = L.wshift;
    const int l = threadIdx.x & (W - 1);
    const int64_t rows = nunits * L.nf;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> ws;
    for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> ws; it < rows; it += stride) {
        const int64_t u = it / L.nf;
        const int p = (int)(it - u * L.nf);
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) s = fma(__ldg(dphi + p * DIM + d), __ldg(flux + (u * DIM + d) * W + l), s);
        b[it * W + l] = s;
    }
}
int launch_rhs_flux(int dim, const LevelView& L, int64_t nunits, const double* dphi, const double* flux, double* b, cudaStream_t st) {
    if (nunits == 0) return 0;
    const unsigned grid = grid_for(nunits * L.nf * L.W, 256, 148 * 32);
    if (dim == 3) rhs_flux_kernel<3><<<grid, 256, 0, st>>>(L, nunits, dphi, flux, b);
    else rhs_flux_kernel<2><<<grid, 256, 0, st>>>(L, nunits, dphi, flux, b);
    return 1;
}
// sum_e w_e sum_p (v[p,e] + v2[p,e]) * (dot(dphi[p], flux_e) + Mv[p,e]),  w_e = |J_e| if the element's global index
// is below nsubset else 0; v2 / flux may be null.  integrate_first_term / integrate_terms.
template <int DIM>
__global__ void __launch_bounds__(256) integrate_kernel(const Reducer R, const LevelView L, int64_t nunits, int64_t nsubset,
                                                        const int32_t* __restrict__ gidx, const double* __restrict__ coef,
                                                        const double* __restrict__ dphi, const double* __restrict__ flux,
                                                        const double* __restrict__ v, const double* __restrict__ v2,
                                                        const double* __restrict__ Mv) {
    using D = Dims<DIM>;
    const int W = L.W, ws = L.wshift;
    const int l = threadIdx.x & (W - 1);
    const int64_t rows = nunits * L.nf;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> ws;
    double acc = 0.0;
    for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> ws; it < rows; it += stride) {
        const int64_t u = it / L.nf;
        const int p = (int)(it - u * L.nf);
        const int32_t gi = __ldg(gidx + u * W + l);
        if (gi < 0 || gi >= nsubset) continue;
        const double w = __ldg(coef + (u * D::CS + D::NC - 1) * W + l);
        double t = Mv[it * W + l];
        if (flux) {
#pragma unroll
            for (int d = 0; d < DIM; ++d) t = fma(__ldg(dphi + p * DIM + d), __ldg(flux + (u * DIM + d) * W + l), t);
        }
        double a = v[it * W + l];
        if (v2) a += v2[it * W + l];
        acc = fma(w * a, t, acc);
    }
    block_reduce_finish(acc, R, POST_STORE, S_TMP);
}
int launch_integrate(int dim, const Reducer& R, const LevelView& L, int64_t nunits, int64_t nsubset, const int32_t* gidx,
                     const double* coef, const double* dphi, const double* flux, const double* v, const double* v2,
                     const double* Mv, cudaStream_t st) {
    const unsigned grid = grid_for(std::max<int64_t>(1, nunits * L.nf * L.W), 256, R.max_blocks);
    if (dim == 3) integrate_kernel<3><<<grid, 256, 0, st>>>(R, L, nunits, nsubset, gidx, coef, dphi, flux, v, v2, Mv);
    else integrate_kernel<2><<<grid, 256, 0, st>>>(R, L, nunits, nsubset, gidx, coef, dphi, flux, v, v2, Mv);
    return 1;
}

// ------------------------------------------------------------------------------------------
// host layout (hierarchical rows, unpadded columns) <-> device layout (lattice rows, interleaved)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) permute_in_kernel(const LevelView L, const int32_t* __restrict__ h2l,
                                                         const double* __restrict__ staged, int64_t lds,
                                                         double* __restrict__ dst, int64_t e0, int64_t ncols) {
    const int64_t total = ncols * L.nf;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / L.nf;
        const int h = (int)(t - c * L.nf);
        const int64_t e = e0 + c;
        dst[((e >> L.wshift) * (int64_t)L.nf + h2l[h]) * L.W + (e & (L.W - 1))] = staged[c * lds + h];
    }
}
__global__ void __launch_bounds__(256) permute_out_kernel(const LevelView L, const int32_t* __restrict__ h2l,
                                                          const double* __restrict__ src, double* __restrict__ staged,
                                                          int64_t lds, int64_t e0, int64_t ncols) {
    const int64_t total = ncols * L.nf;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / L.nf;
        const int h = (int)(t - c * L.nf);
        const int64_t e = e0 + c;
        staged[c * lds + h] = src[((e >> L.wshift) * (int64_t)L.nf + h2l[h]) * L.W + (e & (L.W - 1))];
    }
}
// W = 32, e0 a multiple of 32: 32 x 32 tiles through shared memory, so that both sides move whole 256-byte lines
// (the host side is contiguous along a column, the device side along the 32 columns of a unit)
template <bool IN>
__global__ void __launch_bounds__(256) permute_tile_kernel(int nf, const int32_t* __restrict__ h2l, double* __restrict__ staged,
                                                           int64_t lds, double* __restrict__ dev, int64_t e0, int64_t ncols) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t ntr = (nf + 31) / 32;
    const int64_t ntiles = ((ncols + 31) / 32) * ntr;
    for (int64_t b = blockIdx.x; b < ntiles; b += gridDim.x) {
        const int64_t ub = b / ntr;                       // unit of the chunk
        const int h0 = (int)(b - ub * ntr) * 32;
        const int64_t cb = ub * 32;                       // first column of the unit inside the chunk
        double* unit = dev + ((e0 + cb) >> 5) * (int64_t)nf * 32;
        if (IN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t c = cb + ty + 8 * j;
                const int h = h0 + tx;
                tile[ty + 8 * j][tx] = (c < ncols && h < nf) ? staged[c * lds + h] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int h = h0 + ty + 8 * j;
                if (h < nf) unit[(int64_t)__ldg(h2l + h) * 32 + tx] = tile[tx][ty + 8 * j];     // padding columns: zero
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int h = h0 + ty + 8 * j;
                tile[tx][ty + 8 * j] = h < nf ? unit[(int64_t)__ldg(h2l + h) * 32 + tx] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t c = cb + ty + 8 * j;
                const int h = h0 + tx;
                if (c < ncols && h < nf) staged[c * lds + h] = tile[ty + 8 * j][tx];
            }
        }
        __syncthreads();
    }
}
int launch_permute_in(const LevelView& L, const int32_t* h2l, const double* staged, int64_t lds, double* dst, int64_t e0,
                      int64_t ncols, cudaStream_t st) {
    if (ncols == 0) return 0;
    if (L.W == 32 && (e0 & 31) == 0) {
        const int64_t ntiles = ((ncols + 31) / 32) * ((L.nf + 31) / 32);
        permute_tile_kernel<true><<<grid_for(ntiles, 1, 148 * 16), 256, 0, st>>>(L.nf, h2l, const_cast<double*>(staged), lds, dst, e0, ncols);
    } else {
        permute_in_kernel<<<grid_for(ncols * L.nf, 256), 256, 0, st>>>(L, h2l, staged, lds, dst, e0, ncols);
    }
    return 1;
}
int launch_permute_out(const LevelView& L, const int32_t* h2l, const double* src, double* staged, int64_t lds, int64_t e0,
                       int64_t ncols, cudaStream_t st) {
    if (ncols == 0) return 0;
    if (L.W == 32 && (e0 & 31) == 0) {
        const int64_t ntiles = ((ncols + 31) / 32) * ((L.nf + 31) / 32);
        permute_tile_kernel<false><<<grid_for(ntiles, 1, 148 * 16), 256, 0, st>>>(L.nf, h2l, staged, lds, const_cast<double*>(src), e0, ncols);
    } else {
        permute_out_kernel<<<grid_for(ncols * L.nf, 256), 256, 0, st>>>(L, h2l, src, staged, lds, e0, ncols);
    }
    return 1;
}

// The first `nrows` hierarchical rows of the columns e0 .. e0+ncols (the nodes of a coarser level, src/examples/
// homogenized_coefficients.jl:84: x[1 : nnodes(refined_mesh(implicit, level)), :]); lanes run over the columns, so the
// big array is read in whole 256-byte lines.
__global__ void __launch_bounds__(256) permute_rows_out_kernel(const LevelView L, const int32_t* __restrict__ h2l,
                                                               const double* __restrict__ src, double* __restrict__ staged,
                                                               int64_t lds, int nrows, int64_t e0, int64_t ncols) {
    const int64_t total = ncols * nrows;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(t / ncols);
        const int64_t c = t - (int64_t)h * ncols;
        const int64_t e = e0 + c;
        staged[c * lds + h] = src[((e >> L.wshift) * (int64_t)L.nf + __ldg(h2l + h)) * L.W + (e & (L.W - 1))];
    }
}
int launch_permute_rows_out(const LevelView& L, const int32_t* h2l, const double* src, double* staged, int64_t lds, int nrows,
                            int64_t e0, int64_t ncols, cudaStream_t st) {
    if (ncols == 0 || nrows == 0) return 0;
    permute_rows_out_kernel<<<grid_for(ncols * nrows, 256), 256, 0, st>>>(L, h2l, src, staged, lds, nrows, e0, ncols);
    return 1;
}

// ------------------------------------------------------------------------------------------
// level 1 <-> base vector, coarse solve helpers
// ------------------------------------------------------------------------------------------
__global__ void copy_to_base_kernel(const LevelView L1, int64_t nn, const int32_t* __restrict__ first,
                                    const double* __restrict__ v, double* __restrict__ u) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
        const int32_t id = first[n];
        if (id >= 0) {
            const int64_t e = id >> 3;
            u[n] = v[((e >> L1.wshift) * (int64_t)L1.nf + L1.vpos[id & 7]) * L1.W + (e & (L1.W - 1))];
        }
    }
}
__global__ void distribute_kernel(const LevelView L1, int nv, int64_t ne, const int32_t* __restrict__ elems,
                                  const double* __restrict__ u, double* __restrict__ v) {
    const int64_t total = ne * nv;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / nv;
        const int a = (int)(t - e * nv);
        v[((e >> L1.wshift) * (int64_t)L1.nf + L1.vpos[a]) * L1.W + (e & (L1.W - 1))] = u[elems[t]];
    }
}
__global__ void gather_kernel(const int64_t* __restrict__ idx, int64_t n, const double* __restrict__ src, double* __restrict__ dst) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) dst[t] = src[idx[t]];
}
__global__ void scatter_kernel(const int64_t* __restrict__ idx, int64_t n, const double* __restrict__ src, double* __restrict__ dst) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) dst[idx[t]] = src[t];
}
// multi-GPU: every base node is reported by exactly one rank (the others contribute zero to the reduction)
__global__ void masked_copy_to_base_kernel(const LevelView L1, int64_t nn, const int32_t* __restrict__ first,
                                           const uint8_t* __restrict__ contrib, const double* __restrict__ v, double* __restrict__ u) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
        const int32_t id = first[n];
        double val = 0.0;
        if (id >= 0 && contrib[n]) {
            const int64_t e = id >> 3;
            val = v[((e >> L1.wshift) * (int64_t)L1.nf + L1.vpos[id & 7]) * L1.W + (e & (L1.W - 1))];
        }
        u[n] = val;
    }
}
int launch_masked_copy_to_base(const LevelView& L1, int64_t nn, const int32_t* node_first, const uint8_t* contrib, const double* v,
                               double* u, cudaStream_t st) {
    masked_copy_to_base_kernel<<<grid_for(nn, 256), 256, 0, st>>>(L1, nn, node_first, contrib, v, u);
    return 1;
}
int launch_copy_to_base(const LevelView& L1, int64_t nn, const int32_t* first, const double* v, double* u, cudaStream_t st) {
    copy_to_base_kernel<<<grid_for(nn, 256), 256, 0, st>>>(L1, nn, first, v, u);
    return 1;
}
int launch_distribute(int dim, const LevelView& L1, int64_t ne, const int32_t* elems, const double* u, double* v, cudaStream_t st) {
    if (ne == 0) return 0;
    distribute_kernel<<<grid_for(ne * (dim + 1), 256), 256, 0, st>>>(L1, dim + 1, ne, elems, u, v);
    return 1;
}
int launch_gather(const int64_t* idx, int64_t n, const double* src, double* dst, cudaStream_t st) {
    if (n == 0) return 0;
    gather_kernel<<<grid_for(n, 256), 256, 0, st>>>(idx, n, src, dst);
    return 1;
}
int launch_scatter(const int64_t* idx, int64_t n, const double* src, double* dst, cudaStream_t st) {
    if (n == 0) return 0;
    scatter_kernel<<<grid_for(n, 256), 256, 0, st>>>(idx, n, src, dst);
    return 1;
}

// A (column-major, lower triangle valid) -> full symmetric
__global__ void symmetrize_kernel(double* __restrict__ A, int64_t n) {
    const int64_t total = n * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / n, r = t - c * n;
        if (r < c) A[c * n + r] = A[r * n + c];
    }
}
int launch_symmetrize_lower(double* A, int64_t n, cudaStream_t st) {
    symmetrize_kernel<<<grid_for(n * n, 256), 256, 0, st>>>(A, n);
    return 1;
}
__global__ void set_diagonal_kernel(double* __restrict__ A, int64_t n, double v) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) A[t * n + t] = v;
}
int launch_set_diagonal(double* A, int64_t n, double v, cudaStream_t st) {
    if (n == 0) return 0;
    set_diagonal_kernel<<<grid_for(n, 256), 256, 0, st>>>(A, n, v);
    return 1;
}
// y = A x for a full symmetric column-major matrix: one warp per column (= row), coalesced
__global__ void __launch_bounds__(256) symv_kernel(const double* __restrict__ A, int64_t n, const double* __restrict__ x, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < n; c += nwarps) {
        const double* col = A + c * n;
        double s = 0.0;
        for (int64_t r = lane; r < n; r += 32) s = fma(col[r], x[r], s);
        s = warp_sum(s);
        if (lane == 0) y[c] = s;
    }
}
// y = A x reading only the lower-triangular 128 x 128 tiles of the (fully stored) symmetric matrix: a tile (I, J),
// I >= J, contributes T x_J to the rows of tile I and T' x_I to the rows of tile J.  Both land in per-tile slots
// (PR[J][rows of I], PC[I][rows of J]) that a second kernel adds in a fixed order: half the traffic of symv_kernel,
// deterministic, no atomics.
constexpr int SYMV_TS = 128;
__global__ void __launch_bounds__(256) symv_tiles_kernel(const double* __restrict__ A, int64_t n, int nt, const double* __restrict__ x,
                                                         double* __restrict__ PR, double* __restrict__ PC) {
    __shared__ double us[8][SYMV_TS];
    // triangular decode: block b -> (I, J) with I >= J, b = I (I + 1) / 2 + J
    const int64_t b = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((int64_t)(I + 1) * (I + 2) / 2 <= b) ++I;
    while ((int64_t)I * (I + 1) / 2 > b) --I;
    const int J = (int)(b - (int64_t)I * (I + 1) / 2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)I * SYMV_TS, c0 = (int64_t)J * SYMV_TS;
    double xi[4], u[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = r0 + lane + 32 * k;
        xi[k] = r < n ? x[r] : 0.0;
    }
    for (int cc = warp; cc < SYMV_TS; cc += 8) {
        const int64_t c = c0 + cc;
        if (c >= n) break;
        const double xj = x[c];
        const double* col = A + c * n + r0;
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rr = lane + 32 * k;
            const double t = r0 + rr < n ? col[rr] : 0.0;
            u[k] = fma(t, xj, u[k]);
            v = fma(t, xi[k], v);
        }
        if (I != J) {
            v = warp_sum(v);
            if (lane == 0) PC[(int64_t)I * n + c] = v;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) us[warp][lane + 32 * k] = u[k];
    __syncthreads();
    if (threadIdx.x < SYMV_TS) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += us[w][threadIdx.x];
        const int64_t r = r0 + threadIdx.x;
        if (r < n) PR[(int64_t)J * n + r] = sum;
    }
}
__global__ void __launch_bounds__(256) symv_combine_kernel(int64_t n, int nt, const double* __restrict__ PR,
                                                           const double* __restrict__ PC, double* __restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int T = (int)(i / SYMV_TS);
        double s = 0.0;
        for (int J = 0; J <= T; ++J) s += PR[(int64_t)J * n + i];
        for (int I = T + 1; I < nt; ++I) s += PC[(int64_t)I * n + i];
        y[i] = s;
    }
}
// work: 2 * nt * n doubles (PR then PC), nt = ceil(n / 128)
int launch_symv_half(const double* A, int64_t n, const double* x, double* y, double* work, cudaStream_t st) {
    if (n == 0) return 0;
    const int nt = (int)((n + SYMV_TS - 1) / SYMV_TS);
    double* PR = work;
    double* PC = work + (int64_t)nt * n;
    symv_tiles_kernel<<<(unsigned)((int64_t)nt * (nt + 1) / 2), 256, 0, st>>>(A, n, nt, x, PR, PC);
    symv_combine_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, nt, PR, PC, y);
    return 2;
}
int launch_symv_full(const double* A, int64_t n, const double* x, double* y, cudaStream_t st) {
    if (n == 0) return 0;
    symv_kernel<<<grid_for(n * 32, 256), 256, 0, st>>>(A, n, x, y);
    return 1;
}

}  // namespace hmg
