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
#include <algorithm>

#include "kernels.cuh"
#include "lattice.hpp"

namespace hmg {

constexpr int TASK_INTS = 52, SLOT_INTS = 12;   // = PLAN_TASK_INTS / PLAN_SLOT_INTS of hmg_host.hpp
constexpr int MAX_SLOTS = 8;

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

// ------------------------------------------------------------------------------------------
// K1: local operator apply
// ------------------------------------------------------------------------------------------
// Face set a stencil direction points out of: a neighbour n+d of a node of class `cls` lies outside
// the simplex iff (cls & out_mask(d)) != 0 (the lattice simplex is convex and d has entries in {-1,0,1}).
template <int DIM> __host__ __device__ constexpr int out_mask(int d) {
    if (DIM == 3) {
        constexpr int I[15] = {0, 1, -1, 0, 0, 0, 0, -1, 1, -1, 1, 0, 0, 1, -1};
        constexpr int J[15] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, -1, 1, -1, 1};
        constexpr int K[15] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1};
        return (K[d] < 0 ? 1 : 0) | (J[d] < 0 ? 2 : 0) | (I[d] < 0 ? 4 : 0) | (I[d] + J[d] + K[d] > 0 ? 8 : 0);
    }
    constexpr int I2[7] = {0, 1, -1, 0, 0, -1, 1};
    constexpr int J2[7] = {0, 0, 0, 1, -1, 1, -1};
    return (J2[d] < 0 ? 1 : 0) | (I2[d] < 0 ? 2 : 0) | (I2[d] + J2[d] > 0 ? 4 : 0);
}
__host__ __device__ constexpr int opp_dir(int d) { return d == 0 ? 0 : ((d & 1) ? d + 1 : d - 1); }
// weight of direction d in the stencil of a node of a single-face class, relative to the interior
// stencil: 0 dropped (points outside), 1 = one half (direction inside the face: half of the fine
// elements around that edge exist), 2 = full (points inwards).  Verified against the assembled
// tables at setup (reference.cpp).
template <int DIM> __host__ __device__ constexpr int wcode(int cls, int d) {
    return cls == 0 ? 2
                    : ((cls & out_mask<DIM>(d)) ? 0 : (d == 0 ? 1 : ((cls & out_mask<DIM>(opp_dir(d))) ? 2 : 1)));
}

// directions of the taps of a line sweep along the fastest lattice coordinate.  Centre line: -k, 0,
// +k; NP "minus" lines with taps at k and k-1; NP "plus" lines with taps at k and k+1; tap t of
// minus line q is the opposite direction of tap t of plus line q.
template <int DIM> struct Sweep;
template <> struct Sweep<3> {
    static constexpr int NP = 3, KP = 5, KM = 6, FIRST = 1, LAST = 8, FACE_A = 2, FACE_B = 4;
    __host__ __device__ static constexpr int m0(int q) { return q == 0 ? 1 : (q == 1 ? 3 : 7); }     // (1,0,0) (0,1,0) (-1,1,0)
    __host__ __device__ static constexpr int m1(int q) { return q == 0 ? 10 : (q == 1 ? 12 : 14); }  // ... + (0,0,-1)
    __host__ __device__ static constexpr int p0(int q) { return q == 0 ? 2 : (q == 1 ? 4 : 8); }     // (-1,0,0) (0,-1,0) (1,-1,0)
    __host__ __device__ static constexpr int p1(int q) { return q == 0 ? 9 : (q == 1 ? 11 : 13); }   // ... + (0,0,1)
};
template <> struct Sweep<2> {
    static constexpr int NP = 1, KP = 3, KM = 4, FIRST = 1, LAST = 4, FACE_A = 2, FACE_B = 2;
    __host__ __device__ static constexpr int m0(int) { return 1; }   // (1,0)
    __host__ __device__ static constexpr int m1(int) { return 6; }   // (1,-1)
    __host__ __device__ static constexpr int p0(int) { return 2; }   // (-1,0)
    __host__ __device__ static constexpr int p1(int) { return 5; }   // (-1,1)
};

// per-lane (= per coarse element) operator data
template <int DIM> struct LaneOp {
    double ec[Dims<DIM>::NC];                      // |J| P (upper triangle), lambda |J|
    double c0, cz, ca[Sweep<DIM>::NP], cb[Sweep<DIM>::NP];   // interior stencil, one value per +-pair
    unsigned cm;                                   // Dirichlet class mask
};

template <int DIM, int CLS>
__device__ __forceinline__ double eval_node(const LaneOp<DIM>& op, double xm, double x0, double xp,
                                            const double* Mm, const double* Mk, const double* Pk, const double* Pp) {
    using S = Sweep<DIM>;
    if (CLS == 0) {
        double a = op.c0 * x0;
        a = fma(op.cz, xm + xp, a);
#pragma unroll
        for (int q = 0; q < S::NP; ++q) {
            a = fma(op.ca[q], Mk[q] + Pk[q], a);
            a = fma(op.cb[q], Mm[q] + Pp[q], a);
        }
        return a;
    }
    double a1 = 0.0, ah = op.c0 * x0;    // the diagonal of a face node is one half of the interior one
    {
        constexpr int wp = wcode<DIM>(CLS, S::KP), wm = wcode<DIM>(CLS, S::KM);
        if (wp == 2) a1 = fma(op.cz, xp, a1); else if (wp == 1) ah = fma(op.cz, xp, ah);
        if (wm == 2) a1 = fma(op.cz, xm, a1); else if (wm == 1) ah = fma(op.cz, xm, ah);
    }
#pragma unroll
    for (int q = 0; q < S::NP; ++q) {
        const int w0 = wcode<DIM>(CLS, S::m0(q)), w1 = wcode<DIM>(CLS, S::m1(q));
        const int w2 = wcode<DIM>(CLS, S::p0(q)), w3 = wcode<DIM>(CLS, S::p1(q));
        if (w0 == 2) a1 = fma(op.ca[q], Mk[q], a1); else if (w0 == 1) ah = fma(op.ca[q], Mk[q], ah);
        if (w2 == 2) a1 = fma(op.ca[q], Pk[q], a1); else if (w2 == 1) ah = fma(op.ca[q], Pk[q], ah);
        if (w1 == 2) a1 = fma(op.cb[q], Mm[q], a1); else if (w1 == 1) ah = fma(op.cb[q], Mm[q], ah);
        if (w3 == 2) a1 = fma(op.cb[q], Pp[q], a1); else if (w3 == 1) ah = fma(op.cb[q], Pp[q], ah);
    }
    return fma(0.5, ah, a1);
}

// output of one node: AX  y = fixed ? 0 : acc ; RESIDUAL  r = fixed ? 0 : b - acc ; MULADD  y += alpha acc
struct OutCtx {
    double* y;             // lane pointer at packed node 0 of the unit
    const double* t;       // b (RESIDUAL) or y (MULADD) or nullptr (AX)
    double sa;             // 1, -1 or alpha
};
__device__ __forceinline__ void store_node(const OutCtx& o, int64_t off, double acc, bool fixed, double t) {
    o.y[off] = fixed ? 0.0 : fma(o.sa, acc, t);
}

template <int W> __device__ __forceinline__ int ref_offset(int ref, int b0, int b1, int b2, int bz, int l) {
    const unsigned sel = (unsigned)ref >> 28;
    const int base = sel == 0 ? b0 : (sel == 1 ? b1 : (sel == 2 ? b2 : bz));
    return base + (ref & 0x0fffffff) * W + l;
}

// one line sweep per row slot: nodes kstart .. kstart+cnt-1 of a lattice line, classes RCLS (|FIRST/LAST)
template <int DIM, int W, int RCLS>
__device__ __forceinline__ void sweep_task(const double* __restrict__ sm, const int* __restrict__ d, int b0, int b1, int b2,
                                           int bz, int l, const LaneOp<DIM>& op, const OutCtx& out) {
    using S = Sweep<DIM>;
    constexpr int NP = S::NP;
    const int cnt = d[7];
    const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
    const int oc = ref_offset<W>(d[0], b0, b1, b2, bz, l);
    int om[NP], opl[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        om[q] = ref_offset<W>(d[1 + q], b0, b1, b2, bz, l);
        opl[q] = ref_offset<W>(d[4 + q], b0, b1, b2, bz, l);
    }
    const int64_t pout = (int64_t)d[8] * W;
    const bool ff = d[9] & 1, lf = d[9] & 2;
    double xm = sm[oc - W], x0 = sm[oc];
    double Mm[NP], Pk[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) { Mm[q] = sm[om[q] - W]; Pk[q] = sm[opl[q]]; }
#pragma unroll 2
    for (int s = 0; s < maxcnt; ++s) {
        if (s < cnt) {
            const double t = out.t ? out.t[pout + (int64_t)s * W] : 0.0;
            const double xp = sm[oc + (s + 1) * W];
            double Mk[NP], Pp[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) { Mk[q] = sm[om[q] + s * W]; Pp[q] = sm[opl[q] + (s + 1) * W]; }
            double acc;
            int cls = RCLS;
            if (RCLS == 0 && ff && s == 0) {
                acc = eval_node<DIM, S::FIRST>(op, xm, x0, xp, Mm, Mk, Pk, Pp);
                cls = S::FIRST;
            } else if (RCLS == 0 && lf && s == cnt - 1) {
                acc = eval_node<DIM, S::LAST>(op, xm, x0, xp, Mm, Mk, Pk, Pp);
                cls = S::LAST;
            } else {
                acc = eval_node<DIM, RCLS>(op, xm, x0, xp, Mm, Mk, Pk, Pp);
            }
            store_node(out, pout + (int64_t)s * W, acc, (op.cm >> cls) & 1u, t);
            xm = x0; x0 = xp;
#pragma unroll
            for (int q = 0; q < NP; ++q) { Mm[q] = Mk[q]; Pk[q] = Pp[q]; }
        }
    }
}

// generic path for the nodes on reference edges / vertices: coefficients from the class table
template <int DIM, int W>
__device__ __forceinline__ void node_task(const double* __restrict__ sm, int idx, const uint32_t* __restrict__ nodetab,
                                          const double* __restrict__ G, int b0, int b1, int b2, int bz, int l,
                                          const LaneOp<DIM>& op, const OutCtx& out) {
    using D = Dims<DIM>;
    if (idx < 0) return;
    const uint32_t* e = nodetab + (size_t)idx * 16;
    const uint32_t info = __ldg(e + 15);
    const int p = info & 0xffff, cls = info >> 16;
    const double t = out.t ? out.t[(int64_t)p * W] : 0.0;
    const double* g = G + (size_t)cls * D::NDIR * D::NC;
    double acc = 0.0;
#pragma unroll 1
    for (int dd = 0; dd < D::NDIR; ++dd) {
        const uint32_t ref = __ldg(e + dd);
        if (ref == 0xFFFFFFFFu) continue;
        double c = 0.0;
#pragma unroll
        for (int q = 0; q < D::NC; ++q) c = fma(op.ec[q], __ldg(g + dd * D::NC + q), c);
        acc = fma(c, sm[ref_offset<W>((int)ref, b0, b1, b2, bz, l)], acc);
    }
    store_node(out, (int64_t)p * W, acc, (op.cm >> cls) & 1u, t);
}

// Persistent kernel.  Warp `nwarps` is the TMA producer, warps 0..nwarps-1 consume.  CTA b handles the
// units b, b + gridDim.x, ...; the chunks of consecutive units form one stream through the ring.
template <int DIM, int W, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) apply_kernel(const ApplyArgs a) {
    using D = Dims<DIM>;
    using S = Sweep<DIM>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[MAX_SLOTS], empty_bar[MAX_SLOTS];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const ApplyPlanView& P = a.P;
    const int K = P.nslots, nch = P.nchunks, NW = P.nwarps;
    const int SD = P.slot_doubles, ZD = P.zero_doubles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nf = a.L.nf;

    for (int q = threadIdx.x; q < ZD + K * SD; q += blockDim.x) sm[q] = 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < K; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NW); }
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();

    if (warp == NW) {
        // ---------------- producer ----------------
        if (lane == 0) {
            unsigned n = 0;
            for (int64_t u = blockIdx.x; u < a.nunits; u += gridDim.x) {
                const double* src = a.x + u * (int64_t)nf * W;
                for (int c = 0; c < nch; ++c, ++n) {
                    const unsigned s = n % K, use = n / K;
                    if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1u);
                    const int c0 = __ldg(P.chunk_start + c), c1 = __ldg(P.chunk_start + c + 1);
                    const uint32_t bytes = (uint32_t)(c1 - c0) * W * 8u;
                    mbar_expect_tx(&full_bar[s], bytes);
                    bulk_g2s(sm + ZD + (size_t)s * SD, src + (int64_t)c0 * W, bytes, &full_bar[s]);
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int l = lane % W, slot = lane / W;
    unsigned n0 = 0, waited = 0, released = 0;
    OutCtx out;
    out.sa = a.mode == APPLY_AX ? 1.0 : (a.mode == APPLY_RESIDUAL ? -1.0 : a.alpha);
    for (int64_t u = blockIdx.x; u < a.nunits; u += gridDim.x, n0 += nch) {
        LaneOp<DIM> op;
#pragma unroll
        for (int q = 0; q < D::NC; ++q) op.ec[q] = __ldg(a.coef + (u * D::CS + q) * W + l);
        op.ec[D::NC - 1] *= a.lambda;
        op.cm = a.mode == APPLY_MULADD ? 0u : (unsigned)__ldg(a.cmask + u * W + l);
        {
            auto coefI = [&](int dir) {
                double c = 0.0;
#pragma unroll
                for (int q = 0; q < D::NC; ++q) c = fma(op.ec[q], __ldg(a.L.G + dir * D::NC + q), c);
                return c;
            };
            op.c0 = coefI(0);
            op.cz = coefI(S::KP);
#pragma unroll
            for (int q = 0; q < S::NP; ++q) { op.ca[q] = coefI(S::m0(q)); op.cb[q] = coefI(S::m1(q)); }
        }
        const int64_t ubase = u * (int64_t)nf * W + l;
        out.y = a.y + ubase;
        out.t = a.mode == APPLY_AX ? nullptr : (a.mode == APPLY_RESIDUAL ? a.b + ubase : a.y + ubase);

        for (int t = warp; t < P.ntasks; t += NW) {
            const int32_t* T = P.tasks + (size_t)t * TASK_INTS;
            const int type = __ldg(T), clo = __ldg(T + 1), chi = __ldg(T + 2);
            // Advance this warp's view of the chunk stream: observe every chunk up to n0+chi, hand back
            // every chunk below n0+clo.  A slot is only released after its chunk was seen (an early
            // arrival would be counted in the previous phase of the empty barrier), and releases are
            // not postponed behind a wait (the producer may need them to load what we wait for).
            {
                const unsigned need = n0 + chi, lo = n0 + clo;
                while (waited <= need || released < lo) {
                    if (released < lo && released < waited) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[released % K]);
                        ++released;
                    } else {
                        mbar_wait(&full_bar[waited % K], (waited / K) & 1u);
                        ++waited;
                    }
                }
            }
            const int b0 = ZD + (int)((n0 + clo) % K) * SD;
            const int b1 = ZD + (int)((n0 + clo + 1) % K) * SD;
            const int b2 = ZD + (int)((n0 + clo + 2) % K) * SD;
            int d[SLOT_INTS];
            {
                const int4* src = reinterpret_cast<const int4*>(T + 4 + slot * SLOT_INTS);
                const int4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
                d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w;
                d[4] = v1.x; d[5] = v1.y; d[6] = v1.z; d[7] = v1.w;
                d[8] = v2.x; d[9] = v2.y; d[10] = v2.z; d[11] = v2.w;
            }
            if (type == 0) sweep_task<DIM, W, 0>(sm, d, b0, b1, b2, 0, l, op, out);
            else if (type == 1) sweep_task<DIM, W, S::FACE_A>(sm, d, b0, b1, b2, 0, l, op, out);
            else if (type == 2) sweep_task<DIM, W, S::FACE_B>(sm, d, b0, b1, b2, 0, l, op, out);
            else node_task<DIM, W>(sm, d[0], P.nodetab, a.L.G, b0, b1, b2, 0, l, op, out);
        }
    }
    // hand every remaining slot back: other warps may still need chunks that reuse them
    while (released < n0) {
        if (waited <= released) {
            mbar_wait(&full_bar[waited % K], (waited / K) & 1u);
            ++waited;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[released % K]);
        ++released;
    }
}

template <int DIM, int W>
static int launch_apply_w(const ApplyArgs& a, int ctas_per_sm, size_t smem, cudaStream_t st) {
    auto kern = ctas_per_sm >= 2 ? apply_kernel<DIM, W, 288, 2> : apply_kernel<DIM, W, 288, 1>;
    static size_t configured[2] = {0, 0};
    const int ki = ctas_per_sm >= 2 ? 1 : 0;
    if (smem > configured[ki]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
        configured[ki] = smem;
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int threads = (a.P.nwarps + 1) * 32;
    const int64_t maxgrid = (int64_t)sms * (ctas_per_sm >= 2 ? 2 : 1);
    dim3 grid((unsigned)std::min<int64_t>(a.nunits, maxgrid));
    kern<<<grid, threads, smem, st>>>(a);
    return 1;
}

int launch_apply(int dim, const ApplyArgs& a, int ctas_per_sm, size_t smem_bytes, cudaStream_t st) {
    if (a.nunits == 0) return 0;
    const int W = a.L.W;
    if (dim == 3) return W == 16 ? launch_apply_w<3, 16>(a, ctas_per_sm, smem_bytes, st) : launch_apply_w<3, 8>(a, ctas_per_sm, smem_bytes, st);
    return W == 16 ? launch_apply_w<2, 16>(a, ctas_per_sm, smem_bytes, st) : launch_apply_w<2, 8>(a, ctas_per_sm, smem_bytes, st);
}

// ------------------------------------------------------------------------------------------
// K2: interface sums
// ------------------------------------------------------------------------------------------
// Codimension-1 cells (3D faces, 2D edges) have exactly two owners: lanes are the elements of a
// unit, the lower owner of a pair reads both copies, adds them in ascending owner order
// (src/implicit_fine_grid.jl:219-244) and writes both.  Cells with more owners (3D edges, vertices)
// are processed cell by cell: one thread per shared fine node.  OP 0: sum + broadcast; OP 1: zero
// all but the first owner (src/implicit_fine_grid.jl:334-386).
template <int DIM, int OP>
__global__ void __launch_bounds__(256) interface_kernel(const LevelView L, const TopoView T, int64_t npair_blocks,
                                                        double* __restrict__ x) {
    constexpr int NF = DIM == 3 ? 4 : 3;
    const int W = L.W, ws = L.wshift, nf = L.nf;
    if ((int64_t)blockIdx.x < npair_blocks) {
        const int npc = DIM == 3 ? L.npf : L.npe;
        const uint16_t* tab = L.iface_idx;   // 3D: faces first; 2D: edges first
        const int l = threadIdx.x & (W - 1), ks = threadIdx.x >> ws, nk = blockDim.x >> ws;
        const int64_t u = blockIdx.x / NF;
        const int f = (int)(blockIdx.x - u * NF);
        const int64_t e = u * W + l;
        if (e >= T.ne) return;
        const int32_t pr = T.partner[e * 4 + f];
        if (pr < 0 || (pr >> 3) < e) return;       // the lower owner drives the pair
        const int64_t pe = pr >> 3;
        const int pf = pr & 7;
        double* A = x + u * (int64_t)nf * W + l;
        double* B = x + (pe >> ws) * (int64_t)nf * W + (pe & (W - 1));
        const uint16_t* ta = tab + f * npc;
        const uint16_t* tb = tab + pf * npc;
        for (int k = ks; k < npc; k += nk) {
            const int64_t oa = (int64_t)__ldg(ta + k) * W, ob = (int64_t)__ldg(tb + k) * W;
            if (OP == 0) {
                const double s = A[oa] + B[ob];
                A[oa] = s;
                B[ob] = s;
            } else {
                B[ob] = 0.0;
            }
        }
        return;
    }
    // multi-owner cells
    const int nel = DIM == 3 ? 6 : 3, nfl = DIM == 3 ? 4 : 0;
    const int64_t nedge_items = DIM == 3 ? T.nedges * L.npe : 0;
    const int64_t total = nedge_items + T.nverts;
    const int64_t nb = gridDim.x - npair_blocks;
    for (int64_t t = ((int64_t)blockIdx.x - npair_blocks) * blockDim.x + threadIdx.x; t < total; t += nb * blockDim.x) {
        const int64_t* off;
        const int32_t* own;
        const uint16_t* tab;
        int64_t cell;
        int k, npc;
        if (t < nedge_items) {
            cell = t / L.npe;
            k = (int)(t - cell * L.npe);
            npc = L.npe;
            off = T.edge_off; own = T.edge_own;
            tab = L.iface_idx + nfl * L.npf;
        } else {
            cell = t - nedge_items;
            k = 0; npc = 1;
            off = T.vert_off; own = T.vert_own;
            tab = L.iface_idx + nfl * L.npf + nel * L.npe;
        }
        const int64_t b = off[cell], en = off[cell + 1];
        double s = 0.0;
        for (int64_t o = b; o < en; ++o) {
            const int32_t id = own[o];
            const int64_t el = id >> 3;
            double* ptr = x + ((el >> ws) * (int64_t)nf + __ldg(tab + (id & 7) * npc + k)) * W + (el & (W - 1));
            if (OP == 0) s += *ptr;
            else if (o > b) *ptr = 0.0;
        }
        if (OP == 0)
            for (int64_t o = b; o < en; ++o) {
                const int32_t id = own[o];
                const int64_t el = id >> 3;
                x[((el >> ws) * (int64_t)nf + __ldg(tab + (id & 7) * npc + k)) * W + (el & (W - 1))] = s;
            }
    }
}

static unsigned grid_for(int64_t n, int block, int max_blocks = 148 * 16) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (unsigned)g;
}

template <int OP>
static int launch_interface(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st) {
    const int npc = dim == 3 ? L.npf : L.npe;
    const int64_t nunits = (T.ne + L.W - 1) / L.W;
    const int64_t npair_blocks = npc > 0 ? nunits * (dim == 3 ? 4 : 3) : 0;
    const int64_t multi = (dim == 3 ? T.nedges * L.npe : 0) + T.nverts;
    const int64_t nmulti_blocks = multi > 0 ? grid_for(multi, 256) : 0;
    if (npair_blocks + nmulti_blocks == 0) return 0;
    const unsigned grid = (unsigned)(npair_blocks + nmulti_blocks);
    if (dim == 3) interface_kernel<3, OP><<<grid, 256, 0, st>>>(L, T, npair_blocks, x);
    else interface_kernel<2, OP><<<grid, 256, 0, st>>>(L, T, npair_blocks, x);
    return 1;
}
int launch_interface_sum(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st) {
    return launch_interface<0>(dim, L, T, x, st);
}
int launch_zero_all_but_one(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st) {
    return launch_interface<1>(dim, L, T, x, st);
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
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic: fixed grid, fixed per-block tree, partials summed in block order by the last block
__device__ __forceinline__ void block_reduce_finish(double v, const Reducer& R, int post, int slot) {
    __shared__ double wsum[8];
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
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += wsum[w];
        double* S = R.scalars;
        if (post == POST_STORE) S[slot] = tot;
        else if (post == POST_RHO) S[S_RHO] = tot;
        else if (post == POST_PAP) { S[S_PAP] = tot; S[S_ALPHA] = S[S_RHO] / tot; }
        else if (post == POST_RSQR) { S[S_RSQR] = tot; S[S_BETA] = tot / S[S_RHO]; S[S_RHO] = tot; }
        *R.ticket = 0u;
    }
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
__global__ void __launch_bounds__(256) copy_dot_kernel(const Reducer R, const double* __restrict__ r, double* __restrict__ p, int64_t n) {
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
    block_reduce_finish(s, R, POST_RHO, 0);
}

// x += alpha p ; r -= alpha Ap ; rsqr = dot(r, r) -> beta, rho   (src/multigrid.jl:64-68)
__global__ void __launch_bounds__(256) cg_update_kernel(const Reducer R, double* __restrict__ x, const double* __restrict__ p,
                                                        double* __restrict__ r, const double* __restrict__ Ap, int64_t n) {
    const double alpha = R.scalars[S_ALPHA];
    double s = 0.0;
    const int64_t n2 = n >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2* q2 = reinterpret_cast<const double2*>(Ap);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        double2 xv = x2[t], rv = r2[t];
        const double2 pv = p2[t], qv = q2[t];
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
        x2[t] = xv; r2[t] = rv;
        s = fma(rv.x, rv.x, s); s = fma(rv.y, rv.y, s);
    }
    block_reduce_finish(s, R, POST_RSQR, 0);
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
int launch_copy_dot(const Reducer& R, const double* r, double* p, int64_t n, cudaStream_t st) {
    copy_dot_kernel<<<grid_for(n / 2 + 1, 256, R.max_blocks), 256, 0, st>>>(R, r, p, n);
    return 1;
}
int launch_cg_update(const Reducer& R, double* x, const double* p, double* r, const double* Ap, int64_t n, cudaStream_t st) {
    cg_update_kernel<<<grid_for(n / 2 + 1, 256, R.max_blocks), 256, 0, st>>>(R, x, p, r, Ap, n);
    return 1;
}
int launch_p_update(const Reducer& R, double* p, const double* r, int64_t n, cudaStream_t st) {
    p_update_kernel<<<grid_for(n / 2 + 1, 256, kVecBlocks), 256, 0, st>>>(R.scalars, p, r, n);
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
int launch_permute_in(const LevelView& L, const int32_t* h2l, const double* staged, int64_t lds, double* dst, int64_t e0,
                      int64_t ncols, cudaStream_t st) {
    if (ncols == 0) return 0;
    permute_in_kernel<<<grid_for(ncols * L.nf, 256), 256, 0, st>>>(L, h2l, staged, lds, dst, e0, ncols);
    return 1;
}
int launch_permute_out(const LevelView& L, const int32_t* h2l, const double* src, double* staged, int64_t lds, int64_t e0,
                       int64_t ncols, cudaStream_t st) {
    if (ncols == 0) return 0;
    permute_out_kernel<<<grid_for(ncols * L.nf, 256), 256, 0, st>>>(L, h2l, src, staged, lds, e0, ncols);
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
int launch_symv_full(const double* A, int64_t n, const double* x, double* y, cudaStream_t st) {
    if (n == 0) return 0;
    symv_kernel<<<grid_for(n * 32, 256), 256, 0, st>>>(A, n, x, y);
    return 1;
}

}  // namespace hmg
