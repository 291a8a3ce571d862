// Hand-written sm_100a kernels of libhmg_b200 (fp64, HBM-bound; tensor cores are not used).
//
//   K1 apply_kernel        y = A x per coarse element: constant-coefficient lattice stencil on the
//                          refined reference simplex; the element's node block is staged in shared
//                          memory by a TMA bulk copy (cp.async.bulk + mbarrier).  Replaces the
//                          dim^2+1 CSC scatter-SpMVs of src/apply_local_operators.jl:93-133.
//   K2 interface_sum       sums the owners' copies of every shared face/edge/vertex node and writes
//                          the sum back (src/implicit_fine_grid.jl:209-328); gather form, no atomics.
//   K3 vector kernels      fused CG updates with device-resident scalars (src/multigrid.jl:50-69).
//   K4 transfer kernels    restriction / interpolation in lattice form (src/interpolation.jl:52-74),
//                          level-1 gather/scatter (src/implicit_fine_grid.jl:148-202).
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "kernels.cuh"
#include "lattice.hpp"

namespace hmg {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
#define d_tri lat_tri
#define d_tot3 lat_tot3
#define d_pack2 lat_pack2
#define d_pack3 lat_pack3

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// K1: local operator apply (+ fused interface sum)
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

template <int DIM, int D> struct StencilSum {
    // acc += c[d] * x[p + off[d]] for every direction whose neighbour is inside (class test is
    // warp-uniform for uniform tasks)
    static __device__ __forceinline__ double run(const double* c, const double* xp, const int* off, int cls, double acc) {
        acc = StencilSum<DIM, D - 1>::run(c, xp, off, cls, acc);
        if ((cls & out_mask<DIM>(D)) == 0) acc = fma(c[D], xp[off[D]], acc);
        return acc;
    }
};
template <int DIM> struct StencilSum<DIM, 0> {
    static __device__ __forceinline__ double run(const double* c, const double* xp, const int*, int, double acc) {
        return fma(c[0], xp[0], acc);
    }
};

// Persistent, double-buffered: gridDim.x CTAs loop over groups of EPB consecutive coarse elements;
// while a group is processed, the TMA bulk copy of the CTA's next group is already in flight.
// blockDim = (TX, EPB): TX threads (TX/32 warps) per element.
template <int DIM, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) apply_kernel(const ApplyArgs a) {
    using D = Dims<DIM>;
    constexpr int NCELL = DIM == 3 ? 14 : 6;
    constexpr int NFL = DIM == 3 ? 4 : 0, NEL = DIM == 3 ? 6 : 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ int s_cell[8][16], s_beg[8][16], s_cnt[8][16];
    const LevelView& L = a.L;
    const int ld = L.ld;
    const int epb = blockDim.y;
    const int buf = epb * ld;
    double* xs0 = reinterpret_cast<double*>(smem_raw);         // [2][epb][ld]
    double* coef = xs0 + 2 * (size_t)buf;                      // [epb][NCLS][NDIR]
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int el = threadIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t ngroups = (a.ne + epb - 1) / epb;
    const int m = L.m;
    const int mode = a.mode;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int64_t g, int b) {
        const int64_t e0 = g * epb;
        const uint32_t bytes = (uint32_t)min((int64_t)epb, a.ne - e0) * ld * 8u;
        mbar_expect_tx(&mbar[b], bytes);
        bulk_g2s(xs0 + (size_t)b * buf, a.x + e0 * ld, bytes, &mbar[b]);
    };
    if (tid == 0 && (int64_t)blockIdx.x < ngroups) issue(blockIdx.x, 0);

    int it = 0;
    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
        const int cur = it & 1;
        const int64_t e0 = g * epb;
        const int nel = (int)min((int64_t)epb, a.ne - e0);
        const int64_t e = e0 + el;
        // prefetch the next group: the other buffer was released by the barrier ending the last iteration
        if (tid == 0 && g + gridDim.x < ngroups) issue(g + gridDim.x, cur ^ 1);

        // per-element stencil coefficients for every node class, while the copies are in flight
        double* ce = coef + el * (D::NCLS * D::NDIR);
        if (el < nel) {
            double ec[D::NC];
#pragma unroll
            for (int c = 0; c < D::NC; ++c) ec[c] = __ldg(a.elem_coef + e * D::CS + c);
            ec[D::NC - 1] *= a.lambda;
            for (int t = threadIdx.x; t < D::NCLS * D::NDIR; t += blockDim.x) {
                const double* gt = L.G + t * D::NC;
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < D::NC; ++c) s = fma(ec[c], __ldg(gt + c), s);
                ce[t] = s;
            }
        }
        mbar_wait(&mbar[cur], (uint32_t)((it >> 1) & 1));
        __syncthreads();

        if (el < nel) {
            const double* xe = xs0 + (size_t)cur * buf + (size_t)el * ld;
            double* ye = a.y + e * ld;
            const double* be = a.b ? a.b + e * ld : nullptr;
            const unsigned cm = a.cmask[e];
            // uniform tasks: 32 nodes of one class, coefficients in registers; contiguous ranges per warp
            {
                const int t0 = (int)(((long)L.n_uniform * warp) / nwarps), t1 = (int)(((long)L.n_uniform * (warp + 1)) / nwarps);
                int curc = -1;
                double c[D::NDIR];
                bool fixed = false;
                for (int t = t0; t < t1; ++t) {
                    const int cls = L.task_cls[t];
                    if (cls != curc) {
#pragma unroll
                        for (int d = 0; d < D::NDIR; ++d) c[d] = ce[cls * D::NDIR + d];
                        fixed = (cm >> cls) & 1u;
                        curc = cls;
                    }
                    const uint32_t u = __ldg(L.tasks + t * 32 + lane);
                    const int p = u & 0x3fff, i = (u >> 14) & 255, j = (u >> 22) & 255;
                    int off[D::NDIR];
                    neighbour_offsets<DIM>(m, i, j, off);
                    const double acc = StencilSum<DIM, D::NDIR - 1>::run(c, xe + p, off, cls, 0.0);
                    if (mode == APPLY_AX) ye[p] = fixed ? 0.0 : acc;
                    else if (mode == APPLY_RESIDUAL) ye[p] = fixed ? 0.0 : be[p] - acc;
                    else ye[p] += a.alpha * acc;
                }
            }
            // mixed tasks (class remainders): class and coefficients per lane
            for (int t = L.n_uniform + warp; t < L.ntasks; t += nwarps) {
                const uint32_t u = __ldg(L.tasks + t * 32 + lane);
                if (u == 0xFFFFFFFFu) continue;
                const int p = u & 0x3fff, i = (u >> 14) & 255, j = (u >> 22) & 255;
                const int cls = __ldg(L.nodeinfo + p) >> 24;
                int off[D::NDIR];
                neighbour_offsets<DIM>(m, i, j, off);
                const double acc = StencilSum<DIM, D::NDIR - 1>::run(ce + cls * D::NDIR, xe + p, off, cls, 0.0);
                const bool fixed = (cm >> cls) & 1u;
                if (mode == APPLY_AX) ye[p] = fixed ? 0.0 : acc;
                else if (mode == APPLY_RESIDUAL) ye[p] = fixed ? 0.0 : be[p] - acc;
                else ye[p] += a.alpha * acc;
            }
        }

        if (a.fused) {
            // ---- fused interface sum: the last CTA to arrive at a shared cell sums the owners' partial
            // results in ascending owner order and writes the sum to every owner.  No waiting anywhere.
            __threadfence();                 // publish this thread's partial results device-wide
            __syncthreads();
            if (el < nel && threadIdx.x < NCELL) {
                const int cell = a.F.elem_cells[e * 16 + threadIdx.x];
                int mine = -1, bo = 0, cn = 0;
                if (cell >= 0) {
                    const int64_t b0 = a.F.cell_off[cell];
                    cn = (int)(a.F.cell_off[cell + 1] - b0);
                    bo = (int)b0;
                    const unsigned old = atomicAdd(a.F.arrive + cell, 1u);
                    if (old + 1u == (unsigned)cn) {
                        mine = cell;
                        a.F.arrive[cell] = 0u;   // nobody else touches this counter before the next launch
                    }
                }
                s_cell[el][threadIdx.x] = mine;
                s_beg[el][threadIdx.x] = bo;
                s_cnt[el][threadIdx.x] = cn;
            }
            __syncthreads();
            if (el < nel) {
                __threadfence();
                const int npf = L.npf, npe = L.npe;
                int total = 0;
#pragma unroll
                for (int c = 0; c < NCELL; ++c)
                    total += s_cell[el][c] >= 0 ? (c < NFL ? npf : (c < NFL + NEL ? npe : 1)) : 0;
                for (int w = threadIdx.x; w < total; w += blockDim.x) {
                    // locate the (cell, node) item in the flat list of the cells finished by this element
                    int c = 0, t = w, npc = 0, base = 0;
#pragma unroll
                    for (int q = 0; q < NCELL; ++q) {
                        const int n = q < NFL ? npf : (q < NFL + NEL ? npe : 1);
                        const int cnt = s_cell[el][q] >= 0 ? n : 0;
                        if (t >= 0 && t < cnt && npc == 0) {
                            c = q; npc = n;
                            base = q < NFL ? 0 : (q < NFL + NEL ? NFL * npf : NFL * npf + NEL * npe);
                        }
                        if (npc == 0) t -= cnt;
                    }
                    const int b0 = s_beg[el][c], cn = s_cnt[el][c];
                    const uint16_t* tab = L.iface_idx + base + t;
                    double sum = 0.0;
                    for (int o = 0; o < cn; ++o) {
                        const int32_t id = __ldg(a.F.cell_own + b0 + o);
                        sum += __ldcg(a.y + (int64_t)(id >> 3) * ld + __ldg(tab + (id & 7) * npc));
                    }
                    for (int o = 0; o < cn; ++o) {
                        const int32_t id = __ldg(a.F.cell_own + b0 + o);
                        __stcg(a.y + (int64_t)(id >> 3) * ld + __ldg(tab + (id & 7) * npc), sum);
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with this buffer and the coefficient table
    }
}

static void apply_block_shape(const LevelView& L, int& tx, int& epb) {
    // threads per element ~ nodes/4..16, a warp multiple; fill the CTA with elements on small levels
    if (L.nf >= 4096) { tx = 512; epb = 1; }
    else if (L.nf >= 768) { tx = 256; epb = 1; }
    else if (L.nf >= 384) { tx = 128; epb = 2; }
    else if (L.nf >= 128) { tx = 64; epb = 4; }
    else { tx = 32; epb = 8; }
}

template <int DIM>
static int launch_apply_t(const ApplyArgs& a, cudaStream_t st) {
    using D = Dims<DIM>;
    int tx, epb;
    apply_block_shape(a.L, tx, epb);
    const size_t smem = (size_t)epb * (2 * a.L.ld + D::NCLS * D::NDIR) * sizeof(double);
    // register budget: 64/thread (more resident CTAs) by default, 85/thread with HMG_APPLY_REGS=85
    static const bool wide = [] { const char* v = getenv("HMG_APPLY_REGS"); return v && atoi(v) > 64; }();
    const bool big = tx * epb > 256;
    auto kern = big ? (wide ? apply_kernel<DIM, 512, 1> : apply_kernel<DIM, 512, 2>)
                    : (wide ? apply_kernel<DIM, 256, 3> : apply_kernel<DIM, 256, 4>);
    static size_t configured[4] = {0, 0, 0, 0};
    const int ki = (big ? 1 : 0) + (wide ? 2 : 0);
    if (smem > configured[ki]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[ki] = smem;
    }
    // persistent grid: as many CTAs as fit on the device at once
    static int cached_key = -1, cached_blocks = 0, sms = 0;
    const int key = tx * 64 + epb * 4096 * 64 + (int)(smem / 64) % 64 + a.L.ld;
    if (key != cached_key) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_blocks, kern, tx * epb, smem);
        if (cached_blocks < 1) cached_blocks = 1;
        cached_key = key;
    }
    const int64_t ngroups = (a.ne + epb - 1) / epb;
    dim3 block(tx, epb);
    dim3 grid((unsigned)std::min<int64_t>(ngroups, (int64_t)sms * cached_blocks));
    kern<<<grid, block, smem, st>>>(a);
    return 1;
}

int launch_apply(int dim, const ApplyArgs& a, cudaStream_t st) {
    if (a.ne == 0) return 0;
    return dim == 3 ? launch_apply_t<3>(a, st) : launch_apply_t<2>(a, st);
}

// ------------------------------------------------------------------------------------------
// K2: interface sums (gather form: one thread per shared fine node, owners in ascending order)
// ------------------------------------------------------------------------------------------
// OP 0: sum + broadcast; OP 1: zero all but the first owner
template <int DIM, int OP>
__global__ void __launch_bounds__(256) interface_kernel(const LevelView L, const TopoView T, double* __restrict__ x) {
    const int m = L.m;
    const int64_t nface_items = DIM == 3 ? T.nfaces * L.npf : 0;
    const int64_t nedge_items = T.nedges * L.npe;
    const int64_t total = nface_items + nedge_items + T.nverts;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t* off;
        const int32_t* own;
        int64_t cell;
        int kind, q = 0;
        unsigned ab = 0;
        if (t < nface_items) {
            cell = t / L.npf;
            ab = L.face_bary[(int)(t - cell * L.npf)];
            kind = 0; off = T.face_off; own = T.face_own;
        } else if (t < nface_items + nedge_items) {
            const int64_t r = t - nface_items;
            cell = r / L.npe;
            q = (int)(r - cell * L.npe) + 1;                  // weight on the edge's second vertex
            kind = 1; off = T.edge_off; own = T.edge_own;
        } else {
            cell = t - nface_items - nedge_items;
            kind = 2; off = T.vert_off; own = T.vert_own;
        }
        const int64_t b = off[cell], e = off[cell + 1];
        double s = 0.0;
        for (int64_t o = b; o < e; ++o) {
            const int32_t id = own[o];
            double* ptr = x + (int64_t)(id >> 3) * L.ld + interface_node<DIM>(m, kind, id & 7, q, ab);
            if (OP == 0) s += *ptr;
            else if (o > b) *ptr = 0.0;
        }
        if (OP == 0)
            for (int64_t o = b; o < e; ++o) {
                const int32_t id = own[o];
                x[(int64_t)(id >> 3) * L.ld + interface_node<DIM>(m, kind, id & 7, q, ab)] = s;
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
    const int64_t total = (dim == 3 ? T.nfaces * L.npf : 0) + T.nedges * L.npe + T.nverts;
    if (total == 0) return 0;
    if (dim == 3) interface_kernel<3, OP><<<grid_for(total, 256), 256, 0, st>>>(L, T, x);
    else interface_kernel<2, OP><<<grid_for(total, 256), 256, 0, st>>>(L, T, x);
    return 1;
}
int launch_interface_sum(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st) {
    return launch_interface<0>(dim, L, T, x, st);
}
int launch_zero_all_but_one(int dim, const LevelView& L, const TopoView& T, double* x, cudaStream_t st) {
    return launch_interface<1>(dim, L, T, x, st);
}

// apply_constraint!: zero every stored node whose class is on the domain boundary
__global__ void __launch_bounds__(256) constraint_kernel(const LevelView L, int64_t ne, const uint16_t* __restrict__ cmask,
                                                         double* __restrict__ x) {
    const int64_t total = ne * L.n_boundary;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / L.n_boundary;
        const int q = (int)(t - e * L.n_boundary);
        const uint32_t u = L.boundary[q];
        if ((cmask[e] >> (u >> 14)) & 1u) x[e * L.ld + (u & 0x3fff)] = 0.0;
    }
}
int launch_apply_constraint(int, const LevelView& L, int64_t ne, const uint16_t* cmask, double* x, cudaStream_t st) {
    if (ne * L.n_boundary == 0) return 0;
    constraint_kernel<<<grid_for(ne * L.n_boundary, 256), 256, 0, st>>>(L, ne, cmask, x);
    return 1;
}

// ------------------------------------------------------------------------------------------
// K4: restriction / interpolation (column-local, lattice form)
// ------------------------------------------------------------------------------------------
// table-driven, one CTA per group of elements, 32-bit index arithmetic
template <int NDIR>
__global__ void __launch_bounds__(256) restrict_kernel(const LevelView Lf, const LevelView Lc, int64_t ne, int epb,
                                                       const double* __restrict__ rf, double* __restrict__ bc) {
    const int64_t e0 = (int64_t)blockIdx.x * epb;
    const int nel = (int)min((int64_t)epb, ne - e0);
    const int nfc = Lc.nf, total = nel * nfc;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int el = idx / nfc, pc = idx - el * nfc;
        const uint16_t* tab = Lf.restrict_tab + pc * NDIR;
        const double* r = rf + (e0 + el) * Lf.ld;
        double s = 0.0;
#pragma unroll
        for (int d = 1; d < NDIR; ++d) {
            const unsigned q = __ldg(tab + d);
            if (q != 0xFFFFu) s += r[q];
        }
        bc[(e0 + el) * Lc.ld + pc] = r[__ldg(tab)] + 0.5 * s;
    }
}

__global__ void __launch_bounds__(256) interp_kernel(const LevelView Lf, const LevelView Lc, int64_t ne, int epb,
                                                     double* __restrict__ xf, const double* __restrict__ xc) {
    const int64_t e0 = (int64_t)blockIdx.x * epb;
    const int nel = (int)min((int64_t)epb, ne - e0);
    const int nff = Lf.nf, total = nel * nff;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int el = idx / nff, p = idx - el * nff;
        const unsigned u = __ldg(Lf.interp_tab + p);
        const double* c = xc + (e0 + el) * Lc.ld;
        xf[(e0 + el) * Lf.ld + p] += 0.5 * c[u & 0xFFFFu] + 0.5 * c[u >> 16];
    }
}

static int elements_per_block(int nf) {
    int epb = 4096 / (nf > 0 ? nf : 1);
    return epb < 1 ? 1 : (epb > 64 ? 64 : epb);
}
int launch_restrict(int dim, const LevelView& Lf, const LevelView& Lc, int64_t ne, const double* rf, double* bc, cudaStream_t st) {
    if (ne == 0) return 0;
    const int epb = elements_per_block(Lc.nf);
    const unsigned grid = (unsigned)((ne + epb - 1) / epb);
    if (dim == 3) restrict_kernel<15><<<grid, 256, 0, st>>>(Lf, Lc, ne, epb, rf, bc);
    else restrict_kernel<7><<<grid, 256, 0, st>>>(Lf, Lc, ne, epb, rf, bc);
    return 1;
}
int launch_interp_add(int, const LevelView& Lf, const LevelView& Lc, int64_t ne, double* xf, const double* xc, cudaStream_t st) {
    if (ne == 0) return 0;
    const int epb = elements_per_block(Lf.nf);
    interp_kernel<<<(unsigned)((ne + epb - 1) / epb), 256, 0, st>>>(Lf, Lc, ne, epb, xf, xc);
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

// ------------------------------------------------------------------------------------------
// host layout (hierarchical rows, unpadded) <-> device layout (lattice rows, padded)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) permute_in_kernel(int nf, int ld, const int32_t* __restrict__ h2l,
                                                         const double* __restrict__ staged, int64_t lds,
                                                         double* __restrict__ dst, int64_t ncols) {
    const int64_t total = ncols * nf;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / nf;
        const int h = (int)(t - c * nf);
        dst[c * ld + h2l[h]] = staged[c * lds + h];
    }
}
__global__ void __launch_bounds__(256) permute_out_kernel(int nf, int ld, const int32_t* __restrict__ h2l,
                                                          const double* __restrict__ src, double* __restrict__ staged,
                                                          int64_t lds, int64_t ncols) {
    const int64_t total = ncols * nf;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / nf;
        const int h = (int)(t - c * nf);
        staged[c * lds + h] = src[c * ld + h2l[h]];
    }
}
int launch_permute_in(const LevelView& L, const int32_t* h2l, const double* staged, int64_t lds, double* dst, int64_t ncols, cudaStream_t st) {
    if (ncols == 0) return 0;
    permute_in_kernel<<<grid_for(ncols * L.nf, 256), 256, 0, st>>>(L.nf, L.ld, h2l, staged, lds, dst, ncols);
    return 1;
}
int launch_permute_out(const LevelView& L, const int32_t* h2l, const double* src, double* staged, int64_t lds, int64_t ncols, cudaStream_t st) {
    if (ncols == 0) return 0;
    permute_out_kernel<<<grid_for(ncols * L.nf, 256), 256, 0, st>>>(L.nf, L.ld, h2l, src, staged, lds, ncols);
    return 1;
}

// ------------------------------------------------------------------------------------------
// level 1 <-> base vector, coarse solve helpers
// ------------------------------------------------------------------------------------------
__global__ void copy_to_base_kernel(const LevelView L1, int64_t nn, const int32_t* __restrict__ first,
                                    const double* __restrict__ v, double* __restrict__ u) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
        const int32_t id = first[n];
        if (id >= 0) u[n] = v[(int64_t)(id >> 3) * L1.ld + L1.vpos[id & 7]];
    }
}
__global__ void distribute_kernel(const LevelView L1, int nv, int64_t ne, const int32_t* __restrict__ elems,
                                  const double* __restrict__ u, double* __restrict__ v) {
    const int64_t total = ne * nv;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / nv;
        const int a = (int)(t - e * nv);
        v[e * L1.ld + L1.vpos[a]] = u[elems[t]];
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
