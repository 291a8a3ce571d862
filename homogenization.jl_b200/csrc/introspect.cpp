// Host-only verification entry points (no GPU needed, nothing here computes on behalf of the
// product path).  They expose the tables the kernels consume, expanded through the SAME index
// functions the kernels use (lattice.hpp), so that the CPU test-suite can compare them with the
// oracle's explicit sparse operators and interface maps.
#include <cmath>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hmg.h"
#include "hmg_host.hpp"
#include "lattice.hpp"
#include "apply_core.cuh"
#include "kernels.cuh"

using namespace hmg;

namespace {
thread_local std::string g_host_err;
const RefLevel& get_level(const RefElement& ref, int level) {
    HMG_CHECK(level >= 1 && level <= ref.nlevels, "level out of range");
    return ref.lv[level - 1];
}
template <int DIM>
void local_matrix_t(const RefElement& ref, const RefLevel& L, const double* coef, double* dense) {
    using D = Dims<DIM>;
    const int nf = L.nf;
    std::vector<int> l2h(nf);
    for (int h = 0; h < nf; ++h) l2h[L.hier2lat[h]] = h;
    std::memset(dense, 0, sizeof(double) * nf * nf);
    for (int p = 0; p < nf; ++p) {
        const uint32_t info = L.nodeinfo[p];
        const int i = info & 255, j = (info >> 8) & 255, k = (info >> 16) & 255, cls = info >> 24;
        int off[D::NDIR];
        neighbour_offsets<DIM>(L.m, i, j, off);
        for (int d = 0; d < D::NDIR; ++d) {
            if (!neighbour_inside<DIM>(L.m, i, j, k, d)) continue;
            double s = 0.0;
            for (int c = 0; c < D::NC; ++c) s += coef[c] * L.G[((size_t)cls * D::NDIR + d) * D::NC + c];
            const int q = p + off[d];
            HMG_CHECK(q >= 0 && q < nf, "neighbour offset leaves the element");
            dense[(size_t)l2h[q] * nf + l2h[p]] = s;   // column-major, row = p, col = q
        }
    }
    (void)ref;
}
template <int DIM>
void transfer_matrix_t(const RefLevel& Lf, const RefLevel& Lc, double* dense) {
    std::vector<int> l2h_f(Lf.nf), l2h_c(Lc.nf);
    for (int h = 0; h < Lf.nf; ++h) l2h_f[Lf.hier2lat[h]] = h;
    for (int h = 0; h < Lc.nf; ++h) l2h_c[Lc.hier2lat[h]] = h;
    std::memset(dense, 0, sizeof(double) * Lf.nf * Lc.nf);
    for (int p = 0; p < Lf.nf; ++p) {
        const uint32_t info = Lf.nodeinfo[p];
        const int i = info & 255, j = (info >> 8) & 255, k = (info >> 16) & 255;
        int pa, pb;
        const int n = interp_parents<DIM>(Lc.m, i, j, k, pa, pb);
        if (n == 1) dense[(size_t)l2h_c[pa] * Lf.nf + l2h_f[p]] = 1.0;
        else {
            dense[(size_t)l2h_c[pa] * Lf.nf + l2h_f[p]] += 0.5;
            dense[(size_t)l2h_c[pb] * Lf.nf + l2h_f[p]] += 0.5;
        }
    }
}

// Runs the apply kernel's task enumeration and line sweeps (apply_core.cuh: the SAME templates the
// device kernel instantiates -- class dispatch, compile-time weights, sliding windows, row geometry)
// for ONE element on the host, values addressed by packed row.  y (lattice order) = A x.  Verifies
// everything of K1 except the shared-memory ring / TMA plumbing, which the -m gpu parity tests cover.
struct HostMem {
    const double* x;
    int nf;
    double operator()(int addr) const {
        HMG_CHECK(addr >= 0 && addr < nf, "sweep reads a row outside the element");
        return x[addr];
    }
    const double* ptr(int addr) const { return x + addr; }       // (may point past the element until it is read)
    double ld(const double* p) const {
        HMG_CHECK(p >= x && p < x + nf, "sweep reads a row outside the element");
        return *p;
    }
};
struct HostOut {
    static constexpr bool UNROLL4 = true;
    double* y;
    int* written;
    const uint32_t* nodeinfo;
    int row0, nf;
    template <int CLS> void put(int k, double acc, double) {
        const int p = row0 + k;
        HMG_CHECK(p >= 0 && p < nf, "sweep writes a row outside the element");
        HMG_CHECK((int)(nodeinfo[p] >> 24) == CLS, "sweep visits a node with the wrong class");
        y[p] = acc;
        written[p]++;
    }
};
template <int DIM>
void emulate_sweep_t(const RefLevel& L, int seg_shift, const double* coef, const double* x, double* y, int64_t* info) {
    using D = Dims<DIM>;
    StencilTab<DIM> T;
    static_assert(sizeof(T) == sizeof(double) * (Pairs<DIM>::N + 2 * D::NCLS) * D::NC, "table layout");
    HMG_CHECK(L.gi.size() + L.gc.size() + L.ge.size() == sizeof(T) / sizeof(double), "stencil table size");
    std::vector<double> flat = L.gi;
    flat.insert(flat.end(), L.gc.begin(), L.gc.end());
    flat.insert(flat.end(), L.ge.begin(), L.ge.end());
    std::memcpy(&T, flat.data(), sizeof(T));
    LaneOp<DIM> op;
    for (int q = 0; q < D::NC; ++q) op.ec[q] = coef[q];
    interior_coefficients(op, T);
    std::vector<int> written(L.nf, 0);
    HostMem mem{x, L.nf};
    HostOut out{y, written.data(), L.nodeinfo.data(), 0, L.nf};
    const int m = L.m;
    int64_t ntasks = 0, window = 0;
    if constexpr (DIM == 3) {
        for (int t = 0; t <= m; ++t)
            for (int i = 0; i <= t; ++i) {
                const LineRows3 r = line_rows3(m, t, i, lat_off3(m, t));
                LineGeo<3> g;
                g.L = r.L; g.k0 = 0; g.k1 = r.L; g.bc = r.c;
                for (int q = 0; q < 3; ++q) { g.bm[q] = r.rm[q]; g.bp[q] = r.rp[q]; }
                out.row0 = r.c;
                run_line3(op, T, mem, 1, g, t, i, out);
                window = std::max<int64_t>(window, r.need - r.behind + 1);
                ++ntasks;
            }
    } else {
        const int SEG = 1 << seg_shift;
        for (int i = 0; i <= m; ++i)
            for (int k0 = 0; k0 < m - i + 1; k0 += SEG) {
                const int k1 = std::min(m - i + 1, k0 + SEG);
                const LineRows2 r = line_rows2(m, i, k0, k1);
                LineGeo<2> g;
                g.L = r.L; g.k0 = k0; g.k1 = k1; g.bc = r.c; g.bm[0] = r.rm[0]; g.bp[0] = r.rp[0];
                out.row0 = r.c;
                run_line2(op, T, mem, 1, g, i, out);
                window = std::max<int64_t>(window, r.need - r.behind + 1);
                ++ntasks;
            }
    }
    for (int p = 0; p < L.nf; ++p) HMG_CHECK(written[p] == 1, "sweeps do not write every node exactly once");
    if (info) {
        const ApplyConfig cfg = make_apply_config(DIM, m, L.nf, 32);
        info[0] = ntasks; info[1] = window; info[2] = cfg.ring_rows; info[3] = (int64_t)cfg.smem_bytes;
    }
}
}  // namespace

#define HOST_BEGIN try {
#define HOST_END                                     \
    return 0;                                        \
    }                                                \
    catch (const std::exception& ex) {               \
        g_host_err = ex.what();                      \
        return 1;                                    \
    }

extern "C" {

const char* hmg_host_last_error(void) { return g_host_err.c_str(); }

// sizes[8] = m, nf, ld, n_interior, n_boundary, npf, ndir, nc ; hier2lat[nf] / G[ncls*ndir*nc] may be NULL
int hmg_host_reference(int dim, int nlevels, int level, int64_t* sizes, int32_t* hier2lat, double* G,
                       double* mass_total) {
    HOST_BEGIN
    RefElement ref = build_reference(dim, nlevels);
    const RefLevel& L = get_level(ref, level);
    if (sizes) {
        sizes[0] = L.m; sizes[1] = L.nf; sizes[2] = L.ld; sizes[3] = (int64_t)L.interior.size();
        sizes[4] = (int64_t)L.boundary.size(); sizes[5] = (int64_t)L.face_bary.size();
        sizes[6] = ref.ndir; sizes[7] = ref.nc;
    }
    if (hier2lat) std::copy(L.hier2lat.begin(), L.hier2lat.end(), hier2lat);
    if (G) std::copy(L.G.begin(), L.G.end(), G);
    if (mass_total) *mass_total = L.mass_total;
    HOST_END
}

// refined_mesh(implicit, level): nodes dim x nf (reference coordinates, hierarchical order), elements (dim+1) x nel, 1-based
int hmg_host_refined_mesh(int dim, int nlevels, int level, double* nodes, int64_t* elems1, int64_t* nel) {
    HOST_BEGIN
    RefElement ref = build_reference(dim, nlevels);
    const RefLevel& L = get_level(ref, level);
    if (nel) *nel = (int64_t)L.cells.size() / (dim + 1);
    if (nodes)
        for (int n = 0; n < L.nf; ++n)
            for (int d = 0; d < dim; ++d) nodes[(size_t)n * dim + d] = (double)L.hier_coords[(size_t)n * 3 + d] / (double)L.m;
    if (elems1)
        for (size_t q = 0; q < L.cells.size(); ++q) elems1[q] = (int64_t)L.cells[q] + 1;
    HOST_END
}

// dense nf x nf (column-major, hierarchical order) matrix of  sum_c coef[c] * (stencil table c)
int hmg_host_local_matrix(int dim, int nlevels, int level, const double* coef, double* dense) {
    HOST_BEGIN
    RefElement ref = build_reference(dim, nlevels);
    const RefLevel& L = get_level(ref, level);
    if (dim == 3) local_matrix_t<3>(ref, L, coef, dense);
    else local_matrix_t<2>(ref, L, coef, dense);
    HOST_END
}

// dense nf(level) x nf(level-1) interpolation matrix, hierarchical order, column-major
int hmg_host_transfer_matrix(int dim, int nlevels, int level_fine, double* dense) {
    HOST_BEGIN
    RefElement ref = build_reference(dim, nlevels);
    HMG_CHECK(level_fine >= 2, "level_fine must be >= 2");
    const RefLevel& Lf = get_level(ref, level_fine);
    const RefLevel& Lc = get_level(ref, level_fine - 1);
    if (dim == 3) transfer_matrix_t<3>(Lf, Lc, dense);
    else transfer_matrix_t<2>(Lf, Lc, dense);
    HOST_END
}

// hierarchical row of the q-th paired node of a local face (kind 0) / edge (1) / vertex (2)
int hmg_host_interface_rows(int dim, int nlevels, int level, int kind, int lid, int32_t* rows, int64_t* count) {
    HOST_BEGIN
    RefElement ref = build_reference(dim, nlevels);
    const RefLevel& L = get_level(ref, level);
    std::vector<int> l2h(L.nf);
    for (int h = 0; h < L.nf; ++h) l2h[L.hier2lat[h]] = h;
    const int n = kind == 0 ? (int)L.face_bary.size() : kind == 1 ? L.m - 1 : 1;
    if (count) *count = n;
    if (rows)
        for (int t = 0; t < n; ++t) {
            const unsigned ab = kind == 0 ? L.face_bary[t] : 0;
            const int q = kind == 1 ? t + 1 : 0;
            const int p = dim == 3 ? interface_node<3>(L.m, kind, lid, q, ab) : interface_node<2>(L.m, kind, lid, q, ab);
            rows[t] = l2h[p];
        }
    HOST_END
}

// kind 0 faces, 1 edges, 2 interface vertices, 3 all nodes.  Two-pass: arrays may be NULL.
int hmg_host_topology(int dim, int64_t ne, int64_t nn, const int64_t* elems1, int kind, int64_t* ncells,
                      int64_t* nentries, int64_t* offset, int64_t* element, int64_t* local_id) {
    HOST_BEGIN
    std::vector<int64_t> el((size_t)ne * (dim + 1));
    for (size_t q = 0; q < el.size(); ++q) el[q] = elems1[q] - 1;
    Topology T = build_topology(dim, ne, nn, el.data());
    const std::vector<int64_t>* off;
    const std::vector<int32_t>* own;
    if (kind == 3) { off = &T.nodeown_off; own = &T.nodeown; }
    else {
        const CellMap& m = kind == 0 ? T.faces : kind == 1 ? T.edges : T.verts;
        off = &m.offset; own = &m.owner;
    }
    if (ncells) *ncells = off->empty() ? 0 : (int64_t)off->size() - 1;
    if (nentries) *nentries = (int64_t)own->size();
    if (offset) std::copy(off->begin(), off->end(), offset);
    if (element) for (size_t q = 0; q < own->size(); ++q) element[q] = (*own)[q] >> 3;
    if (local_id) for (size_t q = 0; q < own->size(); ++q) local_id[q] = (*own)[q] & 7;
    HOST_END
}

// cmask[ne] (bit c set <=> node class c of that element lies on the domain boundary),
// interior[nn] flags (1 = interior node)
int hmg_host_boundary(int dim, int64_t ne, int64_t nn, const int64_t* elems1, uint16_t* cmask, uint8_t* interior) {
    HOST_BEGIN
    std::vector<int64_t> el((size_t)ne * (dim + 1));
    for (size_t q = 0; q < el.size(); ++q) el[q] = elems1[q] - 1;
    Topology T = build_topology(dim, ne, nn, el.data());
    if (cmask) std::copy(T.cmask.begin(), T.cmask.end(), cmask);
    if (interior) for (int64_t n = 0; n < nn; ++n) interior[n] = T.node_boundary[n] ? 0 : 1;
    HOST_END
}

// ---- partition of the coarse elements over ranks (multi-GPU host logic, no GPU needed) ----
namespace {
Partition host_partition(int dim, int64_t ne, int64_t nn, const int64_t* elems1, const int32_t* owner_rank, int rank, int nranks) {
    std::vector<int64_t> el((size_t)ne * (dim + 1));
    for (size_t q = 0; q < el.size(); ++q) el[q] = elems1[q] - 1;
    Topology T = build_topology(dim, ne, nn, el.data());
    return build_partition(T, owner_rank, rank, nranks);
}
}  // namespace

// local elements of `rank`: local_to_global[ne_local] (0-based), per local element the Dirichlet class mask and
// the owner counts mult[ne_local][16], per base node the first local owner (element*8+local id, -1 none) and
// whether this rank reports the node to the coarse solve.  Arrays may be NULL.
int hmg_host_partition_elements(int dim, int64_t ne, int64_t nn, const int64_t* elems1, const int32_t* owner_rank, int rank,
                                int nranks, int64_t* ne_local, int64_t* local_to_global, uint16_t* cmask, uint8_t* mult,
                                int32_t* node_first, uint8_t* node_contrib) {
    HOST_BEGIN
    const Partition P = host_partition(dim, ne, nn, elems1, owner_rank, rank, nranks);
    if (ne_local) *ne_local = (int64_t)P.local_to_global.size();
    if (local_to_global) std::copy(P.local_to_global.begin(), P.local_to_global.end(), local_to_global);
    if (cmask) std::copy(P.cmask.begin(), P.cmask.end(), cmask);
    if (mult) std::copy(P.mult.begin(), P.mult.end(), mult);
    if (node_first) std::copy(P.node_first.begin(), P.node_first.end(), node_first);
    if (node_contrib) std::copy(P.node_contrib.begin(), P.node_contrib.end(), node_contrib);
    HOST_END
}
// interface cells of one kind (0 faces, 1 edges, 2 vertices) as seen by `rank`.
//   cut == 0: cells whose owners are all local: sizes[2] = cells, entries; CSR offset / element (LOCAL index) / local_id
//   cut == 1: cut cells this rank takes part in: sizes[3] = cells, entries, cut cells of the kind on ALL ranks;
//             slot[cells] = ordinal in the global cut enumeration, first_local[cells]
int hmg_host_partition_cells(int dim, int64_t ne, int64_t nn, const int64_t* elems1, const int32_t* owner_rank, int rank,
                             int nranks, int kind, int cut, int64_t* sizes, int64_t* offset, int64_t* element, int64_t* local_id,
                             int64_t* slot, uint8_t* first_local) {
    HOST_BEGIN
    HMG_CHECK(kind >= 0 && kind < 3, "kind must be 0, 1 or 2");
    const Partition P = host_partition(dim, ne, nn, elems1, owner_rank, rank, nranks);
    const std::vector<int64_t>* off;
    const std::vector<int32_t>* own;
    if (cut) {
        const CutCells& C = P.cut[kind];
        off = &C.offset; own = &C.owner;
        if (sizes) { sizes[0] = C.ncells(); sizes[1] = (int64_t)C.owner.size(); sizes[2] = C.nglobal; }
        if (slot) std::copy(C.slot.begin(), C.slot.end(), slot);
        if (first_local) std::copy(C.first_local.begin(), C.first_local.end(), first_local);
    } else {
        const CellMap& m = kind == 0 ? P.faces : kind == 1 ? P.edges : P.verts;
        off = &m.offset; own = &m.owner;
        if (sizes) { sizes[0] = m.ncells(); sizes[1] = (int64_t)m.owner.size(); }
    }
    if (offset) std::copy(off->begin(), off->end(), offset);
    if (element) for (size_t q = 0; q < own->size(); ++q) element[q] = (*own)[q] >> 3;
    if (local_id) for (size_t q = 0; q < own->size(); ++q) local_id[q] = (*own)[q] & 7;
    HOST_END
}

// neighbour exchange of the cut cells of one kind: per participating cell (same order as hmg_host_partition_cells
// with cut == 1) the other ranks that share it, ascending, the ordinal of the cell in the message exchanged with
// each of them, and the position of the own partial sum in the rank-ordered total; shared_with[nranks * 3] =
// cut cells of every kind shared with every rank.  sizes[2] = cells, peer entries.  Arrays may be NULL.
int hmg_host_partition_peers(int dim, int64_t ne, int64_t nn, const int64_t* elems1, const int32_t* owner_rank, int rank,
                             int nranks, int kind, int64_t* sizes, int64_t* peer_off, int32_t* peer_rank, int32_t* peer_idx,
                             int32_t* my_pos, int64_t* shared_with) {
    HOST_BEGIN
    HMG_CHECK(kind >= 0 && kind < 3, "kind must be 0, 1 or 2");
    const Partition P = host_partition(dim, ne, nn, elems1, owner_rank, rank, nranks);
    const CutCells& C = P.cut[kind];
    if (sizes) { sizes[0] = C.ncells(); sizes[1] = (int64_t)C.peer_rank.size(); }
    if (peer_off) std::copy(C.peer_off.begin(), C.peer_off.end(), peer_off);
    if (peer_rank) std::copy(C.peer_rank.begin(), C.peer_rank.end(), peer_rank);
    if (peer_idx) std::copy(C.peer_idx.begin(), C.peer_idx.end(), peer_idx);
    if (my_pos) std::copy(C.my_pos.begin(), C.my_pos.end(), my_pos);
    if (shared_with) std::copy(P.shared_with.begin(), P.shared_with.end(), shared_with);
    HOST_END
}

// class bitmask of a local face / edge / vertex
int hmg_host_class_of(int dim, int kind, int lid) {
    return kind == 0 ? class_of_face(lid) : kind == 1 ? class_of_edge(dim, lid) : class_of_vertex(dim, lid);
}

// coef[ne][stride] = |J| P (symmetric components) , |J|
int hmg_host_element_coefficients(int dim, int64_t ne, int64_t nn, const double* nodes, const int64_t* elems1,
                                  const double* sigma, double* coef, int stride) {
    HOST_BEGIN
    (void)nn;
    std::vector<int64_t> el((size_t)ne * (dim + 1));
    for (size_t q = 0; q < el.size(); ++q) el[q] = elems1[q] - 1;
    std::vector<double> c;
    element_coefficients(dim, ne, nodes, el.data(), sigma, c, stride);
    std::copy(c.begin(), c.end(), coef);
    HOST_END
}

// y = A x for one element through the apply kernel's sweeps (lattice order); 2D lines are split into
// segments of 2^seg_shift nodes; info[4] = tasks, largest row window of a task, ring rows, smem bytes
int hmg_host_apply_sweep(int dim, int nlevels, int level, int seg_shift, const double* coef, const double* x, double* y,
                         int64_t* info) {
    HOST_BEGIN
    RefElement ref = build_reference(dim, nlevels);
    const RefLevel& L = get_level(ref, level);
    HMG_CHECK(seg_shift >= 1 && seg_shift <= 8, "bad segment size");
    if (dim == 3) emulate_sweep_t<3>(L, seg_shift, coef, x, y, info);
    else emulate_sweep_t<2>(L, seg_shift, coef, x, y, info);
    HOST_END
}

}  // extern "C"
