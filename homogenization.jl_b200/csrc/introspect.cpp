// Host-only verification entry points (no GPU needed, nothing here computes on behalf of the
// product path).  They expose the tables the kernels consume, expanded through the SAME index
// functions the kernels use (lattice.hpp), so that the CPU test-suite can compare them with the
// oracle's explicit sparse operators and interface maps.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hmg.h"
#include "hmg_host.hpp"
#include "lattice.hpp"

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

// Executes the apply plan of one level for ONE element on the host, task by task and slot by slot,
// with the same reference decoding, tap directions and face weights as apply_kernel (kernels.cu),
// all chunks resident.  y (lattice order) = A x.  Verifies the plan tables without a GPU; the
// device code itself is checked by the -m gpu parity tests.
template <int DIM>
void emulate_plan_t(const RefLevel& L, const ApplyPlan& P, const double* coef, const double* x, double* y, int64_t* info) {
    using D = Dims<DIM>;
    constexpr int NP = DIM == 3 ? 3 : 1;
    const int KP = DIM == 3 ? 5 : 3, KM = DIM == 3 ? 6 : 4;
    const int M0[3] = {1, DIM == 3 ? 3 : 0, 7}, M1[3] = {DIM == 3 ? 10 : 6, 12, 14};
    const int P0[3] = {2, 4, 8}, P1[3] = {DIM == 3 ? 9 : 5, 11, 13};
    const int FIRST = 1, LAST = DIM == 3 ? 8 : 4, FACE_A = 2, FACE_B = 4;
    // image: zero line, then the chunks back to back with two nodes of slack each, poisoned slack
    const double poison = std::nan("");
    const int Z = P.zero_nodes;
    std::vector<int> cbase(P.nchunks);
    int total = Z;
    for (int c = 0; c < P.nchunks; ++c) { cbase[c] = total; total += P.chunk_start[c + 1] - P.chunk_start[c] + 2; }
    std::vector<double> sm(total, poison);
    for (int q = 0; q < Z; ++q) sm[q] = 0.0;
    for (int c = 0; c < P.nchunks; ++c)
        for (int p = P.chunk_start[c]; p < P.chunk_start[c + 1]; ++p) sm[cbase[c] + p - P.chunk_start[c]] = x[p];
    auto coefI = [&](int cls, int d) {
        double s = 0.0;
        for (int q = 0; q < D::NC; ++q) s = std::fma(coef[q], L.G[((size_t)cls * D::NDIR + d) * D::NC + q], s);
        return s;
    };
    std::vector<int> written(L.nf, 0);
    int64_t nsweep_nodes = 0, nspecial = 0;
    for (int t = 0; t < P.ntasks; ++t) {
        const int32_t* T = &P.tasks[(size_t)t * PLAN_TASK_INTS];
        const int type = T[0], clo = T[1], chi = T[2];
        HMG_CHECK(clo >= 0 && chi < P.nchunks && chi - clo <= 2, "bad chunk window");
        auto off = [&](uint32_t ref) -> int {
            const unsigned sel = ref >> 28;
            if (sel == 3) return (int)(ref & 0x0fffffff);
            HMG_CHECK((int)(clo + sel) <= chi, "reference beyond the task's chunk window");
            return cbase[clo + sel] + (int)(ref & 0x0fffffff);
        };
        for (int s = 0; s < P.spw; ++s) {
            const int32_t* d = T + 4 + s * PLAN_SLOT_INTS;
            if (type == TASK_NODES) {
                if (d[0] < 0) continue;
                const uint32_t* e = &P.nodetab[(size_t)d[0] * 16];
                const int p = e[15] & 0xffff, cls = e[15] >> 16;
                double acc = 0.0;
                for (int dd = 0; dd < D::NDIR; ++dd)
                    if (e[dd] != 0xFFFFFFFFu) acc = std::fma(coefI(cls, dd), sm[off(e[dd])], acc);
                y[p] = acc;
                written[p]++;
                ++nspecial;
                continue;
            }
            const int rcls = type == TASK_SWEEP_INTERIOR ? 0 : (type == TASK_SWEEP_FACE_A ? FACE_A : FACE_B);
            const int cnt = d[7];
            if (cnt == 0) continue;
            const int oc = off((uint32_t)d[0]);
            int om[3], op[3];
            for (int q = 0; q < NP; ++q) { om[q] = off((uint32_t)d[1 + q]); op[q] = off((uint32_t)d[4 + q]); }
            for (int k = 0; k < cnt; ++k) {
                int cls = rcls;
                if (rcls == 0 && (d[9] & 1) && k == 0) cls = FIRST;
                else if (rcls == 0 && (d[9] & 2) && k == cnt - 1) cls = LAST;
                const int p = d[8] + k;
                HMG_CHECK((int)(L.nodeinfo[p] >> 24) == cls, "sweep visits a node of another class");
                double a1 = 0.0, ah = 0.0;
                auto term = [&](int dir, double c, int o) {
                    const double w = face_weight(DIM, cls, dir);
                    if (w == 1.0) a1 = std::fma(c, sm[o], a1);
                    else if (w == 0.5) ah = std::fma(c, sm[o], ah);
                };
                term(0, coefI(0, 0), oc + k);
                term(KP, coefI(0, KP), oc + k + 1);
                term(KM, coefI(0, KP), oc + k - 1);
                for (int q = 0; q < NP; ++q) {
                    term(M0[q], coefI(0, M0[q]), om[q] + k);
                    term(M1[q], coefI(0, M1[q]), om[q] + k - 1);
                    term(P0[q], coefI(0, M0[q]), op[q] + k);
                    term(P1[q], coefI(0, M1[q]), op[q] + k + 1);
                }
                y[p] = a1 + 0.5 * ah;
                written[p]++;
                ++nsweep_nodes;
            }
        }
    }
    for (int p = 0; p < L.nf; ++p) HMG_CHECK(written[p] == 1, "plan does not write every node exactly once");
    if (info) {
        info[0] = P.nchunks; info[1] = P.nslots; info[2] = P.slot_nodes; info[3] = P.ntasks; info[4] = P.nwarps;
        info[5] = P.ctas_per_sm; info[6] = (int64_t)P.smem_bytes; info[7] = nsweep_nodes; info[8] = nspecial;
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

// y = A x for one element through the apply plan (lattice order), info[9] = nchunks, nslots, slot_nodes,
// ntasks, nwarps, ctas_per_sm, smem_bytes, nodes computed by sweeps, nodes computed by the generic path
int hmg_host_apply_plan(int dim, int nlevels, int level, int W, const double* coef, const double* x, double* y,
                        int64_t* info) {
    HOST_BEGIN
    RefElement ref = build_reference(dim, nlevels);
    const RefLevel& L = get_level(ref, level);
    const ApplyPlan P = build_apply_plan(dim, L, W);
    if (dim == 3) emulate_plan_t<3>(L, P, coef, x, y, info);
    else emulate_plan_t<2>(L, P, coef, x, y, info);
    HOST_END
}

}  // extern "C"
