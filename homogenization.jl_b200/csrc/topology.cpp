// Base-mesh topology: interface cells with their owners, domain-boundary classes, all-nodes map,
// per-element geometry.  Built once on the host (O(Ne log Ne)), uploaded by the context.
//
// Semantics follow src/interface.jl:65-117 (interfaces: cells shared by >= 2 elements, owners in
// ascending element index), src/interface.jl:207-284 (boundary faces -> their edges -> their nodes,
// each with ALL owners), src/grid.jl:176-202 (interior nodes) -- implemented with comparison sorts
// on packed keys instead of the reference's radix-sort pipeline.
#include <algorithm>
#include <array>
#include <cmath>

#include "hmg_host.hpp"

namespace hmg {

int class_of_face(int lf) { return 1 << lf; }
int class_of_vertex(int dim, int lv) { return dim == 3 ? (15 & ~(8 >> lv)) : (7 & ~(4 >> lv)); }
int class_of_edge(int dim, int le) {
    if (dim == 2) return 1 << le;
    static const int E3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    int mask = 0;
    for (int v = 0; v < 4; ++v)
        if (v != E3[le][0] && v != E3[le][1]) mask |= 8 >> v;
    return mask;
}

namespace {

struct Entry {
    std::array<int64_t, 3> key;
    int32_t elem;
    int8_t lid;
    bool operator<(const Entry& o) const {
        if (key != o.key) return key < o.key;
        return elem < o.elem;    // ascending element index inside a cell (stable radix sort in the reference)
    }
};

template <class F>   // F(group_begin, group_end)
void for_groups(const std::vector<Entry>& v, F f) {
    size_t i = 0;
    while (i < v.size()) {
        size_t j = i + 1;
        while (j < v.size() && v[j].key == v[i].key) ++j;
        f(i, j);
        i = j;
    }
}

CellMap interface_map(const std::vector<Entry>& v) {
    CellMap m;
    m.offset.push_back(0);
    for_groups(v, [&](size_t i, size_t j) {
        if (j - i < 2) return;                       // remove_singletons!
        for (size_t q = i; q < j; ++q) m.owner.push_back(v[q].elem * 8 + v[q].lid);
        m.cell_key.push_back(v[i].key[0]);
        m.offset.push_back((int64_t)m.owner.size());
    });
    return m;
}

}  // namespace

Topology build_topology(int dim, int64_t ne, int64_t nn, const int64_t* elems) {
    HMG_CHECK(ne < (int64_t(1) << 27), "too many coarse elements for the packed owner ids");
    Topology T;
    T.dim = dim;
    T.ne = ne;
    T.nn = nn;
    const int nv = dim + 1;
    for (int64_t e = 0; e < ne; ++e)
        for (int a = 0; a < nv; ++a) {
            int64_t v = elems[e * nv + a];
            HMG_CHECK(v >= 0 && v < nn, "element refers to a node outside the mesh");
            if (a) HMG_CHECK(elems[e * nv + a - 1] < v, "base elements must be sorted ascending (src/implicit_fine_grid.jl:14)");
        }
    static const int F3[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
    static const int E3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    static const int E2[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    const int nedge = dim == 3 ? 6 : 3;

    std::vector<Entry> verts, edges, faces;
    verts.reserve(ne * nv);
    edges.reserve(ne * nedge);
    for (int64_t e = 0; e < ne; ++e) {
        const int64_t* el = elems + e * nv;
        for (int a = 0; a < nv; ++a) verts.push_back({{el[a], -1, -1}, (int32_t)e, (int8_t)a});
        for (int q = 0; q < nedge; ++q) {
            const int* pr = dim == 3 ? E3[q] : E2[q];
            edges.push_back({{el[pr[0]], el[pr[1]], -1}, (int32_t)e, (int8_t)q});
        }
        if (dim == 3)
            for (int q = 0; q < 4; ++q)
                faces.push_back({{el[F3[q][0]], el[F3[q][1]], el[F3[q][2]]}, (int32_t)e, (int8_t)q});
    }
    std::sort(verts.begin(), verts.end());
    std::sort(edges.begin(), edges.end());
    std::sort(faces.begin(), faces.end());

    T.verts = interface_map(verts);
    T.edges = interface_map(edges);
    T.faces = interface_map(faces);

    // unified cell numbering + per-element cell table (used by the fused apply / interface kernel)
    {
        T.cell_off.push_back(0);
        T.elem_cells.assign((size_t)ne * 16, -1);
        const CellMap* maps[3] = {&T.faces, &T.edges, &T.verts};
        const int slot0[3] = {0, dim == 3 ? 4 : 0, dim == 3 ? 10 : 3};
        for (int kind = dim == 3 ? 0 : 1; kind < 3; ++kind) {
            const CellMap& m = *maps[kind];
            for (int64_t c = 0; c < m.ncells(); ++c) {
                const int32_t id = (int32_t)(T.cell_off.size() - 1);
                for (int64_t o = m.offset[c]; o < m.offset[c + 1]; ++o) {
                    T.cell_own.push_back(m.owner[o]);
                    T.elem_cells[(size_t)(m.owner[o] >> 3) * 16 + slot0[kind] + (m.owner[o] & 7)] = id;
                }
                T.cell_off.push_back((int64_t)T.cell_own.size());
            }
        }
    }

    // all_nodes map
    T.node_first.assign(nn, -1);
    T.nodeown_off.assign(nn + 1, 0);
    T.nodeown.reserve(verts.size());
    for (const auto& v : verts) T.nodeown_off[v.key[0] + 1]++;
    for (int64_t n = 0; n < nn; ++n) T.nodeown_off[n + 1] += T.nodeown_off[n];
    for (const auto& v : verts) {
        if (T.node_first[v.key[0]] < 0) T.node_first[v.key[0]] = v.elem * 8 + v.lid;
        T.nodeown.push_back(v.elem * 8 + v.lid);
    }

    // domain boundary
    T.cmask.assign(ne, 0);
    T.node_boundary.assign(nn, 0);
    std::vector<std::array<int64_t, 2>> bedges;
    if (dim == 3) {
        for_groups(faces, [&](size_t i, size_t j) {
            if ((j - i) % 2 == 0) return;            // remove_repeated_pairs!: a lone face is a boundary face
            const Entry& f = faces[j - 1];
            T.cmask[f.elem] |= (uint16_t)(1u << class_of_face(f.lid));
            bedges.push_back({f.key[0], f.key[1]});
            bedges.push_back({f.key[0], f.key[2]});
            bedges.push_back({f.key[1], f.key[2]});
        });
        std::sort(bedges.begin(), bedges.end());
        bedges.erase(std::unique(bedges.begin(), bedges.end()), bedges.end());
        for_groups(edges, [&](size_t i, size_t j) {
            std::array<int64_t, 2> k = {edges[i].key[0], edges[i].key[1]};
            if (!std::binary_search(bedges.begin(), bedges.end(), k)) return;
            for (size_t q = i; q < j; ++q)
                T.cmask[edges[q].elem] |= (uint16_t)(1u << class_of_edge(3, edges[q].lid));
        });
    } else {
        for_groups(edges, [&](size_t i, size_t j) {
            if ((j - i) % 2 == 0) return;
            const Entry& ed = edges[j - 1];
            T.cmask[ed.elem] |= (uint16_t)(1u << class_of_edge(2, ed.lid));
            bedges.push_back({ed.key[0], ed.key[1]});
        });
    }
    for (const auto& be : bedges) T.node_boundary[be[0]] = T.node_boundary[be[1]] = 1;
    for (const auto& v : verts)
        if (T.node_boundary[v.key[0]]) T.cmask[v.elem] |= (uint16_t)(1u << class_of_vertex(dim, v.lid));
    for (int64_t n = 0; n < nn; ++n)
        if (!T.node_boundary[n]) T.interior_nodes.push_back(n);
    return T;
}

Partition build_partition(const Topology& T, const int32_t* owner_rank, int rank, int nranks) {
    HMG_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / number of ranks");
    Partition P;
    P.rank = rank;
    P.nranks = nranks;
    const int dim = T.dim;
    const int64_t ne = T.ne;
    P.global_to_local.assign(ne, -1);
    for (int64_t e = 0; e < ne; ++e) {
        const int r = owner_rank ? owner_rank[e] : 0;
        HMG_CHECK(r >= 0 && r < nranks, "owner_rank out of range");
        if (r == rank) {
            P.global_to_local[e] = (int32_t)P.local_to_global.size();
            P.local_to_global.push_back(e);
        }
    }
    const int64_t nel = (int64_t)P.local_to_global.size();
    auto rank_of = [&](int32_t id) { return owner_rank ? owner_rank[id >> 3] : 0; };
    const CellMap* maps[3] = {&T.faces, &T.edges, &T.verts};
    CellMap* locals[3] = {&P.faces, &P.edges, &P.verts};
    P.shared_with.assign((size_t)nranks * 3, 0);
    std::vector<int> cell_ranks;
    for (int kind = 0; kind < 3; ++kind) {
        const CellMap& m = *maps[kind];
        CellMap& loc = *locals[kind];
        CutCells& cut = P.cut[kind];
        loc.offset.assign(1, 0);
        cut.offset.assign(1, 0);
        cut.peer_off.assign(1, 0);
        for (int64_t c = 0; c < m.ncells(); ++c) {
            const int64_t b = m.offset[c], en = m.offset[c + 1];
            int nlocal = 0;
            bool is_cut = false;
            for (int64_t o = b; o < en; ++o) {
                if (rank_of(m.owner[o]) == rank) ++nlocal;
                if (rank_of(m.owner[o]) != rank_of(m.owner[b])) is_cut = true;
            }
            if (!is_cut) {
                if (nlocal == 0) continue;
                for (int64_t o = b; o < en; ++o)
                    loc.owner.push_back(P.global_to_local[m.owner[o] >> 3] * 8 + (m.owner[o] & 7));
                loc.cell_key.push_back(m.cell_key[c]);
                loc.offset.push_back((int64_t)loc.owner.size());
                continue;
            }
            const int64_t slot = cut.nglobal++;
            if (nlocal == 0) continue;
            cut.slot.push_back(slot);
            cut.first_local.push_back(rank_of(m.owner[b]) == rank ? 1 : 0);
            for (int64_t o = b; o < en; ++o)
                if (rank_of(m.owner[o]) == rank)
                    cut.owner.push_back(P.global_to_local[m.owner[o] >> 3] * 8 + (m.owner[o] & 7));
            cut.offset.push_back((int64_t)cut.owner.size());
            // the other ranks that share the cell
            cell_ranks.clear();
            for (int64_t o = b; o < en; ++o) cell_ranks.push_back(rank_of(m.owner[o]));
            std::sort(cell_ranks.begin(), cell_ranks.end());
            cell_ranks.erase(std::unique(cell_ranks.begin(), cell_ranks.end()), cell_ranks.end());
            int below = 0;
            for (int q : cell_ranks) {
                if (q == rank) continue;
                if (q < rank) ++below;
                cut.peer_rank.push_back(q);
                cut.peer_idx.push_back((int32_t)P.shared_with[(size_t)q * 3 + kind]++);
            }
            cut.my_pos.push_back(below);
            cut.peer_off.push_back((int64_t)cut.peer_rank.size());
        }
    }
    // Dirichlet classes and owner counts of the local elements (from the GLOBAL topology)
    P.cmask.resize(nel);
    P.mult.assign((size_t)nel * 16, 1);
    const int nv = dim + 1, nfl = dim == 3 ? 4 : 0, ned = dim == 3 ? 6 : 3;
    for (int64_t l = 0; l < nel; ++l) {
        const int64_t e = P.local_to_global[l];
        P.cmask[l] = T.cmask[e];
        auto owners = [&](int slot) -> int {
            const int32_t id = T.elem_cells[(size_t)e * 16 + slot];
            if (id < 0) return 1;
            const int64_t n = T.cell_off[id + 1] - T.cell_off[id];
            HMG_CHECK(n <= 255, "a base-mesh cell has more than 255 owners");
            return (int)n;
        };
        uint8_t* dst = &P.mult[(size_t)l * 16];
        for (int q = 0; q < nfl; ++q) dst[class_of_face(q)] = (uint8_t)owners(q);
        for (int q = 0; q < ned; ++q) dst[class_of_edge(dim, q)] = (uint8_t)owners(nfl + q);
        for (int q = 0; q < nv; ++q) dst[class_of_vertex(dim, q)] = (uint8_t)owners(nfl + ned + q);
    }
    // base nodes: first local owner; the lowest rank that owns a copy reports the node to the coarse solve
    P.node_first.assign(T.nn, -1);
    P.node_contrib.assign(T.nn, 0);
    for (int64_t n = 0; n < T.nn; ++n) {
        int lowest = nranks;
        for (int64_t o = T.nodeown_off[n]; o < T.nodeown_off[n + 1]; ++o) {
            const int32_t id = T.nodeown[o];
            const int r = rank_of(id);
            lowest = std::min(lowest, r);
            if (r == rank && P.node_first[n] < 0) P.node_first[n] = P.global_to_local[id >> 3] * 8 + (id & 7);
        }
        P.node_contrib[n] = lowest == rank ? 1 : 0;
    }
    return P;
}

void element_flux_vectors(int dim, int64_t ne, const double* nodes, const int64_t* elems, const double* sigma,
                          const double* xi, std::vector<double>& flux) {
    const int nv = dim + 1;
    flux.assign((size_t)ne * dim, 0.0);
    for (int64_t e = 0; e < ne; ++e) {
        const int64_t* el = elems + e * nv;
        double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int c = 0; c < dim; ++c)
            for (int r = 0; r < dim; ++r) J[r][c] = nodes[el[c + 1] * dim + r] - nodes[el[0] * dim + r];
        double s[3] = {0, 0, 0}, y[3] = {0, 0, 0}, det;
        for (int d = 0; d < dim; ++d) s[d] = sigma[e * dim + d] * xi[d];
        // y = J^-1 s by Cramer's rule (adjugate / det)
        if (dim == 2) {
            det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            y[0] = (J[1][1] * s[0] - J[0][1] * s[1]) / det;
            y[1] = (-J[1][0] * s[0] + J[0][0] * s[1]) / det;
        } else {
            const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                         c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
            det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
            const double a[3][3] = {{c00, J[0][2] * J[2][1] - J[0][1] * J[2][2], J[0][1] * J[1][2] - J[0][2] * J[1][1]},
                                    {c01, J[0][0] * J[2][2] - J[0][2] * J[2][0], J[0][2] * J[1][0] - J[0][0] * J[1][2]},
                                    {c02, J[0][1] * J[2][0] - J[0][0] * J[2][1], J[0][0] * J[1][1] - J[0][1] * J[1][0]}};
            for (int k = 0; k < 3; ++k) y[k] = (a[k][0] * s[0] + a[k][1] * s[1] + a[k][2] * s[2]) / det;
        }
        HMG_CHECK(det != 0.0 && std::isfinite(det), "degenerate base element");
        for (int k = 0; k < dim; ++k) flux[(size_t)e * dim + k] = -std::fabs(det) * y[k];
    }
}

void element_coefficients(int dim, int64_t ne, const double* nodes, const int64_t* elems,
                          const double* sigma, std::vector<double>& coef, int stride) {
    const int nv = dim + 1;
    coef.assign((size_t)ne * stride, 0.0);
    for (int64_t e = 0; e < ne; ++e) {
        const int64_t* el = elems + e * nv;
        double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int c = 0; c < dim; ++c)
            for (int r = 0; r < dim; ++r) J[r][c] = nodes[el[c + 1] * dim + r] - nodes[el[0] * dim + r];
        double det, inv[3][3];   // inv = J^-1
        if (dim == 2) {
            det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            inv[0][0] = J[1][1] / det; inv[0][1] = -J[0][1] / det;
            inv[1][0] = -J[1][0] / det; inv[1][1] = J[0][0] / det;
        } else {
            double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
            double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
            double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
            det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
            inv[0][0] = c00 / det;
            inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
            inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
            inv[1][0] = c01 / det;
            inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
            inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
            inv[2][0] = c02 / det;
            inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
            inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
        }
        HMG_CHECK(det != 0.0 && std::isfinite(det), "degenerate base element");
        const double adet = std::fabs(det);
        // P = J^-1 diag(sigma) J^-T :  P[k][l] = sum_r inv[k][r] sigma_r inv[l][r]
        double* out = &coef[(size_t)e * stride];
        int c = 0;
        for (int k = 0; k < dim; ++k)
            for (int l = k; l < dim; ++l, ++c) {
                double s = 0.0;
                for (int r = 0; r < dim; ++r) s += inv[k][r] * sigma[e * dim + r] * inv[l][r];
                out[c] = adet * s;
            }
        out[c] = adet;
    }
}

}  // namespace hmg
