// Streaming plan of the apply kernel: chunks (TMA units), ring geometry and the warp task list
// of one level of the refined reference simplex.  Pure host code, integer-exact, verified for
// coverage (every node is computed by exactly one task) before it is uploaded.
//
// Why it looks like this: the reference applies dim^2+1 CSC scatter-SpMVs per coarse element
// (src/apply_local_operators.jl:93-133).  Here the operator is a lattice stencil and the device
// layout interleaves W coarse elements, so a warp lane is an ELEMENT: all lanes of a row slot
// walk the same lattice line (fastest coordinate) with a register sliding window.  Nodes in the
// interior of the simplex or of one of its faces use coefficients derived from the interior
// stencil (face_weight); the few nodes on reference edges / vertices go through a table-driven
// generic path.
#include <algorithm>
#include <array>
#include <cstdio>

#include "hmg_host.hpp"
#include "lattice.hpp"

namespace hmg {

namespace {

bool dir_outside(int dim, int cls, int d) {
    const int* v = dim == 3 ? DIRS3[d] : DIRS2[d];
    if (dim == 3) {
        const int mask = (v[2] < 0 ? 1 : 0) | (v[1] < 0 ? 2 : 0) | (v[0] < 0 ? 4 : 0) | (v[0] + v[1] + v[2] > 0 ? 8 : 0);
        return (cls & mask) != 0;
    }
    const int mask = (v[1] < 0 ? 1 : 0) | (v[0] < 0 ? 2 : 0) | (v[0] + v[1] > 0 ? 4 : 0);
    return (cls & mask) != 0;
}
int opposite(int d) { return d == 0 ? 0 : (d & 1 ? d + 1 : d - 1); }

struct Builder {
    int dim, m, nf, W, spw;
    std::vector<int> pstart;           // packed index of the first node of plane i (dim 3) / row i (dim 2)
    std::vector<int> chunk_of_plane;
    std::vector<int32_t> chunk_start;

    int pack(int i, int j, int k) const { return dim == 3 ? lat_pack3(m, i, j, k) : lat_pack2(m, i, j); }
    int plane_of_packed(int p) const {
        return (int)(std::upper_bound(pstart.begin(), pstart.end(), p) - pstart.begin()) - 1;
    }
    // absolute reference: chunk << 28 | offset (made relative to the task's first chunk later)
    struct Ref { int chunk; int off; bool zero; };
    Ref ref_node(int plane, int p_line_start, int kstart) const {
        if (plane < 0 || plane > m) return {0, 1, true};
        const int c = chunk_of_plane[plane];
        return {c, p_line_start + kstart - chunk_start[c], false};
    }
};

struct SlotSweep {
    Builder::Ref c, mi[3], pl[3];
    int cnt, pout, flags, plane;
};

}  // namespace

double face_weight(int dim, int cls, int d) {
    if (cls == 0) return 1.0;
    if (dir_outside(dim, cls, d)) return 0.0;
    if (d == 0) return 0.5;
    return dir_outside(dim, cls, opposite(d)) ? 1.0 : 0.5;
}

ApplyPlan build_apply_plan(int dim, const RefLevel& L, int W) {
    HMG_CHECK(W == 8 || W == 16 || W == 32, "group width must be 8, 16 or 32");
    ApplyPlan P;
    P.W = W;
    P.spw = 32 / W;
    Builder B;
    B.dim = dim; B.m = L.m; B.nf = L.nf; B.W = W; B.spw = P.spw;
    const int m = L.m;
    for (int i = 0; i <= m; ++i) B.pstart.push_back(B.pack(i, 0, 0));
    B.pstart.push_back(L.nf);

    // ---- chunks: whole planes, merged while smaller than ~4 KB ----
    const int min_nodes = std::max(8, 4096 / (W * 8));
    B.chunk_of_plane.assign(m + 1, 0);
    B.chunk_start.push_back(0);
    {
        int cur = 0;
        for (int i = 0; i <= m; ++i) {
            B.chunk_of_plane[i] = (int)B.chunk_start.size() - 1;
            cur += B.pstart[i + 1] - B.pstart[i];
            if (cur >= min_nodes && i < m) {
                B.chunk_start.push_back(B.pstart[i + 1]);
                cur = 0;
            }
        }
        B.chunk_start.push_back(L.nf);
        // a tiny trailing chunk joins its predecessor
        const int nc = (int)B.chunk_start.size() - 1;
        if (nc >= 2 && B.chunk_start[nc] - B.chunk_start[nc - 1] < min_nodes) {
            B.chunk_start.erase(B.chunk_start.end() - 2);
            for (int i = 0; i <= m; ++i) B.chunk_of_plane[i] = std::min(B.chunk_of_plane[i], nc - 2);
        }
    }
    P.chunk_start = B.chunk_start;
    P.nchunks = (int)B.chunk_start.size() - 1;
    int maxchunk = 0;
    for (int c = 0; c < P.nchunks; ++c) maxchunk = std::max(maxchunk, B.chunk_start[c + 1] - B.chunk_start[c]);
    P.slot_nodes = maxchunk + 2;
    P.zero_nodes = m + 4;
    {
        const size_t S = (size_t)P.slot_nodes * W * 8, Z = (size_t)P.zero_nodes * W * 8;
        const size_t two = 112 * 1024, one = 224 * 1024;
        if (Z + 5 * S <= two) {
            P.ctas_per_sm = 2;
            P.nslots = (int)std::min<size_t>(6, (two - Z) / S);
        } else {
            P.ctas_per_sm = 1;
            P.nslots = (int)std::min<size_t>(8, (one - Z) / S);
        }
        HMG_CHECK(P.nslots >= 4, "level too large for the shared-memory ring of the apply kernel");
        P.smem_bytes = Z + (size_t)P.nslots * S;
    }
    P.nwarps = 8;

    // ---- tasks ----
    std::vector<std::array<int32_t, PLAN_TASK_INTS>> tasks;
    std::vector<char> covered(L.nf, 0);
    auto cover = [&](int p) {
        HMG_CHECK(p >= 0 && p < L.nf && !covered[p], "apply plan covers a node twice");
        covered[p] = 1;
    };
    auto enc = [&](const Builder::Ref& r, int clo) -> int32_t {
        if (r.zero) return (int32_t)((3u << 28) | 1u);
        HMG_CHECK(r.chunk >= clo && r.chunk - clo <= 2 && r.off >= 0 && r.off < (1 << 28), "apply plan reference out of range");
        return (int32_t)(((uint32_t)(r.chunk - clo) << 28) | (uint32_t)r.off);
    };
    auto emit_sweeps = [&](int type, std::vector<SlotSweep>& rows) {
        for (size_t q = 0; q < rows.size(); q += P.spw) {
            const size_t qe = std::min(rows.size(), q + P.spw);
            int clo = 1 << 30, chi = -1;
            for (size_t r = q; r < qe; ++r) {
                const int pl = rows[r].plane;
                clo = std::min(clo, B.chunk_of_plane[std::max(pl - 1, 0)]);
                chi = std::max(chi, B.chunk_of_plane[std::min(pl + 1, m)]);
            }
            HMG_CHECK(chi - clo <= 2, "apply plan task spans more than three chunks");
            std::array<int32_t, PLAN_TASK_INTS> t{};
            t[0] = type; t[1] = clo; t[2] = chi;
            for (int s = 0; s < 4; ++s) {
                int32_t* d = &t[4 + s * PLAN_SLOT_INTS];
                if (q + s < qe && s < P.spw) {
                    const SlotSweep& r = rows[q + s];
                    d[0] = enc(r.c, clo);
                    for (int a = 0; a < 3; ++a) { d[1 + a] = enc(r.mi[a], clo); d[4 + a] = enc(r.pl[a], clo); }
                    d[7] = r.cnt; d[8] = r.pout; d[9] = r.flags;
                } else {
                    d[0] = (int32_t)((3u << 28) | 1u);
                    for (int a = 0; a < 3; ++a) d[1 + a] = d[4 + a] = d[0];
                    d[7] = 0; d[8] = 0; d[9] = 0;
                }
            }
            tasks.push_back(t);
        }
        rows.clear();
    };
    struct Special { int p, cls, i, j, k; };
    std::vector<Special> specials;
    auto add_special = [&](int i, int j, int k) {
        const int p = B.pack(i, j, k);
        cover(p);
        specials.push_back({p, (int)(L.nodeinfo[p] >> 24), i, j, k});
    };
    const Builder::Ref ZERO{0, 1, true};

    bool has_interior = false;
    for (int p = 0; p < L.nf; ++p) has_interior |= (L.nodeinfo[p] >> 24) == 0;
    if (!has_interior) {
        // coarsest levels: no interior node, hence no interior stencil to derive the face stencils from
        for (int p = 0; p < L.nf; ++p) {
            const uint32_t info = L.nodeinfo[p];
            add_special(info & 255, (info >> 8) & 255, (info >> 16) & 255);
        }
    } else if (dim == 3) {
        auto make = [&](int i, int j, int kstart, int cnt, int flags) {
            SlotSweep r;
            r.plane = i; r.cnt = cnt; r.flags = flags; r.pout = B.pack(i, j, kstart);
            auto line = [&](int ii, int jj) -> Builder::Ref {
                if (ii < 0 || ii > m || jj < 0 || jj > m - ii) return ZERO;
                return B.ref_node(ii, B.pack(ii, jj, 0), kstart);
            };
            r.c = line(i, j);
            r.mi[0] = line(i + 1, j); r.mi[1] = line(i, j + 1); r.mi[2] = line(i - 1, j + 1);
            r.pl[0] = line(i - 1, j); r.pl[1] = line(i, j - 1); r.pl[2] = line(i + 1, j - 1);
            for (int k = 0; k < cnt; ++k) cover(r.pout + k);
            return r;
        };
        for (int i = 0; i <= m; ++i) {
            std::vector<SlotSweep> r0, ra, rb;
            for (int j = 0; j <= m - i; ++j) {
                const int len = m - i - j + 1;
                const int R = (j == 0 ? 2 : 0) | (i == 0 ? 4 : 0);
                if (R == 0) {
                    if (len >= 2) r0.push_back(make(i, j, 0, len, 3));
                    else add_special(i, j, 0);
                } else if (R == 2 || R == 4) {
                    if (len >= 3) (R == 2 ? ra : rb).push_back(make(i, j, 1, len - 2, 0));
                    add_special(i, j, 0);
                    if (len >= 2) add_special(i, j, len - 1);
                } else {
                    for (int k = 0; k < len; ++k) add_special(i, j, k);
                }
            }
            emit_sweeps(TASK_SWEEP_INTERIOR, r0);
            emit_sweeps(TASK_SWEEP_FACE_A, ra);
            emit_sweeps(TASK_SWEEP_FACE_B, rb);
        }
    } else {
        const int SEG = 16;
        auto make = [&](int i, int jstart, int cnt, int flags) {
            SlotSweep r;
            r.plane = i; r.cnt = cnt; r.flags = flags; r.pout = B.pack(i, jstart, 0);
            auto line = [&](int ii) -> Builder::Ref {
                if (ii < 0 || ii > m) return ZERO;
                return B.ref_node(ii, B.pack(ii, 0, 0), jstart);
            };
            r.c = line(i);
            r.mi[0] = line(i + 1); r.mi[1] = r.mi[2] = ZERO;
            r.pl[0] = line(i - 1); r.pl[1] = r.pl[2] = ZERO;
            for (int k = 0; k < cnt; ++k) cover(r.pout + k);
            return r;
        };
        for (int i = 0; i <= m; ++i) {
            std::vector<SlotSweep> r0, ra;
            const int len = m - i + 1;
            if (i >= 1) {
                if (len >= 2) {
                    for (int j0 = 0; j0 < len; j0 += SEG) {
                        int cnt = std::min(SEG, len - j0);
                        if (len - (j0 + cnt) == 1) cnt += 1;      // do not leave a one-node tail segment
                        r0.push_back(make(i, j0, cnt, (j0 == 0 ? 1 : 0) | (j0 + cnt == len ? 2 : 0)));
                        if (j0 + cnt == len) break;
                    }
                } else {
                    add_special(i, 0, 0);
                }
            } else {
                add_special(0, 0, 0);
                if (len >= 2) add_special(0, len - 1, 0);
                for (int j0 = 1; j0 < len - 1; j0 += SEG) ra.push_back(make(0, j0, std::min(SEG, len - 1 - j0), 0));
            }
            emit_sweeps(TASK_SWEEP_INTERIOR, r0);
            emit_sweeps(TASK_SWEEP_FACE_A, ra);
        }
    }
    for (int p = 0; p < L.nf; ++p) HMG_CHECK(covered[p], "apply plan misses a node");

    // special nodes -> nodetab + node tasks (grouped by the chunk of plane max(i-1, 0))
    {
        const int ndir = dim == 3 ? NDIR3 : NDIR2;
        std::stable_sort(specials.begin(), specials.end(), [&](const Special& a, const Special& b) { return a.i < b.i; });
        P.nodetab.assign(specials.size() * 16, 0xFFFFFFFFu);
        std::vector<int> base_chunk(specials.size()), hi_chunk(specials.size());
        for (size_t q = 0; q < specials.size(); ++q) {
            const Special& s = specials[q];
            const int clo = B.chunk_of_plane[std::max(s.i - 1, 0)];
            base_chunk[q] = clo;
            hi_chunk[q] = B.chunk_of_plane[std::min(s.i + 1, m)];
            uint32_t* e = &P.nodetab[q * 16];
            for (int d = 0; d < ndir; ++d) {
                const int* v = dim == 3 ? DIRS3[d] : DIRS2[d];
                const int ni = s.i + v[0], nj = s.j + v[1], nk = s.k + (dim == 3 ? v[2] : 0);
                const bool inside = ni >= 0 && nj >= 0 && nk >= 0 && ni + nj + nk <= m;
                HMG_CHECK(inside == !dir_outside(dim, s.cls, d), "class does not describe the neighbourhood");
                if (!inside) continue;
                const int q2 = B.pack(ni, nj, nk);
                const int c = B.chunk_of_plane[ni];
                HMG_CHECK(c >= clo && c - clo <= 2, "special node neighbour outside the chunk window");
                e[d] = ((uint32_t)(c - clo) << 28) | (uint32_t)(q2 - B.chunk_start[c]);
            }
            e[15] = (uint32_t)s.p | ((uint32_t)s.cls << 16);
        }
        size_t q = 0;
        while (q < specials.size()) {
            size_t qe = q;
            while (qe < specials.size() && qe - q < (size_t)P.spw && base_chunk[qe] == base_chunk[q]) ++qe;
            std::array<int32_t, PLAN_TASK_INTS> t{};
            int chi = 0;
            for (size_t r = q; r < qe; ++r) chi = std::max(chi, hi_chunk[r]);
            t[0] = TASK_NODES; t[1] = base_chunk[q]; t[2] = chi;
            HMG_CHECK(chi - base_chunk[q] <= 2, "node task spans more than three chunks");
            for (int s = 0; s < 4; ++s) t[4 + s * PLAN_SLOT_INTS] = (q + s < qe && s < P.spw) ? (int32_t)(q + s) : -1;
            tasks.push_back(t);
            q = qe;
        }
    }
    std::stable_sort(tasks.begin(), tasks.end(), [](const auto& a, const auto& b) {
        if (a[1] != b[1]) return a[1] < b[1];
        return a[2] < b[2];
    });
    P.ntasks = (int)tasks.size();
    P.nwarps = std::max(1, std::min(8, P.ntasks));   // every consumer warp must own at least one task per unit
    P.tasks.reserve(tasks.size() * PLAN_TASK_INTS);
    for (const auto& t : tasks) P.tasks.insert(P.tasks.end(), t.begin(), t.end());
    return P;
}

}  // namespace hmg
