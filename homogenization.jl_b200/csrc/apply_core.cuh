// Core of the local-operator apply (K1), shared by the device kernel (kernels.cu) and the host-side
// emulation used by the CPU test-suite (introspect.cpp): node classes, the compile-time weight rule,
// the per-node evaluation and the line sweep with register sliding windows.
//
// What it replaces: the dim^2+1 CSC scatter-SpMVs per coarse element of
// src/apply_local_operators.jl:93-133.  On the refined reference simplex the local operator
//     A_e = sum_kl |J| P_kl ops[k,l] + lambda |J| mass
// is a constant-coefficient lattice stencil whose coefficient on the lattice segment (n, n+d) only
// depends on the set of reference faces that CONTAIN the segment:
//     no face    -> the interior coefficient c_d                      (all fine elements around it exist)
//     one face   -> c_d / 2                                           (the inner half exists; point symmetry)
//     two faces  -> a "special" coefficient of that reference edge    (tabulated, Ge)
// and the diagonal of a node is c_0 (interior), c_0 / 2 (one face) or a special value per class (Gc).
// reference.cpp verifies this rule node by node against the assembled integer stencil at setup.
#pragma once
#include <cmath>

#include "lattice.hpp"

namespace hmg {

// Face set a stencil direction points out of: a neighbour n+d of a node of class `cls` lies outside
// the simplex iff (cls & out_mask(d)) != 0 (the lattice simplex is convex and d has entries in {-1,0,1}).
// class bits 3D: 1 = K (k = 0), 2 = J (j = 0), 4 = I (i = 0), 8 = S (i + j + k = m);  2D: 1 = J, 2 = I, 4 = S.
template <int DIM> HMG_HD constexpr int out_mask(int d) {
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
HMG_HD constexpr int opp_dir(int d) { return d == 0 ? 0 : ((d & 1) ? d + 1 : d - 1); }
HMG_HD constexpr int popc4(int c) { return (c & 1) + ((c >> 1) & 1) + ((c >> 2) & 1) + ((c >> 3) & 1); }

enum WeightCode { W_OUT = 0, W_HALF = 1, W_FULL = 2, W_SPEC_C = 3, W_SPEC_E = 4 };
// faces of `cls` that contain the segment (n, n + d), for a direction that does not point outside
template <int DIM> HMG_HD constexpr int seg_class(int cls, int d) { return cls & ~out_mask<DIM>(opp_dir(d)); }
template <int DIM> HMG_HD constexpr int wcode(int cls, int d) {
    if (d == 0) return cls == 0 ? W_FULL : (popc4(cls) == 1 ? W_HALF : W_SPEC_C);
    if (cls & out_mask<DIM>(d)) return W_OUT;
    const int s = seg_class<DIM>(cls, d);
    return s == 0 ? W_FULL : (popc4(s) == 1 ? W_HALF : W_SPEC_E);
}

// Taps of a line sweep along the fastest lattice coordinate k.  Centre line: k-1, k, k+1.  NP "minus"
// lines with taps at k (m0) and k-1 (m1); NP "plus" lines with taps at k (p0) and k+1 (p1); tap t of
// minus line q is the opposite direction of tap t of plus line q.
//   3D (node (t = i+j, i, k)):  q = 0: minus = line (t+1, i+1), plus = line (t-1, i-1)
//                               q = 1: minus = line (t+1, i),   plus = line (t-1, i)
//                               q = 2: minus = line (t,   i-1), plus = line (t,   i+1)
//   2D (node (i, k = j)):       q = 0: minus = line i+1,        plus = line i-1
template <int DIM> struct Sweep;
template <> struct Sweep<3> {
    static constexpr int NP = 3, KP = 5, KM = 6;
    static constexpr int CK = 1, CJ = 2, CI = 4, CS = 8;      // class bits
    HMG_HD static constexpr int m0(int q) { return q == 0 ? 1 : (q == 1 ? 3 : 7); }     // (1,0,0) (0,1,0) (-1,1,0)
    HMG_HD static constexpr int m1(int q) { return q == 0 ? 10 : (q == 1 ? 12 : 14); }  // ... + (0,0,-1)
    HMG_HD static constexpr int p0(int q) { return q == 0 ? 2 : (q == 1 ? 4 : 8); }     // (-1,0,0) (0,-1,0) (1,-1,0)
    HMG_HD static constexpr int p1(int q) { return q == 0 ? 9 : (q == 1 ? 11 : 13); }   // ... + (0,0,1)
};
template <> struct Sweep<2> {
    static constexpr int NP = 1, KP = 3, KM = 4;
    static constexpr int CK = 1, CI = 2, CS = 4, CJ = 0;      // first node of a line (j = 0), line 0, last node
    HMG_HD static constexpr int m0(int) { return 1; }   // (1,0)
    HMG_HD static constexpr int m1(int) { return 6; }   // (1,-1)
    HMG_HD static constexpr int p0(int) { return 2; }   // (-1,0)
    HMG_HD static constexpr int p1(int) { return 5; }   // (-1,1)
};
// number of +-direction pairs + centre: rows of the interior coefficient table
template <int DIM> struct Pairs { static constexpr int N = 2 + 2 * Sweep<DIM>::NP; };
// direction whose coefficient is row r of the interior table: centre, +k, then (m0(q), m1(q)) per q
template <int DIM> HMG_HD constexpr int pair_dir(int r) {
    return r == 0 ? 0 : (r == 1 ? Sweep<DIM>::KP : (((r - 2) & 1) ? Sweep<DIM>::m1((r - 2) >> 1) : Sweep<DIM>::m0((r - 2) >> 1)));
}

// stencil tables of one level, small enough to travel in the kernel parameter block (constant bank)
template <int DIM> struct StencilTab {
    double gi[Pairs<DIM>::N][Dims<DIM>::NC];          // interior: one row per +-pair (and the centre)
    double gc[Dims<DIM>::NCLS][Dims<DIM>::NC];        // diagonal of the classes on >= 2 faces
    double ge[Dims<DIM>::NCLS][Dims<DIM>::NC];        // segment along the reference edge of a 2-face class
};

// per-lane (= per coarse element) operator data.  The interior stencil lives in registers; the element's
// |J| P / lambda |J| components are only needed again for the few "special" coefficients of edge and
// vertex nodes, so a kernel short of registers may leave them in memory (LaneOpMem).
template <int DIM> struct LaneOp {
    static constexpr int dim = DIM;
    double ec[Dims<DIM>::NC];                             // |J| P (upper triangle), lambda |J|
    double c0, cz, ca[Sweep<DIM>::NP], cb[Sweep<DIM>::NP];   // interior stencil, one value per +-pair
    HMG_HD double coef(int q) const { return ec[q]; }
};
template <int DIM> struct LaneOpMem {
    static constexpr int dim = DIM;
    const double* ecp;                                    // component q at ecp[q * stride]
    int stride;
    double lambda;                                        // scales the last component
    double c0, cz, ca[Sweep<DIM>::NP], cb[Sweep<DIM>::NP];
    HMG_HD double coef(int q) const {
#ifdef __CUDA_ARCH__
        const double v = __ldg(ecp + q * stride);
#else
        const double v = ecp[q * stride];
#endif
        return q == Dims<DIM>::NC - 1 ? v * lambda : v;
    }
};

template <class Op> HMG_HD double combine(const Op& op, const double* g) {
    double c = 0.0;
#pragma unroll
    for (int q = 0; q < Dims<Op::dim>::NC; ++q) c = fma(op.coef(q), g[q], c);
    return c;
}
template <class Op> HMG_HD void interior_coefficients(Op& op, const StencilTab<Op::dim>& T) {
    op.c0 = combine(op, T.gi[0]);
    op.cz = combine(op, T.gi[1]);
#pragma unroll
    for (int q = 0; q < Sweep<Op::dim>::NP; ++q) {
        op.ca[q] = combine(op, T.gi[2 + 2 * q]);
        op.cb[q] = combine(op, T.gi[3 + 2 * q]);
    }
}

template <int DIM, int CLS, int DIR, class Op>
HMG_HD void acc_tap(const Op& op, const StencilTab<DIM>& T, double c, double v, double& a1, double& ah) {
    constexpr int w = wcode<DIM>(CLS, DIR);
    if (w == W_FULL) a1 = fma(c, v, a1);
    else if (w == W_HALF) ah = fma(c, v, ah);
    else if (w == W_SPEC_E) a1 = fma(combine(op, T.ge[seg_class<DIM>(CLS, DIR)]), v, a1);
}
template <int DIM, int CLS, int Q, class Op>
HMG_HD void acc_lines(const Op& op, const StencilTab<DIM>& T, const double* Mm, const double* Mk,
                      const double* Pk, const double* Pp, double& a1, double& ah) {
    using S = Sweep<DIM>;
    acc_tap<DIM, CLS, S::m0(Q)>(op, T, op.ca[Q], Mk[Q], a1, ah);
    acc_tap<DIM, CLS, S::p0(Q)>(op, T, op.ca[Q], Pk[Q], a1, ah);
    acc_tap<DIM, CLS, S::m1(Q)>(op, T, op.cb[Q], Mm[Q], a1, ah);
    acc_tap<DIM, CLS, S::p1(Q)>(op, T, op.cb[Q], Pp[Q], a1, ah);
    if constexpr (Q + 1 < S::NP) acc_lines<DIM, CLS, Q + 1>(op, T, Mm, Mk, Pk, Pp, a1, ah);
}

// (A x) at one node of class CLS.  xm/x0/xp: centre line at k-1, k, k+1; Mm/Mk: minus lines at k-1, k;
// Pk/Pp: plus lines at k, k+1.  Taps that do not exist for the class are never read.
template <int DIM, int CLS, class Op>
HMG_HD double eval_node(const Op& op, const StencilTab<DIM>& T, double xm, double x0, double xp,
                        const double* Mm, const double* Mk, const double* Pk, const double* Pp) {
    using S = Sweep<DIM>;
    if (CLS == 0) {
        double a = op.c0 * x0, b = op.cz * (xm + xp);      // two chains for latency
#pragma unroll
        for (int q = 0; q < S::NP; ++q) {
            a = fma(op.ca[q], Mk[q] + Pk[q], a);
            b = fma(op.cb[q], Mm[q] + Pp[q], b);
        }
        return a + b;
    }
    double a1 = 0.0, ah = 0.0;
    constexpr int wc = wcode<DIM>(CLS, 0);
    if (wc == W_HALF) ah = op.c0 * x0;
    else a1 = combine(op, T.gc[CLS]) * x0;
    acc_tap<DIM, CLS, S::KP>(op, T, op.cz, xp, a1, ah);
    acc_tap<DIM, CLS, S::KM>(op, T, op.cz, xm, a1, ah);
    acc_lines<DIM, CLS, 0>(op, T, Mm, Mk, Pk, Pp, a1, ah);
    return fma(0.5, ah, a1);
}

template <int DIM, int CLS> HMG_HD constexpr bool uses_minus_k(int q) {   // tap of minus line q at k
    return wcode<DIM>(CLS, Sweep<DIM>::m0(q)) != W_OUT;
}
template <int DIM, int CLS> HMG_HD constexpr bool uses_minus_km(int q) {  // ... at k-1
    return wcode<DIM>(CLS, Sweep<DIM>::m1(q)) != W_OUT;
}
template <int DIM, int CLS> HMG_HD constexpr bool uses_plus_k(int q) { return wcode<DIM>(CLS, Sweep<DIM>::p0(q)) != W_OUT; }
template <int DIM, int CLS> HMG_HD constexpr bool uses_plus_kp(int q) { return wcode<DIM>(CLS, Sweep<DIM>::p1(q)) != W_OUT; }

// where the rows of a line and of its neighbour lines start (element offsets into the value store)
template <int DIM> struct LineGeo {
    int L;                                    // nodes of the line
    int k0, k1;                               // nodes [k0, k1) are computed by this task
    int bc;                                   // centre line, node 0
    int bm[Sweep<DIM>::NP], bp[Sweep<DIM>::NP];   // minus / plus lines, node 0 (unused ones: anything)
};

// Sweep nodes [k0, k1) of a line with L >= 2 nodes.  Classes: FIRST at k = 0, LAST at k = L - 1, MID
// in between.  RS = distance (in doubles) between consecutive nodes of a line.  Mem(addr) reads the
// input value, out.template put<CLS>(k, acc, x0) consumes the result.
template <int DIM, int MID, int FIRST, int LAST, class Op, class Mem, class Out>
HMG_HD void sweep_line(const Op& op, const StencilTab<DIM>& T, const Mem& mem, int RS, const LineGeo<DIM>& g, Out& out) {
    using S = Sweep<DIM>;
    constexpr int NP = S::NP;
    int k = g.k0;
    double xm = 0.0, x0, xp = 0.0, Mm[NP], Mk[NP], Pk[NP], Pp[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) { Mm[q] = 0.0; Mk[q] = 0.0; Pk[q] = 0.0; Pp[q] = 0.0; }
    x0 = mem(g.bc + k * RS);
    if (k == 0) {
        xp = mem(g.bc + RS);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (uses_minus_k<DIM, FIRST>(q) || uses_minus_km<DIM, MID>(q) || uses_minus_km<DIM, LAST>(q)) Mk[q] = mem(g.bm[q]);
            if (uses_plus_k<DIM, FIRST>(q)) Pk[q] = mem(g.bp[q]);
            if (uses_plus_kp<DIM, FIRST>(q) || uses_plus_k<DIM, MID>(q) || uses_plus_k<DIM, LAST>(q)) Pp[q] = mem(g.bp[q] + RS);
        }
        out.template put<FIRST>(0, eval_node<DIM, FIRST>(op, T, 0.0, x0, xp, Mm, Mk, Pk, Pp), x0);
        xm = x0; x0 = xp;
#pragma unroll
        for (int q = 0; q < NP; ++q) { Mm[q] = Mk[q]; Pk[q] = Pp[q]; }
        k = 1;
    } else {
        xm = mem(g.bc + (k - 1) * RS);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (uses_minus_km<DIM, MID>(q) || uses_minus_km<DIM, LAST>(q)) Mm[q] = mem(g.bm[q] + (k - 1) * RS);
            if (uses_plus_k<DIM, MID>(q) || uses_plus_k<DIM, LAST>(q)) Pk[q] = mem(g.bp[q] + k * RS);
        }
    }
    const int kend = g.k1 < g.L - 1 ? g.k1 : g.L - 1;
    // running pointers to what node k loads next -- xp on the centre line, the minus lines at k, the plus lines at
    // k + 1: one add per line and TWO nodes (the second node of the unrolled pair is an immediate offset) instead of
    // a multiply-add from k per load (15 integer instructions per two nodes in the 3D loop of round 1)
    const double* ac = mem.ptr(g.bc + (k + 1) * RS);
    const double *am[NP], *ap[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) { am[q] = mem.ptr(g.bm[q] + k * RS); ap[q] = mem.ptr(g.bp[q] + (k + 1) * RS); }
    auto node = [&](int kk) {
        xp = mem.ld(ac);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (uses_minus_k<DIM, MID>(q) || uses_minus_km<DIM, MID>(q) || uses_minus_km<DIM, LAST>(q)) Mk[q] = mem.ld(am[q]);
            if (uses_plus_kp<DIM, MID>(q) || uses_plus_k<DIM, MID>(q) || uses_plus_k<DIM, LAST>(q)) Pp[q] = mem.ld(ap[q]);
        }
        out.template put<MID>(kk, eval_node<DIM, MID>(op, T, xm, x0, xp, Mm, Mk, Pk, Pp), x0);
        xm = x0; x0 = xp;
        ac += RS;
#pragma unroll
        for (int q = 0; q < NP; ++q) { Mm[q] = Mk[q]; Pk[q] = Pp[q]; am[q] += RS; ap[q] += RS; }
    };
    // four nodes per trip where the registers allow it (Out::UNROLL4): one pointer add per line and four nodes, more
    // loads in flight.  Measured on C4 / C2 (profiles/r02t_unroll_AB.jsonl): product 4.54 -> 4.38 ms, residual 6.74 ->
    // 6.33 ms, 2D residual 2.84 -> 2.43 ms; the 3D variants with the fused reduction lose 2 % (registers) and keep two.
    if constexpr (Out::UNROLL4) {
#pragma unroll 1
        for (; k + 3 < kend; k += 4) { node(k); node(k + 1); node(k + 2); node(k + 3); }
    }
#pragma unroll 1
    for (; k + 1 < kend; k += 2) { node(k); node(k + 1); }
    if (k < kend) { node(k); ++k; }
    if (g.k1 == g.L) {     // k == L - 1
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (uses_minus_k<DIM, LAST>(q)) Mk[q] = mem.ld(am[q]);
            if (uses_plus_kp<DIM, LAST>(q)) Pp[q] = mem.ld(ap[q]);
        }
        out.template put<LAST>(k, eval_node<DIM, LAST>(op, T, xm, x0, 0.0, Mm, Mk, Pk, Pp), x0);
    }
}

// a line with a single node (the last plane / line of the simplex)
template <int DIM, int CLS, class Op, class Mem, class Out>
HMG_HD void single_node(const Op& op, const StencilTab<DIM>& T, const Mem& mem, int RS, const LineGeo<DIM>& g, Out& out) {
    using S = Sweep<DIM>;
    constexpr int NP = S::NP;
    double Mm[NP], Mk[NP], Pk[NP], Pp[NP];
    const double x0 = mem(g.bc);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        Mm[q] = 0.0;
        Mk[q] = uses_minus_k<DIM, CLS>(q) ? mem(g.bm[q]) : 0.0;
        Pk[q] = uses_plus_k<DIM, CLS>(q) ? mem(g.bp[q]) : 0.0;
        Pp[q] = uses_plus_kp<DIM, CLS>(q) ? mem(g.bp[q] + RS) : 0.0;
    }
    out.template put<CLS>(0, eval_node<DIM, CLS>(op, T, 0.0, x0, 0.0, Mm, Mk, Pk, Pp), x0);
}

// Rows (packed indices relative to the element) of a line and of its neighbour lines, and the row
// window [behind, need] the task reads.  3D: line i of diagonal plane t, po = lat_off3(m, t).
struct LineRows3 { int L, c, rm[3], rp[3], need, behind; };
HMG_HD LineRows3 line_rows3(int m, int t, int i, int po) {
    LineRows3 r;
    const int L = m - t + 1;
    const int pn = po + (t + 1) * L, pp = po - t * (L + 1);
    r.L = L;
    r.c = po + i * L;
    r.rm[0] = pn + (i + 1) * (L - 1); r.rm[1] = pn + i * (L - 1); r.rm[2] = r.c - L;
    r.rp[0] = pp + (i - 1) * (L + 1); r.rp[1] = pp + i * (L + 1); r.rp[2] = r.c + L;
    r.need = t < m ? r.rm[0] + L - 2 : r.c + (i < t ? 1 : 0);
    r.behind = t > 0 ? pp + (i > 0 ? i - 1 : 0) * (L + 1) : po;
    return r;
}
// 2D: nodes [k0, k1) of line i
struct LineRows2 { int L, c, rm[1], rp[1], need, behind; };
HMG_HD LineRows2 line_rows2(int m, int i, int k0, int k1) {
    LineRows2 r;
    const int L = m - i + 1;
    r.L = L;
    r.c = lat_off2(m, i);
    r.rm[0] = r.c + L;
    r.rp[0] = r.c - (L + 1);
    const int kn = k1 - 1 < L - 2 ? k1 - 1 : L - 2;            // last node of the next line that is read
    r.need = i < m ? (kn >= 0 ? r.rm[0] + kn : r.c + k1 - 1) : r.c;
    if (r.need < r.c + (k1 < L ? k1 : L - 1)) r.need = r.c + (k1 < L ? k1 : L - 1);
    // oldest row read: node k0 - 1 of the previous line; monotone in task order (line 0 stays resident until line 1 starts)
    r.behind = i > 0 ? r.rp[0] + (k0 > 0 ? k0 - 1 : 0) : r.c;
    return r;
}

// dispatch of one task to the sweep of its line type
template <class Op, class Mem, class Out>
HMG_HD void run_line3(const Op& op, const StencilTab<3>& T, const Mem& mem, int RS, const LineGeo<3>& g, int t, int i, Out& out) {
    constexpr int K = 1, J = 2, I = 4, S = 8;
    if (g.L == 1) {
        if (i == 0) single_node<3, I | K | S>(op, T, mem, RS, g, out);
        else if (i == t) single_node<3, J | K | S>(op, T, mem, RS, g, out);
        else single_node<3, K | S>(op, T, mem, RS, g, out);
    } else if (t == 0) sweep_line<3, I | J, I | J | K, I | J | S>(op, T, mem, RS, g, out);
    else if (i == 0) sweep_line<3, I, I | K, I | S>(op, T, mem, RS, g, out);
    else if (i == t) sweep_line<3, J, J | K, J | S>(op, T, mem, RS, g, out);
    else sweep_line<3, 0, K, S>(op, T, mem, RS, g, out);
}
template <class Op, class Mem, class Out>
HMG_HD void run_line2(const Op& op, const StencilTab<2>& T, const Mem& mem, int RS, const LineGeo<2>& g, int i, Out& out) {
    constexpr int J = 1, I = 2, S = 4;
    if (g.L == 1) single_node<2, J | S>(op, T, mem, RS, g, out);
    else if (i == 0) sweep_line<2, I, I | J, I | S>(op, T, mem, RS, g, out);
    else sweep_line<2, 0, J, S>(op, T, mem, RS, g, out);
}

}  // namespace hmg
