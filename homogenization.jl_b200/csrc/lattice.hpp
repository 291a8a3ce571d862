// Lattice index arithmetic shared by the kernels and the host-side verification code.
//
// The refined reference simplex of level l is the set of integer points (i, j[, k]) >= 0 with
// i + j [+ k] <= m, m = 2^(l-1) (reference vertex 1 = origin, vertex 2 = m e_1, ...).  The operator
// couples a node with its neighbours along 7 (3D) / 3 (2D) lattice directions and their opposites.
//
// "Packed" node order of the device vectors:
//   2D  lexicographic (i, j), j fastest: line i has m - i + 1 nodes.
//   3D  DIAGONAL-PLANE order (t = i + j, then i, then k), k fastest: plane t holds the t + 1 lines
//       (i, t - i), i = 0..t, which ALL have the same length m - t + 1.  The stencil of a node only
//       reaches planes t - 1, t, t + 1, and a plane has at most (m/2 + 1)^2 nodes (vs (m+1)(m+2)/2 for
//       planes i = const), so the streaming window of the apply kernel is half as large, and the
//       lines of a plane are interchangeable work items.
#pragma once
#ifdef __CUDACC__
#define HMG_HD __host__ __device__ __forceinline__
#else
#define HMG_HD inline
#endif

namespace hmg {

HMG_HD int lat_tri(int q) { return (q + 1) * (q + 2) / 2; }
HMG_HD int lat_tot3(int q) { return (q + 1) * (q + 2) * (q + 3) / 6; }
HMG_HD int lat_pack2(int m, int i, int j) { return lat_tri(m) - lat_tri(m - i) + j; }
// first packed index of diagonal plane t: sum_{s<t} (s + 1)(m - s + 1)
HMG_HD int lat_off3(int m, int t) { return (m + 1) * t * (t + 1) / 2 - (t - 1) * t * (t + 1) / 3; }
HMG_HD int lat_pack3(int m, int i, int j, int k) {
    const int t = i + j;
    return lat_off3(m, t) + i * (m - t + 1) + k;
}
// first packed index of line i (2D)
HMG_HD int lat_off2(int m, int i) { return i * (m + 1) - i * (i - 1) / 2; }

template <int DIM> struct Dims;
template <> struct Dims<3> { static constexpr int NDIR = 15, NC = 7, NCLS = 16, CS = 8; };
template <> struct Dims<2> { static constexpr int NDIR = 7, NC = 4, NCLS = 8, CS = 4; };

// direction d of the stencil (0 = centre); must match DIRS3 / DIRS2 of reference.cpp
HMG_HD void lat_dir3(int d, int& di, int& dj, int& dk) {
    // packed as 2-bit fields (value + 1) to stay in registers / immediates
    const int I[15] = {0, 1, -1, 0, 0, 0, 0, -1, 1, -1, 1, 0, 0, 1, -1};
    const int J[15] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, -1, 1, -1, 1};
    const int K[15] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1};
    di = I[d]; dj = J[d]; dk = K[d];
}
HMG_HD void lat_dir2(int d, int& di, int& dj) {
    const int I[7] = {0, 1, -1, 0, 0, -1, 1};
    const int J[7] = {0, 0, 0, 1, -1, 1, -1};
    di = I[d]; dj = J[d];
}

// packed-index offsets of the stencil neighbours of node (i, j, *) at lattice size m
template <int DIM> HMG_HD void neighbour_offsets(int m, int i, int j, int* off);
template <> HMG_HD void neighbour_offsets<3>(int m, int i, int j, int* off) {
    const int t = i + j;
    const int L = m - t + 1;                        // length of every line of plane t
    const int D = (t + 2) * L - 1 - i;              // (i+1, j, k) - (i, j, k): plane t+1, line i+1
    const int E = (t + 1) * L - i;                  // (i, j+1, k) - (i, j, k): plane t+1, line i
    const int F = i - (t + 1) * (L + 1);            // (i-1, j, k) - (i, j, k): plane t-1, line i-1
    const int G = i - t * (L + 1);                  // (i, j-1, k) - (i, j, k): plane t-1, line i
    off[0] = 0;       off[1] = D;      off[2] = F;
    off[3] = E;       off[4] = G;      off[5] = 1;       off[6] = -1;
    off[7] = -L;      off[8] = L;      off[9] = F + 1;   off[10] = D - 1;
    off[11] = G + 1;  off[12] = E - 1; off[13] = L + 1;  off[14] = -L - 1;
}
template <> HMG_HD void neighbour_offsets<2>(int m, int i, int, int* off) {
    const int B = m - i + 1;                        // length of row i
    off[0] = 0; off[1] = B; off[2] = -(B + 1); off[3] = 1; off[4] = -1; off[5] = -B; off[6] = B - 1;
}

template <int DIM> HMG_HD bool neighbour_inside(int m, int i, int j, int k, int d) {
    int di, dj, dk = 0;
    if (DIM == 3) lat_dir3(d, di, dj, dk); else lat_dir2(d, di, dj);
    const int ni = i + di, nj = j + dj, nk = k + dk;
    return ni >= 0 && nj >= 0 && nk >= 0 && ni + nj + nk <= m;
}

// local vertex ids of the reference faces / edges (src/grid.jl:89-91, 0-based)
HMG_HD void face_vertices(int lf, int* v) {
    const int F[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
    v[0] = F[lf][0]; v[1] = F[lf][1]; v[2] = F[lf][2];
}
template <int DIM> HMG_HD void edge_vertices(int le, int* v) {
    const int E3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    const int E2[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    v[0] = DIM == 3 ? E3[le][0] : E2[le][0];
    v[1] = DIM == 3 ? E3[le][1] : E2[le][1];
}

// packed index of the node with barycentric weights w[0..n) on the local vertices lv[0..n)
template <int DIM> HMG_HD int bary_to_packed(int m, const int* lv, const int* w, int n) {
    int lam[4] = {0, 0, 0, 0};
    for (int q = 0; q < 3; ++q)
        if (q < n) lam[lv[q]] = w[q];
    return DIM == 3 ? lat_pack3(m, lam[1], lam[2], lam[3]) : lat_pack2(m, lam[1], lam[2]);
}

// kind 0 = face (ab = barycentric weights a | b<<8 on the face's first two vertices), 1 = edge
// (q = weight on the edge's second vertex), 2 = vertex.  Written without indexed local arrays:
// lattice coordinates (i, j, k) are the barycentric weights of local vertices 1, 2, 3.
template <int DIM> HMG_HD int interface_node(int m, int kind, int lid, int q, unsigned ab) {
    int i = 0, j = 0, k = 0;
    if (kind == 0) {            // faces (0,1,2) (0,1,3) (0,2,3) (1,2,3)
        const int w0 = ab & 255, w1 = ab >> 8, w2 = m - w0 - w1;
        i = lid == 0 ? w1 : (lid == 1 ? w1 : (lid == 2 ? 0 : w0));
        j = lid == 0 ? w2 : (lid == 1 ? 0 : w1);
        k = lid == 0 ? 0 : w2;
    } else if (kind == 1) {
        const int w0 = m - q, w1 = q;
        if (DIM == 3) {         // edges (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
            i = lid == 0 ? w1 : ((lid == 3 || lid == 4) ? w0 : 0);
            j = lid == 1 ? w1 : (lid == 3 ? w1 : (lid == 5 ? w0 : 0));
            k = lid == 2 ? w1 : ((lid == 4 || lid == 5) ? w1 : 0);
        } else {                // edges (0,1) (0,2) (1,2)
            i = lid == 0 ? w1 : (lid == 2 ? w0 : 0);
            j = lid == 0 ? 0 : w1;
        }
    } else {
        i = lid == 1 ? m : 0;
        j = lid == 2 ? m : 0;
        k = lid == 3 ? m : 0;
    }
    return DIM == 3 ? lat_pack3(m, i, j, k) : lat_pack2(m, i, j);
}

// interpolation parents of fine node (i, j, k): returns 1 (coincides with a coarse node) or 2
// (midpoint of two coarse nodes, weights 1/2); pa/pb are packed indices on the coarse lattice mc.
// The parity -> direction table is verified against the edge graph in reference.cpp.
template <int DIM> HMG_HD int interp_parents(int mc, int i, int j, int k, int& pa, int& pb) {
    if (DIM == 3) {
        const int par = (i & 1) * 4 + (j & 1) * 2 + (k & 1);
        if (par == 0) { pa = pb = lat_pack3(mc, i >> 1, j >> 1, k >> 1); return 1; }
        const int dtab[8] = {0, 5, 3, 11, 1, 9, 7, 13};
        int di, dj, dk;
        lat_dir3(dtab[par], di, dj, dk);
        pa = lat_pack3(mc, (i + di) >> 1, (j + dj) >> 1, (k + dk) >> 1);
        pb = lat_pack3(mc, (i - di) >> 1, (j - dj) >> 1, (k - dk) >> 1);
        return 2;
    }
    const int par = (i & 1) * 2 + (j & 1);
    if (par == 0) { pa = pb = lat_pack2(mc, i >> 1, j >> 1); return 1; }
    const int dtab[4] = {0, 3, 1, 5};
    int di, dj;
    lat_dir2(dtab[par], di, dj);
    pa = lat_pack2(mc, (i + di) >> 1, (j + dj) >> 1);
    pb = lat_pack2(mc, (i - di) >> 1, (j - dj) >> 1);
    return 2;
}

}  // namespace hmg
