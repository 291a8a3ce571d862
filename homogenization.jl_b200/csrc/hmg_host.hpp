// Host-side setup structures of libhmg_b200 (plain C++17, no CUDA types).
//
// The reference builds, per level, dim^2 sparse matrices ops[k,l] and a mass matrix on the
// refined reference simplex (src/build_local_operators.jl:51-141) in a *hierarchical* node
// order (src/multilevel_reference.jl:41-61).  This library never stores those matrices: the
// refined reference simplex is a uniform simplicial lattice, so the operator of one coarse
// element is a constant-coefficient stencil (15-point in 3D, 7-point in 2D) whose truncation
// on the simplex boundary depends only on WHICH reference faces a node lies on (its "class").
// Everything here is integer-exact; the structural assumptions are verified at setup and the
// library refuses to run if one does not hold.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace hmg {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define HMG_CHECK(cond, msg)                                                      \
    do {                                                                          \
        if (!(cond)) throw ::hmg::Error(std::string("hmg: ") + (msg) + " [" #cond "]"); \
    } while (0)

// ---- lattice helpers -----------------------------------------------------------------
inline int tri(int q) { return (q + 1) * (q + 2) / 2; }            // #points of a 2-simplex lattice
inline int tot3(int q) { return (q + 1) * (q + 2) * (q + 3) / 6; } // #points of a 3-simplex lattice
// packed (lexicographic, last coordinate fastest) index of a lattice point
inline int pack2(int m, int i, int j) { return tri(m) - tri(m - i) + j; }
// 3D: diagonal-plane order (t = i + j, then i, then k), see lattice.hpp
inline int pack3(int m, int i, int j, int k) {
    const int t = i + j;
    return (m + 1) * t * (t + 1) / 2 - (t - 1) * t * (t + 1) / 3 + i * (m - t + 1) + k;
}

// stencil directions (index 0 = centre); the set is verified against the assembled stencil
constexpr int NDIR3 = 15, NDIR2 = 7;
extern const int DIRS3[NDIR3][3];
extern const int DIRS2[NDIR2][2];
// symmetric tensor components: 3D (00,01,02,11,12,22) + mass = 7; 2D (00,01,11) + mass = 4
constexpr int NC3 = 7, NC2 = 4;
constexpr int NCLS3 = 16, NCLS2 = 8;   // class = bitmask of reference faces the node lies on

struct RefLevel {
    int m = 0;     // lattice size 2^(level-1)
    int nf = 0;    // nodes of the refined reference element
    int ld = 0;    // = nf (kept for the introspection ABI; the device layout is element-interleaved)
    std::vector<int32_t> hier2lat;   // hierarchical row -> packed lattice index
    std::vector<uint32_t> nodeinfo;  // per packed index: i | j<<8 | k<<16 | class<<24
    std::vector<uint32_t> interior;  // interior nodes (class 0), packed p | i<<14 | j<<22
    std::vector<uint32_t> boundary;  // boundary nodes sorted by class, p | class<<14 ; (i,j,k) via nodeinfo
    // transfer tables between this level (fine) and the next coarser one (levels >= 2):
    std::vector<uint32_t> interp_tab;   // per fine node: coarse parents pa | pb<<16 (pa == pb: coincident)
    std::vector<uint16_t> restrict_tab; // per coarse node: [ndir] fine indices (centre first), 0xFFFF = outside
    std::vector<double> G;           // [ncls][ndir][nc] stencil table, scale factors folded in
    // the tables the apply kernel is launched with (StencilTab of apply_core.cuh): interior rows
    // [npair][nc], diagonal of every class [ncls][nc], reference-edge segment of 2-face classes [ncls][nc]
    std::vector<double> gi, gc, ge;
    std::vector<double> dphi;        // [nf][dim] int d phi_p / d x_j over the refined reference element (lattice order)
    std::vector<uint16_t> face_bary; // 3D: interior nodes of a face, barycentric a | b<<8
    // packed index of the t-th paired node of every local cell: faces [4][npf] (3D), then edges
    // [6|3][npe], then vertices [4|3]
    std::vector<uint16_t> iface_idx;
    double mass_total = 0.0;         // sum of all entries of the reference mass matrix
    // refined_mesh(implicit, level) itself (src/implicit_fine_grid.jl:24), for construct_full_grid / the VTK export
    // only: lattice coordinates of every hierarchical row [nf][3] and the fine elements [(dim+1) x nel] in hierarchical
    // node ids, each index-sorted as refined_element leaves them (src/multilevel_reference.jl:56-58)
    std::vector<int32_t> hier_coords;
    std::vector<int32_t> cells;
};

struct RefElement {
    int dim = 0, nlevels = 0;
    int ndir = 0, nc = 0, ncls = 0;
    std::vector<RefLevel> lv;        // lv[l-1] = level l
};

RefElement build_reference(int dim, int nlevels);

// ---- base-mesh topology ---------------------------------------------------------------
struct CellMap {                     // CSR cell -> (element, local id), owners ascending
    std::vector<int64_t> offset;     // ncells + 1
    std::vector<int32_t> owner;      // element * 8 + local id
    std::vector<int64_t> cell_key;   // first global vertex id of the cell (diagnostics / partition)
    int64_t ncells() const { return offset.empty() ? 0 : (int64_t)offset.size() - 1; }
};

struct Topology {
    int dim = 0;
    int64_t ne = 0, nn = 0;
    CellMap faces, edges, verts;         // interface cells (>= 2 owners), src/interface.jl:65-117
    std::vector<int32_t> node_first;     // per base node: first owner (element*8+local), -1 if unused
    std::vector<int64_t> nodeown_off;    // all_nodes CSR (src/interface.jl:82-88)
    std::vector<int32_t> nodeown;
    std::vector<uint16_t> cmask;         // per element: bit c set <=> class c is on the domain boundary
    std::vector<uint8_t> node_boundary;  // per base node: 1 if on the domain boundary
    std::vector<int64_t> interior_nodes; // complement (sorted), src/grid.jl:176-202
    // unified numbering of the interface cells: faces, then edges, then vertices
    std::vector<int64_t> cell_off;       // CSR over all interface cells
    std::vector<int32_t> cell_own;
    std::vector<int32_t> elem_cells;     // [ne][16]: unified cell id of the element's local faces (0..3),
                                         // edges (4..9), vertices (10..13) in 3D; edges (0..2), vertices (3..5)
                                         // in 2D; -1 if the cell has a single owner
    // cut cells (multi-GPU): filled by the partitioner
};

// elems: (dim+1) x ne, 0-based, sorted per element
Topology build_topology(int dim, int64_t ne, int64_t nn, const int64_t* elems);

// ---- partition of the coarse elements over ranks (one process per GPU) ---------------------
// A rank keeps the columns of its own elements (global order preserved).  Interface cells whose owners
// all live on one rank are summed locally; CUT cells (owners on >= 2 ranks) are summed through a packed
// buffer that every rank fills with the partial sum over ITS owners and that is all-reduced: slot
// layout per level = [cut faces x npf][cut edges x npe][cut vertices], cells in global cell order, so
// every rank derives the same layout from the global topology without communication.
struct CutCells {                      // the cut cells of one kind this rank takes part in
    int64_t nglobal = 0;               // cut cells of this kind on all ranks (slot space)
    std::vector<int64_t> slot;         // per participating cell: ordinal among the global cut cells of the kind
    std::vector<int64_t> offset;       // CSR over the LOCAL owners of the cell
    std::vector<int32_t> owner;        // local element * 8 + local id, ascending global element
    std::vector<uint8_t> first_local;  // 1 if the globally first owner of the cell is owner[offset[c]]
    // neighbour exchange: the OTHER ranks that share the cell, ascending, with the ordinal of the cell among
    // the cut cells of this kind the two ranks share (both enumerate the global cells in the same order)
    std::vector<int64_t> peer_off;     // CSR over the peers of the cell
    std::vector<int32_t> peer_rank;
    std::vector<int32_t> peer_idx;
    std::vector<int32_t> my_pos;       // number of peers with a smaller rank (position of the own partial sum)
    int64_t ncells() const { return (int64_t)slot.size(); }
};
struct Partition {
    int rank = 0, nranks = 1;
    std::vector<int64_t> local_to_global;   // local element -> global element
    std::vector<int32_t> global_to_local;   // global element -> local element, -1 if remote
    CellMap faces, edges, verts;            // interface cells with all owners on this rank (local ids)
    CutCells cut[3];                        // 0 faces (3D), 1 edges, 2 vertices
    std::vector<int64_t> shared_with;       // [nranks][3] cut cells of every kind shared with each other rank
    std::vector<uint16_t> cmask;            // per local element
    std::vector<uint8_t> mult;              // [ne_local][16] owners (on all ranks) of the cell behind every node class
    std::vector<int32_t> node_first;        // per base node: first LOCAL owner (local element*8+local id), -1 if none
    std::vector<uint8_t> node_contrib;      // per base node: 1 if this rank reports the node to the coarse solve
};
Partition build_partition(const Topology& T, const int32_t* owner_rank, int rank, int nranks);

// class bitmask of the reference-face set containing a local cell
int class_of_face(int lf);              // 3D face 0..3
int class_of_edge(int dim, int le);     // edge local id
int class_of_vertex(int dim, int lv);

// per element geometry: coef[e][c] = |J| * P_c (c < nc-1), coef[e][nc-1] = |J| with
// P = J^-1 diag(sigma) J^-T (src/apply_local_operators.jl:101-105, src/cell_values.jl:104-127)
// flux[e][k] = -|J| (J^-1 (sigma_e .* xi))_k : the vector P of rhs_a_xi_grad_v! / integrate_first_term
// (src/examples/homogenized_coefficients.jl:468, 612)
void element_flux_vectors(int dim, int64_t ne, const double* nodes, const int64_t* elems, const double* sigma,
                          const double* xi, std::vector<double>& flux /*ne x dim*/);
void element_coefficients(int dim, int64_t ne, const double* nodes /*dim x nn*/,
                          const int64_t* elems /*0-based*/, const double* sigma /*dim x ne*/,
                          std::vector<double>& coef /*ne x stride*/, int stride);

}  // namespace hmg
