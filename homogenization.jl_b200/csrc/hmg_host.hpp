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
inline int pack3(int m, int i, int j, int k) {
    int n1 = m - i;
    return tot3(m) - tot3(n1) + tri(n1) - tri(n1 - j) + k;
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
    std::vector<uint16_t> face_bary; // 3D: interior nodes of a face, barycentric a | b<<8
    // packed index of the t-th paired node of every local cell: faces [4][npf] (3D), then edges
    // [6|3][npe], then vertices [4|3]
    std::vector<uint16_t> iface_idx;
    double mass_total = 0.0;         // sum of all entries of the reference mass matrix
};

struct RefElement {
    int dim = 0, nlevels = 0;
    int ndir = 0, nc = 0, ncls = 0;
    std::vector<RefLevel> lv;        // lv[l-1] = level l
};

RefElement build_reference(int dim, int nlevels);

// ---- streaming plan of the apply kernel (plan.cpp) ---------------------------------------
// Device vectors are ELEMENT-INTERLEAVED: W consecutive coarse elements form a group ("unit"),
// entry (element e, packed node p) lives at ((e / W) * nf + p) * W + e % W.  A warp lane is an
// element, so every stencil access of a warp is one coalesced, bank-conflict-free line and the
// control flow (node class, neighbour offsets) is warp-uniform.  A unit is streamed through shared
// memory in CHUNKS (whole lattice planes i = const, merged when tiny) by TMA bulk copies into a
// ring of slots; warps consume TASKS that only need chunks clo..chi (at most 3 consecutive).
constexpr int PLAN_SLOT_INTS = 12;
constexpr int PLAN_TASK_INTS = 4 + 4 * PLAN_SLOT_INTS;
enum PlanTaskType { TASK_SWEEP_INTERIOR = 0, TASK_SWEEP_FACE_A = 1, TASK_SWEEP_FACE_B = 2, TASK_NODES = 3 };

struct ApplyPlan {
    int W = 16;            // elements per unit (lanes per row slot)
    int spw = 2;           // row slots per warp = 32 / W
    int nchunks = 0;       // chunks per unit
    int nslots = 0;        // ring slots
    int slot_nodes = 0;    // nodes per slot (largest chunk + slack)
    int zero_nodes = 0;    // nodes of the zero line in front of the ring
    int nwarps = 8;        // consumer warps per CTA (one more warp produces)
    int ctas_per_sm = 1;
    int ntasks = 0;
    size_t smem_bytes = 0;
    std::vector<int32_t> chunk_start;   // [nchunks + 1] packed node offsets
    // task t: [type, clo, chi, 0] + 4 slots x 12 ints.
    //  sweep slot: [0] centre ref, [1..3] "minus" line refs, [4..6] "plus" line refs, [7] count,
    //              [8] packed index of the first node, [9] flags (1: first node is a face end, 2: last)
    //  node slot:  [0] index into nodetab (-1: idle)
    //  ref = (chunk - clo) << 28 | node offset inside the chunk; (3 << 28 | 1) = the zero line
    std::vector<int32_t> tasks;
    // special nodes (edge / vertex classes): [0] centre ref, [d] ref of neighbour d (0xFFFFFFFF outside),
    // [15] packed index | class << 16; refs relative to the chunk of plane max(i - 1, 0)
    std::vector<uint32_t> nodetab;
};
ApplyPlan build_apply_plan(int dim, const RefLevel& L, int W);
// single-face classes derive their coefficients from the interior ones: weight of direction d
double face_weight(int dim, int cls, int d);

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

// class bitmask of the reference-face set containing a local cell
int class_of_face(int lf);              // 3D face 0..3
int class_of_edge(int dim, int le);     // edge local id
int class_of_vertex(int dim, int lv);

// per element geometry: coef[e][c] = |J| * P_c (c < nc-1), coef[e][nc-1] = |J| with
// P = J^-1 diag(sigma) J^-T (src/apply_local_operators.jl:101-105, src/cell_values.jl:104-127)
void element_coefficients(int dim, int64_t ne, const double* nodes /*dim x nn*/,
                          const int64_t* elems /*0-based*/, const double* sigma /*dim x ne*/,
                          std::vector<double>& coef /*ne x stride*/, int stride);

}  // namespace hmg
