/*
 * hmg_introspect.h -- host-only verification entry points of libhmg_b200.so.
 *
 * They need no GPU and compute nothing for the product path: they expand the tables the kernels
 * consume (through the same index functions, csrc/lattice.hpp) so that the CPU test-suite can
 * compare them with the oracle's explicit operators and maps.  All return 0 on success;
 * hmg_host_last_error() describes the last failure.  Index arrays crossing this boundary are
 * 1-based Int64 for `elems1`, 0-based for everything returned.
 */
#ifndef HMG_B200_INTROSPECT_H
#define HMG_B200_INTROSPECT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* hmg_host_last_error(void);
/* sizes[8] = m, nf, ld, n_interior, n_boundary, npf, ndir, nc; hier2lat[nf], G[ncls*ndir*nc] may be NULL.
 * Restates what refined_element() fixes (src/multilevel_reference.jl:41-61). */
int hmg_host_reference(int dim, int nlevels, int level, int64_t* sizes, int32_t* hier2lat, double* G,
                       double* mass_total);
/* refined_mesh(implicit, level) (src/implicit_fine_grid.jl:24, built by refined_element, src/multilevel_reference.jl:41-61):
 * nodes[dim x nf] reference coordinates in hierarchical row order, elems1[(dim+1) x nel] 1-based and index-sorted per
 * element; arrays may be NULL (nel is always written).  Used by construct_full_grid for the VTK export. */
int hmg_host_refined_mesh(int dim, int nlevels, int level, double* nodes, int64_t* elems1, int64_t* nel);
/* dense nf x nf column-major, hierarchical order: sum_c coef[c] * table_c, i.e. the matrix
 * sum_kl |J| P_kl ops[k,l] + lambda |J| mass of src/apply_local_operators.jl:105-118 */
int hmg_host_local_matrix(int dim, int nlevels, int level, const double* coef, double* dense);
/* dense nf(level) x nf(level-1) interpolation operator (src/interpolation.jl:7-50) */
int hmg_host_transfer_matrix(int dim, int nlevels, int level_fine, double* dense);
/* hierarchical rows of the paired nodes of a local face (kind 0) / edge (1) / vertex (2):
 * must equal numbering.faces_interior / edges_interior / nodes (src/multilevel_reference.jl:125-203) */
int hmg_host_interface_rows(int dim, int nlevels, int level, int kind, int lid, int32_t* rows, int64_t* count);
/* kind 0 faces, 1 edges, 2 interface vertices, 3 all nodes (src/interface.jl:65-117); arrays may be NULL */
int hmg_host_topology(int dim, int64_t ne, int64_t nn, const int64_t* elems1, int kind, int64_t* ncells,
                      int64_t* nentries, int64_t* offset, int64_t* element, int64_t* local_id);
/* Dirichlet classes per element and interior-node flags (src/interface.jl:207-284, src/grid.jl:176-202) */
int hmg_host_boundary(int dim, int64_t ne, int64_t nn, const int64_t* elems1, uint16_t* cmask, uint8_t* interior);
int hmg_host_class_of(int dim, int kind, int lid);
/* partition of the coarse elements over ranks (the host logic behind hmg_create_partitioned): local elements of
 * `rank` with their Dirichlet class masks, owner counts mult[ne_local][16] (owners on all ranks of the cell behind
 * every node class), per base node the first local owner (element*8+local id, -1 none) and whether this rank
 * reports the node to the coarse solve on rank 0.  Arrays may be NULL. */
int hmg_host_partition_elements(int dim, int64_t ne, int64_t nn, const int64_t* elems1, const int32_t* owner_rank, int rank,
                                int nranks, int64_t* ne_local, int64_t* local_to_global, uint16_t* cmask, uint8_t* mult,
                                int32_t* node_first, uint8_t* node_contrib);
/* interface cells of one kind (0 faces, 1 edges, 2 vertices) as seen by `rank`.  cut == 0: cells whose owners are
 * all local, sizes[2] = cells, entries, CSR offset / element (LOCAL index) / local_id.  cut == 1: cut cells (owners
 * on several ranks) this rank takes part in, sizes[3] = cells, entries, cut cells of the kind on ALL ranks;
 * slot[cells] = ordinal in the global cut enumeration (the packed exchange buffer), first_local[cells] = 1 if the
 * globally first owner is this rank's first entry. */
int hmg_host_partition_cells(int dim, int64_t ne, int64_t nn, const int64_t* elems1, const int32_t* owner_rank, int rank,
                             int nranks, int kind, int cut, int64_t* sizes, int64_t* offset, int64_t* element, int64_t* local_id,
                             int64_t* slot, uint8_t* first_local);
int hmg_host_element_coefficients(int dim, int64_t ne, int64_t nn, const double* nodes, const int64_t* elems1,
                                  const double* sigma, double* coef, int stride);

/* neighbour exchange of the cut cells of one kind: per participating cell (same order as hmg_host_partition_cells
 * with cut == 1) the other ranks that share it, ascending (CSR peer_off / peer_rank), the ordinal of the cell in the
 * message exchanged with each of them (peer_idx) and the position of the own partial sum in the rank-ordered total
 * (my_pos); shared_with[nranks * 3] = cut cells of every kind shared with every rank (the message sizes).
 * sizes[2] = cells, peer entries.  Arrays may be NULL. */
int hmg_host_partition_peers(int dim, int64_t ne, int64_t nn, const int64_t* elems1, const int32_t* owner_rank, int rank,
                             int nranks, int kind, int64_t* sizes, int64_t* peer_off, int32_t* peer_rank, int32_t* peer_idx,
                             int32_t* my_pos, int64_t* shared_with);

/* runs the apply kernel's task enumeration and line sweeps (csrc/apply_core.cuh, the templates the
 * device kernel instantiates) for one element on the host: y = A x in lattice order, coef = |J| P (upper
 * triangle), lambda |J|; 2D lines are split into segments of 2^seg_shift nodes; info[4] = tasks, largest
 * row window of a task, ring rows, shared-memory bytes of the launch configuration */
int hmg_host_apply_sweep(int dim, int nlevels, int level, int seg_shift, const double* coef, const double* x,
                         double* y, int64_t* info);

#ifdef __cplusplus
}
#endif
#endif
