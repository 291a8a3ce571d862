/*
 * hmg.h -- C ABI of libhmg_b200.so: the B200-native implementation of the hot path of
 * haampie/Homogenization.jl (matrix-free operator on the implicit fine grid + geometric
 * multigrid V-cycle).  This is the drop-in boundary: plain pointers and sizes, no C++ or
 * torch types.  The reference has no FFI of its own (it is pure Julia); every entry point
 * below names the Julia function (file:line, relative to the reference repository root)
 * whose device method it backs -- see INTEGRATION.md for the `ccall` shim.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; hmg_last_error() describes
 *     the last failure of the calling thread.  No exception crosses this boundary.
 *   - indices that cross the boundary are Julia-native: 1-based Int64.
 *   - levels are 1-based as in the reference: level 1 = base mesh, level `nlevels` = finest.
 *   - host matrices are the reference's layout: column-major Nf(level) x Ne Float64 with the
 *     rows in the reference's HIERARCHICAL node order (src/multilevel_reference.jl:41-61);
 *     the library permutes to its lattice order and interleaves groups of columns internally.
 *   - a context is owned by the caller (hmg_destroy), is not thread-safe, and all calls are
 *     stream-ordered on the context's CUDA stream; calls that return host data synchronise.
 *     The library keeps per-device launch attributes in process-wide tables without a lock: make the calls
 *     of one process from one host thread at a time (one process per GPU is the multi-GPU model).
 *   - there is no CPU fallback: without a CUDA device hmg_create fails.
 */
#ifndef HMG_B200_H
#define HMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hmg_ctx hmg_ctx;

/* the five state vectors of LevelState (src/multigrid.jl:7-25) plus two scratch vectors
 * (V = v_prev of src/examples/homogenized_coefficients.jl:243, W = work). */
enum hmg_vec { HMG_X = 0, HMG_B = 1, HMG_R = 2, HMG_P = 3, HMG_AP = 4, HMG_V = 5, HMG_W = 6,
               HMG_NVEC = 7 };

const char* hmg_last_error(void);
int hmg_version(void);

/* ImplicitFineGrid(base, levels) (src/implicit_fine_grid.jl:13-18) + ZeroDirichletConstraint(
 * list_boundary_nodes_edges_faces(base)...) (src/interface.jl:207-284) + L2PlusDivAGrad(diff,
 * mass, constraint, lambda, sigma) (src/build_local_operators.jl:26-32) + LevelState per
 * level (src/multigrid.jl:18-25), all on `device`.
 *   base_nodes: dim x nn column-major; base_elems: (dim+1) x ne, 1-based, each element sorted
 *   ascending (asserted, src/implicit_fine_grid.jl:14); sigma: dim x ne (diagonal tensors). */
int hmg_create(int dim, int nlevels, int64_t ne, int64_t nn, const double* base_nodes,
               const int64_t* base_elems, const double* sigma, double lambda, int device,
               hmg_ctx** out);
int hmg_destroy(hmg_ctx* ctx);

/* multi-GPU: one process per GPU; coarse elements are partitioned by `owner_rank[ne]`; this
 * process keeps the columns with owner_rank == rank (global order preserved).  nccl_id is the
 * 128-byte ncclUniqueId produced by rank 0.  Only interface partial sums and scalars move.
 * Creation and hmg_destroy of a partitioned context are COLLECTIVE (every rank maps the communication buffers of all
 * others and nobody may free one a peer still has mapped): call them on all ranks, in the same order; so is every
 * operation that sums over interfaces or reduces a scalar (broadcast_interfaces, zero_out_all_but_one, dot, apply_global,
 * smoothing_steps, vcycle(s), the integrals) and the coarse-matrix setup (hmg_assemble_coarse / hmg_set_coarse_matrix:
 * rank 0 factorises, the others wait in a barrier).  A rank that stops answering makes the others' kernels trap after
 * about a minute. */
int hmg_create_partitioned(int dim, int nlevels, int64_t ne, int64_t nn, const double* base_nodes,
                           const int64_t* base_elems, const double* sigma, double lambda,
                           int device, int rank, int nranks, const int32_t* owner_rank,
                           const void* nccl_id, hmg_ctx** out);
int hmg_nccl_unique_id(void* out128);

/* sizes: nnodes(refined_mesh(implicit, level)), local column count, rows of a device column
 * (= nf), and the interleave width W of the device layout: entry (column e, lattice node p) of
 * a level vector lives at ((e / W) * nf + p) * W + e % W */
int64_t hmg_nf(const hmg_ctx* ctx, int level);
int64_t hmg_ne_local(const hmg_ctx* ctx);
int64_t hmg_ld(const hmg_ctx* ctx, int level);
int hmg_group_width(const hmg_ctx* ctx);
/* how the ranks of a partitioned context talk: 0 = one rank, 1 = NCCL (all-reduce of the CG scalars, grouped send/recv
 * of the cut cells), 2 = peer memory over NVLink (CUDA IPC: the reduction kernels sum over the ranks themselves, the
 * pack kernel stores into the neighbours' buffers).  2 is the default where every rank can map every other's buffer;
 * HMG_PEER=0 forces 1.  The coarse-level reduce / broadcast is NCCL in both. */
int hmg_comm_mode(const hmg_ctx* ctx);
/* global element index (1-based) of local column j (1-based) */
int hmg_local_elements(const hmg_ctx* ctx, int64_t* out);

/* operator.lambda = lambda (src/examples/homogenized_coefficients.jl:331) / new sigma */
int hmg_set_lambda(hmg_ctx* ctx, double lambda);
int hmg_set_sigma(hmg_ctx* ctx, const double* sigma);

/* host <-> device copies of one state matrix (reference layout on the host side);
 * ld_host >= Nf(level).  Only the local columns are transferred (ne_local of them). */
int hmg_upload(hmg_ctx* ctx, int level, int which, const double* host, int64_t ld_host);
int hmg_download(hmg_ctx* ctx, int level, int which, double* host, int64_t ld_host);
/* the first `nrows` rows of every local column: x[1 : nnodes(refined_mesh(implicit, k)), :] with nrows = Nf(k) are the
 * values on the nodes of the coarser level k (export_unknown, src/examples/homogenized_coefficients.jl:81-87); only
 * nrows x ne_local doubles cross PCIe.  ld_host >= nrows. */
int hmg_download_rows(hmg_ctx* ctx, int level, int which, int64_t nrows, double* host, int64_t ld_host);
/* domain shrink without a host round trip: dst.vec[dst_which] = src.vec[src_which][:, OneTo(ne_local(dst))]
 * (shrink_level_state, src/examples/homogenized_coefficients.jl:54-60, caller :325-327).  Both contexts live on the same
 * device, are not partitioned, and `dst` was created on an element prefix of `src`'s base mesh. */
int hmg_copy_columns_from(hmg_ctx* dst, int level, int dst_which, hmg_ctx* src, int src_which);
int hmg_fill(hmg_ctx* ctx, int level, int which, double value);          /* fill!           */
int hmg_copy(hmg_ctx* ctx, int level, int dst, int src);                 /* copyto!         */
int hmg_axpy(hmg_ctx* ctx, int level, double alpha, int x, int y);       /* axpy!           */
int hmg_dot(hmg_ctx* ctx, int level, int a, int b, double* out);         /* dot (all stored entries, src/multigrid.jl:54) */

/* mul!(alpha, base, A, x, y): y <- alpha*A*x + y, column-local, no constraint, no interface
 * sum (src/apply_local_operators.jl:85-120). */
int hmg_mul(hmg_ctx* ctx, int level, double alpha, int x, int y);
/* the "global product" of src/multigrid.jl:58-61 fused: y = broadcast(constraint(A*x)) */
int hmg_apply_global(hmg_ctx* ctx, int level, int x, int y);
/* apply_constraint!(x, level, z, implicit) (src/implicit_fine_grid.jl:94-139) */
int hmg_apply_constraint(hmg_ctx* ctx, int level, int which);
/* broadcast_interfaces!(x, implicit, level) (src/implicit_fine_grid.jl:209-328) */
int hmg_broadcast_interfaces(hmg_ctx* ctx, int level, int which);
/* zero_out_all_but_one!(x, implicit, level) (src/implicit_fine_grid.jl:334-386) */
int hmg_zero_out_all_but_one(hmg_ctx* ctx, int level, int which);
/* local_residual!(implicit, A, curr, k): r = b - A*x, constraint (src/apply_local_operators.jl:18-27) */
int hmg_local_residual(hmg_ctx* ctx, int level);
/* restrict_to!(levels[k-1].b, P, levels[k].r) (src/interpolation.jl:64-74) */
int hmg_restrict(hmg_ctx* ctx, int level_fine);
/* interpolate_and_sum_to!(levels[k].x, P, levels[k-1].x) (src/interpolation.jl:52-62) */
int hmg_interpolate_add(hmg_ctx* ctx, int level_fine);
/* smoothing_steps!(steps, implicit, ops, curr, k) (src/multigrid.jl:46-71).  x and r end up as in the reference;
 * p and Ap are scratch (the reference's last, dead update of p is skipped; the copy p = r is folded into the first
 * CG update; the p buffer may be swapped with an internal one by the fused direction update). */
int hmg_smoothing_steps(hmg_ctx* ctx, int level, int steps);

/* BaseLevel (src/multigrid.jl:30-41): the caller's factorisation F = cholesky(A[interior,
 * interior]) (src/examples/homogenized_coefficients.jl:259-261) is replaced by handing the
 * CSC matrix itself (1-based colptr/rowval) and the 1-based interior node list. */
int hmg_set_coarse_matrix(hmg_ctx* ctx, int64_t n_interior, const int64_t* colptr,
                          const int64_t* rowval, const double* nzval,
                          const int64_t* interior_nodes);
/* or: assemble_checkerboard(base, sigma, lambda)[interior, interior] + list_interior_nodes
 * inside the library (src/examples/homogenized_coefficients.jl:358-402, src/grid.jl:176-202).
 * Size ceiling of both: the inverse is DENSE on GPU 0, 8 n^2 bytes for n interior base nodes (twice that while it is
 * factorised), next to the state vectors: n = 29 791 (C4) takes 7 GB, n = 65 025 (C2 at 256^2 cells) 34 GB; the call
 * fails with a message when the device memory left does not suffice. */
int hmg_assemble_coarse(hmg_ctx* ctx);
/* copy_to_base!(u, v, implicit) / distribute!(v, u, implicit) on level 1
 * (src/implicit_fine_grid.jl:148-202); u has nn entries (host) */
int hmg_copy_to_base(hmg_ctx* ctx, int which, double* u_host);
int hmg_distribute(hmg_ctx* ctx, int which, const double* u_host);

/* vcycle!(implicit, base, ops, levels, k, steps) (src/multigrid.jl:73-119).  Levels below
 * `top_level` use 2 smoothing steps (the reference does not forward `steps`, :109).
 * out_resnorm (may be NULL): norm(zero_out_all_but_one!(r_top)) -- the logged residual of
 * src/examples/homogenized_coefficients.jl:286-287.  With out_resnorm != NULL r_top is zeroed all-but-one IN PLACE, as
 * the reference's driver does; with out_resnorm == NULL it is left as the CG recurrence left it.
 * After the call x_top and r_top are the reference's; everything else (p, Ap on every level; x, b, r below the top) is
 * scratch: where nothing reads them, the r-update, rho', the interface sum of Ap and Ap itself of the LAST CG step of a
 * smoothing call are not computed (before the restriction local_residual! recomputes r; below the top nobody reads r).
 * x is bit-identical either way.
 * If lambda or sigma changed since the coarse matrix was factorised (hmg_set_lambda / hmg_set_sigma), a matrix from
 * hmg_assemble_coarse is re-assembled and re-factorised first; one from hmg_set_coarse_matrix makes the call fail
 * (the caller owns that matrix). */
int hmg_vcycle(hmg_ctx* ctx, int top_level, int steps, double* out_resnorm);
/* `ncycles` V-cycles back to back without host synchronisation (benchmark / batch use);
 * resnorms[ncycles] may be NULL. */
int hmg_vcycles(hmg_ctx* ctx, int top_level, int steps, int ncycles, double* resnorms);

/* driver integrals evaluated on the device (next row N1 of SURVEY.md 8f):
 * rhs_a_xi_grad_v!, integrate_first_term, integrate_terms, integrate_area, next_rhs!
 * (src/examples/homogenized_coefficients.jl:449-474, 592-632, 634-667, 673-689, 695-713).
 * nsubset = length of the element prefix to integrate over. */
int hmg_rhs_axi_grad(hmg_ctx* ctx, const double* xi, int which_b);
int hmg_integrate_first_term(hmg_ctx* ctx, int which_v, const double* xi, int64_t nsubset, double* out);
int hmg_integrate_terms(hmg_ctx* ctx, int which_vk, int which_vkm1, int64_t nsubset, double* out);
int hmg_integrate_area(hmg_ctx* ctx, int64_t nsubset, double* out);
int hmg_next_rhs(hmg_ctx* ctx, int which_b, int which_x);

/* generate_field(ns, T, threads, alpha, p) (tools/generate_st1_field.jl:86-120): white Gaussian
 * noise -> real FFT -> division by (1 + |k|)^p (:41-84) -> inverse real FFT -> exp(alpha |G|), on the device with cuFFT.
 * n[dim] even extents; the field comes back as n[0] x n[1] (x n[2]) doubles, last index fastest (the transpose of the
 * Julia array).  normalize != 0: G is scaled to unit (population) standard deviation before the exponential (the inputs
 * of BASELINE.json configs[4]; 0 = the tool as written).  noise_or_null: the caller's standard-normal samples in the same
 * layout, or NULL to draw them on the device (Philox4x32-10, key = seed, counter = cell index, Box-Muller).  No context
 * is needed. */
int hmg_generate_field(int dim, const int* n, uint64_t seed, double alpha, double p, int normalize,
                       const double* noise_or_null, double* out_host, int device);

/* refined_mesh(implicit, level) (src/implicit_fine_grid.jl:24): the refined reference element that construct_full_grid
 * (src/implicit_fine_grid.jl:41-78) maps into every coarse element for the VTK export.  nodes[dim x Nf(level)] reference
 * coordinates in hierarchical row order, elems1[(dim+1) x nel] 1-based, index-sorted per element; arrays may be NULL,
 * *nel is always written.  Host data only. */
int hmg_refined_mesh(const hmg_ctx* ctx, int level, double* nodes, int64_t* elems1, int64_t* nel);

/* stream control / measurement */
int hmg_synchronize(hmg_ctx* ctx);
/* device time in milliseconds of `reps` repetitions of an operation, measured with CUDA events
 * on the context's stream: op 0 = hmg_apply_global(top, P -> AP), 1 = hmg_vcycle(top, steps),
 * 2 = hmg_mul(top, 1.0, P, AP), 3 = local apply AP = constraint(A P), 4 = interface sum of AP,
 * 5 = local residual, 6 / 7 / 8 = the fused CG vector kernels (x,r update; p update; p = r with
 * rho), 9 = restriction level -> level-1, 10 = interpolation level-1 -> level, 11 = local apply with
 * the fused owner-weighted dot, 12 = the fused direction update + product of a CG step (p' = R + beta P
 * formed inside the apply kernel, AP = broadcast(constraint(A p'))), 13 / 14 = the interface kernel restricted to the
 * two-owner cells (faces in 3D, edges in 2D) / to the cells with more owners (local cells only), 16 = x += alpha P, 17 = dot(P, AP) summed over
 * the ranks (the pattern of every CG scalar), 18 = the cut-cell exchange of AP alone (partitioned contexts).  The
 * operation is
 * launched `reps` times back to back. */
int hmg_time_op(hmg_ctx* ctx, int op, int level, int steps, int reps, float* ms_out);
/* number of kernel launches issued on the context's stream since creation */
int64_t hmg_launch_count(const hmg_ctx* ctx);
/* raw device pointer of a state vector (element-interleaved, lattice row order; see hmg_group_width).  The pointer of
 * HMG_P is invalidated by every hmg_smoothing_steps / hmg_vcycle(s) call (the fused direction update swaps the p buffer
 * with an internal one): fetch it again after such a call. */
void* hmg_device_ptr(hmg_ctx* ctx, int level, int which);
/* permutation: lattice position (0-based) of hierarchical row i (0-based) at `level` */
int hmg_hier_to_lattice(const hmg_ctx* ctx, int level, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* HMG_B200_H */
