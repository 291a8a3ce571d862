/*
 * hmg_ref_cpu.c -- CPU restatement of the reference's hot loops.  TEST INFRASTRUCTURE ONLY:
 * it is the checker's twin and the CPU baseline that bench.py times beside the GPU numbers
 * ("restatement, not Julia": Julia is not installed in this image).  Nothing in the product
 * links or calls this file.
 *
 * Loop structure follows the reference (paths relative to the reference repository root):
 *   ref_mul                  src/apply_local_operators.jl:85-133  (dim^2 + 1 CSC scatter-SpMVs per
 *                            coarse element, cyclic distribution of elements over threads)
 *   ref_cell_sum / _zero     src/implicit_fine_grid.jl:209-328, 94-139, 334-386 (serial)
 *   ref_interpolate / _restrict  src/interpolation.jl:52-74 (cyclic over columns)
 *   ref_dot / ref_axpy / ref_xpby  BLAS-1 passes of src/multigrid.jl:54,64-68
 * Indices are 0-based int64 (the reference uses 1-based Int64).
 */
#include <stdint.h>
#include <string.h>
#include <omp.h>

typedef int64_t i64;

/* y[:, off] += alpha * A * x[:, off]   -- my_A_mul_B!, src/apply_local_operators.jl:125-133 */
static inline void a_mul_b(double alpha, i64 n, const i64* colptr, const i64* rowval, const double* nzval,
                           const double* x, double* y) {
    for (i64 j = 0; j < n; ++j) {
        const double axj = alpha * x[j];
        for (i64 q = colptr[j]; q < colptr[j + 1]; ++q) y[rowval[q]] += nzval[q] * axj;
    }
}

/* mul!(alpha, base, A::L2PlusDivAGrad, x, y).  P: ne x dim x dim (row-major per element), det: ne,
 * ops: dim*dim CSC matrices given as arrays of pointers (k*dim + l). */
void ref_mul(double alpha, int dim, i64 ne, i64 nf, const double* P, const double* det, double lambda,
             const i64* const* op_colptr, const i64* const* op_rowval, const double* const* op_nzval,
             const i64* m_colptr, const i64* m_rowval, const double* m_nzval, const double* x, double* y,
             int nthreads) {
#pragma omp parallel num_threads(nthreads)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        for (i64 e = t; e < ne; e += nt) {                 /* thread_id : nthreads : nelements */
            const double detJ = det[e];
            const double* Pe = P + e * dim * dim;
            const double* xe = x + e * nf;
            double* ye = y + e * nf;
            for (int i = 0; i < dim; ++i)
                for (int j = 0; j < dim; ++j)
                    a_mul_b(alpha * detJ * Pe[i * dim + j], nf, op_colptr[i * dim + j], op_rowval[i * dim + j],
                            op_nzval[i * dim + j], xe, ye);
            if (alpha * lambda * detJ != 0.0) a_mul_b(alpha * lambda * detJ, nf, m_colptr, m_rowval, m_nzval, xe, ye);
        }
    }
}

/* One cell family (faces, edges or nodes) of broadcast_interfaces!: rows[local_id][k] are the local
 * node lists (npc nodes each).  mode 0: sum and broadcast; 1: zero every owner (constraint);
 * 2: zero all but the first owner.  Serial, like the reference. */
void ref_cells(int mode, i64 ncells, const i64* offset, const i64* element, const i64* local_id, i64 npc,
               const i64* rows, i64 nf, double* x, double* buffer) {
    for (i64 c = 0; c < ncells; ++c) {
        if (mode == 0) {
            memset(buffer, 0, sizeof(double) * npc);
            for (i64 o = offset[c]; o < offset[c + 1]; ++o) {
                const i64* nodes = rows + local_id[o] * npc;
                const double* col = x + element[o] * nf;
                for (i64 k = 0; k < npc; ++k) buffer[k] += col[nodes[k]];
            }
            for (i64 o = offset[c]; o < offset[c + 1]; ++o) {
                const i64* nodes = rows + local_id[o] * npc;
                double* col = x + element[o] * nf;
                for (i64 k = 0; k < npc; ++k) col[nodes[k]] = buffer[k];
            }
        } else {
            for (i64 o = offset[c] + (mode == 2 ? 1 : 0); o < offset[c + 1]; ++o) {
                const i64* nodes = rows + local_id[o] * npc;
                double* col = x + element[o] * nf;
                for (i64 k = 0; k < npc; ++k) col[nodes[k]] = 0.0;
            }
        }
    }
}

/* y[:, c] += P x[:, c]  (CSC scatter), cyclic over columns -- src/interpolation.jl:52-62 */
void ref_interpolate(i64 ne, i64 nff, i64 nfc, const i64* colptr, const i64* rowval, const double* nzval,
                     const double* xc, double* yf, int nthreads) {
#pragma omp parallel num_threads(nthreads)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        for (i64 e = t; e < ne; e += nt) a_mul_b(1.0, nfc, colptr, rowval, nzval, xc + e * nfc, yf + e * nff);
    }
}

/* y[:, c] = P' x[:, c] -- src/interpolation.jl:64-74 */
void ref_restrict(i64 ne, i64 nff, i64 nfc, const i64* colptr, const i64* rowval, const double* nzval,
                  const double* xf, double* yc, int nthreads) {
#pragma omp parallel num_threads(nthreads)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        for (i64 e = t; e < ne; e += nt) {
            const double* x = xf + e * nff;
            double* y = yc + e * nfc;
            for (i64 j = 0; j < nfc; ++j) {
                double s = 0.0;
                for (i64 q = colptr[j]; q < colptr[j + 1]; ++q) s += nzval[q] * x[rowval[q]];
                y[j] = s;
            }
        }
    }
}

double ref_dot(i64 n, const double* a, const double* b, int nthreads) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) num_threads(nthreads) schedule(static)
    for (i64 i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
void ref_axpy(i64 n, double alpha, const double* x, double* y, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (i64 i = 0; i < n; ++i) y[i] += alpha * x[i];
}
/* p = r + beta p : single-threaded broadcast in the reference (src/multigrid.jl:68) */
void ref_xpby(i64 n, const double* r, double beta, double* p) {
    for (i64 i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
}
void ref_copy(i64 n, const double* src, double* dst, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (i64 i = 0; i < n; ++i) dst[i] = src[i];
}
void ref_fill(i64 n, double v, double* dst, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (i64 i = 0; i < n; ++i) dst[i] = v;
}
int ref_max_threads(void) { return omp_get_max_threads(); }
