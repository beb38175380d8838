/*
 * csr_matvec.c — plain-C restatement of scipy.sparse._sparsetools csr_matvec, the native
 * loop behind `self.matrix @ data` (reference filters/fields/regrid.py:309-310).
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): the checker and the multi-threaded CPU
 * baseline of bench.py; never linked into the product.
 *
 * scipy's algorithm (published in scipy/sparse/sparsetools/csr.h, csr_matvec):
 *     for i in rows:  sum = y[i];  for jj in [Ap[i], Ap[i+1]):  sum += Ax[jj] * Xx[Aj[jj]];  y[i] = sum;
 * with y zero-initialised by the Python caller.  Accumulation is sequential in storage
 * order; scipy wheels target baseline x86-64, so the multiply and the add are rounded
 * separately — compile this file with -ffp-contract=off to keep that.
 * Pinned bit for bit against scipy in tests/test_oracle_spmm.py.
 */
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void csr_matvec_f32(int64_t n_row, const int32_t* Ap, const int32_t* Aj, const float* Ax, const float* Xx,
                           float* Yx) {
    for (int64_t i = 0; i < n_row; i++) {
        float sum = 0.0f;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) sum += Ax[jj] * Xx[Aj[jj]];
        Yx[i] = sum;
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Field-major batches: X[n_fields][n_src] -> Y[n_fields][n_tgt], one csr_matvec per field
 * (the reference's per-field loop, regrid.py:204-208), fields spread over threads. */
void oracle_csr_matvecs_f32(int64_t n_row, const int32_t* Ap, const int32_t* Aj, const float* Ax, const float* X,
                            float* Y, int64_t n_fields, int64_t n_src, int64_t n_tgt, int n_threads) {
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int64_t f = 0; f < n_fields; f++) csr_matvec_f32(n_row, Ap, Aj, Ax, X + f * n_src, Y + f * n_tgt);
}
