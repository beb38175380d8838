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

/* ---- GRIB-backed fields: decode + matvec per field (CPU baseline of bench.py's e2e_grib) ----
 * What RegridFilter.forward costs the reference on a GRIB FieldList (regrid.py:309-310): per
 * field `to_numpy(flatten=True)` — 16-bit simple packing decoded to float64 in one pass,
 * ((X * s) + R) * d as in oracle/grib.py — then csr_matvec in float64 (scipy widens a float32
 * matrix exactly when the vector is float64).  Fields spread over threads, one scratch vector
 * per thread. */
#include <stdlib.h>

static void csr_matvec_f32w_f64(int64_t n_row, const int32_t* Ap, const int32_t* Aj, const float* Ax, const double* Xx,
                                double* Yx) {
    for (int64_t i = 0; i < n_row; i++) {
        double sum = 0.0;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; jj++) sum += (double)Ax[jj] * Xx[Aj[jj]];
        Yx[i] = sum;
    }
}

void oracle_grib16_regrid_f64(int64_t n_row, const int32_t* Ap, const int32_t* Aj, const float* Ax,
                              const uint8_t* const* packed, const double* R, const double* s, const double* d, double* Y,
                              int64_t n_fields, int64_t n_src, int64_t n_tgt, int n_threads) {
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel num_threads(n_threads)
#endif
    {
        double* x = (double*)malloc((size_t)n_src * sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int64_t f = 0; f < n_fields; f++) {
            const uint8_t* p = packed[f];
            const double r = R[f], sf = s[f], df = d[f];
            for (int64_t i = 0; i < n_src; i++) {
                const double v = (double)(((unsigned)p[2 * i] << 8) | p[2 * i + 1]);
                x[i] = ((v * sf) + r) * df;
            }
            csr_matvec_f32w_f64(n_row, Ap, Aj, Ax, x, Y + f * n_tgt);
        }
        free(x);
    }
}
