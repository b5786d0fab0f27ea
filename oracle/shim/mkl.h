/* Shim for <mkl.h>: forwards the five BLAS/LAPACK entry points the reference's executor calls
 * (parallel_PB_Cholesky_05.h:160,173,204,218; Triangular_BCSC.h:203,212) to the Fortran symbols of
 * the OpenBLAS bundled in the image. TEST/BASELINE INFRASTRUCTURE ONLY (oracle/_ref build). */
#ifndef PARSY_ORACLE_MKL_SHIM_H
#define PARSY_ORACLE_MKL_SHIM_H
typedef int MKL_INT;
#define MKL_DOMAIN_BLAS 1
extern "C" {
void dsyrk_(const char*, const char*, const int*, const int*, const double*, const double*, const int*,
            const double*, double*, const int*);
void dgemm_(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
            const int*, const double*, const int*, const double*, double*, const int*);
void dpotrf_(const char*, const int*, double*, const int*, int*);
void dtrsm_(const char*, const char*, const char*, const char*, const int*, const int*, const double*,
            const double*, const int*, double*, const int*);
void dgemv_(const char*, const int*, const int*, const double*, const double*, const int*, const double*,
            const int*, const double*, double*, const int*);
void openblas_set_num_threads(int);
}
static inline void dsyrk(const char* u, const char* t, const int* n, const int* k, const double* al,
                         const double* a, const int* lda, const double* be, double* c, const int* ldc) {
  dsyrk_(u, t, n, k, al, a, lda, be, c, ldc);
}
static inline void dgemm(const char* ta, const char* tb, const int* m, const int* n, const int* k,
                         const double* al, const double* a, const int* lda, const double* b, const int* ldb,
                         const double* be, double* c, const int* ldc) {
  dgemm_(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}
static inline void dpotrf(const char* u, const int* n, double* a, const int* lda, int* info) {
  dpotrf_(u, n, a, lda, info);
}
static inline void dtrsm(const char* s, const char* u, const char* t, const char* d, const int* m, const int* n,
                         const double* al, const double* a, const int* lda, double* b, const int* ldb) {
  dtrsm_(s, u, t, d, m, n, al, a, lda, b, ldb);
}
static inline void dgemv(const char* t, const int* m, const int* n, const double* al, const double* a,
                         const int* lda, const double* x, const int* incx, const double* be, double* y,
                         const int* incy) {
  dgemv_(t, m, n, al, a, lda, x, incx, be, y, incy);
}
static inline void MKL_Domain_Set_Num_Threads(int nt, int) { openblas_set_num_threads(nt); }
static inline void MKL_Set_Num_Threads(int nt) { openblas_set_num_threads(nt); }
#endif
