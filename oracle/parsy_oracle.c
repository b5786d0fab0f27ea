/* oracle/parsy_oracle.c — CPU restatement of ParSy's numeric hot path in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
 * call this; the product (parsy_bench_b200/) never does and has no CPU path.
 *
 * Every function restates one reference routine (file:line given) with scalar loops.  Where the reference calls
 * vendor BLAS/LAPACK (MKL in the original, OpenBLAS in oracle/_ref: dsyrk/dgemm/dpotrf/dtrsm at
 * cholesky/parallel_PB_Cholesky_05.h:160,173,204,218) the arithmetic restated here is the reference's own scalar
 * alternative (cholesky/MyBLAS.h:10-35) and textbook SYRK/GEMM.  Parity of this file is PINNED by
 * tests/test_oracle.py: against the committed golden vectors produced by the compiled reference
 * (tests/golden/, generator tests/golden/make_golden.py) and, where oracle/_ref exists, against the reference run
 * live on the same inputs.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* common/Reach.h:112-143 — descendants of the supernode holding columns [col1,col2) on the supernodal etree.
 * s/w are work arrays of n ints (w all zero on entry and on exit); returns top, result in s[top..n). */
#define OR_FLIP(i) (-(i)-2)
#define OR_MARKED(w, j) ((w)[j] < 0)
#define OR_MARK(w, j) { (w)[j] = OR_FLIP((w)[j]); }
int oracle_ereach_sn(int n, const int* Ap, const int* Ai, int col1, int col2, const int* col2sup, const int* parent,
                     int* s, int* w) {
  int top = n;
  for (int k = col1; k < col2; ++k) {
    if (k == col1) OR_MARK(w, col2sup[k]);
    for (int p = Ap[k]; p < Ap[k + 1]; ++p) {
      int i = col2sup[Ai[p]];
      int len = 0;
      if (Ai[p] > k) continue; /* only the upper triangular part */
      for (; !OR_MARKED(w, i); i = parent[i]) { s[len++] = i; OR_MARK(w, i); }
      while (len > 0) s[--top] = s[--len];
    }
  }
  for (int p = top; p < n; ++p) OR_MARK(w, s[p]);
  OR_MARK(w, col2sup[col1]);
  return top;
}

/* cholesky/MyBLAS.h:10-25 — column Cholesky of the dim x dim leading block of a panel with leading dimension n.
 * Returns 0, or j+1 if pivot j is not positive (LAPACK dpotrf info convention, parallel_PB_Cholesky_05.h:204-207). */
static int cholesky_col(int n, int dim, double* a) {
  for (int j = 0; j < dim; ++j) {
    for (int k = 0; k < j; ++k) {
      const double t = a[(size_t)k * n + j];
      for (int i = j; i < dim; ++i) a[(size_t)j * n + i] -= a[(size_t)k * n + i] * t;
    }
    if (!(a[(size_t)j * n + j] > 0.0)) return j + 1;
    const double d = sqrt(a[(size_t)j * n + j]);
    for (int k = j + 1; k < dim; ++k) a[(size_t)j * n + k] /= d;
    a[(size_t)j * n + j] = d;
  }
  return 0;
}

/* cholesky/MyBLAS.h:27-35 — one row of  X * L11' = A21  (rhs strided by the panel's leading dimension) */
static void lsolve_dense_col(int colSize, int col, const double* M, double* rhs) {
  for (int i = 0; i < col; ++i) {
    rhs[(size_t)i * colSize] /= M[(size_t)i * colSize + i];
    for (int j = i + 1; j < col; ++j) rhs[(size_t)j * colSize] -= M[(size_t)i * colSize + j] * rhs[(size_t)i * colSize];
  }
}

/* One supernode of cholesky/parallel_PB_Cholesky_05.h:96-219 (s is 0-based here).  map: n ints, contribs:
 * >= maxCol*maxSupWid doubles, xi: 2*supNo ints (zero).  Returns 0 or the failing global column + 1. */
static int factor_supernode(int s, const int* c, const int* r, const double* values, const size_t* lC, const int* lR,
                            const size_t* Li_ptr, double* lValues, const int* blockSet, int supNo, const int* aTree,
                            const int* cT, const int* rT, const int* col2Sup, int* map, double* contribs, int* xi) {
  const int curCol = blockSet[s], nxtCol = blockSet[s + 1];
  const int supWdt = nxtCol - curCol;
  const int nSupR = (int)(Li_ptr[nxtCol] - Li_ptr[curCol]);
  int cnt = 0;
  for (size_t i = Li_ptr[curCol]; i < Li_ptr[nxtCol]; ++i) map[lR[i]] = cnt++;   /* :100-102 */
  for (int i = curCol; i < nxtCol; ++i)                                           /* :104-112 */
    for (int j = c[i]; j < c[i + 1]; ++j) lValues[lC[i] + map[r[j]]] = values[j];
  double* cur = &lValues[lC[curCol]];
  const int top = oracle_ereach_sn(supNo, cT, rT, curCol, nxtCol, col2Sup, aTree, xi, xi + supNo);   /* :115 */
  for (int q = top; q < supNo; ++q) {                                                                 /* :117 */
    const int lSN = xi[q];
    const int cSN = blockSet[lSN], cNSN = blockSet[lSN + 1];
    const size_t p0 = Li_ptr[cSN], p1 = Li_ptr[cNSN];
    const int nSNRCur = (int)(p1 - p0), supWdts = cNSN - cSN;
    int lb = 0, ub = 0, sw = 1;
    for (size_t j = p0; j < p1; ++j) {                                                                /* :137-149 */
      if (lR[j] >= curCol && sw) { lb = (int)(j - p0); sw = 0; }
      if (lR[j] < curCol + supWdt && !sw) ub = (int)(j - p0);
      if (lR[j] >= curCol + supWdt) break;
    }
    const int nSupRs = nSNRCur - lb, ndrow1 = ub - lb + 1;
    const double* src = &lValues[lC[cSN] + lb];
    /* dsyrk("L","N") :160 and dgemm("N","C") :173 — contribs(j,i) = sum_k src(j,k) src(i,k), j >= i, ld nSupRs */
    for (int i = 0; i < ndrow1; ++i)
      for (int j = i; j < nSupRs; ++j) {
        double acc = 0.0;
        for (int k = 0; k < supWdts; ++k) acc += src[(size_t)k * nSNRCur + j] * src[(size_t)k * nSNRCur + i];
        contribs[(size_t)i * nSupRs + j] = acc;
      }
    for (int i = 0; i < ndrow1; ++i) {                                                                /* :190-197 */
      const int col = map[lR[p0 + i + lb]];
      for (int j = i; j < nSupRs; ++j) {
        const int cRow = lR[p0 + j + lb];
        cur[(size_t)col * nSupR + map[cRow]] -= contribs[(size_t)i * nSupRs + j];
      }
    }
  }
  const int info = cholesky_col(nSupR, supWdt, cur);                                                  /* :204 */
  if (info) return curCol + info;
  for (int i = supWdt; i < nSupR; ++i) lsolve_dense_col(nSupR, supWdt, cur, &cur[i]);                  /* :218 */
  return 0;
}

/* cholesky/parallel_PB_Cholesky_05.h:27-425 run by ONE thread: H-levels in order, w-partitions in order,
 * supernodes of a partition in list order.  lValues must be zeroed by the caller (choleskyTest01.cpp:202).
 * Returns 1 on success, 0 on a non-positive pivot. */
int oracle_cholesky_left_par_05(int n, const int* c, const int* r, const double* values, const size_t* lC,
                                const int* lR, const size_t* Li_ptr, double* lValues, const int* blockSet, int supNo,
                                const int* aTree, const int* cT, const int* rT, const int* col2Sup, int nLevels,
                                const int* levelPtr, const int* parPtr, const int* partition, int super_max,
                                int col_max) {
  int* map = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
  int* xi = (int*)calloc((size_t)(2 * supNo + 2), sizeof(int));
  double* contribs = (double*)calloc((size_t)super_max * (size_t)col_max + 1, sizeof(double));
  int ok = 1;
  for (int l = 0; l < nLevels && ok; ++l)
    for (int j1 = levelPtr[l]; j1 < levelPtr[l + 1] && ok; ++j1)
      for (int k1 = parPtr[j1]; k1 < parPtr[j1 + 1]; ++k1)
        if (factor_supernode(partition[k1], c, r, values, lC, lR, Li_ptr, lValues, blockSet, supNo, aTree, cT, rT,
                             col2Sup, map, contribs, xi)) { ok = 0; break; }
  free(map); free(xi); free(contribs);
  return ok;
}

/* cholesky/PB_Cholesky.h:16-154 — serial twin: supernodes 0..supNo-1 in index order */
int oracle_cholesky_left_sn(int n, const int* c, const int* r, const double* values, const size_t* lC, const int* lR,
                            const size_t* Li_ptr, double* lValues, const int* blockSet, int supNo, const int* aTree,
                            const int* cT, const int* rT, const int* col2Sup, int super_max, int col_max) {
  int* map = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
  int* xi = (int*)calloc((size_t)(2 * supNo + 2), sizeof(int));
  double* contribs = (double*)calloc((size_t)super_max * (size_t)col_max + 1, sizeof(double));
  int ok = 1;
  for (int s = 0; s < supNo; ++s)
    if (factor_supernode(s, c, r, values, lC, lR, Li_ptr, lValues, blockSet, supNo, aTree, cT, rT, col2Sup, map,
                         contribs, xi)) { ok = 0; break; }
  free(map); free(xi); free(contribs);
  return ok;
}

/* triangularSolve/BLAS.h:8-103 — dense lower non-unit solve of the ncol x ncol leading block, column oriented
 * (the 8/4/2/1 unrolling of the reference only regroups the same operations) */
static void dlsolve_nonunit(int ldm, int ncol, const double* M, double* rhs) {
  for (int j = 0; j < ncol; ++j) {
    const double xj = rhs[j] / M[(size_t)j * ldm + j];
    rhs[j] = xj;
    for (int k = j + 1; k < ncol; ++k) rhs[k] -= xj * M[(size_t)j * ldm + k];
  }
}
/* triangularSolve/BLAS.h:119-191 — Mxvec += M * vec */
static void dmatvec(int ldm, int nrow, int ncol, const double* M, const double* vec, double* Mxvec) {
  for (int j = 0; j < ncol; ++j) {
    const double v = vec[j];
    for (int k = 0; k < nrow; ++k) Mxvec[k] += v * M[(size_t)j * ldm + k];
  }
}
static void solve_supernode_fwd(int i, const size_t* Lp, const int* Li, const double* Lx, const size_t* Li_ptr,
                                const int* sup2col, double* x, double* tempVec) {
  const int curCol = sup2col[i], nxtCol = sup2col[i + 1], supWdt = nxtCol - curCol;
  const int nSupR = (int)(Li_ptr[nxtCol] - Li_ptr[curCol]);
  dlsolve_nonunit(nSupR, supWdt, &Lx[Lp[curCol]], &x[curCol]);                       /* Triangular_BCSC.h:32 */
  dmatvec(nSupR, nSupR - supWdt, supWdt, &Lx[Lp[curCol] + supWdt], &x[curCol], tempVec);   /* :35 */
  size_t l = Li_ptr[curCol] + supWdt;
  for (int k = 0; l < Li_ptr[nxtCol]; ++l, ++k) { x[Li[l]] -= tempVec[k]; tempVec[k] = 0; }   /* :36-39 */
}

/* triangularSolve/Triangular_BCSC.h:14-49 */
int oracle_blockedLsolve(int n, const size_t* Lp, const int* Li, const double* Lx, const size_t* Li_ptr,
                         const int* sup2col, int supNo, double* x) {
  if (!Lp || !Li || !x) return 0;
  double* tempVec = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  for (int i = 0; i < supNo; ++i) solve_supernode_fwd(i, Lp, Li, Lx, Li_ptr, sup2col, x, tempVec);
  free(tempVec);
  return 1;
}
/* triangularSolve/Triangular_BCSC.h:171-232 (and :115-164, :238-348, which only differ in the schedule handed
 * in): levels in order, w-partitions in order, supernodes in list order — one thread */
int oracle_H2LeveledBlockedLsolve(int n, const size_t* Lp, const int* Li, const double* Lx, const size_t* Li_ptr,
                                  const int* sup2col, int supNo, double* x, int levels, const int* levelPtr,
                                  const int* parPtr, const int* partition) {
  (void)supNo;
  if (!Lp || !Li || !x) return 0;
  double* tempVec = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  for (int l = 0; l < levels; ++l)
    for (int j1 = levelPtr[l]; j1 < levelPtr[l + 1]; ++j1)
      for (int k1 = parPtr[j1]; k1 < parPtr[j1 + 1]; ++k1)
        solve_supernode_fwd(partition[k1], Lp, Li, Lx, Li_ptr, sup2col, x, tempVec);
  free(tempVec);
  return 1;
}
/* NEW (not in the reference): backward sweep L' x = b, supernodes in reverse order:
 * x_s -= L21' x[rows]; x_s <- L11^-T x_s.  Parity of this routine is pinned only by the residual test. */
int oracle_blockedLtsolve(int n, const size_t* Lp, const int* Li, const double* Lx, const size_t* Li_ptr,
                          const int* sup2col, int supNo, double* x) {
  (void)n;
  if (!Lp || !Li || !x) return 0;
  for (int i = supNo - 1; i >= 0; --i) {
    const int curCol = sup2col[i], nxtCol = sup2col[i + 1], supWdt = nxtCol - curCol;
    const int nSupR = (int)(Li_ptr[nxtCol] - Li_ptr[curCol]);
    const double* M = &Lx[Lp[curCol]];
    const int* rows = &Li[Li_ptr[curCol]];
    for (int j = supWdt - 1; j >= 0; --j) {
      double acc = x[curCol + j];
      for (int k = supWdt; k < nSupR; ++k) acc -= M[(size_t)j * nSupR + k] * x[rows[k]];
      for (int k = j + 1; k < supWdt; ++k) acc -= M[(size_t)j * nSupR + k] * x[curCol + k];
      x[curCol + j] = acc / M[(size_t)j * nSupR + j];
    }
  }
  return 1;
}
/* triangularSolve/Triangular_CSC.h:14-27 */
int oracle_lsolve(int n, const int* Lp, const int* Li, const double* Lx, double* x) {
  if (!Lp || !Li || !x) return 0;
  for (int j = 0; j < n; ++j) {
    x[j] /= Lx[Lp[j]];
    for (int p = Lp[j] + 1; p < Lp[j + 1]; ++p) x[Li[p]] -= Lx[p] * x[j];
  }
  return 1;
}
/* common/Util.h:277-288 — b = L * 1 */
void oracle_rhsInitBlocked(size_t n, const size_t* Ap, const int* Ai, const size_t* AiP, const double* Ax, double* b) {
  for (size_t j = 0; j < n; ++j) b[j] = 0;
  for (size_t c = 0; c < n; ++c) {
    size_t j = 0;
    for (size_t cc = Ap[c]; cc < Ap[c + 1]; ++cc, ++j) b[Ai[AiP[c] + j]] += Ax[cc];
  }
}
/* common/Util.h:294-306 — one-sided check 1 - x[i] < 0.001 for every i */
int oracle_testTriangular(size_t n, const double* x) {
  size_t test = 0;
  for (size_t i = 0; i < n; ++i) if (1 - x[i] < 0.001) test++;
  return n - test > 0 ? 0 : 1;
}

/* ---- full system A x = b around the sweeps (test oracle for parsy_cuda_solve_system) -----------------------------
 * The reference has no such routine: its driver only sketches the right-hand side b_i = 1 + i/n and a CHOLMOD solve
 * (examples/choleskyTest01.cpp:408-432).  res = b - A x for the symmetric A whose lower half is stored by columns
 * (c, r, values as handed to cholesky_left_par_05, parallel_PB_Cholesky_05.h:27-28). */
void oracle_residual_sym_lower(int n, const int* c, const int* r, const double* values, const double* x,
                               const double* b, double* res) {
  for (int j = 0; j < n; ++j) res[j] = b[j];
  for (int j = 0; j < n; ++j)
    for (int p = c[j]; p < c[j + 1]; ++p) {
      const int i = r[p];
      res[i] -= values[p] * x[j];
      if (i != j) res[j] -= values[p] * x[i];
    }
}
