// oracle/_ref driver — TEST / BASELINE INFRASTRUCTURE ONLY, never on the product path.
//
// Compiles the UNMODIFIED reference headers (from a temp copy that only adds the eight missing
// `return` statements) into one binary that
//   * generates the synthetic Laplacian of BASELINE.json in memory (lower-half CSC, diagonal first),
//   * runs the reference inspector  analyze_p2            (cholesky/LSparsity.h:256),
//   * runs the reference executor   cholesky_left_par_05  (cholesky/parallel_PB_Cholesky_05.h:27),
//   * runs the reference solves     blockedLsolve / leveledBlockedLsolve / H2LeveledBlockedLsolve[_Peeled]
//                                   (triangularSolve/Triangular_BCSC.h:14,115,171,238) and the CSC
//                                   variants lsolve / lsolvePar / lsolveParH2 (Triangular_CSC.h:14,50,76),
//   * dumps every symbolic / schedule / numeric array as raw little-endian files for the parity tests,
//   * prints one JSON line with sizes and timings (the CPU baseline of bench.py).
// Call-site conventions follow examples/choleskyTest01.cpp:123-222 and examples/triangularTest02.cpp:86-266.
#include <iostream>
#include <fstream>
#include <chrono>
#include <algorithm>
#include <climits>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <cmath>
#include <omp.h>
#include "Ordering.h"
#include "Inspection_Prune.h"
#include "Inspection_Block.h"
#include "Util.h"
#include "PB_Cholesky.h"
#include "LSparsity.h"
#include "mkl.h"
#include "parallel_PB_Cholesky_05.h"
#include "Triangular_BCSC.h"
#include "Triangular_CSC.h"
#include "Inspection_Level.h"
#include "DFS.h"
#include "InspectionDAG_03.h"
#include "Inspection_DAG_02.h"
#ifdef PARSY_GPU_FORWARD
// The same driver with the reference's call sites forwarded to libparsy_cuda (include/parsy_cuda_dropin.h, the header
// INTEGRATION.md gives to a maintainer): the reference's inspector and harness drive the CUDA executor — parsy_ref_gpu.
#include "parsy_cuda_dropin.h"
#endif

static double now() {
  return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
}

// kind: 0 = 2D 5-point, 1 = 3D 7-point, 2 = 3D 27-point; Dirichlet, SURVEY.md §8(d)
static void gen_laplacian(int kind, int N, std::vector<int>& p, std::vector<int>& i, std::vector<double>& x) {
  long n = kind == 0 ? (long)N * N : (long)N * N * N;
  p.assign(n + 1, 0);
  i.clear(); x.clear();
  int NZ = kind == 0 ? 1 : N;
  for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < N; ++y)
      for (int xx = 0; xx < N; ++xx) {
        long v = ((long)z * N + y) * N + xx;
        double diag = kind == 0 ? 4.0 : (kind == 1 ? 6.0 : 26.0);
        i.push_back((int)v); x.push_back(diag);
        // lower neighbours (index > v) in ascending order
        for (int dz = 0; dz <= (kind == 0 ? 0 : 1); ++dz)
          for (int dy = (dz == 0 ? 0 : -1); dy <= 1; ++dy)
            for (int dx = ((dz == 0 && dy == 0) ? 1 : -1); dx <= 1; ++dx) {
              int nz = z + dz, ny = y + dy, nx = xx + dx;
              if (nz < 0 || nz >= NZ || ny < 0 || ny >= N || nx < 0 || nx >= N) continue;
              int manh = std::abs(dx) + std::abs(dy) + std::abs(dz);
              if (kind != 2 && manh != 1) continue;
              long u = ((long)nz * N + ny) * N + nx;
              i.push_back((int)u); x.push_back(-1.0);
            }
        p[v + 1] = (int)i.size();
      }
}

template <typename T> static void dump(const std::string& dir, const char* name, const T* a, size_t cnt) {
  if (dir.empty()) return;
  std::string f = dir + "/" + name;
  FILE* fp = fopen(f.c_str(), "wb");
  if (!fp) { perror(f.c_str()); exit(2); }
  if (cnt) fwrite(a, sizeof(T), cnt, fp);
  fclose(fp);
}

int main(int argc, char** argv) {
  int kind = 0, N = 100, costParam = 8, levelParam = 1, divRate = 2, threads = 1, blasThreads = -1, iters = 1;
  int doFactor = 1, doSolve = 1, dumpL = 1, chunk = 1, triOnly = 0, doCsc = 1;
  std::string dir, mtx;
  for (int a = 1; a < argc; ++a) {
    std::string s = argv[a];
    auto nxt = [&]() { return std::string(argv[++a]); };
    if (s == "--kind") { std::string k = nxt(); kind = k == "2d5" ? 0 : k == "3d7" ? 1 : 2; }
    else if (s == "--N") N = atoi(nxt().c_str());
    else if (s == "--cost") costParam = atoi(nxt().c_str());
    else if (s == "--level") levelParam = atoi(nxt().c_str());
    else if (s == "--div") divRate = atoi(nxt().c_str());
    else if (s == "--threads") threads = atoi(nxt().c_str());
    else if (s == "--blas-threads") blasThreads = atoi(nxt().c_str());
    else if (s == "--iters") iters = atoi(nxt().c_str());
    else if (s == "--dump") dir = nxt();
    else if (s == "--no-factor") doFactor = 0;
    else if (s == "--no-solve") doSolve = 0;
    else if (s == "--no-dump-values") dumpL = 0;
    else if (s == "--mtx") mtx = nxt();
    else if (s == "--tri-only") triOnly = 1;
    else if (s == "--no-csc") doCsc = 0;   // skip bcsc2csc + lsolve (full-size configs: saves nnz(L) * 12 B of dump)
    else { fprintf(stderr, "unknown arg %s\n", s.c_str()); return 2; }
  }
  if (blasThreads < 0) blasThreads = threads;
  omp_set_num_threads(threads);
  openblas_set_num_threads(1);

  std::vector<int> Ap, Ai; std::vector<double> Ax;
  if (!mtx.empty()) {
    // Matrix-Market input through the reference's own reader (common/Util.h:77 readMatrix)
    size_t rn = 0, rnnz = 0; int *rc = NULL, *rr = NULL; double* rv = NULL;
    if (!readMatrix(mtx, rn, rnnz, rc, rr, rv)) { fprintf(stderr, "readMatrix failed for %s\n", mtx.c_str()); return 3; }
    Ap.assign(rc, rc + rn + 1); Ai.assign(rr, rr + rnnz); Ax.assign(rv, rv + rnnz);
    delete[] rc; delete[] rr; delete[] rv;
    kind = -1; N = (int)rn;
  } else
  gen_laplacian(kind, N, Ap, Ai, Ax);
  size_t n = Ap.size() - 1, nnzA = Ai.size();
  dump(dir, "A_p.i32", Ap.data(), n + 1); dump(dir, "A_i.i32", Ai.data(), nnzA); dump(dir, "A_x.f64", Ax.data(), nnzA);
  if (triOnly) {
    // The input itself is taken as a lower-triangular matrix in CSC (diagonal first): level sets of its DAG by the
    // reference's buildLevelSet_CSC (triangularSolve/Inspection_Level.h:12) and the column solves of
    // Triangular_CSC.h:14,50 on b_i = 1 + i/n, as examples/triangularTest_DAG.cpp:171-175 does.
    int *lp = NULL, *ls = NULL;
    int levels = buildLevelSet_CSC(n, nnzA, Ap.data(), Ai.data(), lp, ls);
    dump(dir, "tri_levelPtr.i32", lp, (size_t)levels + 1); dump(dir, "tri_levelSet.i32", ls, n);
    std::vector<double> x(n);
    for (size_t i2 = 0; i2 < n; ++i2) x[i2] = 1.0 + (double)i2 / (double)n;
    lsolve((int)n, Ap.data(), Ai.data(), Ax.data(), x.data());
    dump(dir, "tri_x.f64", x.data(), n);
    for (size_t i2 = 0; i2 < n; ++i2) x[i2] = 1.0 + (double)i2 / (double)n;
    lsolvePar((int)n, Ap.data(), Ai.data(), Ax.data(), x.data(), levels, lp, ls, chunk);
    dump(dir, "tri_x_par.f64", x.data(), n);
    // LBC on the DAG of a general lower-triangular matrix (cholesky/InspectionDAG_03.h:14, call site
    // examples/triangularTest_DAG_nonChordal.cpp:343-360: unit node costs) and lsolveParH2 on its schedule (:405)
    int hLevels = 0, hParts = 0, *hLevelPtr = NULL, *hLevelSet = NULL, *hParPtr = NULL, *hPartition = NULL;
    std::vector<double> nodeCost(n, 1.0);
    int avgcc = getCoarseLevelSet_DAG_CSC03(n, Ap.data(), Ai.data(), hLevels, hLevelPtr, hLevelSet, hParts, hParPtr,
                                            hPartition, costParam, levelParam, divRate, nodeCost.data());
    int nparts = hLevelPtr[hLevels];
    dump(dir, "dag_levelPtr.i32", hLevelPtr, (size_t)hLevels + 1); dump(dir, "dag_parPtr.i32", hParPtr, (size_t)nparts + 1);
    dump(dir, "dag_partition.i32", hPartition, n);
    for (size_t i2 = 0; i2 < n; ++i2) x[i2] = 1.0 + (double)i2 / (double)n;
    lsolveParH2((int)n, Ap.data(), Ai.data(), Ax.data(), x.data(), hLevels, hLevelPtr, hLevelSet, hParts, hParPtr,
                hPartition, chunk);
    dump(dir, "tri_x_h2.f64", x.data(), n);
    printf("{\"n\": %zu, \"nnz\": %zu, \"levels\": %d, \"dag_levels\": %d, \"dag_parts\": %d, \"dag_avg_cc\": %d}\n", n,
           nnzA, levels, hLevels, nparts, avgcc);
    return 0;
  }

  int *prunePtr = NULL, *pruneSet = NULL, *levelPtr = NULL, *levelSet = NULL, *parPtr = NULL, *partition = NULL;
  int nLevels = 0, nPar = 0, status = 0, maxSupWid = 0, maxCol = 0;
  double orderingTime = 0;
  int nrelax[3] = {4, 16, 48};
  double zrelax[3] = {0.8, 0.1, 0.05};
  CSC* Amat = new CSC;
  Amat->nzmax = nnzA; Amat->ncol = Amat->nrow = n;
  Amat->stype = -1; Amat->xtype = CHOLMOD_REAL; Amat->packed = TRUE;
  Amat->p = Ap.data(); Amat->i = Ai.data(); Amat->x = Ax.data(); Amat->nz = NULL; Amat->sorted = TRUE;
  double t0 = now();
  BCSC* L = analyze_p2(1, Amat, NULL, NULL, nrelax, zrelax, n, prunePtr, pruneSet, nLevels, levelPtr, levelSet,
                       nPar, parPtr, partition, costParam, levelParam, divRate, status, maxSupWid, maxCol,
                       orderingTime);
  double tSym = now() - t0;
  size_t nsuper = L->nsuper;
  double flops = 0;
  for (size_t j = 0; j < n; ++j) flops += (double)L->ColCount[j] * (double)L->ColCount[j];

  int nParts = levelPtr[nLevels];
  dump(dir, "Perm.i32", L->Perm, n); dump(dir, "ColCount.i32", L->ColCount, n);
  dump(dir, "super.i32", L->super, nsuper + 1); dump(dir, "sParent.i32", L->sParent, nsuper);
  dump(dir, "col2Sup.i32", L->col2Sup, n); dump(dir, "pi.u64", L->pi, nsuper + 1);
  dump(dir, "s.i32", L->s, L->ssize); dump(dir, "p.u64", L->p, n + 1); dump(dir, "i_ptr.u64", L->i_ptr, n + 1);
  dump(dir, "levelPtr.i32", levelPtr, nLevels + 1); dump(dir, "parPtr.i32", parPtr, nParts + 1);
  dump(dir, "partition.i32", partition, nsuper);

  {
    // DAG-based LBC over the factor's blocks (cholesky/Inspection_DAG_02.h:15; as analyze_DAG calls it,
    // cholesky/LSparsity.h:1412, with computeCostperBlock = width x rows as node cost)
    int dLevels = 0, dParts = 0, *dLevelPtr = NULL, *dLevelSet = NULL, *dParPtr = NULL, *dPartition = NULL;
    std::vector<double> blockCost(nsuper);
    for (size_t s2 = 0; s2 < nsuper; ++s2)
      blockCost[s2] = (double)(L->super[s2 + 1] - L->super[s2]) * (double)(L->i_ptr[L->super[s2 + 1]] - L->i_ptr[L->super[s2]]);
    getCoarseLevelSet_DAG_BCSC02(nsuper, L->p, L->i_ptr, L->s, L->super, L->col2Sup, dLevels, dLevelPtr, dLevelSet, dParts,
                                 dParPtr, dPartition, costParam, levelParam, divRate, blockCost.data());
    dump(dir, "dagb_levelPtr.i32", dLevelPtr, (size_t)dLevels + 1);
    dump(dir, "dagb_parPtr.i32", dParPtr, (size_t)dLevelPtr[dLevels] + 1);
    dump(dir, "dagb_partition.i32", dPartition, nsuper);
  }
  CSC* A1 = ptranspose(Amat, 2, L->Perm, NULL, 0, status);   // triu(PAP') (choleskyTest01.cpp:190)
  CSC* A2 = ptranspose(A1, 2, NULL, NULL, 0, status);        // tril(PAP') (choleskyTest01.cpp:191)
  dump(dir, "A1_p.i32", A1->p, n + 1); dump(dir, "A1_i.i32", A1->i, nnzA); dump(dir, "A1_x.f64", A1->x, nnzA);
  dump(dir, "A2_p.i32", A2->p, n + 1); dump(dir, "A2_i.i32", A2->i, nnzA); dump(dir, "A2_x.f64", A2->x, nnzA);

  printf("{\"kind\": %d, \"N\": %d, \"n\": %zu, \"nnzA\": %zu, \"nsuper\": %zu, \"xsize\": %zu, \"ssize\": %zu, "
         "\"nLevels\": %d, \"nParts\": %d, \"maxSupWid\": %d, \"maxCol\": %d, \"flops\": %.17g, \"threads\": %d, "
         "\"cost\": %d, \"level\": %d, \"div\": %d, \"t_inspector\": %.6f, \"t_ordering\": %.6f",
         kind, N, n, nnzA, nsuper, (size_t)L->xsize, (size_t)L->ssize, nLevels, nParts, maxSupWid, maxCol, flops,
         threads, costParam, levelParam, divRate, tSym, orderingTime);

  double* valL = NULL;
  if (doFactor) {
    valL = new double[L->xsize]();
    double* timing = new double[4 + threads + 64]();
    std::vector<double> tf, tl0, tl1;
    bool ok = true;
    for (int k = 0; k < iters; ++k) {
      memset(valL, 0, sizeof(double) * L->xsize);
      for (int i = 0; i < 4 + threads; ++i) timing[i] = 0;
      openblas_set_num_threads(1);   // the reference never lowers the BLAS thread count again (App. C.5)
      double t = now();
      // the reference hands the LAST-level BLAS thread count via `threads` (parallel_PB_Cholesky_05.h:271)
      ok = cholesky_left_par_05((int)n, A2->p, A2->i, A2->x, L->p, L->s, L->i_ptr, valL, L->super, (int)nsuper,
                                timing, L->sParent, A1->p, A1->i, L->col2Sup, nLevels, levelPtr, levelSet, nPar,
                                parPtr, partition, chunk, blasThreads, maxSupWid + 1, maxCol + 1);
      tf.push_back(now() - t); tl0.push_back(timing[0]); tl1.push_back(timing[1]);
      if (!ok) break;
    }
    openblas_set_num_threads(1);
    std::vector<size_t> idx(tf.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return tf[a] < tf[b]; });
    size_t m = idx[idx.size() / 2];
    double fro = 0; for (size_t i = 0; i < L->xsize; ++i) fro += valL[i] * valL[i];
    double tr = 0; for (size_t j = 0; j < n; ++j) tr += Ax[Ap[j]];
    printf(", \"factor_ok\": %d, \"t_factor\": %.6f, \"t_levels\": %.6f, \"t_last\": %.6f, \"fro2\": %.17g, \"trace\": %.17g",
           (int)ok, tf[m], tl0[m], tl1[m], fro, tr);
    printf(", \"t_factor_all\": [");
    for (size_t i = 0; i < tf.size(); ++i) printf("%s%.6f", i ? ", " : "", tf[i]);
    printf("]");
    if (dumpL) dump(dir, "valL.f64", valL, L->xsize);
  }

  if (doFactor && doSolve) {
    double* x = new double[n];
    int nnz = (int)L->xsize;
    int* sup2col = L->super;
    auto med = [&](std::vector<double>& v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
    std::vector<double> t1, t2, t3, t4;
    int okall = 1;
    // b = L*1 (common/Util.h:277); expected x == 1 (common/Util.h:294)
    rhsInitBlocked(n, nsuper, L->p, L->s, L->i_ptr, valL, x);
    dump(dir, "b_L1.f64", x, n);
    for (int k = 0; k < iters; ++k) {
      rhsInitBlocked(n, nsuper, L->p, L->s, L->i_ptr, valL, x);
      double t = now();
      blockedLsolve((int)n, L->p, L->s, valL, nnz, L->i_ptr, L->col2Sup, sup2col, (int)nsuper, x);
      t1.push_back(now() - t); okall &= testTriangular(n, x);
    }
    dump(dir, "x_blocked.f64", x, n);
    int* levelbPtr = new int[nsuper + 1]();
    int* levelbSet = new int[nsuper]();
    int blevels = getLevelSet(nsuper, L->sParent, levelbPtr, levelbSet);   // triangularTest02.cpp:218
    dump(dir, "etree_levelPtr.i32", levelbPtr, blevels + 1); dump(dir, "etree_levelSet.i32", levelbSet, nsuper);
    for (int k = 0; k < iters; ++k) {
      rhsInitBlocked(n, nsuper, L->p, L->s, L->i_ptr, valL, x);
      double t = now();
      leveledBlockedLsolve((int)n, L->p, L->s, valL, nnz, L->i_ptr, L->col2Sup, sup2col, (int)nsuper, x, blevels,
                           levelbPtr, levelbSet, chunk);
      t2.push_back(now() - t); okall &= testTriangular(n, x);
    }
    for (int k = 0; k < iters; ++k) {
      rhsInitBlocked(n, nsuper, L->p, L->s, L->i_ptr, valL, x);
      double t = now();
      H2LeveledBlockedLsolve((int)n, L->p, L->s, valL, nnz, L->i_ptr, L->col2Sup, sup2col, (int)nsuper, x, nLevels,
                             levelPtr, levelSet, nPar, parPtr, partition, chunk);
      t3.push_back(now() - t); okall &= testTriangular(n, x);
    }
    dump(dir, "x_h2.f64", x, n);
    for (int k = 0; k < iters; ++k) {
      rhsInitBlocked(n, nsuper, L->p, L->s, L->i_ptr, valL, x);
      openblas_set_num_threads(1);
      double t = now();
      H2LeveledBlockedLsolve_Peeled((int)n, L->p, L->s, valL, nnz, L->i_ptr, L->col2Sup, sup2col, (int)nsuper, x,
                                    nLevels, levelPtr, levelSet, nPar, parPtr, partition, chunk, blasThreads);
      t4.push_back(now() - t); okall &= testTriangular(n, x);
      openblas_set_num_threads(1);
    }
    // general-RHS forward solve (b_i = 1 + i/n in the permuted ordering) for elementwise parity
    for (size_t i2 = 0; i2 < n; ++i2) x[i2] = 1.0 + (double)i2 / (double)n;
    blockedLsolve((int)n, L->p, L->s, valL, nnz, L->i_ptr, L->col2Sup, sup2col, (int)nsuper, x);
    dump(dir, "y_ramp.f64", x, n);
    // CSC variants (Triangular_CSC.h) on bcsc2csc(L) (Util.h:311)
    size_t nnzC = 0;
    for (size_t s = 0; s < nsuper; ++s) {
      size_t w = L->super[s + 1] - L->super[s], r = L->i_ptr[L->super[s + 1]] - L->i_ptr[L->super[s]];
      nnzC += w * r - w * (w - 1) / 2;
    }
    double tcsc = -1;
    if (doCsc && nnzC < (size_t)INT_MAX) {
      int* Cp = new int[n + 1]; int* Ci = new int[nnzC]; double* Cx = new double[nnzC];
      bcsc2csc(n, nsuper, L->p, L->s, L->i_ptr, sup2col, valL, Cp, Ci, Cx);
      dump(dir, "Lcsc_p.i32", Cp, n + 1);
      if (dumpL) { dump(dir, "Lcsc_i.i32", Ci, nnzC); dump(dir, "Lcsc_x.f64", Cx, nnzC); }
      for (size_t i2 = 0; i2 < n; ++i2) x[i2] = 1.0 + (double)i2 / (double)n;
      double t = now();
      lsolve((int)n, Cp, Ci, Cx, x);
      tcsc = now() - t;
      dump(dir, "y_ramp_csc.f64", x, n);
      delete[] Cp; delete[] Ci; delete[] Cx;
    }
    printf(", \"t_h2_all\": [");
    for (size_t i = 0; i < t3.size(); ++i) printf("%s%.6f", i ? ", " : "", t3[i]);
    printf("]");
    printf(", \"solve_ok\": %d, \"t_blockedLsolve\": %.6f, \"t_leveled\": %.6f, \"t_h2\": %.6f, \"t_h2_peeled\": %.6f, "
           "\"t_lsolve_csc\": %.6f, \"etree_levels\": %d, \"nnzLcsc\": %zu",
           okall, med(t1), med(t2), med(t3), med(t4), tcsc, blevels, nnzC);
  }
  printf("}\n");
  return 0;
}
