// Integration example (INTEGRATION.md §1 made concrete): the flow of the reference's driver
// examples/choleskyTest01.cpp — read the matrix (:118-127), inspect (:150-165), five timed factorizations with the
// reference's argument list (:199-232), print the median run (:271-276), then the solve the reference only sketches
// (:408-432) — over the C ABI of this repository.  C++ host code only: everything numeric happens in libparsy_cuda.
//
//   choleskyTest_b200 <lower-half.mtx | 2d5:N | 3d7:N | 3d27:N> [costParam=592] [levelParam=1] [divRate=4] [refine=1]
//
// Output: the reference's CSV prefix  name,threads,chunk,costParam,levelParam,blasThreads,finalSeqNode,
//         t_factor,t_levels,t_last_level,t_symbolic,t_ordering,  followed by  residual=<||Ax-b||/||b||>.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/parsy_cuda.h"
#include "../include/parsy_inspector.h"

namespace {

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Synthetic inputs of BASELINE.json: lower half, diagonal first, rows ascending.
void laplacian(int kind, int N, std::vector<int>& p, std::vector<int>& i, std::vector<double>& x) {
  const int NZ = kind == 0 ? 1 : N;
  const long n = (long)N * N * NZ;
  p.assign(n + 1, 0);
  const double diag = kind == 0 ? 4.0 : (kind == 1 ? 6.0 : 26.0);
  for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < N; ++y)
      for (int xx = 0; xx < N; ++xx) {
        const long v = ((long)z * N + y) * N + xx;
        i.push_back((int)v); x.push_back(diag);
        for (int dz = 0; dz <= (kind == 0 ? 0 : 1); ++dz)
          for (int dy = (dz == 0 ? 0 : -1); dy <= 1; ++dy)
            for (int dx = ((dz == 0 && dy == 0) ? 1 : -1); dx <= 1; ++dx) {
              const int nz = z + dz, ny = y + dy, nx = xx + dx;
              if (nz < 0 || nz >= NZ || ny < 0 || ny >= N || nx < 0 || nx >= N) continue;
              if (kind != 2 && std::abs(dx) + std::abs(dy) + std::abs(dz) != 1) continue;
              i.push_back((int)(((long)nz * N + ny) * N + nx)); x.push_back(-1.0);
            }
        p[v + 1] = (int)i.size();
      }
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <lower-half.mtx | 2d5:N | 3d7:N | 3d27:N> [costParam] [levelParam] [divRate] [refine]\n", argv[0]);
    return 2;
  }
  const std::string f1 = argv[1];
  const int costParam = argc > 2 ? std::atoi(argv[2]) : 592, levelParam = argc > 3 ? std::atoi(argv[3]) : 1;
  const int divRate = argc > 4 ? std::atoi(argv[4]) : 4, refine = argc > 5 ? std::atoi(argv[5]) : 1;
  const int chunk = 1, numThread = 1, blasThreads = 1, finalSeqNode = 0;   // printed for CSV compatibility only

  // --- input (choleskyTest01.cpp:118-127) ---------------------------------------------------------------------
  std::vector<int> Ap, Ai;
  std::vector<double> Ax;
  int n = 0;
  const size_t colon = f1.find(':');
  if (colon != std::string::npos && (f1.compare(0, colon, "2d5") == 0 || f1.compare(0, colon, "3d7") == 0 || f1.compare(0, colon, "3d27") == 0)) {
    const int kind = f1[0] == '2' ? 0 : (f1.compare(0, colon, "3d7") == 0 ? 1 : 2);
    laplacian(kind, std::atoi(f1.c_str() + colon + 1), Ap, Ai, Ax);
    n = (int)Ap.size() - 1;
  } else {
    int64_t nnz = 0; int *c = nullptr, *r = nullptr; double* v = nullptr;
    if (parsy_read_matrix(f1.c_str(), &n, &nnz, &c, &r, &v) != 0) { std::fprintf(stderr, "%s\n", parsy_inspector_last_error()); return -1; }
    Ap.assign(c, c + n + 1); Ai.assign(r, r + nnz); Ax.assign(v, v + nnz);
    parsy_matrix_free(c, r, v);
  }

  // --- inspector (analyze_p2 + the two ptranspose calls, choleskyTest01.cpp:150-191) ----------------------------
  parsy_symbolic* L = nullptr;
  if (parsy_inspect(n, Ap.data(), Ai.data(), Ax.data(), costParam, levelParam, divRate, nullptr, &L) != 0) {
    std::fprintf(stderr, "inspector: %s\n", parsy_inspector_last_error());
    return -1;
  }

  // --- executor: the reference's call, five times, median reported (choleskyTest01.cpp:199-276) --------------------
  std::vector<double> valL((size_t)L->xsize);
  struct Run { double all, levels, last; };
  std::vector<Run> runs;
  const int iterNo = 5;
  for (int k = 0; k < iterNo; ++k) {
    std::fill(valL.begin(), valL.end(), 0.0);
    double timing[8] = {0};
    const double t0 = now();
    const int ok = parsy_cuda_cholesky_left_par_05(n, L->A2_p, L->A2_i, L->A2_x, L->p, L->s, L->i_ptr, valL.data(), L->super,
                                                   L->nsuper, timing, L->sParent, L->A1_p, L->A1_i, L->col2Sup, L->nLevels,
                                                   L->levelPtr, nullptr, 0, L->parPtr, L->partition, chunk, numThread,
                                                   L->maxSupWid + 1, L->maxCol + 1, nullptr);
    if (!ok) { std::fprintf(stderr, "cholesky_left_par_05: %s\n", parsy_cuda_last_error()); return -1; }
    runs.push_back({now() - t0, timing[0], timing[1]});
  }
  std::sort(runs.begin(), runs.end(), [](const Run& a, const Run& b) { return a.all < b.all; });
  const Run& mid = runs[iterNo / 2];
  std::printf("%s,%d,%d,%d,%d,%d,%d,%g,%g,%g,%g,%g,", f1.c_str(), numThread, chunk, costParam, levelParam, blasThreads,
              finalSeqNode, mid.all, mid.levels, mid.last, L->t_total, L->t_ordering);

  // --- A x = b with the driver's right-hand side b_i = 1 + i/n (choleskyTest01.cpp:428-432), resident handle -------------
  parsy_cuda_solver* h = nullptr;
  if (parsy_cuda_create(&h, n, L->A2_p, L->A2_i, L->p, L->s, L->i_ptr, L->super, L->nsuper, L->sParent, L->col2Sup, L->nLevels,
                        L->levelPtr, L->parPtr, L->partition, nullptr) != 0 ||
      parsy_cuda_set_values(h, L->A2_x) != 0 || parsy_cuda_factor(h) != 0 || parsy_cuda_sync(h) != 0 ||
      parsy_cuda_set_permutation(h, L->Perm) != 0) {
    std::fprintf(stderr, "handle: %s\n", parsy_cuda_last_error());
    return -1;
  }
  std::vector<double> b(n), x(n), rel(refine + 1);
  for (int i = 0; i < n; ++i) b[i] = 1.0 + (double)i / n;
  if (parsy_cuda_solve_system(h, b.data(), x.data(), 1, n, refine, rel.data()) != 0) { std::fprintf(stderr, "solve: %s\n", parsy_cuda_last_error()); return -1; }
  // residual in the caller's ordering, on the host, from the input arrays (independent of the device's own figure)
  std::vector<double> res(b);
  for (int j = 0; j < n; ++j)
    for (int q = Ap[j]; q < Ap[j + 1]; ++q) {
      res[Ai[q]] -= Ax[q] * x[j];
      if (Ai[q] != j) res[j] -= Ax[q] * x[Ai[q]];
    }
  double rr = 0, bb = 0;
  for (int i = 0; i < n; ++i) { rr += res[i] * res[i]; bb += b[i] * b[i]; }
  std::printf("residual=%.3e device_residual=%.3e\n", std::sqrt(rr / bb), rel[refine]);
  parsy_cuda_destroy(h);
  parsy_symbolic_free(L);
  return std::sqrt(rr / bb) < 1e-9 ? 0 : 1;
}
