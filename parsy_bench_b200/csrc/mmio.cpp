// Matrix-Market input for the host inspector (part of libparsy_inspector.so).
//
//   parsy_read_matrix      — what common/Util.h:77-179 `readMatrix` delivers to the drivers
//                            (examples/choleskyTest01.cpp:118-127): the lower half of a symmetric matrix, stored
//                            column-ordered in a coordinate file, as 0-based CSC (int col pointers, int rows, doubles).
//   parsy_make_lower_half  — examples/MakingLowerHalf.cpp:10-100 `printLower`: full symmetric coordinate file ->
//                            lower-half file with the diagonal moved away from zero by `tol`.
//
// Same accept/reject decisions as the reference for well-formed input (header checks in the same order, the
// `y > n` test); where the reference would silently produce a wrong CSC (a skipped column, a short file, a row
// index outside the matrix) this reader fails with a message instead.
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/parsy_inspector.h"

void parsy_inspector_set_error(const std::string& msg);   // inspector.cpp

// C stdio on purpose: this library carries a static libstdc++ (toolchain default here) and is loaded into processes
// that already hold another copy; iostream/locale state must not be shared between the two.
namespace {

enum { MM_OK = 0, MM_HEADER = 1, MM_BANNER = 2, MM_NOT_MATRIX = 3, MM_NOT_COORD = 4, MM_ARITH = 5, MM_DIMS = 6,
       MM_EMPTY = 7, MM_RANGE = 8, MM_ORDER = 9, MM_SHORT = 10, MM_IO = 11 };

int fail(int code, const std::string& msg) {
  parsy_inspector_set_error(msg);
  return code;
}

struct File {
  FILE* f;
  explicit File(const char* path, const char* mode) : f(std::fopen(path, mode)) {}
  ~File() { if (f) std::fclose(f); }
};

// one line without its terminator; false at end of file (or when the file could not be opened)
bool read_line(FILE* f, std::string& line) {
  line.clear();
  if (!f) return false;
  int ch;
  bool any = false;
  while ((ch = std::fgetc(f)) != EOF) {
    any = true;
    if (ch == '\n') break;
    line.push_back((char)ch);
  }
  return any;
}

std::vector<std::string> tokens(const std::string& line) {
  std::vector<std::string> t;
  size_t i = 0;
  while (i < line.size()) {
    while (i < line.size() && std::isspace((unsigned char)line[i])) ++i;
    size_t j = i;
    while (j < line.size() && !std::isspace((unsigned char)line[j])) ++j;
    if (j > i) t.push_back(line.substr(i, j - i));
    i = j;
  }
  return t;
}

// Header + size line, shared by the two entry points (Util.h:88-143, MakingLowerHalf.cpp:21-77).
int read_header(FILE* in, size_t& n, size_t& nnz) {
  std::string line;
  read_line(in, line);   // a missing file gives an empty header line, as getline on a closed ifstream does
  for (char& ch : line) ch = (char)std::tolower((unsigned char)ch);
  const std::vector<std::string> h = tokens(line);
  if (h.size() < 5) return fail(MM_HEADER, "Invalid header (first line does not contain 5 tokens)");
  if (h[0] != "%%matrixmarket") return fail(MM_BANNER, "Invalid header (first token is not \"%%MatrixMarket\")");
  if (h[1] != "matrix") return fail(MM_NOT_MATRIX, "Not a matrix; this driver cannot handle that.");
  if (h[2] != "coordinate") return fail(MM_NOT_COORD, "Not in coordinate format; this driver cannot handle that.");
  if (h[3] != "real") {
    if (h[3] == "complex") return fail(MM_ARITH, "Complex matrix; use zreadMM instead!");
    if (h[3] == "pattern") return fail(MM_ARITH, "Pattern matrix; values are needed!");
    return fail(MM_ARITH, "Unknown arithmetic");
  }
  bool more = true;
  while (more && line.compare(0, 1, "%") == 0) more = read_line(in, line);
  long long a = 0, b = 0, c = 0;
  if (!more || std::sscanf(line.c_str(), "%lld %lld %lld", &a, &b, &c) != 3) return fail(MM_DIMS, "The matrix dimension is missing");
  if (b <= 0 || c <= 0) return fail(MM_EMPTY, "empty matrix");
  n = (size_t)b;   // the reference reads both dimensions into n and keeps the second (Util.h:139)
  nnz = (size_t)c;
  return MM_OK;
}

}  // namespace

extern "C" int parsy_read_matrix(const char* path, int* n_out, int64_t* nnz_out, int** col_out, int** row_out,
                                 double** val_out) {
  if (!path || !n_out || !nnz_out || !col_out || !row_out || !val_out) return fail(MM_IO, "NULL argument");
  File in(path, "r");
  size_t n = 0, nnz = 0;
  int rc = read_header(in.f, n, nnz);
  if (rc) return rc;
  if (n > (size_t)INT32_MAX - 1 || nnz > (size_t)INT32_MAX) return fail(MM_RANGE, "matrix too large for int column pointers");
  int* col = (int*)std::calloc(n + 1, sizeof(int));
  int* row = (int*)std::malloc(nnz * sizeof(int));
  double* val = (double*)std::malloc(nnz * sizeof(double));
  if (!col || !row || !val) { std::free(col); std::free(row); std::free(val); return fail(MM_IO, "out of memory"); }
  auto bail = [&](int code, const std::string& msg) { std::free(col); std::free(row); std::free(val); return fail(code, msg); };
  size_t cur = 0, cnt = 0;   // current column and entries seen in it (Util.h:154-176)
  for (size_t k = 0; k < nnz; ++k) {
    long long x, y;
    double v;
    if (std::fscanf(in.f, "%lld %lld %lf", &x, &y, &v) != 3) return bail(MM_SHORT, "file ends before the announced number of entries");
    --x; --y;
    if (y < 0 || (size_t)y >= n) return bail(MM_RANGE, "column index outside the matrix");
    if (x < 0 || (size_t)x >= n) return bail(MM_RANGE, "row index outside the matrix");
    if ((size_t)y != cur) {
      // the reference advances by exactly one column whenever the column index changes (Util.h:167-170)
      if ((size_t)y != cur + 1 || cnt == 0) return bail(MM_ORDER, "entries must be ordered by column and no column may be empty");
      col[cur + 1] = col[cur] + (int)cnt;
      ++cur;
      cnt = 0;
    }
    row[k] = (int)x;
    val[k] = v;
    ++cnt;
  }
  if (cur != n - 1) return bail(MM_ORDER, "entries must be ordered by column and no column may be empty");
  col[n] = col[n - 1] + (int)cnt;
  *n_out = (int)n; *nnz_out = (int64_t)nnz; *col_out = col; *row_out = row; *val_out = val;
  return MM_OK;
}

extern "C" void parsy_matrix_free(int* col, int* row, double* val) {
  std::free(col); std::free(row); std::free(val);
}

extern "C" int parsy_make_lower_half(const char* in_path, const char* out_path, double tol) {
  if (!in_path || !out_path) return fail(MM_IO, "NULL argument");
  File in(in_path, "r");
  File out(out_path, "w");
  if (!out.f) return fail(MM_IO, std::string("cannot write ") + out_path);
  // the reference prints the new banner before it validates the old one (MakingLowerHalf.cpp:33)
  std::fputs("%%MatrixMarket matrix coordinate real symmetric\n", out.f);
  size_t n = 0, nnz = 0;
  int rc = read_header(in.f, n, nnz);
  if (rc) return rc;
  std::fprintf(out.f, "%zu %zu %zu\n", n, n, (nnz - n) / 2 + n);
  for (size_t k = 0; k < nnz; ++k) {
    long long x, y;
    double v;
    if (std::fscanf(in.f, "%lld %lld %lf", &x, &y, &v) != 3) return fail(MM_SHORT, "file ends before the announced number of entries");
    if (y > (long long)n) return fail(MM_RANGE, "column index outside the matrix");
    if (x >= y) {
      // "%g" = the default ostream formatting of a double (6 significant digits), as std::cout in the reference
      if (x == y) v = v >= 0 ? v + tol : v - tol;
      std::fprintf(out.f, "%lld %lld %g\n", x, y, v);
    }
  }
  return std::ferror(out.f) ? fail(MM_IO, "write failed") : MM_OK;
}
