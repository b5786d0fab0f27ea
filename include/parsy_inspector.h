/* libparsy_inspector — host-side symbolic inspector (C ABI).
 *
 * Re-statement of the reference's analyze_p2 (cholesky/LSparsity.h:256-842) and of the two ptranspose calls the
 * drivers make (examples/choleskyTest01.cpp:190-191): METIS nested dissection, elimination tree, postorders,
 * column counts, relaxed supernodes, supernodal row patterns and the LBC schedule
 * (cholesky/InspectionLevel_06.h:18).  Every array it returns must equal the reference's bit for bit
 * (tests/test_inspector.py compares them against the compiled reference).  Plain host arrays, owned by the
 * returned object; the executor (include/parsy_cuda.h) consumes them unchanged.
 */
#ifndef PARSY_INSPECTOR_H
#define PARSY_INSPECTOR_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors the fields of BCSC (common/def.h:117-204) the executor reads, plus the schedule and P A P'. */
typedef struct parsy_symbolic {
  int n;              /* order of A                                                          */
  int nsuper;         /* number of supernodes                                                */
  int64_t nnzA;       /* entries of tril(A)                                                  */
  int64_t xsize;      /* doubles in the supernodal factor (sum width*rows)                   */
  int64_t ssize;      /* entries of the concatenated row lists                               */
  int maxSupWid;      /* widest supernode                                                    */
  int maxCol;         /* longest supernode (rows)                                            */
  double flops;       /* sum_j ColCount[j]^2  (cholesky/ColumnCount.h:486-498)               */
  double t_ordering;  /* seconds in METIS_NodeND                                             */
  double t_total;     /* seconds in the whole inspector                                      */
  int* Perm;          /* n: Perm[k] = original index placed at position k                    */
  int* ColCount;      /* n: column counts of the simplicial factor, post-ordered             */
  int* Parent;        /* n: column elimination tree, post-ordered                            */
  int* super;         /* nsuper+1: first column of each supernode            (blockSet)      */
  int* sParent;       /* nsuper: supernodal etree                            (aTree)         */
  int* col2Sup;       /* n                                                                  */
  size_t* pi;         /* nsuper+1: start of each supernode's row list in s                   */
  int* s;             /* ssize: row lists                                    (lR)            */
  size_t* p;          /* n+1: offset of every column in the factor           (lC)            */
  size_t* i_ptr;      /* n+1: start of the row list of the column's supernode (Li_ptr)       */
  int nLevels;        /* LBC H-levels                                                        */
  int nParts;         /* total number of w-partitions                                        */
  int* levelPtr;      /* nLevels+1                                                           */
  int* parPtr;        /* nParts+1                                                            */
  int* partition;     /* nsuper                                                              */
  int* A1_p; int* A1_i;                /* triu(P A P') pattern by columns (cT, rT)           */
  int* A2_p; int* A2_i; double* A2_x;  /* tril(P A P') by columns, rows ascending (c, r, values) */
  int64_t* A2_src;    /* nnzA: position in the input arrays of every A2 entry (to re-permute new values) */
} parsy_symbolic;

/* A = lower-half CSC (rows ascending, diagonal first), as common/Util.h:77 `readMatrix` delivers it.
 * costParam / levelParam / divRate = innerParts / minLevelDist / divRate of getCoarseLevelSet_6.
 * userPerm: NULL -> METIS_NodeND with default options on the diagonal-free full graph (LSparsity.h:534-613);
 *           otherwise a permutation of 0..n-1 used in place of the METIS result.
 * Returns 0 on success. */
int parsy_inspect(int n, const int* Ap, const int* Ai, const double* Ax, int costParam, int levelParam, int divRate,
                  const int* userPerm, parsy_symbolic** out);
void parsy_symbolic_free(parsy_symbolic* s);
const char* parsy_inspector_last_error(void);

/* Descendant supernodes of supernode s in topological order, exactly what ereach_sn (common/Reach.h:112-143)
 * leaves in xi[top..supNo): fills out[0..count) and returns count (or -1 on bad input). Work arrays are internal. */
int parsy_ereach_sn(const parsy_symbolic* sym, int s, int* out);

/* Supernodal etree level sets as the reference's getLevelSet (common/TreeUtils.h:119) builds them for
 * leveledBlockedLsolve (examples/triangularTest02.cpp:218): returns the number of levels. */
int parsy_etree_level_set(int nsuper, const int* sParent, int* levelPtr /*nsuper+1*/, int* levelSet /*nsuper*/);

/* Level sets of the dependence DAG of a lower-triangular CSC matrix (diagonal first in each column) — replaces
 * buildLevelSet_CSC (triangularSolve/Inspection_Level.h:12-59; call site examples/triangularTest_DAG.cpp:171-175), the
 * inspector of lsolvePar for inputs that are not Cholesky factors.  levelPtr has n+1 entries (levels+1 are used),
 * levelSet n; columns of a level are listed in increasing order, as the reference does.  Returns the number of levels,
 * or -1 (missing diagonal, entry above the diagonal). */
int parsy_build_level_set_csc(int n, const int* Lp, const int* Li, int* levelPtr, int* levelSet);

/* Load-balanced level coarsening on the DAG of a general lower-triangular CSC matrix (diagonal first in each column) —
 * replaces getCoarseLevelSet_DAG_CSC03 (cholesky/InspectionDAG_03.h:14; call site
 * examples/triangularTest_DAG_nonChordal.cpp:343-360), the inspector of lsolveParH2 (Triangular_CSC.h:76) for inputs that
 * are not Cholesky factors.  innerParts / minLevelDist / divRate as in the reference (its drivers pass costParam /
 * levelParam / divRate); nodeCost: n doubles or NULL for unit costs (what the reference driver uses).  Outputs, bit for bit
 * the reference's: *nLevels H-levels, levelPtr (caller: n+1 ints; nLevels+1 used) into parPtr (caller: n+1 ints) into
 * partition (n columns).  Returns 0, 2 for bad arguments, 3 where the reference itself would index out of range. */
int parsy_dag_lbc_csc(int n, const int* Lp, const int* Li, int innerParts, int minLevelDist, int divRate,
                      const double* nodeCost, int* nLevels, int* levelPtr, int* parPtr, int* partition);

/* The supernodal twin: LBC on the DAG of the BLOCKS of a BCSC factor — replaces getCoarseLevelSet_DAG_BCSC02
 * (cholesky/Inspection_DAG_02.h:15; call site cholesky/LSparsity.h:1412 in analyze_DAG, the inspector behind
 * examples/triangularTest_DAG.cpp).  Li_ptr / lR / blk2col (`super`) / col2blk (`col2Sup`) as in the BCSC structure;
 * nodeCost: nblocks doubles or NULL for unit costs (analyze_DAG passes computeCostperBlock = width x rows).  Outputs as
 * parsy_dag_lbc_csc over block ids: a schedule that cholesky_left_par_05 and H2LeveledBlockedLsolve accept. */
int parsy_dag_lbc_bcsc(int nblocks, const size_t* Li_ptr, const int* lR, const int* blk2col, const int* col2blk,
                       int innerParts, int minLevelDist, int divRate, const double* nodeCost, int* nLevels, int* levelPtr,
                       int* parPtr, int* partition);

/* BCSC -> CSC conversion of a supernodal factor (common/Util.h:311 bcsc2csc); Cp has n+1 entries; returns nnz.
 * Pass Ci = Cx = NULL to only count. */
int64_t parsy_bcsc2csc(const parsy_symbolic* sym, const double* Lx, int* Cp, int* Ci, double* Cx);

/* Matrix-Market input.
 *
 * parsy_read_matrix replaces `readMatrix` (common/Util.h:77-179; call site examples/choleskyTest01.cpp:118-127):
 * a coordinate file holding the LOWER half of a symmetric matrix, entries ordered by column, becomes 0-based CSC.
 * Row order inside a column is kept as written (the inspector wants the diagonal first, as the reference does).
 * Returns 0, or: 1 header has fewer than 5 tokens, 2 banner, 3 not "matrix", 4 not "coordinate", 5 arithmetic is not
 * "real", 6 size line missing, 7 n or nnz <= 0, 8 index outside the matrix, 9 entries not ordered by column / empty
 * column, 10 file shorter than announced, 11 I/O or NULL argument (text in parsy_inspector_last_error; the reference
 * prints the same messages for codes 1-6 and returns false). The arrays are malloc'ed; release with parsy_matrix_free. */
int parsy_read_matrix(const char* path, int* n, int64_t* nnz, int** col, int** row, double** val);
void parsy_matrix_free(int* col, int* row, double* val);

/* Replaces `printLower` of examples/MakingLowerHalf.cpp:10-100: reads a FULL symmetric coordinate file and writes the
 * entries with row >= col to out_path under a "symmetric" banner, size line "n n (nnz-n)/2+n", 1-based indices, the
 * diagonal shifted by +tol (or -tol when negative; the reference hard-codes tol = 0.1), values in the default ostream
 * format (6 significant digits) so that the output is byte-identical to the reference's stdout. Same return codes. */
int parsy_make_lower_half(const char* in_path, const char* out_path, double tol);

#ifdef __cplusplus
}
#endif
#endif
