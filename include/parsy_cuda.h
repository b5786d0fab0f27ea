/* libparsy_cuda — C ABI of the B200 (sm_100a) executor for ParSy's numeric hot path.
 *
 * Two layers:
 *   1. DROP-IN entry points: the reference's free-function executor signatures, argument for argument
 *      (SURVEY.md §8(b)).  Every pointer is a HOST pointer owned by the caller; the call uploads, runs
 *      the CUDA executor and downloads the result (lValues / x).  Names carry a parsy_cuda_ prefix so
 *      the library can be linked next to the reference headers; INTEGRATION.md shows the one-line
 *      forwarding stubs.
 *   2. RESIDENT handle API: structure and values stay in HBM between calls so that the numeric
 *      factorization and the triangular sweeps can be timed (and re-run) without PCIe traffic.
 *
 * Plain C types only.  No CPU fallback exists: every entry point returns an error if no CUDA device
 * is usable (PARSY_CUDA_ERR_NO_DEVICE).
 */
#ifndef PARSY_CUDA_H
#define PARSY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------ */
/* status codes of the handle API                                                                    */
/* ------------------------------------------------------------------------------------------------ */
#define PARSY_CUDA_OK 0
#define PARSY_CUDA_ERR_NOT_SPD 1        /* a diagonal block was not positive definite (dpotrf info != 0) */
#define PARSY_CUDA_ERR_BAD_ARG 2        /* NULL / inconsistent structure arrays                          */
#define PARSY_CUDA_ERR_BAD_SCHEDULE 3   /* LBC schedule is not a legal topological order                 */
#define PARSY_CUDA_ERR_NO_DEVICE 4      /* no CUDA device / driver                                       */
#define PARSY_CUDA_ERR_CUDA 5           /* a CUDA runtime call failed (see parsy_cuda_last_error)        */
#define PARSY_CUDA_ERR_STATE 6          /* call order violated (e.g. solve before factor)                */
#define PARSY_CUDA_ERR_NO_MEMORY 7      /* host memory exhausted while building the plan                 */

const char* parsy_cuda_last_error(void);
int parsy_cuda_device_count(void);
/* library version: major*10000 + minor*100 + patch */
int parsy_cuda_version(void);

/* ------------------------------------------------------------------------------------------------ */
/* 1. drop-in entry points (host pointers, reference signatures)                                     */
/* ------------------------------------------------------------------------------------------------ */

/* Replaces  bool cholesky_left_par_05(...)   cholesky/parallel_PB_Cholesky_05.h:27-39
 * (call sites examples/choleskyTest01.cpp:213-222, examples/triangularTest02.cpp:106-115).
 *   c,r,values   tril(P A P') in CSC (int pointers);        lC,lR,Li_ptr  BCSC factor structure
 *   lValues      xsize doubles; the reference requires the caller to zero it (choleskyTest01.cpp:202);
 *                this implementation zeroes on the device, so the host content on entry is ignored
 *   blockSet     `super`, supNo+1 first-columns;            aTree  supernodal etree (sParent)
 *   cT,rT        triu(P A P') pattern — the reference feeds it to ereach_sn (common/Reach.h:112); here the
 *                descendant lists are derived from the factor structure itself (identical sets), so cT/rT
 *                may be NULL
 *   levelPtr/parPtr/partition   the LBC schedule (H-levels x w-partitions); levelSet/nPar unused (NULL/0)
 *   timing       >= 2 doubles: [0] = seconds in H-levels 0..nLevels-2, [1] = seconds in the last H-level
 *                (parallel_PB_Cholesky_05.h:266,418), measured with CUDA events on the device
 *   chunk, threads, super_max, col_max, nodCost  accepted for signature parity, not used
 * Returns 1 on success, 0 if a diagonal block is not positive definite or on any error
 * (the reference returns false on dpotrf info != 0, :206-207,252-253). */
int parsy_cuda_cholesky_left_par_05(int n, int* c, int* r, double* values, size_t* lC, int* lR, size_t* Li_ptr,
                                    double* lValues, int* blockSet, int supNo, double* timing, int* aTree, int* cT,
                                    int* rT, int* col2Sup, int nLevels, int* levelPtr, int* levelSet, int nPar,
                                    int* parPtr, int* partition, int chunk, int threads, int super_max, int col_max,
                                    double* nodCost);

/* The drop-in entry points keep the resident state of the last few structures they were called with (keyed by a hash
 * over the content of every structure array), because the reference's drivers call them repeatedly on one structure
 * (examples/choleskyTest01.cpp:199-222): a repeated call only moves values.  This releases that state (device memory);
 * PARSY_CUDA_DROPIN_CACHE=0 in the environment disables the cache altogether. */
void parsy_cuda_dropin_cache_clear(void);

/* Replaces  bool cholesky_left_sn_07(...)   cholesky/PB_Cholesky.h:16-19  (serial twin driven by a prune set).
 * prunePtr/pruneSet (descendant lists per supernode) are accepted and validated against the factor
 * structure; map/contribs scratch arguments are ignored (may be NULL). Runs the same device executor with
 * a single H-level whose single w-partition is 0..supNo-1. */
int parsy_cuda_cholesky_left_sn_07(int n, int* c, int* r, double* values, size_t* lC, int* lR, size_t* Li_ptr,
                                   double* lValues, int* blockSet, int supNo, double* timing, int* prunePtr,
                                   int* pruneSet, int* map, double* contribs);

/* Replace the supernodal forward solves of triangularSolve/Triangular_BCSC.h (:14, :115, :171, :238).
 * x is solved in place (host pointer, n doubles).  Return 1, or 0 when Lp/Li/x is NULL (:24,:185) or on error.
 * All four run the same device sweep; they differ in the schedule they are handed, exactly as the reference:
 * none (supernode order), etree level sets, LBC schedule, LBC schedule with a peeled last level. */
int parsy_cuda_blockedLsolve(int n, size_t* Lp, int* Li, double* Lx, int NNZ, size_t* Li_ptr, int* col2sup,
                             int* sup2col, int supNo, double* x);
int parsy_cuda_leveledBlockedLsolve(int n, size_t* Lp, int* Li, double* Lx, int NNZ, size_t* Li_ptr, int* col2sup,
                                    int* sup2col, int supNo, double* x, int levels, int* levelPtr, int* levelSet,
                                    int chunk);
int parsy_cuda_H2LeveledBlockedLsolve(int n, size_t* Lp, int* Li, double* Lx, int NNZ, size_t* Li_ptr,
                                      int* col2sup, int* sup2col, int supNo, double* x, int levels, int* levelPtr,
                                      int* levelSet, int parts, int* parPtr, int* partition, int chunk);
int parsy_cuda_H2LeveledBlockedLsolve_Peeled(int n, size_t* Lp, int* Li, double* Lx, int NNZ, size_t* Li_ptr,
                                             int* col2sup, int* sup2col, int supNo, double* x, int levels,
                                             int* levelPtr, int* levelSet, int parts, int* parPtr, int* partition,
                                             int chunk, int threads);
/* NEW (the reference has no backward sweep): solves L' x = b in place on the same arrays. */
int parsy_cuda_blockedLtsolve(int n, size_t* Lp, int* Li, double* Lx, int NNZ, size_t* Li_ptr, int* col2sup,
                              int* sup2col, int supNo, double* x);

/* Replace the column (CSC) forward solves of triangularSolve/Triangular_CSC.h (:14, :50, :76):
 * Lp/Li/Lx is a CSC lower-triangular matrix with the diagonal first in every column. */
int parsy_cuda_lsolve(int n, int* Lp, int* Li, double* Lx, double* x);
int parsy_cuda_lsolvePar(int n, int* Lp, int* Li, double* Lx, double* x, int levels, int* levelPtr, int* levelSet,
                         int chunk);
int parsy_cuda_lsolveParH2(int n, int* Lp, int* Li, double* Lx, double* x, int levels, int* levelPtr,
                           int* levelSet, int parts, int* parPtr, int* partition, int chunk);

/* Resident form of the column solves: structure (and the column order of the schedule) once, values and right-hand sides
 * per call.  `order` = the schedule flattened to a column order (levelSet of lsolvePar, partition of lsolveParH2, NULL for
 * 0..n-1); it must be a topological order of the column DAG (checked: PARSY_CUDA_ERR_BAD_SCHEDULE otherwise).  One kernel
 * launch per solve: warps take the columns in that order and wait on per-column dependency counters. */
typedef struct parsy_cuda_csc parsy_cuda_csc;
int parsy_cuda_csc_create(parsy_cuda_csc** out, int n, const int* Lp, const int* Li, const int* order, int device);
void parsy_cuda_csc_destroy(parsy_cuda_csc* h);
int parsy_cuda_csc_set_values(parsy_cuda_csc* h, const double* Lx);            /* nnz doubles, host */
/* x: host, n doubles, solved in place; device_ms (may be NULL): kernel time of the sweep from CUDA events */
int parsy_cuda_csc_solve(parsy_cuda_csc* h, double* x, double* device_ms);

/* ------------------------------------------------------------------------------------------------ */
/* 2. resident handle API                                                                            */
/* ------------------------------------------------------------------------------------------------ */
typedef struct parsy_cuda_solver parsy_cuda_solver;

/* Options: initialise with parsy_cuda_options_default() (a zero-filled struct would switch the CUDA graphs off), then
 * override what you need; passing NULL to parsy_cuda_create means the defaults. */
typedef struct parsy_cuda_options {
  int device;          /* CUDA device ordinal                                                     */
  int block_cols;      /* NB: block-column width wide supernodes are factored in (0 = default 128) */
  int use_graph;       /* 1 = capture the factorization / sweeps into CUDA graphs (default)        */
  int ignore_hlevels;  /* 1 = schedule by etree dependencies only, not by LBC H-level barriers     */
  int rank;            /* parsy_cuda_sharded_create: this process' rank ...                        */
  int world;           /* ... of `world` ranks (parsy_cuda_create: must be 0 or 1)                 */
  int reserved[10];    /* [0]=1 no look-ahead stream, [1]=1 per-step sweeps, [2] internal (phase), [3] top H-levels kept
                          shared, [4]=1 replicate the top instead of distributing it,
                          [5]=1 run the leaf region of the sweeps on the general dataflow kernel too,
                          [6]=1 keep the kernel classes of a step on one stream (no fan-out over auxiliary streams),
                          [7] sharded: consecutive top block columns per owner (0 = default 4),
                          [8] sharded: broadcast lanes (communicator + stream each; 0 = default 2, at most 4),
                          [9] sharded: 1 = broadcast whole block columns (no packing), 2 = also a dense factor buffer */
} parsy_cuda_options;
void parsy_cuda_options_default(parsy_cuda_options* opt);

/* Builds the device-resident symbolic state from the inspector's arrays (all HOST pointers, copied):
 *   n, c, r           tril(P A P') pattern (values come later)
 *   lC, lR, Li_ptr    BCSC structure (common/def.h:117-204)
 *   blockSet, supNo   supernode partition;  aTree = sParent;  col2Sup
 *   nLevels, levelPtr, parPtr, partition    LBC schedule (cholesky/InspectionLevel_06.h:18)
 * The derived task lists (descendant pairs with (lb, ndrow1, ndrow3), relative row indices, per-level
 * batches) are computed here once per structure, on up to PARSY_PLAN_THREADS host threads (environment; default
 * min(8, hardware threads); the lists do not depend on the count).  PARSY_PLAN_TIMING=1 prints the planner's laps. */
int parsy_cuda_create(parsy_cuda_solver** out, int n, const int* c, const int* r, const size_t* lC, const int* lR,
                      const size_t* Li_ptr, const int* blockSet, int supNo, const int* aTree, const int* col2Sup,
                      int nLevels, const int* levelPtr, const int* parPtr, const int* partition,
                      const parsy_cuda_options* opt);
void parsy_cuda_destroy(parsy_cuda_solver* s);

/* A values: nnz(A) doubles in the order of c/r. Host -> device, asynchronous on the handle's stream: the host buffer
 * must stay untouched until parsy_cuda_sync (or any download) returns.  The same holds for parsy_cuda_set_rhs and
 * parsy_cuda_set_factor. */
int parsy_cuda_set_values(parsy_cuda_solver* s, const double* values);
/* Numeric factorization on the device: zero L, scatter A, run every H-level. Asynchronous on the solver's
 * stream; parsy_cuda_sync() or any download waits for it. */
int parsy_cuda_factor(parsy_cuda_solver* s);
/* Waits for the stream; returns PARSY_CUDA_ERR_NOT_SPD if the last factorization hit a bad pivot. */
int parsy_cuda_sync(parsy_cuda_solver* s);
/* Device factor -> host lValues (xsize doubles), layout identical to the reference's valL. */
int parsy_cuda_get_factor(parsy_cuda_solver* s, double* lValues);
/* Host lValues -> device (lets the solves run on a factor produced elsewhere). */
int parsy_cuda_set_factor(parsy_cuda_solver* s, const double* lValues);

/* Triangular sweeps on the resident factor; rhs lives on the device. */
#define PARSY_CUDA_SOLVE_FWD 1  /* L y = b   */
#define PARSY_CUDA_SOLVE_BWD 2  /* L' x = y  */
int parsy_cuda_set_rhs(parsy_cuda_solver* s, const double* b);      /* host n doubles -> device */
int parsy_cuda_get_rhs(parsy_cuda_solver* s, double* x);            /* device -> host           */
int parsy_cuda_solve(parsy_cuda_solver* s, int which);              /* FWD, BWD or FWD|BWD      */

/* Full system A x = b on the factored handle.  The reference itself stops at the forward
 * sweep (fact 2 of SURVEY.md); its driver only sketches the CHOLMOD-style right-hand side b_i = 1 + i/n and solve
 * (examples/choleskyTest01.cpp:408-432), so parity is pinned by the oracle's restated sweeps and the residual.
 *
 * parsy_cuda_set_permutation: perm[k] = index, in the caller's ordering, of the unknown at position k of the factored
 *   matrix — L->Perm of the inspector (cholesky/LSparsity.h:613, choleskyTest01.cpp:190).  NULL = identity.  Returns
 *   PARSY_CUDA_ERR_BAD_ARG unless perm is a permutation of 0..n-1.
 * parsy_cuda_solve_system: for each of the nrhs columns of b (host, caller's ordering, column j at b + j*ld):
 *   y = P b;  L z = y;  L' w = z;  then refine_steps times { r = y - (P A P') w;  w += (L L')^{-1} r };  x = P' w.
 *   x (host) may alias b.  rel_residual, if not NULL, receives (refine_steps+1) values per column:
 *   ||y - (P A P') w||_2 / ||y||_2 before each refinement step and after the last one.  The residual uses the values
 *   last given to parsy_cuda_set_values; without them (handle built from set_factor only) refine_steps must be 0 and
 *   rel_residual NULL, else PARSY_CUDA_ERR_STATE. */
int parsy_cuda_set_permutation(parsy_cuda_solver* s, const int* perm);
int parsy_cuda_solve_system(parsy_cuda_solver* s, const double* b, double* x, int nrhs, int64_t ld, int refine_steps,
                            double* rel_residual);

/* Device-side timing of the last parsy_cuda_factor call (CUDA events):
 * out[0] = all H-levels but the last, out[1] = last H-level, out[2] = assembly (zero + scatter A), seconds. */
int parsy_cuda_factor_times(parsy_cuda_solver* s, double* out3);

/* One factorization without CUDA graphs, every kernel launch bracketed by CUDA events on the solver's stream;
 * returns, per kernel class, the summed device time (ms), the number of launches and the algorithmic flops:
 * 0 factor_small, 1 potrf_block, 2 TRSM tiles (DMMA), 3 update tiles 128 (DMMA), 4 update tiles 64 (DMMA),
 * 5 update_small.  This is where bench.py's roofline numbers come from. */
int parsy_cuda_factor_profiled(parsy_cuda_solver* s, double* class_ms6, int64_t* class_launches6, double* class_flops6);
/* The same pass, one record per kernel launch: dependency step, kernel class (numbering above), device time in ms.
 * Returns the number of launches (records beyond max_records are dropped), -1 on error. */
int parsy_cuda_factor_trace(parsy_cuda_solver* s, int max_records, int* step, int* cls, float* ms);

/* Introspection used by bench.py / tests (counts are exact, computed by the planner). */
/* Diagnostics: timeline of the general sweep kernel (which = PARSY_CUDA_SOLVE_FWD or PARSY_CUDA_SOLVE_BWD).  Runs one
 * un-captured sweep on the current right-hand side and reports, for each of its CTAs in plan order: kind (0 = up to
 * eight narrow supernodes, 1 = one slice of a block column, 2 = one narrow supernode with a long panel), rows of the
 * slice (kind 0: supernodes in the CTA), and the microseconds — relative to the first CTA's start — at which it started,
 * saw its inputs complete (kinds 0 and 2: = start), and finished.  Returns the number of CTAs, or -1. */
int parsy_cuda_sweep_trace(parsy_cuda_solver* s, int which, int max_records, int* kind, int* nrows, double* t_start_us,
                           double* t_ready_us, double* t_end_us);

typedef struct parsy_cuda_stats {
  int64_t n, nsuper, xsize, ssize, nnzA;
  int64_t n_pairs;            /* (supernode, descendant) update pairs, = sum of ereach_sn sizes   */
  int64_t n_pairs_small;      /* pairs run by the warp-cooperative FMA kernel                      */
  int64_t n_pairs_tiled;      /* pairs run by the DMMA tile kernel                                 */
  int64_t n_steps;            /* dependency steps (kernel-batch boundaries)                        */
  int64_t n_block_cols;       /* block columns of wide supernodes                                  */
  int64_t rel_entries;        /* relative-index entries resident on the device                     */
  int64_t launches_factor;    /* kernel launches per factorization                                 */
  int64_t launches_fwd, launches_bwd;
  double flops_potrf, flops_trsm, flops_update;   /* executed supernodal flops (F_sn)             */
  double bytes_solve;         /* algorithmic bytes of one sweep                  */
  int64_t device_bytes;       /* HBM held by the handle                                            */
  int64_t reserved[8];
} parsy_cuda_stats;
int parsy_cuda_get_stats(parsy_cuda_solver* s, parsy_cuda_stats* out);

/* HOST-ONLY: runs the planner (schedule validation, descendant pairs, step assignment) without touching a device
 * and fills the structural fields of `out`; used by the CPU test-suite.  Same return codes as parsy_cuda_create. */
int parsy_cuda_plan_check(int n, const size_t* lC, const int* lR, const size_t* Li_ptr, const int* blockSet, int supNo,
                          const int* col2Sup, int nLevels, const int* levelPtr, const int* parPtr,
                          const int* partition, const parsy_cuda_options* opt, parsy_cuda_stats* out);

/* HOST-ONLY: 64-bit digest of everything the planner hands to the executor for these arrays and options (task lists,
 * step table, sweep plan, ownership, broadcast lists): equal digests = the same launches on the same operands.  The CPU
 * test-suite pins it for a set of matrices, so that planner changes that alter the device work are seen without a GPU. */
int parsy_cuda_plan_digest(int n, const size_t* lC, const int* lR, const size_t* Li_ptr, const int* blockSet, int supNo,
                           const int* col2Sup, int nLevels, const int* levelPtr, const int* parPtr,
                           const int* partition, const parsy_cuda_options* opt, uint64_t* digest);

/* Raw device pointers for callers that keep data on the GPU (e.g. bench.py with torch tensors). */
double* parsy_cuda_device_factor(parsy_cuda_solver* s);   /* xsize doubles */
double* parsy_cuda_device_rhs(parsy_cuda_solver* s);      /* n doubles     */
double* parsy_cuda_device_values(parsy_cuda_solver* s);   /* nnzA doubles  */
void* parsy_cuda_stream(parsy_cuda_solver* s);            /* cudaStream_t  */

/* ------------------------------------------------------------------------------------------------ */
/* 3. sharded factorization + solve over the GPUs of one node (one process per GPU)     */
/* ------------------------------------------------------------------------------------------------ */
/* The reference has no multi-device path; this layer follows SURVEY.md §8(e).  The bottom of the tree —
 * every LBC H-level except the last options.reserved[3] (default 1) — is a forest of subtrees hanging off the top
 * separators (cholesky/InspectionLevel_06.h:208-216); they are dealt to the ranks in contiguous column order, balanced
 * by the flops their owner executes.  Per factorization:
 *   phase 1  no communication: each rank zeroes and assembles what it owns, factors its subtrees and applies every update
 *            they generate — also those into the top separators, accumulated in the rank's own copy of the top panels
 *            (fan-in);
 *   sum      one ncclAllReduce per contiguous run of top panels (traffic ~ size of the separators);
 *   top      block columns of the top separators are owned round-robin: the owner applies the updates into its block
 *            columns, factors them (POTRF + TRSM) and ncclBroadcasts the finished panel, step by step with a two-stream
 *            look-ahead.
 * All kernels and collectives of a factorization are captured into CUDA graphs at creation; NCCL (libnccl.so.2) is
 * loaded with dlopen on first use.  Factor memory: a rank touches only its own subtrees and the top.
 *
 * parsy_cuda_nccl_unique_id: call on rank 0, hand the 128 bytes to every rank (MPI / torch.distributed / a file).
 * parsy_cuda_sharded_create: arguments as parsy_cuda_create; options.rank / options.world / options.device select the
 *   rank; options.reserved[3] = top H-levels kept shared, reserved[4] = 1 replicates the top instead of distributing it.
 *   nccl_unique_id == NULL emulates all `world` ranks inside this process on options.device (device copies instead of
 *   NCCL, one stream, no overlap) — used by the single-GPU parity tests.
 * The solve keeps the same ownership: subtree sweeps on the owner, one all-reduce of the top part of the right-hand side,
 * the top separators on every rank, one all-reduce that leaves the full solution on every rank. */
typedef struct parsy_cuda_sharded parsy_cuda_sharded;
int parsy_cuda_nccl_unique_id(void* out128);
int parsy_cuda_sharded_create(parsy_cuda_sharded** out, int n, const int* c, const int* r, const size_t* lC, const int* lR,
                              const size_t* Li_ptr, const int* blockSet, int supNo, const int* aTree, const int* col2Sup,
                              int nLevels, const int* levelPtr, const int* parPtr, const int* partition,
                              const parsy_cuda_options* opt, const void* nccl_unique_id);
void parsy_cuda_sharded_destroy(parsy_cuda_sharded* s);
int parsy_cuda_sharded_set_values(parsy_cuda_sharded* s, const double* values);   /* nnz(A) doubles, host, every rank */
int parsy_cuda_sharded_factor(parsy_cuda_sharded* s);                             /* asynchronous                    */
int parsy_cuda_sharded_sync(parsy_cuda_sharded* s);                               /* PARSY_CUDA_ERR_NOT_SPD on a bad pivot seen by this rank */
int parsy_cuda_sharded_set_rhs(parsy_cuda_sharded* s, const double* b);           /* n doubles, host, every rank     */
int parsy_cuda_sharded_solve(parsy_cuda_sharded* s, int which);                   /* FWD|BWD: x on every rank        */
int parsy_cuda_sharded_get_rhs(parsy_cuda_sharded* s, double* x);
/* writes the panels this process holds complete (its subtrees + the top separators) into the host array, reference layout */
int parsy_cuda_sharded_get_factor(parsy_cuda_sharded* s, double* lValues);
/* seconds of the last factorization on this rank: [0] phase 1, [1] sum of the top panels, [2] distributed top */
int parsy_cuda_sharded_phase_times(parsy_cuda_sharded* s, double* out3);
/* [0] kernel launches per factorization, [1]/[2] NCCL broadcasts / all-reduces per factorization, [3]/[4] their bytes,
 * [5] HBM held on this device, [6] dependency steps of the top chain, [7] NCCL version, [8]/[9] launches per forward /
 * backward sweep, [10] supernodes owned by this rank, [11] shared top supernodes */
int parsy_cuda_sharded_stats(parsy_cuda_sharded* s, int64_t* out12);
/* Diagnostics: one factorization without graphs, timing events around every part of every step of the distributed top.
 * out[7*k + j] = milliseconds since the top phase started on this rank, step k: j = 0 F begin, 1 F end, 2 A end (chain
 * stream), 3 / 4 broadcast begin / end (communication stream), 5 / 6 R begin / end (bulk stream); owner_of_step[k] = rank
 * that factors the step's first block column.  Returns the number of steps, -1 on error. */
int parsy_cuda_sharded_trace_top(parsy_cuda_sharded* s, int max_steps, float* out, int* owner_of_step);
double* parsy_cuda_sharded_device_factor(parsy_cuda_sharded* s);
double* parsy_cuda_sharded_device_rhs(parsy_cuda_sharded* s);
void* parsy_cuda_sharded_stream(parsy_cuda_sharded* s);
/* the phase-1 / phase-2 plan of this process' rank (emulation: of `emulated_rank`), for parsy_cuda_get_stats and
 * parsy_cuda_owned_ranges; owned by the sharded handle */
parsy_cuda_solver* parsy_cuda_sharded_plan(parsy_cuda_sharded* s, int emulated_rank, int phase);
/* contiguous runs of lValues owned by `rank` (begin, end pairs in doubles); returns their number */
int parsy_cuda_owned_ranges(parsy_cuda_solver* s, int rank, int64_t* begin_end_pairs, int max_pairs);
/* HOST-ONLY twin (no device): plans for `world` ranks and returns the runs owned by `for_rank`; -1 on error. */
int parsy_cuda_plan_owned_ranges(int n, const size_t* lC, const int* lR, const size_t* Li_ptr, const int* blockSet,
                                 int supNo, const int* col2Sup, int nLevels, const int* levelPtr, const int* parPtr,
                                 const int* partition, int world, int top_levels, int for_rank,
                                 int64_t* begin_end_pairs, int max_pairs);

#ifdef __cplusplus
}
#endif
#endif /* PARSY_CUDA_H */
