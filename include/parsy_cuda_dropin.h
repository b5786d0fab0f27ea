/* parsy_cuda_dropin.h — forwarding header for the reference's own drivers.
 *
 * Include AFTER the reference's executor headers (cholesky/parallel_PB_Cholesky_05.h,
 * triangularSolve/Triangular_BCSC.h, triangularSolve/Triangular_CSC.h) and link libparsy_cuda.so: every call site of the
 * reference's free functions (examples/choleskyTest01.cpp:213-222, examples/triangularTest02.cpp:106-115,195-266,
 * examples/triangularTest_DAG.cpp:171-175) then runs the CUDA executor with unchanged arguments, array ownership and
 * return conventions.  The reference's functions stay defined (and unused) — this header only renames the calls.
 *
 * The trailing NULL of cholesky_left_par_05 supplies the reference's defaulted `nodCost` argument
 * (parallel_PB_Cholesky_05.h:39). */
#ifndef PARSY_CUDA_DROPIN_H
#define PARSY_CUDA_DROPIN_H
#include "parsy_cuda.h"
#define cholesky_left_par_05(...)            (parsy_cuda_cholesky_left_par_05(__VA_ARGS__, NULL) != 0)
#define blockedLsolve(...)                    parsy_cuda_blockedLsolve(__VA_ARGS__)
#define leveledBlockedLsolve(...)             parsy_cuda_leveledBlockedLsolve(__VA_ARGS__)
#define H2LeveledBlockedLsolve(...)           parsy_cuda_H2LeveledBlockedLsolve(__VA_ARGS__)
#define H2LeveledBlockedLsolve_Peeled(...)    parsy_cuda_H2LeveledBlockedLsolve_Peeled(__VA_ARGS__)
#define lsolve(...)                           parsy_cuda_lsolve(__VA_ARGS__)
#define lsolvePar(...)                        parsy_cuda_lsolvePar(__VA_ARGS__)
#define lsolveParH2(...)                      parsy_cuda_lsolveParH2(__VA_ARGS__)
#endif
