/* Shim for <metis.h> matching the CUDA toolkit's libmetis_static.a (64-bit idx_t). Oracle build only. */
#ifndef PARSY_ORACLE_METIS_SHIM_H
#define PARSY_ORACLE_METIS_SHIM_H
#include <stdint.h>
typedef int64_t idx_t;
typedef float real_t;
#define METIS_NOPTIONS 40
#define METIS_OK 1
#ifdef __cplusplus
extern "C" {
#endif
int METIS_NodeND(idx_t* nvtxs, idx_t* xadj, idx_t* adjncy, idx_t* vwgt, idx_t* options, idx_t* perm, idx_t* iperm);
int METIS_SetDefaultOptions(idx_t* options);
int METIS_Free(void* ptr);
#ifdef __cplusplus
}
#endif
#endif
