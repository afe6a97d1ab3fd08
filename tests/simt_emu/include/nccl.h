// TEST INFRASTRUCTURE ONLY: the NCCL types csrc/mif_api.cu names (it binds NCCL with dlopen; single-rank runs under
// the SIMT interpreter never call it).
#ifndef MIF_SIMT_EMU_NCCL_H
#define MIF_SIMT_EMU_NCCL_H
#include <cuda_runtime.h>
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef enum { ncclChar = 0, ncclInt = 2, ncclFloat = 7, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
#endif
