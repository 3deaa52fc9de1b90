// nccl_dl.cuh -- the handful of NCCL entry points the engine group uses, resolved at run time with
// dlopen (the library stays loadable, and single-GPU use stays possible, on a machine without NCCL).
// Declarations follow nccl.h (NCCL 2.x ABI): ncclUniqueId is 128 bytes passed BY VALUE to
// ncclCommInitRank; ncclInt8 = 0; ncclSuccess = 0.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace gh {

typedef struct ncclComm *ncclComm_t;
struct ncclUniqueId { char internal[128]; };

struct NcclApi {
  int (*GetVersion)(int *);
  int (*GetUniqueId)(ncclUniqueId *);
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  int (*CommInitAll)(ncclComm_t *, int, const int *);
  int (*CommDestroy)(ncclComm_t);
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
};
// Loads libnccl (path: explicit, else $GH_NCCL_LIB, else "libnccl.so.2" -- which resolves to the
// copy torch already mapped when torch is in the process).  nullptr + gh_last_error on failure.
const NcclApi *nccl_api(const char *path);

}  // namespace gh
