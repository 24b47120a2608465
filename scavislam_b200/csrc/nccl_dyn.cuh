// nccl_dyn.cuh -- the handful of NCCL entry points the sharded window needs, resolved at
// run time with dlopen so that libsvsb200.so loads on hosts without NCCL and binds to whichever
// libnccl.so.2 the process already carries (e.g. the one torch.distributed loaded).
// Types restated from nccl.h (ABI-stable since NCCL 2.0): opaque communicator, 128-byte unique id.
#pragma once
#include <cuda_runtime.h>

namespace svs {

struct NcclUniqueId { char internal[128]; };
using NcclComm = void*;

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclFloat64 = 8;   // ncclDouble
constexpr int kNcclSum = 0;       // ncclSum

// nullptr when no NCCL library can be loaded
const NcclApi* nccl_api();

}  // namespace svs
