// dist.cuh — multi-GPU plumbing BELOW the C ABI (kc_comm_* / kc_*_dist / kc_gather_edges in kc_b200.h):
// one engine per GPU, one NCCL rank per engine (one process per GPU, or one thread per GPU in the CLI).
//
// The path shards by row blocks of the pair triangle with the owner computing: no collective
// sits between the kernels of a step.  What crosses NVLink:
//   staging   every rank uploads 1 / world of the residue stream over its own PCIe link and the slices are
//             all-gathered in place (ncclAllGather): the upload cost of a step no longer grows with the ranks
//   counters  one ncclAllReduce of the index counters, one of the pair counters (the numbers the reference
//             prints at src/graph/mod.rs:50-51, :695, :545 are whole-set numbers)
//   edges     the per-rank sorted runs, laid out in block order on rank 0 by ONE group of ncclSend / ncclRecv,
//             or copied by every rank straight into a host buffer all ranks map (kc_gather_edges_shared)
// NCCL is loaded at run time (dlopen "libnccl.so.2": in a Python process that is the copy torch already
// mapped), so the single-GPU library has no NCCL dependency.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace kc {

struct NcclApi {
  void* handle = nullptr;
  std::string err;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return handle != nullptr && err.empty(); }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.err = "libnccl.so.2 not found (dlopen)";
    return api;
  }
  auto sym = [&](const char* n) -> void* {
    void* p = dlsym(api.handle, n);
    if (!p) api.err = std::string("NCCL symbol missing: ") + n;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return api;
}

// first edge of the sorted list whose `a` is >= row: the split between a rank's early-block and late-block run
__global__ void edge_lower_bound_kernel(const uint4* __restrict__ edges, unsigned long long n, uint32_t row,
                                        unsigned long long* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long lo = 0, hi = n;
  while (lo < hi) {
    const unsigned long long mid = (lo + hi) >> 1;
    if (edges[mid].x < row) lo = mid + 1; else hi = mid;
  }
  *out = lo;
}

}  // namespace kc
