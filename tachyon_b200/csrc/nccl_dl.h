// NCCL, resolved at run time.  The only exchange of a point-range-sharded MSM is one
// all-gather of a 128 / 192 / 256 / 384-byte partial sum per rank (the T partial
// sums of pippenger_adapter.h:110-113 taken across GPUs).  It is issued from C++ on the
// engine's own stream, right behind the kernel that produced the partial, so no host code runs
// between the last kernel of an MSM and the collective.
//
// libnccl is dlopen()ed instead of linked: a single-GPU user of this library needs no NCCL,
// and inside a process that already carries a libnccl.so.2 (torch bundles one) the same copy
// is reused.  Only the handful of entry points below are used; the declarations follow the
// public nccl.h (ncclUniqueId is 128 opaque bytes, ncclUint8 == 1).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>
#include <cstdlib>
#include <string>

namespace tb200 {

struct NcclUniqueId {
  char internal[128];
};

class NcclApi {
 public:
  using Comm = void*;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  static constexpr int kUint8 = 1;  // ncclUint8

  // nullptr (and *why set) when no usable libnccl is found
  static const NcclApi* Get(std::string* why) {
    static NcclApi api;
    static std::string error;
    static bool tried = false;
    if (!tried) {
      tried = true;
      const char* names[] = {getenv("TACHYON_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
      for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle_ = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle_) break;
        error = dlerror();
      }
      if (api.handle_) {
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle_, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle_, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle_, "ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle_, "ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle_, "ncclGetErrorString"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather ||
            !api.GetErrorString) {
          error = "libnccl lacks an expected symbol";
          api.handle_ = nullptr;
        }
      }
    }
    if (!api.handle_) {
      if (why) *why = "NCCL unavailable: " + error;
      return nullptr;
    }
    return &api;
  }

 private:
  void* handle_ = nullptr;
};

}  // namespace tb200
