// C ABI of the engine, common part: library-wide state, the field-element parity hooks and
// the global helpers.  The per-curve, per-group entry points — the reference's MSM-GPU API
// (tachyon/c/math/elliptic_curves/generator/msm_gpu.cc.tpl:8-33,
//  tachyon/c/math/elliptic_curves/msm/msm_gpu.h:22-122) and this library's extensions — are
// instantiated in msm_api_<curve>_<group>.cu.  See include/tachyon_msm_b200.h.
#include "msm_api_common.cuh"

namespace tb200 {
std::atomic<uint64_t> g_kernel_launches{0};
thread_local std::string g_last_error;
}  // namespace tb200

extern "C" {

TB200_DEFINE_FIELD_API(bn254, Bn254Curve)
TB200_DEFINE_FIELD_API(bls12_381, Bls381Curve)

int tachyon_b200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return -(int)e;
  }
  return n;
}

uint32_t tachyon_b200_window_bits(size_t n, uint32_t scalar_bits) {
  return ChooseWindowBits(n, scalar_bits);
}
uint32_t tachyon_b200_window_count(uint32_t scalar_bits, uint32_t window_bits) {
  return WindowsFor(scalar_bits, window_bits);
}

const char* tachyon_b200_last_error(void) { return g_last_error.c_str(); }

int tachyon_b200_nccl_unique_id(void* out128) {
  if (!out128) return -1;
  std::string why;
  const NcclApi* nccl = NcclApi::Get(&why);
  if (!nccl) {
    g_last_error = why;
    return -1;
  }
  NcclUniqueId id;
  int rc = nccl->GetUniqueId(&id);
  if (rc != 0) {
    g_last_error = nccl->GetErrorString(rc);
    return -1;
  }
  memcpy(out128, &id, sizeof(id));
  return 0;
}

uint64_t tachyon_b200_kernel_launch_count(void) { return g_kernel_launches.load(); }

void* tachyon_b200_alloc_host(size_t bytes, int write_combined) {
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1,
                                write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_last_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e);
    return nullptr;
  }
  return p;
}

void tachyon_b200_free_host(void* p) {
  if (p) cudaFreeHost(p);
}

double tachyon_b200_imad_peak(int device, int variant, int repeats) {
  try {
    TB_CUDA(cudaSetDevice(device));
    int sms = 0;
    TB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    uint32_t* out;
    TB_CUDA(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1;
    TB_CUDA(cudaEventCreate(&e0));
    TB_CUDA(cudaEventCreate(&e1));
    const uint32_t iters = 4096, blocks = sms * 8, threads = 256;
    double best = 0;
    for (int r = 0; r < repeats + 1; ++r) {
      TB_CUDA(cudaEventRecord(e0));
      if (variant == 0)
        imad_peak_kernel<0><<<blocks, threads>>>(iters, 12345u + r, out);
      else if (variant == 1)
        imad_peak_kernel<1><<<blocks, threads>>>(iters, 12345u + r, out);
      else
        imad_peak_kernel<2><<<blocks, threads>>>(iters, 12345u + r, out);
      TB_CUDA(cudaEventRecord(e1));
      TB_CUDA(cudaEventSynchronize(e1));
      g_kernel_launches.fetch_add(1);
      float ms;
      TB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      double products = (double)blocks * threads * iters * 16.0;
      double rate = products / (ms * 1e-3);
      if (r > 0 && rate > best) best = rate;  // r == 0 is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
  } catch (const CudaError& e) {
    return (double)Fail(e);
  }
}

}  // extern "C"
