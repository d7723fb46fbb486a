// Bare pinned host -> device copy rate of ONE GPU while k - 1 other processes do the same on
// their GPUs: separates a platform limit (host memory / PCIe root complex shared by the GPUs of a
// box) from anything this library does when `e2e.h2d_ms` of bench.py drops at N >= 4.
//   nvcc -O2 -o tools/probe/h2d_probe tools/probe/h2d_probe.cu
//   tools/probe/h2d_probe <device> <MiB per copy> <copies> <mode> [start_epoch_s]
// mode 0: cudaMallocHost; 1: cudaHostAlloc write-combined; 2: malloc + cudaHostRegister (first
// touched by this process, so it sits on the NUMA node the process runs on).
// Prints one line: device, mode, GB/s of the best and of the mean copy.  tools/h2d_probe.sh
// starts k copies at the same wall-clock second.
#include <cuda_runtime.h>
#include <sys/time.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                  \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main(int argc, char** argv) {
  int dev = argc > 1 ? atoi(argv[1]) : 0;
  size_t mib = argc > 2 ? (size_t)atol(argv[2]) : 256;
  int copies = argc > 3 ? atoi(argv[3]) : 20;
  int mode = argc > 4 ? atoi(argv[4]) : 0;
  double start = argc > 5 ? atof(argv[5]) : 0;
  size_t bytes = mib << 20;
  CK(cudaSetDevice(dev));
  void *h = nullptr, *d = nullptr;
  if (mode == 0) {
    CK(cudaMallocHost(&h, bytes));
  } else if (mode == 1) {
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocWriteCombined));
  } else {
    h = aligned_alloc(4096, bytes);
    memset(h, 1, bytes);
    CK(cudaHostRegister(h, bytes, cudaHostRegisterDefault));
  }
  if (mode != 2) memset(h, 1, bytes);
  CK(cudaMalloc(&d, bytes));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));  // warm-up
  CK(cudaStreamSynchronize(s));
  if (start > 0) {  // all processes start their timed copies at the same wall-clock instant
    for (;;) {
      timeval tv;
      gettimeofday(&tv, nullptr);
      if (tv.tv_sec + tv.tv_usec * 1e-6 >= start) break;
      usleep(200);
    }
  }
  double best = 0, sum = 0;
  for (int i = 0; i < copies; ++i) {
    CK(cudaEventRecord(e0, s));
    CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double gbs = bytes / (ms * 1e-3) / 1e9;
    sum += gbs;
    if (gbs > best) best = gbs;
  }
  printf("device %d mode %d: best %.1f GB/s, mean %.1f GB/s (%zu MiB x %d)\n", dev, mode, best, sum / copies,
         mib, copies);
  return 0;
}
