// Shared by the translation units of the C ABI (one per curve and group, so that nvcc
// compiles them in parallel): error plumbing, the context behind tachyon_<c>_<g>_msm_gpu_ptr
// (tachyon/c/math/elliptic_curves/msm/msm_gpu.h:22-122), host-side helpers and the macros
// that stamp out the entry points of include/tachyon_msm_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tachyon_msm_b200.h"
#include "msm_engine.cuh"

namespace tb200 {

extern thread_local std::string g_last_error;

static int Fail(const CudaError& e) {
  char buf[512];
  if (e.device >= 0)
    snprintf(buf, sizeof(buf), "%s:%d: device %d: %s -> %s (%d)", e.file, e.line, e.device, e.what,
             cudaGetErrorString(e.code), (int)e.code);
  else
    snprintf(buf, sizeof(buf), "%s:%d: %s -> %s (%d)", e.file, e.line, e.what,
             cudaGetErrorString(e.code), (int)e.code);
  g_last_error = buf;
  return -(int)(e.code ? e.code : 1);
}

[[noreturn]] static void Die(const CudaError& e) {
  // The reference CHECK()s at the C boundary (msm_gpu.h:79, :59-62): abort loudly.
  Fail(e);
  fprintf(stderr, "tachyon_msm_b200: fatal: %s\n", g_last_error.c_str());
  abort();
}

// Context behind tachyon_<curve>_g1_msm_gpu_ptr (MSMGpuApi of msm_gpu.h:22-67).
template <class C>
struct MsmGpuContext {
  using Engine = MsmEngine<C>;
  using Point = typename Engine::Point;

  std::vector<std::unique_ptr<Engine>> engines;  // [0] = primary device
  int primary_device = 0;
  MsmTiming timing;
  std::string input_dir;  // TACHYON_MSM_GPU_INPUT_DIR (msm_gpu.h:44-48)
  bool log_msm = false;   // TACHYON_LOG_MSM          (msm_gpu.h:49-52)
  size_t idx = 0;

  explicit MsmGpuContext(int device) : primary_device(device) {
    engines.emplace_back(new Engine(device));
  }

  void SetDevices(int k) {
    int avail = 0;
    TB_CUDA(cudaGetDeviceCount(&avail));
    if (k < 1 || k > avail) throw CudaError{cudaErrorInvalidDevice, "devices option", __FILE__, __LINE__};
    while ((int)engines.size() > k) engines.pop_back();
    while ((int)engines.size() < k) {
      int dev = (primary_device + (int)engines.size()) % avail;
      try {
        engines.emplace_back(new Engine(dev));
      } catch (CudaError& e) {
        e.device = dev;
        throw;
      }
      engines.back()->options() = engines[0]->options();
      // bases registered before the device count was raised follow the new engines (the order
      // of register_bases and set_option("devices") does not matter to the caller)
      if (engines[0]->registered_size())
        engines.back()->RegisterBases(engines[0]->registered_bases(), engines[0]->registered_size());
    }
    TB_CUDA(cudaSetDevice(primary_device));
  }

  // The same bases on every engine's device (kzg.h:91-113: the SRS is uploaded once).
  void RegisterBases(const void* bases, size_t n) {
    for (auto& e : engines) e->RegisterBases(bases, n);
    TB_CUDA(cudaSetDevice(primary_device));
  }

  // `count` MSMs, MSM i over bases[i] (nullptr: the registered bases) — the commit loop of
  // kzg.h:217-313, or the four G1 queries of groth16/prove.h:100-131 with one base set each.
  // With k devices MSM i runs on device i mod k — whole MSMs are independent, so they are
  // dealt out rather than sharded; each device keeps two of its MSMs in flight.  Explicit
  // device pointers belong to the primary device, so they confine the batch to it.
  void RunBatch(const void* const* bases, const void* const* scalars, const size_t* sizes,
                size_t count, Point* out) {
    size_t G = engines.size();
    for (size_t i = 0; i < count && G > 1; ++i) {
      cudaPointerAttributes a;
      for (const void* p : {bases[i], scalars[i]}) {
        if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice)
          G = 1;
        cudaGetLastError();
      }
    }
    if (G == 1 || count == 1) {
      engines[0]->RunBatch(bases, scalars, sizes, count, out);
      timing = engines[0]->timing();
      return;
    }
    std::vector<CudaError> errs(G, CudaError{cudaSuccess, "", "", 0});
    std::vector<std::thread> threads;
    auto wall0 = std::chrono::steady_clock::now();
    for (size_t g = 0; g < G; ++g) {
      threads.emplace_back([&, g] {
        std::vector<const void*> bs, sc;
        std::vector<size_t> sz, idx;
        for (size_t i = g; i < count; i += G) {
          bs.push_back(bases[i]);
          sc.push_back(scalars[i]);
          sz.push_back(sizes[i]);
          idx.push_back(i);
        }
        std::vector<Point> res(sc.size());
        try {
          engines[g]->RunBatch(bs.data(), sc.data(), sz.data(), sc.size(), res.data());
          for (size_t k = 0; k < idx.size(); ++k) out[idx[k]] = res[k];
        } catch (const CudaError& e) {
          errs[g] = e;
          errs[g].device = engines[g]->device();
        }
      });
    }
    for (auto& t : threads) t.join();
    for (auto& e : errs)
      if (e.code != cudaSuccess) throw e;
    timing = MsmTiming{};
    for (size_t g = 0; g < G; ++g) {
      const MsmTiming& t = engines[g]->timing();
      timing.sort_ms = std::max(timing.sort_ms, t.sort_ms);
      timing.accumulate_ms = std::max(timing.accumulate_ms, t.accumulate_ms);
      timing.reduce_ms = std::max(timing.reduce_ms, t.reduce_ms);
      timing.h2d_ms = std::max(timing.h2d_ms, t.h2d_ms);
      timing.host_ms = std::max(timing.host_ms, t.host_ms);
      timing.window_bits = t.window_bits;
      timing.windows = t.windows;
      timing.tasks += t.tasks;
      timing.entries += t.entries;
      timing.kernel_launches += t.kernel_launches;
      timing.ranges = std::max(timing.ranges, t.ranges);
      timing.acc_kernel_ms = std::max(timing.acc_kernel_ms, t.acc_kernel_ms);
      timing.acc_kernel_entries += t.acc_kernel_entries;
    }
    timing.total_ms =
        std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    timing.devices = (uint32_t)G;
    TB_CUDA(cudaSetDevice(primary_device));
  }

  // Point-range sharding (the split of pippenger_adapter.h:82-113 across GPUs
  // instead of threads): device g takes [g*n/G, (g+1)*n/G); partial sums are
  // added on the host.  Device-resident inputs stay on the primary device.
  Point Run(const void* bases, const void* scalars, size_t n) {
    size_t G = engines.size();
    bool host_inputs = true;
    {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, bases) == cudaSuccess &&
          (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged))
        host_inputs = false;
      else
        cudaGetLastError();
      if (cudaPointerGetAttributes(&a, scalars) == cudaSuccess &&
          (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged))
        host_inputs = false;
      else
        cudaGetLastError();
    }
    if (G == 1 || !host_inputs || n < 2 * G) {
      Point r = engines[0]->Run(bases, scalars, n);
      timing = engines[0]->timing();
      return r;
    }
    std::vector<Point> parts(G);
    std::vector<CudaError> errs(G, CudaError{cudaSuccess, "", "", 0});
    std::vector<std::thread> threads;
    for (size_t g = 0; g < G; ++g) {
      threads.emplace_back([&, g] {
        size_t lo = n * g / G, hi = n * (g + 1) / G;
        try {
          parts[g] = engines[g]->Run(
              static_cast<const char*>(bases) + lo * Engine::kAffineBytes,
              static_cast<const char*>(scalars) + lo * Engine::kScalarBytes, hi - lo);
        } catch (const CudaError& e) {
          errs[g] = e;
          errs[g].device = engines[g]->device();
        }
      });
    }
    for (auto& t : threads) t.join();
    for (auto& e : errs)
      if (e.code != cudaSuccess) throw e;
    Point total = parts[0];
    for (size_t g = 1; g < G; ++g) total = total.Add(parts[g]);
    timing = MsmTiming{};
    for (size_t g = 0; g < G; ++g) {
      const MsmTiming& t = engines[g]->timing();
      timing.h2d_ms = std::max(timing.h2d_ms, t.h2d_ms);
      timing.sort_ms = std::max(timing.sort_ms, t.sort_ms);
      timing.accumulate_ms = std::max(timing.accumulate_ms, t.accumulate_ms);
      timing.reduce_ms = std::max(timing.reduce_ms, t.reduce_ms);
      timing.total_ms = std::max(timing.total_ms, t.total_ms);
      timing.host_ms = std::max(timing.host_ms, t.host_ms);
      timing.window_bits = t.window_bits;
      timing.windows = t.windows;
      timing.tasks += t.tasks;
      timing.entries += t.entries;
      timing.kernel_launches += t.kernel_launches;
      timing.ranges = std::max(timing.ranges, t.ranges);
      timing.acc_kernel_ms = std::max(timing.acc_kernel_ms, t.acc_kernel_ms);
      timing.acc_kernel_entries += t.acc_kernel_entries;
      timing.low_windows = t.low_windows;
      timing.combine_ms = std::max(timing.combine_ms, t.combine_ms);
    }
    timing.devices = (uint32_t)G;
    TB_CUDA(cudaSetDevice(primary_device));
    return total;
  }
};

template <class F>
static std::string FpHex(const HostFp<F>& montgomery) {
  // canonical value, big-endian hex (what ToHexString prints)
  HostFp<F> one = HostFp<F>::Zero();
  one.v[0] = 1;
  HostFp<F> c = montgomery.Mul(one);
  std::string s;
  char buf[17];
  for (int i = HostFp<F>::N; i-- > 0;) {
    snprintf(buf, sizeof(buf), "%016llx", (unsigned long long)c.v[i]);
    s += buf;
  }
  size_t nz = s.find_first_not_of('0');  // ToHexString(pad_zero = false), big_int.cc:55-58
  return "0x" + (nz == std::string::npos ? std::string("0") : s.substr(nz));
}

// Dump format of msm_gpu.h:99-119 / msm_gpu_replay.cc:19-37: u64 count, then
// every field element as canonical little-endian u64 limbs.
template <class F>
static std::string ElHex(const HostFp<F>& e) {
  return FpHex<F>(e);
}
template <class F>
static std::string ElHex(const HostFp2<F>& e) {
  return "(" + FpHex<F>(e.c0) + ", " + FpHex<F>(e.c1) + ")";
}

template <class C>
static void MaybeLogAndDump(MsmGpuContext<C>& ctx, const void* bases, const void* scalars,
                            size_t size, const HostPointJacobian<HostElT<C>>& jac) {
  using Fq = typename C::Fq;
  using Fr = typename C::Fr;
  if (ctx.log_msm) {
    std::cout << "\033[33mDoMSMGpu()" << ctx.idx << "\033[0m" << std::endl;
    std::cout << "(" << ElHex(jac.x) << ", " << ElHex(jac.y) << ", " << ElHex(jac.z) << ")"
              << std::endl;
  }
  ctx.idx++;
  if (!ctx.input_dir.empty()) {
    cudaPointerAttributes a;
    bool dev = cudaPointerGetAttributes(&a, bases) == cudaSuccess && a.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (dev) return;  // device-resident SRS is not dumped
    auto write = [&](const char* stem, auto tag, const void* data, size_t per) {
      using F = decltype(tag);
      std::string path = ctx.input_dir + "/" + stem + std::to_string(ctx.idx - 1) + ".txt";
      {
        std::ofstream f(path, std::ios::binary);
        uint64_t count = size;
        f.write(reinterpret_cast<const char*>(&count), 8);
      }
      std::ofstream f(path, std::ios::binary | std::ios::app);
      HostFp<F> one = HostFp<F>::Zero();
      one.v[0] = 1;
      const uint64_t* src = static_cast<const uint64_t*>(data);
      for (size_t i = 0; i < size * per; ++i) {
        HostFp<F> m;
        memcpy(m.v, src + i * HostFp<F>::N, sizeof(m.v));
        HostFp<F> c = m.Mul(one);
        f.write(reinterpret_cast<const char*>(c.v), sizeof(c.v));
      }
    };
    write("bases", Fq{}, bases, 2 * C::Field::kDegree);
    write("scalars", Fr{}, scalars, 1);
  }
}

template <class C, class CJacobian>
static CJacobian* DoMsmGpu(MsmGpuContext<C>* ctx, const void* bases, const void* scalars,
                           size_t size) {
  try {
    auto sum = ctx->Run(bases, scalars, size);
    HostPointJacobian<HostElT<C>> jac = ToJacobian(sum);
    static_assert(sizeof(CJacobian) == sizeof(jac), "layout");
    CJacobian* ret = new CJacobian();  // caller deletes (msm_gpu.h:81)
    memcpy(ret, &jac, sizeof(jac));
    MaybeLogAndDump<C>(*ctx, bases, scalars, size, jac);
    return ret;
  } catch (const CudaError& e) {
    Die(e);
  }
}

template <class Ctx>
static Ctx* CreateContext(int device, bool banner, int degree = 0) {
  if (banner) {
    // msm_gpu.h:36-42
    std::cout << "\033[32mCreateMSMGpuApi()\033[0m" << std::endl;
  }
  std::unique_ptr<Ctx> ctx(new Ctx(device));
  if (const char* d = getenv("TACHYON_MSM_GPU_INPUT_DIR")) ctx->input_dir = d;
  if (const char* l = getenv("TACHYON_LOG_MSM")) ctx->log_msm = std::string(l) == "1";
  if (const char* w = getenv("TACHYON_B200_MSM_WINDOW_BITS"))
    ctx->engines[0]->options().window_bits = (uint32_t)atoi(w);
  if (const char* g = getenv("TACHYON_B200_MSM_DEVICES")) ctx->SetDevices(atoi(g));
  // The reference entry point passes degree = log2(max size): reserve for it now, so the
  // first MSM (the only one the reference's benchmark times, msm_runner.h:46-61) is warm.
  // Capped at 2^24 points (5 GB) unless TACHYON_B200_MSM_PREWARM_DEGREE says otherwise; 0 = off.
  int prewarm = banner ? degree : 0;
  if (const char* pw = getenv("TACHYON_B200_MSM_PREWARM_DEGREE")) prewarm = atoi(pw);
  else if (prewarm > 24) prewarm = 24;
  if (prewarm > 0 && prewarm <= 26) {
    // Best effort: the reference ignores `degree` and allocates lazily (msm_gpu.h:35), so a busy
    // or smaller GPU must not make create() fail where the reference would not.  On
    // out-of-memory the partial reservation is released and the first MSM allocates what it
    // needs (or runs in more point ranges).
    size_t per = (size_t(1) << prewarm) / ctx->engines.size();
    for (auto& e : ctx->engines) {
      try {
        e->Prewarm(per ? per : 1);
      } catch (const CudaError& err) {
        if (err.code != cudaErrorMemoryAllocation) throw;
        cudaGetLastError();
        e->ReleaseWorkspace();
      }
    }
    TB_CUDA(cudaSetDevice(device));
  }
  return ctx.release();
}

// ---- element-wise hooks ------------------------------------------------------
template <class F>
static int FieldOpGpu(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  try {
    if (n == 0) return 0;
    size_t bytes = n * F::kLimbs64 * 8;
    uint32_t *da, *db, *dout;
    TB_CUDA(cudaMalloc(&da, bytes));
    TB_CUDA(cudaMalloc(&db, bytes));
    TB_CUDA(cudaMalloc(&dout, bytes));
    TB_CUDA(cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice));
    TB_CUDA(cudaMemcpy(db, b ? b : a, bytes, cudaMemcpyHostToDevice));
    field_op_kernel<F><<<(uint32_t)((n + 127) / 128), 128>>>(op, da, db, dout, (uint32_t)n);
    g_kernel_launches.fetch_add(1);
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return 0;
  } catch (const CudaError& e) {
    return Fail(e);
  }
}

template <class K>
static int ExtFieldOpGpu(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  try {
    if (n == 0) return 0;
    if (op < 0 || op > 6) throw CudaError{cudaErrorInvalidValue, "fq2 op", __FILE__, __LINE__};
    size_t bytes = n * K::kWords * 4;
    uint32_t *da, *db, *dout;
    TB_CUDA(cudaMalloc(&da, bytes));
    TB_CUDA(cudaMalloc(&db, bytes));
    TB_CUDA(cudaMalloc(&dout, bytes));
    TB_CUDA(cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice));
    TB_CUDA(cudaMemcpy(db, b ? b : a, bytes, cudaMemcpyHostToDevice));
    ext_field_op_kernel<K><<<(uint32_t)((n + 63) / 64), 64>>>(op, da, db, dout, (uint32_t)n);
    g_kernel_launches.fetch_add(1);
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return 0;
  } catch (const CudaError& e) {
    return Fail(e);
  }
}

template <class C>
static int PointOpGpu(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  try {
    if (n == 0) return 0;
    size_t fe = C::Field::kWords * 4;
    size_t abytes = n * 4 * fe, bbytes = n * (op == 0 ? 4 : 2) * fe;
    uint32_t *da, *db, *dout;
    TB_CUDA(cudaMalloc(&da, abytes));
    TB_CUDA(cudaMalloc(&db, bbytes));
    TB_CUDA(cudaMalloc(&dout, abytes));
    TB_CUDA(cudaMemcpy(da, a, abytes, cudaMemcpyHostToDevice));
    if (b) TB_CUDA(cudaMemcpy(db, b, bbytes, cudaMemcpyHostToDevice));
    point_op_kernel<C><<<(uint32_t)((n + 63) / 64), 64>>>(op, da, db, dout, (uint32_t)n);
    g_kernel_launches.fetch_add(1);
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpy(out, dout, abytes, cudaMemcpyDeviceToHost));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return 0;
  } catch (const CudaError& e) {
    return Fail(e);
  }
}

template <class C>
static int GenerateBases(uint64_t seed, size_t first, size_t n, void* device_out) {
  try {
    if (n == 0) return 0;
    const size_t chain = size_t(1) << kChainLog;
    if (first % chain) throw CudaError{cudaErrorInvalidValue, "first must be a multiple of 4096", __FILE__, __LINE__};
    // in slabs, so the XYZZ scratch stays bounded
    const size_t slab = size_t(1) << 22;
    uint32_t* scratch;
    size_t scratch_pts = n < slab ? n : slab;
    TB_CUDA(cudaMalloc(&scratch, scratch_pts * 4 * C::Field::kWords * 4));
    for (size_t off = 0; off < n; off += slab) {
      size_t len = n - off < slab ? n - off : slab;
      uint32_t chains = (uint32_t)((len + chain - 1) / chain);
      generate_chains_kernel<C><<<(chains + 31) / 32, 32>>>(
          seed, (uint32_t)((first + off) / chain), chains, (uint32_t)len, scratch);
      TB_CUDA(cudaGetLastError());
      normalize_kernel<C><<<(uint32_t)((len + 127) / 128), 128>>>(
          scratch, (uint32_t)len,
          reinterpret_cast<uint32_t*>(static_cast<char*>(device_out) + off * 2 * C::Field::kWords * 4));
      TB_CUDA(cudaGetLastError());
      g_kernel_launches.fetch_add(2);
    }
    TB_CUDA(cudaDeviceSynchronize());
    cudaFree(scratch);
    return 0;
  } catch (const CudaError& e) {
    return Fail(e);
  }
}

template <class C>
static int GenerateScalars(uint64_t seed, int dist, size_t first, size_t n, void* device_out) {
  try {
    if (n == 0) return 0;
    generate_scalars_kernel<C><<<(uint32_t)((n + 255) / 256), 256>>>(
        seed, dist, (uint64_t)first, (uint32_t)n, static_cast<uint32_t*>(device_out));
    g_kernel_launches.fetch_add(1);
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaDeviceSynchronize());
    return 0;
  } catch (const CudaError& e) {
    return Fail(e);
  }
}

}  // namespace tb200

using namespace tb200;

// One translation unit per (curve, group): the context type and its entry points.
#define TB200_INSTANTIATE_GROUP(CN, G, CURVE)                               \
  struct tachyon_##CN##_##G##_msm_gpu : public MsmGpuContext<CURVE> {       \
    using MsmGpuContext<CURVE>::MsmGpuContext;                              \
  };                                                                        \
  extern "C" {                                                              \
  TB200_DEFINE_GROUP_API(CN, G, CURVE)                                      \
  }

#define TB200_DEFINE_FIELD_API(CN, CURVE)                                                      \
  int tachyon_##CN##_fq_op_b200(int op, const uint64_t* a, const uint64_t* b, uint64_t* out,   \
                                size_t n) {                                                    \
    return FieldOpGpu<CURVE::Fq>(op, a, b, out, n);                                            \
  }                                                                                            \
  int tachyon_##CN##_fr_op_b200(int op, const uint64_t* a, const uint64_t* b, uint64_t* out,   \
                                size_t n) {                                                    \
    return FieldOpGpu<CURVE::Fr>(op, a, b, out, n);                                            \
  }                                                                                            \
  int tachyon_##CN##_fq2_op_b200(int op, const uint64_t* a, const uint64_t* b, uint64_t* out,  \
                                 size_t n) {                                                   \
    return ExtFieldOpGpu<Fp2Field<CURVE::Fq>>(op, a, b, out, n);                               \
  }

#define TB200_DEFINE_GROUP_API(CN, G, CURVE)                                                      \
  void tachyon_##CN##_##G##_init(void) {}                                                         \
  tachyon_##CN##_##G##_msm_gpu_ptr tachyon_##CN##_##G##_create_msm_gpu(uint8_t degree) {             \
    /* degree is advisory (the reference never reads it, msm_gpu.h:35): used to pre-reserve */ \
    try {                                                                                      \
      int dev = 0;                                                                             \
      if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;                                         \
      return CreateContext<tachyon_##CN##_##G##_msm_gpu>(dev, true, degree);                  \
    } catch (const CudaError& e) {                                                             \
      Die(e);                                                                                  \
    }                                                                                          \
  }                                                                                            \
  tachyon_##CN##_##G##_msm_gpu_ptr tachyon_##CN##_##G##_create_msm_gpu_b200(uint8_t degree,          \
                                                                     int device) {             \
    (void)degree;                                                                              \
    try {                                                                                      \
      return CreateContext<tachyon_##CN##_##G##_msm_gpu>(device, false);  \
    } catch (const CudaError& e) {                                                             \
      Fail(e);                                                                                 \
      return nullptr;                                                                          \
    }                                                                                          \
  }                                                                                            \
  void tachyon_##CN##_##G##_destroy_msm_gpu(tachyon_##CN##_##G##_msm_gpu_ptr ptr) {                  \
    delete ptr;                                            \
  }                                                                                            \
  tachyon_##CN##_##G##_jacobian* tachyon_##CN##_##G##_point2_msm_gpu(                                \
      tachyon_##CN##_##G##_msm_gpu_ptr ptr, const tachyon_##CN##_##G##_point2* bases,                \
      const tachyon_##CN##_fr* scalars, size_t size) {                                         \
    return DoMsmGpu<CURVE, tachyon_##CN##_##G##_jacobian>(ptr, bases, scalars, size);             \
  }                                                                                            \
  tachyon_##CN##_##G##_jacobian* tachyon_##CN##_##G##_affine_msm_gpu(                                \
      tachyon_##CN##_##G##_msm_gpu_ptr ptr, const tachyon_##CN##_##G##_affine* bases,                \
      const tachyon_##CN##_fr* scalars, size_t size) {                                         \
    return DoMsmGpu<CURVE, tachyon_##CN##_##G##_jacobian>(ptr, bases, scalars, size);             \
  }                                                                                            \
  int tachyon_##CN##_##G##_msm_gpu_set_stream_b200(tachyon_##CN##_##G##_msm_gpu_ptr ptr,             \
                                                void* cuda_stream) {                           \
    if (!ptr || ptr->engines.size() != 1) return -1;                                           \
    ptr->engines[0]->SetStream(static_cast<cudaStream_t>(cuda_stream));                        \
    return 0;                                                                                  \
  }                                                                                            \
  int tachyon_##CN##_##G##_msm_gpu_join_ranks_b200(tachyon_##CN##_##G##_msm_gpu_ptr ptr,          \
                                                   const void* nccl_unique_id, int rank,       \
                                                   int world) {                                \
    if (!ptr || !nccl_unique_id || world < 1 || rank < 0 || rank >= world) return -1;          \
    if (ptr->engines.size() != 1) {                                                            \
      g_last_error = "join_ranks: a context sharded over in-process devices cannot join ranks"; \
      return -1;                                                                               \
    }                                                                                          \
    try {                                                                                      \
      if (world == 1) ptr->engines[0]->LeaveRanks();                                           \
      else ptr->engines[0]->JoinRanks(nccl_unique_id, rank, world);                            \
      return 0;                                                                                \
    } catch (const CudaError& e) {                                                             \
      return Fail(e);                                                                          \
    }                                                                                          \
  }                                                                                            \
  int tachyon_##CN##_##G##_msm_gpu_set_option_b200(tachyon_##CN##_##G##_msm_gpu_ptr ptr,             \
                                                const char* name, long value) {                \
    if (!ptr || !name) return -1;                                                              \
    try {                                                                                      \
      std::string k(name);                                                                     \
      if (k == "window_bits") {                                                                \
        for (auto& e : ptr->engines) e->options().window_bits = (uint32_t)value;               \
      } else if (k == "segment") {                                                             \
        for (auto& e : ptr->engines) e->options().segment = (uint32_t)value;                   \
      } else if (k == "aggregate") {                                                           \
        for (auto& e : ptr->engines) e->options().aggregate = (int)value;                      \
      } else if (k == "sample_scalars") {                                                      \
        for (auto& e : ptr->engines) e->options().sample_scalars = (int)value;                 \
      } else if (k == "sort_mode") {                                                           \
        for (auto& e : ptr->engines) e->options().sort_mode = (int)value;                      \
      } else if (k == "release_workspace") {                                                   \
        for (auto& e : ptr->engines) e->ReleaseWorkspace();                                    \
      } else if (k == "prewarm") { /* value = number of points to reserve for */                \
        for (auto& e : ptr->engines) e->Prewarm((size_t)value / ptr->engines.size());          \
      } else if (k == "pair_rounds") {                                                         \
        for (auto& e : ptr->engines) e->options().pair_rounds = (int)value;                    \
      } else if (k == "host_ranges") {                                                         \
        for (auto& e : ptr->engines) e->options().host_ranges = (uint32_t)(value < 1 ? 1 : value); \
      } else if (k == "reduce_mode") {                                                         \
        for (auto& e : ptr->engines) e->options().reduce_mode = (int)value;                    \
      } else if (k == "level_fill") {                                                          \
        for (auto& e : ptr->engines) e->options().level_fill = (uint32_t)value;                \
      } else if (k == "balance") {                                                             \
        for (auto& e : ptr->engines) e->options().balance = (int)value;                        \
      } else if (k == "precompute") {                                                          \
        for (auto& e : ptr->engines) e->options().precompute = (int)value;                     \
      } else if (k == "reduce_inline") {                                                       \
        for (auto& e : ptr->engines) e->options().reduce_inline = (int)value;                  \
      } else if (k == "reduce_roll") {                                                         \
        for (auto& e : ptr->engines) e->options().reduce_roll = (int)value;                    \
      } else if (k == "acc_lockstep") {                                                        \
        for (auto& e : ptr->engines) e->options().acc_lockstep = (int)value;                   \
      } else if (k == "acc_variant") {                                                         \
        for (auto& e : ptr->engines) e->options().acc_variant = (int)value;                    \
      } else if (k == "stage_points") {                                                        \
        for (auto& e : ptr->engines) e->options().stage_points = (int)value;                   \
      } else if (k == "device_ladder") {                                                       \
        for (auto& e : ptr->engines) e->options().device_ladder = (int)value;                  \
      } else if (k == "low_windows") {                                                         \
        for (auto& e : ptr->engines) e->options().low_windows = (int)value;                    \
      } else if (k == "ranges") {                                                              \
        for (auto& e : ptr->engines) e->options().ranges = (uint32_t)value;                    \
      } else if (k == "devices") {                                                             \
        ptr->SetDevices((int)value);                                                           \
      } else {                                                                                 \
        g_last_error = "unknown option " + k;                                                  \
        return -1;                                                                             \
      }                                                                                        \
      return 0;                                                                                \
    } catch (const CudaError& e) {                                                             \
      return Fail(e);                                                                          \
    }                                                                                          \
  }                                                                                            \
  int tachyon_##CN##_##G##_msm_gpu_xyzz_b200(tachyon_##CN##_##G##_msm_gpu_ptr ptr,                   \
                                          const tachyon_##CN##_##G##_affine* bases,               \
                                          const tachyon_##CN##_fr* scalars, size_t size,       \
                                          tachyon_##CN##_##G##_xyzz* out) {                       \
    if (!ptr || !out) return -1;                                                               \
    try {                                                                                      \
      auto sum = ptr->Run(bases, scalars, size);                                               \
      static_assert(sizeof(*out) == sizeof(sum), "layout");                                    \
      memcpy(out, &sum, sizeof(sum));                                                          \
      return 0;                                                                                \
    } catch (const CudaError& e) {                                                             \
      return Fail(e);                                                                          \
    }                                                                                          \
  }                                                                                            \
  int tachyon_##CN##_##G##_msm_gpu_register_bases_b200(tachyon_##CN##_##G##_msm_gpu_ptr ptr,         \
                                                    const tachyon_##CN##_##G##_affine* bases,     \
                                                    size_t size) {                             \
    if (!ptr || (!bases && size)) return -1;                                                   \
    try {                                                                                      \
      ptr->RegisterBases(bases, size);                                                         \
      return 0;                                                                                \
    } catch (const CudaError& e) {                                                             \
      return Fail(e);                                                                          \
    }                                                                                          \
  }                                                                                            \
  int tachyon_##CN##_##G##_msm_gpu_commit_batch_b200(                                             \
      tachyon_##CN##_##G##_msm_gpu_ptr ptr, const tachyon_##CN##_fr* const* scalars,              \
      const size_t* sizes, size_t count, tachyon_##CN##_##G##_xyzz* out) {                        \
    if (!ptr || (count && (!scalars || !sizes || !out))) return -1;                            \
    try {                                                                                      \
      static_assert(sizeof(*out) == sizeof(MsmGpuContext<CURVE>::Point), "layout");            \
      std::vector<const void*> none(count, nullptr);                                           \
      ptr->RunBatch(none.data(), reinterpret_cast<const void* const*>(scalars), sizes, count,  \
                    reinterpret_cast<MsmGpuContext<CURVE>::Point*>(out));                      \
      return 0;                                                                                \
    } catch (const CudaError& e) {                                                             \
      return Fail(e);                                                                          \
    }                                                                                          \
  }                                                                                            \
  int tachyon_##CN##_##G##_msm_gpu_batch_b200(                                                    \
      tachyon_##CN##_##G##_msm_gpu_ptr ptr, const tachyon_##CN##_##G##_affine* const* bases,         \
      const tachyon_##CN##_fr* const* scalars, const size_t* sizes, size_t count,              \
      tachyon_##CN##_##G##_xyzz* out) {                                                           \
    if (!ptr || (count && (!bases || !scalars || !sizes || !out))) return -1;                  \
    try {                                                                                      \
      ptr->RunBatch(reinterpret_cast<const void* const*>(bases),                               \
                    reinterpret_cast<const void* const*>(scalars), sizes, count,               \
                    reinterpret_cast<MsmGpuContext<CURVE>::Point*>(out));                      \
      return 0;                                                                                \
    } catch (const CudaError& e) {                                                             \
      return Fail(e);                                                                          \
    }                                                                                          \
  }                                                                                            \
  void tachyon_##CN##_##G##_xyzz_batch_normalize_b200(const tachyon_##CN##_##G##_xyzz* in, size_t n, \
                                                   tachyon_##CN##_##G##_affine* out) {            \
    BatchNormalize(reinterpret_cast<const HostPointXYZZ<HostElT<CURVE>>*>(in), n,             \
                              reinterpret_cast<HostPointAffine<HostElT<CURVE>>*>(out));                  \
  }                                                                                            \
  int tachyon_##CN##_##G##_msm_gpu_last_timing_b200(tachyon_##CN##_##G##_msm_gpu_ptr ptr,            \
                                                 tachyon_b200_msm_timing* out) {               \
    if (!ptr || !out) return -1;                                                               \
    const MsmTiming& t = ptr->timing;                                                          \
    out->h2d_ms = t.h2d_ms;                                                                    \
    out->sort_ms = t.sort_ms;                                                                  \
    out->accumulate_ms = t.accumulate_ms;                                                      \
    out->reduce_ms = t.reduce_ms;                                                              \
    out->total_ms = t.total_ms;                                                                \
    out->host_ms = t.host_ms;                                                                  \
    out->window_bits = t.window_bits;                                                          \
    out->windows = t.windows;                                                                  \
    out->tasks = t.tasks;                                                                      \
    out->entries = t.entries;                                                                  \
    out->kernel_launches = t.kernel_launches;                                                  \
    out->devices = t.devices;                                                                  \
    out->ranges = t.ranges;                                                                    \
    out->enqueue_ms = t.enqueue_ms;                                                            \
    out->wait_ms = t.wait_ms;                                                                  \
    out->pair_rounds = t.pair_rounds;                                                          \
    out->acc_kernel_ms = t.acc_kernel_ms;                                                      \
    out->acc_kernel_entries = t.acc_kernel_entries;                                            \
    out->low_windows = t.low_windows;                                                          \
    out->combine_ms = t.combine_ms;                                                            \
    return 0;                                                                                  \
  }                                                                                            \
  int tachyon_##CN##_##G##_generate_bases_b200(uint64_t seed, size_t first, size_t n,             \
                                            void* device_out) {                                \
    return GenerateBases<CURVE>(seed, first, n, device_out);                                   \
  }                                                                                            \
  int tachyon_##CN##_##G##_generate_scalars_b200(uint64_t seed, int dist, size_t first, size_t n, \
                                              void* device_out) {                              \
    return GenerateScalars<CURVE>(seed, dist, first, n, device_out);                           \
  }                                                                                            \
  int tachyon_##CN##_##G##_point_op_b200(int op, const uint64_t* a, const uint64_t* b,            \
                                      uint64_t* out, size_t n) {                               \
    return PointOpGpu<CURVE>(op, a, b, out, n);                                                \
  }                                                                                            \
  void tachyon_##CN##_##G##_xyzz_add_b200(const tachyon_##CN##_##G##_xyzz* a,                        \
                                       const tachyon_##CN##_##G##_xyzz* b,                        \
                                       tachyon_##CN##_##G##_xyzz* out) {                          \
    HostPointXYZZ<HostElT<CURVE>> x, y;                                                                  \
    memcpy(&x, a, sizeof(x));                                                                  \
    memcpy(&y, b, sizeof(y));                                                                  \
    HostPointXYZZ<HostElT<CURVE>> r = x.Add(y);                                                          \
    memcpy(out, &r, sizeof(r));                                                                \
  }                                                                                            \
  void tachyon_##CN##_##G##_xyzz_to_jacobian_b200(const tachyon_##CN##_##G##_xyzz* a,                \
                                               tachyon_##CN##_##G##_jacobian* out) {              \
    HostPointXYZZ<HostElT<CURVE>> x;                                                                     \
    memcpy(&x, a, sizeof(x));                                                                  \
    HostPointJacobian<HostElT<CURVE>> j = ToJacobian(x);                                      \
    memcpy(out, &j, sizeof(j));                                                                \
  }

