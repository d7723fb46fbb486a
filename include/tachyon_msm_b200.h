/*
 * tachyon_msm_b200.h — C ABI of the B200-native variable-base MSM engine.
 *
 * Part 1 is the drop-in boundary: exactly the symbols and struct layouts that
 * the reference generates from
 *   tachyon/c/math/elliptic_curves/generator/msm_gpu.h.tpl:17-56   (GPU MSM API)
 *   tachyon/c/math/elliptic_curves/generator/point.h.tpl:22-117    (point structs, _init)
 *   tachyon/c/math/finite_fields/generator/.../prime_field.h.tpl:36-38 (limb structs)
 * instantiated for %{type} in {bn254, bls12_381}
 * (tachyon/c/math/elliptic_curves/generator/build_defs.bzl:93-170).
 * A caller of libtachyon's MSM-GPU API (vendors/scroll_halo2/src/bn254_msm_gpu.cc:11-34,
 * benchmark/msm/msm_benchmark_gpu.cc:61-66) links against this library unchanged.
 *
 * Part 2 (prefix tachyon_b200_ / *_b200) is this engine's own extension:
 * device-resident inputs, explicit device/stream selection, stage timings,
 * synthetic test-set generation and the element-wise parity hooks.
 *
 * All field elements are little-endian u64 limbs in Montgomery form
 * (R = 2^(64*limbs)), fully reduced; the affine identity is (0, 0); a Jacobian
 * point with z == 0 is the identity.
 */
#ifndef TACHYON_MSM_B200_H_
#define TACHYON_MSM_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define TACHYON_C_EXPORT __attribute__((visibility("default")))
#else
#define TACHYON_C_EXPORT
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* Part 1 — the reference's MSM-GPU C API                                     */
/* ------------------------------------------------------------------------- */

/* Field element structs of one curve (prime_field.h.tpl:36-38; Fq2 = Fq[u]/(u^2 + 1) as in   \
   ext_field.h.tpl, c0 first) and the element-wise parity hooks on HOST arrays (copied to the  \
   current device and back).  op: 0 add 1 sub 2 mul 3 square 4 neg 5 double 6 inverse          \
   7 from_mont 8 to_mont (fq2: 0..6). */
#define TACHYON_B200_DECLARE_FIELDS(C, FQ_LIMBS)                                             \
  struct tachyon_##C##_fr { uint64_t limbs[4]; };                                            \
  struct tachyon_##C##_fq { uint64_t limbs[FQ_LIMBS]; };                                     \
  struct tachyon_##C##_fq2 { struct tachyon_##C##_fq c0, c1; };                              \
  TACHYON_C_EXPORT int tachyon_##C##_fq_op_b200(int op, const uint64_t* a, const uint64_t* b, \
                                                uint64_t* out, size_t n);                    \
  TACHYON_C_EXPORT int tachyon_##C##_fr_op_b200(int op, const uint64_t* a, const uint64_t* b, \
                                                uint64_t* out, size_t n);                    \
  TACHYON_C_EXPORT int tachyon_##C##_fq2_op_b200(int op, const uint64_t* a, const uint64_t* b, \
                                                 uint64_t* out, size_t n);

#define TACHYON_B200_DECLARE_GROUP(C, G, FE)                                              \
  struct tachyon_##C##_##G##_affine { struct FE x, y; };     /* point.h.tpl:22-25 */ \
  struct tachyon_##C##_##G##_point2 { struct FE x, y; };     /* point.h.tpl:78-81 */ \
  struct tachyon_##C##_##G##_jacobian { struct FE x, y, z; };/* point.h.tpl:50-54 */ \
  struct tachyon_##C##_##G##_xyzz { struct FE x, y, zz, zzz; };/* point.h.tpl:64-69 */ \
  typedef struct tachyon_##C##_##G##_msm_gpu* tachyon_##C##_##G##_msm_gpu_ptr; /* msm_gpu.h.tpl:17 */ \
                                                                                             \
  /* point.h.tpl:117 — idempotent, cheap; callers invoke it before anything else */         \
  TACHYON_C_EXPORT void tachyon_##C##_##G##_init(void);                                         \
  /* msm_gpu.h.tpl:26 — degree = log2(max size), advisory (reference never reads it,        \
     c/math/elliptic_curves/msm/msm_gpu.h:35).  Prints "CreateMSMGpuApi()" and honours      \
     TACHYON_MSM_GPU_INPUT_DIR / TACHYON_LOG_MSM like msm_gpu.h:36-52.  Aborts on CUDA       \
     failure (msm_gpu.h:59-62). */                                                           \
  TACHYON_C_EXPORT tachyon_##C##_##G##_msm_gpu_ptr tachyon_##C##_##G##_create_msm_gpu(uint8_t degree); \
  /* msm_gpu.h.tpl:32 */                                                                     \
  TACHYON_C_EXPORT void tachyon_##C##_##G##_destroy_msm_gpu(tachyon_##C##_##G##_msm_gpu_ptr ptr);  \
  /* msm_gpu.h.tpl:42-44.  bases/scalars: `size` elements each, host memory (pageable or    \
     pinned) or device memory (icicle_msm_bn254_g1.cc:38-45).  Returns a heap object made   \
     with C++ `new`, owned by the caller (msm_gpu.h:81).  Any failure aborts (msm_gpu.h:79). */ \
  TACHYON_C_EXPORT struct tachyon_##C##_##G##_jacobian* tachyon_##C##_##G##_point2_msm_gpu(        \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, const struct tachyon_##C##_##G##_point2* bases,         \
      const struct tachyon_##C##_fr* scalars, size_t size);                                  \
  /* msm_gpu.h.tpl:54-56 — same code path as point2 (msm_input_provider.h:23-29) */          \
  TACHYON_C_EXPORT struct tachyon_##C##_##G##_jacobian* tachyon_##C##_##G##_affine_msm_gpu(        \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, const struct tachyon_##C##_##G##_affine* bases,         \
      const struct tachyon_##C##_fr* scalars, size_t size);                                  \
                                                                                             \
  /* ---- Part 2: extensions ---------------------------------------------------------- */  \
  /* Context on a given CUDA device, silent (no banner).  NULL on failure. */               \
  TACHYON_C_EXPORT tachyon_##C##_##G##_msm_gpu_ptr tachyon_##C##_##G##_create_msm_gpu_b200(        \
      uint8_t degree, int device);                                                           \
  /* Run all work of this context on an existing CUDA stream (a cudaStream_t) of the        \
     context's device instead of its own stream.  0 on success. */                           \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_msm_gpu_set_stream_b200(                             \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, void* cuda_stream);                                  \
  /* name: "window_bits" (0 = automatic), "segment" (max entries per accumulation task),    \
     "devices" (point-range sharding over the first k devices; 1 = this context's device),  \
     "sample_scalars" (1 = choose the window from the bit lengths of 1024 sampled scalars    \
     and flag skewed vectors for the two-level sort, the default; 0 = from the size alone),  \
     "sort_mode" (0 = one-level atomic counting sort, 1 = two-level shared-memory sort       \
     where eligible, -1 = automatic),                                                        \
     "balance" (1 = balanced window widths, the default; 0 = equal widths),                  \
     "reduce_mode" (1 = two threads — G2: two lane pairs — per block of buckets in the       \
     running-sum level, the default; 2 = two threads for G2 as well; 0 = one thread),        \
     "host_ranges" (most ranges of the automatic host-input pipeline),                       \
     "prewarm" (reserve workspace, staging and copy threads for an MSM of `value` points now),\
     "release_workspace" (free the grow-only workspace; registered bases are kept),           \
     "precompute" (1 = the next register_bases call also builds the table of window           \
     multiples 2^(bit offset of window w) * P — the full precompute_factor of                 \
     icicle_msm.h:21 — so that MSMs over the registered bases use one bucket set for all      \
     windows: W times the base memory, a W times smaller bucket reduction, no ladder),        \
     "acc_variant" (G2 groups: 1 = one lane pair per accumulation task, a lane per Fq2         \
     component — the default; 2 = the same at the other register budget; 0 = one thread per   \
     task.  G1 groups: 3 = the warps of a CTA walk their tasks in step, the default for       \
     BLS12-381; 0 = free-running warps, the default for BN254),                               \
     "acc_lockstep" (G2 lane-pair kernel: 1 = warps of a CTA in step, the default; 0 = free), \
     "reduce_roll" (field multiplications of the running-sum kernel: 0 = unrolled, 1 = as a   \
     loop over row pairs, 2 = squarings through that loop too, -1 = per-curve default),       \
     "reduce_inline" (G1 running-sum kernel: 1 = both roles share one inlined call site of    \
     the point addition, the default for BN254; 0 = out-of-line addition; -1 = per curve),    \
     "stage_points" (1 = the accumulation stages the next point through shared memory with    \
     cp.async instead of registers),                                                          \
     "device_ladder" (where the final ladder over the W window sums runs — ~255 strictly     \
     sequential point doublings: 0 = on the host, the default, 65 us; 1 = as a kernel,        \
     430 us for BN254, hidden behind the accumulation of the low windows where possible),    \
     "low_windows" (device ladder: windows accumulated last while the merge tree and doubling \
     chain of the others run on a second stream; -1 = cost model, 0 = no split),              \
     "ranges" (point ranges one MSM is pipelined over; 0 = automatic), "pair_rounds"        \
     (experimental batched-affine rounds before the XYZZ accumulation; -1 = none, the       \
     default; -2 = chosen from the bucket occupancy; 0..4 = forced). */                       \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_msm_gpu_set_option_b200(                             \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, const char* name, long value);                       \
  /* Point-range sharding over `world` PROCESSES, one GPU each (the split of      \
     pippenger_adapter.h:82-113 across GPUs).  Rank 0 obtains 128 bytes from                  \
     tachyon_b200_nccl_unique_id() and distributes them by any means; after every rank has    \
     joined, each MSM call on this context takes the rank's OWN point range and returns the   \
     sum over all ranks on every rank: the ranks' XYZZ partials are exchanged with one        \
     ncclAllGather issued on the context's stream behind the last kernel, and the `world`     \
     points are added on the host.  Batch calls stay local.  world == 1 leaves.  NCCL is      \
     dlopen()ed (libnccl.so.2, or $TACHYON_B200_NCCL_LIB).  0 or a negative error code. */     \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_msm_gpu_join_ranks_b200(                             \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, const void* nccl_unique_id, int rank, int world);   \
  /* MSM returning the un-normalised XYZZ sum by value into *out; returns 0 or a negative   \
     error code instead of aborting.  Pointers as for *_affine_msm_gpu. */                   \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_msm_gpu_xyzz_b200(                                   \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, const struct tachyon_##C##_##G##_affine* bases,         \
      const struct tachyon_##C##_fr* scalars, size_t size, struct tachyon_##C##_##G##_xyzz* out); \
  /* Device-resident bases + batched commitments (the SRS handling of           \
     tachyon/crypto/commitments/kzg/kzg.h:91-113 and the commit loop of :217-313).           \
     register_bases copies `size` bases (host or device source) into memory the context      \
     owns, on every device of the context; it replaces an earlier registration.              \
     commit_batch runs `count` MSMs, MSM i over the first sizes[i] registered bases and       \
     scalars[i] (host or device), and writes the un-normalised sums to out[i]; consecutive    \
     MSMs are pipelined (scalars of the next one cross PCIe while the current one runs) and   \
     with "devices" = k they are dealt out over k GPUs.  0 or a negative error code. */       \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_msm_gpu_register_bases_b200(                         \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, const struct tachyon_##C##_##G##_affine* bases,         \
      size_t size);                                                                          \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_msm_gpu_commit_batch_b200(                           \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, const struct tachyon_##C##_fr* const* scalars,       \
      const size_t* sizes, size_t count, struct tachyon_##C##_##G##_xyzz* out);                 \
  /* The general batch: MSM i over bases[i] (NULL = the registered bases) and scalars[i],    \
     sizes[i] elements each, host or device memory — e.g. the A, B1, L and H queries of a    \
     Groth16 proof (zk/r1cs/groth16/prove.h:100-131) in one call.  Pipelined and dealt out   \
     like commit_batch (explicit device pointers keep the batch on the context's device). */ \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_msm_gpu_batch_b200(                                  \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, const struct tachyon_##C##_##G##_affine* const* bases,  \
      const struct tachyon_##C##_fr* const* scalars, const size_t* sizes, size_t count,      \
      struct tachyon_##C##_##G##_xyzz* out);                                                    \
  /* Host-only: n XYZZ points -> affine with one field inversion (point_xyzz.h:109-163       \
     BatchNormalize); the identity becomes (0, 0). */                                        \
  TACHYON_C_EXPORT void tachyon_##C##_##G##_xyzz_batch_normalize_b200(                          \
      const struct tachyon_##C##_##G##_xyzz* in, size_t n, struct tachyon_##C##_##G##_affine* out); \
  /* Stage timings of the last call on this context (CUDA events on its stream). */         \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_msm_gpu_last_timing_b200(                            \
      tachyon_##C##_##G##_msm_gpu_ptr ptr, struct tachyon_b200_msm_timing* out);                \
  /* Deterministic synthetic test set written to DEVICE memory of the current device:       \
     points first..first+n of the doubling-chain stream, scalars of distribution dist       \
     (0 uniform, 1 non_uniform, 2 witness).  Mirrors msm/test/variable_base_msm_test_set.h. */ \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_generate_bases_b200(uint64_t seed, size_t first,     \
                                                            size_t n, void* device_out);     \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_generate_scalars_b200(uint64_t seed, int dist,       \
                                                              size_t first, size_t n,        \
                                                              void* device_out);             \
  /* Point-op parity hook on HOST arrays: 0 xyzz+xyzz 1 xyzz+affine 2 xyzz-affine 3 double. */ \
  TACHYON_C_EXPORT int tachyon_##C##_##G##_point_op_b200(int op, const uint64_t* a,             \
                                                      const uint64_t* b, uint64_t* out,      \
                                                      size_t n);                             \
  /* Host-only helpers (no GPU needed): out = a + b on XYZZ points — how per-GPU / per-rank  \
     partial sums are combined (pippenger_adapter.h:110-113) — and XYZZ -> Jacobian           \
     (point_xyzz.h:228-237), the conversion applied to every MSM result. */                  \
  TACHYON_C_EXPORT void tachyon_##C##_##G##_xyzz_add_b200(const struct tachyon_##C##_##G##_xyzz* a, \
                                                       const struct tachyon_##C##_##G##_xyzz* b, \
                                                       struct tachyon_##C##_##G##_xyzz* out);   \
  TACHYON_C_EXPORT void tachyon_##C##_##G##_xyzz_to_jacobian_b200(                              \
      const struct tachyon_##C##_##G##_xyzz* a, struct tachyon_##C##_##G##_jacobian* out);

struct tachyon_b200_msm_timing {
  float h2d_ms;         /* host->device copies of bases/scalars on the copy stream (0 for device
                           inputs); overlaps sort/accumulate of the earlier point ranges */
  float sort_ms;        /* recode + histogram + scan + task build + scatter */
  float accumulate_ms;  /* bucket accumulation (+ folding of split buckets) */
  float reduce_ms;      /* what remains after the last accumulation: bucket reduction, merge
                           tree and window combination (of the low windows when split) + the
                           device->host copy of the one result point */
  float total_ms;       /* first enqueue .. last device->host copy */
  float host_ms;        /* host epilogue (reading the result point; partial adds), wall clock */
  uint32_t window_bits;
  uint32_t windows;
  uint32_t tasks;       /* accumulation tasks (threads of the hot kernel) */
  uint32_t entries;     /* non-zero digits = mixed additions performed */
  uint32_t kernel_launches; /* kernels of this library launched by the call */
  uint32_t devices;
  uint32_t ranges;      /* point ranges the call was pipelined over (1 for device inputs) */
  float enqueue_ms;     /* host wall clock spent queueing copies and kernels */
  float wait_ms;        /* host wall clock blocked waiting for the device */
  uint32_t pair_rounds; /* batched-affine pair rounds run before the XYZZ accumulation */
  float acc_kernel_ms;  /* the accumulate launches that ran alone on the device (all but the low
                           window group's, which overlaps the high group's reduction) ... */
  uint32_t acc_kernel_entries; /* ... and the mixed additions they performed */
  uint32_t low_windows; /* windows accumulated last, behind which the high windows' reduction
                           and window combination are hidden (0 = windows not split) */
  float combine_ms;     /* window_combine_kernel of the high group (or of all windows) */
};

TACHYON_B200_DECLARE_FIELDS(bn254, 4)
TACHYON_B200_DECLARE_FIELDS(bls12_381, 6)
TACHYON_B200_DECLARE_GROUP(bn254, g1, tachyon_bn254_fq)
TACHYON_B200_DECLARE_GROUP(bls12_381, g1, tachyon_bls12_381_fq)
/* G2: the reference has the g2 point structs (point.h.tpl instantiated for g2,
   build_defs.bzl:93-170) but reaches its G2 MSM only through C++
   (VariableBaseMSMGpu<G2AffinePoint>, zk/r1cs/groth16/prove.h:129-131 ->
   icicle_msm_bn254_g2.cc); the entry points below give it the same C shape as G1. */
TACHYON_B200_DECLARE_GROUP(bn254, g2, tachyon_bn254_fq2)
TACHYON_B200_DECLARE_GROUP(bls12_381, g2, tachyon_bls12_381_fq2)

/* Groth16 proof assembly over this library's MSMs (the MSM part of
   tachyon/zk/r1cs/groth16/prove.h:33-165 CreateProofWithAssignment): five MSMs — L, H, A, B1 as
   one batch on the G1 context, B2 concurrently on the G2 context — and the r / s blinding
   arithmetic.  Query arrays may be host or device memory.  Sizes: a, b1, b2 queries have
   full_size + 1 points (element 0 is added as is, prove.h:46), the l query witness_size, the h
   query at least h_size - 1 (prove.h:100-112).  r == 0 selects the non-ZK branch (prove.h:135).
   0 or a negative error code. */
#define TACHYON_B200_DECLARE_GROTH16(C)                                                       \
  struct tachyon_##C##_groth16_proving_key_b200 {                                             \
    struct tachyon_##C##_g1_affine alpha_g1, beta_g1, delta_g1;                               \
    struct tachyon_##C##_g2_affine beta_g2, delta_g2;                                         \
    const struct tachyon_##C##_g1_affine* a_g1_query; size_t a_g1_size;                       \
    const struct tachyon_##C##_g1_affine* b_g1_query; size_t b_g1_size;                       \
    const struct tachyon_##C##_g2_affine* b_g2_query; size_t b_g2_size;                       \
    const struct tachyon_##C##_g1_affine* h_g1_query; size_t h_g1_size;                       \
    const struct tachyon_##C##_g1_affine* l_g1_query; size_t l_g1_size;                       \
  };                                                                                          \
  struct tachyon_##C##_groth16_proof_b200 {                                                   \
    struct tachyon_##C##_g1_affine a;                                                         \
    struct tachyon_##C##_g2_affine b;                                                         \
    struct tachyon_##C##_g1_affine c;                                                         \
  };                                                                                          \
  TACHYON_C_EXPORT int tachyon_##C##_groth16_prove_b200(                                      \
      tachyon_##C##_g1_msm_gpu_ptr g1, tachyon_##C##_g2_msm_gpu_ptr g2,                       \
      const struct tachyon_##C##_groth16_proving_key_b200* pk, const struct tachyon_##C##_fr* r, \
      const struct tachyon_##C##_fr* s, const struct tachyon_##C##_fr* h_coefficients,        \
      size_t h_size, const struct tachyon_##C##_fr* witness_assignments, size_t witness_size, \
      const struct tachyon_##C##_fr* full_assignments, size_t full_size,                      \
      struct tachyon_##C##_groth16_proof_b200* out);                                          \
  /* The whole prover entry of vendors/circom/prover_main.cc:81-186 (CreateProof): parse a      \
     snarkjs .zkey (circomlib/zkey/zkey.h:88-317, v1; the query sections are used zero-copy   \
     out of the memory-mapped file as MSM bases, zkey.h:176-183) and a .wtns witness          \
     (circomlib/wtns/wtns.h:66-154, v2), run the QAP witness map on the host                  \
     (circomlib/circuit/quadratic_arithmetic_program.h:25-118), the five MSMs on the two      \
     contexts, and write — when the paths are not NULL — snarkjs's proof.json and public.json \
     (circomlib/json/groth16_proof.h, points.h, prime_field.h).  r, s: the blinding scalars   \
     in Montgomery form; NULL = 0, the --no_zk mode (prove.h:177-187).  0 or negative. */      \
  TACHYON_C_EXPORT int tachyon_##C##_groth16_prove_from_files_b200(                           \
      tachyon_##C##_g1_msm_gpu_ptr g1, tachyon_##C##_g2_msm_gpu_ptr g2, const char* zkey_path,  \
      const char* wtns_path, const struct tachyon_##C##_fr* r, const struct tachyon_##C##_fr* s, \
      struct tachyon_##C##_groth16_proof_b200* out, const char* proof_json_path,              \
      const char* public_json_path);                                                          \
  /* Host-only half of the above (no GPU): the files' domain size and public-input count, and \
     — when h_out is not NULL — the h scalars of the witness map in Montgomery form. */       \
  TACHYON_C_EXPORT int tachyon_##C##_groth16_witness_map_from_files_b200(                     \
      const char* zkey_path, const char* wtns_path, struct tachyon_##C##_fr* h_out,           \
      size_t capacity, size_t* domain_size, size_t* num_public);

TACHYON_B200_DECLARE_GROTH16(bn254)
TACHYON_B200_DECLARE_GROTH16(bls12_381)

/* Number of CUDA devices visible, or a negative error. */
TACHYON_C_EXPORT int tachyon_b200_device_count(void);
/* 128-byte ncclUniqueId for *_msm_gpu_join_ranks_b200 (call on one rank, hand to all). */
TACHYON_C_EXPORT int tachyon_b200_nccl_unique_id(void* out128);
/* Text of the last error recorded by an extension call on this thread. */
TACHYON_C_EXPORT const char* tachyon_b200_last_error(void);
/* Measured INT32 multiply-pipe peak of `device`: 32x32->64 multiply-adds per second
   (best of `repeats`).  variant 0 = IMAD.WIDE.U32(.X) carry chains, 1 = IMAD.WIDE.U32 with
   64-bit addend, 2 = IMAD + IMAD.HI pairs.  Negative on error. */
TACHYON_C_EXPORT double tachyon_b200_imad_peak(int device, int variant, int repeats);
/* Window size the engine picks for an n-point MSM over a scalar field of `scalar_bits`
   bits, and the matching window count (host-only). */
TACHYON_C_EXPORT uint32_t tachyon_b200_window_bits(size_t n, uint32_t scalar_bits);
TACHYON_C_EXPORT uint32_t tachyon_b200_window_count(uint32_t scalar_bits, uint32_t window_bits);
/* Total kernels launched by this library in this process (all contexts). */
TACHYON_C_EXPORT uint64_t tachyon_b200_kernel_launch_count(void);
/* Page-locked host memory for MSM inputs (cudaHostAlloc / cudaFreeHost).  write_combined = 1
   asks for write-combined pages: the CPU should only fill them front to back (reads are slow),
   but with several GPUs copying at once the DMA engines read them markedly faster than ordinary
   pinned memory (bare copy probe on an 8 x B200 box, profiles/r2_h2d_probe_8gpu.txt: 4 GPUs
   49 vs 31 GB/s each, 8 GPUs 39-49 vs 29-50).  NULL on failure. */
TACHYON_C_EXPORT void* tachyon_b200_alloc_host(size_t bytes, int write_combined);
TACHYON_C_EXPORT void tachyon_b200_free_host(void* p);

#ifdef __cplusplus
}
#endif

#endif /* TACHYON_MSM_B200_H_ */
