// Groth16 proof assembly over this library's MSMs (the part that is MSM work):
// the five multi-scalar multiplications and the blinding arithmetic of
//   tachyon/zk/r1cs/groth16/prove.h:33-165 (CalculateCoeff, CreateProofWithAssignment)
// given a proving key, the assignments and the quotient coefficients h.  Circuit synthesis,
// the QAP witness map (FFTs) and zkey / wtns parsing stay with the caller.
//
// The four G1 MSMs (L, H, A, B1) run as one pipelined batch on the G1 context (dealt over its
// devices when it has several); the G2 MSM (B2) runs concurrently on the G2 context from a
// second host thread.
//
// This translation unit contains no kernels: the MSMs go through the C ABI of the G1 / G2
// contexts (msm_gpu_batch_b200, msm_gpu_xyzz_b200), the rest is host arithmetic.
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <thread>

#include "../../include/tachyon_msm_b200.h"
#include "groth16_files.h"
#include "host_math.h"

namespace tb200 {

extern thread_local std::string g_last_error;

struct Groth16Error {
  std::string what;
};

// T supplies the field parameter structs and the C entry points of one curve.
template <class T>
struct Groth16 {
  using E1 = HostFp<typename T::Fq>;
  using E2 = HostFp2<typename T::Fq>;
  using P1 = HostPointXYZZ<E1>;
  using P2 = HostPointXYZZ<E2>;
  using A1 = HostPointAffine<E1>;
  using A2 = HostPointAffine<E2>;
  using FrEl = HostFp<typename T::Fr>;

  struct Key {
    A1 alpha_g1, beta_g1, delta_g1;
    A2 beta_g2, delta_g2;
    const void* a_g1_query;
    size_t a_g1_size;
    const void* b_g1_query;
    size_t b_g1_size;
    const void* b_g2_query;
    size_t b_g2_size;
    const void* h_g1_query;
    size_t h_g1_size;
    const void* l_g1_query;
    size_t l_g1_size;
  };
  struct Proof {
    A1 a;
    A2 b;
    A1 c;
  };

  // First element of a query array that may live on the host or on a device.
  template <class A>
  static A Head(const void* query) {
    A out;
    if (cudaMemcpy(&out, query, sizeof(A), cudaMemcpyDefault) != cudaSuccess)
      throw Groth16Error{"groth16: cannot read the head of a query array"};
    return out;
  }

  static int Prove(typename T::G1Ctx g1, typename T::G2Ctx g2, const Key& pk, const void* r_mont,
                   const void* s_mont, const void* h, size_t h_size, const void* witness,
                   size_t witness_size, const void* full, size_t full_size, Proof* out) {
    try {
      if (pk.a_g1_size != full_size + 1 || pk.b_g1_size != full_size + 1 ||
          pk.b_g2_size != full_size + 1 || pk.l_g1_size != witness_size)
        throw Groth16Error{"groth16: query and assignment sizes differ"};
      // prove.h:100-112: h may carry one coefficient more than the query has points
      size_t h_n = h_size > pk.h_g1_size ? h_size - 1 : h_size;
      if (h_n > pk.h_g1_size)
        throw Groth16Error{"groth16: h longer than the h query"};
      FrEl one = FrEl::Zero(), r, s;
      one.v[0] = 1;
      memcpy(&r, r_mont, sizeof(r));
      memcpy(&s, s_mont, sizeof(s));
      FrEl rc = r.Mul(one), sc = s.Mul(one);  // canonical
      const bool blind = !r.IsZero();          // prove.h:135

      // ---- the MSMs: B2 on the G2 context, {L, H, A, B1} batched on the G1 context ----------
      P2 b2_acc;
      int g2_rc = 0;
      std::string g2_msg;
      std::thread g2_thread([&] {
        g2_rc = T::G2Msm(g2, static_cast<const char*>(pk.b_g2_query) + sizeof(A2), full, full_size,
                         &b2_acc);
        if (g2_rc) g2_msg = g_last_error;  // thread-local in the worker
      });
      // the blinding terms that do not depend on the MSMs are computed on a third host thread
      // while the GPU works: r delta, s delta, s (r delta) in G1 and s delta2 in G2
      P1 delta1 = FromAffine(pk.delta_g1), r_delta, s_delta, sr_delta;
      P2 s_delta2;
      std::thread pre_thread([&] {
        r_delta = ScalarMul(delta1, rc.v, FrEl::N);
        s_delta2 = ScalarMul(FromAffine(pk.delta_g2), sc.v, FrEl::N);
        if (blind) {
          s_delta = ScalarMul(delta1, sc.v, FrEl::N);
          sr_delta = ScalarMul(r_delta, sc.v, FrEl::N);
        }
      });
      const void* bases[4] = {pk.l_g1_query, pk.h_g1_query,
                              static_cast<const char*>(pk.a_g1_query) + sizeof(A1),
                              static_cast<const char*>(pk.b_g1_query) + sizeof(A1)};
      const void* scalars[4] = {witness, h, full, full};
      size_t sizes[4] = {witness_size, h_n, full_size, full_size};
      P1 acc[4];
      int g1_rc = T::G1Batch(g1, bases, scalars, sizes, blind ? 4 : 3, acc);
      g2_thread.join();
      pre_thread.join();
      if (g1_rc) return g1_rc;  // g_last_error already set on this thread
      if (g2_rc) {
        g_last_error = g2_msg;
        return g2_rc;
      }
      const P1 &witness_acc = acc[0], &h_acc = acc[1];

      // ---- assembly (prove.h:113-160) -------------------------------------------------------
      // [A]1 = r delta + a_query[0] + sum x_i a_i + alpha        (CalculateCoeff, :33-52)
      P1 a = r_delta.Add(FromAffine(Head<A1>(pk.a_g1_query))).Add(acc[2]).Add(FromAffine(pk.alpha_g1));
      // [B]2 = s delta2 + b2_query[0] + sum x_i b_i + beta2
      P2 b2 = s_delta2.Add(FromAffine(Head<A2>(pk.b_g2_query))).Add(b2_acc).Add(FromAffine(pk.beta_g2));
      // [C]1 = s A (+ r B1 - s (r delta)) + witness_acc + h_acc
      P1 c = ScalarMul(a, sc.v, FrEl::N);
      if (blind) {
        P1 b1 = s_delta.Add(FromAffine(Head<A1>(pk.b_g1_query))).Add(acc[3]).Add(FromAffine(pk.beta_g1));
        c = c.Add(ScalarMul(b1, rc.v, FrEl::N));
        P1 sub = sr_delta;
        sub.y = E1::Zero().Sub(sub.y);
        c = c.Add(sub);
      }
      c = c.Add(witness_acc).Add(h_acc);
      P1 ac[2] = {a, c};
      A1 ac_aff[2];
      BatchNormalize(ac, 2, ac_aff);  // :156-157
      A2 b_aff;
      BatchNormalize(&b2, 1, &b_aff);
      out->a = ac_aff[0];
      out->b = b_aff;
      out->c = ac_aff[1];
      return 0;
    } catch (const Groth16Error& e) {
      g_last_error = e.what;
      return -1;
    }
  }
};

// Proof from a .zkey and a .wtns file (vendors/circom/prover_main.cc:81-186 CreateProof): parse
// both, run the QAP witness map on the host, hand the five MSMs to the contexts (query sections
// zero-copy out of the mapped file), write the snarkjs JSON files when asked to.
template <class T>
static int ProveFromFiles(typename T::G1Ctx g1, typename T::G2Ctx g2, const char* zkey_path,
                          const char* wtns_path, const void* r_mont, const void* s_mont,
                          typename Groth16<T>::Proof* out, const char* proof_json,
                          const char* public_json) {
  using G = Groth16<T>;
  using FrEl = typename G::FrEl;
  try {
    MappedFile zfile(zkey_path), wfile(wtns_path);
    Zkey<typename T::Fq, typename T::Fr> zk(zfile);
    std::vector<FrEl> full = ReadWitness<typename T::Fr>(wfile);
    if (full.size() != zk.num_vars) throw FileError{"wtns: witness count differs from the zkey's variable count"};
    std::vector<FrEl> h = WitnessMap(zk, full, T::kFrGenerator);
    typename G::Key pk;
    memcpy(&pk.alpha_g1, zk.alpha_g1, sizeof(pk.alpha_g1));
    memcpy(&pk.beta_g1, zk.beta_g1, sizeof(pk.beta_g1));
    memcpy(&pk.delta_g1, zk.delta_g1, sizeof(pk.delta_g1));
    memcpy(&pk.beta_g2, zk.beta_g2, sizeof(pk.beta_g2));
    memcpy(&pk.delta_g2, zk.delta_g2, sizeof(pk.delta_g2));
    pk.a_g1_query = zk.a_g1;
    pk.a_g1_size = zk.num_vars;
    pk.b_g1_query = zk.b_g1;
    pk.b_g1_size = zk.num_vars;
    pk.b_g2_query = zk.b_g2;
    pk.b_g2_size = zk.num_vars;
    pk.h_g1_query = zk.h_g1;
    pk.h_g1_size = zk.domain_size;
    pk.l_g1_query = zk.c_g1;
    pk.l_g1_size = zk.num_vars - zk.num_public - 1;
    FrEl zero = FrEl::Zero();
    const size_t n_inst = (size_t)zk.num_public + 1;  // instance variables incl. the constant 1
    // prover_main.cc:150-165: instance = full[1, n_inst), witness = full[n_inst, ..), full[1, ..)
    int rc = G::Prove(g1, g2, pk, r_mont ? r_mont : &zero, s_mont ? s_mont : &zero, h.data(), h.size(),
                      full.data() + n_inst, full.size() - n_inst, full.data() + 1, full.size() - 1, out);
    if (rc) return rc;
    if (proof_json) WriteTextFile(proof_json, ProofJson<typename T::Fq>(out->a, out->b, out->c, T::kSnarkjsName));
    if (public_json) WriteTextFile(public_json, PublicJson<typename T::Fr>(full.data() + 1, n_inst - 1));
    return 0;
  } catch (const FileError& e) {
    g_last_error = e.what;
    return -1;
  }
}

// Host-only half of the above, for parity tests without a GPU: the h scalars (Montgomery form).
template <class T>
static int WitnessMapFromFiles(const char* zkey_path, const char* wtns_path, void* h_out,
                               size_t capacity, size_t* domain_size, size_t* num_public) {
  try {
    MappedFile zfile(zkey_path), wfile(wtns_path);
    Zkey<typename T::Fq, typename T::Fr> zk(zfile);
    auto full = ReadWitness<typename T::Fr>(wfile);
    if (full.size() != zk.num_vars) throw FileError{"wtns: witness count differs from the zkey's variable count"};
    if (domain_size) *domain_size = zk.domain_size;
    if (num_public) *num_public = zk.num_public;
    if (!h_out) return 0;
    if (capacity < zk.domain_size) throw FileError{"h buffer too small"};
    auto h = WitnessMap(zk, full, T::kFrGenerator);
    memcpy(h_out, h.data(), h.size() * sizeof(h[0]));
    return 0;
  } catch (const FileError& e) {
    g_last_error = e.what;
    return -1;
  }
}

}  // namespace tb200

using namespace tb200;

#define TB200_DEFINE_GROTH16(CN, FQ, FR, FR_GENERATOR, SNARKJS_NAME)                            \
  struct Groth16Traits_##CN {                                                                    \
    using Fq = FQ;                                                                               \
    using Fr = FR;                                                                               \
    static constexpr uint32_t kFrGenerator = FR_GENERATOR; /* multiplicative generator of Fr */  \
    static constexpr const char* kSnarkjsName = SNARKJS_NAME; /* groth16_proof.h:32-39 */        \
    using G1Ctx = tachyon_##CN##_g1_msm_gpu_ptr;                                                 \
    using G2Ctx = tachyon_##CN##_g2_msm_gpu_ptr;                                                 \
    static int G1Batch(G1Ctx c, const void* const* bases, const void* const* scalars,            \
                       const size_t* sizes, size_t count, void* out) {                           \
      return tachyon_##CN##_g1_msm_gpu_batch_b200(                                               \
          c, reinterpret_cast<const tachyon_##CN##_g1_affine* const*>(bases),                    \
          reinterpret_cast<const tachyon_##CN##_fr* const*>(scalars), sizes, count,              \
          static_cast<tachyon_##CN##_g1_xyzz*>(out));                                            \
    }                                                                                            \
    static int G2Msm(G2Ctx c, const void* bases, const void* scalars, size_t n, void* out) {     \
      return tachyon_##CN##_g2_msm_gpu_xyzz_b200(                                                \
          c, static_cast<const tachyon_##CN##_g2_affine*>(bases),                                \
          static_cast<const tachyon_##CN##_fr*>(scalars), n,                                     \
          static_cast<tachyon_##CN##_g2_xyzz*>(out));                                            \
    }                                                                                            \
  };                                                                                             \
  extern "C" int tachyon_##CN##_groth16_prove_b200(                                              \
      tachyon_##CN##_g1_msm_gpu_ptr g1, tachyon_##CN##_g2_msm_gpu_ptr g2,                        \
      const tachyon_##CN##_groth16_proving_key_b200* pk, const tachyon_##CN##_fr* r,             \
      const tachyon_##CN##_fr* s, const tachyon_##CN##_fr* h_coefficients, size_t h_size,        \
      const tachyon_##CN##_fr* witness_assignments, size_t witness_size,                         \
      const tachyon_##CN##_fr* full_assignments, size_t full_size,                               \
      tachyon_##CN##_groth16_proof_b200* out) {                                                  \
    if (!g1 || !g2 || !pk || !r || !s || !out) return -1;                                        \
    using G = Groth16<Groth16Traits_##CN>;                                                       \
    static_assert(sizeof(G::Key) == sizeof(*pk), "proving key layout");                          \
    static_assert(sizeof(G::Proof) == sizeof(*out), "proof layout");                             \
    return G::Prove(g1, g2, *reinterpret_cast<const G::Key*>(pk), r, s, h_coefficients, h_size,  \
                    witness_assignments, witness_size, full_assignments, full_size,              \
                    reinterpret_cast<G::Proof*>(out));                                           \
  }                                                                                              \
  extern "C" int tachyon_##CN##_groth16_prove_from_files_b200(                                   \
      tachyon_##CN##_g1_msm_gpu_ptr g1, tachyon_##CN##_g2_msm_gpu_ptr g2, const char* zkey_path, \
      const char* wtns_path, const tachyon_##CN##_fr* r, const tachyon_##CN##_fr* s,             \
      tachyon_##CN##_groth16_proof_b200* out, const char* proof_json_path,                       \
      const char* public_json_path) {                                                            \
    if (!g1 || !g2 || !zkey_path || !wtns_path || !out) return -1;                               \
    using G = Groth16<Groth16Traits_##CN>;                                                       \
    return ProveFromFiles<Groth16Traits_##CN>(g1, g2, zkey_path, wtns_path, r, s,                \
                                              reinterpret_cast<G::Proof*>(out), proof_json_path, \
                                              public_json_path);                                 \
  }                                                                                              \
  extern "C" int tachyon_##CN##_groth16_witness_map_from_files_b200(                             \
      const char* zkey_path, const char* wtns_path, tachyon_##CN##_fr* h_out, size_t capacity,   \
      size_t* domain_size, size_t* num_public) {                                                 \
    if (!zkey_path || !wtns_path) return -1;                                                     \
    return WitnessMapFromFiles<Groth16Traits_##CN>(zkey_path, wtns_path, h_out, capacity,        \
                                                   domain_size, num_public);                     \
  }

TB200_DEFINE_GROTH16(bn254, Bn254FqParams, Bn254FrParams, 5, "bn128")
TB200_DEFINE_GROTH16(bls12_381, Bls381FqParams, Bls381FrParams, 7, "bls12381")
