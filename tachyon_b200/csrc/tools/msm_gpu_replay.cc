// Replays MSM inputs dumped by the C API (TACHYON_MSM_GPU_INPUT_DIR) through
// tachyon_<curve>_g1_affine_msm_gpu and prints the elapsed time and the affine result.
//
// Same flags, file names and wire format as the reference's tool
//   tachyon/c/math/elliptic_curves/msm/msm_gpu_replay.cc:19-37 (readers), :40-88 (main)
//   tachyon/c/math/elliptic_curves/msm/msm_gpu.h:99-119 (writer)
// File = u64 element count, then every field element as little-endian u64 limbs in
// CANONICAL form (prime_field_base.h:198-213: ToBigInt before writing); a point is x then y
// (short_weierstrass/affine_point.h:221-223).  `--curve bls12_381` is this tool's addition.
//
//   msm_gpu_replay --idx 0,1,2 --degree 20 --input_dir /path [--curve bn254]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/tachyon_msm_b200.h"
#include "../host_math.h"

namespace {

using namespace tb200;

[[noreturn]] void Die(const std::string& msg) {
  std::cerr << msg << std::endl;
  exit(1);
}

// Reads `path` as count + count * per field elements and converts them to Montgomery form.
template <class F>
std::vector<HostFp<F>> ReadElements(const std::string& path, size_t per, size_t* count) {
  std::ifstream f(path, std::ios::binary);
  if (!f) Die("cannot open " + path);
  uint64_t n = 0;
  f.read(reinterpret_cast<char*>(&n), 8);
  if (!f) Die(path + ": truncated header");
  std::vector<HostFp<F>> out(n * per);
  HostFp<F> r2;
  memcpy(r2.v, F::kR2_64, sizeof(r2.v));
  for (auto& e : out) {
    HostFp<F> c;
    f.read(reinterpret_cast<char*>(c.v), sizeof(c.v));
    if (!f) Die(path + ": truncated body");
    e = c.Mul(r2);  // canonical -> Montgomery
  }
  if (f.peek() != std::ifstream::traits_type::eof()) Die(path + ": trailing bytes");  // buffer.Done()
  *count = n;
  return out;
}

template <class F>
std::string Hex(const HostFp<F>& montgomery) {
  HostFp<F> one = HostFp<F>::Zero();
  one.v[0] = 1;
  HostFp<F> c = montgomery.Mul(one);
  std::string s;
  char buf[17];
  for (int i = HostFp<F>::N; i-- > 0;) {
    snprintf(buf, sizeof(buf), "%016llx", (unsigned long long)c.v[i]);
    s += buf;
  }
  size_t nz = s.find_first_not_of('0');  // ToHexString(pad_zero = false)
  return "0x" + (nz == std::string::npos ? std::string("0") : s.substr(nz));
}

std::vector<int> ParseIdx(const std::string& v) {
  std::vector<int> out;
  size_t pos = 0;
  while (pos <= v.size()) {
    size_t comma = v.find(',', pos);
    if (comma == std::string::npos) comma = v.size();
    if (comma > pos) out.push_back(atoi(v.substr(pos, comma - pos).c_str()));
    pos = comma + 1;
  }
  return out;
}

template <class Fq, class Fr, class Jacobian, class Ctx, class Create, class Destroy, class Msm>
int Replay(const std::vector<int>& idxes, int degree, const std::string& dir, Create create,
           Destroy destroy, Msm msm) {
  Ctx ctx = create((uint8_t)degree);
  for (int idx : idxes) {
    size_t nb = 0, ns = 0;
    auto bases = ReadElements<Fq>(dir + "/bases" + std::to_string(idx) + ".txt", 2, &nb);
    auto scalars = ReadElements<Fr>(dir + "/scalars" + std::to_string(idx) + ".txt", 1, &ns);
    if (nb != ns) Die("bases and scalars differ in size");  // CHECK_EQ of :76
    auto t0 = std::chrono::steady_clock::now();
    std::unique_ptr<Jacobian> ret(msm(ctx, bases.data(), scalars.data(), ns));
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << s << " s" << std::endl;
    // Jacobian -> affine (jacobian_point.h:201-213): (X / Z^2, Y / Z^3), identity -> (0, 0)
    HostJacobian<Fq> j;
    memcpy(&j, ret.get(), sizeof(j));
    HostFp<Fq> x = HostFp<Fq>::Zero(), y = HostFp<Fq>::Zero();
    if (!j.z.IsZero()) {
      HostFp<Fq> zi = j.z.Inv(), zi2 = zi.Sqr();
      x = j.x.Mul(zi2);
      y = j.y.Mul(zi2.Mul(zi));
    }
    std::cout << "(" << Hex<Fq>(x) << ", " << Hex<Fq>(y) << ")" << std::endl;
  }
  destroy(ctx);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (getenv("TACHYON_MSM_GPU_INPUT_DIR")) {
    std::cerr << "If this is set, the log is overwritten" << std::endl;  // :41-44
    return 1;
  }
  std::string idx, dir, curve = "bn254";
  int degree = -1;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i], v;
    size_t eq = a.find('=');
    if (eq != std::string::npos) {
      v = a.substr(eq + 1);
      a = a.substr(0, eq);
    } else if (i + 1 < argc) {
      v = argv[++i];
    } else {
      Die("missing value for " + a);
    }
    if (a == "--idx") idx = v;
    else if (a == "--degree") degree = atoi(v.c_str());
    else if (a == "--input_dir") dir = v;
    else if (a == "--curve") curve = v;
    else Die("unknown flag " + a);
  }
  if (idx.empty() || degree < 0 || dir.empty())
    Die("usage: msm_gpu_replay --idx i[,j...] --degree k --input_dir DIR [--curve bn254|bls12_381]");
  std::vector<int> idxes = ParseIdx(idx);
  if (curve == "bn254") {
    tachyon_bn254_g1_init();
    return Replay<Bn254FqParams, Bn254FrParams, tachyon_bn254_g1_jacobian, tachyon_bn254_g1_msm_gpu_ptr>(
        idxes, degree, dir, tachyon_bn254_g1_create_msm_gpu, tachyon_bn254_g1_destroy_msm_gpu,
        [](tachyon_bn254_g1_msm_gpu_ptr c, const void* b, const void* s, size_t n) {
          return tachyon_bn254_g1_affine_msm_gpu(c, static_cast<const tachyon_bn254_g1_affine*>(b),
                                                 static_cast<const tachyon_bn254_fr*>(s), n);
        });
  }
  if (curve == "bls12_381") {
    tachyon_bls12_381_g1_init();
    return Replay<Bls381FqParams, Bls381FrParams, tachyon_bls12_381_g1_jacobian,
                  tachyon_bls12_381_g1_msm_gpu_ptr>(
        idxes, degree, dir, tachyon_bls12_381_g1_create_msm_gpu, tachyon_bls12_381_g1_destroy_msm_gpu,
        [](tachyon_bls12_381_g1_msm_gpu_ptr c, const void* b, const void* s, size_t n) {
          return tachyon_bls12_381_g1_affine_msm_gpu(
              c, static_cast<const tachyon_bls12_381_g1_affine*>(b),
              static_cast<const tachyon_bls12_381_fr*>(s), n);
        });
  }
  Die("unknown curve " + curve);
}
