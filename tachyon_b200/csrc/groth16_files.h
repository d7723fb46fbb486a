// circom / snarkjs file formats around the Groth16 MSMs: the .zkey proving key,
// the .wtns witness, the QAP witness map that turns them into the h scalars, and the proof /
// public-input JSON.  What the reference does in
//   vendors/circom/circomlib/zkey/zkey.h:56-317      (ParseZKey, v1 sections)
//   vendors/circom/circomlib/wtns/wtns.h:44-160      (ParseWtns, v2 sections)
//   vendors/circom/circomlib/base/sections.h, modulus.h
//   vendors/circom/circomlib/circuit/quadratic_arithmetic_program.h:25-118 (WitnessMapFromMatrices)
//   vendors/circom/circomlib/json/{groth16_proof,points,prime_field}.h, prover_main.cc:81-186
// Host code only.  The point sections are used ZERO-COPY as MSM bases straight out of the
// memory-mapped file, as the reference does (zkey.h:176-183): they are affine points in
// Montgomery form with the layout of tachyon_<c>_g1_affine / _g2_affine.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "host_math.h"

namespace tb200 {

struct FileError {
  std::string what;
};

// Read-only memory mapping of a whole file.
class MappedFile {
 public:
  explicit MappedFile(const char* path) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) throw FileError{std::string("cannot open ") + path};
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) {
      close(fd);
      throw FileError{std::string("cannot stat ") + path};
    }
    size_ = (size_t)st.st_size;
    void* p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) throw FileError{std::string("cannot map ") + path};
    madvise(p, size_, MADV_SEQUENTIAL);  // zkey.h:62-65
    data_ = static_cast<const uint8_t*>(p);
  }
  ~MappedFile() {
    if (data_) munmap(const_cast<uint8_t*>(data_), size_);
  }
  MappedFile(const MappedFile&) = delete;
  MappedFile& operator=(const MappedFile&) = delete;
  const uint8_t* data() const { return data_; }
  size_t size() const { return size_; }

 private:
  const uint8_t* data_ = nullptr;
  size_t size_ = 0;
};

// iden3 binary container: magic[4], u32 version, u32 section count, then per section
// u32 type, u64 size, payload (sections.h:26-60).  Little endian.
class BinSections {
 public:
  BinSections(const MappedFile& f, const char magic[4], uint32_t version) : file_(f) {
    Cursor c{f.data(), f.size(), 0};
    char m[4];
    c.Read(m, 4);
    if (memcmp(m, magic, 4) != 0) throw FileError{std::string("bad magic, expected ") + std::string(magic, 4)};
    uint32_t v = c.U32();
    if (v != version) throw FileError{"unsupported file version " + std::to_string(v)};
    uint32_t count = c.U32();
    for (uint32_t i = 0; i < count; ++i) {
      uint32_t type = c.U32();
      uint64_t size = c.U64();
      if (size > f.size() - c.pos) throw FileError{"section runs past the end of the file"};
      sections_.push_back({type, c.pos, (size_t)size});
      c.pos += (size_t)size;
    }
  }
  struct Cursor {
    const uint8_t* base;
    size_t size, pos;
    void Read(void* out, size_t n) {
      if (n > size - pos) throw FileError{"truncated file"};
      memcpy(out, base + pos, n);
      pos += n;
    }
    uint32_t U32() {
      uint32_t v;
      Read(&v, 4);
      return v;
    }
    uint64_t U64() {
      uint64_t v;
      Read(&v, 8);
      return v;
    }
    // pointer to n bytes inside the mapping (zero copy)
    const uint8_t* Take(size_t n) {
      if (n > size - pos) throw FileError{"truncated section"};
      const uint8_t* p = base + pos;
      pos += n;
      return p;
    }
  };
  // cursor over the first section of the given type (sections.h:41-51 MoveTo)
  Cursor Open(uint32_t type) const {
    for (const auto& s : sections_)
      if (s.type == type) return Cursor{file_.data(), s.offset + s.size, s.offset};
    throw FileError{"section " + std::to_string(type) + " is missing"};
  }

 private:
  struct Section {
    uint32_t type;
    size_t offset, size;
  };
  const MappedFile& file_;
  std::vector<Section> sections_;
};

// Field size + modulus as stored in the headers (modulus.h:41-53); must equal the curve's.
template <class F>
static void ExpectModulus(BinSections::Cursor& c, const char* what) {
  uint32_t bytes = c.U32();
  if (bytes != (uint32_t)F::kLimbs64 * 8) throw FileError{std::string(what) + ": field size does not match the curve"};
  uint64_t limbs[F::kLimbs64];
  c.Read(limbs, bytes);
  for (int i = 0; i < F::kLimbs64; ++i)
    if (limbs[i] != F::kMod64[i]) throw FileError{std::string(what) + ": modulus does not match the curve"};
}

#pragma pack(push, 1)
template <class Fr>
struct ZkeyCoefficient {  // coefficient.h:31-46
  uint32_t matrix, constraint, signal;
  uint64_t value[Fr::kLimbs64];  // the value times R^2 (snarkjs convention; zkey.h:216-219 strips one R)
};
#pragma pack(pop)

// zkey v1 (zkey.h:88-317).  Fq / Fr: parameter structs of field_constants.h.
template <class Fq, class Fr>
struct Zkey {
  static constexpr size_t kG1 = 2 * Fq::kLimbs64 * 8, kG2 = 2 * kG1;
  uint32_t num_vars = 0, num_public = 0, domain_size = 0;
  const uint8_t *alpha_g1, *beta_g1, *beta_g2, *gamma_g2, *delta_g1, *delta_g2;  // header points
  const uint8_t *ic, *a_g1, *b_g1, *b_g2, *c_g1, *h_g1;                         // query sections
  const ZkeyCoefficient<Fr>* coefficients = nullptr;
  uint32_t num_coefficients = 0;

  explicit Zkey(const MappedFile& f) {
    BinSections s(f, "zkey", 1);
    {
      auto c = s.Open(1);  // header: prover type 1 = groth16 (zkey.h:118-126)
      if (c.U32() != 1) throw FileError{"zkey: not a groth16 key"};
    }
    auto c = s.Open(2);  // groth16 header (zkey.h:152-160) + verifying key (verifying_key.h:39-44)
    ExpectModulus<Fq>(c, "zkey q");
    ExpectModulus<Fr>(c, "zkey r");
    num_vars = c.U32();
    num_public = c.U32();
    domain_size = c.U32();
    if (num_vars < num_public + 1 || domain_size == 0 || (domain_size & (domain_size - 1)))
      throw FileError{"zkey: inconsistent header"};
    alpha_g1 = c.Take(kG1);
    beta_g1 = c.Take(kG1);
    beta_g2 = c.Take(kG2);
    gamma_g2 = c.Take(kG2);
    delta_g1 = c.Take(kG1);
    delta_g2 = c.Take(kG2);
    ic = s.Open(3).Take((size_t)(num_public + 1) * kG1);
    {
      auto cc = s.Open(4);
      num_coefficients = cc.U32();
      coefficients = reinterpret_cast<const ZkeyCoefficient<Fr>*>(
          cc.Take((size_t)num_coefficients * sizeof(ZkeyCoefficient<Fr>)));
    }
    a_g1 = s.Open(5).Take((size_t)num_vars * kG1);
    b_g1 = s.Open(6).Take((size_t)num_vars * kG1);
    b_g2 = s.Open(7).Take((size_t)num_vars * kG2);
    c_g1 = s.Open(8).Take((size_t)(num_vars - num_public - 1) * kG1);
    h_g1 = s.Open(9).Take((size_t)domain_size * kG1);
  }
};

// wtns v2 (wtns.h:66-154): canonical (non-Montgomery) scalars; returned in Montgomery form.
template <class Fr>
static std::vector<HostFp<Fr>> ReadWitness(const MappedFile& f) {
  BinSections s(f, "wtns", 2);
  auto h = s.Open(1);
  ExpectModulus<Fr>(h, "wtns");
  uint32_t n = h.U32();
  const uint8_t* raw = s.Open(2).Take((size_t)n * sizeof(HostFp<Fr>));
  std::vector<HostFp<Fr>> out(n);
  HostFp<Fr> r2;
  for (int i = 0; i < HostFp<Fr>::N; ++i) r2.v[i] = Fr::kR2_64[i];
  for (uint32_t i = 0; i < n; ++i) {
    HostFp<Fr> x;
    memcpy(&x, raw + (size_t)i * sizeof(x), sizeof(x));
    out[i] = x.Mul(r2);  // to Montgomery (wtns.h:104-106)
  }
  return out;
}

template <class Fr>
static HostFp<Fr> FrPow(HostFp<Fr> base, const uint64_t* e, int limbs) {
  HostFp<Fr> acc = HostFp<Fr>::One();
  for (int i = limbs * 64; i-- > 0;) {
    acc = acc.Sqr();
    if ((e[i >> 6] >> (i & 63)) & 1) acc = acc.Mul(base);
  }
  return acc;
}

// Primitive n-th root of unity of Fr, n a power of two: g^((r - 1) / n) for the multiplicative
// generator g of the field (5 for BN254, 7 for BLS12-381).
template <class Fr>
static HostFp<Fr> RootOfUnity(uint64_t n, uint32_t generator) {
  using E = HostFp<Fr>;
  uint64_t e[E::N];
  memcpy(e, Fr::kMod64, sizeof(e));
  e[0] -= 1;  // r - 1 (r is odd)
  unsigned shift = 0;
  while ((uint64_t(1) << shift) < n) ++shift;
  for (int i = 0; i < E::N; ++i) {  // e >>= shift (exact: the two-adicity was checked by the caller)
    uint64_t hi = i + 1 < E::N ? e[i + 1] : 0;
    e[i] = shift ? (e[i] >> shift) | (hi << (64 - shift)) : e[i];
  }
  E g = E::One();
  E acc = E::Zero();
  for (uint32_t k = 0; k < generator; ++k) acc = acc.Add(g);  // small integer -> Montgomery
  E w = FrPow<Fr>(acc, e, E::N);
  // w^(n/2) must be -1
  E t = w;
  for (uint64_t m = n; m > 2; m >>= 1) t = t.Sqr();
  if (n >= 2 && !t.Add(E::One()).IsZero()) throw FileError{"no root of unity of that order"};
  return w;
}

// In-place radix-2 NTT over Fr (decimation in time, bit-reversed input order handled here).
template <class Fr>
static void Ntt(std::vector<HostFp<Fr>>& a, const HostFp<Fr>& root) {
  using E = HostFp<Fr>;
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    E wl = root;
    for (size_t m = n; m > len; m >>= 1) wl = wl.Sqr();
    for (size_t i = 0; i < n; i += len) {
      E w = E::One();
      for (size_t k = 0; k < len / 2; ++k) {
        E u = a[i + k], v = a[i + k + len / 2].Mul(w);
        a[i + k] = u.Add(v);
        a[i + k + len / 2] = u.Sub(v);
        w = w.Mul(wl);
      }
    }
  }
}

// QAP witness map (quadratic_arithmetic_program.h:25-118, after rapidsnark): a = A z, b = B z
// from the coefficient list, c = a o b; all three to coefficient form (inverse NTT), onto the
// coset of the 2n-th root of unity, back to evaluations; h_i = a_i b_i - c_i.
template <class Fq, class Fr>
static std::vector<HostFp<Fr>> WitnessMap(const Zkey<Fq, Fr>& zk, const std::vector<HostFp<Fr>>& full,
                                          uint32_t generator) {
  using E = HostFp<Fr>;
  const size_t n = zk.domain_size;
  std::vector<E> a(n, E::Zero()), b(n, E::Zero()), c(n);
  E one_canonical = E::Zero();
  one_canonical.v[0] = 1;
  for (uint32_t i = 0; i < zk.num_coefficients; ++i) {
    ZkeyCoefficient<Fr> co;
    memcpy(&co, &zk.coefficients[i], sizeof(co));  // packed, unaligned in the mapping
    if (co.constraint >= n || co.signal >= full.size()) throw FileError{"zkey: coefficient out of range"};
    E v;
    memcpy(&v, co.value, sizeof(v));
    v = v.Mul(one_canonical);  // stored as value * R^2: one reduction leaves Montgomery form
    std::vector<E>& ab = co.matrix == 0 ? a : b;
    ab[co.constraint] = ab[co.constraint].Add(v.Mul(full[co.signal]));
  }
  for (size_t i = 0; i < n; ++i) c[i] = a[i].Mul(b[i]);
  const E w = RootOfUnity<Fr>(n, generator), w_inv = w.Inv(), g = RootOfUnity<Fr>(2 * n, generator);
  E n_mont = E::Zero();
  {
    E one = E::One(), acc = E::Zero(), p = one;  // n as a field element: double-and-add
    for (size_t bit = 1; bit <= n; bit <<= 1, p = p.Dbl())
      if (n & bit) acc = acc.Add(p);
    n_mont = acc;
  }
  const E n_inv = n_mont.Inv();
  auto coset = [&](std::vector<E>& x) {
    Ntt<Fr>(x, w_inv);  // inverse transform (scaled below together with the coset powers)
    E p = n_inv;
    for (size_t i = 0; i < n; ++i) {
      x[i] = x[i].Mul(p);
      p = p.Mul(g);
    }
    Ntt<Fr>(x, w);
  };
  std::thread ta([&] { coset(a); }), tb([&] { coset(b); });
  coset(c);
  ta.join();
  tb.join();
  for (size_t i = 0; i < n; ++i) a[i] = a[i].Mul(b[i]).Sub(c[i]);
  return a;
}

// Decimal string of a canonical little-endian big integer (what PrimeField::ToString prints).
static inline std::string DecimalString(const uint64_t* limbs, int n) {
  std::vector<uint64_t> v(limbs, limbs + n);
  std::string out;
  for (;;) {
    bool zero = true;
    unsigned __int128 rem = 0;
    for (int i = n; i-- > 0;) {
      unsigned __int128 cur = (rem << 64) | v[i];
      v[i] = (uint64_t)(cur / 1000000000000000000ull);
      rem = cur % 1000000000000000000ull;
      zero = zero && v[i] == 0;
    }
    char buf[32];
    if (zero) {
      snprintf(buf, sizeof(buf), "%llu", (unsigned long long)(uint64_t)rem);
      out.insert(0, buf);
      break;
    }
    snprintf(buf, sizeof(buf), "%018llu", (unsigned long long)(uint64_t)rem);
    out.insert(0, buf);
  }
  return out;
}

template <class F>
static std::string FieldDecimal(const HostFp<F>& montgomery) {
  HostFp<F> one = HostFp<F>::Zero();
  one.v[0] = 1;
  HostFp<F> c = montgomery.Mul(one);
  return DecimalString(c.v, HostFp<F>::N);
}

static inline void WriteTextFile(const char* path, const std::string& text) {
  FILE* f = fopen(path, "wb");
  if (!f) throw FileError{std::string("cannot write ") + path};
  size_t n = fwrite(text.data(), 1, text.size(), f);
  fclose(f);
  if (n != text.size()) throw FileError{std::string("short write to ") + path};
}

// snarkjs proof.json (groth16_proof.h:21-44, points.h:17-55): affine coordinates as decimal
// strings, G1 as [x, y, "1"], G2 as [[x.c0, x.c1], [y.c0, y.c1], ["1", "0"]].
template <class Fq>
static std::string ProofJson(const HostPointAffine<HostFp<Fq>>& a, const HostPointAffine<HostFp2<Fq>>& b,
                             const HostPointAffine<HostFp<Fq>>& c, const char* curve_name) {
  auto g1 = [](const HostPointAffine<HostFp<Fq>>& p) {
    return "[\"" + FieldDecimal<Fq>(p.x) + "\",\"" + FieldDecimal<Fq>(p.y) + "\",\"1\"]";
  };
  std::string s = "{\"pi_a\":" + g1(a) + ",\"pi_b\":[[\"" + FieldDecimal<Fq>(b.x.c0) + "\",\"" +
                  FieldDecimal<Fq>(b.x.c1) + "\"],[\"" + FieldDecimal<Fq>(b.y.c0) + "\",\"" +
                  FieldDecimal<Fq>(b.y.c1) + "\"],[\"1\",\"0\"]],\"pi_c\":" + g1(c) +
                  ",\"protocol\":\"groth16\",\"curve\":\"" + curve_name + "\"}";
  return s;
}

// public.json (prime_field.h:19-34): the public inputs as decimal strings.
template <class Fr>
static std::string PublicJson(const HostFp<Fr>* inputs, size_t n) {
  std::string s = "[";
  for (size_t i = 0; i < n; ++i) s += (i ? ",\"" : "\"") + FieldDecimal<Fr>(inputs[i]) + "\"";
  return s + "]";
}

}  // namespace tb200
