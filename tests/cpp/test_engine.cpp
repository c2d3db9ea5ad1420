// test_engine.cpp -- the reference's own engine and curve tests (src/tests/engine.rs, src/tests/curve.rs), restated in
// batch form on the C++ host mirror (include/pairing_b200.hpp), plus bit-exact comparison with the CPU oracle.
// Built and run by tests/test_gpu_cpp.py on the GPU box:  g++ -std=c++17 ... -lpairing_b200 -lbls_oracle
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "pairing_b200.hpp"

using namespace pairing_b200;

// the oracle's C API (oracle/bls_oracle.c); its structs have the ABI's layout.  TEST CODE ONLY.
extern "C" {
int oracle_pairing(const bls_g1_affine*, const bls_g2_affine*, bls_fq12*, size_t, int);
int oracle_g1_op(int op, const bls_g1*, const void* b, const bls_fr_repr* k, bls_g1* out, size_t n, int window, int threads);
int oracle_g2_op(int op, const bls_g2*, const void* b, const bls_fr_repr* k, bls_g2* out, size_t n, int window, int threads);
int oracle_g1_into_affine(const bls_g1*, bls_g1_affine*, size_t);
int oracle_g2_into_affine(const bls_g2*, bls_g2_affine*, size_t);
int oracle_g1_from_affine(const bls_g1_affine*, bls_g1*, size_t);
int oracle_g2_from_affine(const bls_g2_affine*, bls_g2*, size_t);
int oracle_g1_batch_normalization(bls_g1*, size_t);
int oracle_g2_batch_normalization(bls_g2*, size_t);
void oracle_generators(bls_g1_affine*, bls_g2_affine*);
}
enum { O_DOUBLE = 0, O_ADD = 1, O_ADD_MIXED = 2, O_NEGATE = 3, O_MUL = 4, O_WNAF = 5, O_SUB = 6 };   // oracle PT_* codes
static const int TH = 8;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)

static uint64_t sm_state = 0x5dbe62598d313d76ull;
static uint64_t splitmix() {
  uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static std::vector<FrRepr> rand_scalars(size_t n) {   // < 2^254 < r, never 0 or 1
  std::vector<FrRepr> k(n);
  for (auto& s : k) {
    for (int i = 0; i < 4; i++) s.l[i] = splitmix();
    s.l[3] &= (1ull << 62) - 1;
    s.l[0] |= 2;
  }
  return k;
}
template <class T> static bool same(const std::vector<T>& a, const std::vector<T>& b) {
  return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0);
}
// G::rand stand-in: [k] * generator as a non-normalised Jacobian point (oracle double-and-add)
static std::vector<G1Point> rand_g1(size_t n) {
  bls_g1_affine g1; bls_g2_affine g2; oracle_generators(&g1, &g2);
  std::vector<G1AffinePoint> ga(n, g1);
  std::vector<G1Point> base(n), out(n);
  oracle_g1_from_affine(ga.data(), base.data(), n);
  auto k = rand_scalars(n);
  oracle_g1_op(O_MUL, base.data(), nullptr, k.data(), out.data(), n, 0, TH);
  return out;
}
static std::vector<G2Point> rand_g2(size_t n) {
  bls_g1_affine g1; bls_g2_affine g2; oracle_generators(&g1, &g2);
  std::vector<G2AffinePoint> ga(n, g2);
  std::vector<G2Point> base(n), out(n);
  oracle_g2_from_affine(ga.data(), base.data(), n);
  auto k = rand_scalars(n);
  oracle_g2_op(O_MUL, base.data(), nullptr, k.data(), out.data(), n, 0, TH);
  return out;
}

// src/tests/engine.rs:5-47
// device count without CUDA headers in this host-only program: the runtime is already loaded through libpairing_b200.so
#include <dlfcn.h>
static void cudaGetDeviceCount_shim(int* n) {
  *n = 0;
  typedef int (*fn_t)(int*);
  fn_t fn = (fn_t)dlsym(RTLD_DEFAULT, "cudaGetDeviceCount");
  if (!fn) {
    void* h = dlopen("libcudart.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libcudart.so.12", RTLD_NOW | RTLD_GLOBAL);
    if (h) fn = (fn_t)dlsym(h, "cudaGetDeviceCount");
  }
  if (fn) fn(n);
}

static void engine_tests(Gpu& g) {
  const size_t n = 10;
  auto a = G1::into_affine(g, rand_g1(n));
  auto b = G2::into_affine(g, rand_g2(n));
  CHECK(same(G1Affine::pairing_with(g, a, b), G2Affine::pairing_with(g, b, a)));
  CHECK(same(G1Affine::pairing_with(g, a, b), Bls12::pairing(g, a, b)));
  // zero-point handling: a pair with an infinity member contributes the factor one
  const size_t m = 24;
  G1AffinePoint z1; std::memset(&z1, 0, sizeof z1); z1.y = detail::fq_one(); z1.infinity = 1;     // G1Affine::zero(), ec.rs:158-164
  G2AffinePoint z2; std::memset(&z2, 0, sizeof z2); z2.y.c0 = detail::fq_one(); z2.infinity = 1;
  auto pa = G1::into_affine(g, rand_g1(m)), pc = G1::into_affine(g, rand_g1(m));
  auto qb = G2::into_affine(g, rand_g2(m)), qd = G2::into_affine(g, rand_g2(m));
  auto pb = G2Affine::prepare(g, qb), pd = G2Affine::prepare(g, qd);
  auto pz2 = G2Affine::prepare(g, {z2});
  const Fq12 one = fq12_one();
  for (size_t i = 0; i < m; i++) {
    auto fe = [&](const std::vector<G1PreparedPoint>& p, const std::vector<G2PreparedPoint>& q) {
      return Bls12::final_exponentiation(g, Bls12::miller_loop(g, p, q)).value();
    };
    CHECK(fe({z1}, {pb[i]}) == one);
    CHECK(fe({pa[i]}, {pz2[0]}) == one);
    CHECK(fe({z1, pc[i]}, {pb[i], pd[i]}) == fe({pa[i], pc[i]}, {pz2[0], pd[i]}));
    CHECK(fe({pa[i], z1}, {pb[i], pd[i]}) == fe({pa[i], pc[i]}, {pb[i], pz2[0]}));
  }
  // final_exponentiation(0) is None (mod.rs:105-106)
  Fq12 zero; std::memset(&zero, 0, sizeof zero);
  CHECK(!Bls12::final_exponentiation(g, zero).has_value());
}

// src/tests/engine.rs:50-91
static void random_miller_loop_tests(Gpu& g) {
  const size_t n = 200;
  auto a = rand_g1(n); auto b = rand_g2(n); auto c = rand_g1(n); auto d = rand_g2(n);
  auto aa = G1::into_affine(g, a), ca = G1::into_affine(g, c);
  auto ba = G2::into_affine(g, b), da = G2::into_affine(g, d);
  auto p2 = Bls12::pairing(g, aa, ba);
  auto fe = Bls12::final_exponentiation(g, Bls12::miller_loop_batch(g, aa, ba));
  for (size_t i = 0; i < n; i++) CHECK(fe[i].has_value() && *fe[i] == p2[i]);
  // the oracle agrees bit for bit
  std::vector<Fq12> want(n);
  oracle_pairing(aa.data(), ba.data(), want.data(), n, TH);
  CHECK(same(p2, want));
  // double Miller loop, with prepared G2 inputs as in the reference
  auto cd = Bls12::pairing(g, ca, da);
  auto pb = G2Affine::prepare(g, ba), pd = G2Affine::prepare(g, da);
  for (size_t i = 0; i < 24; i++) {
    Fq12 abcd = Bls12::product(g, {p2[i], cd[i]});
    auto dbl = Bls12::final_exponentiation(g, Bls12::miller_loop(g, {aa[i], ca[i]}, {pb[i], pd[i]}));
    CHECK(dbl.has_value() && *dbl == abcd);
  }
  // many G1 points against one prepared G2 point (the use G2Prepared exists for)
  {
    std::vector<G2AffinePoint> rep(n, ba[3]);
    CHECK(same(Bls12::pairing_shared_q(g, aa, pb[3]), Bls12::pairing(g, aa, rep)));
  }
  // one multi-Miller loop over all 2n pairs == product of everything
  std::vector<G1AffinePoint> allp(aa); allp.insert(allp.end(), ca.begin(), ca.end());
  std::vector<G2AffinePoint> allq(ba); allq.insert(allq.end(), da.begin(), da.end());
  std::vector<Fq12> all(p2); all.insert(all.end(), cd.begin(), cd.end());
  CHECK(*Bls12::final_exponentiation(g, Bls12::miller_loop(g, allp, allq)) == Bls12::product(g, all));
  // final_exponentiation(miller_loop(pairs)) in one call, and the same product with the pairs sharded over every visible
  // device inside the library (bls_mgpu_*): one GT element, bit-equal
  CHECK(*Bls12::pairing_product(g, allp, allq) == Bls12::product(g, all));
  int ndev = 0;
  cudaGetDeviceCount_shim(&ndev);
  std::vector<int> devs;
  for (int i = 0; i < (ndev > 0 ? ndev : 1); i++) devs.push_back(i);
  MultiGpu mg(devs);
  CHECK(mg.device_count() == (int)devs.size());
  CHECK(Bls12::miller_loop(mg, allp, allq) == Bls12::miller_loop(g, allp, allq));
  CHECK(*Bls12::pairing_product(mg, allp, allq) == Bls12::product(g, all));
  CHECK(same(Bls12::pairing(mg, allp, allq), all));
}

// src/tests/engine.rs:93-126
static void random_bilinearity_tests(Gpu& g) {
  const size_t n = 100;
  auto a = rand_g1(n); auto b = rand_g2(n);
  auto c = rand_scalars(n), d = rand_scalars(n);
  auto ac = G1::mul_assign(g, a, c), ad = G1::mul_assign(g, a, d);
  auto bc = G2::mul_assign(g, b, c), bd = G2::mul_assign(g, b, d);
  auto acbd = Bls12::pairing(g, G1::into_affine(g, ac), G2::into_affine(g, bd));
  auto adbc = Bls12::pairing(g, G1::into_affine(g, ad), G2::into_affine(g, bc));
  // let mut cd = c; cd.mul_assign(&d); e(a, b).pow(cd.into_repr())
  auto cd = FrField::into_repr(g, FrField::mul_assign(g, FrField::from_repr(g, c), FrField::from_repr(g, d)));
  auto eab = Bls12::pairing(g, G1::into_affine(g, a), G2::into_affine(g, b));
  auto abcd = Bls12::pow(g, eab, cd);
  CHECK(same(acbd, adbc));
  CHECK(same(acbd, abcd));
  CHECK(same(abcd, Bls12::pow(g, Bls12::pow(g, eab, c), d)));
  // a * a^-1 == 1 and 0 has no inverse (fr.rs:1342-1357)
  auto cm = FrField::from_repr(g, c);
  std::vector<uint8_t> ok;
  auto ci = FrField::inverse(g, cm, &ok);
  auto one = FrField::from_repr(g, std::vector<FrRepr>(n, FrRepr{{1, 0, 0, 0}}));
  CHECK(same(FrField::mul_assign(g, cm, ci), one));
  FrField::inverse(g, std::vector<bls_fr>(1, bls_fr{{0, 0, 0, 0}}), &ok);
  CHECK(ok[0] == 0);
  CHECK(!(acbd[0] == fq12_one()));
}

// src/tests/curve.rs:68-92 (wNAF == mul_assign), 347-388 (batch_normalization == into_affine with sprinkled entries)
template <class G, class RandFn, class OracleOp, class OracleNorm>
static void curve_tests(Gpu& g, RandFn rand_pts, OracleOp oracle_op, OracleNorm oracle_norm) {
  using Proj = typename G::Projective;
  const size_t n = 96;
  auto base = rand_pts(n);
  auto k = rand_scalars(n);
  k[0] = FrRepr{{0, 0, 0, 0}}; k[1] = FrRepr{{1, 0, 0, 0}}; k[2] = FrRepr{{1ull << 33, 0, 0, 0}};   // windows 2/3 and the zero scalar
  auto w = Wnaf<G>().scalar(k).base(g, base);
  auto mres = G::mul_assign(g, base, k);
  CHECK(same(G::into_affine(g, w), G::into_affine(g, mres)));
  std::vector<Proj> want(n);
  oracle_op(O_WNAF, base.data(), nullptr, k.data(), want.data(), n, 0, TH);
  CHECK(same(w, want));                                   // the Jacobian triples themselves are the reference's
  // fixed-base mode
  for (size_t num : {1u, 50u, 700u}) {
    auto wb = Wnaf<G>().base(base[5], num);
    std::vector<Proj> rep(n, base[5]);
    oracle_op(O_WNAF, rep.data(), nullptr, k.data(), want.data(), n, wb.window_size(), TH);
    CHECK(same(wb.scalar(g, k), want));
  }
  // batch normalisation with infinity and already-normalised entries sprinkled in
  auto v = rand_pts(n);
  std::memset(&v[3], 0, sizeof(Proj)); std::memset(&v[40], 0, sizeof(Proj));
  {
    std::vector<Proj> some(v.begin() + 10, v.begin() + 20);
    auto aff = G::into_affine(g, some);
    auto back = G::into_projective(aff);
    std::copy(back.begin(), back.end(), v.begin() + 10);
  }
  auto expected = G::into_affine(g, v);
  auto ref = v;
  oracle_norm(ref.data(), n);
  G::batch_normalization(g, v);
  CHECK(same(v, ref));
  CHECK(same(G::into_affine(g, v), expected));
  // group law against the oracle: add / double / sub / mixed
  auto x = rand_pts(n), y = rand_pts(n);
  y[7] = x[7];
  oracle_op(O_ADD, x.data(), y.data(), nullptr, want.data(), n, 0, TH);
  CHECK(same(G::add_assign(g, x, y), want));
  oracle_op(O_DOUBLE, x.data(), nullptr, nullptr, want.data(), n, 0, TH);
  CHECK(same(G::double_(g, x), want));
  oracle_op(O_SUB, x.data(), y.data(), nullptr, want.data(), n, 0, TH);
  CHECK(same(G::sub_assign(g, x, y), want));
  auto ya = G::into_affine(g, y);
  oracle_op(O_ADD_MIXED, x.data(), ya.data(), nullptr, want.data(), n, 0, TH);
  CHECK(same(G::add_assign_mixed(g, x, ya), want));
}

template <class Enc, class G, class RandFn> static void encoding_tests(Gpu& g, RandFn rand_pts) {
  auto aff = G::into_affine(g, rand_pts(50));
  aff[4].infinity = 1; std::memset(&aff[4].x, 0, sizeof(aff[4].x)); std::memset(&aff[4].y, 0, sizeof(aff[4].y));
  reinterpret_cast<Fq*>(&aff[4].y)[0] = detail::fq_one();
  auto bytes = Enc::from_affine(g, aff);
  CHECK(bytes.size() == aff.size() * Enc::size());
  auto dec = Enc::into_affine(g, bytes);
  CHECK(same(dec.points, aff));
  for (auto s : dec.status) CHECK(s == BLS_DEC_OK);
  bytes[0] ^= 0x80;                                     // wrong compression flag on the first element
  auto bad = Enc::into_affine(g, bytes);
  CHECK(bad.status[0] == BLS_DEC_UNEXPECTED_COMPRESSION_MODE && bad.status[1] == BLS_DEC_OK);
}

int main(int argc, char** argv) {
  try {
    Gpu g(0);
    // src/bls12_381/tests/mod.rs:5-53: e(g1, g2) against the RELIC value (golden file written from the reference's decimal constants)
    if (argc > 1) {
      bls_g1_affine g1; bls_g2_affine g2; oracle_generators(&g1, &g2);
      std::ifstream f(argv[1], std::ios::binary);
      Fq12 want;
      f.read(reinterpret_cast<char*>(&want), sizeof want);
      CHECK(f.gcount() == (std::streamsize)sizeof want);
      CHECK(Bls12::pairing(g, {g1}, {g2})[0] == want);
    }
    engine_tests(g);
    random_miller_loop_tests(g);
    random_bilinearity_tests(g);
    curve_tests<G1>(g, rand_g1, oracle_g1_op, oracle_g1_batch_normalization);
    curve_tests<G2>(g, rand_g2, oracle_g2_op, oracle_g2_batch_normalization);
    encoding_tests<G1Compressed, G1>(g, rand_g1);
    encoding_tests<G1Uncompressed, G1>(g, rand_g1);
    encoding_tests<G2Compressed, G2>(g, rand_g2);
    encoding_tests<G2Uncompressed, G2>(g, rand_g2);
    CHECK(G1::recommended_wnaf_for_num_scalars(1000000) == 16 && G2::recommended_wnaf_for_num_scalars(1000000) == 15);
    CHECK(G1::recommended_wnaf_for_scalar(FrRepr{{0, 0, 4, 0}}) == 4 && G1::recommended_wnaf_for_scalar(FrRepr{{0, 0, 1, 0}}) == 3);
  } catch (const Error& e) {
    printf("FAIL: pairing_b200::Error %d: %s\n", e.status, e.what());
    return 2;
  }
  if (failures) { printf("%d check(s) failed\n", failures); return 1; }
  printf("cpp engine tests ok\n");
  return 0;
}
