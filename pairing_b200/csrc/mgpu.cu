// mgpu.cu -- several GPUs of one node behind ONE call of the C ABI (include/pairing_b200.h, "several GPUs").
//
// The reference's Engine::miller_loop takes every pair of a product in one call (bls12_381/mod.rs:40-102) and its caller
// is one process, so the multi-device path lives inside the library: one host thread per device for the duration of a
// call, contiguous shards, and -- for the product -- ONE exchange step: every device folds its shard to a single 576-byte
// Fq12 (k_pair_multi_miller + k_pair_product_tail), copies it into a gather buffer on the first device (peer copy over
// NVLink, ordered on the producing device's stream), and the first device folds the n partials and runs the single final
// exponentiation.  (Under torchrun -- one process per GPU -- the same two kernels are used around an NCCL all-gather:
// pairing_b200/dist.py.)  Host code only: no kernels in this translation unit.
#include <chrono>
#include <new>
#include <thread>
#include <vector>

#include "abi_common.cuh"

struct bls_mgpu {
  std::vector<bls_ctx*> ctx;
  void* gather = nullptr;        // on device 0: n partials + the result + the is_some byte
  double phase_ms[3] = {0, 0, 0};
};

namespace {
struct Shard { size_t lo, n; };
Shard shard_of(size_t n, int parts, int i) {   // contiguous [n i / parts, n (i + 1) / parts): pairing_b200/dist.py::shard_range
  const size_t lo = (size_t)((unsigned __int128)n * (unsigned)i / (unsigned)parts);
  const size_t hi = (size_t)((unsigned __int128)n * (unsigned)(i + 1) / (unsigned)parts);
  return Shard{lo, hi - lo};
}
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// run f(i) for every device on its own host thread (device 0 on the calling thread); first failure wins
template <class F> int for_each_device(bls_mgpu* m, F f) {
  const int nd = (int)m->ctx.size();
  std::vector<int> rc(nd, BLS_OK);
  std::vector<std::thread> th;
  for (int i = 1; i < nd; i++) th.emplace_back([&, i] { rc[i] = f(i); });
  rc[0] = f(0);
  for (auto& t : th) t.join();
  for (int i = 0; i < nd; i++)
    if (rc[i] != BLS_OK) return rc[i];
  return BLS_OK;
}

// one device's share of a product: H2D, Miller kernel, per-device fold, then the partial into slot i of the gather buffer
int product_shard(bls_mgpu* m, int i, const bls_g1_affine* p, const bls_g2_affine* q, size_t n) {
  bls_ctx* ctx = m->ctx[i];
  const Shard sh = shard_of(n, (int)m->ctx.size(), i);
  USE_DEVICE(ctx);
  H2D(dp, p + sh.lo, sh.n * sizeof(*p));
  H2D(dq, q + sh.lo, sh.n * sizeof(*q));
  DALLOC(dscr, bls_multi_miller_scratch_bytes(ctx, sh.n));
  DALLOC(dpart, sizeof(bls_fq12));
  TRY(bls_multi_miller_loop_dev(ctx, (const bls_g1_affine*)dp.p, (const bls_g2_affine*)dq.p, sh.n, (bls_fq12*)dpart.p, dscr.p, nullptr));
  bls_fq12* slot = (bls_fq12*)m->gather + i;
  if (i == 0) CK(cudaMemcpyAsync(slot, dpart.p, sizeof(bls_fq12), cudaMemcpyDeviceToDevice, ctx->stream));
  else CK(cudaMemcpyPeerAsync(slot, m->ctx[0]->device, dpart.p, ctx->device, sizeof(bls_fq12), ctx->stream));
  SYNC();
  return BLS_OK;
}

int product_call(bls_mgpu* m, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1, int final_exp, uint8_t* is_some) {
  if (!m || !out1 || (n && (!p || !q))) return BLS_ERR_INVALID_ARGUMENT;
  const int nd = (int)m->ctx.size();
  const double t0 = now_ms();
  int rc = for_each_device(m, [&](int i) { return product_shard(m, i, p, q, n); });
  if (rc != BLS_OK) return rc;
  const double t1 = now_ms();
  bls_ctx* ctx = m->ctx[0];
  USE_DEVICE(ctx);
  bls_fq12* res = (bls_fq12*)m->gather + nd;
  uint8_t* dsome = (uint8_t*)(res + 1);
  TRY(bls_fq12_product_tail_dev(ctx, (const bls_fq12*)m->gather, (size_t)nd, res, final_exp, dsome, nullptr));
  CK(cudaMemcpyAsync(out1, res, sizeof(*out1), cudaMemcpyDeviceToHost, ctx->stream));
  if (final_exp && is_some) CK(cudaMemcpyAsync(is_some, dsome, 1, cudaMemcpyDeviceToHost, ctx->stream));
  SYNC();
  const double t2 = now_ms();
  m->phase_ms[0] = t1 - t0; m->phase_ms[1] = t2 - t1; m->phase_ms[2] = t2 - t0;
  return BLS_OK;
}
}  // namespace

extern "C" {

bls_mgpu* bls_mgpu_create(const int* devices, int n_devices, int* err) {
  if (n_devices <= 0 || n_devices > 64) { if (err) *err = BLS_ERR_INVALID_ARGUMENT; return nullptr; }
  bls_mgpu* m = new (std::nothrow) bls_mgpu();
  if (!m) { if (err) *err = BLS_ERR_OUT_OF_MEMORY; return nullptr; }
  int e = BLS_OK;
  for (int i = 0; i < n_devices && e == BLS_OK; i++) {
    bls_ctx* c = bls_ctx_create(devices ? devices[i] : i, &e);
    if (c) m->ctx.push_back(c);
  }
  if (e == BLS_OK) {
    DevGuard guard;
    const int d0 = m->ctx[0]->device;
    if (guard.enter(d0) != cudaSuccess || cudaMalloc(&m->gather, (size_t)(n_devices + 1) * sizeof(bls_fq12) + 8) != cudaSuccess) e = BLS_ERR_CUDA;
    // direct peer stores into device 0's gather buffer where the topology allows it (NVLink / NVSwitch); a refusal is
    // not an error: cudaMemcpyPeerAsync then stages through the host
    for (int i = 1; i < n_devices && e == BLS_OK; i++) {
      DevGuard g2;
      int can = 0;
      if (g2.enter(m->ctx[i]->device) == cudaSuccess && cudaDeviceCanAccessPeer(&can, m->ctx[i]->device, d0) == cudaSuccess && can) {
        cudaError_t pe = cudaDeviceEnablePeerAccess(d0, 0);
        if (pe != cudaSuccess) cudaGetLastError();      // already enabled by the host application
      }
    }
  }
  if (e != BLS_OK) { if (err) *err = e; bls_mgpu_destroy(m); return nullptr; }
  if (err) *err = BLS_OK;
  return m;
}

void bls_mgpu_destroy(bls_mgpu* m) {
  if (!m) return;
  if (m->gather && !m->ctx.empty()) {
    DevGuard guard;
    guard.enter(m->ctx[0]->device);
    cudaFree(m->gather);
  }
  for (bls_ctx* c : m->ctx) bls_ctx_destroy(c);
  delete m;
}

int bls_mgpu_device_count(const bls_mgpu* m) { return m ? (int)m->ctx.size() : 0; }
bls_ctx* bls_mgpu_ctx(bls_mgpu* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

int bls_mgpu_multi_miller_loop(bls_mgpu* m, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1) {
  return product_call(m, p, q, n, out1, 0, nullptr);
}
int bls_mgpu_pairing_product(bls_mgpu* m, const bls_g1_affine* p, const bls_g2_affine* q, size_t n, bls_fq12* out1, uint8_t* is_some) {
  return product_call(m, p, q, n, out1, 1, is_some);
}

int bls_mgpu_pairing_batch(bls_mgpu* m, const bls_g1_affine* p, const bls_g2_affine* q, bls_fq12* out, size_t n) {
  if (!m || (n && (!p || !q || !out))) return BLS_ERR_INVALID_ARGUMENT;
  const int nd = (int)m->ctx.size();
  return for_each_device(m, [&](int i) {
    const Shard sh = shard_of(n, nd, i);
    return bls_pairing_batch(m->ctx[i], p + sh.lo, q + sh.lo, out + sh.lo, sh.n);
  });
}
int bls_mgpu_g1_wnaf_mul_batch(bls_mgpu* m, const bls_g1* bases, const bls_fr_repr* k, bls_g1* out, size_t n) {
  if (!m || (n && (!bases || !k || !out))) return BLS_ERR_INVALID_ARGUMENT;
  const int nd = (int)m->ctx.size();
  return for_each_device(m, [&](int i) {
    const Shard sh = shard_of(n, nd, i);
    return bls_g1_wnaf_mul_batch(m->ctx[i], bases + sh.lo, k + sh.lo, out + sh.lo, sh.n);
  });
}
int bls_mgpu_g2_wnaf_mul_batch(bls_mgpu* m, const bls_g2* bases, const bls_fr_repr* k, bls_g2* out, size_t n) {
  if (!m || (n && (!bases || !k || !out))) return BLS_ERR_INVALID_ARGUMENT;
  const int nd = (int)m->ctx.size();
  return for_each_device(m, [&](int i) {
    const Shard sh = shard_of(n, nd, i);
    return bls_g2_wnaf_mul_batch(m->ctx[i], bases + sh.lo, k + sh.lo, out + sh.lo, sh.n);
  });
}

int bls_mgpu_last_phase_ms(const bls_mgpu* m, double* ms3) {
  if (!m || !ms3) return BLS_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < 3; k++) ms3[k] = m->phase_ms[k];
  return BLS_OK;
}

}  // extern "C"
