// imad_forms.cu -- which operand form keeps the IMAD pipe full?  Standalone probe (not part of the library).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_forms imad_forms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pairing_b200/csrc/fp.cuh"
using namespace bls;

__constant__ uint32_t c_q[12] = {BLS_Q0, BLS_Q1, BLS_Q2, BLS_Q3, BLS_Q4, BLS_Q5, BLS_Q6, BLS_Q7, BLS_Q8, BLS_Q9, BLS_Q10, BLS_Q11};

// reduction row with q supplied as registers / constant-bank values
__device__ __forceinline__ void redc_row_q(uint32_t (&e)[12], uint32_t (&o)[12], const uint32_t (&q)[12]) {
  uint32_t m = e[0] * BLS_NINV;
  uint32_t dummy = 0;
  fp_cmad_row(o, &q[1], m, dummy);
  fp_cmad_row(e, &q[0], m, o[11]);
}
template <int MODE>
__device__ __forceinline__ Fp mul_t(const Fp& a, const Fp& b, const uint32_t (&q)[12]) {
  uint32_t e[12], o[12];
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    uint64_t t = (uint64_t)a.v[j] * b.v[0];
    e[j] = (uint32_t)t; e[j + 1] = (uint32_t)(t >> 32);
    uint64_t u = (uint64_t)a.v[j + 1] * b.v[0];
    o[j] = (uint32_t)u; o[j + 1] = (uint32_t)(u >> 32);
  }
  if (MODE == 0) fp_redc_row(e, o); else redc_row_q(e, o, q);
#pragma unroll
  for (int i = 1; i < 12; i += 2) {
    fp_madc_rshift_row(o[0], e[1], e, &a.v[1], b.v[i]);
    fp_cmad_row(o, &a.v[0], b.v[i], e[11]);
    if (MODE == 0) fp_redc_row(o, e); else redc_row_q(o, e, q);
    if (i + 1 < 12) {
      fp_madc_rshift_row(e[0], o[1], o, &a.v[1], b.v[i + 1]);
      fp_cmad_row(e, &a.v[0], b.v[i + 1], o[11]);
      if (MODE == 0) fp_redc_row(e, o); else redc_row_q(e, o, q);
    }
  }
  return fp_merge(o, e);
}
// MODE 0: immediates, 1: constant bank, 2: registers (opaque mov), 3: like 2 but no final conditional subtract
template <int MODE>
__global__ void __launch_bounds__(256) k_mul(uint32_t seed, int iters, uint32_t* sink, uint32_t zero) {
  uint32_t q[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    if (MODE == 2) q[i] = c_q[i] ^ (zero * threadIdx.x);   // per-thread value: forces vector registers
    else q[i] = c_q[i];
  }
  Fp x = fp_one(), y = fp_r2();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed;
#pragma unroll 1
  for (int it = 0; it < iters; it++) { x = mul_t<MODE>(x, y, q); y = mul_t<MODE>(y, x, q); }
  if (x.v[0] == 0x1234567u && y.v[3] == 7u) sink[0] = x.v[1];
}
// rows only, multiplier form varies
template <int MODE>
__global__ void __launch_bounds__(256) k_rows(uint32_t seed, int iters, uint32_t* sink, uint32_t zero) {
  uint32_t q[12], e[12], o[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    if (MODE == 2) q[i] = c_q[i] ^ (zero * threadIdx.x);
    else q[i] = c_q[i];
    e[i] = seed + i; o[i] = seed * 3 + i;
  }
  uint32_t m = seed ^ (threadIdx.x * 2654435761u);
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MODE == 0) { fp_cmad_q_odd(o, m); fp_cmad_q_even(e, m, o[11]); }
      else { uint32_t d = 0; fp_cmad_row(o, &q[1], m, d); fp_cmad_row(e, &q[0], m, o[11]); }
      m += e[3];
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) s ^= e[k] ^ o[k];
  if (s == 0x1234567u) sink[0] = s;
}

// rolled: 6 iterations of two rows; b is rotated down two limbs per iteration (compact code: fits the L0 I-cache)
__device__ __forceinline__ Fp mul_rolled(const Fp& a, const Fp& bb) {
  uint32_t e[12], o[12], b[12];
#pragma unroll
  for (int j = 0; j < 12; j++) { e[j] = 0; o[j] = 0; b[j] = bb.v[j]; }
#pragma unroll 1
  for (int it = 0; it < 6; it++) {
    // even row: e is word-0 aligned
    fp_madc_rshift_row(e[0], o[1], o, &a.v[1], b[0]);   // o = (o >> 64) + a_odd * b0, carry from e[0] += o[1]
    fp_cmad_row(e, &a.v[0], b[0], o[11]);
    fp_redc_row(e, o);
    // odd row: roles swapped
    fp_madc_rshift_row(o[0], e[1], e, &a.v[1], b[1]);
    fp_cmad_row(o, &a.v[0], b[1], e[11]);
    fp_redc_row(o, e);
#pragma unroll
    for (int j = 0; j < 10; j++) b[j] = b[j + 2];
  }
  return fp_merge(o, e);
}
__global__ void __launch_bounds__(256) k_mul_rolled(uint32_t seed, int iters, uint32_t* sink, uint32_t zero) {
  Fp x = fp_one(), y = fp_r2();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed;
#pragma unroll 1
  for (int it = 0; it < iters; it++) { x = mul_rolled(x, y); y = mul_rolled(y, x); }
  if (x.v[0] == 0x1234567u && y.v[3] == 7u) sink[0] = x.v[1];
}
// one unrolled product per loop iteration (half the loop body of k_mul<0>)
__global__ void __launch_bounds__(256) k_mul_one(uint32_t seed, int iters, uint32_t* sink, uint32_t zero) {
  uint32_t q[12] = {0};
  Fp x = fp_one(), y = fp_r2();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed;
#pragma unroll 1
  for (int it = 0; it < 2 * iters; it++) { x = mul_t<0>(x, y, q); }
  if (x.v[0] == 0x1234567u && y.v[3] == 7u) sink[0] = x.v[1];
}
// cycles per product seen by ONE warp per SM sub-partition (latency view) and by a full SM
__global__ void __launch_bounds__(256) k_mul_cycles(uint32_t seed, int iters, long long* out) {
  uint32_t q[12] = {0};
  Fp x = fp_one(), y = fp_r2();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) { x = mul_t<0>(x, y, q); y = mul_t<0>(y, x, q); }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (x.v[0] == 0x1234567u && y.v[3] == 7u) out[1] = x.v[1];
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// mode 0: fp_mul chain, mode 1: 8 independent mad.wide chains.  out[0]=cycles, out[1]=ns (block 0, thread 0)
template <int MODE>
__global__ void __launch_bounds__(256) k_clock_probe(uint32_t seed, int iters, unsigned long long* out) {
  uint32_t q[12] = {0};
  Fp x = fp_one(), y = fp_r2();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed;
  uint64_t acc[8]; uint32_t a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { acc[k] = seed + k; a[k] = (seed ^ threadIdx.x) * (k + 3) + 1; }
  unsigned long long g0 = gtime(); long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) { x = mul_t<0>(x, y, q); y = mul_t<0>(y, x, q); }
    else {
#pragma unroll
      for (int u = 0; u < 75; u++)
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[k]), "r"(seed));
    }
  }
  long long t1 = clock64(); unsigned long long g1 = gtime();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = g1 - g0; }
  uint64_t sacc = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) sacc ^= acc[k];
  if (x.v[0] == 0x1234567u && y.v[3] == 7u && sacc == 99) out[2] = x.v[1];
}

// ablations of the product: FLAGS bit0 = no final conditional subtract, bit1 = no merge at all,
// bit2 = m does not depend on the accumulator, bit3 = no m*q rows at all (a*b rows only)
template <int FLAGS>
__device__ __forceinline__ Fp mul_abl(const Fp& a, const Fp& b, uint32_t mconst) {
  uint32_t e[12], o[12];
#pragma unroll
  for (int j = 0; j < 12; j++) { e[j] = a.v[j]; o[j] = b.v[j]; }
#pragma unroll
  for (int i = 0; i < 12; i += 2) {
    fp_madc_rshift_row(e[0], o[1], o, &a.v[1], b.v[i]);
    fp_cmad_row(e, &a.v[0], b.v[i], o[11]);
    if (!(FLAGS & 8)) {
      uint32_t m = (FLAGS & 4) ? mconst : e[0] * BLS_NINV;
      fp_cmad_q_odd(o, m); fp_cmad_q_even(e, m, o[11]);
    }
    fp_madc_rshift_row(o[0], e[1], e, &a.v[1], b.v[i + 1]);
    fp_cmad_row(o, &a.v[0], b.v[i + 1], e[11]);
    if (!(FLAGS & 8)) {
      uint32_t m = (FLAGS & 4) ? mconst + i : o[0] * BLS_NINV;
      fp_cmad_q_odd(e, m); fp_cmad_q_even(o, m, e[11]);
    }
  }
  Fp r;
  if (FLAGS & 2) {
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = e[j] ^ o[j];
    return r;
  }
  if (FLAGS & 1) {
    asm("add.cc.u32 %0, %12, %23;\n\taddc.cc.u32 %1, %13, %24;\n\taddc.cc.u32 %2, %14, %25;\n\taddc.cc.u32 %3, %15, %26;\n\t"
        "addc.cc.u32 %4, %16, %27;\n\taddc.cc.u32 %5, %17, %28;\n\taddc.cc.u32 %6, %18, %29;\n\taddc.cc.u32 %7, %19, %30;\n\t"
        "addc.cc.u32 %8, %20, %31;\n\taddc.cc.u32 %9, %21, %32;\n\taddc.cc.u32 %10, %22, %33;\n\taddc.u32 %11, %34, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(r.v[8]), "=r"(r.v[9]), "=r"(r.v[10]), "=r"(r.v[11])
        : "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]),
          "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]), "r"(e[10]), "r"(e[11]));
    return r;
  }
  return fp_merge(o, e);
}
template <int FLAGS>
__global__ void __launch_bounds__(256, 2) k_abl(uint32_t seed, int iters, uint32_t* sink, uint32_t zero) {
  Fp x = fp_one(), y = fp_r2();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed;
#pragma unroll 1
  for (int it = 0; it < iters; it++) { x = mul_abl<FLAGS>(x, y, seed + it); y = mul_abl<FLAGS>(y, x, seed ^ it); }
  if (x.v[0] == 0x1234567u && y.v[3] == 7u) sink[0] = x.v[1];
}

// row-form probes.  FORM 0: cmad rows, same b.  1: cmad rows, b varies per row.  2: rshift rows only.
// 3: rshift + cmad alternating (the a*b rows of the product).  4: like 3 but the rshift row starts a fresh carry chain.
__device__ __forceinline__ void rshift_row_nocarry(uint32_t (&acc)[12], const uint32_t* x, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %12, %18, %2;\n\tmadc.hi.cc.u32 %1, %12, %18, %3;\n\t"
      "madc.lo.cc.u32 %2, %13, %18, %4;\n\tmadc.hi.cc.u32 %3, %13, %18, %5;\n\t"
      "madc.lo.cc.u32 %4, %14, %18, %6;\n\tmadc.hi.cc.u32 %5, %14, %18, %7;\n\t"
      "madc.lo.cc.u32 %6, %15, %18, %8;\n\tmadc.hi.cc.u32 %7, %15, %18, %9;\n\t"
      "madc.lo.cc.u32 %8, %16, %18, %10;\n\tmadc.hi.cc.u32 %9, %16, %18, %11;\n\t"
      "madc.lo.cc.u32 %10, %17, %18, 0;\n\tmadc.hi.u32 %11, %17, %18, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
        "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11])
      : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(x[8]), "r"(x[10]), "r"(b));
}
template <int FORM>
__global__ void __launch_bounds__(256, 2) k_rowform(uint32_t seed, int iters, uint32_t* sink, uint32_t zero) {
  uint32_t e[12], o[12], x[12], b[12];
#pragma unroll
  for (int k = 0; k < 12; k++) { e[k] = seed + k; o[k] = seed * 3 + k; x[k] = (seed ^ threadIdx.x) * (2 * k + 1) + 1; b[k] = x[k] * 77 + k; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 12; i += 2) {
      if (FORM == 0) { fp_cmad_row(e, &x[0], b[0], o[11]); fp_cmad_row(o, &x[1], b[0], e[11]); fp_cmad_row(e, &x[0], b[0], o[11]); fp_cmad_row(o, &x[1], b[0], e[11]); }
      if (FORM == 1) { fp_cmad_row(e, &x[0], b[i], o[11]); fp_cmad_row(o, &x[1], b[i], e[11]); fp_cmad_row(e, &x[0], b[i + 1], o[11]); fp_cmad_row(o, &x[1], b[i + 1], e[11]); }
      if (FORM == 2) { fp_madc_rshift_row(e[0], o[1], o, &x[1], b[i]); fp_madc_rshift_row(o[0], e[1], e, &x[1], b[i]);
                       fp_madc_rshift_row(e[0], o[1], o, &x[1], b[i + 1]); fp_madc_rshift_row(o[0], e[1], e, &x[1], b[i + 1]); }
      if (FORM == 3) { fp_madc_rshift_row(e[0], o[1], o, &x[1], b[i]); fp_cmad_row(e, &x[0], b[i], o[11]);
                       fp_madc_rshift_row(o[0], e[1], e, &x[1], b[i + 1]); fp_cmad_row(o, &x[0], b[i + 1], e[11]); }
      if (FORM == 4) { rshift_row_nocarry(o, &x[1], b[i]); fp_cmad_row(e, &x[0], b[i], o[11]);
                       rshift_row_nocarry(e, &x[1], b[i + 1]); fp_cmad_row(o, &x[0], b[i + 1], e[11]); }
    }
  }
  uint32_t s2 = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) s2 ^= e[k] ^ o[k];
  if (s2 == 0x1234567u) sink[0] = s2;
}

// occupancy / ILP sweep: NCHAIN independent fp_mul chains per thread, MINB blocks per SM requested
template <int NCHAIN, int MINB>
__global__ void __launch_bounds__(256, MINB) k_ilp(uint32_t seed, int iters, uint32_t* sink, uint32_t zero) {
  uint32_t q[12] = {0};
  Fp x[NCHAIN], y[NCHAIN];
#pragma unroll
  for (int c = 0; c < NCHAIN; c++) { x[c] = fp_one(); y[c] = fp_r2(); x[c].v[0] ^= seed ^ threadIdx.x ^ c; y[c].v[1] ^= seed + c; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < NCHAIN; c++) x[c] = mul_t<0>(x[c], y[c], q);
#pragma unroll
    for (int c = 0; c < NCHAIN; c++) y[c] = mul_t<0>(y[c], x[c], q);
  }
  uint32_t s2 = 0;
#pragma unroll
  for (int c = 0; c < NCHAIN; c++) s2 ^= x[c].v[0] ^ y[c].v[3];
  if (s2 == 0x1234567u) sink[0] = s2;
}

template <class K> static void run(const char* name, K kern, double macs_per_thread_iter, int iters) {
  uint32_t* sink; cudaMalloc(&sink, 64);
  int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(a);
    kern<<<sm * 8, 256>>>(12345u, iters, sink, 0u);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (r && ms < best) best = ms;
  }
  double total = macs_per_thread_iter * iters * 256.0 * sm * 8;
  printf("%-44s %8.3f T MAC/s  (%.2f ms) %s\n", name, total / best * 1e-9, best, cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink);
}
int main() {
  run("rows, q immediate", k_rows<0>, 8 * 24.0, 4000);
  run("rows, q uniform registers", k_rows<1>, 8 * 24.0, 4000);
  run("rows, q vector registers", k_rows<2>, 8 * 24.0, 4000);
  run("fp_mul, q immediate", k_mul<0>, 600.0, 2000);
  run("fp_mul, q uniform registers", k_mul<1>, 600.0, 2000);
  run("fp_mul, q vector registers", k_mul<2>, 600.0, 2000);
  run("rowform 0: cmad rows, same b", k_rowform<0>, 288.0, 2000);
  run("rowform 1: cmad rows, b varies", k_rowform<1>, 288.0, 2000);
  run("rowform 2: rshift rows only", k_rowform<2>, 288.0, 2000);
  run("rowform 3: rshift + cmad (a*b rows)", k_rowform<3>, 288.0, 2000);
  run("rowform 4: rshift(no carry-in) + cmad", k_rowform<4>, 288.0, 2000);
  run("ablation: full (uniform rows)", k_abl<0>, 600.0, 2000);
  run("ablation: no final sub", k_abl<1>, 600.0, 2000);
  run("ablation: no merge", k_abl<2>, 600.0, 2000);
  run("ablation: m independent, no merge", k_abl<6>, 600.0, 2000);
  run("ablation: a*b rows only (300->144 MAC)", k_abl<10>, 288.0, 2000);
  run("fp_mul x1 chain, minBlocks=1", k_ilp<1, 1>, 600.0, 2000);
  run("fp_mul x1 chain, minBlocks=2", k_ilp<1, 2>, 600.0, 2000);
  run("fp_mul x1 chain, minBlocks=3", k_ilp<1, 3>, 600.0, 2000);
  run("fp_mul x1 chain, minBlocks=4", k_ilp<1, 4>, 600.0, 2000);
  run("fp_mul x2 chains, minBlocks=1", k_ilp<2, 1>, 1200.0, 1000);
  run("fp_mul x2 chains, minBlocks=2", k_ilp<2, 2>, 1200.0, 1000);
  run("fp_mul x3 chains, minBlocks=1", k_ilp<3, 1>, 1800.0, 700);
  run("fp_mul x3 chains, minBlocks=2", k_ilp<3, 2>, 1800.0, 700);
  run("fp_mul x4 chains, minBlocks=1", k_ilp<4, 1>, 2400.0, 500);
  run("fp_mul rolled (6 x 2 rows)", k_mul_rolled, 600.0, 2000);
  run("fp_mul unrolled, one per iteration", k_mul_one, 600.0, 2000);
  {
    unsigned long long* dd; cudaMalloc(&dd, 32); unsigned long long hh[2];
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    for (int mode = 0; mode < 2; mode++)
      for (int iters : {200, 2000, 20000}) {
        for (int bps : {1, 4, 8}) {
          cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
          cudaEventRecord(a);
          if (mode == 0) k_clock_probe<0><<<sm * bps, 256>>>(7u, iters, dd); else k_clock_probe<1><<<sm * bps, 256>>>(7u, iters, dd);
          cudaEventRecord(b); cudaEventSynchronize(b);
          float ms; cudaEventElapsedTime(&ms, a, b);
          cudaMemcpy(hh, dd, 16, cudaMemcpyDeviceToHost);
          double macs = 600.0 * iters * 256.0 * sm * bps;
          printf("%s iters=%6d blocks/SM=%d: event %.3f ms, block0 %.3f ms, %llu cycles -> %.0f MHz; %.2f T MAC/s; %.1f MAC/clk/SM\n",
                 mode == 0 ? "fp_mul  " : "mad.wide", iters, bps, ms, hh[1] * 1e-6, hh[0], hh[0] / (hh[1] * 1e-3), macs / ms * 1e-9,
                 macs / sm / (double)hh[0]);
        }
      }
  }
  long long* d; cudaMalloc(&d, 16); long long h[2];
  int cfgs[][2] = {{1, 32}, {1, 128}, {1, 256}, {1, 512}, {1, 1024}, {148 * 4, 256}};
  for (auto& c : cfgs) {
    k_mul_cycles<<<c[0], c[1]>>>(1u, 500, d); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("blocks=%d threads=%d: %.1f cycles per fp_mul per warp (block 0)\n", c[0], c[1], h[0] / 1000.0);
  }
  return 0;
}
