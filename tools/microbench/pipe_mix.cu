// pipe_mix.cu -- (1) clean issue rates of IMAD / IMAD.HI / IMAD.WIDE (no alu ops in the loop: the product feeds the
// next multiplier), (2) do the fp64 pipe (DFMA) and the fma pipe (the fp_mul carry chains) run concurrently when
// different warps of one SM sub-partition use them?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pairing_b200/csrc/fp.cuh"
using namespace bls;

#define NCH 8
#define UNR 16
// FORM 0: mul.wide.u32 (lo,hi) = lo * hi         IMAD.WIDE.U32 R, R.lo, R.hi, RZ
// FORM 1: mul.lo.u32   r = a * r                 IMAD
// FORM 2: mul.hi.u32   r = a * r (hi)            IMAD.HI.U32
// FORM 3: mad.lo.cc/madc.hi pair, no carry chain IMAD.WIDE.U32 R, a, R.lo, R  (accumulating, plain)
// FORM 4: fma.rn.f64 chain                       DFMA
template <int FORM>
__global__ void __launch_bounds__(256) k_rate(uint32_t seed, int iters, uint64_t* sink) {
  uint32_t lo[NCH], hi[NCH], a[NCH];
  double d[NCH];
  uint32_t b0 = seed ^ (threadIdx.x * 2654435761u);
  double db = 1.0 + 1e-9 * (double)(threadIdx.x & 7);
#pragma unroll
  for (int i = 0; i < NCH; i++) { lo[i] = seed + i + b0; hi[i] = b0 ^ i; a[i] = b0 * (2 * i + 3) + 1; d[i] = 1.0 + i; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < UNR; u++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) {
        if (FORM == 0) asm volatile("{.reg .u64 t; mul.wide.u32 t, %1, %0; mov.b64 {%0, %1}, t;}" : "+r"(lo[i]), "+r"(hi[i]));
        if (FORM == 1) asm volatile("mul.lo.u32 %0, %1, %0;" : "+r"(lo[i]) : "r"(a[i]));
        if (FORM == 2) asm volatile("mul.hi.u32 %0, %1, %0;" : "+r"(lo[i]) : "r"(a[i]));
        if (FORM == 3) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(lo[(i + 1) % NCH]));
        if (FORM == 4) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(db));
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s ^= lo[i] ^ ((uint64_t)hi[i] << 32) ^ (uint64_t)__double_as_longlong(d[i]);
  if (s == 0x1234567ull) sink[0] = s;
}

// MODE bit0: integer warps run the fp_mul chain; bit1: the other warps run DFMA chains.
// warp w of a block: (w & 1) == 0 -> integer role, == 1 -> fp64 role (each SMSP gets both kinds with 8 warps per block)
// mode 1: only integer-role warps work (others exit), mode 2: only fp64-role warps work, mode 3: both.
__global__ void __launch_bounds__(256) k_mix(int mode, uint32_t seed, int it_int, int it_dfma, uint64_t* sink) {
  const int warp = threadIdx.x >> 5;
  const bool int_role = ((warp >> 2) & 1) == 0;   // warps 0-3 integer (one per SMSP), 4-7 fp64 (one per SMSP)
  if (int_role) {
    if (!(mode & 1)) return;
    Fp x = fp_one(), y = fp_r2();
    x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed;
#pragma unroll 1
    for (int it = 0; it < it_int; it++) { x = fp_mul_inline(x, y); y = fp_mul_inline(y, x); }
    if (x.v[0] == 0x1234567u && y.v[3] == 7u) sink[0] = x.v[1];
  } else {
    if (!(mode & 2)) return;
    double d[NCH];
    double db = 1.0 + 1e-9 * (double)(threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < NCH; i++) d[i] = 1.0 + i;
#pragma unroll 1
    for (int it = 0; it < it_dfma; it++) {
#pragma unroll
      for (int u = 0; u < UNR; u++)
#pragma unroll
        for (int i = 0; i < NCH; i++) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(db));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += d[i];
    if (s == 0.12345) sink[1] = 1;
  }
}

template <class K> static void run(const char* name, K kern, int iters) {
  uint64_t* sink; cudaMalloc(&sink, 64);
  int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(a);
    kern<<<sm * 8, 256>>>(12345u, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (r && ms < best) best = ms;
  }
  double steps = (double)NCH * UNR * iters * 256.0 * sm * 8;
  double cyc = best * 1e-3 * 1.965e9 / ((double)NCH * UNR * iters * 16.0);
  printf("%-44s %8.3f T op/s  (%.2f ms)  %.2f cyc/op/SMSP  %s\n", name, steps / best * 1e-9, best, cyc, cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink);
}

template <class K> static void run_occ(const char* name, K kern, int iters, int threads, int blocks_per_sm) {
  uint64_t* sink; cudaMalloc(&sink, 64);
  int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(a);
    kern<<<sm * blocks_per_sm, threads>>>(12345u, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (r && ms < best) best = ms;
  }
  double warps_per_smsp = threads / 32.0 * blocks_per_sm / 4.0;
  double cyc = best * 1e-3 * 1.965e9 / ((double)NCH * UNR * iters * warps_per_smsp);
  printf("%-40s %4d thr x %d blk/SM (%.1f warps/SMSP): %.2f cyc/op/SMSP\n", name, threads, blocks_per_sm, warps_per_smsp, cyc);
  cudaFree(sink);
}
__global__ void __launch_bounds__(256) k_fpmul_chain(uint32_t seed, int iters, uint64_t* sink) {
  Fp x = fp_one(), y = fp_r2();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed;
#pragma unroll 1
  for (int it = 0; it < iters; it++) { x = fp_mul_inline(x, y); y = fp_mul_inline(y, x); }
  if (x.v[0] == 0x1234567u && y.v[3] == 7u) sink[0] = x.v[1];
}
__global__ void __launch_bounds__(256) k_fpmul2_chain(uint32_t seed, int iters, uint64_t* sink) {
  Fp x = fp_one(), y = fp_r2(), z = fp_r2();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed; z.v[2] ^= seed;
#pragma unroll 1
  for (int it = 0; it < iters; it++) { x = fp_mul2_inline(x, y, z, x); y = fp_mul2_inline(y, x, z, y); }
  if (x.v[0] == 0x1234567u && y.v[3] == 7u) sink[0] = x.v[1];
}
// two independent products per step (ILP 2)
__global__ void __launch_bounds__(256) k_fpmul_ilp2(uint32_t seed, int iters, uint64_t* sink) {
  Fp x = fp_one(), y = fp_r2(), u = fp_r2(), v = fp_one();
  x.v[0] ^= seed ^ threadIdx.x; y.v[1] ^= seed; u.v[2] ^= threadIdx.x; v.v[3] ^= seed;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    Fp x2 = fp_mul_inline(x, y), u2 = fp_mul_inline(u, v);
    y = fp_mul_inline(y, x2); v = fp_mul_inline(v, u2);
    x = x2; u = u2;
  }
  if (x.v[0] == 0x1234567u && y.v[3] == 7u && u.v[1] == 5u && v.v[2] == 9u) sink[0] = x.v[1];
}
static float run_mix(int mode, int it_int, int it_dfma) {
  uint64_t* sink; cudaMalloc(&sink, 64);
  int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(a);
    k_mix<<<sm * 8, 256>>>(mode, 12345u, it_int, it_dfma, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (r && ms < best) best = ms;
  }
  cudaFree(sink);
  return best;
}
int main() {
  run("0 IMAD.WIDE.U32 (mul.wide, RZ addend)", k_rate<0>, 2000);
  run("1 IMAD (mul.lo)", k_rate<1>, 2000);
  run("2 IMAD.HI.U32 (mul.hi)", k_rate<2>, 2000);
  run("3 IMAD.WIDE.U32 accumulate, no carry link", k_rate<3>, 2000);
  run("4 DFMA", k_rate<4>, 2000);
  for (int thr : {128, 256, 384, 512}) run_occ("IMAD.WIDE.U32 chain x8", k_rate<0>, 2000, thr, 1);
  // per fp_mul: iters*2 products of 300 MAC; report cycles per MAC: scale NCH*UNR=128 -> 600 MACs per iter
  for (int thr : {128, 256, 384, 512}) run_occ("fp_mul chain (x 600/128 = cyc/MAC)", k_fpmul_chain, 500, thr, 1);
  for (int thr : {128, 256, 384, 512}) run_occ("fp_mul2 chain (x 888/128)", k_fpmul2_chain, 500, thr, 1);
  for (int thr : {128, 256, 384}) run_occ("fp_mul ilp2 (x 1200/128)", k_fpmul_ilp2, 500, thr, 1);
  // integer warps: 2*it_int fp_mul of 300 MAC32; fp64 warps: 128*it_dfma DFMA.  Balance so that each alone takes similar time.
  int it_int = 1000, it_dfma = 0;
  float t_int = run_mix(1, it_int, 0);
  // DFMA alone at 2 cyc/op with 8 warps/SMSP ... pick it_dfma so t_dfma ~ t_int
  float t_probe = run_mix(2, 0, 1000);
  it_dfma = (int)(1000.0 * t_int / t_probe);
  float t_dfma = run_mix(2, 0, it_dfma);
  float t_both = run_mix(3, it_int, it_dfma);
  printf("mix: fp_mul warps alone %.2f ms, DFMA warps alone %.2f ms, both %.2f ms  (serial sum %.2f, perfect overlap %.2f)\n",
         t_int, t_dfma, t_both, t_int + t_dfma, t_int > t_dfma ? t_int : t_dfma);
  return 0;
}
