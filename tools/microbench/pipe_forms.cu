// pipe_forms.cu -- issue rates of the integer-multiply forms a Montgomery product can be built from,
// and whether the alu pipe (IADD3/LOP3/SHF) and the fp64 pipe co-issue with the fma pipe (IMAD.WIDE).
// Standalone probe: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_forms pipe_forms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCH 8
#define UNR 16

// FORM 0: plain mad.wide.u32                      (IMAD.WIDE.U32)
// FORM 1: carry-out only, carry counted on alu    (IMAD.WIDE.U32 R,P + IADD3.X)
// FORM 2: carry-in only from an alu add           (IADD3 P + IMAD.WIDE.U32.X)
// FORM 3: carry-in and carry-out chain of NCH     (IMAD.WIDE.U32.X R,P ... P)
// FORM 4: plain mad.wide + 1 independent LOP3 per MAC
// FORM 5: plain mad.wide + 1 independent IADD3 per MAC
// FORM 6: plain mad.wide + 1 independent SHF per MAC
// FORM 7: DFMA only
// FORM 8: DFMA + mad.wide 1:1
// FORM 9: mad.wide.s32
// FORM 10: plain mad.wide + 2 alu ops per MAC
// FORM 11: alu only (IADD3)
// FORM 12: mad.lo.u32 (IMAD) + mad.wide 1:1
// FORM 13: 64-bit add only (IADD3 + IADD3.X pairs)
template <int FORM>
__global__ void __launch_bounds__(256) k(uint32_t seed, int iters, uint64_t* sink) {
  uint64_t acc[NCH];
  uint32_t a[NCH], cnt[NCH], x[NCH];
  double d[NCH];
  uint32_t b0 = seed ^ (threadIdx.x * 2654435761u);
  double db = 1.0 + 1e-9 * (double)(threadIdx.x & 7);
#pragma unroll
  for (int i = 0; i < NCH; i++) { acc[i] = seed + i + b0; a[i] = b0 * (i + 3) + 1; cnt[i] = i; x[i] = b0 + 77 * i; d[i] = 1.0 + i; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      if (FORM == 3) {
        uint32_t lo[NCH], hi[NCH];
        const uint32_t b = (uint32_t)acc[u % NCH] | 1u;
#pragma unroll
        for (int i = 0; i < NCH; i++) { lo[i] = (uint32_t)acc[i]; hi[i] = (uint32_t)(acc[i] >> 32); }
        asm volatile("mad.lo.cc.u32 %0, %16, %24, %0;\n\tmadc.hi.cc.u32 %1, %16, %24, %1;\n\t"
            "madc.lo.cc.u32 %2, %17, %24, %2;\n\tmadc.hi.cc.u32 %3, %17, %24, %3;\n\t"
            "madc.lo.cc.u32 %4, %18, %24, %4;\n\tmadc.hi.cc.u32 %5, %18, %24, %5;\n\t"
            "madc.lo.cc.u32 %6, %19, %24, %6;\n\tmadc.hi.cc.u32 %7, %19, %24, %7;\n\t"
            "madc.lo.cc.u32 %8, %20, %24, %8;\n\tmadc.hi.cc.u32 %9, %20, %24, %9;\n\t"
            "madc.lo.cc.u32 %10, %21, %24, %10;\n\tmadc.hi.cc.u32 %11, %21, %24, %11;\n\t"
            "madc.lo.cc.u32 %12, %22, %24, %12;\n\tmadc.hi.cc.u32 %13, %22, %24, %13;\n\t"
            "madc.lo.cc.u32 %14, %23, %24, %14;\n\tmadc.hi.u32 %15, %23, %24, %15;"
            : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
              "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b));
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = ((uint64_t)hi[i] << 32) | lo[i];
      } else {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
          // multiplier varies every step (low word of the neighbouring accumulator), so ptxas cannot hoist the product
          const uint32_t b = (uint32_t)acc[(i + 1) % NCH] | 1u;
          if (FORM == 0 || FORM == 4 || FORM == 5 || FORM == 6 || FORM == 8 || FORM == 10 || FORM == 12)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b));
          if (FORM == 9) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b));
          if (FORM == 1) {
            uint32_t lo = (uint32_t)acc[i], hi = (uint32_t)(acc[i] >> 32);
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                : "+r"(lo), "+r"(hi), "+r"(cnt[i]) : "r"(a[i]), "r"(b));
            acc[i] = ((uint64_t)hi << 32) | lo;
          }
          if (FORM == 2) {
            uint32_t lo = (uint32_t)acc[i], hi = (uint32_t)(acc[i] >> 32);
            asm volatile("add.cc.u32 %2, %2, %3;\n\tmadc.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.u32 %1, %3, %4, %1;"
                : "+r"(lo), "+r"(hi), "+r"(cnt[i]) : "r"(a[i]), "r"(b));
            acc[i] = ((uint64_t)hi << 32) | lo;
          }
          if (FORM == 4 || FORM == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a[i]), "r"(b));
          if (FORM == 5 || FORM == 10 || FORM == 11) asm volatile("add.u32 %0, %0, %1;" : "+r"(cnt[i]) : "r"(a[i]));
          if (FORM == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(a[i]));
          if (FORM == 7 || FORM == 8) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(db));
          if (FORM == 12) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a[i]), "r"(b));
          if (FORM == 13) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(cnt[i]) : "r"(a[i]), "r"(b));
        }
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s ^= acc[i] ^ cnt[i] ^ x[i] ^ (uint64_t)__double_as_longlong(d[i]);
  if (s == 0x1234567ull) sink[0] = s;
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
  double steps = (double)NCH * UNR * iters * 256.0 * sm * 8;   // "MAC slots" (one per chain per unroll step)
  // cycles per step per SMSP at 1.965 GHz: warps per SMSP = 16
  double cyc = best * 1e-3 * 1.965e9 / ((double)NCH * UNR * iters * 16.0);
  printf("%-52s %8.3f T steps/s  (%.2f ms)  %.2f cyc/step/SMSP  %s\n", name, steps / best * 1e-9, best, cyc, cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink);
}
int main() {
  const int it = 2000;
  run("0 mad.wide.u32 plain", k<0>, it);
  run("1 carry-out only + IADD3.X count", k<1>, it);
  run("2 IADD3 carry + carry-in-only IMAD.WIDE.X", k<2>, it);
  run("3 carry chain in+out (8 long)", k<3>, it);
  run("4 mad.wide + LOP3", k<4>, it);
  run("5 mad.wide + IADD", k<5>, it);
  run("6 mad.wide + SHF", k<6>, it);
  run("7 DFMA", k<7>, it);
  run("8 DFMA + mad.wide", k<8>, it);
  run("9 mad.wide.s32", k<9>, it);
  run("10 mad.wide + LOP3 + IADD", k<10>, it);
  run("11 IADD only", k<11>, it);
  run("12 mad.wide + mad.lo", k<12>, it);
  run("13 64-bit add (IADD3 + IADD3.X)", k<13>, it);
  return 0;
}
