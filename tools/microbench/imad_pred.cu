// imad_pred.cu -- IMAD.WIDE throughput vs carry-predicate usage (standalone probe).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NACC 8
#define UNROLL 16
// MODE 0: mad.wide (no predicate)                        -> IMAD.WIDE.U32
// MODE 1: carry-out only, consumed by an ALU addc        -> IMAD.WIDE.U32 Rd, P, ... ; IADD3.X
// MODE 2: carry-in only (from an ALU add.cc)             -> IADD3 P ; IMAD.WIDE.U32.X Rd, ..., P
// MODE 3: carry-in and carry-out (chained)               -> IMAD.WIDE.U32.X Rd, P, ..., P
// MODE 4: like 1 but the carry is dropped into a per-accumulator counter (deferred carries)
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(uint32_t seed, int iters, uint32_t* sink) {
  uint32_t lo[NACC], hi[NACC], c[NACC], a[NACC];
  uint32_t b = seed ^ (threadIdx.x * 2654435761u);
#pragma unroll
  for (int i = 0; i < NACC; i++) { lo[i] = seed + i; hi[i] = seed * 7 + i; c[i] = i; a[i] = b * (2 * i + 3) + 1; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
          uint64_t t = ((uint64_t)hi[i] << 32) | lo[i];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[i]), "r"(b));
          lo[i] = (uint32_t)t; hi[i] = (uint32_t)(t >> 32);
        }
      } else if (MODE == 1 || MODE == 4) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
          asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                       : "+r"(lo[i]), "+r"(hi[i]), "+r"(c[i]) : "r"(a[i]), "r"(b));
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
          asm volatile("add.cc.u32 %2, %2, %3;\n\tmadc.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.u32 %1, %3, %4, %1;"
                       : "+r"(lo[i]), "+r"(hi[i]), "+r"(c[i]) : "r"(a[i]), "r"(b));
      } else if (MODE == 3) {
        // one chain through all accumulators: carry-in and carry-out on every link
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
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s ^= lo[i] ^ hi[i] ^ c[i];
  if (s == 0x1234567u) sink[0] = s;
}

template <class K> static void run(const char* name, K kern, int iters) {
  uint32_t* sink; cudaMalloc(&sink, 64);
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
  double total = (double)NACC * UNROLL * iters * 256.0 * sm * 8;
  printf("%-52s %8.3f T MAC/s  (%.2f ms) %s\n", name, total / best * 1e-9, best, cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink);
}
int main() {
  run("0: IMAD.WIDE.U32, no predicate", k<0>, 4000);
  run("1: carry-out only (+ IADD3.X consumer on ALU pipe)", k<1>, 4000);
  run("2: carry-in only (+ IADD3 producer on ALU pipe)", k<2>, 4000);
  run("3: carry-in and carry-out (chained)", k<3>, 4000);
  return 0;
}
