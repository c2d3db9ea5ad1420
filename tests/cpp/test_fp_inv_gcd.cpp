// Host build of pairing_b200/csrc/fp_inv_gcd.cuh (the division-step Fq inversion the CUDA kernels run): reads canonical
// operands as 96 hex digits per line on stdin, prints a^-1 * 2^768 mod q and the number of 30-step batches it took.
// Driven by tests/test_host_logic.py, which compares with Python's pow(a, -1, q).
#include <cstdio>
#include <cstring>
#include "../../pairing_b200/csrc/fp_inv_gcd.cuh"

int main() {
  char line[256];
  while (fgets(line, sizeof line, stdin)) {
    if (strlen(line) < 96) continue;
    uint32_t a[12], out[12];
    for (int j = 0; j < 12; j++) {               // most significant word first on the line
      unsigned v;
      sscanf(line + 8 * j, "%8x", &v);
      a[11 - j] = v;
    }
    const int batches = bls::gcd30::invert_words(out, a);
    for (int j = 11; j >= 0; j--) printf("%08x", out[j]);
    printf(" %d\n", batches);
  }
  return 0;
}
