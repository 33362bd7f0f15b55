// Host harness of the DEVICE generator: csrc/rng.cuh::Philox::gen is __host__ __device__, so the very code k_rng_fill runs
// is compiled for the CPU here and checked against the Random123 known-answer vectors (tests/test_rng_kat.py).
#include <cstdio>
#include "rng.cuh"
int main(int argc, char** argv) {
  // stdin: lines "seed i agent step stream" (hex) -> 4 hex words
  unsigned long long seed; unsigned i, a, s, t;
  while (scanf("%llx %x %x %x %x", &seed, &i, &a, &s, &t) == 5) {
    uint32_t o[4]; saceo::Philox::gen(seed, i, a, s, t, o);
    printf("%08x %08x %08x %08x\n", o[0], o[1], o[2], o[3]);
  }
  return 0;
}
