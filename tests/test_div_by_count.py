"""The segmentation scan divides the running centroid by the small integer cnt + 1 with one reciprocal and Markstein's
FMA correction (csrc/segment.cuh, the merged-centroid update of the scan) instead of 24 IEEE divisions per lane and frame.  NumPy divides with
the correctly rounded quotient, so the sequence must equal `a / n` bit for bit: replayed here in C with hardware FMA
for every n <= 4096 against random numerators over 80 binades (the theorem covers all n whose significand is not all
ones, i.e. every integer below 2^24 - 1)."""
import os
import subprocess
import tempfile

SRC = r"""
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void) {
  long bad = 0;
  for (int n = 1; n <= 4096; ++n) {
    volatile float fn = (float)n;
    const float r = 1.0f / fn;                       /* __frcp_rn */
    for (int k = 0; k < 3000; ++k) {
      uint32_t bits = (uint32_t)rnd();
      const uint32_t e = 87 + (bits >> 8) % 80;
      bits = (bits & 0x807fffffu) | (e << 23);
      float a;
      memcpy(&a, &bits, 4);
      const float q0 = a * r;
      const float rem = fmaf(-q0, fn, a);
      const float q = fmaf(rem, r, q0);
      if (q != a / fn) ++bad;
    }
  }
  printf("%ld\n", bad);
  return 0;
}
"""


def test_reciprocal_plus_fma_correction_is_the_rounded_quotient():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "d.c"), os.path.join(d, "d")
        open(src, "w").write(SRC)
        subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, src, "-lm"])
        assert subprocess.check_output([exe]).strip() == b"0"
