"""csrc/glibc_pow.cuh (the operation-by-operation restatement of glibc 2.39's FMA pow that the
probability kernel uses) compiled for the HOST and compared bit for bit with the libm the
reference links against.  The GPU twin of this test is test_gpu_parity.py::test_pow_matches_host_libm."""
import os
import platform
import subprocess
import tempfile

import pytest

import parity_lib as pl

SRC = r"""
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "glibc_pow.cuh"
static unsigned long long s = 88172645463325252ull;
static unsigned long long rnd(void){ s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(int argc, char** argv){
  long n = atol(argv[1]), bad = 0, tested = 0;
  for (long i = 0; i < n; i++) {
    double u = (rnd() >> 11) / 9007199254740992.0, v = (rnd() >> 11) / 9007199254740992.0;
    double dens = (i & 1) ? pow(10.0, -9.0 + 8.7 * u) : u;
    double x = 1.0 - dens;
    if (!(x > 0)) continue;
    int c = (int)pow(10.0, 5.3 * v); if (c < 1) c = 1;
    if (i % 7 == 0) c = 1 + (int)(rnd() % 3000);
    double want = pow(x, (double)c), got = glibc_pow_pos(x, (double)c);
    tested++;
    if (want != got) {
      if (fabs(c * log(x)) >= 512 && (1.0 - want) == (1.0 - got)) continue;
      if (bad < 5) printf("MISMATCH x=%a c=%d want=%a got=%a\n", x, c, want, got);
      bad++;
    }
  }
  printf("tested %ld mismatches %ld\n", tested, bad);
  return bad != 0;
}
"""


def _host_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read() and " avx2 " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(platform.machine() != "x86_64" or not _host_has_fma(),
                    reason="glibc selects its FMA pow only on x86-64 hosts with AVX2+FMA")
def test_pow_port_is_bit_exact_against_host_libm():
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(SRC)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-I", os.path.join(pl.REPO, "i-adhore_b200", "csrc"),
                        "-o", exe, src, "-lm"], check=True)
        p = subprocess.run([exe, "4000000"], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout[-500:]
        assert "mismatches 0" in p.stdout
