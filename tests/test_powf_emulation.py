"""The fp64 replay of glibc powf(x, .5f) used by the CUDA segmentation kernel, checked on the host: the same
operation sequence written with fma() must equal libm's powf on this machine (the reference's NumPy calls it)."""
import ctypes
import math
import os
import re
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUH = os.path.join(ROOT, "sylber_b200", "csrc", "powf_tables.cuh")


def _tables():
    src = open(CUH).read()

    def arr(name):
        body = re.search(name + r"\[\d+\] = \{(.*?)\};", src, re.S).group(1)
        return [t.strip() for t in body.replace("\n", " ").split(",") if t.strip()]

    log_tab = [float.fromhex(t) for t in arr("kPowfLogTab")]
    log_poly = [float.fromhex(t) for t in arr("kPowfLogPoly")]
    exp_tab = [int(t.rstrip("ull"), 16) for t in arr("kExp2fTab")]
    shift = float.fromhex(re.search(r"kExp2fShiftScaled = (\S+);", src).group(1))
    exp_poly = [float.fromhex(t) for t in arr("kExp2fPolyScaled")]
    return log_tab, log_poly, exp_tab, shift, exp_poly


def _fma(a, b, c):
    return math.fma(a, b, c) if hasattr(math, "fma") else float(np.longdouble(a) * np.longdouble(b) + np.longdouble(c))


def powf_half_replay(x, T):
    log_tab, A, exp_tab, shift, C = T
    ix = struct.unpack("<I", struct.pack("<f", x))[0]
    tmp = (ix - 0x3F330000) & 0xFFFFFFFF
    i = (tmp >> 19) & 15
    top = tmp & 0xFF800000
    iz = (ix - top) & 0xFFFFFFFF
    k = struct.unpack("<i", struct.pack("<I", top))[0] >> 23
    z = struct.unpack("<f", struct.pack("<I", iz))[0]
    r = _fma(z, log_tab[2 * i], -1.0)
    y0 = log_tab[2 * i + 1] + float(k)
    y = _fma(A[0], r, A[1])
    p = _fma(A[2], r, A[3])
    r2 = r * r
    q = _fma(r, A[4], y0)
    r4 = r2 * r2
    q = _fma(r2, p, q)
    y = _fma(y, r4, q)
    ylogx = 0.5 * y
    kd = ylogx + shift
    ki = struct.unpack("<Q", struct.pack("<d", kd))[0]
    kd -= shift
    rr = ylogx - kd
    t = (exp_tab[ki & 31] + (ki << 47)) & 0xFFFFFFFFFFFFFFFF
    s = struct.unpack("<d", struct.pack("<Q", t))[0]
    zz = _fma(rr, C[0], C[1])
    rr2 = rr * rr
    yy = _fma(rr, C[2], 1.0)
    yy = _fma(zz, rr2, yy)
    yy = yy * s
    return float(np.float32(yy))


def test_replay_matches_libm_powf():
    if not hasattr(math, "fma"):
        import pytest
        pytest.skip("math.fma needs Python >= 3.13; covered by the C check below")
    T = _tables()
    libm = ctypes.CDLL("libm.so.6")
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    rng = np.random.default_rng(0)
    xs = np.concatenate([(rng.random(20000) * 4000 + 1e-8).astype(np.float32),
                         np.float32([float.fromhex("0x1.66bf82p+8"), float.fromhex("0x1.aaf7eap+9"), 1e-8, 768.0 * 9])])
    for x in xs:
        assert powf_half_replay(float(x), T) == libm.powf(float(x), 0.5), float(x).hex()


def test_c_replay_matches_libm_on_a_dense_range(tmp_path):
    """Compile the replay as C (fma from libm, no contraction) and sweep 64 M consecutive floats around the values
    the segmentation produces (sum of squares of 768 features ~ 1e-2 .. 1e5)."""
    T = _tables()
    log_tab, A, exp_tab, shift, C = T
    c_src = tmp_path / "replay.c"
    c_src.write_text(r'''
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static const double LT[32] = {%s};
static const double A[5] = {%s};
static const uint64_t ET[32] = {%s};
static const double SH = %s;
static const double C[3] = {%s};
static double asd(uint64_t u){double d; memcpy(&d,&u,8); return d;}
static uint64_t asu(double d){uint64_t u; memcpy(&u,&d,8); return u;}
static float replay(float x){
  uint32_t ix; memcpy(&ix,&x,4);
  uint32_t tmp = ix - 0x3f330000u; int i = (tmp >> 19) & 15; uint32_t top = tmp & 0xff800000u; uint32_t iz = ix - top;
  int k = (int32_t)top >> 23; float zf; memcpy(&zf,&iz,4); double z = zf;
  double r = fma(z, LT[2*i], -1.0), y0 = LT[2*i+1] + (double)k;
  double y = fma(A[0], r, A[1]), p = fma(A[2], r, A[3]), r2 = r*r, q = fma(r, A[4], y0), r4 = r2*r2;
  q = fma(r2, p, q); y = fma(y, r4, q);
  double ylogx = 0.5*y, kd = ylogx + SH; uint64_t ki = asu(kd); kd -= SH; double rr = ylogx - kd;
  double s = asd(ET[ki & 31] + (ki << 47));
  double zz = fma(rr, C[0], C[1]), rr2 = rr*rr, yy = fma(rr, C[2], 1.0); yy = fma(zz, rr2, yy); yy *= s;
  return (float)yy;
}
int main(void){
  uint64_t bad = 0, n = 0, differs_from_sqrt = 0;
  float lo = 1e-2f, hi = 1e5f; uint32_t a, b; memcpy(&a,&lo,4); memcpy(&b,&hi,4);
  for (uint32_t u = a; u < b; u += 3) { float x; memcpy(&x,&u,4); float w = powf(x, 0.5f);
    if (replay(x) != w) bad++; if (w != sqrtf(x)) differs_from_sqrt++; n++; }
  printf("%%llu %%llu %%llu\n", (unsigned long long)n, (unsigned long long)bad, (unsigned long long)differs_from_sqrt);
  return 0;
}
''' % (", ".join(x.hex() for x in log_tab), ", ".join(x.hex() for x in A), ", ".join(f"0x{v:016x}ull" for v in exp_tab),
       shift.hex(), ", ".join(x.hex() for x in C)))
    exe = tmp_path / "replay"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", str(exe), str(c_src), "-lm"])
    n, bad, dsq = map(int, subprocess.check_output([str(exe)]).split())
    assert n > 60_000_000
    assert bad == 0
    assert dsq > 0          # powf really is not sqrtf on this libm: the replay is needed


def test_tables_match_this_libm():
    """The constants baked into powf_tables.cuh are the ones inside this machine's libm.so.6."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import extract_powf_tables as E
    rd = E.vaddr_reader(E.LIBM)
    log_tab, log_poly, exp_tab, shift, exp_poly = _tables()
    blob = rd(0xB7F80, 256)
    if struct.unpack("<d", blob[:8])[0] != log_tab[0]:
        import pytest
        pytest.skip("different glibc build layout: table addresses moved (emulation still checked above)")
    assert list(struct.unpack("<32d", blob)) == log_tab
    assert list(struct.unpack("<5d", rd(0xB8080, 40))) == log_poly
    assert list(struct.unpack("<32Q", rd(0xB7BE0, 256))) == exp_tab
