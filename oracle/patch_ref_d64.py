#!/usr/bin/env python3
"""oracle/patch_ref_d64.py — TEST INFRASTRUCTURE ONLY.

Writes a PATCHED copy of two reference files into a scratch directory (never into the repository) so that
oracle/Makefile can compile oracle/_ref/libhvref_d64.so, the checker for BASELINE config 5 (DZ2019-MC, d = 50),
which lies outside the reference's domain (SURVEY fact 1, §8c "Oracle for d=50"):

  c/config.h    MOOCORE_HVAPPROX_DIMENSION_MAX 31 -> 64                                   (c/config.h:28)
  c/hvapprox.c  sphere_area_div_2_pow_d_times_d[] gets entries d = 32..64                 (c/hvapprox.c:168-201)
                recip_phi_d[] gets entries d = 32..63                                     (c/hvapprox.c:700-732)

The appended constants come from the closed forms the reference's own generators use —
tools/hypersphere_volume.py:12-15  (2 pi^(d/2) / Gamma(d/2)) / (d 2^d)  and
tools/rphi_seq.py:21-27            1 / (positive root of x^(d+1) = x + 1) —
evaluated with mpmath at 400 bits and printed as hex long double literals like the reference's.  Everything else
in the two files is byte-for-byte the reference; all other sources are used from where they lie.  DZ2019-HW stays
limited by the reference's 32 closed forms of int sin^m (c/hvapprox.c:367-477) and is not extended.

usage: patch_ref_d64.py <reference c dir> <output dir>
"""
import os
import re
import sys

import mpmath as mp

mp.mp.prec = 400
NEW_MAX = 64


def ld_hex(x, digits=32):
    m, e = mp.frexp(x)
    m *= 2
    e -= 1
    n = int(mp.floor((m - 1) * mp.mpf(16) ** digits))
    return "0x1.%0*xp%+dL" % (digits, n, e)


def c_d(d):
    return 2 * mp.pi ** (mp.mpf(d) / 2) / (mp.gamma(mp.mpf(d) / 2) * d * mp.mpf(2) ** d)


def recip_phi(m):
    return 1 / mp.findroot(lambda x: x ** (m + 1) - x - 1, mp.mpf(1) + mp.mpf(1) / m)


def main():
    src, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)

    cfg = open(os.path.join(src, "config.h")).read()
    cfg2, n = re.subn(r"(#define\s+MOOCORE_HVAPPROX_DIMENSION_MAX\s+)31\b", r"\g<1>%d" % NEW_MAX, cfg)
    assert n == 1, "config.h: MOOCORE_HVAPPROX_DIMENSION_MAX not found"
    open(os.path.join(out, "config.h"), "w").write(cfg2)

    hv = open(os.path.join(src, "hvapprox.c")).read()
    # last entry of the c_d table (d = 31) and of the recip_phi_d table (d = 31): append behind them
    m = re.search(r"(static const long double sphere_area_div_2_pow_d_times_d\[\] = \{.*?// d = 31, value = [^\n]*\n)(\};)", hv, re.S)
    assert m, "hvapprox.c: c_d table not found"
    extra = "".join("    %s, // d = %d (appended by oracle/patch_ref_d64.py)\n" % (ld_hex(c_d(d)), d) for d in range(32, NEW_MAX + 1))
    hv = hv[:m.end(1)] + extra + hv[m.start(2):]
    m = re.search(r"(static const long double recip_phi_d\[\] = \{.*?// d = 31, value = [^\n]*\n)(\s*\};)", hv, re.S)
    assert m, "hvapprox.c: recip_phi_d table not found"
    extra = "".join("        %s, // d = %d (appended by oracle/patch_ref_d64.py)\n" % (ld_hex(recip_phi(d)), d) for d in range(32, NEW_MAX))
    hv = hv[:m.end(1)] + extra + hv[m.start(2):]
    open(os.path.join(out, "hvapprox.c"), "w").write(hv)


if __name__ == "__main__":
    main()
