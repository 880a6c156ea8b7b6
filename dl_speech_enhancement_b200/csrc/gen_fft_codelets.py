#!/usr/bin/env python3
"""Generator for the in-register FFT codelets used by the STFT kernels.

Every lane of a warp holds R complex points in registers and runs one of these
straight-line codelets on them (R = 16, 32, 64).  The codelets are forward DFTs
(kernel exp(-2*pi*i*j*k/R)), natural order in, natural order out; the inverse
(un-normalised, exp(+...)) is obtained by calling the same codelet with the
real and imaginary arrays swapped, so no second set is generated.

Twiddle constants are rounded to fp32 from fp64 values here, at generation time
(the reference path runs torch.stft -> pocketfft/MKL/cuFFT, all of which use
correctly rounded tables; SURVEY.md section 7 "hard parts").

Usage:  python gen_fft_codelets.py            # rewrites fft_codelets.cuh
        python gen_fft_codelets.py --check    # numerically verifies the IR vs numpy.fft

The emitted header is committed; the build does not run this script.
"""
import math
import os
import sys
from fractions import Fraction

import numpy as np


class Val:
    """A real scalar: +/- name."""
    __slots__ = ("name", "sign")

    def __init__(self, name, sign=1):
        self.name, self.sign = name, sign

    def neg(self):
        return Val(self.name, -self.sign)


class Prog:
    """Tiny SSA IR: ('add', dst, a, sa, b, sb) | ('mul', dst, a, c) | ('fma', dst, a, c, b, sb)."""

    def __init__(self):
        self.ops = []
        self.n = 0

    def _new(self):
        self.n += 1
        return "t%d" % self.n

    def add(self, a, b):
        # result = a + b (signed operands). Keep result positive where possible.
        if a.sign < 0 and b.sign < 0:
            d = self._new()
            self.ops.append(("add", d, a.name, 1, b.name, 1))
            return Val(d, -1)
        d = self._new()
        self.ops.append(("add", d, a.name, a.sign, b.name, b.sign))
        return Val(d, 1)

    def sub(self, a, b):
        return self.add(a, b.neg())

    def mul(self, a, c):
        d = self._new()
        self.ops.append(("mul", d, a.name, float(np.float32(c * a.sign))))
        return Val(d, 1)

    def fma(self, a, c, b):
        # a*c + b
        d = self._new()
        self.ops.append(("fma", d, a.name, float(np.float32(c * a.sign)), b.name, b.sign))
        return Val(d, 1)


def cadd(p, x, y):
    return (p.add(x[0], y[0]), p.add(x[1], y[1]))


def csub(p, x, y):
    return (p.sub(x[0], y[0]), p.sub(x[1], y[1]))


def mul_neg_i(x):  # x * (-i) = (im, -re)
    return (x[1], x[0].neg())


def mul_i(x):  # x * i = (-im, re)
    return (x[1].neg(), x[0])


SQ = math.sqrt(0.5)


def ctwiddle(p, x, k, n):
    """x * exp(-2*pi*i*k/n), specialised for the trivial angles."""
    fr = Fraction(k % n, n)
    if fr == 0:
        return x
    if fr == Fraction(1, 4):
        return mul_neg_i(x)
    if fr == Fraction(1, 2):
        return (x[0].neg(), x[1].neg())
    if fr == Fraction(3, 4):
        return mul_i(x)
    if fr.denominator == 8:
        # odd multiples of 1/8: sqrt(1/2) * (+-1 +- i)
        a, b = x
        m = fr.numerator  # 1,3,5,7
        if m == 1:   # (1 - i)/sqrt2 : (a+b, b-a)
            re, im = p.add(a, b), p.sub(b, a)
        elif m == 3:  # (-1 - i)/sqrt2 : (b-a, -(a+b))
            re, im = p.sub(b, a), p.add(a, b).neg()
        elif m == 5:  # (-1 + i)/sqrt2 : (-(a+b), a-b)
            re, im = p.add(a, b).neg(), p.sub(a, b)
        else:        # (1 + i)/sqrt2 : (a-b, a+b)
            re, im = p.sub(a, b), p.add(a, b)
        return (p.mul(re, SQ), p.mul(im, SQ))
    th = 2.0 * math.pi * (k % n) / n
    c, s = math.cos(th), -math.sin(th)       # W = c + i*s
    a, b = x
    # (a + ib)(c + is) = (ac - bs) + i(as + bc)
    t = p.mul(b, -s)
    re = p.fma(a, c, t)
    t2 = p.mul(b, c)
    im = p.fma(a, s, t2)
    return (re, im)


def fft(p, x):
    n = len(x)
    if n == 1:
        return list(x)
    if n == 2:
        return [cadd(p, x[0], x[1]), csub(p, x[0], x[1])]
    if n == 4:
        t0, t1 = cadd(p, x[0], x[2]), csub(p, x[0], x[2])
        t2, t3 = cadd(p, x[1], x[3]), mul_neg_i(csub(p, x[1], x[3]))
        return [cadd(p, t0, t2), cadd(p, t1, t3), csub(p, t0, t2), csub(p, t1, t3)]
    if n == 8:
        a, b = 2, 4
    elif n == 16:
        a, b = 4, 4
    elif n == 32:
        a, b = 4, 8
    elif n == 64:
        a, b = 8, 8
    else:
        raise ValueError(n)
    # n = a*b, input index j + b*i (j<b, i<a), output index i' + a*j'
    cols = []
    for j in range(b):
        sub = fft(p, [x[j + b * i] for i in range(a)])
        cols.append([ctwiddle(p, sub[i2], j * i2, n) for i2 in range(a)])
    out = [None] * n
    for i2 in range(a):
        sub = fft(p, [cols[j][i2] for j in range(b)])
        for j2 in range(b):
            out[i2 + a * j2] = sub[j2]
    return out


def build(n):
    p = Prog()
    x = [(Val("re[%d]" % i), Val("im[%d]" % i)) for i in range(n)]
    out = fft(p, x)
    return p, out


def fl(c):
    r = repr(float(np.float32(c)))
    if "e" not in r and "." not in r and "inf" not in r:
        r += ".0"
    return r + "f"


def render_c(n):
    p, out = build(n)
    L = []
    L.append("// %d-point forward DFT, in place, natural order. %d fp32 ops." % (n, len(p.ops)))
    L.append("SPL_DEVICE void fft%d(float (&re)[%d], float (&im)[%d]) {" % (n, n, n))
    for op in p.ops:
        if op[0] == "add":
            _, d, a, sa, b, sb = op
            if sa > 0:
                L.append("  const float %s = %s %s %s;" % (d, a, "+" if sb > 0 else "-", b))
            else:
                assert sb > 0
                L.append("  const float %s = %s - %s;" % (d, b, a))
        elif op[0] == "mul":
            _, d, a, c = op
            L.append("  const float %s = %s * %s;" % (d, a, fl(c)))
        else:
            _, d, a, c, b, sb = op
            L.append("  const float %s = fmaf(%s, %s, %s%s);" % (d, a, fl(c), "" if sb > 0 else "-", b))
    for k, (r, i) in enumerate(out):
        L.append("  re[%d] = %s%s; im[%d] = %s%s;" % (k, "" if r.sign > 0 else "-", r.name,
                                                        k, "" if i.sign > 0 else "-", i.name))
    L.append("}")
    return "\n".join(L), len(p.ops)


def evaluate(n, re, im):
    """Run the IR in fp32 numpy on (batch, n) arrays; returns (re, im)."""
    p, out = build(n)
    env = {}
    for i in range(n):
        env["re[%d]" % i] = re[:, i].astype(np.float32)
        env["im[%d]" % i] = im[:, i].astype(np.float32)
    f32 = np.float32
    for op in p.ops:
        if op[0] == "add":
            _, d, a, sa, b, sb = op
            env[d] = (f32(sa) * env[a] + f32(sb) * env[b]).astype(f32)
        elif op[0] == "mul":
            _, d, a, c = op
            env[d] = (env[a] * f32(c)).astype(f32)
        else:
            _, d, a, c, b, sb = op
            # emulate fused multiply-add in fp64 then round once
            env[d] = (env[a].astype(np.float64) * np.float64(f32(c)) + sb * env[b].astype(np.float64)).astype(f32)
    ro = np.stack([o[0].sign * env[o[0].name] for o in out], axis=1)
    io = np.stack([o[1].sign * env[o[1].name] for o in out], axis=1)
    return ro, io


def check():
    rng = np.random.default_rng(0)
    ok = True
    for n in (16, 32, 64):
        re = rng.standard_normal((64, n)).astype(np.float32)
        im = rng.standard_normal((64, n)).astype(np.float32)
        ro, io = evaluate(n, re, im)
        ref = np.fft.fft(re.astype(np.float64) + 1j * im.astype(np.float64), axis=1)
        err = np.abs((ro + 1j * io) - ref).max() / np.abs(ref).max()
        # inverse by swapping the arrays
        io2, ro2 = evaluate(n, im, re)
        refi = np.fft.ifft(re.astype(np.float64) + 1j * im.astype(np.float64), axis=1) * n
        erri = np.abs((ro2 + 1j * io2) - refi).max() / np.abs(refi).max()
        print("fft%d: ops=%d fwd_err=%.2e inv_err=%.2e" % (n, len(build(n)[0].ops), err, erri))
        ok &= err < 2e-6 and erri < 2e-6
    return ok


HEADER = """// GENERATED by gen_fft_codelets.py -- do not edit by hand.
// In-register forward DFT codelets (16/32/64 points) for the warp-per-frame STFT kernels.
// Inverse (un-normalised): call with the arrays swapped, fftN(im, re).
#pragma once
#ifndef SPL_DEVICE
#define SPL_DEVICE __device__ __forceinline__
#endif
"""


def main():
    if "--check" in sys.argv:
        sys.exit(0 if check() else 1)
    parts = [HEADER]
    for n in (16, 32, 64):
        src, nops = render_c(n)
        parts.append(src)
        parts.append("")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fft_codelets.cuh")
    with open(path, "w") as f:
        f.write("\n".join(parts))
    print("wrote", path)


if __name__ == "__main__":
    main()
