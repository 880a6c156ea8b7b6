#!/usr/bin/env python3
"""Generator for the in-register FFT codelets used by the STFT kernels.

Every lane of a warp holds R complex points in registers and runs one of these straight-line
codelets on them (R = 16, 32, 64): forward DFTs (kernel exp(-2*pi*i*j*k/R)), natural order in,
natural order out.  The inverse (un-normalised) is the same codelet on component-swapped data.

The codelets are written for Blackwell's packed fp32 pipe (sm_100: FADD2 / FMUL2 / FFMA2, exposed
as __fadd2_rn / __fmul2_rn / __ffma2_rn): a complex number is one float2 register pair, so
  * a complex add/sub -- including operands multiplied by -1, -i or +i, which are free operand
    modifiers (half swap `.LO_HI`, per-half negate `.NP`) -- is ONE instruction,
  * a twiddle multiplication  x*(c + i s) = c*(a, b) + s*(-b, a)  is TWO (FMUL2 + FFMA2).
fft32 is 224 packed instructions instead of 432 scalar ones.

Twiddle constants are rounded to fp32 from fp64 values here, at generation time (the reference
path runs torch.stft -> pocketfft/MKL/cuFFT, all of which use correctly rounded tables).

Usage:  python gen_fft_codelets.py            # rewrites fft_codelets.cuh
        python gen_fft_codelets.py --check    # numerically verifies the IR vs numpy.fft

The emitted header is committed; the build does not run this script.
"""
import math
import os
import sys
from fractions import Fraction

import numpy as np


class CVal:
    """A complex value: (-i)^u * register."""
    __slots__ = ("name", "u")

    def __init__(self, name, u=0):
        self.name, self.u = name, u % 4

    def rot(self, k):          # multiply by (-i)^k
        return CVal(self.name, self.u + k)


class Prog:
    """SSA IR at the complex level: ('add', dst, a, ua, b, ub) | ('tw', dst, a, c, s)  (dst = a * (c + i s)) |
    ('mul', dst, a, ua, c)  (dst = c * a, c real) | ('fma', dst, a, ua, c, b, ub)  (dst = c * a + b, c real)."""

    def __init__(self):
        self.ops, self.n = [], 0

    def _new(self):
        self.n += 1
        return "t%d" % self.n

    def add(self, a, b):
        d = self._new()
        self.ops.append(("add", d, a.name, a.u, b.name, b.u))
        return CVal(d)

    def sub(self, a, b):
        return self.add(a, b.rot(2))

    def scale(self, a, c):
        d = self._new()
        self.ops.append(("mul", d, a.name, a.u, float(np.float32(c))))
        return CVal(d)

    def fma(self, a, c, b):
        d = self._new()
        self.ops.append(("fma", d, a.name, a.u, float(np.float32(c)), b.name, b.u))
        return CVal(d)

    def twiddle(self, a, k, n):
        """a * exp(-2*pi*i*k/n); multiples of a quarter turn are free rotations."""
        fr = Fraction(k % n, n)
        if fr.denominator in (1, 2, 4):
            return a.rot(int(fr * 4))
        th = 2.0 * math.pi * (k % n) / n
        w = complex(math.cos(th), -math.sin(th)) * (-1j) ** a.u      # fold the pending rotation into the constant
        d = self._new()
        self.ops.append(("tw", d, a.name, float(np.float32(w.real)), float(np.float32(w.imag))))
        return CVal(d)


def fft(p, x):
    n = len(x)
    if n == 1:
        return list(x)
    if n == 2:
        return [p.add(x[0], x[1]), p.sub(x[0], x[1])]
    if n == 4:
        t0, t1 = p.add(x[0], x[2]), p.sub(x[0], x[2])
        t2, t3 = p.add(x[1], x[3]), p.sub(x[1], x[3]).rot(1)
        return [p.add(t0, t2), p.add(t1, t3), p.sub(t0, t2), p.sub(t1, t3)]
    if n == 5:
        # radix-5 butterfly: 8 FADD2 + 2 FMUL2 + 8 FFMA2 (real constants on the packed pipe), +-i rotations free
        c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
        s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
        a1, a2 = p.add(x[1], x[4]), p.add(x[2], x[3])
        b1, b2 = p.sub(x[1], x[4]), p.sub(x[2], x[3])
        y0 = p.add(p.add(x[0], a1), a2)
        m1 = p.fma(a2, c2, p.fma(a1, c1, x[0]))
        m2 = p.fma(a2, c1, p.fma(a1, c2, x[0]))
        n1 = p.fma(b2, s2, p.scale(b1, s1))
        n2 = p.fma(b2, -s1, p.scale(b1, s2))
        return [y0, p.add(m1, n1.rot(1)), p.add(m2, n2.rot(1)), p.add(m2, n2.rot(3)), p.add(m1, n1.rot(3))]
    a, b = {8: (2, 4), 16: (4, 4), 32: (4, 8), 64: (8, 8), 25: (5, 5)}[n]
    # n = a*b, input index j + b*i (j<b, i<a), output index i' + a*j'
    cols = []
    for j in range(b):
        sub = fft(p, [x[j + b * i] for i in range(a)])
        cols.append([p.twiddle(sub[i2], j * i2, n) for i2 in range(a)])
    out = [None] * n
    for i2 in range(a):
        sub = fft(p, [cols[j][i2] for j in range(b)])
        for j2 in range(b):
            out[i2 + a * j2] = sub[j2]
    return out


def build(n):
    p = Prog()
    out = fft(p, [CVal("v[%d]" % i) for i in range(n)])
    return p, out


def fl(c):
    r = repr(float(np.float32(c)))
    if "e" not in r and "." not in r and "inf" not in r:
        r += ".0"
    return r + "f"


def operand(name, u):
    if u == 0:
        return name
    if u == 1:   # * -i : (y, -x)
        return "make_float2(%s.y, -%s.x)" % (name, name)
    if u == 2:   # * -1
        return "make_float2(-%s.x, -%s.y)" % (name, name)
    return "make_float2(-%s.y, %s.x)" % (name, name)          # * +i : (-y, x)


def render_c(n):
    p, out = build(n)
    n_add = sum(1 for o in p.ops if o[0] == "add")
    n_tw = sum(1 for o in p.ops if o[0] == "tw")
    n_real = len(p.ops) - n_add - n_tw
    L = ["// %d-point forward DFT, in place, natural order: %d FADD2 + %d twiddles (FMUL2 + FFMA2)%s = %d packed instructions."
         % (n, n_add, n_tw, " + %d real-constant FMUL2/FFMA2" % n_real if n_real else "", n_add + 2 * n_tw + n_real),
         "SPL_DEVICE void fft%d(float2 (&v)[%d]) {" % (n, n)]
    for op in p.ops:
        if op[0] == "add":
            _, d, a, ua, b, ub = op
            L.append("  const float2 %s = __fadd2_rn(%s, %s);" % (d, operand(a, ua), operand(b, ub)))
        elif op[0] == "mul":
            _, d, a, ua, c = op
            L.append("  const float2 %s = __fmul2_rn(%s, make_float2(%s, %s));" % (d, operand(a, ua), fl(c), fl(c)))
        elif op[0] == "fma":
            _, d, a, ua, c, b, ub = op
            L.append("  const float2 %s = __ffma2_rn(%s, make_float2(%s, %s), %s);" % (d, operand(a, ua), fl(c), fl(c), operand(b, ub)))
        else:
            _, d, a, c, s = op
            L.append("  const float2 %s = __ffma2_rn(make_float2(-%s.y, %s.x), make_float2(%s, %s), "
                     "__fmul2_rn(%s, make_float2(%s, %s)));" % (d, a, a, fl(s), fl(s), a, fl(c), fl(c)))
    for k, o in enumerate(out):
        L.append("  v[%d] = %s;" % (k, operand(o.name, o.u)))
    L.append("}")
    return "\n".join(L), n_add + 2 * n_tw + n_real


def evaluate(n, z):
    """Run the IR in fp32 numpy on a (batch, n) complex array; returns complex64-precision results."""
    p, out = build(n)
    f32 = np.float32
    env = {"v[%d]" % i: (z[:, i].real.astype(f32), z[:, i].imag.astype(f32)) for i in range(n)}

    def opnd(name, u):
        x, y = env[name]
        return [(x, y), (y, -x), (-x, -y), (-y, x)][u]

    for op in p.ops:
        if op[0] == "add":
            _, d, a, ua, b, ub = op
            (ax, ay), (bx, by) = opnd(a, ua), opnd(b, ub)
            env[d] = ((ax + bx).astype(f32), (ay + by).astype(f32))
        elif op[0] == "mul":
            _, d, a, ua, c = op
            ax, ay = opnd(a, ua)
            env[d] = ((ax * f32(c)).astype(f32), (ay * f32(c)).astype(f32))
        elif op[0] == "fma":
            _, d, a, ua, c, b, ub = op
            (ax, ay), (bx, by) = opnd(a, ua), opnd(b, ub)
            env[d] = ((ax.astype(np.float64) * np.float64(f32(c)) + bx).astype(f32),
                      (ay.astype(np.float64) * np.float64(f32(c)) + by).astype(f32))
        else:
            _, d, a, c, s = op
            x, y = env[a]
            tx, ty = (x * f32(c)).astype(f32), (y * f32(c)).astype(f32)
            # fused multiply-add: one rounding
            env[d] = ((-y.astype(np.float64) * np.float64(f32(s)) + tx).astype(f32),
                      (x.astype(np.float64) * np.float64(f32(s)) + ty).astype(f32))
    res = []
    for o in out:
        x, y = opnd(o.name, o.u)
        res.append(x.astype(np.float64) + 1j * y.astype(np.float64))
    return np.stack(res, axis=1)


def check():
    rng = np.random.default_rng(0)
    ok = True
    for n in (16, 25, 32, 64):
        z = rng.standard_normal((64, n)) + 1j * rng.standard_normal((64, n))
        z = z.astype(np.complex64).astype(np.complex128)
        got = evaluate(n, z)
        ref = np.fft.fft(z, axis=1)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        zs = z.imag + 1j * z.real                                 # inverse through swapped components
        gi = evaluate(n, zs)
        gi = gi.imag + 1j * gi.real
        refi = np.fft.ifft(z, axis=1) * n
        erri = np.abs(gi - refi).max() / np.abs(refi).max()
        p, _ = build(n)
        n_add = sum(1 for o in p.ops if o[0] == "add")
        n_tw = sum(1 for o in p.ops if o[0] == "tw")
        print("fft%d: %d FADD2 + %d twiddles + %d real-constant ops, fwd_err=%.2e inv_err=%.2e"
              % (n, n_add, n_tw, len(p.ops) - n_add - n_tw, err, erri))
        ok &= err < 2e-6 and erri < 2e-6
    return ok


HEADER = """// GENERATED by gen_fft_codelets.py -- do not edit by hand.
// In-register forward DFT codelets (16/25/32/64 points) for the warp-per-frame STFT kernels, written for
// the packed fp32 pipe of sm_100 (FADD2 / FMUL2 / FFMA2): one float2 register pair per complex point.
// Inverse (un-normalised): run the same codelet on component-swapped data, (y, x) in -> (y, x) out.
#pragma once
#ifndef SPL_DEVICE
#define SPL_DEVICE __device__ __forceinline__
#endif
"""


def main():
    if "--check" in sys.argv:
        sys.exit(0 if check() else 1)
    parts = [HEADER]
    for n in (16, 25, 32, 64):
        src, _ = render_c(n)
        parts.append(src)
        parts.append("")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fft_codelets.cuh")
    with open(path, "w") as f:
        f.write("\n".join(parts))
    print("wrote", path)


if __name__ == "__main__":
    main()
