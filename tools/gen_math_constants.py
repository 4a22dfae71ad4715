"""Generate the polynomial / reduction constants of the shared arithmetic specification as hex-float
literals (exactly the same doubles on the GPU and in the CPU oracle).

Writes   toss_b200/csrc/tmath_constants.h   (TM_*  ; `TM_CONST` qualifier, __constant__ on the device)
         oracle/orc_math_constants.h         (ORC_* ; static const)

All series are plain Taylor series of the function named; the number of terms is fixed by the argument
range each call site guarantees (see tmath.cuh / orc_math.h):
  ATANH[k] = 1/(2k+3),            k = 0..10   2 atanh(z) = 2z (1 + z^2 sum ATANH[k] z^(2k)),   z^2 < 0.04
  ATAN[k]  = (-1)^(k+1)/(2k+3),   k = 0..21   atan(t) = t (1 + t^2 sum ATAN[k] t^(2k));  8 terms for |t| <= 0.1,
                                               22 terms for |t| <= tan(pi/8)
  ATANH_MM[k], k = 0..6 / ATAN_MM[k], k = 0..4: minimax (Remez, weight u) polynomials in u = z^2 of
      (atanh(z)/z - 1)/u on u <= 0.0401 and (atan(t)/t - 1)/u on u <= 0.01005 -- the common path of the face-pair
      sum; their approximation error (1.6e-17, 4.2e-17 of the function value) is below the rounding error
  ATANH_FAR[k], ATAN_FAR[k], k = 0..3: the same on u <= 0.0025 (errors 2.0e-17, 2.0e-17) -- rows of 32 pairs whose
      every argument is that small (84 % .. 97 % of the rows a trajectory population evaluates)
  ATANH_MID[k], k = 0..4: the same on u <= 0.01005 (error 4.6e-17) -- rows whose every z^2 is below 0.01 (another 12 %
      of the rows of an init-like population)
  ATANH_WIDE[k], k = 0..10: the same on u <= 0.1601 (error 4.0e-18) -- the remaining rows of the lane-per-pair kernels:
      the wider range leaves a third as many rows with terms that need the ln-based evaluation
  SIN[k]   = (-1)^(k+1)/(2k+3)!,  k = 0..8    sin(r) = r + r^3 sum SIN[k] r^(2k),          |r| <= pi/4
  COS[k]   = (-1)^(k+1)/(2k+2)!,  k = 0..8    cos(r) = 1 + r^2 sum COS[k] r^(2k)
"""
from fractions import Fraction as Fr
from math import factorial
from pathlib import Path

import mpmath as mp

mp.mp.dps = 80


def dbl(x) -> float:
    """nearest double of a Fraction / mpf"""
    if isinstance(x, Fr):
        return float(mp.mpf(x.numerator) / mp.mpf(x.denominator))
    return float(x)


def split_hi_lo(x, zero_low_bits=0):
    hi = float(x)
    if zero_low_bits:
        import struct
        b = struct.unpack("<Q", struct.pack("<d", hi))[0]
        b &= ~((1 << zero_low_bits) - 1)
        hi = struct.unpack("<d", struct.pack("<Q", b))[0]
    lo = float(x - mp.mpf(hi))
    return hi, lo


def remez(g, U, n, iters=30):
    """p(u) = sum_{k<n} c_k u^k minimising max |u (p(u) - g(u))| on (0, U]: exchange on n+1 alternation points."""
    U = mp.mpf(U)
    xs = [U * (1 - mp.cos(mp.pi * (i + 1) / (n + 1))) / 2 for i in range(n + 1)]
    c = None
    for _ in range(iters):
        A = mp.matrix(n + 1, n + 1)
        b = mp.matrix(n + 1, 1)
        for i, x in enumerate(xs):
            for k in range(n):
                A[i, k] = x ** (k + 1)
            A[i, n] = (-1) ** i
            b[i] = x * g(x)
        sol = mp.lu_solve(A, b)
        c, E = [sol[k] for k in range(n)], sol[n]

        def err(u):
            return u * (sum(c[k] * u ** k for k in range(n)) - g(u))

        M = 4000
        grid = [U * i / M for i in range(1, M + 1)]
        vals = [err(u) for u in grid]
        ext = []
        for i in range(M):
            left = vals[i - 1] if i > 0 else mp.mpf(0)
            v = vals[i]
            if i + 1 == M:
                if abs(v) > abs(left):
                    ext.append(grid[i])
            elif (v - left) * (vals[i + 1] - v) <= 0 and abs(v) >= abs(left) and abs(v) >= abs(vals[i + 1]):
                lo, hi = (grid[i - 1] if i > 0 else mp.mpf(0)), grid[i + 1]
                for _ in range(60):   # golden-section refinement of the extremum
                    m1, m2 = lo + (hi - lo) * mp.mpf("0.381966"), lo + (hi - lo) * mp.mpf("0.618034")
                    if abs(err(m1)) > abs(err(m2)):
                        hi = m2
                    else:
                        lo = m1
                ext.append((lo + hi) / 2)
        if len(ext) != n + 1:
            ext = sorted(sorted(ext, key=lambda u: -abs(err(u)))[: n + 1])
        done = max(abs(abs(err(u)) - abs(E)) for u in ext) < abs(E) * mp.mpf("1e-6")
        xs = ext
        if done:
            break
    return [float(x) for x in c]


def g_atanh(u):
    if u == 0:
        return mp.mpf(1) / 3
    z = mp.sqrt(u)
    return (mp.atanh(z) / z - 1) / u


def g_atan(u):
    if u == 0:
        return -mp.mpf(1) / 3
    t = mp.sqrt(u)
    return (mp.atan(t) / t - 1) / u


TABLES = {
    "ATANH_MM": remez(g_atanh, "0.0401", 7),
    "ATAN_MM": remez(g_atan, "0.01005", 5),
    "ATANH_FAR": remez(g_atanh, "0.0025", 4),
    "ATAN_FAR": remez(g_atan, "0.0025", 4),
    "ATANH_MID": remez(g_atanh, "0.01005", 5),
    "ATANH_WIDE": remez(g_atanh, "0.1601", 11),
    "ATANH": [dbl(Fr(1, 2 * k + 3)) for k in range(11)],
    "ATAN": [dbl(Fr((-1) ** (k + 1), 2 * k + 3)) for k in range(22)],
    "SIN": [dbl(Fr((-1) ** (k + 1), factorial(2 * k + 3))) for k in range(9)],
    "COS": [dbl(Fr((-1) ** (k + 1), factorial(2 * k + 2))) for k in range(9)],
}
pio2_hi, pio2_lo = split_hi_lo(mp.pi / 2)
ln2_hi, ln2_lo = split_hi_lo(mp.log(2), zero_low_bits=21)     # e * LN2_HI exact for |e| < 2^21
SCALARS = {
    "TWO_OVER_PI": float(2 / mp.pi), "PIO2_HI": pio2_hi, "PIO2_LO": pio2_lo,
    "PI": float(mp.pi), "PIO2": float(mp.pi / 2), "PIO4": float(mp.pi / 4),
    "LN2_HI": ln2_hi, "LN2_LO": ln2_lo, "SQRT2": float(mp.sqrt(2)), "TAN_PIO8": float(mp.tan(mp.pi / 8)),
}


def emit(prefix, qual):
    out = [f"/* GENERATED by tools/gen_math_constants.py -- do not edit. */"]
    for name, vals in TABLES.items():
        out.append(f"{qual} double {prefix}_{name}[{len(vals)}] = {{")
        out += [f"    {v.hex()},   /* {v!r} */" for v in vals]
        out.append("};")
    for name, v in SCALARS.items():
        out.append(f"#define {prefix}_{name} {v.hex()}   /* {v!r} */")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    root = Path(__file__).resolve().parent.parent
    (root / "toss_b200" / "csrc" / "tmath_constants.h").write_text(emit("TM", "static __constant__"))
    (root / "oracle" / "orc_math_constants.h").write_text(emit("ORC", "static const"))
    print("constants written")
