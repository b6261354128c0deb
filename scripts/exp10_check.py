"""Accuracy of k_comp's gain routine 10^x (x = -att/20 <= 0), simulated operation by operation with exact FMA
semantics (fractions) and compared with 60-digit arithmetic (mpmath).  Prints the constants of the variant as hex
floats for csrc/b200m_kernels.cuh (c_exp10).

    python scripts/exp10_check.py [samples]

Two variants: "deg13" (reduction by whole powers of two, |r| <= 0.1506, degree-13 Taylor polynomial) and "t8"
(reduction by eighths of a power of two: 10^x = 2^n * 2^(j/8) * 10^r with |r| <= 0.0188, an 8-entry table and a
degree-8 polynomial: five fused multiply-adds fewer per gain)."""
import math, random, struct, sys
from fractions import Fraction as Fr
import mpmath

mpmath.mp.prec = 200
MAGIC = 6755399441055744.0      # 1.5 * 2^52


def rn(fr):                      # round a Fraction to the nearest double (ties to even)
    return float(fr)             # int / int true division is correctly rounded


def fma(a, b, c):
    return rn(Fr(a) * Fr(b) + Fr(c))


def hexf(x):
    return float.hex(x)


LN10 = mpmath.log(10)
TAYLOR = [float(LN10 ** k / mpmath.factorial(k)) for k in range(14)]
LOG2_10 = float(mpmath.log(10, 2))
NEG_LOG10_2_HI = -float(mpmath.log10(2))
NEG_LOG10_2_LO = float(-mpmath.log10(2) - mpmath.mpf(NEG_LOG10_2_HI))
T8 = [float(mpmath.mpf(2) ** (mpmath.mpf(j) / 8)) for j in range(8)]


def insert_exp(p, n):
    bits = struct.unpack("<q", struct.pack("<d", p))[0] + (n << 52)
    return struct.unpack("<d", struct.pack("<q", bits))[0]


def exp10_deg13(x):
    t = fma(x, LOG2_10, MAGIC)
    nf = t - MAGIC
    r = fma(nf, NEG_LOG10_2_HI, x)
    r = fma(nf, NEG_LOG10_2_LO, r)
    p = TAYLOR[13]
    for k in range(12, -1, -1):
        p = fma(p, r, TAYLOR[k])
    return insert_exp(p, int(nf))


def exp10_t8(x, deg=8):
    t = fma(x, 8.0 * LOG2_10, MAGIC)
    mf = t - MAGIC
    r = fma(mf, NEG_LOG10_2_HI / 8.0, x)
    r = fma(mf, NEG_LOG10_2_LO / 8.0, r)
    p = TAYLOR[deg]
    for k in range(deg - 1, -1, -1):
        p = fma(p, r, TAYLOR[k])
    m = int(mf)
    v = rn(Fr(p) * Fr(T8[m & 7]))
    return insert_exp(v, m >> 3)


def exp10_tf(x, nt, deg):
    """exp10_t8f with a table of nt entries (a power of two)"""
    T = [float(mpmath.mpf(2) ** (mpmath.mpf(j) / nt)) for j in range(nt)] if nt not in _TABS else _TABS[nt]
    _TABS[nt] = T
    t = fma(x, nt * LOG2_10, MAGIC)
    mf = t - MAGIC
    r = fma(mf, NEG_LOG10_2_HI / nt, x)
    r = fma(mf, NEG_LOG10_2_LO / nt, r)
    p = TAYLOR[deg]
    for k in range(deg - 1, 0, -1):
        p = fma(p, r, TAYLOR[k])
    s = rn(Fr(p) * Fr(r))
    m = int(mf)
    return insert_exp(fma(T[m % nt], s, T[m % nt]), m // nt)


_TABS = {}


def exp10_t8f(x, deg=8):
    """the same reduction; 10^r - 1 = r * q(r) and the result 2^(j/8) + 2^(j/8) * (r q(r)) in one FMA, so the table
    entry's rounding and the final rounding are the only errors of size"""
    t = fma(x, 8.0 * LOG2_10, MAGIC)
    mf = t - MAGIC
    r = fma(mf, NEG_LOG10_2_HI / 8.0, x)
    r = fma(mf, NEG_LOG10_2_LO / 8.0, r)
    p = TAYLOR[deg]
    for k in range(deg - 1, 0, -1):
        p = fma(p, r, TAYLOR[k])
    s = rn(Fr(p) * Fr(r))
    m = int(mf)
    T = T8[m & 7]
    return insert_exp(fma(T, s, T), m >> 3)


def ulp_err(got, x):
    true = mpmath.mpf(10) ** mpmath.mpf(x)
    e = math.frexp(float(true))[1]
    ulp = mpmath.mpf(2) ** (e - 53)
    return float(abs(mpmath.mpf(got) - true) / ulp)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    rnd = random.Random(7)
    worst = {"deg13": 0.0, "t8": 0.0, "t8f": 0.0, "t8f7": 0.0, "t16d7": 0.0, "t32d6": 0.0, "t32d5": 0.0}
    for i in range(n):
        # attenuations as the compressor produces them: 0 .. 60 dB densely, up to 6000 dB sparsely
        att = rnd.random() * (60.0 if i % 4 else 6000.0) if i % 16 else rnd.random() * 1e-6
        q = att * 0.05
        rr = fma(-q, 20.0, att)
        x = -fma(rr, 0.05, q)
        if not x > -300.0:
            continue
        worst["deg13"] = max(worst["deg13"], ulp_err(exp10_deg13(x), x))
        worst["t8"] = max(worst["t8"], ulp_err(exp10_t8(x), x))
        worst["t8f"] = max(worst["t8f"], ulp_err(exp10_t8f(x), x))
        worst["t8f7"] = max(worst["t8f7"], ulp_err(exp10_t8f(x, 7), x))
        worst["t16d7"] = max(worst["t16d7"], ulp_err(exp10_tf(x, 16, 7), x))
        worst["t32d6"] = max(worst["t32d6"], ulp_err(exp10_tf(x, 32, 6), x))
        worst["t32d5"] = max(worst["t32d5"], ulp_err(exp10_tf(x, 32, 5), x))
    assert exp10_t8(-0.0) == 1.0 and exp10_t8f(-0.0) == 1.0 and exp10_t8f(0.0) == 1.0 and exp10_deg13(-0.0) == 1.0
    print("max error in ulp over", n, "exponents:", worst)
    print("taylor:", ", ".join(hexf(c) for c in TAYLOR[:9]))
    print("8 log2(10):", hexf(8.0 * LOG2_10), " -log10(2)/8 hi:", hexf(NEG_LOG10_2_HI / 8.0), " lo:", hexf(NEG_LOG10_2_LO / 8.0))
    print("2^(j/8):", ", ".join(hexf(c) for c in T8))
    print("32 log2(10):", hexf(32.0 * LOG2_10), " -log10(2)/32 hi:", hexf(NEG_LOG10_2_HI / 32.0), " lo:", hexf(NEG_LOG10_2_LO / 32.0))
    print("2^(j/32):", ", ".join(hexf(c) for c in _TABS[32]))


if __name__ == "__main__":
    main()
