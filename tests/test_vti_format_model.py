"""CPU-side checks of the device "%g" formatter of the VTI writer (csrc/vti.cu):
 * the 128-bit power-of-ten table it includes is an exact bracket of 10^k,
 * an integer model of its algorithm (same table, same decisions) reproduces printf("%g") on random and
   adversarial doubles -- so the GPU tests only have to show that the CUDA code implements the model."""
import math
import os
import random
import re
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "pd_mg_pin_corrosion_b200", "csrc", "pow10_table.inc")


def _table():
    text = open(INC).read()
    kmin = int(re.search(r"PD_POW10_KMIN \((-?\d+)\)", text).group(1))
    rows = re.findall(r"\{0x([0-9a-f]{16})ull, 0x([0-9a-f]{16})ull, (-?\d+), (\d)\}", text)
    return kmin, [((int(h, 16) << 64) | int(lo, 16), int(b), int(e)) for h, lo, b, e in rows]


def test_pow10_table_brackets_powers_of_ten():
    kmin, rows = _table()
    assert len(rows) == 641 and kmin == -320
    for idx, (T, b, exact) in enumerate(rows):
        k = kmin + idx
        assert (1 << 127) <= T < (1 << 128)
        # T 2^b <= 10^k < (T+1) 2^b, compared as integers after clearing the denominators
        if k >= 0:
            lhs, mid, rhs = (T << b, 10 ** k, (T + 1) << b) if b >= 0 else (T, 10 ** k << -b, T + 1)
        else:
            lhs, mid, rhs = (T * 10 ** -k, 1 << -b, (T + 1) * 10 ** -k)
        assert lhs <= mid < rhs, k
        assert (lhs == mid) == bool(exact), k


def _model_g(v: float, table) -> str:
    """Integer model of fmt_g/dec6 in csrc/vti.cu."""
    kmin, rows = table
    if math.isnan(v) or math.isinf(v):
        v = 0.0
    if v != 0.0 and abs(v) < 1e-300:
        v = 0.0
    bits = struct.unpack("<Q", struct.pack("<d", v))[0]
    out = "-" if bits >> 63 else ""
    if v == 0.0:
        return out + "0"
    be = (bits >> 52) & 0x7FF
    m = (bits & ((1 << 52) - 1)) | (1 << 52)
    e = be - 1075
    X = ((be - 1023) * 78913) >> 18
    flag = False
    for _ in range(3):
        k = 5 - X
        T, b, exact = rows[k - kmin]
        P = m * T
        s = -(e + b)
        I = P >> s
        if I >= 1000000:
            X += 1
            continue
        if I < 99999:
            X -= 1
            continue
        F = P & ((1 << s) - 1)
        half = 1 << (s - 1)
        if exact:
            up = (I & 1) == 1 if F == half else F > half
        elif F >= half:
            up = True
        else:
            Pu = P + m
            Iu, Fu = Pu >> s, Pu & ((1 << s) - 1)
            if Iu == I and Fu <= half:
                up = False
            else:
                tie = False
                if k < 0 and -k <= 22:          # exact tie: m 2^(e+1) == (2I+1) 10^-k
                    R, sl = (2 * I + 1) * 10 ** -k, e + 1
                    tie = (m << sl) == R if sl >= 0 else (m % (1 << -sl) == 0 and (m >> -sl) == R)
                if tie:
                    up = (I & 1) == 1
                else:
                    up, flag = True, True
        if I + (1 if up else 0) < 100000:
            X -= 1
            continue
        break
    assert not flag, repr(v)
    q = I + (1 if up else 0)
    if q >= 1000000:
        q, X = 100000, X + 1
    d = "%06d" % q
    nd = len(d.rstrip("0")) or 1
    if X < -4 or X >= 6:
        mant = d[0] + ("." + d[1:nd] if nd > 1 else "")
        return out + mant + "e" + ("-" if X < 0 else "+") + ("%02d" % abs(X))
    if X >= 0:
        return out + d[:X + 1] + ("." + d[X + 1:nd] if nd > X + 1 else "")
    return out + "0." + "0" * (-X - 1) + d[:nd]


def _printf_g(v: float) -> str:
    if math.isnan(v) or math.isinf(v):
        v = 0.0
    if v != 0.0 and abs(v) < 1e-300:
        v = 0.0
    return "%g" % v


def test_model_matches_printf():
    table = _table()
    rng = random.Random(11)
    vals = [0.0, -0.0, 1.0, 1e6, 1e5, 999999.5, 100000.5, 1000005.0, 1234565.0, 9.999995e-5, 1e-4, 1e-5, 1e22, 1e23,
            1e-300, 5e-324, 1.7976931348623157e308, float("nan"), float("inf"), 0.1, 1 / 3, 2.5, 123456789012345678.0]
    vals += [float(10 ** k) for k in range(23)] + [float(2 ** k) for k in range(-70, 71)]
    vals += [(2 * n + 1) * 5.0 ** j * 2.0 ** (j - 1) for n in (100000, 123456, 499999, 314159) for j in range(1, 12)]
    for _ in range(200000):
        vals.append(struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0])
    for _ in range(100000):   # near the decimal half-way points of 6-digit numbers
        base = (rng.randrange(100000, 1000000) + 0.5) * 10.0 ** rng.randrange(-15, 15)
        vals.append(base * (1.0 + rng.randrange(-3, 4) * 2.0 ** -52))
    for v in vals:
        assert _model_g(v, table) == _printf_g(v), (repr(v), _model_g(v, table), _printf_g(v))
