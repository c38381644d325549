"""Generates arkflow_b200/csrc/ryu_tables.h (128-bit powers of five for shortest double→decimal, Ryu
[Adams 2018]) and self-checks a line-by-line Python model of the device algorithm against Python's
own shortest repr on random doubles.  The CUDA code in arrow_to_json.cu is a transcription of
`d2d()` below.  Run: python scripts/gen_ryu_tables.py"""
import os
import random
import struct

POW5_INV_BITCOUNT = 125
POW5_BITCOUNT = 125
INV_SIZE = 342
POW5_SIZE = 326
M64 = (1 << 64) - 1


def pow5bits(e):
    return ((e * 1217359) >> 19) + 1


def log10pow2(e):
    return (e * 78913) >> 18


def log10pow5(e):
    return (e * 732923) >> 20


POW5_INV = []
for i in range(INV_SIZE):
    j = pow5bits(i) - 1 + POW5_INV_BITCOUNT
    POW5_INV.append(((1 << j) // (5 ** i)) + 1)
POW5 = []
for i in range(POW5_SIZE):
    p = 5 ** i
    l = p.bit_length()
    POW5.append(p >> (l - POW5_BITCOUNT) if l >= POW5_BITCOUNT else p << (POW5_BITCOUNT - l))


def mulshift64(m, mul, j):
    # ((m * mul) >> j) with a 128-bit mul, exactly as the device does it with 64x64->128 products
    lo, hi = mul & M64, mul >> 64
    b0 = m * lo
    b2 = m * hi
    s = (b0 >> 64) + b2  # < 2^129, the device keeps the low 128 bits: shift below needs j - 64 >= 0 and result < 2^64
    return (s >> (j - 64)) & M64


def multiple_of_pow5(v, p):
    c = 0
    while v % 5 == 0 and v:
        v //= 5
        c += 1
    return c >= p


def d2d(bits):
    """returns (digits:int, exponent10:int) with value = digits * 10^exponent10, digits shortest"""
    ieee_m = bits & ((1 << 52) - 1)
    ieee_e = (bits >> 52) & 0x7FF
    if ieee_e == 0:
        e2, m2 = 1 - 1023 - 52 - 2, ieee_m
    else:
        e2, m2 = ieee_e - 1023 - 52 - 2, (1 << 52) | ieee_m
    accept = (m2 & 1) == 0
    mv = 4 * m2
    mm_shift = 1 if (ieee_m != 0 or ieee_e <= 1) else 0
    vm_tz = vr_tz = False
    if e2 >= 0:
        q = log10pow2(e2) - (1 if e2 > 3 else 0)
        e10 = q
        k = POW5_INV_BITCOUNT + pow5bits(q) - 1
        i = -e2 + q + k
        vr = mulshift64(4 * m2, POW5_INV[q], i)
        vp = mulshift64(4 * m2 + 2, POW5_INV[q], i)
        vm = mulshift64(4 * m2 - 1 - mm_shift, POW5_INV[q], i)
        if q <= 21:
            if mv % 5 == 0:
                vr_tz = multiple_of_pow5(mv, q)
            elif accept:
                vm_tz = multiple_of_pow5(mv - 1 - mm_shift, q)
            else:
                vp -= 1 if multiple_of_pow5(mv + 2, q) else 0
    else:
        q = log10pow5(-e2) - (1 if -e2 > 1 else 0)
        e10 = q + e2
        i = -e2 - q
        k = pow5bits(i) - POW5_BITCOUNT
        j = q - k
        vr = mulshift64(4 * m2, POW5[i], j)
        vp = mulshift64(4 * m2 + 2, POW5[i], j)
        vm = mulshift64(4 * m2 - 1 - mm_shift, POW5[i], j)
        if q <= 1:
            vr_tz = True
            if accept:
                vm_tz = mm_shift == 1
            else:
                vp -= 1
        elif q < 63:
            vr_tz = (mv & ((1 << q) - 1)) == 0
    removed = 0
    last = 0
    if vm_tz or vr_tz:
        while vp // 10 > vm // 10:
            vm_tz = vm_tz and vm % 10 == 0
            vr_tz = vr_tz and last == 0
            last = vr % 10
            vr //= 10
            vp //= 10
            vm //= 10
            removed += 1
        if vm_tz:
            while vm % 10 == 0:
                vr_tz = vr_tz and last == 0
                last = vr % 10
                vr //= 10
                vp //= 10
                vm //= 10
                removed += 1
        if vr_tz and last == 5 and vr % 2 == 0:
            last = 4
        out = vr + (1 if ((vr == vm and (not accept or not vm_tz)) or last >= 5) else 0)
    else:
        round_up = False
        while vp // 10 > vm // 10:
            round_up = vr % 10 >= 5
            vr //= 10
            vp //= 10
            vm //= 10
            removed += 1
        out = vr + (1 if (vr == vm or round_up) else 0)
    return out, e10 + removed


def fmt_lexical(x: float) -> str:
    """lexical-core's default float text (what arrow-json's LineDelimitedWriter emits for a finite f64):
    shortest digits; positional with at least ".0" when -5 <= sci_exp <= 9, else d.ddde±x (mantissa keeps ".0")."""
    bits = struct.unpack("<Q", struct.pack("<d", x))[0]
    neg = bits >> 63
    bits &= (1 << 63) - 1
    if bits == 0:
        return "-0.0" if neg else "0.0"
    digits, e10 = d2d(bits)
    ds = str(digits)
    sci = e10 + len(ds) - 1
    if -5 <= sci <= 9:
        if e10 >= 0:
            s = ds + "0" * e10 + ".0"
        elif -e10 < len(ds):
            s = ds[: len(ds) + e10] + "." + ds[len(ds) + e10:]
        else:
            s = "0." + "0" * (-e10 - len(ds)) + ds
    else:
        s = ds[0] + "." + (ds[1:] if len(ds) > 1 else "0") + "e" + str(sci)
    return ("-" if neg else "") + s


def self_check(n=200000):
    rnd = random.Random(1)
    cases = [1.0, 10.0, 0.1, 0.3, 1e21, 1e22, 1e23, 5e-324, 1.7976931348623157e308, 2.2250738585072014e-308, 123456789.125,
             20.272727272727273, 28.8, 9007199254740993.0, 4.35, 0.000001, 1e-5, 12345678901.0, 1234567890.0]
    for _ in range(n):
        cases.append(struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(64)))[0])
        cases.append(rnd.random() * 10 ** rnd.randint(-20, 20))
    bad = 0
    for x in cases:
        if x != x or x in (float("inf"), float("-inf")) or x == 0:
            continue
        bits = struct.unpack("<Q", struct.pack("<d", abs(x)))[0]
        digits, e10 = d2d(bits)
        want = repr(abs(x))
        # python repr is the shortest round-trip digit string: compare digit strings and exponents
        mant, _, ex = want.partition("e")
        ip, _, fp = mant.partition(".")
        wd = (ip + fp).lstrip("0")
        we = (int(ex) if ex else 0) - len(fp)
        wd2 = wd.rstrip("0")
        we += len(wd) - len(wd2)
        if str(digits) != wd2 or e10 != we:
            bad += 1
            if bad < 5:
                print("MISMATCH", x, digits, e10, wd2, we)
        assert float(fmt_lexical(x)) == x, (x, fmt_lexical(x))
    return bad, len(cases)


if __name__ == "__main__":
    bad, total = self_check()
    print(f"self-check: {bad} mismatches of {total}")
    assert bad == 0
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "arkflow_b200", "csrc", "ryu_tables.h")
    with open(out, "w") as f:
        f.write("// generated by scripts/gen_ryu_tables.py - do not edit.  128-bit powers of five for Ryu (low word, high word).\n#pragma once\n")
        f.write(f"#define ARK_RYU_POW5_INV_BITCOUNT {POW5_INV_BITCOUNT}\n#define ARK_RYU_POW5_BITCOUNT {POW5_BITCOUNT}\n")
        for name, tab in (("kRyuPow5Inv", POW5_INV), ("kRyuPow5", POW5)):
            f.write(f"static __device__ const unsigned long long {name}[{len(tab)}][2] = {{\n")
            for v in tab:
                f.write("  {0x%016Xull, 0x%016Xull},\n" % (v & M64, v >> 64))
            f.write("};\n")
    print("wrote", out)
    for x in (10.0, 20.272727272727273, 1e21, 1e-7, 0.5, 123456789012.0, 1234567890.0):
        print(x, "->", fmt_lexical(x))
