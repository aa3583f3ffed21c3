from fractions import Fraction
import struct, math
# correctly rounded 2^(i/32) via high-precision integer arithmetic
def rn_pow2(i, N=32):
    # compute 2^(i/N) with 200 bits using integer nth root: 2^(i/N) = (2^i)^(1/N)
    prec = 300
    val = (1 << (i + N * prec))  # (2^(i/N) * 2^prec)^N = 2^i * 2^(N*prec)
    # integer N-th root
    lo, hi = 1 << prec, 1 << (prec + 1)
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if mid ** N <= val: lo = mid
        else: hi = mid - 1
    x = Fraction(lo, 1 << prec)  # floor approx with 300 bits
    # round to double
    f = float(x)  # Fraction -> float is correctly rounded
    return f
tab = []
for i in range(32):
    bits = struct.unpack("<Q", struct.pack("<d", rn_pow2(i)))[0]
    tab.append((bits - (i << 47)) & 0xFFFFFFFFFFFFFFFF)
print(",".join("0x%016xULL" % t for t in tab))
