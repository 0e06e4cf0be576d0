"""Coefficients for the hot-loop exponential (csrc/gpmpc_common.cuh: exp2s / exp2s_x4).

2^(f/T) - 1  ~=  f (c1 + f (c2 + f c3))   on |f| <= 1/2,   T = 2048   (f = t2 - rint(t2), t2 = x * T / ln 2)

The polynomial has no constant term (the error is pinned to 0 at f = 0), so a plain Remez exchange is singular; the
truncation error is dominated by the EVEN quartic term k^4 f^4 / 24 (k = ln2 / T), which only c2 can absorb:
minimising max |d f^2 - q f^4| on [0, F] gives d = (sqrt(8) - 2) q F^2 (error 0.1716 q F^4, 5.8x below Taylor).
c1, c3 stay Taylor (the odd remainder k^5 f^5 / 120 is < 1e-21).  Prints the float64 constants and the verified
maximum relative error of the rounded polynomial in 60-digit arithmetic."""
import mpmath as mp

mp.mp.dps = 60
T = mp.mpf(2048)
k = mp.log(2) / T
F = mp.mpf(1) / 2
q = k ** 4 / 24
c1, c2, c3 = k, k * k / 2 + (mp.sqrt(8) - 2) * q * F * F, k ** 3 / 6
cd = [float(c1), float(c2), float(c3)]
for i, x in enumerate(cd):
    print("c%d = %.20e  (%s)" % (i + 1, x, x.hex()))
worst = mp.mpf(0)
worst_taylor = mp.mpf(0)
for i in range(20001):
    f = -F + mp.mpf(i) / 20000
    exact = mp.power(2, f / T)
    p = f * (mp.mpf(cd[0]) + f * (mp.mpf(cd[1]) + f * mp.mpf(cd[2])))
    pt = f * (k + f * (k * k / 2 + f * k ** 3 / 6))
    worst = max(worst, abs((1 + p) / exact - 1))
    worst_taylor = max(worst_taylor, abs((1 + pt) / exact - 1))
print("max relative truncation error: %.3e  (Taylor: %.3e; half ulp = 1.11e-16)" % (float(worst), float(worst_taylor)))
print("SCALE = T/ln2 = %.20e" % float(T / mp.log(2)))
print("ln2/T        = %.20e" % float(k))


def bank_private(T, deg):
    """exp2b (csrc/gpmpc_common.cuh): T-entry table, one private copy per shared-memory bank pair, polynomial of degree
    `deg` without constant term:  2^(f/T) - 1 ~= f (c1 + f (c2 + ... f c_deg)),  |f| <= 1/2.
    The remainder is dominated by q f^(deg+1), q = k^(deg+1) / (deg+1)!, which c_(deg-1) absorbs: with m = deg - 1,
    minimise max |f^m (d - q f^2)| on [0, F] -- the interior extremum (2 d / (m + 2)) f*^m at f*^2 = m d / ((m + 2) q)
    equals minus the value at F."""
    T = mp.mpf(T)
    k = mp.log(2) / T
    F = mp.mpf(1) / 2
    m = deg - 1
    q = k ** (deg + 1) / mp.factorial(deg + 1)
    d = mp.findroot(lambda d: (2 * d / (m + 2)) * (m * d / ((m + 2) * q)) ** (mp.mpf(m) / 2) + F ** m * (d - q * F * F),
                    q * F * F * mp.mpf("0.8"))
    c = [k ** (i + 1) / mp.factorial(i + 1) for i in range(deg)]
    c[m - 1] += d
    cd = [float(x) for x in c]
    for i, x in enumerate(cd):
        print("b%d = %.20e  (%s)" % (i + 1, x, x.hex()))
    worst = worst_t = mp.mpf(0)
    for i in range(20001):
        f = -F + mp.mpf(i) / 20000
        exact = mp.power(2, f / T)
        p = pt = mp.mpf(0)
        for j in reversed(range(deg)):
            p = f * (mp.mpf(cd[j]) + p)
            pt = f * (k ** (j + 1) / mp.factorial(j + 1) + pt)
        worst = max(worst, abs((1 + p) / exact - 1))
        worst_t = max(worst_t, abs((1 + pt) / exact - 1))
    print("degree %d, T = %d: max relative truncation error %.3e  (Taylor: %.3e)" % (deg, int(T), float(worst), float(worst_t)))
    print("SCALE = T/ln2 = %.20e" % float(T / mp.log(2)))


bank_private(256, 4)
bank_private(128, 5)
