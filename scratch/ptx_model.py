"""Python model of the PTX carry flag (CC.CF) used to validate carry-chain schedules before they are written
as inline PTX in zerokit_b200/csrc/fp.cuh.  Mirrors mul_ptx / sqr_ptx instruction for instruction."""
import random

M = 0xFFFFFFFF


class CC:
    def __init__(self):
        self.cf = None  # None = undefined (reading it is a bug)

    def _rd(self):
        assert self.cf is not None, "carry flag read while undefined"
        return self.cf

    def add_cc(self, a, b):
        s = a + b; self.cf = s >> 32; return s & M

    def addc_cc(self, a, b):
        s = a + b + self._rd(); self.cf = s >> 32; return s & M

    def addc(self, a, b):
        s = a + b + self._rd(); self.cf = None; return s & M

    def mad_lo_cc(self, a, b, c):
        s = ((a * b) & M) + c; self.cf = s >> 32; return s & M

    def madc_lo_cc(self, a, b, c):
        s = ((a * b) & M) + c + self._rd(); self.cf = s >> 32; return s & M

    def madc_hi_cc(self, a, b, c):
        s = ((a * b) >> 32) + c + self._rd(); self.cf = s >> 32; return s & M

    def madc_hi(self, a, b, c):
        s = ((a * b) >> 32) + c + self._rd(); assert s >> 32 == 0, "madc.hi dropped a carry"; self.cf = None; return s & M


def words(v, n=8):
    return [(v >> (32 * i)) & M for i in range(n)]


def val(w):
    return sum(x << (32 * i) for i, x in enumerate(w))


def reduce_row(c, E, O, p, inv):
    m = (E[0] * inv) & M
    O[0] = c.mad_lo_cc(p[1], m, O[0]); O[1] = c.madc_hi_cc(p[1], m, O[1])
    for j in (2, 4, 6):
        O[j] = c.madc_lo_cc(p[j + 1], m, O[j]); O[j + 1] = c.madc_hi_cc(p[j + 1], m, O[j + 1])
    assert c.cf == 0, "O chain overflow"
    E[0] = c.mad_lo_cc(p[0], m, E[0]); E[1] = c.madc_hi_cc(p[0], m, E[1])
    for j in (2, 4, 6):
        E[j] = c.madc_lo_cc(p[j], m, E[j]); E[j + 1] = c.madc_hi_cc(p[j], m, E[j + 1])
    s = O[7] + c._rd(); assert s >> 32 == 0; O[7] = s; c.cf = None
    assert E[0] == 0


def finish(c, E, O, P):
    r = [0] * 8
    r[0] = c.add_cc(E[1], O[0])
    for k in range(1, 7):
        r[k] = c.addc_cc(E[k + 1], O[k])
    r[7] = c.addc(O[7], 0)
    v = val(r)
    assert v < 2 * P
    return v - P if v >= P else v


def mul_model(a, b, P, inv):
    c = CC(); A, B, p = words(a), words(b), words(P)
    E = [0] * 8; O = [0] * 8
    for i in range(8):
        bi = B[i]
        if i == 0:
            for j in (0, 2, 4, 6):
                t = A[j] * bi; E[j], E[j + 1] = t & M, t >> 32
                t = A[j + 1] * bi; O[j], O[j + 1] = t & M, t >> 32
        else:
            nE = [0] * 8; nO = [0] * 8
            nE[0] = c.add_cc(O[0], E[1])
            for j in (0, 2, 4):
                nO[j] = c.madc_lo_cc(A[j + 1], bi, E[j + 2]); nO[j + 1] = c.madc_hi_cc(A[j + 1], bi, E[j + 3])
            nO[6] = c.madc_lo_cc(A[7], bi, 0); nO[7] = c.madc_hi(A[7], bi, 0)
            nE[0] = c.mad_lo_cc(A[0], bi, nE[0]); nE[1] = c.madc_hi_cc(A[0], bi, O[1])
            for j in (2, 4, 6):
                nE[j] = c.madc_lo_cc(A[j], bi, O[j]); nE[j + 1] = c.madc_hi_cc(A[j], bi, O[j + 1])
            s = nO[7] + c._rd(); assert s >> 32 == 0; nO[7] = s; c.cf = None
            E, O = nE, nO
        reduce_row(c, E, O, p, inv)
    return finish(c, E, O, P)


def sqr_model(a, P, inv):
    c = CC(); A, p = words(a), words(P)
    d = [0] * 8; s1 = [0] * 8
    for j in range(1, 8):
        d[j] = ((A[j] << 1) | (A[j - 1] >> 31)) & M
        s1[j] = (A[j] << 1) & M

    E = [0] * 8; O = [0] * 8
    for i in range(8):
        bi = A[i]

        def V(j):
            assert j >= i
            return A[i] if j == i else (s1[j] if j == i + 1 else d[j])
        if i == 0:
            for j in (0, 2, 4, 6):
                t = V(j) * bi; E[j], E[j + 1] = t & M, t >> 32
                t = V(j + 1) * bi; O[j], O[j + 1] = t & M, t >> 32
        else:
            nE = [0] * 8; nO = [0] * 8
            nE[0] = c.add_cc(O[0], E[1])
            for j in (0, 2, 4):
                if j + 1 >= i:
                    nO[j] = c.madc_lo_cc(V(j + 1), bi, E[j + 2]); nO[j + 1] = c.madc_hi_cc(V(j + 1), bi, E[j + 3])
                else:
                    nO[j] = c.addc_cc(E[j + 2], 0); nO[j + 1] = c.addc_cc(E[j + 3], 0)
            nO[6] = c.madc_lo_cc(V(7), bi, 0); nO[7] = c.madc_hi(V(7), bi, 0)
            started = False
            for j in (0, 2, 4, 6):
                if j >= i:
                    if not started:
                        nE[j] = c.mad_lo_cc(V(j), bi, O[j]); started = True
                    else:
                        nE[j] = c.madc_lo_cc(V(j), bi, O[j])
                    nE[j + 1] = c.madc_hi_cc(V(j), bi, O[j + 1])
                else:
                    if j > 0:
                        nE[j] = O[j]
                    nE[j + 1] = O[j + 1]
            if started:
                s = nO[7] + c._rd(); assert s >> 32 == 0; nO[7] = s; c.cf = None
            E, O = nE, nO
        reduce_row(c, E, O, p, inv)
    return finish(c, E, O, P)


def inplace_row(c, E, O, A, bi):
    """E/O += A * bi without shifting (same shape as the reduction row)"""
    O[0] = c.mad_lo_cc(A[1], bi, O[0]); O[1] = c.madc_hi_cc(A[1], bi, O[1])
    for j in (2, 4, 6):
        O[j] = c.madc_lo_cc(A[j + 1], bi, O[j]); O[j + 1] = c.madc_hi_cc(A[j + 1], bi, O[j + 1])
    assert c.cf == 0, "O chain overflow (row)"
    E[0] = c.mad_lo_cc(A[0], bi, E[0]); E[1] = c.madc_hi_cc(A[0], bi, E[1])
    for j in (2, 4, 6):
        E[j] = c.madc_lo_cc(A[j], bi, E[j]); E[j + 1] = c.madc_hi_cc(A[j], bi, E[j + 1])
    s = O[7] + c._rd(); assert s >> 32 == 0; O[7] = s; c.cf = None


def dot2_model(a, b, cc_, d, P, inv):
    """a*b + cc_*d with one interleaved reduction"""
    c = CC(); A, B, C2, D, p = words(a), words(b), words(cc_), words(d), words(P)
    E = [0] * 8; O = [0] * 8
    for i in range(8):
        bi = B[i]
        if i == 0:
            for j in (0, 2, 4, 6):
                t = A[j] * bi; E[j], E[j + 1] = t & M, t >> 32
                t = A[j + 1] * bi; O[j], O[j + 1] = t & M, t >> 32
        else:
            nE = [0] * 8; nO = [0] * 8
            nE[0] = c.add_cc(O[0], E[1])
            for j in (0, 2, 4):
                nO[j] = c.madc_lo_cc(A[j + 1], bi, E[j + 2]); nO[j + 1] = c.madc_hi_cc(A[j + 1], bi, E[j + 3])
            nO[6] = c.madc_lo_cc(A[7], bi, 0); nO[7] = c.madc_hi(A[7], bi, 0)
            nE[0] = c.mad_lo_cc(A[0], bi, nE[0]); nE[1] = c.madc_hi_cc(A[0], bi, O[1])
            for j in (2, 4, 6):
                nE[j] = c.madc_lo_cc(A[j], bi, O[j]); nE[j + 1] = c.madc_hi_cc(A[j], bi, O[j + 1])
            s = nO[7] + c._rd(); assert s >> 32 == 0; nO[7] = s; c.cf = None
            E, O = nE, nO
        inplace_row(c, E, O, C2, D[i])
        reduce_row(c, E, O, p, inv)
    return finish(c, E, O, P)


if __name__ == "__main__":
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    rnd = random.Random(1)
    for P in (R, Q):
        inv = (-pow(P, -1, 1 << 32)) % (1 << 32)
        Rinv = pow(1 << 256, -1, P)
        cases = [0, 1, P - 1, P - 2, (1 << 253), (1 << 254) - 1 if (1 << 254) - 1 < P else P - 3, 0x80000000 * sum(1 << (32 * i) for i in range(7))]
        cases += [rnd.randrange(P) for _ in range(3000)]
        # words with top bits set in every limb (exercise the doubling carries)
        cases += [val([rnd.choice([M, 0x80000000, 0x7fffffff, rnd.randrange(1 << 32)]) for _ in range(7)] + [rnd.randrange(0x30000000)]) for _ in range(2000)]
        for a in cases:
            a %= P
            b = rnd.randrange(P)
            assert mul_model(a, b, P, inv) == a * b * Rinv % P
            assert sqr_model(a, P, inv) == a * a * Rinv % P, hex(a)
            c2 = rnd.choice([P, P - 1, 0, rnd.randrange(P)]); d2 = rnd.choice([P - 1, rnd.randrange(P)])
            assert dot2_model(a, b, c2, d2, P, inv) == (a * b + c2 * d2) * Rinv % P
    for P in (R, Q):
        inv = (-pow(P, -1, 1 << 32)) % (1 << 32); Rinv = pow(1 << 256, -1, P)
        for a, b, c2, d2 in ((P - 1, P - 1, P, P - 1), (P - 1, P - 1, P - 1, P - 1), (P, P, P, P)):
            assert dot2_model(a, b, c2, d2, P, inv) == (a * b + c2 * d2) * Rinv % P
    print("mul/sqr/dot2 carry-chain schedules OK")


# ---- lazy-reduction Fq2 arithmetic (round 2, G2 accumulate): unreduced 512-bit products, one reduction per output coordinate ----
class CCs(CC):
    def sub_cc(self, a, b):
        s = a - b; self.cf = 1 if s < 0 else 0; return s & M      # cf models the BORROW here (sub.cc/subc.cc pair up consistently)

    def subc_cc(self, a, b):
        s = a - b - self._rd(); self.cf = 1 if s < 0 else 0; return s & M

    def subc(self, a, b):
        s = a - b - self._rd(); self.cf = None; return s & M


def shift_row(c, E, O, A, bi):
    """T = (T >> 32) + A*bi; E[0] of the old T is dropped by the caller's contract (it has been emitted or cancelled)"""
    nE = [0] * 8; nO = [0] * 8
    nE[0] = c.add_cc(O[0], E[1])
    for j in (0, 2, 4):
        nO[j] = c.madc_lo_cc(A[j + 1], bi, E[j + 2]); nO[j + 1] = c.madc_hi_cc(A[j + 1], bi, E[j + 3])
    nO[6] = c.madc_lo_cc(A[7], bi, 0); nO[7] = c.madc_hi(A[7], bi, 0)
    nE[0] = c.mad_lo_cc(A[0], bi, nE[0]); nE[1] = c.madc_hi_cc(A[0], bi, O[1])
    for j in (2, 4, 6):
        nE[j] = c.madc_lo_cc(A[j], bi, O[j]); nE[j + 1] = c.madc_hi_cc(A[j], bi, O[j + 1])
    s = nO[7] + c._rd(); assert s >> 32 == 0; nO[7] = s; c.cf = None
    return nE, nO


def mul_wide_model(a, b):
    """16-word product of two integers below 2^256 (mul_wide_ptx): the rows of mul_model without the reduction rows, the low
    word of every row is emitted instead of cancelled"""
    c = CC(); A, B = words(a), words(b)
    E = [0] * 8; O = [0] * 8; out = [0] * 16
    for j in (0, 2, 4, 6):
        t = A[j] * B[0]; E[j], E[j + 1] = t & M, t >> 32
        t = A[j + 1] * B[0]; O[j], O[j + 1] = t & M, t >> 32
    out[0] = E[0]
    for i in range(1, 8):
        E, O = shift_row(c, E, O, A, B[i])
        out[i] = E[0]
    out[8] = c.add_cc(E[1], O[0])
    for k in range(1, 7):
        out[8 + k] = c.addc_cc(E[k + 1], O[k])
    out[15] = c.addc(O[7], 0)
    return val(out)


def redc_wide_model(w, P, inv):
    """Montgomery reduction of a 16-word integer w < P·2^256 (redc_wide_ptx): reduction rows on the low half, the words of the
    high half enter at the top of the shifted accumulator one per row, the last one in the closing addition"""
    c = CC(); W = words(w, 16); p = words(P)
    E = W[:8]; O = [0] * 8
    for i in range(8):
        reduce_row(c, E, O, p, inv)
        if i == 7:
            break
        nE = [0] * 8; nO = [0] * 8
        nE[0] = c.add_cc(O[0], E[1])
        for j in range(6):
            nO[j] = c.addc_cc(E[j + 2], 0)
        nO[6] = c.addc_cc(W[8 + i], 0)
        nO[7] = c.addc(0, 0)
        for j in range(1, 8):
            nE[j] = O[j]
        E, O = nE, nO
    r = [0] * 8
    r[0] = c.add_cc(E[1], O[0])
    for k in range(1, 7):
        r[k] = c.addc_cc(E[k + 1], O[k])
    s = O[7] + W[15] + c._rd(); assert s >> 32 == 0, "redc_wide: result does not fit 256 bits"; r[7] = s; c.cf = None
    v = val(r)
    assert v < 2 * P, "redc_wide: more than one subtraction needed"
    return v - P if v >= P else v


def wide_add(x, y):
    return (x + y) & ((1 << 512) - 1)


def wide_sub(x, y):
    return (x - y) & ((1 << 512) - 1)


def fq2_mul_lazy_model(a0, a1, b0, b1, P, inv):
    OFF = P << 255
    P0 = mul_wide_model(a0, b0); P1 = mul_wide_model(a1, b1)
    sa = a0 + a1; sb = b0 + b1
    assert sa < 1 << 256 and sb < 1 << 256
    X = mul_wide_model(sa, sb)
    D = wide_add(wide_sub(P0, P1), OFF)
    X = wide_sub(wide_sub(X, P0), P1)
    return redc_wide_model(D, P, inv), redc_wide_model(X, P, inv)


def fq2_sub_prod_lazy_model(x, y, z, w, P, inv):
    """x·y − z·w over Fq2 = Fq[u]/(u² + 1), every argument a pair"""
    OFF = P << 255
    re = OFF; im = OFF
    t = mul_wide_model(x[0], y[0]); re = wide_add(re, t); im = wide_sub(im, t)
    t = mul_wide_model(x[1], y[1]); re = wide_sub(re, t); im = wide_sub(im, t)
    t = mul_wide_model(x[0] + x[1], y[0] + y[1]); im = wide_add(im, t)
    t = mul_wide_model(z[0], w[0]); re = wide_sub(re, t); im = wide_add(im, t)
    t = mul_wide_model(z[1], w[1]); re = wide_add(re, t); im = wide_add(im, t)
    t = mul_wide_model(z[0] + z[1], w[0] + w[1]); im = wide_sub(im, t)
    return redc_wide_model(re, P, inv), redc_wide_model(im, P, inv)


def check_lazy():
    Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    rnd = random.Random(7)
    inv = (-pow(Q, -1, 1 << 32)) % (1 << 32); Rinv = pow(1 << 256, -1, Q)
    edge = [0, 1, Q - 1, Q - 2, (1 << 253), Q - 3, 0xFFFFFFFF, (1 << 224) - 1]
    for it in range(6000):
        pick = (lambda: rnd.choice(edge)) if it < 1500 else (lambda: rnd.randrange(Q))
        a0, a1, b0, b1 = pick(), pick(), pick(), pick()
        assert mul_wide_model(a0 + a1, b0 + b1) == (a0 + a1) * (b0 + b1)
        c0, c1 = fq2_mul_lazy_model(a0, a1, b0, b1, Q, inv)
        assert c0 == (a0 * b0 - a1 * b1) * Rinv % Q and c1 == (a0 * b1 + a1 * b0) * Rinv % Q
        x, y, z, w = (pick(), pick()), (pick(), pick()), (pick(), pick()), (pick(), pick())
        re, im = fq2_sub_prod_lazy_model(x, y, z, w, Q, inv)
        assert re == (x[0] * y[0] - x[1] * y[1] - z[0] * w[0] + z[1] * w[1]) * Rinv % Q
        assert im == (x[0] * y[1] + x[1] * y[0] - z[0] * w[1] - z[1] * w[0]) * Rinv % Q
    # extremes of the bound: everything at q − 1 in the direction that maximises / minimises the wide sums
    m = Q - 1
    for x, y, z, w in (((m, 0), (m, 0), (0, m), (0, m)), ((0, m), (0, m), (m, 0), (m, 0)), ((m, m), (m, m), (m, m), (m, m)),
                       ((m, m), (m, m), (0, 0), (0, 0)), ((0, 0), (0, 0), (m, m), (m, m))):
        re, im = fq2_sub_prod_lazy_model(x, y, z, w, Q, inv)
        assert re == (x[0] * y[0] - x[1] * y[1] - z[0] * w[0] + z[1] * w[1]) * Rinv % Q
        assert im == (x[0] * y[1] + x[1] * y[0] - z[0] * w[1] - z[1] * w[0]) * Rinv % Q
    print("lazy Fq2 product / difference of products: carry-chain schedules OK")


if __name__ == "__main__":
    check_lazy()


# ---- one-level Karatsuba on the 8 x 8 limb product + Montgomery reduction with the shift fused into the reduction rows -------------
def mul4_wide_model(c, x, y):
    """8-word product of two 4-word integers, two interleaved carry chains (16 wide MADs)"""
    E = [0] * 4; O = [0] * 4; out = [0] * 8
    for j in (0, 2):
        t = x[j] * y[0]; E[j], E[j + 1] = t & M, t >> 32
        t = x[j + 1] * y[0]; O[j], O[j + 1] = t & M, t >> 32
    out[0] = E[0]
    for i in range(1, 4):
        bi = y[i]
        nE = [0] * 4; nO = [0] * 4
        nE[0] = c.add_cc(O[0], E[1])
        nO[0] = c.madc_lo_cc(x[1], bi, E[2]); nO[1] = c.madc_hi_cc(x[1], bi, E[3])
        nO[2] = c.madc_lo_cc(x[3], bi, 0); nO[3] = c.madc_hi(x[3], bi, 0)
        nE[0] = c.mad_lo_cc(x[0], bi, nE[0]); nE[1] = c.madc_hi_cc(x[0], bi, O[1])
        nE[2] = c.madc_lo_cc(x[2], bi, O[2]); nE[3] = c.madc_hi_cc(x[2], bi, O[3])
        s = nO[3] + c._rd(); assert s >> 32 == 0; nO[3] = s; c.cf = None
        E, O = nE, nO
        out[i] = E[0]
    out[4] = c.add_cc(E[1], O[0]); out[5] = c.addc_cc(E[2], O[1]); out[6] = c.addc_cc(E[3], O[2]); out[7] = c.addc(O[3], 0)
    return out


def kara_wide_model(a, b):
    """16-word product by one level of subtractive Karatsuba: a0·b0, a1·b1, |a0 − a1|·|b1 − b0| (48 wide MADs)"""
    c = CCs(); A, B = words(a), words(b)
    a0, a1, b0, b1 = A[:4], A[4:], B[:4], B[4:]
    z0 = mul4_wide_model(c, a0, b0); z2 = mul4_wide_model(c, a1, b1)

    def absdiff(x, y):
        d = [0] * 4
        d[0] = c.sub_cc(x[0], y[0])
        for i in (1, 2, 3):
            d[i] = c.subc_cc(x[i], y[i])
        neg = c.subc(0, 0) & 1                       # borrow
        mk = (0 - neg) & M
        d[0] = c.add_cc(d[0] ^ mk, neg)
        for i in (1, 2):
            d[i] = c.addc_cc(d[i] ^ mk, 0)
        d[3] = c.addc(d[3] ^ mk, 0)
        return d, neg
    da, sa = absdiff(a0, a1); db, sb = absdiff(b1, b0)
    assert val(da) == abs(val(a0) - val(a1)) and val(db) == abs(val(b1) - val(b0))
    z1 = mul4_wide_model(c, da, db)
    s = sa ^ sb; mz = (0 - s) & M
    m = [0] * 9
    m[0] = c.add_cc(z0[0], z2[0])
    for i in range(1, 8):
        m[i] = c.addc_cc(z0[i], z2[i])
    m[8] = c.addc(0, 0)
    c.add_cc(s, M)                                   # CF = s
    for i in range(8):
        m[i] = c.addc_cc(m[i], z1[i] ^ mz)
    m[8] = c.addc(m[8], mz)
    assert m[8] <= 1
    w = [0] * 16
    w[0:4] = z0[0:4]
    w[4] = c.add_cc(z0[4], m[0])
    for i in (1, 2, 3):
        w[4 + i] = c.addc_cc(z0[4 + i], m[i])
    for i in range(4):
        w[8 + i] = c.addc_cc(z2[i], m[4 + i])
    w[12] = c.addc_cc(z2[4], m[8]); w[13] = c.addc_cc(z2[5], 0); w[14] = c.addc_cc(z2[6], 0); w[15] = c.addc(z2[7], 0)
    return val(w)


def redc_fused_model(w, P, inv):
    """Montgomery reduction of a 16-word integer: row 0 reduces in place, rows 1..7 shift and reduce in one pass (the shift's
    carry ripple rides in the MAD chains), the high words enter at the top one per row"""
    c = CC(); W = words(w, 16); p = words(P)
    E = W[:8]; O = [0] * 8
    reduce_row(c, E, O, p, inv)
    for i in range(1, 8):
        m = (((O[0] + E[1]) & M) * inv) & M
        nE = [0] * 8; nO = [0] * 8
        nE[0] = c.add_cc(O[0], E[1])
        for j in (0, 2, 4):
            nO[j] = c.madc_lo_cc(p[j + 1], m, E[j + 2]); nO[j + 1] = c.madc_hi_cc(p[j + 1], m, E[j + 3])
        nO[6] = c.madc_lo_cc(p[7], m, W[8 + i - 1]); nO[7] = c.madc_hi(p[7], m, 0)
        nE[0] = c.mad_lo_cc(p[0], m, nE[0]); nE[1] = c.madc_hi_cc(p[0], m, O[1])
        for j in (2, 4, 6):
            nE[j] = c.madc_lo_cc(p[j], m, O[j]); nE[j + 1] = c.madc_hi_cc(p[j], m, O[j + 1])
        s = nO[7] + c._rd(); assert s >> 32 == 0; nO[7] = s; c.cf = None
        assert nE[0] == 0
        E, O = nE, nO
    r = [0] * 8
    r[0] = c.add_cc(E[1], O[0])
    for k in range(1, 7):
        r[k] = c.addc_cc(E[k + 1], O[k])
    s = O[7] + W[15] + c._rd(); assert s >> 32 == 0, "redc_fused: result does not fit 256 bits"; r[7] = s; c.cf = None
    v = val(r)
    assert v < 2 * P, "redc_fused: more than one subtraction needed"
    return v - P if v >= P else v


def check_kara():
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    rnd = random.Random(11)
    for P in (R, Q):
        inv = (-pow(P, -1, 1 << 32)) % (1 << 32); Rinv = pow(1 << 256, -1, P)
        edge = [0, 1, P - 1, P - 2, P, (1 << 253), (1 << 128) - 1, 1 << 128, (1 << 128) + 1, ((1 << 128) - 1) << 96, 0xFFFFFFFF,
                sum(0xFFFFFFFF << (64 * i) for i in range(4)) % P, sum(0x80000000 << (32 * i) for i in range(7))]
        for it in range(8000):
            if it < 2500:
                a, b = rnd.choice(edge), rnd.choice(edge)
            elif it < 4000:
                a = val([rnd.choice([M, 0, 0x80000000, rnd.randrange(1 << 32)]) for _ in range(7)] + [rnd.randrange(0x30000000)])
                b = val([rnd.choice([M, 0, 1, rnd.randrange(1 << 32)]) for _ in range(7)] + [rnd.randrange(0x30000000)])
            else:
                a, b = rnd.randrange(P), rnd.randrange(P)
            assert kara_wide_model(a, b) == a * b, (hex(a), hex(b))
            assert redc_fused_model(a * b, P, inv) == a * b * Rinv % P
            c2, d2 = rnd.choice([P, P - 1, 0, rnd.randrange(P)]), rnd.choice([P - 1, rnd.randrange(P)])
            assert redc_fused_model(a * b + c2 * d2, P, inv) == (a * b + c2 * d2) * Rinv % P
        for n_terms in (3, 4):        # up to 4 products of operands at their maximum (p): 4p² < 0.76·p·2^256
            assert redc_fused_model(n_terms * P * P, P, inv) == 0
    print("Karatsuba product + fused-shift reduction: carry-chain schedules OK")


if __name__ == "__main__":
    check_kara()
