"""ORACLE (test infrastructure, not product code).

BN254 field / curve / pairing arithmetic on Python integers.  Slow and obvious on
purpose: this is the restatement that gets pinned against the reference's golden
vectors (tests/test_oracle_goldens.py); the C++ oracle and the CUDA product are
then compared against it.

Reference call sites whose arithmetic lives in un-vendored crates (ark-ff /
ark-bn254 / ark-ec 0.5.0, Cargo.lock:60-234): constants are the public BN254
parameters, also spelled out in rln/src/circuit/iden3calc/graph.rs:14-15 (Fr
modulus).
"""

# scalar field (Fr) and base field (Fq) moduli
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583

ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE = 63
TWO_ADICITY = 28
FR_GENERATOR = 5


def inv(a, m):
    return pow(a, -1, m)


# ----------------------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_neg(a):
    return ((-a[0]) % Q, (-a[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_sqr(a):
    return f2_mul(a, a)


def f2_inv(a):
    d = inv((a[0] * a[0] + a[1] * a[1]) % Q, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


def f2_scalar(a, k):
    return (a[0] * k % Q, a[1] * k % Q)


F2_ZERO = (0, 0)
F2_ONE = (1, 0)

# G2 curve coefficient b' = 3/(9+u)
B2 = f2_mul((3, 0), f2_inv((9, 1)))

G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


# ----------------------------------------------------------------------------- generic affine curve ops
class Ops:
    """Field-op bundle so G1 (ints mod Q) and G2 (Fq2 tuples) share one group law."""

    def __init__(self, add, sub, mul, inv_, zero, b):
        self.add, self.sub, self.mul, self.inv, self.zero, self.b = add, sub, mul, inv_, zero, b


OPS1 = Ops(lambda a, b: (a + b) % Q, lambda a, b: (a - b) % Q, lambda a, b: a * b % Q,
           lambda a: inv(a, Q), 0, 3)
OPS2 = Ops(f2_add, f2_sub, f2_mul, f2_inv, F2_ZERO, B2)

INF = None  # point at infinity


def pt_neg(o, p):
    if p is INF:
        return INF
    return (p[0], o.sub(o.zero, p[1]))


def pt_double(o, p):
    if p is INF:
        return INF
    x, y = p
    if y == o.zero:
        return INF
    xx = o.mul(x, x)
    lam = o.mul(o.add(o.add(xx, xx), xx), o.inv(o.add(y, y)))
    x3 = o.sub(o.sub(o.mul(lam, lam), x), x)
    y3 = o.sub(o.mul(lam, o.sub(x, x3)), y)
    return (x3, y3)


def pt_add(o, p, q):
    if p is INF:
        return q
    if q is INF:
        return p
    if p[0] == q[0]:
        if p[1] == q[1]:
            return pt_double(o, p)
        return INF
    lam = o.mul(o.sub(q[1], p[1]), o.inv(o.sub(q[0], p[0])))
    x3 = o.sub(o.sub(o.mul(lam, lam), p[0]), q[0])
    y3 = o.sub(o.mul(lam, o.sub(p[0], x3)), p[1])
    return (x3, y3)


def pt_mul(o, p, k):
    k %= R
    acc = INF
    while k:
        if k & 1:
            acc = pt_add(o, acc, p)
        p = pt_double(o, p)
        k >>= 1
    return acc


def on_curve(o, p):
    if p is INF:
        return True
    x, y = p
    return o.mul(y, y) == o.add(o.mul(o.mul(x, x), x), o.b)


# Jacobian arithmetic for the (slow-but-not-hopeless) python MSM -----------------
def _jac_add_affine(o, P, q):
    """P jacobian (X,Y,Z) or None, q affine (never INF)."""
    if P is None:
        one = 1 if o is OPS1 else F2_ONE
        return (q[0], q[1], one)
    X1, Y1, Z1 = P
    Z1Z1 = o.mul(Z1, Z1)
    U2 = o.mul(q[0], Z1Z1)
    S2 = o.mul(o.mul(q[1], Z1), Z1Z1)
    H = o.sub(U2, X1)
    r = o.sub(S2, Y1)
    if H == o.zero:
        if r == o.zero:
            return _jac_double(o, P)
        return None
    HH = o.mul(H, H)
    HHH = o.mul(H, HH)
    V = o.mul(X1, HH)
    X3 = o.sub(o.sub(o.mul(r, r), HHH), o.add(V, V))
    Y3 = o.sub(o.mul(r, o.sub(V, X3)), o.mul(Y1, HHH))
    Z3 = o.mul(Z1, H)
    return (X3, Y3, Z3)


def _jac_double(o, P):
    if P is None:
        return None
    X, Y, Z = P
    if Y == o.zero:
        return None
    A = o.mul(X, X)
    B = o.mul(Y, Y)
    C = o.mul(B, B)
    t = o.add(X, B)
    D = o.sub(o.sub(o.mul(t, t), A), C)
    D = o.add(D, D)
    E = o.add(o.add(A, A), A)
    F = o.mul(E, E)
    X3 = o.sub(F, o.add(D, D))
    C8 = o.add(C, C)
    C8 = o.add(C8, C8)
    C8 = o.add(C8, C8)
    Y3 = o.sub(o.mul(E, o.sub(D, X3)), C8)
    Z3 = o.mul(o.add(Y, Y), Z)
    return (X3, Y3, Z3)


def _jac_to_affine(o, P):
    if P is None:
        return INF
    X, Y, Z = P
    zi = o.inv(Z)
    zi2 = o.mul(zi, zi)
    return (o.mul(X, zi2), o.mul(Y, o.mul(zi2, zi)))


def _jac_add(o, P, Qj):
    if P is None:
        return Qj
    if Qj is None:
        return P
    return pt_to_jac_add_general(o, P, Qj)


def pt_to_jac_add_general(o, P, S):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = S
    Z1Z1 = o.mul(Z1, Z1)
    Z2Z2 = o.mul(Z2, Z2)
    U1 = o.mul(X1, Z2Z2)
    U2 = o.mul(X2, Z1Z1)
    S1 = o.mul(o.mul(Y1, Z2), Z2Z2)
    S2 = o.mul(o.mul(Y2, Z1), Z1Z1)
    H = o.sub(U2, U1)
    r = o.sub(S2, S1)
    if H == o.zero:
        if r == o.zero:
            return _jac_double(o, P)
        return None
    HH = o.mul(H, H)
    HHH = o.mul(H, HH)
    V = o.mul(U1, HH)
    X3 = o.sub(o.sub(o.mul(r, r), HHH), o.add(V, V))
    Y3 = o.sub(o.mul(r, o.sub(V, X3)), o.mul(S1, HHH))
    Z3 = o.mul(o.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def msm(o, points, scalars, c=8):
    """Σ scalars[i]·points[i] (affine result).  Plain windowed bucket method; the result is a
    unique group element so the algorithm choice does not affect parity
    (rln/src/partial_proof.rs:98-104 → ark-ec msm_bigint)."""
    assert len(points) == len(scalars)
    pairs = [(p, s % R) for p, s in zip(points, scalars) if p is not INF and s % R]
    if not pairs:
        return INF
    nwin = (254 + c - 1) // c
    total = None
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            total = _jac_double(o, total)
        buckets = [None] * (1 << c)
        sh = w * c
        for p, s in pairs:
            d = (s >> sh) & ((1 << c) - 1)
            if d:
                buckets[d] = _jac_add_affine(o, buckets[d], p)
        run = None
        acc = None
        for d in range((1 << c) - 1, 0, -1):
            run = _jac_add(o, run, buckets[d])
            acc = _jac_add(o, acc, run)
        total = _jac_add(o, total, acc)
    return _jac_to_affine(o, total)


# ----------------------------------------------------------------------------- Fq12 = Fq[w]/(w^12 - 18 w^6 + 82)
def f12_mul(a, b):
    t = [0] * 23
    for i in range(12):
        ai = a[i]
        if ai:
            for j in range(12):
                t[i + j] += ai * b[j]
    for i in range(22, 11, -1):
        v = t[i]
        if v:
            t[i - 6] += 18 * v
            t[i - 12] -= 82 * v
    return [x % Q for x in t[:12]]


F12_ONE = [1] + [0] * 11


def f12_pow(a, e):
    out = F12_ONE
    while e:
        if e & 1:
            out = f12_mul(out, a)
        a = f12_mul(a, a)
        e >>= 1
    return out


def f12_inv(a):
    # a^(q^12 - 2); only used a handful of times by the oracle verifier
    return f12_pow(a, Q ** 12 - 2)


def f12_sub(a, b):
    return [(x - y) % Q for x, y in zip(a, b)]


def f12_add(a, b):
    return [(x + y) % Q for x, y in zip(a, b)]


def f12_from_int(k):
    return [k % Q] + [0] * 11


W2 = [0, 0, 1] + [0] * 9
W3 = [0, 0, 0, 1] + [0] * 8


def twist(pt):
    """G2 affine over Fq2 → curve over Fq12 (y^2 = x^3 + 3)."""
    (x0, x1), (y0, y1) = pt
    nx = [0] * 12
    ny = [0] * 12
    nx[0], nx[6] = (x0 - 9 * x1) % Q, x1
    ny[0], ny[6] = (y0 - 9 * y1) % Q, y1
    return (f12_mul(nx, W2), f12_mul(ny, W3))


def f2_conj(a):
    return (a[0], (-a[1]) % Q)


def f2_pow(a, e):
    out = F2_ONE
    while e:
        if e & 1:
            out = f2_mul(out, a)
        a = f2_mul(a, a)
        e >>= 1
    return out


XI = (9, 1)
GAMMA2 = f2_pow(XI, (Q - 1) // 3)   # w^(2(q-1))
GAMMA3 = f2_pow(XI, (Q - 1) // 2)   # w^(3(q-1))


def frob_twist(pt):
    """q-power Frobenius of ψ(pt), expressed back in twist (Fq2) coordinates."""
    return (f2_mul(f2_conj(pt[0]), GAMMA2), f2_mul(f2_conj(pt[1]), GAMMA3))


def embed(a, shift):
    """Fq2 element a0 + a1·u (u = w^6 − 9) times w^shift, as an Fq12 coefficient list."""
    out = [0] * 12
    out[shift] = (a[0] - 9 * a[1]) % Q
    out[shift + 6] = a[1]
    return out


def _line(lam, r, p1):
    """Line through ψ(r) with twist-slope lam, evaluated at p1 ∈ G1:
    l = λ·w·(xP − xR·w²) − (yP − yR·w³)."""
    l = embed(f2_scalar(lam, p1[0]), 1)                 # λ w xP
    l = f12_sub(l, embed(f2_mul(lam, r[0]), 3))         # − λ xR w³
    l = f12_add(l, embed(r[1], 3))                      # + yR w³
    l[0] = (l[0] - p1[1]) % Q                           # − yP
    return l


def miller_loop(q2, p1):
    """Optimal-ate Miller loop (no final exponentiation). q2 ∈ G2 affine, p1 ∈ G1 affine."""
    if q2 is INF or p1 is INF:
        return F12_ONE
    o = OPS2
    Rp = q2
    f = F12_ONE

    def dbl_step(f, Rp):
        x, y = Rp
        xx = f2_mul(x, x)
        lam = f2_mul(f2_add(f2_add(xx, xx), xx), f2_inv(f2_add(y, y)))
        f = f12_mul(f12_mul(f, f), _line(lam, Rp, p1))
        return f, pt_double(o, Rp)

    def add_step(f, Rp, S):
        lam = f2_mul(f2_sub(S[1], Rp[1]), f2_inv(f2_sub(S[0], Rp[0])))
        f = f12_mul(f, _line(lam, Rp, p1))
        return f, pt_add(o, Rp, S)

    for i in range(LOG_ATE, -1, -1):
        f, Rp = dbl_step(f, Rp)
        if ATE_LOOP_COUNT & (1 << i):
            f, Rp = add_step(f, Rp, q2)
    Q1 = frob_twist(q2)
    nQ2 = pt_neg(o, frob_twist(Q1))
    f, Rp = add_step(f, Rp, Q1)
    f, Rp = add_step(f, Rp, nQ2)
    return f


def final_exp(f):
    return f12_pow(f, (Q ** 12 - 1) // R)


def pairing_product_is_one(pairs):
    """Π e(P_i, Q_i) == 1 with one shared final exponentiation."""
    f = F12_ONE
    for p1, q2 in pairs:
        f = f12_mul(f, miller_loop(q2, p1))
    return final_exp(f) == F12_ONE
