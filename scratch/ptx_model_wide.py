"""Carry-flag model of fp.cuh dot_wide (the pairing VM's sum of up to 8 products with ONE Montgomery reduction, operands a[k]
anywhere below 2^256): the rows of mul_ptx with a ninth accumulator word X (T = E + O·2^32 + X·2^288), because
T < (Σ a[k] + p)·2^32 no longer fits 2^288 once Σ a[k] + p ≥ 2^256.  Instruction for instruction as written in fp.cuh."""
import random
from ptx_model import CC, M, words, val


def row_first(A, bi):
    E = [0] * 8; O = [0] * 8
    for j in (0, 2, 4, 6):
        t = A[j] * bi; E[j], E[j + 1] = t & M, t >> 32
        t = A[j + 1] * bi; O[j], O[j + 1] = t & M, t >> 32
    return E, O, 0


def row_inplace_w(c, E, O, X, A, bi):
    O[0] = c.mad_lo_cc(A[1], bi, O[0]); O[1] = c.madc_hi_cc(A[1], bi, O[1])
    for j in (2, 4, 6):
        O[j] = c.madc_lo_cc(A[j + 1], bi, O[j]); O[j + 1] = c.madc_hi_cc(A[j + 1], bi, O[j + 1])
    X = c.addc(X, 0)
    E[0] = c.mad_lo_cc(A[0], bi, E[0]); E[1] = c.madc_hi_cc(A[0], bi, E[1])
    for j in (2, 4, 6):
        E[j] = c.madc_lo_cc(A[j], bi, E[j]); E[j + 1] = c.madc_hi_cc(A[j], bi, E[j + 1])
    O[7] = c.addc_cc(O[7], 0)
    X = c.addc(X, 0)
    assert X < (1 << 32)
    return X


def row_shift_w(c, E, O, X, A, bi):
    nE = [0] * 8; nO = [0] * 8
    nE[0] = c.add_cc(O[0], E[1])
    for j in (0, 2, 4):
        nO[j] = c.madc_lo_cc(A[j + 1], bi, E[j + 2]); nO[j + 1] = c.madc_hi_cc(A[j + 1], bi, E[j + 3])
    nO[6] = c.madc_lo_cc(A[7], bi, 0); nO[7] = c.madc_hi_cc(A[7], bi, X)
    nX = c.addc(0, 0)
    nE[0] = c.mad_lo_cc(A[0], bi, nE[0]); nE[1] = c.madc_hi_cc(A[0], bi, O[1])
    for j in (2, 4, 6):
        nE[j] = c.madc_lo_cc(A[j], bi, O[j]); nE[j + 1] = c.madc_hi_cc(A[j], bi, O[j + 1])
    nO[7] = c.addc_cc(nO[7], 0)
    nX = c.addc(nX, 0)
    return nE, nO, nX


def dot_wide_model(As, Bs, P, inv):
    c = CC(); p = words(P)
    Aw = [words(a) for a in As]; Bw = [words(b) for b in Bs]
    E = O = None; X = 0
    for i in range(8):
        if i == 0:
            E, O, X = row_first(Aw[0], Bw[0][0])
        else:
            assert E[0] == 0
            E, O, X = row_shift_w(c, E, O, X, Aw[0], Bw[0][i])
        for k in range(1, len(As)):
            X = row_inplace_w(c, E, O, X, Aw[k], Bw[k][i])
        m = (E[0] * inv) & M
        X = row_inplace_w(c, E, O, X, p, m)          # the reduction row has the shape of an in-place row with a = p, b = m
        assert E[0] == 0
    # finish: (E >> 32) + O, X must be spent
    assert X == 0, "final value does not fit 256 bits"
    r = [0] * 8
    r[0] = c.add_cc(E[1], O[0])
    for k in range(1, 7):
        r[k] = c.addc_cc(E[k + 1], O[k])
    s = O[7] + c._rd(); assert s >> 32 == 0; r[7] = s
    return val(r)


if __name__ == "__main__":
    Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    inv = (-pow(Q, -1, 1 << 32)) % (1 << 32); Rinv = pow(1 << 256, -1, Q)
    rnd = random.Random(3)
    FULL = (1 << 256) - 1
    worst_b = val([M] * 7 + [0x30644e72])   # > q as an integer but the largest word pattern a reduced value can approach
    tests = 0
    for N in range(1, 9):
        for trial in range(400):
            kind = trial % 4
            if kind == 0:   # the extreme: every a = 4·(q) or all ones, b with saturated words
                As = [rnd.choice([FULL, 4 * Q, 4 * Q - 4]) for _ in range(N)]
                Bs = [min(worst_b, Q - 1) if rnd.random() < 0.5 else val([M] * 7 + [0x30644e71]) for _ in range(N)]
            elif kind == 1:
                As = [rnd.choice([Q, Q - 1, 2 * Q, 4 * Q, 0, 1]) for _ in range(N)]
                Bs = [rnd.choice([Q - 1, 0, 1, rnd.randrange(Q)]) for _ in range(N)]
            else:
                As = [rnd.randrange(Q) << rnd.choice([0, 0, 1, 2]) for _ in range(N)]
                Bs = [rnd.randrange(Q) for _ in range(N)]
            W = sum(a / Q for a in As)
            if 0.18903 * W + 1 >= 5.28:   # the caller's contract: the result must fit 256 bits
                continue
            got = dot_wide_model(As, Bs, Q, inv)
            want = sum(a * b for a, b in zip(As, Bs)) * Rinv % Q
            assert got % Q == want, (N, trial)
            assert got < (0.18903 * W + 1) * Q + 1
            tests += 1
    print("dot_wide carry-chain schedule OK:", tests, "cases")
