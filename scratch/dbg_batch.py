import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
c = sys.argv[1]; Bs = [int(x) for x in sys.argv[2:]]
os.environ['RLN_B200_WINDOW_BITS'] = c
import zerokit_b200 as z
from common import *
from oracle import cref_binding as C
from pyref import groth16 as G, poseidon as P
D = 20
rln = z.RLN.new(D)
ctx = C.Ctx(resource(D, 'rln_final.arkzkey'), resource(D, 'graph.bin'))
fs = fr_stream(9)
N = max(Bs)
pe = [P.poseidon([i + 7]) for i in range(D)]
recs, rs, oin = [], [], []
for j in range(N):
    s, x = next(fs), next(fs)
    ix = [(j >> i) & 1 for i in range(D)]
    recs.append(witness_le(s, 100, j % 100, pe, ix, x, 777)); rs += [next(fs), next(fs)]
    if j < 16 or j >= N - 8: oin.append((j, ctx.inputs_buffer(s, 100, j % 100, pe, ix, x, 777)))
rsb = fr_bytes(rs)
chk = [j for j, _ in oin]
want = {}
for j, ib in oin:
    p, pub = ctx.prove_batch(ib, rsb[64 * j:64 * j + 64], 1)
    v = ints(p); proof = ((v[0], v[1]), ((v[2], v[3]), (v[4], v[5])), (v[6], v[7]))
    want[j] = G.proof_to_bytes(proof)
for B in Bs:
    out = rln.prove_batch(b''.join(recs[:B]), B, rsb[:64 * B])
    bad = []
    for j in chk:
        if j >= B: continue
        got = out[290 * j + 1:290 * j + 129]
        if got != want[j]:
            bad.append((j, 'A' if got[:32] != want[j][:32] else '', 'B' if got[32:96] != want[j][32:96] else '', 'C' if got[96:] != want[j][96:] else ''))
    ok = rln.verify_batch(out, B)
    print('c', c, 'B', B, 'mismatch', bad, 'verify_fail', sum(1 for v in ok if v != 1), rln.last_stage_ms(), flush=True)
