# per-level cycle counts of the lane-parallel verifier on one proof, grouped by the level's shape (terms per lane, subtractions,
# combine): the measured cost model behind verify_vm.hpp level_cost()
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
os.environ.setdefault("RLN_B200_WINDOW_BITS", "8")
import zerokit_b200 as z
from pyref import groth16 as G
g = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'reference_kats.json')))["groth16_verifier_single"]
proof = ((int(g["pi_a"][0]), int(g["pi_a"][1])), ((int(g["pi_b"][0][0]), int(g["pi_b"][0][1])), (int(g["pi_b"][1][0]), int(g["pi_b"][1][1]))),
         (int(g["pi_c"][0]), int(g["pi_c"][1])))
rec = G.rln_proof_to_bytes_le(proof, {k: int(g[k]) for k in ("root", "x", "external_nullifier", "y", "nullifier")})
rln = z.RLN.new(20)
for _ in range(3):
    ok, cyc, meta = rln.verify_vm_trace(rec)
assert ok == 1
tot = sum(cyc)
print('levels', len(cyc), 'cycles', tot, 'ms at 1.965 GHz', round(tot / 1.965e6, 3))
groups = collections.defaultdict(list)
for c, m in zip(cyc, meta):
    ns = [(m >> (4 * w)) & 15 for w in range(4)]
    key = (max(ns), (m >> 16) & 3, (m >> 18) & 1, m >> 20)
    groups[key].append(c)
print('maxN nsub comb special : levels  mean  min  max  share')
for k in sorted(groups):
    v = groups[k]
    print(k, len(v), round(sum(v) / len(v)), min(v), max(v), f'{100 * sum(v) / tot:.1f}%')
# by number of active warps at N = 1 (chains) and at the wide levels
aw = collections.defaultdict(list)
for c, m in zip(cyc, meta):
    ns = [(m >> (4 * w)) & 15 for w in range(4)]
    aw[(max(ns), sum(1 for x in ns if x))].append(c)
print('maxN active_warps : levels mean')
for k in sorted(aw):
    print(k, len(aw[k]), round(sum(aw[k]) / len(aw[k])))
json.dump(dict(cycles=cyc, meta=meta), open(os.path.join(ROOT, 'gpurun_out', 'vm_trace.json'), 'w'))
