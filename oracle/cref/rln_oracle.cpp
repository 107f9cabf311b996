// ORACLE (test infrastructure, not product code).
//
// C++ CPU restatement of the reference's RLN proving path, exported as a C API for ctypes.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  It is validated against oracle/pyref + tests/golden (which are
// pinned to the reference's golden vectors) by tests/test_oracle_cref.py.
//
// Follows (reference file:line):
//   utils/src/poseidon/poseidon_constants.rs:15-263, poseidon_hash.rs:63-135, rln/src/hashers.rs:14-23
//   utils/src/merkle_tree/full_merkle_tree.rs:82-115,197-223,288-304,360-399
//   rln/src/circuit/mod.rs:256-305 (arkzkey), iden3calc/storage.rs:265-302 + proto.rs (graph.bin)
//   rln/src/circuit/iden3calc/graph.rs:71-143,246-272,314-466 (node evaluation)
//   rln/src/circuit/qap.rs:30-98 (CircomReduction::witness_map_from_matrices)
//   rln/src/partial_proof.rs:98-104,182-274 (msm + proof assembly == ark-groth16 0.5.0)
//   rln/src/protocol/proof.rs:856-894 (verifier public-input order)
// Un-vendored algorithms restated from the pinned crates' published behaviour: ark-ec 0.5.0
// msm_bigint (window rule c = 3 if n<32 else ceil(log2 n)*69/100+2, signed digits), ark-poly
// 0.5.0 radix-2 domain (generator 5, two-adicity 28), ark-groth16 0.5.0 verify_proof.
#include "field.hpp"

#include <malloc.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <string>
#include <chrono>
#include <thread>

#include "pairing_consts.inc"

static Fq fq_from_limbs(const u64* l) { return Fq::from_u256({{l[0], l[1], l[2], l[3]}}); }

const PairingConsts& pairing_consts() {
    static PairingConsts pc = [] {
        PairingConsts c;
        c.gamma2 = {fq_from_limbs(K_GAMMA2[0]), fq_from_limbs(K_GAMMA2[1])};
        c.gamma3 = {fq_from_limbs(K_GAMMA3[0]), fq_from_limbs(K_GAMMA3[1])};
        for (int k = 0; k < 6; k++) c.frob2[k] = fq_from_limbs(K_FROB2[k]);
        c.hard.assign(K_HARD, K_HARD + 12);
        return c;
    }();
    return pc;
}

// =============================================================================== Poseidon
struct GrainLFSR {  // poseidon_constants.rs:15-205
    bool st[80];
    int head = 0;
    GrainLFSR(int nbits, int t, int rf, int rp) {
        memset(st, 0, sizeof st);
        st[1] = true;  // field
        auto put = [&](int lo, int hi, u64 v) {
            for (int i = hi; i >= lo; i--) { st[i] = v & 1; v >>= 1; }
        };
        put(6, 17, nbits);
        put(18, 29, t);
        put(30, 39, rf);
        put(40, 49, rp);
        for (int i = 50; i < 80; i++) st[i] = true;
        for (int i = 0; i < 160; i++) update();
    }
    bool update() {
        bool nb = st[(head + 62) % 80] ^ st[(head + 51) % 80] ^ st[(head + 38) % 80] ^ st[(head + 23) % 80] ^
                  st[(head + 13) % 80] ^ st[head];
        st[head] = nb;
        head = (head + 1) % 80;
        return nb;
    }
    U256 get254() {  // first produced bit is the most significant
        U256 v = {{0, 0, 0, 0}};
        for (int k = 0; k < 254; k++) {
            bool b = update();
            while (!b) { update(); b = update(); }
            bool bit = update();
            // v = (v << 1) | bit
            for (int i = 3; i > 0; i--) v.l[i] = (v.l[i] << 1) | (v.l[i - 1] >> 63);
            v.l[0] = (v.l[0] << 1) | (u64)bit;
        }
        return v;
    }
};

struct PoseidonParams {
    int t, rf, rp;
    std::vector<Fr> ark;
    std::vector<Fr> mds;  // row-major t×t
};
static const int ROUND_PARAMS[8][3] = {{2, 8, 56}, {3, 8, 57}, {4, 8, 56}, {5, 8, 60},
                                       {6, 8, 60}, {7, 8, 63}, {8, 8, 64}, {9, 8, 63}};  // hashers.rs:14-23

static PoseidonParams make_params(int t, int rf, int rp) {  // poseidon_constants.rs:207-263 (skip_matrices = 0)
    GrainLFSR l(254, t, rf, rp);
    PoseidonParams p{t, rf, rp, {}, {}};
    while ((int)p.ark.size() < (rf + rp) * t) {
        U256 v = l.get254();
        if (u256_cmp(v, FrParams::MOD) < 0) p.ark.push_back(Fr::from_u256(v));
    }
    auto modp = [&]() {
        U256 v = l.get254();
        while (u256_cmp(v, FrParams::MOD) >= 0) u256_sub(v, v, FrParams::MOD);
        return Fr::from_u256(v);
    };
    std::vector<Fr> xs(t), ys(t);
    for (auto& x : xs) x = modp();
    for (auto& y : ys) y = modp();
    p.mds.resize(t * t);
    for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) p.mds[i * t + j] = (xs[i] + ys[j]).inv();
    return p;
}
static const PoseidonParams& poseidon_params(int t) {
    static PoseidonParams tab[10];
    static std::atomic<int> ready[10];
    static std::atomic_flag lock = ATOMIC_FLAG_INIT;
    if (t < 2 || t > 9) abort();
    if (!ready[t].load(std::memory_order_acquire)) {
        while (lock.test_and_set(std::memory_order_acquire)) {}
        if (!ready[t].load()) {
            tab[t] = make_params(t, ROUND_PARAMS[t - 2][1], ROUND_PARAMS[t - 2][2]);
            ready[t].store(1, std::memory_order_release);
        }
        lock.clear(std::memory_order_release);
    }
    return tab[t];
}

static Fr poseidon(const Fr* in, int n) {  // poseidon_hash.rs:97-135
    const int t = n + 1;
    const PoseidonParams& p = poseidon_params(t);
    Fr st[9], nx[9];
    st[0] = Fr::zero();
    for (int i = 0; i < n; i++) st[i + 1] = in[i];
    for (int r = 0; r < p.rf + p.rp; r++) {
        for (int k = 0; k < t; k++) st[k] += p.ark[r * t + k];
        bool full = r < p.rf / 2 || r >= p.rf / 2 + p.rp;
        for (int k = 0; k < (full ? t : 1); k++) {
            Fr x2 = st[k].sqr();
            st[k] = x2.sqr() * st[k];
        }
        for (int i = 0; i < t; i++) {
            Fr acc = Fr::zero();
            for (int j = 0; j < t; j++) acc += p.mds[i * t + j] * st[j];
            nx[i] = acc;
        }
        for (int i = 0; i < t; i++) st[i] = nx[i];
    }
    return st[0];
}
static Fr poseidon2(const Fr& a, const Fr& b) {
    Fr in[2] = {a, b};
    return poseidon(in, 2);
}

template <class Fn>
static void parallel_for(size_t n, int nthreads, Fn fn) {
    if (nthreads <= 1 || n < 2) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    const size_t grain = std::max<size_t>(1, n / (nthreads * 8));
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&] {
            for (;;) {
                size_t s = next.fetch_add(grain);
                if (s >= n) break;
                size_t e = std::min(n, s + grain);
                for (size_t i = s; i < e; i++) fn(i);
            }
        });
    for (auto& x : th) x.join();
}

// =============================================================================== circuit data
struct SparseRow { std::vector<std::pair<Fr, uint32_t>> e; };
struct ZkeyData {
    G1A alpha_g1, beta_g1, delta_g1;
    G2A beta_g2, gamma_g2, delta_g2;
    std::vector<G1A> gamma_abc, a_query, b_g1, h_query, l_query;
    std::vector<G2A> b_g2;
    u64 num_instance, num_witness, num_constraints;
    std::vector<SparseRow> A, B;
};
struct Rd {
    const uint8_t* p;
    size_t n, o = 0;
    bool fail = false;
    bool need(size_t k) { if (o + k > n) { fail = true; return false; } return true; }
    u64 u64le() { if (!need(8)) return 0; u64 v; memcpy(&v, p + o, 8); o += 8; return v; }
    const uint8_t* take(size_t k) { if (!need(k)) return nullptr; const uint8_t* r = p + o; o += k; return r; }
};
static Fq rd_fq(const uint8_t* b, bool mask) {
    uint8_t t[32];
    memcpy(t, b, 32);
    if (mask) t[31] &= 0x3f;  // ark-serialize SWFlags live in the two top bits
    return Fq::from_le32(t);
}
static G1A rd_g1(Rd& r) {
    const uint8_t* b = r.take(64);
    if (!b) return {Fq::zero(), Fq::zero(), true};
    if (b[63] & 0x40) return {Fq::zero(), Fq::zero(), true};
    return {rd_fq(b, false), rd_fq(b + 32, true), false};
}
static G2A rd_g2(Rd& r) {
    const uint8_t* b = r.take(128);
    if (!b) return {Fq2::zero(), Fq2::zero(), true};
    if (b[127] & 0x40) return {Fq2::zero(), Fq2::zero(), true};
    return {{rd_fq(b, false), rd_fq(b + 32, false)}, {rd_fq(b + 64, false), rd_fq(b + 96, true)}, false};
}
static bool parse_zkey(const uint8_t* data, size_t n, ZkeyData& z) {  // circuit/mod.rs:256-305
    Rd r{data, n};
    z.alpha_g1 = rd_g1(r);
    z.beta_g2 = rd_g2(r);
    z.gamma_g2 = rd_g2(r);
    z.delta_g2 = rd_g2(r);
    auto vec1 = [&](std::vector<G1A>& v) { u64 k = r.u64le(); if (k > n) { r.fail = true; return; } v.resize(k); for (auto& p : v) p = rd_g1(r); };
    vec1(z.gamma_abc);
    z.beta_g1 = rd_g1(r);
    z.delta_g1 = rd_g1(r);
    vec1(z.a_query);
    vec1(z.b_g1);
    { u64 k = r.u64le(); if (k > n) return false; z.b_g2.resize(k); for (auto& p : z.b_g2) p = rd_g2(r); }
    vec1(z.h_query);
    vec1(z.l_query);
    z.num_instance = r.u64le();
    z.num_witness = r.u64le();
    z.num_constraints = r.u64le();
    r.u64le(); r.u64le(); r.u64le();
    auto mat = [&](std::vector<SparseRow>& m) {
        u64 rows = r.u64le();
        if (rows > n) { r.fail = true; return; }
        m.resize(rows);
        for (auto& row : m) {
            u64 k = r.u64le();
            if (k > n) { r.fail = true; return; }
            row.e.resize(k);
            for (auto& e : row.e) {
                const uint8_t* b = r.take(32);
                if (!b) return;
                e.first = Fr::from_le32(b);
                e.second = (uint32_t)r.u64le();
            }
        }
    };
    std::vector<SparseRow> C;
    mat(z.A); mat(z.B); mat(C);
    return !r.fail && r.o == n;
}

enum NodeKind : uint8_t { N_INPUT, N_CONST, N_UNO, N_DUO, N_TRES };
struct Node { NodeKind kind; uint8_t op; uint32_t a, b, c; };
struct GraphData {
    std::vector<Node> nodes;
    std::vector<Fr> consts;  // value for N_CONST nodes, indexed by node.a
    std::vector<uint32_t> signals;
    std::map<std::string, std::pair<uint32_t, uint32_t>> inputs;
    uint32_t inputs_size = 0;
};
static bool varint(const uint8_t* p, size_t n, size_t& o, u64& v) {
    v = 0;
    for (int s = 0; s < 70; s += 7) {
        if (o >= n) return false;
        uint8_t c = p[o++];
        v |= (u64)(c & 0x7f) << s;
        if (!(c & 0x80)) return true;
    }
    return false;
}
struct PbField { u64 tag, wt, val; const uint8_t* ptr; size_t len; };
static bool pb_next(const uint8_t* p, size_t n, size_t& o, PbField& f) {
    u64 key;
    if (!varint(p, n, o, key)) return false;
    f.tag = key >> 3; f.wt = key & 7; f.ptr = nullptr; f.len = 0; f.val = 0;
    if (f.wt == 0) return varint(p, n, o, f.val);
    if (f.wt == 2) { u64 l; if (!varint(p, n, o, l) || o + l > n) return false; f.ptr = p + o; f.len = l; o += l; return true; }
    if (f.wt == 5) { if (o + 4 > n) return false; o += 4; return true; }
    if (f.wt == 1) { if (o + 8 > n) return false; o += 8; return true; }
    return false;
}
static bool parse_graph(const uint8_t* d, size_t n, GraphData& g) {  // storage.rs:265-302
    static const char MAGIC[] = "wtns.graph.001";
    if (n < 14 + 8 || memcmp(d, MAGIC, 14)) return false;
    size_t o = 14;
    u64 cnt;
    memcpy(&cnt, d + o, 8);
    o += 8;
    if (cnt > n) return false;
    g.nodes.reserve(cnt);
    for (u64 i = 0; i < cnt; i++) {
        u64 len;
        if (!varint(d, n, o, len) || o + len > n) return false;
        const uint8_t* m = d + o;
        size_t mo = 0;
        o += len;
        PbField f;
        if (!pb_next(m, len, mo, f) || f.wt != 2) return false;
        Node nd{N_INPUT, 0, 0, 0, 0};
        u64 vals[5] = {0, 0, 0, 0, 0};
        const uint8_t* sub = nullptr;
        size_t sublen = 0;
        size_t bo = 0;
        PbField bf;
        while (bo < f.len) {
            if (!pb_next(f.ptr, f.len, bo, bf)) return false;
            if (bf.tag < 5 && bf.wt == 0) vals[bf.tag] = bf.val;
            if (bf.tag == 1 && bf.wt == 2) { sub = bf.ptr; sublen = bf.len; }
        }
        switch (f.tag) {
            case 1: nd.kind = N_INPUT; nd.a = (uint32_t)vals[1]; break;
            case 2: {
                nd.kind = N_CONST;
                const uint8_t* vb = nullptr; size_t vl = 0; size_t so = 0; PbField sf;
                while (sub && so < sublen) { if (!pb_next(sub, sublen, so, sf)) return false; if (sf.tag == 1 && sf.wt == 2) { vb = sf.ptr; vl = sf.len; } }
                nd.a = (uint32_t)g.consts.size();
                g.consts.push_back(Fr::from_le_bytes_mod(vb, vl));  // storage.rs:45-47
                break;
            }
            case 3: nd.kind = N_UNO; nd.op = (uint8_t)vals[1]; nd.a = (uint32_t)vals[2]; break;
            case 4: nd.kind = N_DUO; nd.op = (uint8_t)vals[1]; nd.a = (uint32_t)vals[2]; nd.b = (uint32_t)vals[3]; break;
            case 5: nd.kind = N_TRES; nd.op = (uint8_t)vals[1]; nd.a = (uint32_t)vals[2]; nd.b = (uint32_t)vals[3]; nd.c = (uint32_t)vals[4]; break;
            default: return false;
        }
        g.nodes.push_back(nd);
    }
    u64 mdlen;
    if (!varint(d, n, o, mdlen) || o + mdlen > n) return false;
    const uint8_t* md = d + o;
    size_t mo = 0;
    PbField f;
    while (mo < mdlen) {
        if (!pb_next(md, mdlen, mo, f)) return false;
        if (f.tag == 1 && f.wt == 2) {
            size_t po = 0; u64 v;
            while (po < f.len) { if (!varint(f.ptr, f.len, po, v)) return false; g.signals.push_back((uint32_t)v); }
        } else if (f.tag == 1 && f.wt == 0) {
            g.signals.push_back((uint32_t)f.val);
        } else if (f.tag == 2 && f.wt == 2) {
            size_t eo = 0; PbField ef; std::string key; u64 off = 0, ln = 0;
            while (eo < f.len) {
                if (!pb_next(f.ptr, f.len, eo, ef)) return false;
                if (ef.tag == 1 && ef.wt == 2) key.assign((const char*)ef.ptr, ef.len);
                if (ef.tag == 2 && ef.wt == 2) {
                    size_t so = 0; PbField sf;
                    while (so < ef.len) { if (!pb_next(ef.ptr, ef.len, so, sf)) return false; if (sf.tag == 1) off = sf.val; if (sf.tag == 2) ln = sf.val; }
                }
            }
            g.inputs[key] = {(uint32_t)off, (uint32_t)ln};
        }
    }
    // iden3calc.rs:106-121 get_inputs_size
    bool started = false; uint32_t mx = 0;
    for (auto& nd : g.nodes) {
        if (nd.kind == N_INPUT) { mx = std::max(mx, nd.a); started = true; }
        else if (started) break;
    }
    g.inputs_size = mx + 1;
    return true;
}

// graph.rs:71-143, 314-466 — only value-level semantics; values are canonical-equivalent Fr
static const U256 HALF_R = {{0xa1f0fac9f8000000ULL, 0x9419f4243cdcb848ULL, 0xdc2822db40c0ac2eULL, 0x183227397098d014ULL}};
static bool eval_duo(int op, const Fr& a, const Fr& b, Fr& out) {
    switch (op) {
        case 0: out = a * b; return true;                                   // Mul
        case 1: out = b.is_zero() ? Fr::zero() : a * b.inv(); return true;  // Div
        case 2: out = a + b; return true;                                   // Add
        case 3: out = a - b; return true;                                   // Sub
        case 4: out = a.pow(b.to_u256()); return true;                      // Pow
        case 7: out = (a == b) ? Fr::one() : Fr::zero(); return true;       // Eq
        case 8: out = (a == b) ? Fr::zero() : Fr::one(); return true;       // Neq
        case 13: out = (a.is_zero() || b.is_zero()) ? Fr::zero() : Fr::one(); return true;  // Land
        case 14: out = (a.is_zero() && b.is_zero()) ? Fr::zero() : Fr::one(); return true;  // Lor
        default: break;
    }
    U256 x = a.to_u256(), y = b.to_u256();
    if (op >= 9 && op <= 12) {  // Lt Gt Leq Geq with the "negative above r/2" convention (graph.rs:410-466)
        bool xn = u256_cmp(x, HALF_R) > 0, yn = u256_cmp(y, HALF_R) > 0;
        int c = u256_cmp(x, y);
        bool res;
        if (xn != yn) { bool lt = xn; res = (op == 9 || op == 11) ? lt : !lt; }
        else res = op == 9 ? c < 0 : op == 10 ? c > 0 : op == 11 ? c <= 0 : c >= 0;
        out = res ? Fr::one() : Fr::zero();
        return true;
    }
    if (op == 16 || op == 15) {  // Shr / Shl
        if (u256_is_zero(y)) { out = a; return true; }
        U256 lim = {{254, 0, 0, 0}};
        if (u256_cmp(y, lim) >= 0) { out = Fr::zero(); return true; }
        int s = (int)(y.l[0] & 0xff);
        U256 r = {{0, 0, 0, 0}};
        for (int i = 0; i < 256; i++) {
            int src = op == 16 ? i + s : i - s;
            if (src >= 0 && src < 256 && u256_bit(x, src)) r.l[i >> 6] |= 1ULL << (i & 63);
        }
        if (u256_cmp(r, FrParams::MOD) >= 0) return false;  // from_bigint failure
        out = Fr::from_u256(r);
        return true;
    }
    if (op >= 17 && op <= 19) {  // Bor Band Bxor (graph.rs:365-408)
        U256 r;
        for (int i = 0; i < 4; i++) r.l[i] = op == 17 ? (x.l[i] | y.l[i]) : op == 18 ? (x.l[i] & y.l[i]) : (x.l[i] ^ y.l[i]);
        if (u256_cmp(r, FrParams::MOD) > 0) u256_sub(r, r, FrParams::MOD);
        if (u256_cmp(r, FrParams::MOD) >= 0) return false;
        out = Fr::from_u256(r);
        return true;
    }
    return false;  // Idiv / Mod: not used by the bundled graphs; not restated
}

static bool evaluate_graph(const GraphData& g, const Fr* inputs, std::vector<Fr>& vals, Fr* out) {  // graph.rs:246-272
    vals.resize(g.nodes.size());
    for (size_t i = 0; i < g.nodes.size(); i++) {
        const Node& nd = g.nodes[i];
        switch (nd.kind) {
            case N_CONST: vals[i] = g.consts[nd.a]; break;
            case N_INPUT: vals[i] = inputs[nd.a]; break;
            case N_DUO: if (!eval_duo(nd.op, vals[nd.a], vals[nd.b], vals[i])) return false; break;
            case N_UNO: if (nd.op != 0) return false; vals[i] = vals[nd.a].neg(); break;
            case N_TRES: vals[i] = vals[nd.a].is_zero() ? vals[nd.c] : vals[nd.b]; break;
        }
    }
    for (size_t i = 0; i < g.signals.size(); i++) out[i] = vals[g.signals[i]];
    return true;
}

// =============================================================================== NTT / QAP
static Fr root_of_unity(u64 n) {  // ark-poly radix-2 domain: generator 5, two-adicity 28
    int lg = 0;
    while ((1ULL << lg) < n) lg++;
    U256 e;  // (r−1) >> 28
    U256 one = {{1, 0, 0, 0}};
    u256_sub(e, FrParams::MOD, one);
    for (int s = 0; s < 28; s++) {
        for (int i = 0; i < 3; i++) e.l[i] = (e.l[i] >> 1) | (e.l[i + 1] << 63);
        e.l[3] >>= 1;
    }
    Fr w = Fr::from_u64(5).pow(e);
    for (int i = 0; i < 28 - lg; i++) w = w.sqr();
    return w;
}
static void ntt_inplace(std::vector<Fr>& a, const Fr& w) {
    size_t n = a.size();
    int lg = 0;
    while ((1ULL << lg) < n) lg++;
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (int b = 0; b < lg; b++) j |= ((i >> b) & 1) << (lg - 1 - b);
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t m = 1; m < n; m <<= 1) {
        Fr wm = w;
        for (size_t k = n / (2 * m); k > 1; k >>= 1) wm = wm.sqr();
        static thread_local std::vector<Fr> tw;
        if (tw.size() < m) tw.resize(m);
        tw[0] = Fr::one();
        for (size_t j = 1; j < m; j++) tw[j] = tw[j - 1] * wm;
        for (size_t k = 0; k < n; k += 2 * m)
            for (size_t j = 0; j < m; j++) {
                Fr u = a[k + j], x = a[k + j + m] * tw[j];
                a[k + j] = u + x;
                a[k + j + m] = u - x;
            }
    }
}
static void intt_inplace(std::vector<Fr>& a, const Fr& w) {
    ntt_inplace(a, w.inv());
    Fr ni = Fr::from_u64(a.size()).inv();
    for (auto& x : a) x *= ni;
}

static void witness_map(const ZkeyData& z, const Fr* w, std::vector<Fr>& h) {  // qap.rs:30-98
    size_t n = 1;
    while (n < z.num_constraints + z.num_instance) n <<= 1;
    static thread_local std::vector<Fr> a, b, c;
    a.assign(n, Fr::zero());
    b.assign(n, Fr::zero());
    c.assign(n, Fr::zero());
    for (size_t i = 0; i < z.num_constraints; i++) {
        Fr sa = Fr::zero(), sb = Fr::zero();
        for (auto& e : z.A[i].e) sa += e.first * w[e.second];
        for (auto& e : z.B[i].e) sb += e.first * w[e.second];
        a[i] = sa; b[i] = sb;
    }
    for (size_t i = 0; i < z.num_instance; i++) a[z.num_constraints + i] = w[i];
    for (size_t i = 0; i < z.num_constraints; i++) c[i] = a[i] * b[i];
    Fr om = root_of_unity(n), g = root_of_unity(2 * n);
    auto to_coset = [&](std::vector<Fr>& v) {
        intt_inplace(v, om);
        Fr t = Fr::one();
        for (size_t i = 0; i < n; i++) { v[i] *= t; t *= g; }
        ntt_inplace(v, om);
    };
    to_coset(a); to_coset(b); to_coset(c);
    h.resize(n);
    for (size_t i = 0; i < n; i++) h[i] = a[i] * b[i] - c[i];
}

// =============================================================================== MSM (ark-ec 0.5 msm_bigint restated)
static int ark_window(size_t n) {
    if (n < 32) return 3;
    int lg = 0;
    while ((1ULL << lg) < n) lg++;  // ark_std::log2 = ceil
    return lg * 69 / 100 + 2;
}
template <class F>
static Jac<F> msm_pippenger(const Affine<F>* pts, const U256* sc, size_t n, int nthreads = 1, int c_override = 0) {
    if (n == 0) return Jac<F>::infinity();
    const int c = c_override ? c_override : ark_window(n);
    const int nwin = (254 + c - 1) / c + 1;  // +1 for the signed-digit carry
    // signed digits in [−2^(c−1), 2^(c−1)]; scratch is reused per worker thread (no allocator traffic
    // when many proofs run side by side)
    static thread_local std::vector<int32_t> digits_tl;
    if (digits_tl.size() < n * (size_t)nwin) digits_tl.resize(n * (size_t)nwin);
    int32_t* digits = digits_tl.data();
    for (size_t i = 0; i < n; i++) {
        int carry = 0;
        for (int w = 0; w < nwin; w++) {
            int bit = w * c;
            int64_t d = carry;
            if (bit < 256) {
                int li = bit >> 6, sh = bit & 63;
                u64 v = sc[i].l[li] >> sh;
                if (sh + c > 64 && li < 3) v |= sc[i].l[li + 1] << (64 - sh);
                d += (int64_t)(v & ((1ULL << c) - 1));
            }
            if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else carry = 0;
            digits[i * nwin + w] = (int32_t)d;
        }
    }
    std::vector<Jac<F>> winsum(nwin);
    parallel_for(nwin, nthreads, [&](size_t w) {
        static thread_local std::vector<Jac<F>> buckets_tl;
        buckets_tl.assign((size_t)1 << (c - 1), Jac<F>::infinity());
        std::vector<Jac<F>>& buckets = buckets_tl;
        for (size_t i = 0; i < n; i++) {
            int d = digits[i * nwin + w];
            if (d == 0 || pts[i].inf) continue;
            if (d > 0) buckets[d - 1] = buckets[d - 1].add_affine(pts[i]);
            else { Affine<F> np = {pts[i].x, pts[i].y.neg(), false}; buckets[-d - 1] = buckets[-d - 1].add_affine(np); }
        }
        Jac<F> run = Jac<F>::infinity(), acc = Jac<F>::infinity();
        for (int k = (1 << (c - 1)) - 1; k >= 0; k--) {
            run = run.add(buckets[k]);
            acc = acc.add(run);
        }
        winsum[w] = acc;
    });
    Jac<F> total = winsum[nwin - 1];
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) total = total.dbl();
        total = total.add(winsum[w]);
    }
    return total;
}

// =============================================================================== context + prove/verify
struct OracleCtx {
    ZkeyData z;
    GraphData g;
    uint32_t depth;
};

struct ProofOut { G1A a; G2A b; G1A c; };

static bool prove_one(const OracleCtx& ctx, const Fr* inputs /*inputs_size*/, const Fr& r, const Fr& s, ProofOut& out,
                      Fr* pub5, int nthreads) {
    const ZkeyData& z = ctx.z;
    size_t nw = ctx.g.signals.size();
    static thread_local std::vector<Fr> vals, w, h;
    static thread_local std::vector<U256> ws, hs;
    w.resize(nw);
    if (!evaluate_graph(ctx.g, inputs, vals, w.data())) return false;
    if (pub5) for (size_t i = 0; i + 1 < z.num_instance; i++) pub5[i] = w[1 + i];  // all public wires (5 single, 15 multi)
    witness_map(z, w.data(), h);
    ws.resize(nw);
    hs.resize(h.size());
    for (size_t i = 0; i < nw; i++) ws[i] = w[i].to_u256();
    for (size_t i = 0; i < h.size(); i++) hs[i] = h[i].to_u256();
    const size_t ni = z.num_instance;
    U256 ru = r.to_u256(), su = s.to_u256();
    // partial_proof.rs:226-267
    G1J a_acc = msm_pippenger<Fq>(z.a_query.data() + 1, ws.data() + 1, nw - 1, nthreads);
    G1J g_a = G1J::from_affine(z.alpha_g1).add_affine(z.a_query[0]).add(a_acc).add(G1J::from_affine(z.delta_g1).mul(ru));
    G1J g1_b = G1J::infinity();
    if (!r.is_zero()) {
        G1J b1 = msm_pippenger<Fq>(z.b_g1.data() + 1, ws.data() + 1, nw - 1, nthreads);
        g1_b = G1J::from_affine(z.beta_g1).add_affine(z.b_g1[0]).add(b1).add(G1J::from_affine(z.delta_g1).mul(su));
    }
    G2J b2 = msm_pippenger<Fq2>(z.b_g2.data() + 1, ws.data() + 1, nw - 1, nthreads);
    G2J g2_b = G2J::from_affine(z.beta_g2).add_affine(z.b_g2[0]).add(b2).add(G2J::from_affine(z.delta_g2).mul(su));
    G1J l_acc = msm_pippenger<Fq>(z.l_query.data(), ws.data() + ni, nw - ni, nthreads);
    G1J h_acc = msm_pippenger<Fq>(z.h_query.data(), hs.data(), std::min(hs.size(), z.h_query.size()), nthreads);
    U256 rs = (r * s).to_u256();
    G1J g_c = g_a.mul(su).add(g1_b.mul(ru)).add(G1J::from_affine(z.delta_g1).mul(rs).neg()).add(l_acc).add(h_acc);
    out.a = g_a.to_affine();
    out.b = g2_b.to_affine();
    out.c = g_c.to_affine();
    return true;
}

static bool verify_one(const ZkeyData& z, const ProofOut& p, const Fr* pub, size_t npub) {
    if (npub + 1 != z.gamma_abc.size()) return false;
    G1J vkx = G1J::from_affine(z.gamma_abc[0]);
    for (size_t i = 0; i < npub; i++) vkx = vkx.add(G1J::from_affine(z.gamma_abc[i + 1]).mul(pub[i].to_u256()));
    G1A na = {p.a.x, p.a.y.neg(), p.a.inf};
    Fq12 f = miller_loop(p.b, na) * miller_loop(z.beta_g2, z.alpha_g1) * miller_loop(z.gamma_g2, vkx.to_affine()) *
             miller_loop(z.delta_g2, p.c);
    return final_exponentiation(f) == Fq12::one();
}

// =============================================================================== C API (all field values: 32-byte LE canonical)
static void g1_out(const G1A& p, uint8_t* o) {  // x|y, all-zero + flag byte 0x40 at [63] for infinity
    memset(o, 0, 64);
    if (p.inf) { o[63] = 0x40; return; }
    p.x.to_le32(o); p.y.to_le32(o + 32);
}
static void g2_out(const G2A& p, uint8_t* o) {
    memset(o, 0, 128);
    if (p.inf) { o[127] = 0x40; return; }
    p.x.a.to_le32(o); p.x.b.to_le32(o + 32); p.y.a.to_le32(o + 64); p.y.b.to_le32(o + 96);
}
static G1A g1_in(const uint8_t* b) {
    if (b[63] & 0x40) return {Fq::zero(), Fq::zero(), true};
    return {Fq::from_le32(b), Fq::from_le32(b + 32), false};
}
static G2A g2_in(const uint8_t* b) {
    if (b[127] & 0x40) return {Fq2::zero(), Fq2::zero(), true};
    return {{Fq::from_le32(b), Fq::from_le32(b + 32)}, {Fq::from_le32(b + 64), Fq::from_le32(b + 96)}, false};
}

extern "C" {

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
// keep big blocks inside the per-thread arenas instead of mmap/munmap round trips (page-fault storms with >100 workers)
static int oracle_malloc_tuning = [] { mallopt(M_MMAP_THRESHOLD, 1 << 30); mallopt(M_TRIM_THRESHOLD, 1 << 30); mallopt(M_ARENA_MAX, 256); return 0; }();

// Poseidon over n inputs (1..8)
void orc_poseidon(const uint8_t* in, int n, uint8_t* out) {
    Fr v[8];
    for (int i = 0; i < n; i++) v[i] = Fr::from_le32(in + 32 * i);
    poseidon(v, n).to_le32(out);
}
// many independent pair hashes (bench helper): in = count × 64 B, out = count × 32 B
void orc_poseidon_pairs(const uint8_t* in, size_t count, uint8_t* out, int nthreads) {
    parallel_for(count, nthreads, [&](size_t i) {
        poseidon2(Fr::from_le32(in + 64 * i), Fr::from_le32(in + 64 * i + 32)).to_le32(out + 32 * i);
    });
}
// ark round constants / MDS for state width t (Montgomery-free canonical bytes)
int orc_poseidon_constants(int t, uint8_t* ark_out, uint8_t* mds_out) {
    const PoseidonParams& p = poseidon_params(t);
    for (size_t i = 0; i < p.ark.size(); i++) p.ark[i].to_le32(ark_out + 32 * i);
    for (size_t i = 0; i < p.mds.size(); i++) p.mds[i].to_le32(mds_out + 32 * i);
    return (int)p.ark.size();
}

// Dense Merkle tree (full_merkle_tree.rs): nodes_out holds 2^(depth+1)−1 values, heap order
// (root at 0, leaf i at 2^depth−1+i).  Leaves [start, start+count) are set over an otherwise
// default-leaf(0) tree.
void orc_merkle_build(uint32_t depth, const uint8_t* leaves, size_t start, size_t count, uint8_t* nodes_out, int nthreads) {
    size_t nleaf = (size_t)1 << depth, total = 2 * nleaf - 1;
    std::vector<Fr> nodes(total);
    std::vector<Fr> zeros(depth + 1);
    zeros[0] = Fr::zero();
    for (uint32_t k = 0; k < depth; k++) zeros[k + 1] = poseidon2(zeros[k], zeros[k]);
    for (uint32_t lvl = 0; lvl <= depth; lvl++) {
        size_t base = ((size_t)1 << lvl) - 1;
        for (size_t i = 0; i < ((size_t)1 << lvl); i++) nodes[base + i] = zeros[depth - lvl];
    }
    if (count) {
        size_t base = nleaf - 1;
        for (size_t i = 0; i < count; i++) nodes[base + start + i] = Fr::from_le32(leaves + 32 * i);
        size_t lo = base + start, hi = base + start + count - 1;
        while (lo > 0) {  // update_hashes: level by level over the touched range (:360-399)
            lo = (lo - 1) / 2; hi = (hi - 1) / 2;
            parallel_for(hi - lo + 1, nthreads, [&](size_t k) {
                size_t p = lo + k;
                nodes[p] = poseidon2(nodes[2 * p + 1], nodes[2 * p + 2]);
            });
        }
    }
    for (size_t i = 0; i < total; i++) nodes[i].to_le32(nodes_out + 32 * i);
}

void* orc_ctx_new(const uint8_t* zkey, size_t zlen, const uint8_t* graph, size_t glen) {
    OracleCtx* c = new OracleCtx();
    if (!parse_zkey(zkey, zlen, c->z) || !parse_graph(graph, glen, c->g)) { delete c; return nullptr; }
    auto it = c->g.inputs.find("pathElements");
    c->depth = it == c->g.inputs.end() ? 0 : it->second.second;
    return c;
}
void orc_ctx_free(void* p) { delete (OracleCtx*)p; }
uint32_t orc_ctx_depth(void* p) { return ((OracleCtx*)p)->depth; }
uint32_t orc_ctx_inputs_size(void* p) { return ((OracleCtx*)p)->g.inputs_size; }
uint32_t orc_ctx_num_public(void* p) { return (uint32_t)((OracleCtx*)p)->z.num_instance - 1; }
uint32_t orc_ctx_num_wires(void* p) { return (uint32_t)((OracleCtx*)p)->g.signals.size(); }
uint32_t orc_ctx_domain(void* p) {
    OracleCtx* c = (OracleCtx*)p; size_t n = 1;
    while (n < c->z.num_constraints + c->z.num_instance) n <<= 1;
    return (uint32_t)n;
}
// offset,len of a named input signal; returns 0 if absent
int orc_ctx_input(void* p, const char* name, uint32_t* off, uint32_t* len) {
    OracleCtx* c = (OracleCtx*)p;
    auto it = c->g.inputs.find(name);
    if (it == c->g.inputs.end()) return 0;
    *off = it->second.first; *len = it->second.second;
    return 1;
}

// inputs: inputs_size × 32 B (slot 0 must be 1).  w_out: num_wires × 32 B.
int orc_witness(void* p, const uint8_t* inputs, uint8_t* w_out) {
    OracleCtx* c = (OracleCtx*)p;
    std::vector<Fr> in(c->g.inputs_size), vals, w(c->g.signals.size());
    for (size_t i = 0; i < in.size(); i++) in[i] = Fr::from_le32(inputs + 32 * i);
    if (!evaluate_graph(c->g, in.data(), vals, w.data())) return 0;
    for (size_t i = 0; i < w.size(); i++) w[i].to_le32(w_out + 32 * i);
    return 1;
}
// h_out: domain × 32 B
void orc_qap_h(void* p, const uint8_t* w, uint8_t* h_out) {
    OracleCtx* c = (OracleCtx*)p;
    std::vector<Fr> wv(c->g.signals.size()), h;
    for (size_t i = 0; i < wv.size(); i++) wv[i] = Fr::from_le32(w + 32 * i);
    witness_map(c->z, wv.data(), h);
    for (size_t i = 0; i < h.size(); i++) h[i].to_le32(h_out + 32 * i);
}

// Batch prove.  inputs: n × inputs_size × 32 B; rs: n × 64 B (r|s); proofs_out: n × 256 B
// (A 64 | B 128 | C 64, affine canonical); pub_out: n × num_public × 32 B (the public wires w[1..], i.e.
// [y, root, nullifier, x, en] for the single circuit) or NULL.
// One worker thread per proof (the model rln/README.md:324-332 recommends); returns #failures.
int orc_prove_batch(void* p, size_t n, const uint8_t* inputs, const uint8_t* rs, uint8_t* proofs_out, uint8_t* pub_out,
                    int nthreads) {
    OracleCtx* c = (OracleCtx*)p;
    std::atomic<int> fails(0);
    const size_t isz = c->g.inputs_size;
    int inner = (n == 1) ? nthreads : 1;
    parallel_for(n, n == 1 ? 1 : nthreads, [&](size_t j) {
        std::vector<Fr> in(isz);
        for (size_t i = 0; i < isz; i++) in[i] = Fr::from_le32(inputs + (j * isz + i) * 32);
        Fr r = Fr::from_le32(rs + 64 * j), s = Fr::from_le32(rs + 64 * j + 32), pub[64];
        const size_t npub = c->z.num_instance - 1;
        ProofOut po;
        if (!prove_one(*c, in.data(), r, s, po, pub, inner)) { fails++; return; }
        g1_out(po.a, proofs_out + 256 * j);
        g2_out(po.b, proofs_out + 256 * j + 64);
        g1_out(po.c, proofs_out + 256 * j + 192);
        if (pub_out) for (size_t i = 0; i < npub; i++) pub[i].to_le32(pub_out + (npub * j + i) * 32);
    });
    return fails.load();
}

// Batch verify: proofs n × 256 B, pub n × npub × 32 B, ok_out n bytes.
void orc_verify_batch(void* p, size_t n, const uint8_t* proofs, const uint8_t* pub, size_t npub, uint8_t* ok_out, int nthreads) {
    OracleCtx* c = (OracleCtx*)p;
    pairing_consts();
    parallel_for(n, nthreads, [&](size_t j) {
        ProofOut po{g1_in(proofs + 256 * j), g2_in(proofs + 256 * j + 64), g1_in(proofs + 256 * j + 192)};
        std::vector<Fr> pv(npub);
        for (size_t i = 0; i < npub; i++) pv[i] = Fr::from_le32(pub + (j * npub + i) * 32);
        ok_out[j] = verify_one(c->z, po, pv.data(), npub) ? 1 : 0;
    });
}

// Variable-base MSMs: points affine canonical (64 B / 128 B each), scalars 32 B LE.
void orc_msm_g1(const uint8_t* pts, const uint8_t* sc, size_t n, uint8_t* out, int nthreads) {
    std::vector<G1A> P(n); std::vector<U256> S(n);
    for (size_t i = 0; i < n; i++) { P[i] = g1_in(pts + 64 * i); memcpy(S[i].l, sc + 32 * i, 32); }
    g1_out(msm_pippenger<Fq>(P.data(), S.data(), n, nthreads).to_affine(), out);
}
void orc_msm_g2(const uint8_t* pts, const uint8_t* sc, size_t n, uint8_t* out, int nthreads) {
    std::vector<G2A> P(n); std::vector<U256> S(n);
    for (size_t i = 0; i < n; i++) { P[i] = g2_in(pts + 128 * i); memcpy(S[i].l, sc + 32 * i, 32); }
    g2_out(msm_pippenger<Fq2>(P.data(), S.data(), n, nthreads).to_affine(), out);
}
// k·G for the G1 generator (1,2): bench/test base generation.  ks: n × 32 B.
void orc_g1_mul_gen(const uint8_t* ks, size_t n, uint8_t* out, int nthreads) {
    G1J g = G1J::from_affine({Fq::from_u64(1), Fq::from_u64(2), false});
    parallel_for(n, nthreads, [&](size_t i) {
        U256 k; memcpy(k.l, ks + 32 * i, 32);
        g1_out(g.mul(k).to_affine(), out + 64 * i);
    });
}
// Σ kᵢ·sᵢ mod r.  Checker for MSMs too large to redo on the CPU: with bases kᵢ·G the MSM must equal (Σ kᵢ·sᵢ)·G, an identity that
// does not go through any group arithmetic of the code under test.  ks, ss: n × 32 B little-endian (reduced mod r on load).
void orc_fr_dot(const uint8_t* ks, const uint8_t* ss, size_t n, uint8_t* out, int nthreads) {
    const size_t T = nthreads > 1 ? (size_t)nthreads : 1;
    std::vector<Fr> part(T, Fr::zero());
    parallel_for(T, (int)T, [&](size_t t) {
        Fr acc = Fr::zero();
        for (size_t i = n * t / T, e = n * (t + 1) / T; i < e; i++) acc = acc + Fr::from_le32(ks + 32 * i) * Fr::from_le32(ss + 32 * i);
        part[t] = acc;
    });
    Fr tot = Fr::zero();
    for (const Fr& v : part) tot = tot + v;
    tot.to_le32(out);
}
// single-thread cost of the port's two primitives, so a reader can scale the CPU baseline against another library's figures:
// what = 0: ns per Fq Montgomery product (dependent chain), 1: ns per G1 mixed addition (Jacobian += affine)
double orc_bench_primitive(int what, int iters) {
    if (iters < 1) iters = 1;
    auto t0 = std::chrono::steady_clock::now();
    if (what == 0) {
        Fq a = Fq::from_u64(0x1234567), b = Fq::from_u64(0x89abcdef);
        for (int i = 0; i < iters; i++) a = a * b;
        volatile uint64_t sink = a.v.l[0];
        (void)sink;
    } else {
        G1J g = G1J::from_affine({Fq::from_u64(1), Fq::from_u64(2), false});
        U256 k; memset(&k, 0, sizeof k); k.l[0] = 0x9e3779b97f4a7c15ull;
        const auto p = g.mul(k).to_affine();
        G1J acc = g;
        for (int i = 0; i < iters; i++) acc = acc.add_affine(p);
        volatile uint64_t sink = acc.to_affine().x.v.l[0];
        (void)sink;
    }
    return std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / iters;
}
// zkey accessors for parity tests: which: 0 a_query 1 b_g1 2 h_query 3 l_query 4 gamma_abc
size_t orc_ctx_g1_vec(void* p, int which, uint8_t* out) {
    OracleCtx* c = (OracleCtx*)p;
    const std::vector<G1A>* v = which == 0 ? &c->z.a_query : which == 1 ? &c->z.b_g1 : which == 2 ? &c->z.h_query : which == 3 ? &c->z.l_query : &c->z.gamma_abc;
    if (out) for (size_t i = 0; i < v->size(); i++) g1_out((*v)[i], out + 64 * i);
    return v->size();
}
size_t orc_ctx_g2_vec(void* p, uint8_t* out) {
    OracleCtx* c = (OracleCtx*)p;
    if (out) for (size_t i = 0; i < c->z.b_g2.size(); i++) g2_out(c->z.b_g2[i], out + 128 * i);
    return c->z.b_g2.size();
}
// forward / inverse NTT of size n (power of two) with ark's root of unity; in-place on 32 B LE values
void orc_ntt(uint8_t* data, size_t n, int inverse) {
    std::vector<Fr> a(n);
    for (size_t i = 0; i < n; i++) a[i] = Fr::from_le32(data + 32 * i);
    Fr w = root_of_unity(n);
    if (inverse) intt_inplace(a, w); else ntt_inplace(a, w);
    for (size_t i = 0; i < n; i++) a[i].to_le32(data + 32 * i);
}

}  // extern "C"
