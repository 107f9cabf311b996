// Host-side helpers of the boundary layer: byte/bigint utilities, Keccak-256 (hash_to_field),
// the arkzkey and witnesscalc-graph parsers.  No proving arithmetic happens here: everything that
// touches field elements beyond byte shuffling is handed to the GPU.
//
//   Keccak-256 / hash_to_field   rln/src/hashers.rs:73-93  (tiny-keccak 2.0.2, padding 0x01…0x80)
//   arkzkey layout               rln/src/circuit/mod.rs:256-305 (ark-serialize uncompressed, unchecked)
//   graph.bin layout             rln/src/circuit/iden3calc/storage.rs:16-22,265-302; proto.rs:7-117
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "vm.cuh"

namespace zk {

// ------------------------------------------------------------------------------- 256-bit byte helpers
static const uint8_t FR_MODULUS_LE[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                                          0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

inline int cmp_le32(const uint8_t* a, const uint8_t* b) {
    for (int i = 31; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
inline bool fr_is_canonical(const uint8_t* a) { return cmp_le32(a, FR_MODULUS_LE) < 0; }
inline void sub_le32(uint8_t* a, const uint8_t* b) {
    int borrow = 0;
    for (int i = 0; i < 32; i++) {
        int d = (int)a[i] - b[i] - borrow;
        borrow = d < 0;
        a[i] = (uint8_t)(d & 0xff);
    }
}
// value mod r for an arbitrary 256-bit little-endian integer (2^256 / r < 6)
inline void fr_reduce(uint8_t* a) {
    while (!fr_is_canonical(a)) sub_le32(a, FR_MODULUS_LE);
}
inline bool is_zero32(const uint8_t* a) {
    for (int i = 0; i < 32; i++)
        if (a[i]) return false;
    return true;
}
// decimal string of a 256-bit little-endian integer (ark-ff Display prints decimal)
inline std::string decimal_le32(const uint8_t* a) {
    uint32_t w[8];
    memcpy(w, a, 32);
    std::string out;
    for (;;) {
        bool nz = false;
        uint64_t rem = 0;
        for (int i = 7; i >= 0; i--) {
            uint64_t cur = (rem << 32) | w[i];
            w[i] = (uint32_t)(cur / 10);
            rem = cur % 10;
            nz = nz || w[i];
        }
        out.push_back((char)('0' + rem));
        if (!nz) break;
    }
    return std::string(out.rbegin(), out.rend());
}

// ------------------------------------------------------------------------------- Keccak-256
inline void keccak_f1600(uint64_t st[25]) {
    static const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
                                    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                                    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
                                    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                                    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
                                    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ ((bc[(i + 1) % 5] << 1) | (bc[(i + 1) % 5] >> 63));
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = PIL[i];
            uint64_t b = st[j];
            st[j] = (t << ROT[i]) | (t >> (64 - ROT[i]));
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[round];
    }
}
inline void keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint64_t st[25];
    memset(st, 0, sizeof st);
    const size_t rate = 136;
    uint8_t block[136];
    size_t off = 0;
    for (;;) {
        size_t take = len - off < rate ? len - off : rate;
        memset(block, 0, rate);
        memcpy(block, data + off, take);
        off += take;
        bool last = take < rate;
        if (last) {
            block[take] ^= 0x01;
            block[rate - 1] ^= 0x80;
        }
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w;
            memcpy(&w, block + 8 * i, 8);
            st[i] ^= w;
        }
        keccak_f1600(st);
        if (last) break;
    }
    memcpy(out, st, 32);
}

// ------------------------------------------------------------------------------- arkzkey
struct ZkeyHost {
    // raw point bytes exactly as in the file (x|y canonical LE, flags in the last byte)
    std::vector<uint8_t> alpha_g1, beta_g1, delta_g1;     // 64 each
    std::vector<uint8_t> beta_g2, gamma_g2, delta_g2;     // 128 each
    std::vector<uint8_t> gamma_abc, a_query, b_g1, h_query, l_query;  // n × 64
    std::vector<uint8_t> b_g2;                                      // n × 128
    uint64_t num_instance = 0, num_witness = 0, num_constraints = 0;
    // CSR matrices, coefficients canonical LE
    std::vector<uint32_t> a_ptr, a_col, b_ptr, b_col;
    std::vector<uint8_t> a_val, b_val;
};

struct ByteReader {
    const uint8_t* p;
    size_t n, o = 0;
    void need(size_t k) const {
        if (k > n - o) throw std::runtime_error("unexpected end of zkey data");
    }
    uint64_t u64() {
        need(8);
        uint64_t v;
        memcpy(&v, p + o, 8);
        o += 8;
        return v;
    }
    void take(std::vector<uint8_t>& dst, size_t k) {
        need(k);
        dst.insert(dst.end(), p + o, p + o + k);
        o += k;
    }
};

inline void parse_zkey(const uint8_t* data, size_t n, ZkeyHost& z) {
    ByteReader r{data, n};
    r.take(z.alpha_g1, 64);
    r.take(z.beta_g2, 128);
    r.take(z.gamma_g2, 128);
    r.take(z.delta_g2, 128);
    auto vec = [&](std::vector<uint8_t>& dst, size_t elem) {
        uint64_t k = r.u64();
        if (k > n / elem) throw std::runtime_error("zkey vector length exceeds file size");
        r.take(dst, k * elem);
    };
    vec(z.gamma_abc, 64);
    r.take(z.beta_g1, 64);
    r.take(z.delta_g1, 64);
    vec(z.a_query, 64);
    vec(z.b_g1, 64);
    vec(z.b_g2, 128);
    vec(z.h_query, 64);
    vec(z.l_query, 64);
    z.num_instance = r.u64();
    z.num_witness = r.u64();
    z.num_constraints = r.u64();
    r.u64();
    r.u64();
    r.u64();
    auto mat = [&](std::vector<uint32_t>& ptr, std::vector<uint32_t>& col, std::vector<uint8_t>& val) {
        uint64_t rows = r.u64();
        if (rows > n) throw std::runtime_error("zkey matrix too large");
        ptr.assign(1, 0);
        for (uint64_t i = 0; i < rows; i++) {
            uint64_t k = r.u64();
            if (k > n / 40) throw std::runtime_error("zkey matrix row too large");
            for (uint64_t e = 0; e < k; e++) {
                r.take(val, 32);
                const uint64_t c = r.u64();
                if (c > 0xffffffffull) throw std::runtime_error("zkey matrix column out of range");
                col.push_back((uint32_t)c);
            }
            ptr.push_back((uint32_t)col.size());
        }
    };
    mat(z.a_ptr, z.a_col, z.a_val);
    mat(z.b_ptr, z.b_col, z.b_val);
    std::vector<uint32_t> cp, cc;
    std::vector<uint8_t> cv;
    mat(cp, cc, cv);
    if (r.o != n) throw std::runtime_error("trailing bytes after zkey");
    if (z.a_ptr.size() != z.num_constraints + 1 || z.b_ptr.size() != z.num_constraints + 1)
        throw std::runtime_error("zkey matrices do not match num_constraints");
}

// ------------------------------------------------------------------------------- graph.bin
struct GraphHost {
    std::vector<VmInstr> prog;
    std::vector<uint8_t> consts;  // canonical LE, 32 bytes each (already reduced mod r)
    std::vector<uint32_t> signals;
    std::map<std::string, std::pair<uint32_t, uint32_t>> inputs;
    uint32_t n_slots = 0;
};

inline bool rd_varint(const uint8_t* p, size_t n, size_t& o, uint64_t& v) {
    v = 0;
    for (int s = 0; s < 70; s += 7) {
        if (o >= n) return false;
        uint8_t c = p[o++];
        v |= (uint64_t)(c & 0x7f) << s;
        if (!(c & 0x80)) return true;
    }
    return false;
}
struct PbField {
    uint64_t tag, wt, val;
    const uint8_t* ptr;
    size_t len;
};
inline bool pb_next(const uint8_t* p, size_t n, size_t& o, PbField& f) {
    uint64_t key;
    if (!rd_varint(p, n, o, key)) return false;
    f.tag = key >> 3;
    f.wt = key & 7;
    f.val = 0;
    f.ptr = nullptr;
    f.len = 0;
    switch (f.wt) {
        case 0: return rd_varint(p, n, o, f.val);
        case 2: {
            uint64_t l;
            if (!rd_varint(p, n, o, l) || l > n - o) return false;
            f.ptr = p + o;
            f.len = l;
            o += l;
            return true;
        }
        case 5: if (o + 4 > n) return false; o += 4; return true;
        case 1: if (o + 8 > n) return false; o += 8; return true;
        default: return false;
    }
}
// little-endian bytes of any length reduced mod r (Fr::from_le_bytes_mod_order, storage.rs:45-47)
inline void le_bytes_mod_r(const uint8_t* b, size_t n, uint8_t out[32]) {
    // Horner over bytes from the most significant end with 320-bit scratch: acc = acc·256 + byte (mod r)
    uint8_t acc[33];
    memset(acc, 0, sizeof acc);
    for (size_t i = n; i-- > 0;) {
        for (int k = 32; k > 0; k--) acc[k] = acc[k - 1];
        acc[0] = b[i];
        // acc < 256·r < 2^264: subtract r·2^j greedily (at most 8 bits worth)
        for (int bit = 8; bit >= 0; bit--) {
            uint8_t m[33];
            memset(m, 0, sizeof m);
            // m = r << bit
            unsigned carry = 0;
            for (int k = 0; k < 32; k++) {
                unsigned v = ((unsigned)FR_MODULUS_LE[k] << bit) | carry;
                m[k] = (uint8_t)(v & 0xff);
                carry = v >> 8;
            }
            m[32] = (uint8_t)carry;
            int c = 0;
            for (int k = 32; k >= 0; k--) {
                if (acc[k] != m[k]) { c = acc[k] < m[k] ? -1 : 1; break; }
            }
            if (c >= 0) {
                int borrow = 0;
                for (int k = 0; k < 33; k++) {
                    int d = (int)acc[k] - m[k] - borrow;
                    borrow = d < 0;
                    acc[k] = (uint8_t)(d & 0xff);
                }
            }
        }
    }
    memcpy(out, acc, 32);
}

inline void parse_graph(const uint8_t* d, size_t n, GraphHost& g) {
    static const char MAGIC[] = "wtns.graph.001";
    if (n < 22 || memcmp(d, MAGIC, 14)) throw std::runtime_error("Invalid magic");
    size_t o = 14;
    uint64_t cnt;
    memcpy(&cnt, d + o, 8);
    o += 8;
    if (cnt > n) throw std::runtime_error("graph node count exceeds file size");
    g.prog.reserve(cnt);
    for (uint64_t i = 0; i < cnt; i++) {
        uint64_t len;
        if (!rd_varint(d, n, o, len) || len > n - o) throw std::runtime_error("Unexpected EOF");
        const uint8_t* m = d + o;
        o += len;
        size_t mo = 0;
        PbField f;
        if (!pb_next(m, len, mo, f) || f.wt != 2) throw std::runtime_error("Proto::Node must have a node field");
        uint64_t vals[5] = {0, 0, 0, 0, 0};
        const uint8_t* sub = nullptr;
        size_t sublen = 0, bo = 0;
        PbField bf;
        while (bo < f.len) {
            if (!pb_next(f.ptr, f.len, bo, bf)) throw std::runtime_error("malformed node");
            if (bf.tag < 5 && bf.wt == 0) vals[bf.tag] = bf.val;
            if (bf.tag == 1 && bf.wt == 2) { sub = bf.ptr; sublen = bf.len; }
        }
        for (int k = 1; k < 5; k++)   // node indices are u32 on the device; the reference indexes a Vec and would panic
            if (vals[k] > 0xffffffffull) throw std::runtime_error("node operand out of range");
        VmInstr in{0, 0, 0, 0};
        switch (f.tag) {
            case 1: in.kind_op = VM_INPUT; in.a = (uint32_t)vals[1]; break;
            case 2: {
                const uint8_t* vb = nullptr;
                size_t vl = 0, so = 0;
                PbField sf;
                if (!sub) throw std::runtime_error("Constant node must have a value");
                while (so < sublen) {
                    if (!pb_next(sub, sublen, so, sf)) throw std::runtime_error("malformed constant");
                    if (sf.tag == 1 && sf.wt == 2) { vb = sf.ptr; vl = sf.len; }
                }
                in.kind_op = VM_CONST;
                in.a = (uint32_t)(g.consts.size() / 32);
                uint8_t c[32];
                le_bytes_mod_r(vb, vl, c);
                g.consts.insert(g.consts.end(), c, c + 32);
                break;
            }
            case 3:
                if (vals[1] > 1) throw std::runtime_error("UnoOp must be valid enum value");
                in.kind_op = VM_UNO | ((uint32_t)vals[1] << 8); in.a = (uint32_t)vals[2]; break;
            case 4:
                if (vals[1] > 19) throw std::runtime_error("DuoOp must be valid enum value");
                in.kind_op = VM_DUO | ((uint32_t)vals[1] << 8); in.a = (uint32_t)vals[2]; in.b = (uint32_t)vals[3]; break;
            case 5:
                if (vals[1] != 0) throw std::runtime_error("TresOp must be valid enum value");
                in.kind_op = VM_TRES; in.a = (uint32_t)vals[2]; in.b = (uint32_t)vals[3]; in.c = (uint32_t)vals[4]; break;
            default: throw std::runtime_error("Proto::Node must have a node field");
        }
        // operands must refer to earlier nodes (topological order is what evaluate() relies on)
        uint32_t kind = in.kind_op & 0xff;
        if ((kind == VM_UNO || kind == VM_DUO || kind == VM_TRES) && in.a >= i) throw std::runtime_error("node operand out of order");
        if ((kind == VM_DUO || kind == VM_TRES) && in.b >= i) throw std::runtime_error("node operand out of order");
        if (kind == VM_TRES && in.c >= i) throw std::runtime_error("node operand out of order");
        g.prog.push_back(in);
    }
    uint64_t mdlen;
    if (!rd_varint(d, n, o, mdlen) || mdlen > n - o) throw std::runtime_error("Unexpected EOF");
    const uint8_t* md = d + o;
    size_t mo = 0;
    PbField f;
    while (mo < mdlen) {
        if (!pb_next(md, mdlen, mo, f)) throw std::runtime_error("malformed graph metadata");
        if (f.tag == 1 && f.wt == 2) {
            size_t po = 0;
            uint64_t v;
            while (po < f.len) {
                if (!rd_varint(f.ptr, f.len, po, v) || v > 0xffffffffull) throw std::runtime_error("malformed witness_signals");
                g.signals.push_back((uint32_t)v);
            }
        } else if (f.tag == 1 && f.wt == 0) {
            if (f.val > 0xffffffffull) throw std::runtime_error("malformed witness_signals");
            g.signals.push_back((uint32_t)f.val);
        } else if (f.tag == 2 && f.wt == 2) {
            size_t eo = 0;
            PbField ef;
            std::string key;
            uint64_t off = 0, ln = 0;
            while (eo < f.len) {
                if (!pb_next(f.ptr, f.len, eo, ef)) throw std::runtime_error("malformed inputs map");
                if (ef.tag == 1 && ef.wt == 2) key.assign((const char*)ef.ptr, ef.len);
                if (ef.tag == 2 && ef.wt == 2) {
                    size_t so = 0;
                    PbField sf;
                    while (so < ef.len) {
                        if (!pb_next(ef.ptr, ef.len, so, sf)) throw std::runtime_error("malformed signal description");
                        if (sf.tag == 1) off = sf.val;
                        if (sf.tag == 2) ln = sf.val;
                    }
                }
            }
            if (off > 0xffffffffull || ln > 0xffffffffull) throw std::runtime_error("input signal out of range");
            g.inputs[key] = {(uint32_t)off, (uint32_t)ln};
        }
    }
    for (uint32_t s : g.signals)
        if (s >= g.prog.size()) throw std::runtime_error("witness signal index out of range");
    bool started = false;  // iden3calc.rs:106-121 get_inputs_size
    uint32_t mx = 0;
    for (auto& in : g.prog) {
        if ((in.kind_op & 0xff) == VM_INPUT) { mx = in.a > mx ? in.a : mx; started = true; }
        else if (started) break;
    }
    g.n_slots = mx + 1;
    for (auto& in : g.prog)
        if ((in.kind_op & 0xff) == VM_INPUT && in.a >= g.n_slots) throw std::runtime_error("input index out of range");
    for (auto& kv : g.inputs)   // populate_inputs (iden3calc.rs:123-146) writes inputs[offset .. offset + len)
        if ((uint64_t)kv.second.first + kv.second.second > g.n_slots) throw std::runtime_error("input signal out of range");
}


// ------------------------------------------------------------------------------- witness VM schedule
// ---- depth reduction of the witness program ---------------------------------------------------------------------------------
// k_witness is paced by the graph's longest dependency chain, not by its node count (one CTA, lone warps: a bundle whose slowest
// slot is a product costs ≈ 1 200 cycles however few nodes it holds).  circom's Poseidon spends, per partial round and on that
// chain, x → x² → x⁴ → x⁵ → x⁵ + c → k·(x⁵ + c) → + a → + b: four products and three additions.  Field arithmetic is exact, so the
// program may be re-associated freely as long as every WIRE keeps its value:
//   k·(t + c)            → k·t + k·c        (constants folded; t + c is repeated in each of its consumers, it is one addition)
//   k·(x⁴·x)             → (k·x)·x⁴         (products re-paired so that the factor that is ready last is multiplied in last)
//   ((late + e₁) + e₂)   → late + (e₁ + e₂) (sums likewise)
// which leaves x → x² → x⁴ → (k·x)·x⁴ → + (k·c + a + b): three products and one addition (−29 % on the depth-20 graph's chain).
// Nodes that are wires, or have several consumers, are kept as they are (and may be computed a second time inside a consumer's
// re-associated form); identical nodes are merged.  The node count grows (23 414 → 37 481 at depth 20: the (k·x) products are
// new work), which is why k_witness runs eight slots per bundle.  Output: a new program, constant table and wire → node map.
struct VmOptimized {
    std::vector<VmInstr> prog;
    std::vector<uint8_t> consts;
    std::vector<uint32_t> signals;
};
class VmOptimizer {
    static constexpr uint64_t COST_MUL = 1170, COST_ADD = 400;   // cycles of a bundle by its slowest slot (measured), for ordering only
    const std::vector<VmInstr>& in_;
    const std::vector<uint8_t>& cin_;
    std::vector<uint8_t> wire_;
    std::vector<uint32_t> uses_;
    std::vector<int64_t> map_;
    std::vector<VmInstr> out_;
    std::vector<uint64_t> ready_;
    std::vector<Fr> cval_;                                       // value of every new constant (Montgomery), by constant index
    std::map<std::vector<uint32_t>, uint32_t> cmap_;             // canonical words → node
    std::map<std::vector<uint32_t>, uint32_t> cse_;
    static uint32_t kind(const VmInstr& i) { return i.kind_op & 0xff; }
    static uint32_t op(const VmInstr& i) { return i.kind_op >> 8; }
    bool is_duo(uint32_t o, uint32_t want) const { return kind(in_[o]) == VM_DUO && op(in_[o]) == want; }
    bool is_const(uint32_t o) const { return kind(in_[o]) == VM_CONST; }
    Fr const_old(uint32_t o) const {
        u32 c[8];
        memcpy(c, cin_.data() + 32 * (size_t)in_[o].a, 32);
        return Fr::from_canonical(c);
    }
    uint32_t add_const(const Fr& v) {
        u32 c[8];
        v.to_canonical(c);
        std::vector<uint32_t> key(c, c + 8);
        auto it = cmap_.find(key);
        if (it != cmap_.end()) return it->second;
        VmInstr n{VM_CONST, (uint32_t)cval_.size(), 0, 0};
        cval_.push_back(v);
        out_.push_back(n);
        ready_.push_back(0);
        cmap_[key] = (uint32_t)out_.size() - 1;
        return (uint32_t)out_.size() - 1;
    }
    uint32_t add_node(VmInstr n) {
        const uint32_t k = kind(n);
        if (k == VM_DUO && (op(n) == OP_ADD || op(n) == OP_MUL) && n.a > n.b) std::swap(n.a, n.b);
        std::vector<uint32_t> key = {n.kind_op, n.a, n.b, n.c};
        if (k != VM_INPUT) {
            auto it = cse_.find(key);
            if (it != cse_.end()) return it->second;
        }
        uint64_t r = 0;
        if (k == VM_UNO || k == VM_DUO || k == VM_TRES) r = ready_[n.a];
        if (k == VM_DUO || k == VM_TRES) r = std::max(r, ready_[n.b]);
        if (k == VM_TRES) r = std::max(r, ready_[n.c]);
        const uint64_t c = k == VM_INPUT ? 0 : (k == VM_DUO && op(n) != OP_ADD && op(n) != OP_SUB) ? COST_MUL : COST_ADD;
        out_.push_back(n);
        ready_.push_back(r + c);
        if (k != VM_INPUT) cse_[key] = (uint32_t)out_.size() - 1;
        return (uint32_t)out_.size() - 1;
    }
    // may the consumer fold old node o (a sum / a product) into its own expression?
    bool absorbable(uint32_t o, uint32_t want) const {
        if (!is_duo(o, want)) return false;
        if (wire_[o]) return false;
        if (want == OP_ADD && (is_const(in_[o].a) || is_const(in_[o].b))) return true;   // t + c: cheap to repeat per consumer
        return uses_[o] == 1;
    }
    void flat(uint32_t o, uint32_t want, std::vector<uint32_t>& leaves) const {
        if (absorbable(o, want)) { flat(in_[o].a, want, leaves); flat(in_[o].b, want, leaves); }
        else leaves.push_back(o);
    }
    uint32_t combine(std::vector<uint32_t> items, uint32_t o) {   // the two operands that are ready first are paired first
        auto by_ready = [&](uint32_t x, uint32_t y) { return ready_[x] != ready_[y] ? ready_[x] < ready_[y] : x < y; };
        std::sort(items.begin(), items.end(), by_ready);
        while (items.size() > 1) {
            const uint32_t n = add_node(VmInstr{VM_DUO | (o << 8), items[0], items[1], 0});
            items.erase(items.begin(), items.begin() + 2);
            items.insert(std::upper_bound(items.begin(), items.end(), n, by_ready), n);
        }
        return items[0];
    }
    uint32_t prod_of(std::vector<uint32_t> new_items, const std::vector<uint32_t>& old_factors) {
        Fr c = Fr::one();
        bool have_c = false;
        for (uint32_t f : old_factors) {
            if (is_const(f)) { c = c * const_old(f); have_c = true; }
            else new_items.push_back(emit(f));
        }
        if (have_c && (c != Fr::one() || new_items.empty())) new_items.push_back(add_const(c));
        if (new_items.empty()) new_items.push_back(add_const(Fr::one()));
        return combine(new_items, OP_MUL);
    }
    // old node o as a sum: new summands + a constant
    void sum_items(uint32_t o, std::vector<uint32_t>& items, Fr& c) {
        if (is_const(o)) { c = c + const_old(o); return; }
        if (absorbable(o, OP_ADD)) { sum_items(in_[o].a, items, c); sum_items(in_[o].b, items, c); return; }
        if (absorbable(o, OP_MUL)) {
            std::vector<uint32_t> fs, rest;
            flat(o, OP_MUL, fs);
            Fr k = Fr::one();
            for (uint32_t f : fs) { if (is_const(f)) k = k * const_old(f); else rest.push_back(f); }
            if (k != Fr::one() && rest.size() == 1 && absorbable(rest[0], OP_ADD)) {
                std::vector<uint32_t> terms, nc;
                flat(rest[0], OP_ADD, terms);
                Fr cs = Fr::zero();
                for (uint32_t t : terms) { if (is_const(t)) cs = cs + const_old(t); else nc.push_back(t); }
                if (nc.size() <= 1) {   // k·(t + c) = k·t + k·c, the product k·t re-paired by readiness
                    if (!nc.empty()) {
                        std::vector<uint32_t> f2;
                        flat(nc[0], OP_MUL, f2);
                        items.push_back(prod_of({add_const(k)}, f2));
                    }
                    c = c + k * cs;
                    return;
                }
            }
        }
        items.push_back(emit(o));
    }
    uint32_t emit(uint32_t o) {
        if (map_[o] >= 0) return (uint32_t)map_[o];
        const VmInstr& n = in_[o];
        uint32_t r;
        if (is_const(o)) {
            r = add_const(const_old(o));
        } else if (is_duo(o, OP_ADD)) {
            std::vector<uint32_t> items;
            Fr c = Fr::zero();
            sum_items(n.a, items, c);
            sum_items(n.b, items, c);
            if (!c.is_zero() || items.empty()) items.push_back(add_const(c));
            r = combine(items, OP_ADD);
        } else if (is_duo(o, OP_MUL)) {
            std::vector<uint32_t> fs;
            flat(n.a, OP_MUL, fs);
            flat(n.b, OP_MUL, fs);
            // k·(t + c) at the top of a product that nobody sums up: the same distribution, as a sum of its own
            std::vector<uint32_t> items;
            Fr c = Fr::zero();
            bool as_sum = false;
            {
                std::vector<uint32_t> rest;
                Fr k = Fr::one();
                for (uint32_t f : fs) { if (is_const(f)) k = k * const_old(f); else rest.push_back(f); }
                if (k != Fr::one() && rest.size() == 1 && absorbable(rest[0], OP_ADD)) {
                    std::vector<uint32_t> terms, nc;
                    flat(rest[0], OP_ADD, terms);
                    Fr cs = Fr::zero();
                    for (uint32_t t : terms) { if (is_const(t)) cs = cs + const_old(t); else nc.push_back(t); }
                    if (nc.size() <= 1) {
                        if (!nc.empty()) {
                            std::vector<uint32_t> f2;
                            flat(nc[0], OP_MUL, f2);
                            items.push_back(prod_of({add_const(k)}, f2));
                        }
                        c = k * cs;
                        as_sum = true;
                    }
                }
            }
            if (as_sum) {
                if (!c.is_zero() || items.empty()) items.push_back(add_const(c));
                r = combine(items, OP_ADD);
            } else {
                r = prod_of({}, fs);
            }
        } else {
            VmInstr m = n;
            const uint32_t k = kind(n);
            if (k == VM_UNO || k == VM_DUO || k == VM_TRES) m.a = emit(n.a);
            if (k == VM_DUO || k == VM_TRES) m.b = emit(n.b);
            if (k == VM_TRES) m.c = emit(n.c);
            r = add_node(m);
        }
        map_[o] = r;
        return r;
    }

public:
    VmOptimizer(const std::vector<VmInstr>& prog, const std::vector<uint8_t>& consts, const std::vector<uint32_t>& signals)
        : in_(prog), cin_(consts), wire_(prog.size(), 0), uses_(prog.size(), 0), map_(prog.size(), -1) {
        for (uint32_t s : signals) wire_[s] = 1;
        for (const VmInstr& n : prog) {
            const uint32_t k = kind(n);
            if (k == VM_UNO || k == VM_DUO || k == VM_TRES) uses_[n.a]++;
            if (k == VM_DUO || k == VM_TRES) uses_[n.b]++;
            if (k == VM_TRES) uses_[n.c]++;
        }
    }
public:
    VmOptimized run(const std::vector<uint32_t>& signals) {
        for (uint32_t o = 0; o < in_.size(); o++) {
            // a node that its consumers fold in is emitted by them (or never); wires, shared nodes and everything else now
            if (!wire_[o] && (absorbable(o, OP_ADD) || absorbable(o, OP_MUL) || is_const(o))) continue;
            if (!wire_[o] && uses_[o] == 0) continue;
            emit(o);
        }
        std::vector<uint32_t> sig;
        for (uint32_t s : signals) sig.push_back(emit(s));
        fuse_and_compact(sig);
        VmOptimized r;
        r.prog = out_;
        r.signals = sig;
        r.consts.resize(32 * cval_.size());
        for (size_t i = 0; i < cval_.size(); i++) {
            u32 c[8];
            cval_[i].to_canonical(c);
            memcpy(r.consts.data() + 32 * i, c, 32);
        }
        return r;
    }

private:
    // late·product + early sum → one node (a·b + c: the addition rides in the product's bundle), then dead nodes are dropped
    void fuse_and_compact(std::vector<uint32_t>& sig) {
        const size_t n = out_.size();
        std::vector<uint32_t> uses(n, 0);
        std::vector<uint8_t> wire(n, 0), dead(n, 0);
        for (uint32_t s : sig) wire[s] = 1;
        auto count = [&](const VmInstr& i, int d) {
            const uint32_t k = kind(i);
            if (k == VM_UNO || k == VM_DUO || k == VM_TRES) uses[i.a] += d;
            if (k == VM_DUO || k == VM_TRES) uses[i.b] += d;
            if (k == VM_TRES) uses[i.c] += d;
        };
        for (const VmInstr& i : out_) count(i, 1);
        for (size_t i = 0; i < n; i++) {
            VmInstr& a = out_[i];
            if (!(kind(a) == VM_DUO && op(a) == OP_ADD)) continue;
            uint32_t p = a.a, q = a.b;
            auto fusable = [&](uint32_t x) { return kind(out_[x]) == VM_DUO && op(out_[x]) == OP_MUL && uses[x] == 1 && !wire[x]; };
            if (fusable(q) && (!fusable(p) || ready_[q] > ready_[p])) std::swap(p, q);
            if (!fusable(p) || ready_[p] < ready_[q]) continue;   // only when the product is the operand that arrives last
            const VmInstr m = out_[p];
            a = VmInstr{VM_TRES | (VM_TRES_FMA << 8), m.a, m.b, q};
            dead[p] = 1;
        }
        for (size_t i = n; i-- > 0;) {   // whatever nobody reads any more
            if (dead[i]) continue;
            if (!wire[i] && uses[i] == 0) dead[i] = 1;
        }
        std::vector<uint32_t> renum(n, 0);
        std::vector<VmInstr> packed;
        for (size_t i = 0; i < n; i++) {
            if (dead[i]) continue;
            VmInstr x = out_[i];
            const uint32_t k = kind(x);
            if (k == VM_UNO || k == VM_DUO || k == VM_TRES) x.a = renum[x.a];
            if (k == VM_DUO || k == VM_TRES) x.b = renum[x.b];
            if (k == VM_TRES) x.c = renum[x.c];
            renum[i] = (uint32_t)packed.size();
            packed.push_back(x);
        }
        for (uint32_t& s : sig) s = renum[s];
        out_.swap(packed);
    }
};
inline VmOptimized vm_optimize_program(const std::vector<VmInstr>& prog, const std::vector<uint8_t>& consts, const std::vector<uint32_t>& signals) {
    VmOptimizer o(prog, consts, signals);
    return o.run(signals);
}

// k_witness evaluates the graph with four warps per 32 proofs: the nodes are list-scheduled into bundles of ≤ VM_SLOTS mutually
// independent nodes (every operand lies in an earlier bundle), warp w executes slot w.  Recent values are also kept in a
// shared-memory ring of VM_RING bundles, so an operand produced ≤ VM_RING − 1 bundles ago is a shared-memory read instead of an
// L2 round trip on the critical path; constants are read from the constant table; anything older comes from vals[node][B].
struct VmRecord {            // 32 bytes: two 128-bit loads per (bundle, slot)
    uint32_t kind_op;        // as VmInstr; 0xffffffff = empty slot
    uint32_t out;            // node index of the result; bit 31: the value is also stored to vals[node][B] (a wire, or a far consumer)
    uint32_t a, b, c;        // operands: source type in the top 2 bits (VM_SRC_*), index below
    uint32_t pad[3];
};
// consts_resident: the kernel keeps the whole constant table in shared memory, so an operand that is a constant node is read from
// there and constant nodes that are not wires need not be evaluated at all; otherwise constants are ordinary nodes (loaded once by
// their own record, then read from the ring or from vals like any other value).  is_signal[i] != 0: node i is a wire of the witness.
inline std::vector<VmRecord> vm_build_schedule(const std::vector<VmInstr>& prog, uint32_t& n_bundles, const std::vector<uint8_t>* is_signal = nullptr,
                                               bool consts_resident = true) {
    const size_t n = prog.size();
    const uint32_t NONE = 0xffffffffu;
    // Slots s and s + VM_SLOTS/2 are warps of the same scheduler.  A product keeps that scheduler's multiplier busy for most of its
    // 865 cycles, so two products on one scheduler take twice as long (measured: 8 free-for-all slots, 1 700 cycles per bundle);
    // an addition next to a product is nearly free.  Hence two classes: heavy nodes (anything with a product in it) go to the
    // first half of the slots, light ones (additions, loads, selections) to the second half.
    const bool two_classes = VM_SLOTS >= 8;   // with one warp per scheduler every slot is as good as any other
    const uint32_t CAP = two_classes ? VM_SLOTS / 2 : VM_SLOTS;
    std::vector<uint32_t> bundle(n, NONE), slot(n, 0);
    std::vector<uint32_t> fill[2];   // per bundle: heavy, light
    auto heavy = [&](size_t i) -> int {
        if (!two_classes) return 0;
        const uint32_t k = prog[i].kind_op & 0xff, o = prog[i].kind_op >> 8;
        if (k == VM_DUO) return (o == OP_ADD || o == OP_SUB) ? 0 : 1;
        if (k == VM_TRES) return o == VM_TRES_FMA ? 1 : 0;
        return 0;
    };
    auto wire = [&](size_t i) { return !is_signal || (*is_signal)[i]; };
    for (size_t i = 0; i < n; i++) {
        const VmInstr& in = prog[i];
        const uint32_t kind = in.kind_op & 0xff;
        if (kind == VM_CONST && consts_resident && !wire(i)) continue;   // dead: every consumer reads the constant table itself
        uint32_t b = 0;
        auto after = [&](uint32_t op) {
            if (bundle[op] == NONE) return;   // a resident constant: available from the start
            if (bundle[op] + 1 > b) b = bundle[op] + 1;
        };
        if (kind == VM_UNO || kind == VM_DUO || kind == VM_TRES) after(in.a);
        if (kind == VM_DUO || kind == VM_TRES) after(in.b);
        if (kind == VM_TRES) after(in.c);
        const int h = heavy(i);
        for (;; b++) {
            if (b >= fill[0].size()) { fill[0].resize(b + 1, 0); fill[1].resize(b + 1, 0); }
            if (fill[h][b] < CAP) break;
        }
        bundle[i] = b;
        fill[h][b]++;
    }
    n_bundles = (uint32_t)fill[0].size();
    // As-soon-as-possible leaves values that are ready early but consumed late (the side sums of every Poseidon round) far from their
    // reader: beyond the ring, i.e. an L2 round trip inside the reader's bundle.  Push every node as late as its readers and the free
    // slots of its class allow (readers have larger indices, so they are final when the node is visited).
    {
        std::vector<uint32_t> first_reader(n, NONE);
        for (size_t i = n; i-- > 0;) {
            if (bundle[i] == NONE) continue;
            const int h = heavy(i);
            if (first_reader[i] != NONE) {
                for (uint32_t b = first_reader[i] - 1; b > bundle[i]; b--)
                    if (fill[h][b] < CAP) { fill[h][bundle[i]]--; fill[h][b]++; bundle[i] = b; break; }
            }
            const VmInstr& in = prog[i];
            const uint32_t kind = in.kind_op & 0xff;
            auto seen_by = [&](uint32_t op) { if (bundle[op] != NONE && bundle[i] < first_reader[op]) first_reader[op] = bundle[i]; };
            if (kind == VM_UNO || kind == VM_DUO || kind == VM_TRES) seen_by(in.a);
            if (kind == VM_DUO || kind == VM_TRES) seen_by(in.b);
            if (kind == VM_TRES) seen_by(in.c);
        }
        std::fill(fill[0].begin(), fill[0].end(), 0);
        std::fill(fill[1].begin(), fill[1].end(), 0);
        for (size_t i = 0; i < n; i++)
            if (bundle[i] != NONE) { const int h = heavy(i); slot[i] = (h || !two_classes ? 0 : CAP) + fill[h][bundle[i]]++; }
    }
    VmRecord empty;
    memset(&empty, 0, sizeof empty);
    empty.kind_op = 0xffffffffu;
    std::vector<VmRecord> recs((size_t)n_bundles * VM_SLOTS, empty);
    std::vector<uint8_t> to_global(n, 0);
    for (size_t i = 0; i < n; i++) to_global[i] = wire(i) ? 1 : 0;
    for (size_t i = 0; i < n; i++) {
        if (bundle[i] == NONE) continue;
        const VmInstr& in = prog[i];
        const uint32_t kind = in.kind_op & 0xff;
        auto enc = [&](uint32_t op) -> uint32_t {
            const uint32_t ok = prog[op].kind_op & 0xff;
            if (ok == VM_CONST && consts_resident) return (VM_SRC_CONST << 30) | prog[op].a;
            if (bundle[i] - bundle[op] <= VM_RING - 1) return (VM_SRC_RING << 30) | ((bundle[op] % VM_RING) * VM_SLOTS + slot[op]);
            to_global[op] = 1;
            return (VM_SRC_GLOBAL << 30) | op;
        };
        VmRecord r = empty;
        r.kind_op = in.kind_op;
        r.out = (uint32_t)i;
        r.a = in.a; r.b = in.b; r.c = in.c;   // INPUT / CONST keep their raw index
        if (kind == VM_UNO || kind == VM_DUO || kind == VM_TRES) r.a = enc(in.a);
        if (kind == VM_DUO || kind == VM_TRES) r.b = enc(in.b);
        if (kind == VM_TRES) r.c = enc(in.c);
        recs[(size_t)bundle[i] * VM_SLOTS + slot[i]] = r;
    }
    for (size_t i = 0; i < n; i++)   // the producers are always earlier in the program than their consumers: flags are final here
        if (bundle[i] != NONE && to_global[i]) recs[(size_t)bundle[i] * VM_SLOTS + slot[i]].out |= 0x80000000u;
    return recs;
}

// ------------------------------------------------------------------------------------------- tree configuration (PmTreeConfig)
// The JSON the reference accepts for its default tree (rln/src/pm_tree_adapter.rs:139-174 `impl FromStr for PmTreeConfig`):
//   { "path": "...", "temporary": bool, "cache_capacity": u64, "flush_every_ms": u64, "mode": "HighThroughput" | "LowSpace",
//     "use_compression": bool, "tree_depth": u64 }      every key optional, unknown keys ignored (serde_json::Value indexing)
// and its path rules (`resolve_path`, :93-100).  cache_capacity / flush_every_ms / mode / use_compression tune sled; they are
// parsed, kept and otherwise meaningless for a tree that lives in HBM.
struct TreeConfig {
    bool has_path = false;
    std::string path;
    bool temporary = true;
    uint64_t cache_capacity = 1073741824ull, flush_every_ms = 500;
    bool low_space = false, use_compression = false;
    bool has_depth = false;
    uint64_t tree_depth = 0;
    bool persistent() const { return !temporary && has_path; }
};
struct JsonCursor {
    const char* p;
    const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("Error while reading pmtree config: ") + what); }
    std::string string() {
        if (p >= end || *p != '"') fail("expected a string");
        p++;
        std::string out;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) fail("EOF while parsing a string");
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {   // \uXXXX: paths are ASCII in practice; keep the low byte of BMP code points below 0x80, reject the rest
                        if (end - p < 5) fail("EOF while parsing a string");
                        unsigned v = 0;
                        for (int i = 1; i <= 4; i++) {
                            const char c = p[i];
                            v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 256);
                        }
                        if (v >= 0x80) fail("unsupported escape in a string");
                        out += (char)v;
                        p += 4;
                        break;
                    }
                    default: out += *p;
                }
                p++;
            } else out += *p++;
        }
        if (p >= end) fail("EOF while parsing a string");
        p++;
        return out;
    }
    void skip_value() {   // any JSON value; nesting by bracket counting (strings skipped properly)
        ws();
        if (p >= end) fail("EOF while parsing a value");
        if (*p == '"') { string(); return; }
        if (*p == '{' || *p == '[') {
            int depth = 0;
            do {
                if (p >= end) fail("EOF while parsing a value");
                if (*p == '"') { string(); continue; }
                if (*p == '{' || *p == '[') depth++;
                if (*p == '}' || *p == ']') depth--;
                p++;
            } while (depth > 0);
            return;
        }
        while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;
    }
};
inline TreeConfig parse_tree_config(const std::string& json) {
    TreeConfig c;
    JsonCursor j{json.data(), json.data() + json.size()};
    j.ws();
    if (j.p >= j.end) j.fail("EOF while parsing a value");
    if (*j.p != '{') {   // serde_json::Value indexing on a non-object yields Null for every key: all defaults, after a syntax check
        j.skip_value();
        j.ws();
        if (j.p != j.end) j.fail("trailing characters");
        return c;
    }
    j.p++;
    j.ws();
    if (j.p < j.end && *j.p == '}') { j.p++; j.ws(); if (j.p != j.end) j.fail("trailing characters"); return c; }
    for (;;) {
        j.ws();
        const std::string key = j.string();
        j.ws();
        if (j.p >= j.end || *j.p != ':') j.fail("expected `:`");
        j.p++;
        j.ws();
        const char* v0 = j.p;
        auto is_lit = [&](const char* lit) { const size_t n = strlen(lit); return (size_t)(j.end - v0) >= n && !memcmp(v0, lit, n); };
        auto as_u64 = [&](uint64_t& out) -> bool {   // as_u64(): non-negative integers only
            const char* q = v0;
            if (q >= j.end || *q < '0' || *q > '9') return false;
            uint64_t v = 0;
            while (q < j.end && *q >= '0' && *q <= '9') { v = v * 10 + (uint64_t)(*q - '0'); q++; }
            if (q < j.end && (*q == '.' || *q == 'e' || *q == 'E')) return false;
            out = v;
            return true;
        };
        if (key == "path" && j.p < j.end && *j.p == '"') { c.path = j.string(); c.has_path = true; }
        else if (key == "mode" && j.p < j.end && *j.p == '"') { c.low_space = j.string() == "LowSpace"; }
        else {
            uint64_t u = 0;
            if (key == "temporary") { if (is_lit("true")) c.temporary = true; else if (is_lit("false")) c.temporary = false; }
            else if (key == "use_compression") { if (is_lit("true")) c.use_compression = true; else if (is_lit("false")) c.use_compression = false; }
            else if (key == "cache_capacity") { if (as_u64(u)) c.cache_capacity = u; }
            else if (key == "flush_every_ms") { if (as_u64(u)) c.flush_every_ms = u; }
            else if (key == "tree_depth") { if (as_u64(u)) { c.tree_depth = u; c.has_depth = true; } }
            j.skip_value();
        }
        j.ws();
        if (j.p < j.end && *j.p == ',') { j.p++; continue; }
        if (j.p < j.end && *j.p == '}') { j.p++; break; }
        j.fail("expected `,` or `}`");
    }
    j.ws();
    if (j.p != j.end) j.fail("trailing characters");
    return c;
}

}  // namespace zk
