// TEST HARNESS (not product): mutation fuzz of the product's host-side file parsers (host_util.hpp: arkzkey, witnesscalc graph,
// VM list schedule), built with -fsanitize=address,undefined by tests/test_host_fuzz.py.  A mutated file must either parse into
// structures whose indices are all in range, or raise std::exception — never read or write out of bounds.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

#include "host_util.hpp"

using namespace zk;

static uint64_t g_s = 1;
static uint64_t rnd() { g_s += 0x9e3779b97f4a7c15ull; uint64_t z = g_s; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static std::vector<uint8_t> slurp(const char* p) {
    std::ifstream f(p, std::ios::binary);
    if (!f) { fprintf(stderr, "cannot open %s\n", p); exit(2); }
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void mutate(std::vector<uint8_t>& b, size_t hot) {   // `hot`: bias positions towards the first `hot` bytes and the tail (headers, metadata)
    auto pos = [&]() -> size_t {
        if (b.empty()) return 0;
        uint64_t k = rnd() % 4;
        if (k == 0) return rnd() % b.size();
        if (k == 1) return rnd() % (hot < b.size() ? hot : b.size());
        size_t tail = b.size() < 2048 ? b.size() : 2048;
        return b.size() - 1 - rnd() % tail;
    };
    static const uint64_t odd[] = {~0ull, 1ull << 63, 1ull << 32, (1ull << 32) + 3, 1ull << 31, 0, 1, 0x7fffffffffffffffull, (1ull << 61) + 1, 0xffffffffull};
    switch (rnd() % 7) {
        case 0: b.resize(pos()); break;
        case 1: for (int i = 0, k = 1 + rnd() % 8; i < k && !b.empty(); i++) b[pos()] ^= (uint8_t)(1u << (rnd() % 8)); break;
        case 2: if (b.size() > 8) { uint64_t v = odd[rnd() % 10]; memcpy(&b[pos() % (b.size() - 8)], &v, 8); } break;
        case 3: if (b.size() > 10) { size_t p = pos() % (b.size() - 10); for (int i = 0; i < 9; i++) b[p + i] = 0xff; b[p + 9] = (uint8_t)(rnd() & 0x7f); } break;   // long varint
        case 4: if (!b.empty()) b[pos()] = (uint8_t)rnd(); break;
        case 5: { size_t p = pos(), k = rnd() % 64; if (p + k <= b.size()) b.erase(b.begin() + p, b.begin() + p + k); } break;
        default: { size_t p = pos(); std::vector<uint8_t> ins(rnd() % 32); for (auto& x : ins) x = (uint8_t)rnd(); b.insert(b.begin() + (p <= b.size() ? p : 0), ins.begin(), ins.end()); }
    }
}
static void check_graph(const GraphHost& g) {   // the invariants the device code relies on
    for (size_t i = 0; i < g.prog.size(); i++) {
        const VmInstr& in = g.prog[i];
        const uint32_t kind = in.kind_op & 0xff;
        if (kind == VM_INPUT && in.a >= g.n_slots) abort();
        if (kind == VM_CONST && (size_t)in.a * 32 + 32 > g.consts.size()) abort();
        if ((kind == VM_UNO || kind == VM_DUO || kind == VM_TRES) && in.a >= i) abort();
        if ((kind == VM_DUO || kind == VM_TRES) && in.b >= i) abort();
        if (kind == VM_TRES && in.c >= i) abort();
    }
    for (uint32_t s : g.signals) if (s >= g.prog.size()) abort();
    for (auto& kv : g.inputs) if ((uint64_t)kv.second.first + kv.second.second > g.n_slots) abort();
    uint32_t nb = 0;
    std::vector<VmRecord> recs = vm_build_schedule(g.prog, nb);
    if (recs.size() != (size_t)nb * VM_SLOTS) abort();
}
static void check_zkey(const ZkeyHost& z) {
    if (z.a_ptr.empty() || z.a_ptr.back() != z.a_col.size() || z.a_val.size() != 32 * z.a_col.size()) abort();
    if (z.b_ptr.empty() || z.b_ptr.back() != z.b_col.size() || z.b_val.size() != 32 * z.b_col.size()) abort();
    if (z.alpha_g1.size() != 64 || z.delta_g2.size() != 128 || z.a_query.size() % 64 || z.b_g2.size() % 128) abort();
}

static void put_varint(std::vector<uint8_t>& b, uint64_t v) {
    while (v >= 0x80) { b.push_back((uint8_t)(v | 0x80)); v >>= 7; }
    b.push_back((uint8_t)v);
}
// hand-made hostile graphs: each must be rejected with an exception
static int crafted_graphs(const std::vector<uint8_t>& graph) {
    int rejected = 0, total = 0;
    auto expect_reject = [&](const std::vector<uint8_t>& b) {
        total++;
        try { GraphHost g; parse_graph(b.data(), b.size(), g); check_graph(g); } catch (const std::exception&) { rejected++; }
    };
    size_t o = 14;
    uint64_t cnt;
    memcpy(&cnt, graph.data() + o, 8);
    o += 8;
    const size_t first_node = o;
    for (uint64_t i = 0; i < cnt; i++) { uint64_t l; if (!rd_varint(graph.data(), graph.size(), o, l)) abort(); o += l; }
    const size_t md_at = o;
    uint64_t mdlen;
    if (!rd_varint(graph.data(), graph.size(), o, mdlen)) abort();
    const std::vector<uint8_t> md(graph.begin() + o, graph.begin() + o + mdlen);
    for (uint64_t huge : {~0ull, ~0ull - 20, 1ull << 63}) {   // node / metadata lengths that wrap a 64-bit offset
        std::vector<uint8_t> b(graph.begin(), graph.begin() + first_node);
        put_varint(b, huge);
        b.insert(b.end(), graph.begin() + first_node + 1, graph.end());
        expect_reject(b);
        std::vector<uint8_t> c(graph.begin(), graph.begin() + md_at);
        put_varint(c, huge);
        c.insert(c.end(), md.begin(), md.end());
        expect_reject(c);
    }
    auto with_input = [&](const char* name, uint64_t off, uint64_t len) {   // a later map entry overrides the real one
        std::vector<uint8_t> sig, ent, nm;
        sig.push_back(0x08); put_varint(sig, off);
        sig.push_back(0x10); put_varint(sig, len);
        ent.push_back(0x0a); put_varint(ent, strlen(name)); ent.insert(ent.end(), name, name + strlen(name));
        ent.push_back(0x12); put_varint(ent, sig.size()); ent.insert(ent.end(), sig.begin(), sig.end());
        nm = md;
        nm.push_back(0x12); put_varint(nm, ent.size()); nm.insert(nm.end(), ent.begin(), ent.end());
        std::vector<uint8_t> b(graph.begin(), graph.begin() + md_at);
        put_varint(b, nm.size());
        b.insert(b.end(), nm.begin(), nm.end());
        return b;
    };
    expect_reject(with_input("pathElements", 40, 1000));
    expect_reject(with_input("x", 1ull << 32, 1));
    expect_reject(with_input("identitySecret", 0xffffffffull, 1));
    expect_reject(with_input("messageId", 5, 1ull << 40));
    {   // sanity of the builder itself: re-stating an existing entry unchanged still parses
        std::vector<uint8_t> b = with_input("x", 1, 1);
        GraphHost g; parse_graph(b.data(), b.size(), g); check_graph(g);
        if (g.inputs.at("x") != std::make_pair(1u, 1u)) abort();
    }
    printf("crafted graphs: %d of %d rejected\n", rejected, total);
    return rejected == total ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: fuzz_parsers zkey graph iters seed\n"); return 2; }
    const std::vector<uint8_t> zkey = slurp(argv[1]), graph = slurp(argv[2]);
    const long iters = atol(argv[3]);
    g_s = strtoull(argv[4], nullptr, 10);
    long ok_g = 0, ok_z = 0;
    {   // the unmodified files parse
        GraphHost g; parse_graph(graph.data(), graph.size(), g); check_graph(g);
        ZkeyHost z; parse_zkey(zkey.data(), zkey.size(), z); check_zkey(z);
    }
    if (crafted_graphs(graph)) return 1;
    for (long it = 0; it < iters; it++) {
        {
            std::vector<uint8_t> b = graph;
            for (int k = 0, m = 1 + rnd() % 3; k < m; k++) mutate(b, 64);
            try { GraphHost g; parse_graph(b.data(), b.size(), g); check_graph(g); ok_g++; } catch (const std::exception&) {}
        }
        if (it % 8 == 0) {   // the key is 3.4 MB: fewer rounds
            std::vector<uint8_t> b = zkey;
            for (int k = 0, m = 1 + rnd() % 3; k < m; k++) mutate(b, 1024);
            try { ZkeyHost z; parse_zkey(b.data(), b.size(), z); check_zkey(z); ok_z++; } catch (const std::exception&) {}
        }
    }
    printf("fuzz done: %ld iterations, %ld graphs and %ld keys still parsed\n", iters, ok_g, ok_z);
    return 0;
}
