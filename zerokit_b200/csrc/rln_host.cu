// Host side of librln_b200.so: the `Rln` object (mirror of rln::public::RLN, rln/src/public.rs:65-771)
// that owns the circuit, the HBM-resident Merkle tree and the proving workspace, and the C ABI
// declared in include/rln_b200.h.  The host only parses files, moves bytes and launches kernels;
// every field / curve operation of the hot path runs on the GPU (there is no CPU fallback — a missing
// device is an error).
#include <dlfcn.h>
#include <sys/stat.h>
#include <sys/types.h>

#include <cerrno>

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <mutex>
#include <random>
#include <chrono>
#include <sstream>
#include <thread>

#include "../../include/rln_b200.h"
#include "device_api.hpp"
#include "coalesce.hpp"
#include "host_util.hpp"
#include "pairing_constants.hpp"
#include "poseidon_constants.hpp"
#include "verify_vm_program.hpp"

namespace zk {

std::atomic<uint64_t> g_launch_count{0};

// ------------------------------------------------------------------------------------------- small RAII helpers
struct DevMem {
    void* p = nullptr;
    size_t bytes = 0;
    DevMem() {}
    DevMem(const DevMem&) = delete;
    DevMem& operator=(const DevMem&) = delete;
    ~DevMem() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    void alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        ZK_CUDA_CHECK(cudaMalloc(&p, n));
        bytes = n;
    }
    void ensure(size_t n) {
        if (n > bytes) alloc(n);
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
    void upload(const void* src, size_t n) {
        alloc(n);
        if (n) ZK_CUDA_CHECK(cudaMemcpy(p, src, n, cudaMemcpyHostToDevice));
    }
};

struct RlnError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// The __constant__ / __device__ tables (Poseidon round constants, pairing constants) exist once per device: a process that
// drives several GPUs (rlnb200_set_device, the multi-device prover) uploads them to each device the first time that device is
// used.  The host copies are computed once.
static std::mutex g_init_mu;
static std::string g_init_error;
static bool g_have_device = false, g_probed = false;
static std::unique_ptr<PoseidonTables> g_host_pt;
static std::unique_ptr<PairingTables> g_host_pair;
static std::vector<char> g_device_ready;
static int global_init() {   // returns the current device, initialised
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (!g_probed) {
        g_probed = true;
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            g_init_error = std::string("no usable CUDA device (librln_b200 has no CPU path): ") + cudaGetErrorString(e);
        } else {
            g_have_device = true;
            g_device_ready.assign((size_t)n, 0);
            g_host_pt = std::make_unique<PoseidonTables>();
            poseidon_fill_tables(*g_host_pt);
            g_host_pair = std::make_unique<PairingTables>();
            pairing_tables_init(*g_host_pair);
        }
    }
    if (!g_have_device) throw RlnError(g_init_error);
    int dev = 0;
    ZK_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || (size_t)dev >= g_device_ready.size()) throw RlnError("CUDA initialisation failed: current device out of range");
    if (!g_device_ready[dev]) {
        try {
            poseidon_upload_tables(*g_host_pt);
            pairing_upload_tables(*g_host_pair);
        } catch (const CudaError& ce) {
            throw RlnError(std::string("CUDA initialisation failed: ") + cudaGetErrorString(ce.code));
        }
        g_device_ready[dev] = 1;
    }
    return dev;
}
// cudaSetDevice is per host thread: every entry point that touches a handle's streams / buffers first makes the handle's
// device current on the calling thread, and puts the caller's device back afterwards
struct DeviceGuard {
    int want, prev = -1;
    explicit DeviceGuard(int device) : want(device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != want) ZK_CUDA_CHECK(cudaSetDevice(want));
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static std::vector<uint8_t> read_file_bytes(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw RlnError("I/O error: cannot open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static void random_fr(uint8_t out[32]) {
    static thread_local std::random_device rd;
    for (;;) {
        for (int i = 0; i < 8; i++) {
            uint32_t w = rd();
            memcpy(out + 4 * i, &w, 4);
        }
        out[31] &= 0x3f;
        if (fr_is_canonical(out)) return;
    }
}

// ------------------------------------------------------------------------------------------- witness record
// RLNWitnessInput (rln/src/protocol/witness.rs:52-60); message mode SingleV1 (k = 1) or MultiV1 (k = max_out)
struct Witness {
    uint8_t secret[32], limit[32], x[32], ext_null[32];
    bool multi = false;
    std::vector<uint8_t> mids;   // k × 32 message ids
    std::vector<uint8_t> sel;    // k selector_used flags (multi mode only)
    std::vector<uint8_t> path;   // depth × 32
    std::vector<uint8_t> index;  // depth
    size_t k() const { return mids.size() / 32; }
};
// RLNProofValues (rln/src/protocol/proof.rs:100-190)
struct ProofValues {
    uint8_t root[32], ext_null[32], x[32];
    bool multi = false;
    std::vector<uint8_t> ys, nulls;  // k × 32 each
    std::vector<uint8_t> sel;        // k (multi only)
    size_t k() const { return ys.size() / 32; }
};
struct RlnProof {
    uint8_t proof[128];  // ark-compressed A|B|C
    ProofValues pv;
};
struct PartialProofHost {  // PartialProof (rln/src/partial_proof.rs:30-43); the mask is a property of the circuit
    uint8_t affine[320];   // partial_pi_a 64 | partial_rho 64 | partial_pi_b 128 | partial_pi_c 64, canonical
    uint8_t comp[160];     // the same, ark-compressed 32 | 32 | 64 | 32
};

static void put_u64(std::vector<uint8_t>& b, uint64_t v) { b.insert(b.end(), (uint8_t*)&v, (uint8_t*)&v + 8); }
static std::string msg_read_len(size_t expected, size_t got) {
    std::ostringstream s;
    s << "Expected to read " << expected << " bytes but read " << got << " bytes";
    return s.str();
}
static std::string msg_mode(uint8_t b) {
    char t[64];
    snprintf(t, sizeof t, "Unknown message mode version byte: %#04x", b);
    return t;
}

// rln_witness_to_bytes_le (witness.rs:369-415)
static std::vector<uint8_t> witness_to_bytes(const Witness& w) {
    std::vector<uint8_t> b;
    b.push_back(w.multi ? 1 : 0);
    b.insert(b.end(), w.secret, w.secret + 32);
    b.insert(b.end(), w.limit, w.limit + 32);
    if (!w.multi) b.insert(b.end(), w.mids.begin(), w.mids.end());
    put_u64(b, w.path.size() / 32);
    b.insert(b.end(), w.path.begin(), w.path.end());
    put_u64(b, w.index.size());
    b.insert(b.end(), w.index.begin(), w.index.end());
    b.insert(b.end(), w.x, w.x + 32);
    b.insert(b.end(), w.ext_null, w.ext_null + 32);
    if (w.multi) {
        put_u64(b, w.k());
        b.insert(b.end(), w.mids.begin(), w.mids.end());
        put_u64(b, w.sel.size());
        b.insert(b.end(), w.sel.begin(), w.sel.end());
    }
    return b;
}
// RLNWitnessInput::new_single / new_multi (witness.rs:78-176)
static void validate_witness(const Witness& w) {
    if (is_zero32(w.limit)) throw RlnError("User message limit cannot be zero");
    if (w.path.size() / 32 != w.index.size()) {
        std::ostringstream s;
        s << "Merkle proof length mismatch: expected " << w.path.size() / 32 << ", got " << w.index.size();
        throw RlnError(s.str());
    }
    auto range_err = [&](const uint8_t* mid) {
        return RlnError("Message id (" + decimal_le32(mid) + ") is not within user_message_limit (" + decimal_le32(w.limit) + ")");
    };
    if (!w.multi) {
        if (cmp_le32(w.mids.data(), w.limit) >= 0) throw range_err(w.mids.data());
        return;
    }
    if (w.k() == 0) throw RlnError("The field message_ids must contain at least one message_id");
    if (w.sel.size() != w.k()) {
        std::ostringstream s;
        s << "The field message_ids has length " << w.k() << ", but the field selector_used has length " << w.sel.size();
        throw RlnError(s.str());
    }
    bool any = false;
    for (uint8_t v : w.sel) any = any || v;
    if (!any) throw RlnError("At least one selector_used value must be true");
    for (size_t i = 0; i < w.k(); i++)
        for (size_t j = 0; j < i; j++)
            if (w.sel[i] && w.sel[j] && !memcmp(&w.mids[32 * i], &w.mids[32 * j], 32)) throw RlnError("Duplicate message ID found in message_ids");
    for (size_t i = 0; i < w.k(); i++)
        if (w.sel[i] && cmp_le32(&w.mids[32 * i], w.limit) >= 0) throw range_err(&w.mids[32 * i]);
}
// bytes_le_to_rln_witness (witness.rs:470-560): returns bytes consumed
static size_t witness_from_bytes(const uint8_t* b, size_t len, Witness& w) {
    size_t o = 0;
    auto need = [&](size_t k) {
        if (o + k > len) throw RlnError(msg_read_len(o + k, len));
    };
    need(1);
    if (b[0] > 1) throw RlnError(msg_mode(b[0]));
    w.multi = b[0] == 1;
    o = 1;
    auto fr = [&](uint8_t* dst) {
        need(32);
        memcpy(dst, b + o, 32);
        if (!fr_is_canonical(dst)) throw RlnError("Non-canonical field element: value is not in [0, r-1]");
        o += 32;
    };
    auto vec_fr = [&](std::vector<uint8_t>& dst) {
        need(8);
        uint64_t n;
        memcpy(&n, b + o, 8);
        o += 8;
        if (n > (len - o) / 32) throw RlnError(msg_read_len(o + n * 32, len));
        dst.assign(b + o, b + o + 32 * n);
        for (uint64_t i = 0; i < n; i++)
            if (!fr_is_canonical(dst.data() + 32 * i)) throw RlnError("Non-canonical field element: value is not in [0, r-1]");
        o += 32 * n;
    };
    auto vec_u8 = [&](std::vector<uint8_t>& dst) {
        need(8);
        uint64_t n;
        memcpy(&n, b + o, 8);
        o += 8;
        if (n > len - o) throw RlnError(msg_read_len(o + n, len));
        dst.assign(b + o, b + o + n);
        o += n;
    };
    fr(w.secret);
    fr(w.limit);
    if (!w.multi) {
        w.mids.resize(32);
        fr(w.mids.data());
    }
    vec_fr(w.path);
    vec_u8(w.index);
    fr(w.x);
    fr(w.ext_null);
    if (w.multi) {
        vec_fr(w.mids);
        vec_u8(w.sel);
        for (auto& v : w.sel) v = v != 0;  // bytes_le_to_vec_bool: any non-zero byte is true (utils.rs:407-410)
    }
    validate_witness(w);
    return o;
}
// rln_proof_values_to_bytes_le (proof.rs:192-236): version | root | external_nullifier | x | y | nullifier, or for MultiV1
// version | root | external_nullifier | x | vec ys | vec nullifiers | vec bool selector_used
static std::vector<uint8_t> proof_values_to_bytes(const ProofValues& pv) {
    std::vector<uint8_t> b;
    b.push_back(pv.multi ? 1 : 0);
    b.insert(b.end(), pv.root, pv.root + 32);
    b.insert(b.end(), pv.ext_null, pv.ext_null + 32);
    b.insert(b.end(), pv.x, pv.x + 32);
    if (!pv.multi) {
        b.insert(b.end(), pv.ys.begin(), pv.ys.end());
        b.insert(b.end(), pv.nulls.begin(), pv.nulls.end());
    } else {
        put_u64(b, pv.k());
        b.insert(b.end(), pv.ys.begin(), pv.ys.end());
        put_u64(b, pv.k());
        b.insert(b.end(), pv.nulls.begin(), pv.nulls.end());
        put_u64(b, pv.sel.size());
        b.insert(b.end(), pv.sel.begin(), pv.sel.end());
    }
    return b;
}
static size_t proof_values_from_bytes(const uint8_t* b, size_t len, ProofValues& pv) {
    if (len < 1) throw RlnError(msg_read_len(1, 0));
    if (b[0] > 1) throw RlnError(msg_mode(b[0]));
    pv.multi = b[0] == 1;
    size_t o = 1;
    auto fr = [&](uint8_t* dst) {
        if (o + 32 > len) throw RlnError("RLN utility error: Input data too short: expected at least 32 bytes, got " + std::to_string(len - o) + " bytes");
        memcpy(dst, b + o, 32);
        if (!fr_is_canonical(dst)) throw RlnError("Non-canonical field element: value is not in [0, r-1]");
        o += 32;
    };
    auto vec_fr = [&](std::vector<uint8_t>& dst) {
        if (o + 8 > len) throw RlnError("RLN utility error: Input data too short: expected at least 8 bytes, got " + std::to_string(len - o) + " bytes");
        uint64_t n;
        memcpy(&n, b + o, 8);
        o += 8;
        if (n > (len - o) / 32) throw RlnError("RLN utility error: Input data too short: expected at least " + std::to_string(8 + n * 32) + " bytes, got " + std::to_string(len - o + 8) + " bytes");
        dst.assign(b + o, b + o + 32 * n);
        for (uint64_t i = 0; i < n; i++)
            if (!fr_is_canonical(dst.data() + 32 * i)) throw RlnError("Non-canonical field element: value is not in [0, r-1]");
        o += 32 * n;
    };
    fr(pv.root);
    fr(pv.ext_null);
    fr(pv.x);
    if (!pv.multi) {
        pv.ys.resize(32);
        pv.nulls.resize(32);
        fr(pv.ys.data());
        fr(pv.nulls.data());
    } else {
        vec_fr(pv.ys);
        vec_fr(pv.nulls);
        if (o + 8 > len) throw RlnError("RLN utility error: Input data too short: expected at least 8 bytes, got " + std::to_string(len - o) + " bytes");
        uint64_t n;
        memcpy(&n, b + o, 8);
        o += 8;
        if (n > len - o) throw RlnError("RLN utility error: Input data too short: expected at least " + std::to_string(8 + n) + " bytes, got " + std::to_string(len - o + 8) + " bytes");
        pv.sel.assign(b + o, b + o + n);
        for (auto& v : pv.sel) v = v != 0;
        o += n;
        if (pv.sel.size() != pv.k()) throw RlnError("The field ys has length " + std::to_string(pv.k()) + ", but the field selector_used has length " + std::to_string(pv.sel.size()));
        if (pv.nulls.size() != pv.ys.size()) throw RlnError("The field ys has length " + std::to_string(pv.k()) + ", but the field nullifiers has length " + std::to_string(pv.nulls.size() / 32));
    }
    return o;
}
// rln_proof_to_bytes_le (proof.rs:413-428): version | proof(128) | proof_values
static std::vector<uint8_t> rln_proof_to_bytes(const RlnProof& p) {
    std::vector<uint8_t> b;
    b.push_back(p.pv.multi ? 1 : 0);
    b.insert(b.end(), p.proof, p.proof + 128);
    std::vector<uint8_t> v = proof_values_to_bytes(p.pv);
    b.insert(b.end(), v.begin(), v.end());
    return b;
}
// public inputs in circuit order (proof.rs:863-884): single [y, root, nullifier, x, en]; multi [ys…, root, nullifiers…, x, en, selectors…]
static std::vector<uint8_t> public_inputs(const ProofValues& pv) {
    std::vector<uint8_t> b(pv.ys);
    b.insert(b.end(), pv.root, pv.root + 32);
    b.insert(b.end(), pv.nulls.begin(), pv.nulls.end());
    b.insert(b.end(), pv.x, pv.x + 32);
    b.insert(b.end(), pv.ext_null, pv.ext_null + 32);
    if (pv.multi)
        for (uint8_t v : pv.sel) {
            uint8_t fr[32] = {0};
            fr[0] = v ? 1 : 0;
            b.insert(b.end(), fr, fr + 32);
        }
    return b;
}

// ------------------------------------------------------------------------------------------- the RLN object
class Rln {
   public:
    Rln(size_t tree_depth, const uint8_t* zkey, size_t zlen, const uint8_t* graph, size_t glen);
    ~Rln();

    int device() const { return device_; }
    size_t depth() const { return depth_; }
    size_t tree_depth() const { return tree_depth_; }
    size_t max_out() const { return max_out_; }
    bool multi() const { return multi_; }
    size_t values_stride() const { return 32 * (3 + 2 * max_out_); }  // root | ext_null | x | ys[k] | nullifiers[k]
    size_t n_public() const { return vk_.n_public; }
    uint32_t n_slots() const { return gh_.n_slots; }
    uint32_t n_wires() const { return (uint32_t)gh_.signals.size(); }
    uint32_t domain() const { return domain_; }
    const GraphHost& graph() const { return gh_; }

    // tree: one HBM heap tree; the bookkeeping (next_index, "is set" flags, override_range, error kinds) follows the
    // reference flavour the handle was built as — PmTree is the reference's default PoseidonTree (rln/src/pm_tree_adapter.rs),
    // FullMerkleTree / OptimalMerkleTree (utils/src/merkle_tree/*.rs) are selectable through the V3 constructors
    enum TreeKind { TREE_PM = 0, TREE_FULL = 1, TREE_OPTIMAL = 2 };
    void set_tree_kind(TreeKind k) { tree_kind_ = k; }
    TreeKind tree_kind() const { return tree_kind_; }
    RlnError oob_error() const {   // pmtree TreeErrorKind::IndexOutOfBounds vs ZerokitMerkleTreeError::InvalidLeaf
        return RlnError(tree_kind_ == TREE_PM ? "Merkle tree error: Pmtree error: Tree error: Index out of bounds" : "Merkle tree error: Leaf index out of bounds");
    }
    void set_tree(size_t depth);
    void set_range_host(size_t start, const uint8_t* leaves, size_t count);
    void set_range_device(size_t start, const uint8_t* d_leaves, size_t count, cudaStream_t s);
    void override_range(size_t start, const uint8_t* leaves, size_t n_leaves, std::vector<size_t> indices);
    void set_leaf(size_t index, const uint8_t* leaf);
    void delete_leaf(size_t index);
    void set_next(const uint8_t* leaf);
    void get_leaf(size_t index, uint8_t out[32]);
    void root(uint8_t out[32]);
    size_t leaves_set() const { return next_index_; }
    size_t capacity() const { return (size_t)1 << tree_depth_; }
    void merkle_proofs(const uint64_t* idx, size_t n, uint8_t* elems, uint8_t* bits);

    // proving
    void reserve(size_t B);
    void prove_device(const uint8_t* d_inputs, const uint8_t* d_rs, size_t n, uint8_t* d_proofs, uint8_t* d_values, uint8_t* d_affine,
                      cudaStream_t s, int phase = MSM_FULL, const uint8_t* d_partial = nullptr, uint8_t* d_partial_affine = nullptr,
                      uint8_t* d_partial_comp = nullptr);
    void prove_host(const std::vector<Witness>& ws, const uint8_t* rs, std::vector<RlnProof>& out, const PartialProofHost* partials = nullptr);
    // the batch path on wire records: n rln_witness_to_bytes_le records in, n rln_proof_to_bytes_le records
    // out, both resident in HBM (k_records.cu parses / formats them on the device); d_rs may be null (fresh r, s per proof)
    void prove_records_device(const uint8_t* d_records, const uint8_t* d_rs, size_t n, uint8_t* d_proof_records, cudaStream_t s);
    // the same from / to host memory: one upload, the device path, one download per super-chunk
    void prove_records_host(const uint8_t* records, const uint8_t* rs, size_t n, uint8_t* proof_records);
    // record lengths of rln_witness_to_bytes_le / rln_proof_to_bytes_le for this circuit (witness.rs:369-415, proof.rs:192-236,413-428)
    size_t witness_record_len() const {
        const size_t d = depth_, k = max_out_;
        return multi_ ? 1 + 64 + (8 + 32 * d) + (8 + d) + 64 + (8 + 32 * k) + (8 + k) : 1 + 32 * (5 + d) + 16 + d;
    }
    size_t proof_record_len() const {
        const size_t k = max_out_;
        return multi_ ? 1 + 128 + 1 + 96 + (8 + 32 * k) * 2 + (8 + k) : 290;
    }
    RecordLayout record_layout() const { return RecordLayout{slots_, (u32)witness_record_len(), (u32)proof_record_len()}; }
    // two-phase proving (rln/src/protocol/proof.rs:783-849): the unknown inputs of `ws` (message_id, x, external_nullifier) are ignored
    void partial_host(const std::vector<Witness>& ws, std::vector<PartialProofHost>& out);
    const std::vector<uint8_t>& partial_mask() const { return mask_; }  // one byte per wire 1..n_wires-1 (1 = known)
    void decompress_partial(const uint8_t comp[160], uint8_t affine[320]);
    void witness_slots(const Witness& w, uint8_t* slots) const;
    void check_witness_shape(const Witness& w) const;
    void debug_w_h(const Witness& w, uint8_t* w_out, uint8_t* h_out);
    void table_info(int* c, int* K, uint64_t* g1, uint64_t* g2, uint64_t* bytes, int* c2, int* K2) const {
        *c = plan_.c; *K = plan_.K; *c2 = plan_.c2; *K2 = plan_.K2;
        *g1 = (uint64_t)plan_.g1[0].n_bases + plan_.g1[1].n_bases + plan_.g1[2].n_bases + plan_.g1[3].n_bases;
        *g2 = plan_.g2.n_bases;
        *bytes = 0;
        for (int i = 0; i < 5; i++) *bytes += d_tab_[i].bytes;
    }
    bool glv() const { return plan_.glv != 0; }
    // cached_leaves_indices (utils/src/merkle_tree/full_merkle_tree.rs:40-45,186-194): 1 = set, 0 = deleted / never set
    std::vector<uint8_t> leaf_set_;
    void mark_leaves(size_t start, size_t count, uint8_t v) {
        if (start + count > leaf_set_.size()) leaf_set_.resize(start + count, 0);
        std::fill(leaf_set_.begin() + start, leaf_set_.begin() + start + count, v);
    }
    std::vector<size_t> empty_leaves_indices() const {
        std::vector<size_t> out;
        for (size_t i = 0; i < next_index_; i++)
            if (i >= leaf_set_.size() || !leaf_set_[i]) out.push_back(i);
        return out;
    }
    // get_subtree_root (full_merkle_tree.rs:157-184): the ancestor at `level` (0 = root, depth = the leaf itself) of leaf `index`
    void subtree_root(size_t level, size_t index, uint8_t out[32]) {
        if (level > tree_depth_) throw RlnError("Merkle tree error: Invalid index");
        if (index >= capacity()) throw RlnError("Merkle tree error: Invalid leaf");
        const size_t node = (capacity() + index) >> (tree_depth_ - level);
        DevMem tmp;
        tmp.alloc(32);
        launch_fr_to_bytes(d_nodes_.as<Fr>() + node, tmp.as<uint8_t>(), 1, stream_);
        g_launch_count++;
        ZK_CUDA_CHECK(cudaMemcpyAsync(out, tmp.p, 32, cudaMemcpyDeviceToHost, stream_));
        ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
    }
    std::vector<uint8_t> metadata;   // set_metadata / get_metadata (rln/src/public.rs:499-515): opaque bytes kept beside the tree
    void set_metadata(const uint8_t* p, size_t n) { metadata.assign(p, p + n); store_dirty_ = true; }
    void sync() { ZK_CUDA_CHECK(cudaStreamSynchronize(stream_)); }
    // The on-disk side of the reference's default tree (PmTree over sled: rln/src/pm_tree_adapter.rs:194-239 reload-or-create,
    // :393-408 metadata, utils/src/pm_tree/sled_adapter.rs:38-103).  The tree itself lives in HBM; a persistent configuration
    // (temporary = false, a path) adds one file under that path — leaves below next_index, next_index, metadata — written by
    // flush() and when the handle is dropped (sled flushes on close / drop), and read back by the next handle opened on the
    // same path, which rebuilds the tree on the GPU (11.6 ms for 2^20 leaves).
    void attach_store(const TreeConfig& cfg);
    void flush_store();
    bool persistent() const { return !store_dir_.empty(); }
    const uint8_t* ext_wires_ = nullptr;   // device pointer: B × n_wires canonical values that replace the graph evaluation
    // generate_rln_proof_with_witness (rln/src/public.rs:643-658): wires = n_wires × 32 canonical bytes calculated by the caller
    void prove_with_wires(const Witness& w, const std::vector<uint8_t>& wires, const uint8_t* rs, RlnProof& out) {
        if (wires.size() != 32 * n_wires())
            throw RlnError("Protocol error: the calculated witness has " + std::to_string(wires.size() / 32) + " elements, the circuit has " +
                           std::to_string(n_wires()) + " wires");
        DevMem d;
        d.upload(wires.data(), wires.size());
        ext_wires_ = d.as<uint8_t>();
        std::vector<RlnProof> o;
        try {
            prove_host(std::vector<Witness>(1, w), rs, o);
        } catch (...) {
            ext_wires_ = nullptr;
            throw;
        }
        ext_wires_ = nullptr;
        out = o[0];
    }
    void verify_batch(const uint8_t* proofs128, const uint8_t* publics_circuit_order, size_t n, uint8_t* ok);
    // one proof through the lane-parallel kernel with a clock sample per level: cycles[levels + 1], meta[levels]
    void verify_vm_trace(const uint8_t* proof128, const uint8_t* publics, long long* cycles, uint32_t* meta, uint8_t* ok) {
        DevMem dp, dv, dok, dt;
        dp.upload(proof128, 128);
        dv.upload(publics, 32 * (size_t)vk_.n_public);
        dok.alloc(1);
        dt.alloc(sizeof(long long) * (vm_.n_levels + 1));
        ZK_CUDA_CHECK(cudaMemset(dt.p, 0, sizeof(long long) * (vm_.n_levels + 1)));
        VerifyVmDev t = vm_;
        t.trace = dt.as<long long>();
        launch_verify_vm(t, vk_, dp.as<uint8_t>(), dv.as<uint8_t>(), 1, dok.as<uint8_t>(), stream_);
        ZK_CUDA_CHECK(cudaMemcpyAsync(cycles, dt.p, sizeof(long long) * (vm_.n_levels + 1), cudaMemcpyDeviceToHost, stream_));
        ZK_CUDA_CHECK(cudaMemcpyAsync(ok, dok.p, 1, cudaMemcpyDeviceToHost, stream_));
        ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
        memcpy(meta, vm_meta_.data(), 4 * vm_meta_.size());
    }
    double verify_vm_est_cycles() const { return vm_est_cycles_; }
    size_t n_public_inputs() const { return vk_.n_public; }
    bool set_verify_vm_max(size_t n) { if (n && !vm_.code) return false; vm_max_batch_ = n; return true; }
    void verify_vm_info(uint32_t* levels, uint32_t* slots, uint32_t* constants) const { *levels = vm_.n_levels; *slots = vm_.n_slots; *constants = vm_.n_const; }
    // witness, qap, g1 accumulate, g1 reduce, g2 accumulate, g2 reduce, assemble, proof values
    float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u32 last_chunks = 1;   // device batches (kernel launches per stage) the stage_ms above are summed over
    std::mutex mu;
    // concurrent single-item calls on one handle are run as batches (coalesce.hpp; RLN_B200_COALESCE=1)
    struct ProveReq { const Witness* w; const uint8_t* rs; RlnProof out; std::string err; bool failed = false; bool done = false; };
    struct PairingReq { const uint8_t* proof128; const uint8_t* pub; uint8_t ok = 0; std::string err; bool failed = false; bool done = false; };
    Coalescer<ProveReq> co_prove;
    Coalescer<PairingReq> co_pairing;
    size_t max_batch() const { return max_batch_; }

   private:
    struct TaskSet {
        DevMem g1, g2;
        u32 n1 = 0, n2 = 0, n_ab = 0;   // n_ab: tasks of the groups A and B₁ (they come first)
    };
    TaskSet& tasks_for(u32 B, int phase);
    void compute_known_mask();
    void build_circuit();
    void build_tables();
    void check_graph_shape();

    int device_ = 0;         // the CUDA device that owns every buffer, stream and event of this handle
    ZkeyHost zk_;
    GraphHost gh_;
    size_t depth_ = 0;       // circuit tree depth (len of pathElements)
    size_t max_out_ = 1;     // message-id slots per proof (rln/src/circuit/mod.rs:181-194)
    bool multi_ = false;
    size_t tree_depth_ = 0;  // depth of the stateful tree
    uint32_t domain_ = 0, log_domain_ = 0;
    InputSlots slots_{};

    // device-resident circuit
    DevMem d_prog_, d_consts_, d_signals_, d_a_ptr_, d_a_col_, d_a_val_, d_b_ptr_, d_b_col_, d_b_val_, d_tw_inv_, d_tw_fwd_, d_coset_, d_tw_tile_dif_, d_tw_tile_dit_, d_tw_mid_;
    CircuitDev circ_{};
    // fixed-base tables
    DevMem d_sched_;
    DevMem d_tab_[5], d_rows_[5], d_gamma_abc_, d_gamma_tab_, d_delta1_tab_, d_delta2_tab_, d_alpha1_tab_, d_beta1_tab_, d_vk_pre_, d_vm_code_, d_vm_consts_, d_vfy_in_, d_vfy_ok_;
    VerifyVmDev vm_{};
    DevMem ws_fold_s_, ws_fold_r_, ws_fold_part_, ws_fold_sum_;   // folded assembly of ≤ FOLD_MAX full proofs
    static constexpr size_t FOLD_MAX = 16;          // measured: the fold wins up to 16 proofs (1: 7.46 → 6.88 ms, 32: 10.47 → 10.68)
    static constexpr size_t SMALL_MAX = 32;         // batches that take the latency-oriented small-batch kernels
    bool fold_on_ = true;
    std::vector<uint32_t> vm_meta_;
    double vm_est_cycles_ = 0;
    size_t vm_max_batch_ = 0;   // verify_batch takes the lane-parallel kernel up to this many proofs (0: never)
    FixedMsmPlan plan_{};
    ProverKeyDev pk_{};
    VerifyKeyDev vk_{};
    // tree
    DevMem d_nodes_;
    size_t next_index_ = 0;
    TreeKind tree_kind_ = TREE_PM;
    std::string store_dir_;
    bool store_dirty_ = false;
    void override_range_dense(size_t start, const uint8_t* leaves, size_t n_leaves, const std::vector<size_t>& indices);
    void download_leaves(size_t first, size_t count, uint8_t* out);
    // workspace
    size_t cap_ = 0, max_batch_ = 4096;
    DevMem ws_inputs_, ws_rs_, ws_vals_, ws_a_, ws_b_, ws_c_, ws_err_, ws_part1_, ws_part2_, ws_sum1_, ws_sum2_, ws_proofs_, ws_values_, ws_affine_;
    std::map<u64, std::unique_ptr<TaskSet>> tasks_;
    std::vector<uint8_t> wire_known_, mask_;
    DevMem ws_partial_, ws_partial_comp_, ws_bad_, ws_recs_in_, ws_recs_out_, ws_rs_all_;
    cudaStream_t stream_ = nullptr, side_ = nullptr, asm_side_ = nullptr;
    cudaEvent_t fork_ = nullptr, join_ = nullptr, asm_fork_ = nullptr, asm_join_ = nullptr;
    cudaEvent_t ev_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t mev_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void destroy_handles();
};

static void upload_points_g1(const std::vector<uint8_t>& raw, const std::vector<uint32_t>& pick, DevMem& out) {
    std::vector<uint8_t> sel(pick.size() * 64);
    for (size_t i = 0; i < pick.size(); i++) memcpy(sel.data() + 64 * i, raw.data() + 64 * (size_t)pick[i], 64);
    DevMem tmp;
    tmp.upload(sel.data(), sel.size());
    out.alloc(sizeof(G1Affine) * pick.size());
    launch_g1_from_bytes(tmp.as<uint8_t>(), out.as<G1Affine>(), pick.size(), 0);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaDeviceSynchronize());
}
static void upload_points_g2(const std::vector<uint8_t>& raw, const std::vector<uint32_t>& pick, DevMem& out) {
    std::vector<uint8_t> sel(pick.size() * 128);
    for (size_t i = 0; i < pick.size(); i++) memcpy(sel.data() + 128 * i, raw.data() + 128 * (size_t)pick[i], 128);
    DevMem tmp;
    tmp.upload(sel.data(), sel.size());
    out.alloc(sizeof(G2Affine) * pick.size());
    launch_g2_from_bytes(tmp.as<uint8_t>(), out.as<G2Affine>(), pick.size(), 0);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaDeviceSynchronize());
}
static G1Affine fetch_g1(const std::vector<uint8_t>& raw64) {
    DevMem d;
    std::vector<uint32_t> one = {0};
    upload_points_g1(raw64, one, d);
    G1Affine h;
    ZK_CUDA_CHECK(cudaMemcpy(&h, d.p, sizeof h, cudaMemcpyDeviceToHost));
    return h;
}
static G2Affine fetch_g2(const std::vector<uint8_t>& raw128) {
    DevMem d;
    std::vector<uint32_t> one = {0};
    upload_points_g2(raw128, one, d);
    G2Affine h;
    ZK_CUDA_CHECK(cudaMemcpy(&h, d.p, sizeof h, cudaMemcpyDeviceToHost));
    return h;
}

Rln::Rln(size_t tree_depth, const uint8_t* zkey, size_t zlen, const uint8_t* graph, size_t glen) {
    device_ = global_init();
    try {
        parse_zkey(zkey, zlen, zk_);
    } catch (const std::exception& e) {
        throw RlnError(std::string("ZKey error: ") + e.what());
    }
    try {
        parse_graph(graph, glen, gh_);
    } catch (const std::exception& e) {
        throw RlnError(std::string("Graph error: ") + e.what());
    }
    check_graph_shape();
    compute_known_mask();   // on the graph as the reference sees it
    if (env_int("RLN_B200_WITNESS_REASSOC", 1)) {
        // the witness program with its sums and products re-associated by readiness (host_util.hpp vm_optimize_program): same wires,
        // a dependency chain a third shorter — k_witness is paced by that chain
        VmOptimized o = vm_optimize_program(gh_.prog, gh_.consts, gh_.signals);
        gh_.prog.swap(o.prog);
        gh_.consts.swap(o.consts);
        gh_.signals.swap(o.signals);
    }
    {
        const int mb = env_int("RLN_B200_MAX_BATCH", 4096);
        if (mb < 1 || mb > (1 << 20)) throw RlnError("Configuration error: RLN_B200_MAX_BATCH must be in [1, 1048576]");
        max_batch_ = (size_t)mb;
        fold_on_ = env_int("RLN_B200_ASSEMBLE_FOLD", 1) != 0;
    }
    try {
        ZK_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        ZK_CUDA_CHECK(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking));
        ZK_CUDA_CHECK(cudaStreamCreateWithFlags(&asm_side_, cudaStreamNonBlocking));
        ZK_CUDA_CHECK(cudaEventCreateWithFlags(&fork_, cudaEventDisableTiming));
        ZK_CUDA_CHECK(cudaEventCreateWithFlags(&join_, cudaEventDisableTiming));
        ZK_CUDA_CHECK(cudaEventCreateWithFlags(&asm_fork_, cudaEventDisableTiming));
        ZK_CUDA_CHECK(cudaEventCreateWithFlags(&asm_join_, cudaEventDisableTiming));
        for (auto& e : ev_) ZK_CUDA_CHECK(cudaEventCreate(&e));
        for (auto& e : mev_) ZK_CUDA_CHECK(cudaEventCreate(&e));
        build_circuit();
        build_tables();
        set_tree(tree_depth);
    } catch (...) {   // a throwing constructor never runs the destructor: release the streams / events created so far
        destroy_handles();
        throw;
    }
}
void Rln::destroy_handles() {
    for (auto& e : ev_) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : mev_) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (cudaEvent_t* e : {&fork_, &join_, &asm_fork_, &asm_join_}) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
    for (cudaStream_t* st : {&stream_, &side_, &asm_side_}) if (*st) { cudaStreamDestroy(*st); *st = nullptr; }
}
Rln::~Rln() {
    int prev = -1;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device_) cudaSetDevice(device_);
    struct Back { int p, d; ~Back() { if (p >= 0 && p != d) cudaSetDevice(p); } } back{prev, device_};
    try {
        if (persistent()) flush_store();
    } catch (...) {   // a destructor cannot report: the previous flush() is what survives
    }
    cudaDeviceSynchronize();
    destroy_handles();
}

// Which wires a partial witness determines (graph.rs:274-312 evaluate_partial): a node is known iff all of its operands
// are; the unknown inputs are messageId, x and externalNullifier (witness.rs:887-931).
void Rln::compute_known_mask() {
    std::vector<uint8_t> slot_unknown(gh_.n_slots, 0);
    for (const char* name : {"messageId", "x", "externalNullifier", "selectorUsed"}) {
        auto it = gh_.inputs.find(name);
        if (it == gh_.inputs.end()) continue;
        for (uint32_t i = 0; i < it->second.second; i++) slot_unknown[it->second.first + i] = 1;
    }
    std::vector<uint8_t> known(gh_.prog.size());
    for (size_t i = 0; i < gh_.prog.size(); i++) {
        const VmInstr& in = gh_.prog[i];
        switch (in.kind_op & 0xff) {
            case VM_CONST: known[i] = 1; break;
            case VM_INPUT: known[i] = !slot_unknown[in.a]; break;
            case VM_UNO: known[i] = known[in.a]; break;
            case VM_DUO: known[i] = known[in.a] && known[in.b]; break;
            default: known[i] = known[in.a] && known[in.b] && known[in.c]; break;
        }
    }
    wire_known_.resize(gh_.signals.size());
    for (size_t i = 0; i < gh_.signals.size(); i++) wire_known_[i] = known[gh_.signals[i]];
    if (!wire_known_[0]) throw RlnError("Graph error: the constant wire depends on an unknown input");
    mask_.assign(wire_known_.begin() + 1, wire_known_.end());
}

void Rln::check_graph_shape() {
    auto need = [&](const char* name) -> std::pair<uint32_t, uint32_t> {
        auto it = gh_.inputs.find(name);
        if (it == gh_.inputs.end()) throw RlnError(std::string("Graph error: missing input signal ") + name);
        return it->second;
    };
    auto pe = need("pathElements"), pi = need("identityPathIndex");
    auto mid = need("messageId");
    multi_ = gh_.inputs.count("selectorUsed") != 0;
    max_out_ = multi_ ? mid.second : 1;
    if (mid.second != max_out_ || max_out_ == 0 || max_out_ > 16) throw RlnError("Graph error: unsupported number of messageId slots");
    if (multi_ && need("selectorUsed").second != max_out_) throw RlnError("Graph error: selectorUsed / messageId length mismatch");
    slots_.selector = multi_ ? need("selectorUsed").first : 0;
    slots_.max_out = (u32)max_out_;
    slots_.multi = multi_ ? 1 : 0;
    depth_ = pe.second;
    if (pi.second != depth_) throw RlnError("Graph error: pathElements / identityPathIndex length mismatch");
    auto scalar = [&](const char* name) {
        auto s = need(name);
        if (s.second != 1) throw RlnError(std::string("Graph error: input signal ") + name + " is not a single field element");
        return s.first;
    };
    slots_.secret = scalar("identitySecret");
    slots_.limit = scalar("userMessageLimit");
    slots_.message_id = mid.first;
    slots_.path = pe.first;
    slots_.index = pi.first;
    slots_.x = scalar("x");
    slots_.ext_null = scalar("externalNullifier");
    slots_.depth = (u32)depth_;
    slots_.n_slots = gh_.n_slots;
    const size_t nw = gh_.signals.size();
    if (nw == 0) throw RlnError("Graph error: no witness signals");
    if (depth_ == 0 || depth_ > 32) throw RlnError("Graph error: unsupported tree depth");
    if (zk_.gamma_abc.size() / 64 != zk_.num_instance) throw RlnError("ZKey error: gamma_abc size does not match the number of public inputs");
    if (zk_.a_query.size() / 64 != nw || zk_.b_g1.size() / 64 != nw || zk_.b_g2.size() / 128 != nw)
        throw RlnError("ZKey error: query sizes do not match the witness graph");
    if (zk_.l_query.size() / 64 + zk_.num_instance != nw) throw RlnError("ZKey error: l_query size does not match the witness graph");
    if (zk_.num_instance != 1 + 3 + 2 * max_out_ + (multi_ ? max_out_ : 0)) throw RlnError("ZKey error: number of public inputs does not match the message mode of the graph");
    for (auto c : zk_.a_col)
        if (c >= nw) throw RlnError("ZKey error: matrix column out of range");
    for (auto c : zk_.b_col)
        if (c >= nw) throw RlnError("ZKey error: matrix column out of range");
    size_t n = 1;
    log_domain_ = 0;
    while (n < zk_.num_constraints + zk_.num_instance) { n <<= 1; log_domain_++; }
    domain_ = (uint32_t)n;
    if (zk_.h_query.size() / 64 < domain_) throw RlnError("ZKey error: h_query shorter than the evaluation domain");
    if (log_domain_ > 16 || log_domain_ < 2) throw RlnError("ZKey error: unsupported evaluation domain size");
}

void Rln::build_circuit() {
    d_prog_.upload(gh_.prog.data(), gh_.prog.size() * sizeof(VmInstr));
    {   // constants → Montgomery on the device
        DevMem raw;
        raw.upload(gh_.consts.data(), gh_.consts.size());
        d_consts_.alloc(sizeof(Fr) * (gh_.consts.size() / 32 + 1));
        launch_fr_from_bytes(raw.as<uint8_t>(), d_consts_.as<Fr>(), gh_.consts.size() / 32, 0);
        g_launch_count++;
    }
    d_signals_.upload(gh_.signals.data(), gh_.signals.size() * 4);
    auto up_mat = [&](std::vector<uint32_t>& ptr, std::vector<uint32_t>& col, std::vector<uint8_t>& val, DevMem& dptr, DevMem& dcol, DevMem& dval) {
        dptr.upload(ptr.data(), ptr.size() * 4);
        dcol.upload(col.data(), col.size() * 4);
        DevMem raw;
        raw.upload(val.data(), val.size());
        dval.alloc(sizeof(Fr) * (val.size() / 32 + 1));
        launch_fr_from_bytes(raw.as<uint8_t>(), dval.as<Fr>(), val.size() / 32, 0);
        g_launch_count++;
        ZK_CUDA_CHECK(cudaDeviceSynchronize());
    };
    up_mat(zk_.a_ptr, zk_.a_col, zk_.a_val, d_a_ptr_, d_a_col_, d_a_val_);
    up_mat(zk_.b_ptr, zk_.b_col, zk_.b_val, d_b_ptr_, d_b_col_, d_b_val_);
    // NTT tables (ark-poly radix-2 domain: generator 5, two-adicity 28).  One-off host arithmetic.
    {
        const u32 n = domain_;
        u32 e[8];  // (r−1) >> 28
        for (int i = 0; i < 8; i++) e[i] = FrCfg::p(i);
        e[0] -= 1;
        for (int sft = 0; sft < 28; sft++) {
            for (int i = 0; i < 7; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 31);
            e[7] >>= 1;
        }
        Fr root28 = Fr::from_u32(5).pow(e);  // primitive 2^28-th root of unity
        auto root_for = [&](u32 lg) {
            Fr w = root28;
            for (u32 i = 0; i < 28 - lg; i++) w = w.sqr();
            return w;
        };
        const Fr om = root_for(log_domain_), om_inv = om.inv(), g = root_for(log_domain_ + 1);
        std::vector<Fr> fwd(n / 2), inv(n / 2), coset(n);
        Fr a = Fr::one(), b = Fr::one();
        for (u32 k = 0; k < n / 2; k++) {
            fwd[k] = a;
            inv[k] = b;
            a = a * om;
            b = b * om_inv;
        }
        const Fr n_inv = Fr::from_u32(n).inv();
        std::vector<Fr> gp(n);
        Fr t = n_inv;
        for (u32 i = 0; i < n; i++) { gp[i] = t; t = t * g; }
        for (u32 p = 0; p < n; p++) {
            u32 r = 0;
            for (u32 bit = 0; bit < log_domain_; bit++) r |= ((p >> bit) & 1) << (log_domain_ - 1 - bit);
            coset[p] = gp[r];
        }
        d_tw_fwd_.upload(fwd.data(), fwd.size() * sizeof(Fr));
        d_tw_inv_.upload(inv.data(), inv.size() * sizeof(Fr));
        d_coset_.upload(coset.data(), coset.size() * sizeof(Fr));
        if (log_domain_ >= 8 && log_domain_ <= 13) {   // tile-major twiddles of the tiled transforms (k_prover.cu k_ntt_outer / k_ntt_middle)
            const u32 MID = 64, R = n / MID;
            std::vector<Fr> t_dif((size_t)MID * R, Fr::one()), t_dit((size_t)MID * R, Fr::one()), mid(2 * MID, Fr::one());
            for (u32 i0 = 0; i0 < MID; i0++)
                for (u32 lh = 1; lh <= R / 2; lh <<= 1)
                    for (u32 kj = 0; kj < lh; kj++) {
                        const u32 e = (i0 + MID * kj) * (R / (2 * lh));   // exponent < n/2: j_global · stride of that stage
                        t_dif[(size_t)i0 * R + R - 2 * lh + kj] = inv[e];
                        t_dit[(size_t)i0 * R + R - 2 * lh + kj] = fwd[e];
                    }
            for (u32 h = 1; h <= MID / 2; h <<= 1)
                for (u32 j = 0; j < h; j++) {
                    mid[MID - 2 * h + j] = inv[(size_t)j * (n / (2 * h))];
                    mid[MID + MID - 2 * h + j] = fwd[(size_t)j * (n / (2 * h))];
                }
            d_tw_tile_dif_.upload(t_dif.data(), t_dif.size() * sizeof(Fr));
            d_tw_tile_dit_.upload(t_dit.data(), t_dit.size() * sizeof(Fr));
            d_tw_mid_.upload(mid.data(), mid.size() * sizeof(Fr));
        }
    }
    circ_.n_nodes = (u32)gh_.prog.size();
    circ_.n_slots = gh_.n_slots;
    circ_.n_wires = (u32)gh_.signals.size();
    circ_.prog = d_prog_.as<VmInstr>();
    circ_.consts = d_consts_.as<Fr>();
    circ_.signals = d_signals_.as<u32>();
    {   // list schedule for k_witness (host_util.hpp): bundles of 4 independent nodes, operand sources resolved (ring / const / global)
        uint32_t nb = 0;
        std::vector<uint8_t> is_signal(gh_.prog.size(), 0);
        for (uint32_t node : gh_.signals) is_signal[node] = 1;
        const uint32_t n_consts = (uint32_t)(gh_.consts.size() / 32);
        const bool consts_resident = n_consts <= vm_const_smem_max();
        circ_.n_consts_smem = consts_resident ? n_consts : 0;
        std::vector<VmRecord> recs = vm_build_schedule(gh_.prog, nb, &is_signal, consts_resident);
        {   // whole blocks for the bulk copies of k_witness: pad with empty bundles
            const uint32_t blk = vm_schedule_block_bundles();
            nb = (nb + blk - 1) / blk * blk;
            VmRecord empty;
            memset(&empty, 0, sizeof empty);
            empty.kind_op = 0xffffffffu;
            recs.resize((size_t)nb * VM_SLOTS, empty);
        }
        d_sched_.upload(recs.data(), recs.size() * sizeof(VmRecord));
        circ_.sched = d_sched_.as<uint4>();
        circ_.n_bundles = nb;
    }
    circ_.n_constraints = (u32)zk_.num_constraints;
    circ_.n_instance = (u32)zk_.num_instance;
    circ_.domain = domain_;
    circ_.log_domain = log_domain_;
    circ_.a_ptr = d_a_ptr_.as<u32>(); circ_.a_col = d_a_col_.as<u32>(); circ_.a_val = d_a_val_.as<Fr>();
    circ_.b_ptr = d_b_ptr_.as<u32>(); circ_.b_col = d_b_col_.as<u32>(); circ_.b_val = d_b_val_.as<Fr>();
    circ_.tw_inv = d_tw_inv_.as<Fr>();
    circ_.tw_fwd = d_tw_fwd_.as<Fr>();
    circ_.coset = d_coset_.as<Fr>();
    if (d_tw_tile_dif_.p) {
        circ_.tw_tile_dif = d_tw_tile_dif_.as<Fr>();
        circ_.tw_tile_dit = d_tw_tile_dit_.as<Fr>();
        circ_.tw_mid = d_tw_mid_.as<Fr>();
    }
    ZK_CUDA_CHECK(cudaDeviceSynchronize());
}

void Rln::build_tables() {
    // window size: as wide as HBM allows (tables are 64·2^(c−1)·K bytes per G1 base, twice that per G2 base)
    size_t free_b = 0, total_b = 0;
    ZK_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    const size_t nw = gh_.signals.size(), ni = zk_.num_instance;
    // bases at infinity contribute nothing and are dropped (47 in a_query, 1 999 in b_query for the bundled key)
    auto non_inf = [](const std::vector<uint8_t>& raw, size_t elem, size_t count, u32 first) {
        std::vector<uint32_t> v;
        for (size_t i = first; i < count; i++)
            if (!(raw[elem * i + elem - 1] & 0x40)) v.push_back((uint32_t)i);
        return v;
    };
    std::vector<uint32_t> pick[5];
    pick[0] = non_inf(zk_.a_query, 64, nw, 0);
    pick[1] = non_inf(zk_.b_g1, 64, nw, 0);
    pick[2] = non_inf(zk_.l_query, 64, nw - ni, 0);
    pick[3] = non_inf(zk_.h_query, 64, domain_, 0);
    pick[4] = non_inf(zk_.b_g2, 128, nw, 0);
    // order every group as [bases of wires a partial witness knows | the rest] so both phases are contiguous ranges
    u32 n_known[5] = {0, 0, 0, 0, 0};
    for (int g : {0, 1, 2, 4}) {
        const size_t shift = g == 2 ? ni : 0;
        auto mid = std::stable_partition(pick[g].begin(), pick[g].end(), [&](uint32_t i) { return wire_known_[i + shift] != 0; });
        n_known[g] = (u32)(mid - pick[g].begin());
    }
    size_t n_g1 = pick[0].size() + pick[1].size() + pick[2].size() + pick[3].size(), n_g2 = pick[4].size();
    // Window widths: G1 tables cost 64·K·2^(c−1) bytes per base, G2 tables twice that but have 6× fewer bases, so G2
    // gets the wider window when HBM allows (fewer additions per term: K = ⌈255/c⌉).
    // G1 scalars are GLV-split (k = k₁ + k₂·λ, |kᵢ| < 2^128): the windows only have to cover 129 bits (c·K > 129) and each
    // base is visited twice, so c = 13 costs 2 × 10 additions per term where the unsplit form needs 22 at c = 12
    const bool glv = env_int("RLN_B200_GLV", 1) != 0;
    auto windows_g1 = [&](int c1) { return glv ? (size_t)(129 / c1 + 1) : (size_t)((255 + c1 - 1) / c1); };
    auto table_bytes = [&](int c1, int c2) {
        size_t k1 = windows_g1(c1), k2 = windows_g1(c2);
        return (n_g1 * 64 * k1 << (c1 - 1)) + (n_g2 * 128 * k2 << (c2 - 1));
    };
    int c = env_int("RLN_B200_WINDOW_BITS", 0), c2 = env_int("RLN_B200_WINDOW_BITS_G2", 0);
    const size_t reserve = (size_t)26 << 30;  // proving workspace, tree, MSM scratch, slack
    const bool c_auto = c == 0;
    if (c == 0) {
        for (c = glv ? 13 : 12; c > 5; c--)
            if (table_bytes(c, c2 ? c2 : c) + reserve < free_b) break;
    }
    if (c2 == 0) {
        const int base2 = glv && c_auto ? 13 : c;   // G2 (GLV through β²): c = 15 → 2 × 9 additions per term, 72.6 GB
        for (c2 = base2 + 2; c2 > base2; c2--)
            if (table_bytes(c, c2) + reserve < free_b) break;
    }
    if (c < 5 || c > 16 || c2 < 5 || c2 > 16) throw RlnError("Configuration error: RLN_B200_WINDOW_BITS[_G2] must be in [5, 16]");
    const int K = (int)windows_g1(c), K2 = (int)windows_g1(c2);
    const int cd2 = c2 < 12 ? c2 : 12, Kd2 = (255 + cd2 - 1) / cd2;
    const int cd = c < 12 ? c : 12, Kd = (255 + cd - 1) / cd;   // δ₁ window table: unsplit scalars
    plan_.c = c;
    plan_.K = K;
    plan_.glv = glv ? 1 : 0;
    plan_.cd = cd;
    plan_.Kd = Kd;
    plan_.cd2 = cd2;
    plan_.Kd2 = Kd2;
    plan_.c2 = c2;
    plan_.K2 = K2;
    // scalar row of each base: A/B use wire i → node signals[i]; L uses wire ni+i; H uses row i of the h matrix
    for (int g = 0; g < 5; g++) {
        std::vector<uint32_t> rows(pick[g].size());
        for (size_t i = 0; i < rows.size(); i++) {
            if (g == 3) rows[i] = pick[g][i];
            else if (g == 2) rows[i] = gh_.signals[ni + pick[g][i]];
            else rows[i] = gh_.signals[pick[g][i]];
        }
        d_rows_[g].upload(rows.data(), rows.size() * 4);
        DevMem bases;
        const size_t half = (size_t)1 << ((g < 4 ? c : c2) - 1);
        if (g < 4) {
            const std::vector<uint8_t>& raw = g == 0 ? zk_.a_query : g == 1 ? zk_.b_g1 : g == 2 ? zk_.l_query : zk_.h_query;
            upload_points_g1(raw, pick[g], bases);
            d_tab_[g].alloc(sizeof(G1Affine) * pick[g].size() * K * half);
            launch_build_table_g1(bases.as<G1Affine>(), (u32)pick[g].size(), c, K, d_tab_[g].as<G1Affine>(), 0);
        } else {
            upload_points_g2(zk_.b_g2, pick[g], bases);
            d_tab_[g].alloc(sizeof(G2Affine) * pick[g].size() * K2 * half);
            launch_build_table_g2(bases.as<G2Affine>(), (u32)pick[g].size(), c2, K2, d_tab_[g].as<G2Affine>(), 0);
        }
        g_launch_count += 2;
        ZK_CUDA_CHECK(cudaDeviceSynchronize());
        MsmGroupDev& dst = g < 4 ? plan_.g1[g] : plan_.g2;
        dst.n_bases = (u32)pick[g].size();
        dst.n_known = n_known[g];
        dst.row = d_rows_[g].as<u32>();
        dst.table = d_tab_[g].p;
        dst.which_src = g == 3 ? 1 : 0;
    }
    {   // window tables of δ₁ and δ₂ for the blinding terms of the assembly
        std::vector<uint32_t> one = {0};
        DevMem b1, b2;
        upload_points_g1(zk_.delta_g1, one, b1);
        upload_points_g2(zk_.delta_g2, one, b2);
        d_delta1_tab_.alloc(sizeof(G1Affine) * Kd * ((size_t)1 << (cd - 1)));
        d_delta2_tab_.alloc(sizeof(G2Affine) * Kd2 * ((size_t)1 << (cd2 - 1)));
        launch_build_table_g1(b1.as<G1Affine>(), 1, cd, Kd, d_delta1_tab_.as<G1Affine>(), 0);
        launch_build_table_g2(b2.as<G2Affine>(), 1, cd2, Kd2, d_delta2_tab_.as<G2Affine>(), 0);
        g_launch_count += 4;
        ZK_CUDA_CHECK(cudaDeviceSynchronize());
        plan_.delta1_table = d_delta1_tab_.as<G1Affine>();
        {   // α₁ and β₁ with the same geometry: s·α₁ and r·β₁ of the folded assembly (a handful of proofs)
            DevMem ba, bb;
            upload_points_g1(zk_.alpha_g1, one, ba);
            upload_points_g1(zk_.beta_g1, one, bb);
            d_alpha1_tab_.alloc(sizeof(G1Affine) * Kd * ((size_t)1 << (cd - 1)));
            d_beta1_tab_.alloc(sizeof(G1Affine) * Kd * ((size_t)1 << (cd - 1)));
            launch_build_table_g1(ba.as<G1Affine>(), 1, cd, Kd, d_alpha1_tab_.as<G1Affine>(), 0);
            launch_build_table_g1(bb.as<G1Affine>(), 1, cd, Kd, d_beta1_tab_.as<G1Affine>(), 0);
            g_launch_count += 4;
            ZK_CUDA_CHECK(cudaDeviceSynchronize());
            plan_.alpha1_table = d_alpha1_tab_.as<G1Affine>();
            plan_.beta1_table = d_beta1_tab_.as<G1Affine>();
        }
        plan_.delta2_table = d_delta2_tab_.as<G2Affine>();
    }
    pk_.alpha_g1 = fetch_g1(zk_.alpha_g1);
    pk_.beta_g1 = fetch_g1(zk_.beta_g1);
    pk_.delta_g1 = fetch_g1(zk_.delta_g1);
    pk_.beta_g2 = fetch_g2(zk_.beta_g2);
    pk_.delta_g2 = fetch_g2(zk_.delta_g2);
    vk_.alpha_g1 = pk_.alpha_g1;
    vk_.beta_g2 = pk_.beta_g2;
    vk_.gamma_g2 = fetch_g2(zk_.gamma_g2);
    vk_.delta_g2 = pk_.delta_g2;
    {
        std::vector<uint32_t> all(zk_.gamma_abc.size() / 64);
        for (size_t i = 0; i < all.size(); i++) all[i] = (uint32_t)i;
        upload_points_g1(zk_.gamma_abc, all, d_gamma_abc_);
        vk_.gamma_abc = d_gamma_abc_.as<G1Affine>();
        vk_.n_public = (u32)all.size() - 1;
        // window tables of gamma_abc[1..]: vk_x = gamma_abc[0] + Σ xᵢ·gamma_abc[i] costs 32 additions per public input
        vk_.gc = 8;
        vk_.gK = (255 + vk_.gc - 1) / vk_.gc;
        d_gamma_tab_.alloc(sizeof(G1Affine) * vk_.n_public * vk_.gK * ((size_t)1 << (vk_.gc - 1)));
        launch_build_table_g1(d_gamma_abc_.as<G1Affine>() + 1, vk_.n_public, vk_.gc, vk_.gK, d_gamma_tab_.as<G1Affine>(), 0);
        g_launch_count += 2;
        ZK_CUDA_CHECK(cudaDeviceSynchronize());
        vk_.gamma_tab = d_gamma_tab_.as<G1Affine>();
    }
    {   // prepare_verifying_key: key-only parts of the pairing check (one-off, host portable arithmetic)
        PairingTables pr;
        pairing_tables_init(pr);
        struct Pre { Fq12 ml; FixedLines g, d; };
        auto pre = std::make_unique<Pre>();
        pre->ml = miller_loop(&pr, vk_.beta_g2, vk_.alpha_g1);
        precompute_lines(&pr, vk_.gamma_g2, pre->g);
        precompute_lines(&pr, vk_.delta_g2, pre->d);
        d_vk_pre_.upload(pre.get(), sizeof(Pre));
        const uint8_t* base = d_vk_pre_.as<uint8_t>();
        vk_.ml_alpha_beta = reinterpret_cast<const Fq12*>(base + offsetof(Pre, ml));
        vk_.gamma_lam = reinterpret_cast<const Fq2*>(base + offsetof(Pre, g) + offsetof(FixedLines, lam));
        vk_.gamma_c = reinterpret_cast<const Fq2*>(base + offsetof(Pre, g) + offsetof(FixedLines, c));
        vk_.delta_lam = reinterpret_cast<const Fq2*>(base + offsetof(Pre, d) + offsetof(FixedLines, lam));
        vk_.delta_c = reinterpret_cast<const Fq2*>(base + offsetof(Pre, d) + offsetof(FixedLines, c));
        // the lane-parallel verifier's program for this key (verify_vm_program.hpp): traced and scheduled here, once
        vm_max_batch_ = (size_t)std::max(0, env_int("RLN_B200_VERIFY_VM_MAX", 4096));
        {
            pvm::VerifyKeyHost h{&pre->g, &pre->d, pre->ml};
            const pvm::Program prog = pvm::build_verify_program(pr, h);
            d_vm_code_.upload(prog.code.data(), prog.code.size() * sizeof(u32));
            d_vm_consts_.upload(prog.consts.data(), prog.consts.size() * sizeof(Fq));
            vm_ = VerifyVmDev{d_vm_code_.as<u32>(), d_vm_consts_.as<Fq>(), prog.n_levels, prog.n_const, prog.n_slots, {}, nullptr};
            memcpy(vm_.exps, prog.exps, sizeof vm_.exps);
            vm_meta_.assign(prog.n_levels, 0);
            for (u32 l = 0; l < prog.n_levels; l++) {   // per level: N of each warp, max nsub, any combine, special id (for reports)
                const u32* rec = prog.code.data() + (size_t)l * pvm::REC_WORDS * pvm::LANES;
                u32 m = 0, nsub = 0, comb = 0;
                for (int w = 0; w < pvm::NW && w < 4; w++) {
                    const u32 w1 = rec[pvm::LANES + 32 * w];
                    m |= (w1 & 15) << (4 * w);
                    if (w1 & 15) { nsub = std::max(nsub, (w1 >> 4) & 3); comb |= (w1 >> 6) & 1; }
                }
                vm_meta_[l] = m | (nsub << 16) | (comb << 18) | (((rec[pvm::LANES] >> 8) & 0xff) << 20);
            }
            vm_est_cycles_ = prog.est_cycles;
        }
    }
}

// ------------------------------------------------------------------------------------------- tree
void Rln::set_tree(size_t depth) {
    if (depth == 0 || depth > 30) throw RlnError("Merkle tree error: Tree depth exceeds maximum allowed (must be < 64)");
    // RLN::set_tree replaces the tree by PoseidonTree::default(depth) (rln/src/public.rs:298-303): the old tree is dropped (its
    // store flushed) and the new one is a temporary tree without a store
    if (persistent()) { flush_store(); store_dir_.clear(); }
    tree_depth_ = depth;
    d_nodes_.alloc(sizeof(Fr) * ((size_t)2 << depth));
    launch_merkle_fill_empty(d_nodes_.as<Fr>(), (u32)depth, stream_);
    g_launch_count += 2;
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
    next_index_ = 0;
    leaf_set_.clear();
}
void Rln::set_range_device(size_t start, const uint8_t* d_leaves, size_t count, cudaStream_t s) {
    if (count == 0) return;
    if (start + count > capacity() || start + count < start) throw RlnError("Merkle tree error: set_range got too many leaves");
    g_launch_count += launch_merkle_set_range(d_nodes_.as<Fr>(), (u32)tree_depth_, start, d_leaves, count, s);
    store_dirty_ = true;
    mark_leaves(start, count, 1);
    if (start + count > next_index_) next_index_ = start + count;
}
void Rln::set_range_host(size_t start, const uint8_t* leaves, size_t count) {
    if (count == 0) return;
    if (start + count > capacity() || start + count < start) throw RlnError("Merkle tree error: set_range got too many leaves");
    DevMem tmp;
    tmp.upload(leaves, 32 * count);
    set_range_device(start, tmp.as<uint8_t>(), count, stream_);
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
}
void Rln::set_leaf(size_t index, const uint8_t* leaf) {
    if (index >= capacity()) throw oob_error();
    set_range_host(index, leaf, 1);
}
// delete never moves next_index.  PmTree: vacp2p_pmtree 2.0.3 `delete` refuses a never-used index (key >= next_index →
// TreeErrorKind::InvalidKey; pm_tree_adapter.rs:365-374 forwards it); FullMerkleTree / OptimalMerkleTree ignore such a call,
// whatever the index (full_merkle_tree.rs:278-286, optimal_merkle_tree.rs:253-260)
void Rln::delete_leaf(size_t index) {
    if (index >= next_index_) {
        if (tree_kind_ != TREE_PM) return;
        if (index >= capacity()) throw oob_error();
        throw RlnError("Merkle tree error: Pmtree error: Tree error: Invalid key");
    }
    uint8_t zero[32] = {0};
    size_t keep = next_index_;
    set_range_host(index, zero, 1);
    mark_leaves(index, 1, 0);
    next_index_ = keep;
}
void Rln::set_next(const uint8_t* leaf) {
    if (next_index_ >= capacity()) throw oob_error();
    set_range_host(next_index_, leaf, 1);
}
void Rln::download_leaves(size_t first, size_t count, uint8_t* out) {
    if (!count) return;
    DevMem tmp;
    tmp.alloc(32 * count);
    launch_fr_to_bytes(d_nodes_.as<Fr>() + capacity() + first, tmp.as<uint8_t>(), count, stream_);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaMemcpyAsync(out, tmp.p, 32 * count, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
}
void Rln::get_leaf(size_t index, uint8_t out[32]) {
    if (index >= capacity()) throw oob_error();
    DevMem tmp;
    tmp.alloc(32);
    launch_fr_to_bytes(d_nodes_.as<Fr>() + capacity() + index, tmp.as<uint8_t>(), 1, stream_);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaMemcpyAsync(out, tmp.p, 32, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
}
void Rln::root(uint8_t out[32]) {
    DevMem tmp;
    tmp.alloc(32);
    launch_fr_to_bytes(d_nodes_.as<Fr>() + 1, tmp.as<uint8_t>(), 1, stream_);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaMemcpyAsync(out, tmp.p, 32, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
}
void Rln::merkle_proofs(const uint64_t* idx, size_t n, uint8_t* elems, uint8_t* bits) {
    for (size_t i = 0; i < n; i++)
        if (idx[i] >= capacity()) throw oob_error();
    DevMem d_idx, d_el, d_bits;
    d_idx.upload(idx, 8 * n);
    d_el.alloc(32 * n * tree_depth_);
    d_bits.alloc(n * tree_depth_);
    launch_merkle_paths(d_nodes_.as<Fr>(), (u32)tree_depth_, d_idx.as<u64>(), n, d_el.as<uint8_t>(), d_bits.as<uint8_t>(), stream_);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaMemcpyAsync(elems, d_el.p, 32 * n * tree_depth_, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaMemcpyAsync(bits, d_bits.p, n * tree_depth_, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
}
// override_range with PmTree's "empty indices allowed" policy (rln/src/pm_tree_adapter.rs:320-356,
// utils/src/merkle_tree/override_range_validation.rs:20-65)
//
// This reproduces the reference's default tree (PmTree) state for state, including its two quirks, because RLN peers must
// agree on roots and on `leaves_set()` (which is where the next seq_atomic_operation starts):
//  * removals only (`remove_indices`, pm_tree_adapter.rs:427-445): the WHOLE span [first index, last index + 1) is reset to the
//    default leaf, not only the listed indices, and next_index = max(next_index, last index + 1);
//  * removals + leaves (`remove_indices_and_set_leaves`, :447-483): set_values = [min_index, start + n) with the listed indices
//    blanked and the leaves at offset start − min_index is written AT `start` (not at min_index): leaves below `start` keep
//    their hashes in the tree, the new leaves land (start − min_index) slots higher, next_index advances to
//    start + (start + n − min_index), and the cached "is set" flags are raised for [start, start + n − min_index) only
//    (rln/tests/poseidon_tree.rs:127-145 pins the resulting get_empty_leaves_indices()).
// vacp2p_pmtree 2.0.3 `set_range` (un-vendored): writes the values at [start, start + len), next_index = max(next_index, end),
// MerkleTreeIsFull when end > capacity.
void Rln::override_range(size_t start, const uint8_t* leaves, size_t n_leaves, std::vector<size_t> indices) {
    // validate_override_range_inputs (override_range_validation.rs:20-65); PmTree allows an empty index list, the dense trees do not
    if (tree_kind_ != TREE_PM && indices.empty()) throw RlnError("Merkle tree error: Invalid indices");
    for (size_t i : indices)
        if (i >= capacity()) throw RlnError("Merkle tree error: Invalid indices");
    std::sort(indices.begin(), indices.end());
    indices.erase(std::unique(indices.begin(), indices.end()), indices.end());
    size_t end = 0;
    if (n_leaves) {
        end = start + n_leaves;
        if (end < start || end > capacity()) throw RlnError("Merkle tree error: set_range got too many leaves");
        if (!indices.empty() && (indices[0] > start || indices[0] >= end)) throw RlnError("Merkle tree error: Invalid indices");
    }
    if (tree_kind_ != TREE_PM) { override_range_dense(start, leaves, n_leaves, indices); return; }
    const size_t n_idx = indices.size();
    if (n_leaves == 0 && n_idx == 0) throw RlnError("Merkle tree error: Leaf index out of bounds");
    if (n_leaves == 1 && n_idx == 0) { set_leaf(start, leaves); return; }
    if (n_leaves == 0 && n_idx == 1) { delete_leaf(indices[0]); return; }
    if (n_idx == 0) { set_range_host(start, leaves, n_leaves); return; }
    if (n_leaves == 0) {   // remove_indices: one set_range of default leaves over the span of the indices
        const size_t first = indices.front(), span = indices.back() + 1 - first;
        DevMem zeros;
        zeros.alloc(32 * span);
        ZK_CUDA_CHECK(cudaMemsetAsync(zeros.p, 0, 32 * span, stream_));
        set_range_device(first, zeros.as<uint8_t>(), span, stream_);
        mark_leaves(first, span, 0);
        ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
        return;
    }
    // remove_indices_and_set_leaves
    const size_t min_index = indices.front(), len = end - min_index, head = start - min_index;
    if (start + len > capacity() || start + len < start) throw RlnError("Merkle tree error: Pmtree error: Tree error: Merkle Tree is full");
    std::vector<uint8_t> set_values(32 * len, 0);
    download_leaves(min_index, head, set_values.data());   // leaves [min_index, start) that are not removed keep their value
    for (size_t i : indices)
        if (i < start) memset(&set_values[32 * (i - min_index)], 0, 32);
    memcpy(&set_values[32 * head], leaves, 32 * n_leaves);
    {
        DevMem tmp;
        tmp.upload(set_values.data(), set_values.size());
        g_launch_count += launch_merkle_set_range(d_nodes_.as<Fr>(), (u32)tree_depth_, start, tmp.as<uint8_t>(), len, stream_);
        ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
        store_dirty_ = true;
    }
    if (start + len > next_index_) next_index_ = start + len;
    if (leaf_set_.size() < next_index_) leaf_set_.resize(next_index_, 0);
    for (size_t i : indices) mark_leaves(i, 1, 0);
    if (len > start) mark_leaves(start, len - start, 1);   // `for i in start..(max_index - min_index)`
}
// FullMerkleTree / OptimalMerkleTree::override_range (full_merkle_tree.rs:226-269, optimal_merkle_tree.rs:197-244): the same
// set_values written at `start`, but the flags follow set_range: cleared for the indices first, then raised for the whole
// written range [start, start + len).  Without leaves max_index is `start`: the surviving leaves of [min_index, start) are
// re-written from `start` on.  min_index > max_index underflows in the reference (a panic); it is InvalidIndices here, as
// OptimalMerkleTree already answers for min_index >= max_index.
void Rln::override_range_dense(size_t start, const uint8_t* leaves, size_t n_leaves, const std::vector<size_t>& indices) {
    const size_t min_index = indices.front(), max_index = n_leaves ? start + n_leaves : start;
    if (min_index > max_index || (tree_kind_ == TREE_OPTIMAL && min_index >= max_index)) throw RlnError("Merkle tree error: Invalid indices");
    const size_t len = max_index - min_index, head = start > min_index ? start - min_index : 0;
    if (start + len > capacity() || start + len < start) throw RlnError("Merkle tree error: set_range got too many leaves");
    std::vector<uint8_t> set_values(32 * len, 0);
    download_leaves(min_index, head < len ? head : len, set_values.data());
    for (size_t i : indices)
        if (i >= min_index && i < start && i - min_index < len) memset(&set_values[32 * (i - min_index)], 0, 32);
    if (n_leaves) memcpy(&set_values[32 * head], leaves, 32 * n_leaves);
    if (leaf_set_.size() < next_index_) leaf_set_.resize(next_index_, 0);
    for (size_t i : indices) mark_leaves(i, 1, 0);
    if (len) set_range_host(start, set_values.data(), len);
}


// ------------------------------------------------------------------------------------------- tree store
static const char STORE_MAGIC[12] = {'R', 'L', 'N', 'B', '2', '0', '0', 'T', 'R', 'E', 'E', 0};
static const char* STORE_FILE = "/rlnb200_tree.bin";
static uint64_t fnv1a(const uint8_t* p, size_t n, uint64_t h = 1469598103934665603ull) {
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
static bool path_exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}
// resolve_path (pm_tree_adapter.rs:93-100) + PmTree::new (:194-239): depth check, load if a store is there, else create
void Rln::attach_store(const TreeConfig& cfg) {
    if (cfg.has_depth && cfg.tree_depth != tree_depth_)
        throw RlnError("Merkle tree error: Tree depth exceeds maximum allowed (must be < 64)");   // ZerokitMerkleTreeError::InvalidDepth
    if (cfg.temporary) {
        if (cfg.has_path && path_exists(cfg.path)) throw RlnError("Configuration error: Error while creating pmtree config: path already exists");
        return;   // a temporary tree: HBM only (the reference's temporary sled database is deleted when the tree is dropped)
    }
    if (!cfg.has_path) throw RlnError("Configuration error: Error while creating pmtree config: missing path");
    const std::string file = cfg.path + STORE_FILE;
    if (path_exists(file)) {
        std::vector<uint8_t> b = read_file_bytes(file);
        const size_t hdr = 12 + 4 + 4 + 8 + 8;
        auto corrupt = [&](const char* why) { return RlnError(std::string("Merkle tree error: Pmtree error: Database error: Cannot load database: ") + why + " (" + file + ")"); };
        if (b.size() < hdr + 8 || memcmp(b.data(), STORE_MAGIC, 12)) throw corrupt("not a tree store");
        uint32_t version, depth;
        uint64_t next, mlen, sum;
        memcpy(&version, &b[12], 4); memcpy(&depth, &b[16], 4); memcpy(&next, &b[20], 8); memcpy(&mlen, &b[28], 8);
        if (version != 1) throw corrupt("unknown store version");
        if (mlen > b.size() || next > ((uint64_t)1 << 30) || hdr + mlen + 32 * next + 8 != b.size()) throw corrupt("truncated store");
        memcpy(&sum, &b[b.size() - 8], 8);
        if (fnv1a(b.data(), b.size() - 8) != sum) throw corrupt("checksum mismatch");
        if (depth != tree_depth_) throw RlnError("Merkle tree error: Tree depth exceeds maximum allowed (must be < 64)");   // InvalidDepth (:205-208)
        if (next > capacity()) throw corrupt("more leaves than the tree holds");
        metadata.assign(b.begin() + hdr, b.begin() + hdr + mlen);
        const uint8_t* leaves = b.data() + hdr + mlen;
        for (uint64_t i = 0; i < next; i++)
            if (!fr_is_canonical(leaves + 32 * i)) throw corrupt("non-canonical leaf");
        if (next) set_range_host(0, leaves, next);
        next_index_ = next;
        // cached_leaves_indices after a reload: set wherever the stored leaf is not the default leaf (:224-233)
        leaf_set_.assign(next, 0);
        for (uint64_t i = 0; i < next; i++) leaf_set_[i] = is_zero32(leaves + 32 * i) ? 0 : 1;
    } else if (::mkdir(cfg.path.c_str(), 0777) != 0 && errno != EEXIST) {
        throw RlnError("Merkle tree error: Pmtree error: Database error: Cannot create database: " + std::string(strerror(errno)) + " (" + cfg.path + ")");
    }
    store_dir_ = cfg.path;
    store_dirty_ = !path_exists(file);   // a new store is written by the first flush even if nothing was inserted
}
void Rln::flush_store() {
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
    if (!persistent() || !store_dirty_) return;
    const uint64_t next = next_index_, mlen = metadata.size();
    const size_t hdr = 12 + 4 + 4 + 8 + 8;
    std::vector<uint8_t> b(hdr + mlen + 32 * next + 8);
    const uint32_t version = 1, depth = (uint32_t)tree_depth_;
    memcpy(&b[0], STORE_MAGIC, 12); memcpy(&b[12], &version, 4); memcpy(&b[16], &depth, 4); memcpy(&b[20], &next, 8); memcpy(&b[28], &mlen, 8);
    if (mlen) memcpy(&b[hdr], metadata.data(), mlen);
    download_leaves(0, next, b.data() + hdr + mlen);
    const uint64_t sum = fnv1a(b.data(), b.size() - 8);
    memcpy(&b[b.size() - 8], &sum, 8);
    // write beside the store, then rename over it: a crash leaves either the old or the new file, never a torn one
    const std::string file = store_dir_ + STORE_FILE, tmp = file + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f || !f.write((const char*)b.data(), (std::streamsize)b.size()) || !f.flush())
            throw RlnError("Merkle tree error: Pmtree error: Database error: Cannot flush database (" + tmp + ")");
    }
    if (::rename(tmp.c_str(), file.c_str()) != 0)
        throw RlnError("Merkle tree error: Pmtree error: Database error: Cannot flush database (" + std::string(strerror(errno)) + ")");
    store_dirty_ = false;
}

// ------------------------------------------------------------------------------------------- proving
Rln::TaskSet& Rln::tasks_for(u32 B, int phase) {
    const u64 key = ((u64)phase << 32) | B;
    auto it = tasks_.find(key);
    if (it != tasks_.end()) return *it->second;
    auto ts = std::make_unique<TaskSet>();
    std::vector<MsmTask> t1 = msm_make_tasks(plan_, B, false, phase), t2 = msm_make_tasks(plan_, B, true, phase);
    ts->g1.upload(t1.data(), t1.size() * sizeof(MsmTask));
    ts->g2.upload(t2.data(), t2.size() * sizeof(MsmTask));
    ts->n1 = (u32)t1.size();
    ts->n2 = (u32)t2.size();
    for (const MsmTask& t : t1) if (t.group <= 1) ts->n_ab++;
    for (size_t i = 0; i < ts->n_ab; i++) if (t1[i].group > 1) throw RlnError("internal: MSM tasks are not ordered by group");
    TaskSet& ref = *ts;
    tasks_[key] = std::move(ts);
    return ref;
}
void Rln::reserve(size_t B) {
    if (B <= cap_) return;
    if (B > max_batch_) B = max_batch_;
    if (B <= cap_) return;
    cap_ = 0;
    // largest task counts over the batch sizes this capacity can serve
    size_t t1 = 0, t2 = 0;
    for (size_t b = 1; b <= B; b <<= 1) {
        size_t a = msm_make_tasks(plan_, (u32)b, false, MSM_FULL).size() * b, c2 = msm_make_tasks(plan_, (u32)b, true, MSM_FULL).size() * b;
        if (a > t1) t1 = a;
        if (c2 > t2) t2 = c2;
    }
    {
        size_t a = msm_make_tasks(plan_, (u32)B, false, MSM_FULL).size() * B, c2 = msm_make_tasks(plan_, (u32)B, true, MSM_FULL).size() * B;
        if (a > t1) t1 = a;
        if (c2 > t2) t2 = c2;
    }
    // non power-of-two batch sizes below B can need more partial slots than the probes above: add headroom
    t1 += t1 / 2;
    t2 += t2 / 2;
    ws_inputs_.alloc(B * (size_t)gh_.n_slots * 32);
    ws_rs_.alloc(B * 64);
    ws_vals_.alloc(sizeof(Fr) * gh_.prog.size() * B);
    ws_a_.alloc(sizeof(Fr) * (size_t)domain_ * B);
    ws_b_.alloc(sizeof(Fr) * (size_t)domain_ * B);
    ws_c_.alloc(sizeof(Fr) * (size_t)domain_ * B);
    ws_err_.alloc(4 * B);
    ws_part1_.alloc(sizeof(G1XYZZ) * t1);
    ws_part2_.alloc(sizeof(G2XYZZ) * t2);
    ws_sum1_.alloc(sizeof(G1XYZZ) * 4 * B);
    ws_sum2_.alloc(sizeof(G2XYZZ) * B);
    ws_proofs_.alloc(128 * B);
    ws_values_.alloc(values_stride() * B);
    ws_affine_.alloc(256 * B);
    ws_partial_.alloc(320 * B);
    ws_partial_comp_.alloc(160 * B);
    cap_ = B;
}

void Rln::prove_device(const uint8_t* d_inputs, const uint8_t* d_rs, size_t n, uint8_t* d_proofs, uint8_t* d_values, uint8_t* d_affine,
                       cudaStream_t s, int phase, const uint8_t* d_partial, uint8_t* d_partial_affine, uint8_t* d_partial_comp) {
    if (n == 0) return;
    reserve(n);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t off = 0; off < n; off += cap_) {
        const u32 B = (u32)(n - off < cap_ ? n - off : cap_);
        TaskSet& ts = tasks_for(B, phase);
        ws_part1_.ensure((size_t)ts.n1 * B * sizeof(G1XYZZ));
        ws_part2_.ensure((size_t)ts.n2 * B * sizeof(G2XYZZ));
        const uint8_t* in = d_inputs + off * (size_t)gh_.n_slots * 32;
        const bool values_from_wires = B <= SMALL_MAX && phase == MSM_FULL && !d_partial && !ext_wires_;   // a handful of full proofs: see below
        ZK_CUDA_CHECK(cudaEventRecord(ev_[0], s));
        if (ext_wires_) launch_scatter_wires(circ_, ext_wires_, ws_vals_.as<Fr>(), B, ws_err_.as<u32>(), s);
        else launch_witness(circ_, in, ws_vals_.as<Fr>(), B, ws_err_.as<u32>(), s);
        if (d_values && phase != MSM_KNOWN && !values_from_wires) {
            // proof values only need the inputs: a latency-bound kernel, run beside the main pipeline — but forked AFTER the witness
            // kernel: started beside it, its CTAs (no shared memory) take 32 SMs out of the 82 KB carve-out the witness CTAs need,
            // a dozen SMs then host two witness CTAs each, and two CTAs on one SM run their products half as fast: the witness
            // kernel took 7.7 ms per 4 096 proofs in the pipeline where it takes 3.9 ms alone (ncu launch list r02at).  Beside the
            // QAP and the MSM launches the 4 096 threads of this kernel disappear.
            ZK_CUDA_CHECK(cudaEventRecord(fork_, s));
            ZK_CUDA_CHECK(cudaStreamWaitEvent(side_, fork_, 0));
            launch_proof_values(in, slots_, B, d_values + values_stride() * off, side_);
            ZK_CUDA_CHECK(cudaEventRecord(join_, side_));
        }
        // for a handful of proofs the values are read off the witness (they are its public signals): the side kernel would hash
        // every Merkle path a second time on one thread per proof, 6 ms — longer than everything else here
        if (d_values && values_from_wires)
            launch_values_from_wires(ws_vals_.as<Fr>(), circ_.signals, B, (u32)max_out_, d_values + values_stride() * off, s);
        ZK_CUDA_CHECK(cudaEventRecord(ev_[1], s));
        // (running the QAP on a side stream beside the A / B1 / L accumulate tasks was measured slower, DESIGN §7b, and removed)
        if (phase != MSM_KNOWN) launch_qap(circ_, ws_vals_.as<Fr>(), ws_a_.as<Fr>(), ws_b_.as<Fr>(), ws_c_.as<Fr>(), B, s);
        ZK_CUDA_CHECK(cudaEventRecord(ev_[2], s));
        MsmWorkspace mw;
        mw.part_g1 = ws_part1_.as<G1XYZZ>();
        mw.part_g2 = ws_part2_.as<G2XYZZ>();
        mw.sum_g1 = ws_sum1_.as<G1XYZZ>();
        mw.sum_g2 = ws_sum2_.as<G2XYZZ>();
        mw.tasks_g1 = ts.g1.as<MsmTask>();
        mw.tasks_g2 = ts.g2.as<MsmTask>();
        mw.n_tasks_g1 = ts.n1;
        mw.n_tasks_g2 = ts.n2;
        mw.ev = mev_;
        mw.side = asm_side_;
        mw.side_fork = asm_fork_;
        mw.side_join = asm_join_;
        if (fold_on_ && B <= FOLD_MAX && phase == MSM_FULL && !d_partial) {
            // a handful of full proofs: s·(ΣzᵢAᵢ) and r·(ΣzᵢB₁ᵢ) as table sums over s·z and r·z instead of the Straus run of the assembly
            ws_fold_s_.ensure(sizeof(Fr) * gh_.prog.size() * FOLD_MAX);
            ws_fold_r_.ensure(sizeof(Fr) * gh_.prog.size() * FOLD_MAX);
            ws_fold_part_.ensure(sizeof(G1XYZZ) * (size_t)ts.n_ab * B);
            ws_fold_sum_.ensure(sizeof(G1XYZZ) * 2 * FOLD_MAX);
            launch_scale_vals(ws_vals_.as<Fr>(), d_rs + 64 * off, (u32)gh_.prog.size(), B, ws_fold_s_.as<Fr>(), ws_fold_r_.as<Fr>(), s);
            mw.fold_s = ws_fold_s_.as<Fr>();
            mw.fold_r = ws_fold_r_.as<Fr>();
            mw.fold_part = ws_fold_part_.as<G1XYZZ>();
            mw.fold_sum = ws_fold_sum_.as<G1XYZZ>();
            mw.n_tasks_ab = ts.n_ab;
            g_launch_count += 3;
        }
        launch_msm_sums(plan_, ws_vals_.as<Fr>(), ws_a_.as<Fr>(), B, mw, s);
        if (phase == MSM_KNOWN) {
            launch_partial_out(pk_, B, mw, d_partial_affine + 320 * off, d_partial_comp + 160 * off, s);
            ZK_CUDA_CHECK(cudaEventRecord(mev_[5], s));
        } else {
            launch_assemble(plan_, pk_, B, d_rs + 64 * off, mw, d_partial ? d_partial + 320 * off : nullptr, d_proofs + 128 * off,
                            d_affine ? d_affine + 256 * off : nullptr, s);
        }
        ZK_CUDA_CHECK(cudaEventRecord(ev_[3], s));
        if (d_values && phase != MSM_KNOWN && !values_from_wires) ZK_CUDA_CHECK(cudaStreamWaitEvent(s, join_, 0));
        ZK_CUDA_CHECK(cudaEventRecord(ev_[4], s));
        g_launch_count += 1 + (phase != MSM_KNOWN ? qap_launch_count(circ_, B) : 0) + 6 + (d_values ? 1 : 0);
        // graph-evaluation failures surface as errors, like WitnessCalcError::GraphEvaluation (rln/src/circuit/iden3calc.rs:52-53)
        std::vector<u32> err(B);
        ZK_CUDA_CHECK(cudaMemcpyAsync(err.data(), ws_err_.p, 4 * B, cudaMemcpyDeviceToHost, s));
        ZK_CUDA_CHECK(cudaStreamSynchronize(s));
        {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev_[0], ev_[1]); acc[0] += ms;
            cudaEventElapsedTime(&ms, ev_[1], ev_[2]);
            acc[1] += ms;
            for (int i = 0; i < 5; i++) { cudaEventElapsedTime(&ms, mev_[i], mev_[i + 1]); acc[2 + i] += ms; }
            cudaEventElapsedTime(&ms, ev_[3], ev_[4]); acc[7] += ms;
        }
        if (phase != MSM_KNOWN)  // a partial witness leaves the unknown wires undefined; only full evaluations can fail
            for (u32 j = 0; j < B; j++)
                if (err[j]) throw RlnError("Protocol error: Error calculating witness: Failed to evaluate witness calculation graph: operator not implemented for Montgomery");
    }
    for (int i = 0; i < 8; i++) stage_ms[i] = acc[i];
}

// zeroes the host staging vector and the device input buffer when the scope ends, normally or by exception (the reference
// zeroises IdSecret on drop on all paths: rln/src/utils.rs:443-527, rln/src/circuit/iden3calc.rs:44-57)
struct SecretScrub {
    std::vector<uint8_t>& host;
    DevMem& dev;
    cudaStream_t s;
    ~SecretScrub() {
        if (!host.empty()) {
            volatile uint8_t* p = host.data();
            for (size_t i = 0; i < host.size(); i++) p[i] = 0;
        }
        if (dev.p) {
            cudaMemsetAsync(dev.p, 0, dev.bytes, s);
            cudaStreamSynchronize(s);
        }
    }
};

void Rln::witness_slots(const Witness& w, uint8_t* slots) const {  // iden3calc.rs:106-181, witness.rs:832-881
    memset(slots, 0, (size_t)gh_.n_slots * 32);
    slots[0] = 1;
    memcpy(slots + 32 * slots_.secret, w.secret, 32);
    memcpy(slots + 32 * slots_.limit, w.limit, 32);
    memcpy(slots + 32 * slots_.message_id, w.mids.data(), 32 * max_out_);
    if (multi_)
        for (size_t i = 0; i < max_out_; i++) slots[32 * (slots_.selector + i)] = w.sel[i] ? 1 : 0;
    memcpy(slots + 32 * slots_.path, w.path.data(), 32 * depth_);
    for (size_t i = 0; i < depth_; i++) slots[32 * (slots_.index + i)] = w.index[i];
    memcpy(slots + 32 * slots_.x, w.x, 32);
    memcpy(slots + 32 * slots_.ext_null, w.ext_null, 32);
}

// validate_witness_against_graph (proof.rs:644-700)
void Rln::check_witness_shape(const Witness& w) const {
    if (w.multi != multi_)
        throw RlnError(std::string("Protocol error: Witness message mode ") + (w.multi ? "MultiV1" : "SingleV1") + " does not match graph mode " +
                       (multi_ ? "MultiV1" : "SingleV1"));
    if (w.k() != max_out_) {
        std::ostringstream s;
        s << "Protocol error: The field message_ids has length " << w.k() << ", but the field max_out has length " << max_out_;
        throw RlnError(s.str());
    }
    if (w.path.size() / 32 != depth_) {
        std::ostringstream s;
        s << "Protocol error: The field path_elements has length " << w.path.size() / 32 << ", but the field tree_depth has length " << depth_;
        throw RlnError(s.str());
    }
    if (w.index.size() != depth_) {
        std::ostringstream s;
        s << "Protocol error: The field identity_path_index has length " << w.index.size() << ", but the field tree_depth has length " << depth_;
        throw RlnError(s.str());
    }
}

void Rln::prove_host(const std::vector<Witness>& wsv, const uint8_t* rs, std::vector<RlnProof>& out, const PartialProofHost* partials) {
    const size_t n = wsv.size();
    out.resize(n);
    if (!n) return;
    for (const Witness& w : wsv) check_witness_shape(w);
    const size_t chunk = n < max_batch_ ? n : max_batch_;
    reserve(chunk);
    const size_t vs = values_stride(), k = max_out_;
    std::vector<uint8_t> slots(chunk * (size_t)gh_.n_slots * 32), rsb(chunk * 64), proofs(chunk * 128), values(chunk * vs);
    SecretScrub scrub{slots, ws_inputs_, stream_};   // identity secrets leave the staging buffers on every path out of here
    for (size_t off = 0; off < n; off += chunk) {
        const size_t B = n - off < chunk ? n - off : chunk;
        for (size_t j = 0; j < B; j++) witness_slots(wsv[off + j], slots.data() + j * (size_t)gh_.n_slots * 32);
        if (rs) memcpy(rsb.data(), rs + 64 * off, 64 * B);
        else
            for (size_t j = 0; j < 2 * B; j++) random_fr(rsb.data() + 32 * j);  // r, s ← rng (proof.rs:743-745)
        ZK_CUDA_CHECK(cudaMemcpyAsync(ws_inputs_.p, slots.data(), B * (size_t)gh_.n_slots * 32, cudaMemcpyHostToDevice, stream_));
        ZK_CUDA_CHECK(cudaMemcpyAsync(ws_rs_.p, rsb.data(), 64 * B, cudaMemcpyHostToDevice, stream_));
        if (partials) {
            std::vector<uint8_t> pb(320 * B);
            for (size_t j = 0; j < B; j++) memcpy(pb.data() + 320 * j, partials[off + j].affine, 320);
            ZK_CUDA_CHECK(cudaMemcpyAsync(ws_partial_.p, pb.data(), pb.size(), cudaMemcpyHostToDevice, stream_));
            ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
        }
        prove_device(ws_inputs_.as<uint8_t>(), ws_rs_.as<uint8_t>(), B, ws_proofs_.as<uint8_t>(), ws_values_.as<uint8_t>(), nullptr, stream_,
                     partials ? MSM_UNKNOWN : MSM_FULL, partials ? ws_partial_.as<uint8_t>() : nullptr);
        ZK_CUDA_CHECK(cudaMemcpyAsync(proofs.data(), ws_proofs_.p, 128 * B, cudaMemcpyDeviceToHost, stream_));
        ZK_CUDA_CHECK(cudaMemcpyAsync(values.data(), ws_values_.p, vs * B, cudaMemcpyDeviceToHost, stream_));
        // secrets do not linger in the staging buffers (reference zeroises them: rln/src/circuit/iden3calc.rs:44-57)
        ZK_CUDA_CHECK(cudaMemsetAsync(ws_inputs_.p, 0, B * (size_t)gh_.n_slots * 32, stream_));
        ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
        for (size_t j = 0; j < B; j++) {
            RlnProof& p = out[off + j];
            memcpy(p.proof, proofs.data() + 128 * j, 128);
            const uint8_t* v = values.data() + vs * j;
            memcpy(p.pv.root, v, 32);
            memcpy(p.pv.ext_null, v + 32, 32);
            memcpy(p.pv.x, v + 64, 32);
            p.pv.multi = multi_;
            p.pv.ys.assign(v + 96, v + 96 + 32 * k);
            p.pv.nulls.assign(v + 96 + 32 * k, v + 96 + 64 * k);
            if (multi_) p.pv.sel = wsv[off + j].sel;
        }
    }
    memset(slots.data(), 0, slots.size());
}


// bulk randomness for r, s (proof.rs:743-745): uniform in [0, r) by rejection on 254 bits, from the OS generator
static void random_fr_bulk(uint8_t* out, size_t count) {
    std::vector<uint8_t> pool;
    size_t used = 0;
    for (size_t i = 0; i < count;) {
        if (used + 32 > pool.size()) {
            pool.resize(32 * (count - i) + 1024);
            used = 0;
            std::ifstream ur("/dev/urandom", std::ios::binary);
            if (!ur.read((char*)pool.data(), (std::streamsize)pool.size())) {   // no OS generator: fall back to std::random_device
                for (; i < count; i++) random_fr(out + 32 * i);
                return;
            }
        }
        memcpy(out + 32 * i, pool.data() + used, 32);
        used += 32;
        out[32 * i + 31] &= 0x3f;
        if (fr_is_canonical(out + 32 * i)) i++;
    }
    if (!pool.empty()) memset(pool.data(), 0, pool.size());
}

void Rln::prove_records_device(const uint8_t* d_records, const uint8_t* d_rs, size_t n, uint8_t* d_out, cudaStream_t s) {
    if (!n) return;
    const size_t chunk = n < max_batch_ ? n : max_batch_;
    reserve(chunk);
    const RecordLayout L = record_layout();
    ws_bad_.ensure(4 * cap_);
    std::vector<u32> bad(cap_);
    std::vector<uint8_t> rsb;
    float acc_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u32 n_chunks = 0;
    struct Scrub {   // secrets never outlive the call, whichever way it ends (reference: IdSecret zeroises on drop)
        Rln& r; cudaStream_t s;
        ~Scrub() { cudaMemsetAsync(r.ws_inputs_.p, 0, r.ws_inputs_.bytes, s); cudaStreamSynchronize(s); }
    } scrub{*this, s};
    for (size_t off = 0; off < n; off += cap_) {
        const size_t B = n - off < cap_ ? n - off : cap_;
        launch_witness_records(d_records + off * (size_t)L.rec_len, B, L, ws_inputs_.as<uint8_t>(), ws_bad_.as<u32>(), s);
        g_launch_count++;
        ZK_CUDA_CHECK(cudaMemcpyAsync(bad.data(), ws_bad_.p, 4 * B, cudaMemcpyDeviceToHost, s));
        const uint8_t* rs_dev = d_rs ? d_rs + 64 * off : nullptr;
        if (!d_rs) {
            rsb.resize(64 * B);
            random_fr_bulk(rsb.data(), 2 * B);
            ZK_CUDA_CHECK(cudaMemcpyAsync(ws_rs_.p, rsb.data(), 64 * B, cudaMemcpyHostToDevice, s));
            rs_dev = ws_rs_.as<uint8_t>();
        }
        ZK_CUDA_CHECK(cudaStreamSynchronize(s));
        for (size_t j = 0; j < B; j++) {
            if (!bad[j]) continue;
            // the reference's own wording: re-parse the refused record with the host parser (same rules, same messages)
            std::vector<uint8_t> rec(L.rec_len);
            ZK_CUDA_CHECK(cudaMemcpy(rec.data(), d_records + (off + j) * (size_t)L.rec_len, L.rec_len, cudaMemcpyDeviceToHost));
            Witness w;
            std::string msg = "witness record " + std::to_string(off + j) + ": ";
            try {
                const size_t used = witness_from_bytes(rec.data(), rec.size(), w);
                memset(rec.data(), 0, rec.size());
                if (w.multi != multi_)
                    throw RlnError(std::string("Protocol error: Witness message mode ") + (w.multi ? "MultiV1" : "SingleV1") + " does not match graph mode " +
                                   (multi_ ? "MultiV1" : "SingleV1"));
                if (used != rec.size() || w.path.size() / 32 != depth_ || w.index.size() != depth_ || w.k() != max_out_)
                    throw RlnError("Protocol error: witness record does not have the shape of the circuit (tree depth " + std::to_string(depth_) +
                                   ", " + std::to_string(max_out_) + " message id slots)");
                throw RlnError("record refused on the device but accepted by the host parser");
            } catch (const RlnError& e) {
                memset(w.secret, 0, 32);
                throw RlnError(msg + e.what());
            }
        }
        prove_device(ws_inputs_.as<uint8_t>(), rs_dev, B, ws_proofs_.as<uint8_t>(), ws_values_.as<uint8_t>(), nullptr, s);
        for (int i = 0; i < 8; i++) acc_ms[i] += stage_ms[i];
        n_chunks++;
        launch_proof_records(ws_proofs_.as<uint8_t>(), ws_values_.as<uint8_t>(), ws_inputs_.as<uint8_t>(), B, L, d_out + off * (size_t)L.proof_rec_len, s);
        g_launch_count++;
    }
    ZK_CUDA_CHECK(cudaStreamSynchronize(s));
    if (!rsb.empty()) memset(rsb.data(), 0, rsb.size());
    for (int i = 0; i < 8; i++) stage_ms[i] = acc_ms[i];   // the whole call, summed over its device batches
    last_chunks = n_chunks;
}

void Rln::prove_records_host(const uint8_t* records, const uint8_t* rs, size_t n, uint8_t* out) {
    if (!n) return;
    const RecordLayout L = record_layout();
    const size_t super = 16 * max_batch_;   // records resident on the device at a time
    const size_t m = n < super ? n : super;
    ws_recs_in_.ensure(m * (size_t)L.rec_len);
    ws_recs_out_.ensure(m * (size_t)L.proof_rec_len);
    if (rs) ws_rs_all_.ensure(64 * m);
    struct Scrub {
        Rln& r;
        ~Scrub() { cudaMemsetAsync(r.ws_recs_in_.p, 0, r.ws_recs_in_.bytes, r.stream_); cudaStreamSynchronize(r.stream_); }
    } scrub{*this};
    for (size_t off = 0; off < n; off += m) {
        const size_t B = n - off < m ? n - off : m;
        ZK_CUDA_CHECK(cudaMemcpyAsync(ws_recs_in_.p, records + off * (size_t)L.rec_len, B * (size_t)L.rec_len, cudaMemcpyHostToDevice, stream_));
        if (rs) ZK_CUDA_CHECK(cudaMemcpyAsync(ws_rs_all_.p, rs + 64 * off, 64 * B, cudaMemcpyHostToDevice, stream_));
        prove_records_device(ws_recs_in_.as<uint8_t>(), rs ? ws_rs_all_.as<uint8_t>() : nullptr, B, ws_recs_out_.as<uint8_t>(), stream_);
        ZK_CUDA_CHECK(cudaMemcpyAsync(out + off * (size_t)L.proof_rec_len, ws_recs_out_.p, B * (size_t)L.proof_rec_len, cudaMemcpyDeviceToHost, stream_));
        ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
    }
}

void Rln::partial_host(const std::vector<Witness>& wsv, std::vector<PartialProofHost>& out) {
    const size_t n = wsv.size();
    out.resize(n);
    if (!n) return;
    for (const Witness& w : wsv)
        if (w.path.size() / 32 != depth_ || w.index.size() != depth_ || w.k() != max_out_)
            throw RlnError("Protocol error: partial witness depth does not match the circuit tree_depth");
    const size_t chunk = n < max_batch_ ? n : max_batch_;
    reserve(chunk);
    std::vector<uint8_t> slots(chunk * (size_t)gh_.n_slots * 32), aff(chunk * 320), comp(chunk * 160);
    SecretScrub scrub{slots, ws_inputs_, stream_};
    for (size_t off = 0; off < n; off += chunk) {
        const size_t B = n - off < chunk ? n - off : chunk;
        for (size_t j = 0; j < B; j++) {
            Witness w = wsv[off + j];
            std::fill(w.mids.begin(), w.mids.end(), 0);  // the unknown inputs are None in the reference; any value works, their wires are not used
            std::fill(w.sel.begin(), w.sel.end(), 0);
            memset(w.x, 0, 32);
            memset(w.ext_null, 0, 32);
            witness_slots(w, slots.data() + j * (size_t)gh_.n_slots * 32);
            memset(w.secret, 0, 32);
        }
        ZK_CUDA_CHECK(cudaMemcpyAsync(ws_inputs_.p, slots.data(), B * (size_t)gh_.n_slots * 32, cudaMemcpyHostToDevice, stream_));
        prove_device(ws_inputs_.as<uint8_t>(), nullptr, B, nullptr, nullptr, nullptr, stream_, MSM_KNOWN, nullptr, ws_partial_.as<uint8_t>(),
                     ws_partial_comp_.as<uint8_t>());
        ZK_CUDA_CHECK(cudaMemcpyAsync(aff.data(), ws_partial_.p, 320 * B, cudaMemcpyDeviceToHost, stream_));
        ZK_CUDA_CHECK(cudaMemcpyAsync(comp.data(), ws_partial_comp_.p, 160 * B, cudaMemcpyDeviceToHost, stream_));
        ZK_CUDA_CHECK(cudaMemsetAsync(ws_inputs_.p, 0, B * (size_t)gh_.n_slots * 32, stream_));
        ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
        for (size_t j = 0; j < B; j++) {
            memcpy(out[off + j].affine, aff.data() + 320 * j, 320);
            memcpy(out[off + j].comp, comp.data() + 160 * j, 160);
        }
    }
    memset(slots.data(), 0, slots.size());
}

// PartialProof::deserialize_compressed validates every point (curve, and the subgroup for G2)
void Rln::decompress_partial(const uint8_t comp[160], uint8_t affine[320]) {
    // reuse the proof decompressor (G1 | G2 | G1): (π_a, π_b, π_c) and (ρ, π_b, π_c)
    uint8_t two[256];
    memcpy(two, comp, 32); memcpy(two + 32, comp + 64, 64); memcpy(two + 96, comp + 128, 32);
    memcpy(two + 128, comp + 32, 32); memcpy(two + 160, comp + 64, 64); memcpy(two + 224, comp + 128, 32);
    DevMem dp, da, dok;
    dp.upload(two, 256);
    da.alloc(512);
    dok.alloc(2);
    launch_decompress(dp.as<uint8_t>(), 2, da.as<uint8_t>(), dok.as<uint8_t>(), stream_);
    g_launch_count++;
    uint8_t ok[2], aff[512];
    ZK_CUDA_CHECK(cudaMemcpyAsync(ok, dok.p, 2, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaMemcpyAsync(aff, da.p, 512, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
    if (!ok[0] || !ok[1]) throw RlnError("Proof serialization error: the input buffer contained invalid data");
    memcpy(affine, aff, 64);             // π_a
    memcpy(affine + 64, aff + 256, 64);  // ρ
    memcpy(affine + 128, aff + 64, 128); // π_b
    memcpy(affine + 256, aff + 192, 64); // π_c
}

void Rln::debug_w_h(const Witness& w, uint8_t* w_out, uint8_t* h_out) {
    reserve(1);
    std::vector<uint8_t> slots((size_t)gh_.n_slots * 32);
    witness_slots(w, slots.data());
    ZK_CUDA_CHECK(cudaMemcpyAsync(ws_inputs_.p, slots.data(), slots.size(), cudaMemcpyHostToDevice, stream_));
    launch_witness(circ_, ws_inputs_.as<uint8_t>(), ws_vals_.as<Fr>(), 1, ws_err_.as<u32>(), stream_);
    launch_qap(circ_, ws_vals_.as<Fr>(), ws_a_.as<Fr>(), ws_b_.as<Fr>(), ws_c_.as<Fr>(), 1, stream_);
    g_launch_count += 1 + qap_launch_count(circ_, 1);
    DevMem vals_b, h_b;
    vals_b.alloc(32 * gh_.prog.size());
    h_b.alloc(32 * (size_t)domain_);
    launch_fr_to_bytes(ws_vals_.as<Fr>(), vals_b.as<uint8_t>(), gh_.prog.size(), stream_);
    launch_fr_to_bytes(ws_a_.as<Fr>(), h_b.as<uint8_t>(), domain_, stream_);
    g_launch_count += 2;
    std::vector<uint8_t> vals(32 * gh_.prog.size());
    ZK_CUDA_CHECK(cudaMemcpyAsync(vals.data(), vals_b.p, vals.size(), cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaMemcpyAsync(h_out, h_b.p, 32 * (size_t)domain_, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
    for (size_t i = 0; i < gh_.signals.size(); i++) memcpy(w_out + 32 * i, vals.data() + 32 * (size_t)gh_.signals[i], 32);
}

void Rln::verify_batch(const uint8_t* proofs128, const uint8_t* publics, size_t n, uint8_t* ok) {
    if (!n) return;
    // one staging buffer (proofs | publics) kept between calls: a single verification is two copies and one launch
    const size_t pb = 128 * n, vb = 32 * n * vk_.n_public;
    d_vfy_in_.ensure(pb + vb);
    d_vfy_ok_.ensure(n);
    uint8_t* dp = d_vfy_in_.as<uint8_t>();
    uint8_t* dv = dp + pb;
    ZK_CUDA_CHECK(cudaMemcpyAsync(dp, proofs128, pb, cudaMemcpyHostToDevice, stream_));
    ZK_CUDA_CHECK(cudaMemcpyAsync(dv, publics, vb, cudaMemcpyHostToDevice, stream_));
    const bool use_vm = vm_max_batch_ && n <= vm_max_batch_ && vk_.n_public <= 32;
    if (use_vm) launch_verify_vm(vm_, vk_, dp, dv, n, d_vfy_ok_.as<uint8_t>(), stream_);
    else launch_verify(vk_, dp, dv, n, d_vfy_ok_.as<uint8_t>(), stream_);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaMemcpyAsync(ok, d_vfy_ok_.p, n, cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
    if (!use_vm) return;
    // proofs the lane-parallel kernel does not decide (a point at infinity, an exceptional addition in the membership test):
    // the one-thread-per-proof kernel has complete formulas
    std::vector<size_t> redo;
    for (size_t j = 0; j < n; j++) if (ok[j] == 3) redo.push_back(j);
    if (redo.empty()) return;
    std::vector<uint8_t> rp(128 * redo.size()), rv(32 * (size_t)vk_.n_public * redo.size()), rok(redo.size());
    for (size_t k = 0; k < redo.size(); k++) {
        memcpy(rp.data() + 128 * k, proofs128 + 128 * redo[k], 128);
        memcpy(rv.data() + 32 * (size_t)vk_.n_public * k, publics + 32 * (size_t)vk_.n_public * redo[k], 32 * (size_t)vk_.n_public);
    }
    DevMem d2, v2, o2;
    d2.upload(rp.data(), rp.size());
    v2.upload(rv.data(), rv.size());
    o2.alloc(redo.size());
    launch_verify(vk_, d2.as<uint8_t>(), v2.as<uint8_t>(), redo.size(), o2.as<uint8_t>(), stream_);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaMemcpyAsync(rok.data(), o2.p, redo.size(), cudaMemcpyDeviceToHost, stream_));
    ZK_CUDA_CHECK(cudaStreamSynchronize(stream_));
    for (size_t k = 0; k < redo.size(); k++) ok[redo[k]] = rok[k];
}

}  // namespace zk

// =============================================================================================== C ABI
using namespace zk;

struct FFI_RLN { std::unique_ptr<Rln> r; };
// copies of a caller's witness are zeroised when they go out of scope, on every path (IdSecret's Drop: rln/src/utils.rs:443-527)
struct WitnessWipe {
    std::vector<Witness>& v;
    ~WitnessWipe() { for (Witness& w : v) memset(w.secret, 0, 32); }
};
// the handle's mutex + its device made current on the calling thread (one handle, many threads, several devices)
struct RlnLock {
    std::lock_guard<std::mutex> lk;
    DeviceGuard dg;
    explicit RlnLock(Rln& r) : lk(r.mu), dg(r.device()) {}
};
struct FFI_RLNProof { RlnProof p; };
struct FFI_RLNProofValues { ProofValues v; };
struct FFI_RLNWitnessInput { Witness w; };
struct RlnB200Msm { VarMsmWorkspace* ws; size_t max_n; int device; };
struct FFI_RLNPartialWitnessInput { Witness w; };
struct FFI_RLNPartialProof { PartialProofHost p; std::vector<uint8_t> mask; };

static RlnString mk_string(const std::string& s) {
    RlnString out;
    out.len = s.size();
    out.cap = s.size() + 1;
    out.ptr = (uint8_t*)malloc(out.cap);
    memcpy(out.ptr, s.data(), s.size());
    out.ptr[s.size()] = 0;
    return out;
}
static RlnString no_string() { return RlnString{nullptr, 0, 0}; }
static Vec_uint8_t mk_vec(const uint8_t* p, size_t n) {
    Vec_uint8_t v;
    v.len = n;
    v.cap = n ? n : 1;
    v.ptr = (uint8_t*)malloc(v.cap);
    if (n) memcpy(v.ptr, p, n);
    return v;
}
static CFr_t* mk_cfr(const uint8_t* b) {
    CFr_t* c = (CFr_t*)malloc(sizeof(CFr_t));
    memcpy(c->bytes, b, 32);
    return c;
}
static std::string describe(const std::exception& e) { return e.what(); }
static std::string describe(const CudaError& e) {
    std::ostringstream s;
    s << "CUDA error: " << cudaGetErrorString(e.code) << " (" << e.expr << " at " << e.file << ":" << e.line << ")";
    return s.str();
}
#define GUARD_BEGIN try {
#define GUARD_END(on_error)                                              \
    }                                                                    \
    catch (const CudaError& e) { std::string m = describe(e); on_error; } \
    catch (const std::exception& e) { std::string m = describe(e); on_error; }

// bundled circuit files live next to the library: <dir of librln_b200.so>/../resources
static std::string resources_dir() {
    const char* env = getenv("RLN_B200_RESOURCES");
    if (env && *env) return env;
    Dl_info info;
    if (dladdr((void*)&resources_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.rfind('/');
        std::string dir = k == std::string::npos ? "." : p.substr(0, k);
        return dir + "/../resources";
    }
    return "resources";
}
static std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw RlnError("I/O error: cannot open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// record lengths of rln_witness_to_bytes_le / rln_proof_to_bytes_le for the handle's circuit
static size_t witness_record_len(const Rln& r) { return r.witness_record_len(); }
static size_t proof_record_len(const Rln& r) { return r.proof_record_len(); }
static std::vector<Witness> parse_records(Rln& r, const uint8_t* witnesses, size_t n, bool partial_phase) {
    const size_t d = r.depth(), rec = witness_record_len(r);
    std::vector<Witness> ws(n);
    for (size_t i = 0; i < n; i++) {
        if (partial_phase) {  // message ids / x / external_nullifier are not looked at: skip the range checks that involve them
            const uint8_t* b = witnesses + rec * i;
            const size_t po = r.multi() ? 65 : 97;  // offset of the path vector (single records carry message_id first)
            memcpy(ws[i].secret, b + 1, 32); memcpy(ws[i].limit, b + 33, 32);
            ws[i].path.assign(b + po + 8, b + po + 8 + 32 * d);
            ws[i].index.assign(b + po + 16 + 32 * d, b + po + 16 + 33 * d);
            ws[i].multi = r.multi();
            ws[i].mids.assign(32 * r.max_out(), 0);
            ws[i].sel.assign(r.multi() ? r.max_out() : 0, 0);
            memset(ws[i].x, 0, 32); memset(ws[i].ext_null, 0, 32);
            if (!fr_is_canonical(ws[i].secret) || !fr_is_canonical(ws[i].limit)) throw RlnError("Non-canonical field element: value is not in [0, r-1]");
        } else {
            size_t used = witness_from_bytes(witnesses + rec * i, rec, ws[i]);
            if (used != rec) throw RlnError(msg_read_len(used, rec));
        }
    }
    return ws;
}

extern "C" {

// ---- RLN object -------------------------------------------------------------------------------
// ffi_rln_new / ffi_rln_new_with_params read the tree configuration from the file `config_path` names; an unreadable file means
// the default configuration (rln/src/ffi/ffi_rln.rs:24-46), one above 1 MiB too (MAX_CONFIG_SIZE → the read fails → default)
static TreeConfig load_tree_config(const char* config_path) {
    if (!config_path || !*config_path) return TreeConfig();
    std::ifstream f(config_path, std::ios::binary);
    if (!f) return TreeConfig();
    std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (text.size() > (1u << 20) || text.empty()) return TreeConfig();
    try {
        return parse_tree_config(text);
    } catch (const std::exception& e) {
        throw RlnError(std::string("Configuration error: ") + e.what());
    }
}
CResult_FFI_RLN_t ffi_rln_new(size_t tree_depth, const char* config_path) {
    GUARD_BEGIN
    const TreeConfig cfg = load_tree_config(config_path);
    // bundled circuit resources, like the reference's include_bytes! (rln/src/circuit/mod.rs:29-78)
    std::ostringstream dir;
    dir << resources_dir() << "/tree_depth_" << tree_depth;
    std::vector<uint8_t> zkey = read_file(dir.str() + "/rln_final.arkzkey"), graph = read_file(dir.str() + "/graph.bin");
    FFI_RLN* h = new FFI_RLN();
    try {
        h->r = std::make_unique<Rln>(tree_depth, zkey.data(), zkey.size(), graph.data(), graph.size());
        DeviceGuard dg(h->r->device());
        h->r->attach_store(cfg);
    } catch (...) {
        delete h;
        throw;
    }
    return CResult_FFI_RLN_t{h, no_string()};
    GUARD_END(return (CResult_FFI_RLN_t{nullptr, mk_string(m)}))
}
CResult_FFI_RLN_t ffi_rln_new_with_params(size_t tree_depth, const Vec_uint8_t* zkey_data, const Vec_uint8_t* graph_data, const char* config_path) {
    GUARD_BEGIN
    const TreeConfig cfg = load_tree_config(config_path);
    FFI_RLN* h = new FFI_RLN();
    try {
        h->r = std::make_unique<Rln>(tree_depth, zkey_data->ptr, zkey_data->len, graph_data->ptr, graph_data->len);
        DeviceGuard dg(h->r->device());
        h->r->attach_store(cfg);
    } catch (...) {
        delete h;
        throw;
    }
    return CResult_FFI_RLN_t{h, no_string()};
    GUARD_END(return (CResult_FFI_RLN_t{nullptr, mk_string(m)}))
}
void ffi_rln_free(FFI_RLN_t* rln) { delete rln; }
size_t ffi_rln_get_tree_depth(FFI_RLN_t* const* rln) { return (*rln)->r->depth(); }
size_t ffi_rln_get_max_out(FFI_RLN_t* const* rln) { return (*rln)->r->max_out(); }
// bundled multi message-id circuit (rln/resources/tree_depth_20/multi_message_id/max_out_4; rln/src/circuit/mod.rs:36-42)
CResult_FFI_RLN_t rlnb200_rln_new_multi(size_t tree_depth, size_t max_out) {
    GUARD_BEGIN
    std::ostringstream dir;
    dir << resources_dir() << "/tree_depth_" << tree_depth << "/multi_message_id/max_out_" << max_out;
    std::vector<uint8_t> zkey = read_file(dir.str() + "/rln_final.arkzkey"), graph = read_file(dir.str() + "/graph.bin");
    FFI_RLN* h = new FFI_RLN();
    try {
        h->r = std::make_unique<Rln>(tree_depth, zkey.data(), zkey.size(), graph.data(), graph.size());
    } catch (...) {
        delete h;
        throw;
    }
    return CResult_FFI_RLN_t{h, no_string()};
    GUARD_END(return (CResult_FFI_RLN_t{nullptr, mk_string(m)}))
}

// ---- tree -------------------------------------------------------------------------------------
#define BOOL_OP(...)                                                      \
    GUARD_BEGIN                                                           \
    RlnLock lk(*(*rln)->r);                        \
    __VA_ARGS__;                                                          \
    return CBoolResult_t{true, no_string()};                              \
    GUARD_END(return (CBoolResult_t{false, mk_string(m)}))

CBoolResult_t ffi_set_tree(FFI_RLN_t** rln, size_t tree_depth) { BOOL_OP((*rln)->r->set_tree(tree_depth)) }
CBoolResult_t ffi_delete_leaf(FFI_RLN_t** rln, size_t index) { BOOL_OP((*rln)->r->delete_leaf(index)) }
CBoolResult_t ffi_set_leaf(FFI_RLN_t** rln, size_t index, const CFr_t* leaf) { BOOL_OP((*rln)->r->set_leaf(index, leaf->bytes)) }
CBoolResult_t ffi_set_next_leaf(FFI_RLN_t** rln, const CFr_t* leaf) { BOOL_OP((*rln)->r->set_next(leaf->bytes)) }
CBoolResult_t ffi_set_leaves_from(FFI_RLN_t** rln, size_t index, const Vec_CFr_t* leaves) {
    BOOL_OP((*rln)->r->override_range(index, (const uint8_t*)leaves->ptr, leaves->len, {}))
}
CBoolResult_t ffi_init_tree_with_leaves(FFI_RLN_t** rln, const Vec_CFr_t* leaves) {
    BOOL_OP((*rln)->r->set_tree((*rln)->r->tree_depth()); (*rln)->r->override_range(0, (const uint8_t*)leaves->ptr, leaves->len, {}))
}
CBoolResult_t ffi_atomic_operation(FFI_RLN_t** rln, size_t index, const Vec_CFr_t* leaves, const Vec_size_t* indices) {
    BOOL_OP((*rln)->r->override_range(index, (const uint8_t*)leaves->ptr, leaves->len, std::vector<size_t>(indices->ptr, indices->ptr + indices->len)))
}
CBoolResult_t ffi_seq_atomic_operation(FFI_RLN_t** rln, const Vec_CFr_t* leaves, const Vec_uint8_t* indices) {
    BOOL_OP(std::vector<size_t> idx(indices->ptr, indices->ptr + indices->len);
            (*rln)->r->override_range((*rln)->r->leaves_set(), (const uint8_t*)leaves->ptr, leaves->len, idx))
}
size_t ffi_leaves_set(FFI_RLN_t* const* rln) { return (*rln)->r->leaves_set(); }
CResult_CFr_t ffi_get_leaf(FFI_RLN_t* const* rln, size_t index) {
    GUARD_BEGIN
    RlnLock lk(*(*rln)->r);
    uint8_t b[32];
    (*rln)->r->get_leaf(index, b);
    return CResult_CFr_t{mk_cfr(b), no_string()};
    GUARD_END(return (CResult_CFr_t{nullptr, mk_string(m)}))
}
CFr_t* ffi_get_root(FFI_RLN_t* const* rln) {
    uint8_t b[32] = {0};
    try {
        RlnLock lk(*(*rln)->r);
        (*rln)->r->root(b);
    } catch (...) {
        abort();  // the reference's get_root is infallible
    }
    return mk_cfr(b);
}
CResult_FFI_MerkleProof_t ffi_get_merkle_proof(FFI_RLN_t* const* rln, size_t index) {
    GUARD_BEGIN
    RlnLock lk(*(*rln)->r);
    const size_t d = (*rln)->r->tree_depth();
    FFI_MerkleProof_t* mp = (FFI_MerkleProof_t*)malloc(sizeof(FFI_MerkleProof_t));
    mp->path_elements.ptr = (CFr_t*)malloc(32 * d);
    mp->path_elements.len = mp->path_elements.cap = d;
    mp->path_index.ptr = (uint8_t*)malloc(d);
    mp->path_index.len = mp->path_index.cap = d;
    uint64_t idx = index;
    try {
        (*rln)->r->merkle_proofs(&idx, 1, (uint8_t*)mp->path_elements.ptr, mp->path_index.ptr);
    } catch (...) {
        free(mp->path_elements.ptr);
        free(mp->path_index.ptr);
        free(mp);
        throw;
    }
    return CResult_FFI_MerkleProof_t{mp, no_string()};
    GUARD_END(return (CResult_FFI_MerkleProof_t{nullptr, mk_string(m)}))
}
void ffi_merkle_proof_free(FFI_MerkleProof_t* mp) {
    if (!mp) return;
    free(mp->path_elements.ptr);
    free(mp->path_index.ptr);
    free(mp);
}

// ---- witness input ----------------------------------------------------------------------------
CResult_FFI_RLNWitnessInput_t ffi_rln_witness_input_new_single(const CFr_t* identity_secret, const CFr_t* user_message_limit,
                                                               const CFr_t* message_id, const Vec_CFr_t* path_elements,
                                                               const Vec_uint8_t* identity_path_index, const CFr_t* x,
                                                               const CFr_t* external_nullifier) {
    GUARD_BEGIN
    auto w = std::make_unique<FFI_RLNWitnessInput>();
    memcpy(w->w.secret, identity_secret->bytes, 32);
    memcpy(w->w.limit, user_message_limit->bytes, 32);
    w->w.mids.assign(message_id->bytes, message_id->bytes + 32);
    memcpy(w->w.x, x->bytes, 32);
    memcpy(w->w.ext_null, external_nullifier->bytes, 32);
    w->w.path.assign((const uint8_t*)path_elements->ptr, (const uint8_t*)path_elements->ptr + 32 * path_elements->len);
    w->w.index.assign(identity_path_index->ptr, identity_path_index->ptr + identity_path_index->len);
    validate_witness(w->w);
    return CResult_FFI_RLNWitnessInput_t{w.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNWitnessInput_t{nullptr, mk_string(m)}))
}
CResult_FFI_RLNWitnessInput_t ffi_rln_witness_input_new_multi(const CFr_t* identity_secret, const CFr_t* user_message_limit,
                                                              const Vec_CFr_t* message_ids, const Vec_CFr_t* path_elements,
                                                              const Vec_uint8_t* identity_path_index, const CFr_t* x,
                                                              const CFr_t* external_nullifier, const Vec_bool_t* selector_used) {
    GUARD_BEGIN
    auto w = std::make_unique<FFI_RLNWitnessInput>();
    w->w.multi = true;
    memcpy(w->w.secret, identity_secret->bytes, 32);
    memcpy(w->w.limit, user_message_limit->bytes, 32);
    memcpy(w->w.x, x->bytes, 32);
    memcpy(w->w.ext_null, external_nullifier->bytes, 32);
    w->w.mids.assign((const uint8_t*)message_ids->ptr, (const uint8_t*)message_ids->ptr + 32 * message_ids->len);
    w->w.path.assign((const uint8_t*)path_elements->ptr, (const uint8_t*)path_elements->ptr + 32 * path_elements->len);
    w->w.index.assign(identity_path_index->ptr, identity_path_index->ptr + identity_path_index->len);
    w->w.sel.resize(selector_used->len);
    for (size_t i = 0; i < selector_used->len; i++) w->w.sel[i] = selector_used->ptr[i] ? 1 : 0;
    validate_witness(w->w);
    return CResult_FFI_RLNWitnessInput_t{w.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNWitnessInput_t{nullptr, mk_string(m)}))
}
CResult_Vec_uint8_t ffi_rln_witness_to_bytes_le(FFI_RLNWitnessInput_t* const* witness) {
    std::vector<uint8_t> b = witness_to_bytes((*witness)->w);
    return CResult_Vec_uint8_t{mk_vec(b.data(), b.size()), no_string()};
}
CResult_FFI_RLNWitnessInput_t ffi_bytes_le_to_rln_witness(const Vec_uint8_t* bytes) {
    GUARD_BEGIN
    auto w = std::make_unique<FFI_RLNWitnessInput>();
    witness_from_bytes(bytes->ptr, bytes->len, w->w);
    return CResult_FFI_RLNWitnessInput_t{w.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNWitnessInput_t{nullptr, mk_string(m)}))
}
void ffi_rln_witness_input_free(FFI_RLNWitnessInput_t* w) {
    if (w) memset(w->w.secret, 0, 32);  // IdSecret is zeroised on drop (rln/src/utils.rs:443-527)
    delete w;
}

// ---- proving / verifying ----------------------------------------------------------------------
// Concurrent single-item calls on one handle are run as device batches (coalesce.hpp).  Measured on a B200 with 16 caller threads
// (profiles/r02a_coalesce.txt): 28 → 158 proofs/s.  RLN_B200_COALESCE=0 puts every call back behind the handle's mutex.
static bool coalesce_enabled() {
    static const bool on = env_int("RLN_B200_COALESCE", 1) != 0;
    return on;
}
static void prove_alone(Rln& R, Rln::ProveReq* q) {   // caller holds R.mu
    try {
        std::vector<Witness> ws(1, *q->w);
        struct Wipe { std::vector<Witness>& v; ~Wipe() { for (Witness& w : v) memset(w.secret, 0, 32); } } wipe{ws};
        std::vector<RlnProof> out;
        R.prove_host(ws, q->rs, out);
        q->out = out[0];
    } catch (const CudaError& e) { q->failed = true; q->err = describe(e); }
    catch (const std::exception& e) { q->failed = true; q->err = describe(e); }
}
static void run_prove_batch(Rln& R, std::vector<Rln::ProveReq*>& b) {
    RlnLock lk(R);
    if (b.size() == 1) { prove_alone(R, b[0]); return; }
    // the cheap host-side checks of prove_host, per request: a caller with a witness of the wrong mode / shape fails alone and
    // never reaches the device batch
    std::vector<Rln::ProveReq*> good;
    good.reserve(b.size());
    for (Rln::ProveReq* q : b) {
        try {
            R.check_witness_shape(*q->w);
            good.push_back(q);
        } catch (const std::exception& e) { q->failed = true; q->err = describe(e); }
    }
    if (good.empty()) return;
    std::vector<Witness> ws;
    struct Wipe { std::vector<Witness>& v; ~Wipe() { for (Witness& w : v) memset(w.secret, 0, 32); } } wipe{ws};
    try {
        ws.reserve(good.size());
        std::vector<uint8_t> rs(64 * good.size());
        for (size_t i = 0; i < good.size(); i++) {
            ws.push_back(*good[i]->w);
            if (good[i]->rs) memcpy(&rs[64 * i], good[i]->rs, 64);
            else { random_fr(&rs[64 * i]); random_fr(&rs[64 * i + 32]); }   // r, s ← rng (proof.rs:743-745)
        }
        std::vector<RlnProof> out;
        R.prove_host(ws, rs.data(), out);
        for (size_t i = 0; i < good.size(); i++) good[i]->out = out[i];
    } catch (const CudaError& e) {   // a device failure is not one request's fault and is not retried: every request of the batch reports it
        for (Rln::ProveReq* q : good) { q->failed = true; q->err = describe(e); }
    } catch (const std::exception&) {
        // a graph-evaluation failure of some item (the only per-item failure left): each request once more on its own, so that
        // only the offending callers see an error.  Bounded: a batch holds at most max_batch requests of concurrent callers.
        for (Rln::ProveReq* q : good) prove_alone(R, q);
    }
}
static void run_pairing_batch(Rln& R, std::vector<Rln::PairingReq*>& b) {
    RlnLock lk(R);
    try {
        const size_t np = 32 * R.n_public();
        std::vector<uint8_t> proofs(128 * b.size()), pubs(np * b.size()), ok(b.size(), 0);
        for (size_t i = 0; i < b.size(); i++) {
            memcpy(&proofs[128 * i], b[i]->proof128, 128);
            memcpy(&pubs[np * i], b[i]->pub, np);
        }
        R.verify_batch(proofs.data(), pubs.data(), b.size(), ok.data());
        for (size_t i = 0; i < b.size(); i++) b[i]->ok = ok[i];
    } catch (const CudaError& e) { for (auto* q : b) { q->failed = true; q->err = describe(e); } }
    catch (const std::exception& e) { for (auto* q : b) { q->failed = true; q->err = describe(e); } }
}
// the pairing check of one proof; pub holds n_public() canonical field elements
static bool pairing_ok(Rln& R, const uint8_t* proof128, const uint8_t* pub) {
    if (coalesce_enabled()) {
        Rln::PairingReq q;
        q.proof128 = proof128; q.pub = pub;
        R.co_pairing.submit(q, R.max_batch(), [&R](std::vector<Rln::PairingReq*>& b) { run_pairing_batch(R, b); });
        if (q.failed) throw RlnError(q.err);
        return q.ok == 1;
    }
    RlnLock lk(R);
    uint8_t ok = 0;
    R.verify_batch(proof128, pub, 1, &ok);
    return ok == 1;
}
static CResult_FFI_RLNProof_t prove_one(FFI_RLN_t* const* rln, FFI_RLNWitnessInput_t* const* witness, const uint8_t* rs) {
    GUARD_BEGIN
    if (coalesce_enabled()) {
        Rln& R = *(*rln)->r;
        Rln::ProveReq q;
        q.w = &(*witness)->w; q.rs = rs;
        R.co_prove.submit(q, R.max_batch(), [&R](std::vector<Rln::ProveReq*>& b) { run_prove_batch(R, b); });
        if (q.failed) throw RlnError(q.err);
        auto p = std::make_unique<FFI_RLNProof>();
        p->p = q.out;
        return CResult_FFI_RLNProof_t{p.release(), no_string()};
    }
    RlnLock lk(*(*rln)->r);
    std::vector<Witness> ws(1, (*witness)->w);
    WitnessWipe wipe{ws};
    std::vector<RlnProof> out;
    (*rln)->r->prove_host(ws, rs, out);
    auto p = std::make_unique<FFI_RLNProof>();
    p->p = out[0];
    return CResult_FFI_RLNProof_t{p.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNProof_t{nullptr, mk_string(m)}))
}
CResult_FFI_RLNProof_t ffi_generate_rln_proof(FFI_RLN_t* const* rln, FFI_RLNWitnessInput_t* const* witness) {
    return prove_one(rln, witness, nullptr);
}
CResult_FFI_RLNProof_t rlnb200_generate_rln_proof_with_rs(FFI_RLN_t* const* rln, FFI_RLNWitnessInput_t* const* witness, const CFr_t* r,
                                                          const CFr_t* s) {
    uint8_t rs[64];
    memcpy(rs, r->bytes, 32);
    memcpy(rs + 32, s->bytes, 32);
    return prove_one(rln, witness, rs);
}

// ---- two-phase proving (rln/src/ffi/ffi_rln.rs:561-712, 238-320, 918-960) ----------------------------
CResult_FFI_RLNPartialWitnessInput_t ffi_rln_partial_witness_input_new(const CFr_t* identity_secret, const CFr_t* user_message_limit,
                                                                       const Vec_CFr_t* path_elements, const Vec_uint8_t* identity_path_index) {
    GUARD_BEGIN
    auto w = std::make_unique<FFI_RLNPartialWitnessInput>();
    memset(&w->w.x, 0, 32); memset(&w->w.ext_null, 0, 32);
    memcpy(w->w.secret, identity_secret->bytes, 32);
    memcpy(w->w.limit, user_message_limit->bytes, 32);
    w->w.path.assign((const uint8_t*)path_elements->ptr, (const uint8_t*)path_elements->ptr + 32 * path_elements->len);
    w->w.index.assign(identity_path_index->ptr, identity_path_index->ptr + identity_path_index->len);
    if (is_zero32(w->w.limit)) throw RlnError("User message limit cannot be zero");   // RLNPartialWitnessInput::new (witness.rs)
    if (w->w.path.size() / 32 != w->w.index.size()) {
        std::ostringstream m;
        m << "Merkle proof length mismatch: expected " << w->w.path.size() / 32 << ", got " << w->w.index.size();
        throw RlnError(m.str());
    }
    return CResult_FFI_RLNPartialWitnessInput_t{w.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNPartialWitnessInput_t{nullptr, mk_string(m)}))
}
void ffi_rln_partial_witness_input_free(FFI_RLNPartialWitnessInput_t* w) {
    if (w) memset(w->w.secret, 0, 32);
    delete w;
}
CResult_FFI_RLNPartialProof_t ffi_generate_partial_zk_proof(FFI_RLN_t* const* rln, FFI_RLNPartialWitnessInput_t* const* partial_witness) {
    GUARD_BEGIN
    RlnLock lk(*(*rln)->r);
    std::vector<Witness> ws(1, (*partial_witness)->w);
    WitnessWipe wipe{ws};
    ws[0].multi = (*rln)->r->multi();   // a partial witness carries no message ids: shape them for the circuit at hand
    ws[0].mids.assign(32 * (*rln)->r->max_out(), 0);
    ws[0].sel.assign(ws[0].multi ? (*rln)->r->max_out() : 0, 0);
    std::vector<PartialProofHost> out;
    (*rln)->r->partial_host(ws, out);
    memset(ws[0].secret, 0, 32);
    auto p = std::make_unique<FFI_RLNPartialProof>();
    p->p = out[0];
    p->mask = (*rln)->r->partial_mask();
    return CResult_FFI_RLNPartialProof_t{p.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNPartialProof_t{nullptr, mk_string(m)}))
}
static CResult_FFI_RLNProof_t finish_one(FFI_RLN_t* const* rln, FFI_RLNPartialProof_t* const* partial, FFI_RLNWitnessInput_t* const* witness,
                                         const uint8_t* rs) {
    GUARD_BEGIN
    RlnLock lk(*(*rln)->r);
    if ((*partial)->mask != (*rln)->r->partial_mask()) throw RlnError("Protocol error: Error producing proof: malformed verifying key");
    std::vector<Witness> ws(1, (*witness)->w);
    WitnessWipe wipe{ws};
    std::vector<RlnProof> out;
    (*rln)->r->prove_host(ws, rs, out, &(*partial)->p);
    auto p = std::make_unique<FFI_RLNProof>();
    p->p = out[0];
    return CResult_FFI_RLNProof_t{p.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNProof_t{nullptr, mk_string(m)}))
}
CResult_FFI_RLNProof_t ffi_finish_rln_proof(FFI_RLN_t* const* rln, FFI_RLNPartialProof_t* const* partial_proof,
                                            FFI_RLNWitnessInput_t* const* witness) {
    return finish_one(rln, partial_proof, witness, nullptr);
}
CResult_FFI_RLNProof_t rlnb200_finish_rln_proof_with_rs(FFI_RLN_t* const* rln, FFI_RLNPartialProof_t* const* partial_proof,
                                                        FFI_RLNWitnessInput_t* const* witness, const CFr_t* r, const CFr_t* s) {
    uint8_t rs[64];
    memcpy(rs, r->bytes, 32);
    memcpy(rs + 32, s->bytes, 32);
    return finish_one(rln, partial_proof, witness, rs);
}
// rln_partial_proof_to_bytes_le (proof.rs:537-547): version | u64 mask length | mask bytes | π_a | ρ | π_b | π_c (ark compressed)
CResult_Vec_uint8_t ffi_rln_partial_proof_to_bytes_le(FFI_RLNPartialProof_t* const* partial_proof) {
    const FFI_RLNPartialProof& p = **partial_proof;
    std::vector<uint8_t> b;
    b.push_back(0);
    uint64_t n = p.mask.size();
    b.insert(b.end(), (uint8_t*)&n, (uint8_t*)&n + 8);
    b.insert(b.end(), p.mask.begin(), p.mask.end());
    b.insert(b.end(), p.p.comp, p.p.comp + 160);
    return CResult_Vec_uint8_t{mk_vec(b.data(), b.size()), no_string()};
}
CResult_FFI_RLNPartialProof_t rlnb200_bytes_le_to_rln_partial_proof(FFI_RLN_t* const* rln, const Vec_uint8_t* bytes) {
    GUARD_BEGIN
    RlnLock lk(*(*rln)->r);
    const uint8_t* b = bytes->ptr;
    const size_t len = bytes->len;
    if (len == 0) throw RlnError(msg_read_len(1, 0));
    if (b[0] != 0) {
        char t[64];
        snprintf(t, sizeof t, "Unknown message mode version byte: %#04x", b[0]);
        throw RlnError(t);
    }
    if (len < 9) throw RlnError("Proof serialization error: io error: failed to fill whole buffer");
    uint64_t n;
    memcpy(&n, b + 1, 8);
    if (n > len || 9 + n + 160 > len) throw RlnError("Proof serialization error: io error: failed to fill whole buffer");
    auto p = std::make_unique<FFI_RLNPartialProof>();
    p->mask.assign(b + 9, b + 9 + n);
    for (uint8_t v : p->mask)
        if (v > 1) throw RlnError("Proof serialization error: the input buffer contained invalid data");
    memcpy(p->p.comp, b + 9 + n, 160);
    (*rln)->r->decompress_partial(p->p.comp, p->p.affine);
    if (9 + n + 160 != len) throw RlnError(msg_read_len(9 + n + 160, len));
    return CResult_FFI_RLNPartialProof_t{p.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNPartialProof_t{nullptr, mk_string(m)}))
}
void ffi_rln_partial_proof_free(FFI_RLNPartialProof_t* p) { delete p; }
// batched two-phase proving on host buffers: witness records as for rlnb200_prove_batch (the unknown fields of the partial
// phase are ignored), partial points n × 320 bytes (canonical affine π_a | ρ | π_b | π_c)
int rlnb200_partial_batch(FFI_RLN_t* const* rln, const uint8_t* witnesses, size_t n, uint8_t* partial_out, RlnString* err);
int rlnb200_finish_batch(FFI_RLN_t* const* rln, const uint8_t* witnesses, size_t n, const uint8_t* partial, const uint8_t* rs,
                         uint8_t* proofs_out, RlnString* err);

// verify_zk_proof (proof.rs:856-894) then root / signal checks (public.rs:725-771)
static CBoolResult_t verify_common(FFI_RLN_t* const* rln, const RlnProof& p, const uint8_t* x, const Vec_CFr_t* roots, bool use_tree_root) {
    GUARD_BEGIN
    Rln& R = *(*rln)->r;
    std::vector<uint8_t> pub = public_inputs(p.pv);  // circuit order (proof.rs:863-884)
    if (pub.size() != 32 * R.n_public())
        throw RlnError("Protocol error: Error producing proof: malformed verifying key");  // SynthesisError::MalformedVerifyingKey
    if (!pairing_ok(R, p.proof, pub.data())) throw RlnError("Verification error: Invalid proof provided");
    if (use_tree_root) {
        uint8_t root[32];
        {
            RlnLock lk(R);
            R.root(root);
        }
        if (memcmp(root, p.pv.root, 32)) throw RlnError("Verification error: Expected one of the provided roots");
    } else if (roots && roots->len) {
        bool found = false;
        for (size_t i = 0; i < roots->len; i++) found = found || !memcmp(roots->ptr[i].bytes, p.pv.root, 32);
        if (!found) throw RlnError("Verification error: Expected one of the provided roots");
    }
    if (memcmp(x, p.pv.x, 32)) throw RlnError("Verification error: Signal value does not match");
    return CBoolResult_t{true, no_string()};
    GUARD_END(return (CBoolResult_t{false, mk_string(m)}))
}
CBoolResult_t ffi_verify_rln_proof(FFI_RLN_t* const* rln, FFI_RLNProof_t* const* rln_proof, const CFr_t* x) {
    return verify_common(rln, (*rln_proof)->p, x->bytes, nullptr, true);
}
CBoolResult_t ffi_verify_with_roots(FFI_RLN_t* const* rln, FFI_RLNProof_t* const* rln_proof, const Vec_CFr_t* roots, const CFr_t* x) {
    return verify_common(rln, (*rln_proof)->p, x->bytes, roots, false);
}

FFI_RLNProofValues_t* ffi_rln_proof_get_values(FFI_RLNProof_t* const* rln_proof) {
    FFI_RLNProofValues* v = new FFI_RLNProofValues();
    v->v = (*rln_proof)->p.pv;
    return v;
}
uint8_t ffi_rln_proof_get_version_byte(FFI_RLNProof_t* const* rln_proof) { return (*rln_proof)->p.pv.multi ? 1 : 0; }
CResult_Vec_uint8_t ffi_rln_proof_to_bytes_le(FFI_RLNProof_t* const* rln_proof) {
    std::vector<uint8_t> b = rln_proof_to_bytes((*rln_proof)->p);
    return CResult_Vec_uint8_t{mk_vec(b.data(), b.size()), no_string()};
}
// proof stays LE (arkworks), values big-endian (rln/src/protocol/proof.rs:430-446)
CResult_Vec_uint8_t ffi_rln_proof_to_bytes_be(FFI_RLNProof_t* const* rln_proof) {
    const ProofValues& pv = (*rln_proof)->p.pv;
    std::vector<uint8_t> b = rln_proof_to_bytes((*rln_proof)->p);
    // every field element and every vector length prefix is big-endian in the BE form (rln/src/utils.rs:141-230)
    size_t o = 130;
    auto rev = [&](size_t n) { std::reverse(b.begin() + o, b.begin() + o + n); o += n; };
    rev(32); rev(32); rev(32);
    if (!pv.multi) { rev(32); rev(32); }
    else {
        for (int v = 0; v < 2; v++) { rev(8); for (size_t i = 0; i < pv.k(); i++) rev(32); }
        rev(8);
    }
    return CResult_Vec_uint8_t{mk_vec(b.data(), b.size()), no_string()};
}
CResult_FFI_RLNProof_t ffi_bytes_le_to_rln_proof(const Vec_uint8_t* bytes) {
    GUARD_BEGIN
    global_init();
    if (bytes->len == 0) throw RlnError(msg_read_len(1, 0));
    if (bytes->ptr[0] > 1) throw RlnError(msg_mode(bytes->ptr[0]));
    if (bytes->len < 129) throw RlnError(msg_read_len(129, bytes->len));
    auto p = std::make_unique<FFI_RLNProof>();
    memcpy(p->p.proof, bytes->ptr + 1, 128);
    {   // Proof::deserialize_compressed validates curve / subgroup membership (proof.rs:469)
        DevMem dp, da, dok;
        dp.upload(p->p.proof, 128);
        da.alloc(256);
        dok.alloc(1);
        launch_decompress(dp.as<uint8_t>(), 1, da.as<uint8_t>(), dok.as<uint8_t>(), 0);
        g_launch_count++;
        uint8_t ok = 0;
        ZK_CUDA_CHECK(cudaMemcpy(&ok, dok.p, 1, cudaMemcpyDeviceToHost));
        if (!ok) throw RlnError("Proof serialization error: the input buffer contained invalid data");
    }
    size_t used = 129 + proof_values_from_bytes(bytes->ptr + 129, bytes->len - 129, p->p.pv);
    if (used != bytes->len) throw RlnError(msg_read_len(used, bytes->len));
    return CResult_FFI_RLNProof_t{p.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNProof_t{nullptr, mk_string(m)}))
}
void ffi_rln_proof_free(FFI_RLNProof_t* p) { delete p; }

CFr_t* ffi_rln_proof_values_get_root(FFI_RLNProofValues_t* const* pv) { return mk_cfr((*pv)->v.root); }
CFr_t* ffi_rln_proof_values_get_x(FFI_RLNProofValues_t* const* pv) { return mk_cfr((*pv)->v.x); }
CFr_t* ffi_rln_proof_values_get_external_nullifier(FFI_RLNProofValues_t* const* pv) { return mk_cfr((*pv)->v.ext_null); }
// variant-mismatched getters return the reference's error (rln/src/error.rs:92-99)
static RlnString variant_err(const char* field, const char* variant) {
    return mk_string(std::string("Field `") + field + "` does not exist on the `" + variant + "` variant");
}
static Vec_CFr_t mk_vec_cfr(const std::vector<uint8_t>& b) {
    Vec_CFr_t v = ffi_vec_cfr_new(b.size() / 32);
    memcpy(v.ptr, b.data(), b.size());
    v.len = b.size() / 32;
    return v;
}
CResult_CFr_t ffi_rln_proof_values_get_y(FFI_RLNProofValues_t* const* pv) {
    if ((*pv)->v.multi) return CResult_CFr_t{nullptr, variant_err("y", "MultiV1")};
    return CResult_CFr_t{mk_cfr((*pv)->v.ys.data()), no_string()};
}
CResult_CFr_t ffi_rln_proof_values_get_nullifier(FFI_RLNProofValues_t* const* pv) {
    if ((*pv)->v.multi) return CResult_CFr_t{nullptr, variant_err("nullifier", "MultiV1")};
    return CResult_CFr_t{mk_cfr((*pv)->v.nulls.data()), no_string()};
}
CResult_Vec_CFr_t ffi_rln_proof_values_get_ys(FFI_RLNProofValues_t* const* pv) {
    if (!(*pv)->v.multi) return CResult_Vec_CFr_t{Vec_CFr_t{nullptr, 0, 0}, variant_err("ys", "SingleV1")};
    return CResult_Vec_CFr_t{mk_vec_cfr((*pv)->v.ys), no_string()};
}
CResult_Vec_CFr_t ffi_rln_proof_values_get_nullifiers(FFI_RLNProofValues_t* const* pv) {
    if (!(*pv)->v.multi) return CResult_Vec_CFr_t{Vec_CFr_t{nullptr, 0, 0}, variant_err("nullifiers", "SingleV1")};
    return CResult_Vec_CFr_t{mk_vec_cfr((*pv)->v.nulls), no_string()};
}
CResult_Vec_uint8_t ffi_rln_proof_values_get_selector_used(FFI_RLNProofValues_t* const* pv) {
    if (!(*pv)->v.multi) return CResult_Vec_uint8_t{Vec_uint8_t{nullptr, 0, 0}, variant_err("selector_used", "SingleV1")};
    return CResult_Vec_uint8_t{mk_vec((*pv)->v.sel.data(), (*pv)->v.sel.size()), no_string()};
}
uint8_t ffi_rln_proof_values_get_version_byte(FFI_RLNProofValues_t* const* pv) { return (*pv)->v.multi ? 1 : 0; }
Vec_uint8_t ffi_rln_proof_values_to_bytes_le(FFI_RLNProofValues_t* const* pv) {
    std::vector<uint8_t> b = proof_values_to_bytes((*pv)->v);
    return mk_vec(b.data(), b.size());
}
CResult_FFI_RLNProofValues_t ffi_bytes_le_to_rln_proof_values(const Vec_uint8_t* bytes) {
    GUARD_BEGIN
    auto v = std::make_unique<FFI_RLNProofValues>();
    size_t used = proof_values_from_bytes(bytes->ptr, bytes->len, v->v);
    if (used != bytes->len) throw RlnError(msg_read_len(used, bytes->len));
    return CResult_FFI_RLNProofValues_t{v.release(), no_string()};
    GUARD_END(return (CResult_FFI_RLNProofValues_t{nullptr, mk_string(m)}))
}
void ffi_rln_proof_values_free(FFI_RLNProofValues_t* v) { delete v; }

// ---- CFr / Vec helpers --------------------------------------------------------------------------
CFr_t* ffi_cfr_zero(void) { uint8_t b[32] = {0}; return mk_cfr(b); }
CFr_t* ffi_cfr_one(void) { uint8_t b[32] = {1}; return mk_cfr(b); }
CFr_t* ffi_uint_to_cfr(uint32_t value) { uint8_t b[32] = {0}; memcpy(b, &value, 4); return mk_cfr(b); }
CResult_Vec_uint8_t ffi_cfr_to_bytes_le(const CFr_t* cfr) { return CResult_Vec_uint8_t{mk_vec(cfr->bytes, 32), no_string()}; }
CResult_Vec_uint8_t ffi_cfr_to_bytes_be(const CFr_t* cfr) {
    uint8_t b[32];
    for (int i = 0; i < 32; i++) b[i] = cfr->bytes[31 - i];
    return CResult_Vec_uint8_t{mk_vec(b, 32), no_string()};
}
CResult_CFr_t ffi_bytes_le_to_cfr(const Vec_uint8_t* bytes) {
    if (bytes->len < 32) return CResult_CFr_t{nullptr, mk_string("io error: failed to fill whole buffer")};
    if (!fr_is_canonical(bytes->ptr)) return CResult_CFr_t{nullptr, mk_string("the input buffer contained invalid data")};
    return CResult_CFr_t{mk_cfr(bytes->ptr), no_string()};
}
CResult_CFr_t ffi_bytes_be_to_cfr(const Vec_uint8_t* bytes) {
    if (bytes->len < 32) return CResult_CFr_t{nullptr, mk_string(msg_read_len(32, bytes->len))};
    uint8_t b[32];
    for (int i = 0; i < 32; i++) b[i] = bytes->ptr[31 - i];
    if (!fr_is_canonical(b)) return CResult_CFr_t{nullptr, mk_string("Non-canonical field element: value is not in [0, r-1]")};
    return CResult_CFr_t{mk_cfr(b), no_string()};
}
RlnString ffi_cfr_debug(const CFr_t* cfr) { return mk_string(cfr ? decimal_le32(cfr->bytes) : std::string("None")); }
void ffi_cfr_free(CFr_t* cfr) { free(cfr); }
Vec_CFr_t ffi_vec_cfr_new(size_t capacity) {
    Vec_CFr_t v;
    v.len = 0;
    v.cap = capacity ? capacity : 1;
    v.ptr = (CFr_t*)malloc(sizeof(CFr_t) * v.cap);
    return v;
}
Vec_CFr_t ffi_vec_cfr_from_cfr(const CFr_t* cfr) {
    Vec_CFr_t v = ffi_vec_cfr_new(1);
    v.ptr[0] = *cfr;
    v.len = 1;
    return v;
}
void ffi_vec_cfr_push(Vec_CFr_t* v, const CFr_t* cfr) {
    if (v->len == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 4;
        v->ptr = (CFr_t*)realloc(v->ptr, sizeof(CFr_t) * v->cap);
    }
    v->ptr[v->len++] = *cfr;
}
size_t ffi_vec_cfr_len(const Vec_CFr_t* v) { return v->len; }
const CFr_t* ffi_vec_cfr_get(const Vec_CFr_t* v, size_t i) { return i < v->len ? &v->ptr[i] : nullptr; }
void ffi_vec_cfr_free(Vec_CFr_t v) { free(v.ptr); }
void ffi_vec_u8_free(Vec_uint8_t v) { free(v.ptr); }
void ffi_c_string_free(RlnString s) { free(s.ptr); }

CFr_t* ffi_hash_to_field_le(const Vec_uint8_t* input) {  // hashers.rs:73-81
    uint8_t h[32];
    keccak256(input->ptr, input->len, h);
    fr_reduce(h);
    return mk_cfr(h);
}
CFr_t* ffi_hash_to_field_be(const Vec_uint8_t* input) { return ffi_hash_to_field_le(input); }  // same Fr (hashers.rs:84-93)

static void device_poseidon(const uint8_t* in, int n, uint8_t* out) {
    global_init();
    DevMem di, dout;
    di.upload(in, 32 * (size_t)n);
    dout.alloc(32);
    launch_poseidon_n(di.as<uint8_t>(), n, dout.as<uint8_t>(), 0);
    g_launch_count++;
    ZK_CUDA_CHECK(cudaMemcpy(out, dout.p, 32, cudaMemcpyDeviceToHost));
}
CFr_t* ffi_poseidon_hash_pair(const CFr_t* a, const CFr_t* b) {
    uint8_t in[64], out[32];
    memcpy(in, a->bytes, 32);
    memcpy(in + 32, b->bytes, 32);
    try {
        device_poseidon(in, 2, out);
    } catch (...) {
        abort();  // the reference's poseidon_hash_pair is infallible
    }
    return mk_cfr(out);
}
Vec_CFr_t ffi_key_gen(void) {  // keygen (rln/src/protocol/keygen.rs:20-30): secret ← rng, commitment = H(secret)
    Vec_CFr_t v = ffi_vec_cfr_new(2);
    random_fr(v.ptr[0].bytes);
    try {
        device_poseidon(v.ptr[0].bytes, 1, v.ptr[1].bytes);
    } catch (...) {
        abort();
    }
    v.len = 2;
    return v;
}

#include "rln_ffi_more.inc"
#include "rln_ffi_v3.inc"

// ---- extensions -------------------------------------------------------------------------------
#define INT_OP(...)                                            \
    GUARD_BEGIN                                                \
    __VA_ARGS__;                                               \
    return 0;                                                  \
    GUARD_END(if (err) *err = mk_string(m); return -1)

int rlnb200_prove_batch(FFI_RLN_t* const* rln, const uint8_t* witnesses, size_t n, const uint8_t* rs, uint8_t* proofs_out, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r); (*rln)->r->prove_records_host(witnesses, rs, n, proofs_out);)
}
int rlnb200_prove_records_device(FFI_RLN_t* const* rln, const void* d_witness_records, const void* d_rs, size_t n, void* d_proof_records,
                                 void* stream, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r);
           (*rln)->r->prove_records_device((const uint8_t*)d_witness_records, (const uint8_t*)d_rs, n, (uint8_t*)d_proof_records, (cudaStream_t)stream);)
}
int rlnb200_partial_batch(FFI_RLN_t* const* rln, const uint8_t* witnesses, size_t n, uint8_t* partial_out, RlnString* err) {
    INT_OP(
        RlnLock lk(*(*rln)->r);
        std::vector<Witness> ws = parse_records(*(*rln)->r, witnesses, n, true);
        std::vector<PartialProofHost> out;
        (*rln)->r->partial_host(ws, out);
        for (auto& w : ws) memset(w.secret, 0, 32);
        for (size_t i = 0; i < n; i++) memcpy(partial_out + 320 * i, out[i].affine, 320);)
}
int rlnb200_finish_batch(FFI_RLN_t* const* rln, const uint8_t* witnesses, size_t n, const uint8_t* partial, const uint8_t* rs,
                         uint8_t* proofs_out, RlnString* err) {
    INT_OP(
        RlnLock lk(*(*rln)->r);
        std::vector<Witness> ws = parse_records(*(*rln)->r, witnesses, n, false);
        std::vector<PartialProofHost> pp(n);
        for (size_t i = 0; i < n; i++) memcpy(pp[i].affine, partial + 320 * i, 320);
        std::vector<RlnProof> out;
        (*rln)->r->prove_host(ws, rs, out, pp.data());
        for (auto& w : ws) memset(w.secret, 0, 32);
        const size_t orec = proof_record_len(*(*rln)->r);
        for (size_t i = 0; i < n; i++) { std::vector<uint8_t> b = rln_proof_to_bytes(out[i]); memcpy(proofs_out + orec * i, b.data(), orec); })
}
int rlnb200_verify_batch(FFI_RLN_t* const* rln, const uint8_t* proofs, size_t n, uint8_t* ok_out, RlnString* err) {
    INT_OP(
        RlnLock lk(*(*rln)->r);
        const size_t orec = proof_record_len(*(*rln)->r), np = (*rln)->r->n_public();
        std::vector<uint8_t> p(128 * n), pub(32 * np * n);
        for (size_t i = 0; i < n; i++) {
            const uint8_t* b = proofs + orec * i;
            memcpy(p.data() + 128 * i, b + 1, 128);
            ProofValues pv;
            proof_values_from_bytes(b + 129, orec - 129, pv);
            std::vector<uint8_t> q = public_inputs(pv);
            if (q.size() != 32 * np) throw RlnError("Protocol error: proof record does not match the circuit's message mode");
            memcpy(pub.data() + 32 * np * i, q.data(), q.size());
        }
        (*rln)->r->verify_batch(p.data(), pub.data(), n, ok_out);)
}
int rlnb200_prove_batch_device(FFI_RLN_t* const* rln, const void* d_inputs, const void* d_rs, size_t n, void* d_proofs, void* d_values,
                               void* d_affine, void* stream, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r);
           (*rln)->r->prove_device((const uint8_t*)d_inputs, (const uint8_t*)d_rs, n, (uint8_t*)d_proofs, (uint8_t*)d_values, (uint8_t*)d_affine,
                                   (cudaStream_t)stream);)
}
int rlnb200_partial_batch_device(FFI_RLN_t* const* rln, const void* d_inputs, size_t n, void* d_partial_affine, void* d_partial_compressed,
                                 void* stream, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r);
           (*rln)->r->prove_device((const uint8_t*)d_inputs, nullptr, n, nullptr, nullptr, nullptr, (cudaStream_t)stream, MSM_KNOWN, nullptr,
                                   (uint8_t*)d_partial_affine, (uint8_t*)d_partial_compressed);)
}
int rlnb200_finish_batch_device(FFI_RLN_t* const* rln, const void* d_inputs, const void* d_rs, const void* d_partial_affine, size_t n,
                                void* d_proofs, void* d_values, void* stream, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r);
           (*rln)->r->prove_device((const uint8_t*)d_inputs, (const uint8_t*)d_rs, n, (uint8_t*)d_proofs, (uint8_t*)d_values, nullptr,
                                   (cudaStream_t)stream, MSM_UNKNOWN, (const uint8_t*)d_partial_affine);)
}
int rlnb200_witness_to_input_slots(FFI_RLN_t* const* rln, const uint8_t* witness_le, size_t len, uint8_t* slots_out, RlnString* err) {
    INT_OP(Witness w; witness_from_bytes(witness_le, len, w);
           if (w.path.size() / 32 != (*rln)->r->depth() || w.index.size() != (*rln)->r->depth()) throw RlnError("Protocol error: witness depth does not match the circuit");
           (*rln)->r->witness_slots(w, slots_out);)
}
size_t rlnb200_input_slots(FFI_RLN_t* const* rln) { return (*rln)->r->n_slots(); }
size_t rlnb200_state_tree_depth(FFI_RLN_t* const* rln) { return (*rln)->r->tree_depth(); }
int rlnb200_input_slot(FFI_RLN_t* const* rln, const char* name, uint32_t* offset, uint32_t* len) {
    auto& m = (*rln)->r->graph().inputs;
    auto it = m.find(name);
    if (it == m.end()) return 0;
    *offset = it->second.first;
    *len = it->second.second;
    return 1;
}
int rlnb200_reserve(FFI_RLN_t* const* rln, size_t max_batch, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r); (*rln)->r->reserve(max_batch);)
}
uint64_t rlnb200_launch_count(void) { return g_launch_count.load(); }
void rlnb200_last_stage_ms(FFI_RLN_t* const* rln, float out[8]) {
    for (int i = 0; i < 8; i++) out[i] = (*rln)->r->stage_ms[i];
}
uint32_t rlnb200_last_stage_batches(FFI_RLN_t* const* rln) { return (*rln)->r->last_chunks; }
int rlnb200_set_device(int device, RlnString* err) {
    INT_OP(ZK_CUDA_CHECK(cudaSetDevice(device));)
}
int rlnb200_table_info(FFI_RLN_t* const* rln, int* window_bits, int* windows, uint64_t* g1_bases, uint64_t* g2_bases, uint64_t* table_bytes,
                       int* window_bits_g2, int* windows_g2) {
    (*rln)->r->table_info(window_bits, windows, g1_bases, g2_bases, table_bytes, window_bits_g2, windows_g2);
    return 0;
}
int rlnb200_set_verify_vm_max(FFI_RLN_t* const* rln, size_t max_batch) {
    std::lock_guard<std::mutex> lk((*rln)->r->mu);
    return (*rln)->r->set_verify_vm_max(max_batch) ? 0 : 1;
}
int rlnb200_verify_vm_trace(FFI_RLN_t* const* rln, const uint8_t* proof_record, long long* cycles, uint32_t* meta, uint8_t* ok_out, RlnString* err) {
    INT_OP(
        RlnLock lk(*(*rln)->r);
        const size_t orec = proof_record_len(*(*rln)->r), np = (*rln)->r->n_public();
        ProofValues pv;
        proof_values_from_bytes(proof_record + 129, orec - 129, pv);
        std::vector<uint8_t> q = public_inputs(pv);
        if (q.size() != 32 * np) throw RlnError("Protocol error: proof record does not match the circuit's message mode");
        (*rln)->r->verify_vm_trace(proof_record + 1, q.data(), cycles, meta, ok_out);)
}
int rlnb200_verify_vm_info(FFI_RLN_t* const* rln, uint32_t* levels, uint32_t* slots, uint32_t* constants) {
    (*rln)->r->verify_vm_info(levels, slots, constants);
    return 0;
}
int rlnb200_set_leaves_from_bytes(FFI_RLN_t** rln, size_t index, const uint8_t* leaves_le, size_t count, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r); (*rln)->r->set_range_host(index, leaves_le, count);)
}
int rlnb200_set_leaves_from_device(FFI_RLN_t** rln, size_t index, const void* d_leaves, size_t count, void* stream, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r); (*rln)->r->set_range_device(index, (const uint8_t*)d_leaves, count, (cudaStream_t)stream);)
}
int rlnb200_get_merkle_proofs(FFI_RLN_t* const* rln, const uint64_t* indices, size_t n, uint8_t* elements_out, uint8_t* index_bits_out, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r); (*rln)->r->merkle_proofs(indices, n, elements_out, index_bits_out);)
}
int rlnb200_debug_witness_and_h(FFI_RLN_t* const* rln, const uint8_t* witness_le, size_t len, uint8_t* w_out, uint8_t* h_out, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r); Witness w; witness_from_bytes(witness_le, len, w); (*rln)->r->debug_w_h(w, w_out, h_out);)
}
int rlnb200_get_subtree_root(FFI_RLN_t* const* rln, size_t level, size_t index, uint8_t* out32, RlnString* err) {
    INT_OP(RlnLock lk(*(*rln)->r); (*rln)->r->subtree_root(level, index, out32))
}
int rlnb200_get_empty_leaves_indices(FFI_RLN_t* const* rln, Vec_size_t* out, RlnString* err) {
    INT_OP(
        RlnLock lk(*(*rln)->r);
        std::vector<size_t> v = (*rln)->r->empty_leaves_indices();
        out->len = v.size();
        out->cap = v.size() ? v.size() : 1;
        out->ptr = (size_t*)malloc(sizeof(size_t) * out->cap);
        if (!v.empty()) memcpy(out->ptr, v.data(), sizeof(size_t) * v.size());
    )
}
void rlnb200_vec_usize_free(Vec_size_t v) { free(v.ptr); }
int rlnb200_glv_enabled(FFI_RLN_t* const* rln) { return (*rln)->r->glv() ? 1 : 0; }
int rlnb200_glv_split(const uint8_t* scalars_le, size_t n, uint8_t* out36, RlnString* err) {
    INT_OP(
        global_init();
        DevMem d_in; DevMem d_out;
        d_in.upload(scalars_le, 32 * n);
        d_out.alloc(36 * n);
        launch_glv_split(d_in.as<uint8_t>(), n, d_out.as<uint8_t>(), 0);
        g_launch_count++;
        ZK_CUDA_CHECK(cudaMemcpy(out36, d_out.p, 36 * n, cudaMemcpyDeviceToHost));
    )
}
int rlnb200_glv_double_mul(const uint8_t* items192, size_t n, int use_q, uint8_t* out64, RlnString* err) {
    INT_OP(
        global_init();
        DevMem d_in; DevMem d_out;
        d_in.upload(items192, 192 * n);
        d_out.alloc(64 * n);
        launch_glv_double_mul(d_in.as<uint8_t>(), n, use_q, d_out.as<uint8_t>(), 0);
        g_launch_count++;
        ZK_CUDA_CHECK(cudaMemcpy(out64, d_out.p, 64 * n, cudaMemcpyDeviceToHost));
    )
}
size_t rlnb200_num_wires(FFI_RLN_t* const* rln) { return (*rln)->r->n_wires(); }
size_t rlnb200_witness_record_len(FFI_RLN_t* const* rln) { return witness_record_len(*(*rln)->r); }
size_t rlnb200_proof_record_len(FFI_RLN_t* const* rln) { return proof_record_len(*(*rln)->r); }
size_t rlnb200_domain_size(FFI_RLN_t* const* rln) { return (*rln)->r->domain(); }


// ---- one batch, every GPU of the box (BASELINE.json configs[4]) -----------------------------------------------------------
// A multi-device prover is one replica of the RLN object per device (immutable data — tables, matrices, graph program — is
// rebuilt locally on every GPU: 2.6 s, cheaper than moving 125 GiB) inside ONE process.  A batch of independent proofs is cut
// into contiguous shards, one worker thread per device runs the device-records path on its shard straight from / to the
// caller's host buffers (every GPU has its own PCIe link, so no GPU relays another GPU's bytes), and the call returns when
// the slowest shard is done.  No collective is needed inside a process; the torchrun / NCCL form of the same sharding for
// one-process-per-GPU deployments is zerokit_b200/sharding.py.
}  // extern "C" (templates below)
struct RlnB200Multi {
    std::vector<std::unique_ptr<FFI_RLN>> reps;
    std::vector<int> devices;
    std::vector<float> shard_ms;   // wall time of each device's shard in the last batch call
};
template <class Fn>
static void multi_for_each(RlnB200Multi& m, Fn fn) {   // fn(replica index) on one thread per device; rethrows the first failure
    std::vector<std::thread> th;
    std::vector<std::string> errs(m.reps.size());
    std::vector<char> failed(m.reps.size(), 0);
    for (size_t i = 0; i < m.reps.size(); i++)
        th.emplace_back([&, i] {
            try {
                fn(i);
            } catch (const CudaError& e) { failed[i] = 1; errs[i] = describe(e); }
            catch (const std::exception& e) { failed[i] = 1; errs[i] = describe(e); }
        });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < m.reps.size(); i++)
        if (failed[i]) throw RlnError("device " + std::to_string(m.devices[i]) + ": " + errs[i]);
}
// contiguous shard of replica i: [lo, hi)
static void multi_shard(size_t n, size_t parts, size_t i, size_t* lo, size_t* hi) {
    const size_t q = n / parts, r = n % parts;
    *lo = i * q + (i < r ? i : r);
    *hi = *lo + q + (i < r ? 1 : 0);
}
extern "C" {
RlnB200Multi_t* rlnb200_multi_new(size_t tree_depth, const int* devices, size_t n_devices, RlnString* err) {
    GUARD_BEGIN
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) throw RlnError("no usable CUDA device (librln_b200 has no CPU path)");
    auto m = std::make_unique<RlnB200Multi>();
    if (!devices || n_devices == 0) {
        for (int d = 0; d < visible; d++) m->devices.push_back(d);
    } else {
        for (size_t i = 0; i < n_devices; i++) {
            if (devices[i] < 0 || devices[i] >= visible) throw RlnError("Configuration error: device " + std::to_string(devices[i]) + " is not visible");
            m->devices.push_back(devices[i]);
        }
    }
    std::ostringstream dir;
    dir << resources_dir() << "/tree_depth_" << tree_depth;
    const std::vector<uint8_t> zkey = read_file(dir.str() + "/rln_final.arkzkey"), graph = read_file(dir.str() + "/graph.bin");
    m->reps.resize(m->devices.size());
    m->shard_ms.assign(m->devices.size(), 0.f);
    multi_for_each(*m, [&](size_t i) {
        DeviceGuard dg(m->devices[i]);
        auto h = std::make_unique<FFI_RLN>();
        h->r = std::make_unique<Rln>(tree_depth, zkey.data(), zkey.size(), graph.data(), graph.size());
        m->reps[i] = std::move(h);
    });
    return m.release();
    GUARD_END(if (err) *err = mk_string(m); return nullptr)
}
void rlnb200_multi_free(RlnB200Multi_t* m) { delete m; }
size_t rlnb200_multi_device_count(const RlnB200Multi_t* m) { return m->reps.size(); }
int rlnb200_multi_device(const RlnB200Multi_t* m, size_t i) { return i < m->devices.size() ? m->devices[i] : -1; }
// borrowed handle of replica i: every single-device call works on it (tree queries, verification, single proofs)
FFI_RLN_t* const* rlnb200_multi_replica(RlnB200Multi_t* m, size_t i) {
    static thread_local FFI_RLN* slot;
    if (i >= m->reps.size()) return nullptr;
    slot = m->reps[i].get();
    return &slot;
}
// tree updates go to every replica (the tree is part of each device's state: membership paths are read where the proofs are made)
int rlnb200_multi_set_leaves_from_bytes(RlnB200Multi_t* m, size_t index, const uint8_t* leaves_le, size_t count, RlnString* err) {
    INT_OP(multi_for_each(*m, [&](size_t i) { RlnLock lk(*m->reps[i]->r); m->reps[i]->r->set_range_host(index, leaves_le, count); });)
}
int rlnb200_multi_set_tree(RlnB200Multi_t* m, size_t tree_depth, RlnString* err) {
    INT_OP(multi_for_each(*m, [&](size_t i) { RlnLock lk(*m->reps[i]->r); m->reps[i]->r->set_tree(tree_depth); });)
}
int rlnb200_multi_atomic_operation(RlnB200Multi_t* m, size_t index, const uint8_t* leaves_le, size_t n_leaves, const size_t* indices, size_t n_indices,
                                   RlnString* err) {
    INT_OP(multi_for_each(*m, [&](size_t i) {
        RlnLock lk(*m->reps[i]->r);
        m->reps[i]->r->override_range(index, leaves_le, n_leaves, std::vector<size_t>(indices, indices + n_indices));
    });)
}
int rlnb200_multi_reserve(RlnB200Multi_t* m, size_t max_batch, RlnString* err) {
    INT_OP(multi_for_each(*m, [&](size_t i) { RlnLock lk(*m->reps[i]->r); m->reps[i]->r->reserve(max_batch); });)
}
// n witness records (host) → n proof records (host), sharded over the devices; rs may be NULL (fresh r, s)
int rlnb200_multi_prove_batch(RlnB200Multi_t* m, const uint8_t* witnesses, size_t n, const uint8_t* rs, uint8_t* proofs_out, RlnString* err) {
    INT_OP(
        const size_t parts = m->reps.size();
        multi_for_each(*m, [&](size_t i) {
            size_t lo, hi;
            multi_shard(n, parts, i, &lo, &hi);
            m->shard_ms[i] = 0.f;
            if (lo == hi) return;
            Rln& R = *m->reps[i]->r;
            RlnLock lk(R);
            const auto t0 = std::chrono::steady_clock::now();
            R.prove_records_host(witnesses + lo * R.witness_record_len(), rs ? rs + 64 * lo : nullptr, hi - lo, proofs_out + lo * R.proof_record_len());
            m->shard_ms[i] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        });)
}
int rlnb200_multi_verify_batch(RlnB200Multi_t* m, const uint8_t* proofs, size_t n, uint8_t* ok_out, RlnString* err) {
    INT_OP(
        const size_t parts = m->reps.size();
        multi_for_each(*m, [&](size_t i) {
            size_t lo, hi;
            multi_shard(n, parts, i, &lo, &hi);
            if (lo == hi) return;
            FFI_RLN* h = m->reps[i].get();
            RlnString e{nullptr, 0, 0};
            if (rlnb200_verify_batch(&h, proofs + lo * h->r->proof_record_len(), hi - lo, ok_out + lo, &e) != 0) {
                std::string msg((const char*)e.ptr, e.len);
                ffi_c_string_free(e);
                throw RlnError(msg);
            }
        });)
}
void rlnb200_multi_last_shard_ms(const RlnB200Multi_t* m, float* out) {
    for (size_t i = 0; i < m->shard_ms.size(); i++) out[i] = m->shard_ms[i];
}

RlnB200Msm_t* rlnb200_msm_new(size_t max_n, RlnString* err) {
    GUARD_BEGIN
    auto m = std::make_unique<RlnB200Msm>();
    m->device = global_init();
    m->ws = var_msm_workspace_create(max_n);
    m->max_n = max_n;
    return m.release();
    GUARD_END(if (err) *err = mk_string(m); return nullptr)
}
void rlnb200_msm_free(RlnB200Msm_t* m) {
    if (!m) return;
    try {
        DeviceGuard dg(m->device);
        var_msm_workspace_destroy(m->ws);
    } catch (...) {
    }
    delete m;
}
int rlnb200_msm_upload_bases(RlnB200Msm_t* m, const uint8_t* bases, size_t n, void* d_bases_out, RlnString* err) {
    INT_OP(DeviceGuard dg(m->device); DevMem raw; raw.upload(bases, 64 * n); launch_g1_from_bytes(raw.as<uint8_t>(), (G1Affine*)d_bases_out, n, 0); g_launch_count++;
           ZK_CUDA_CHECK(cudaDeviceSynchronize());)
}
int rlnb200_msm_gen_bases(RlnB200Msm_t* m, const void* d_scalars, size_t n, void* d_bases_out, void* stream, RlnString* err) {
    INT_OP(DeviceGuard dg(m->device); launch_g1_mul_gen((const uint8_t*)d_scalars, (G1Affine*)d_bases_out, n, (cudaStream_t)stream); g_launch_count++;)
}
int rlnb200_msm_g1_device(RlnB200Msm_t* m, const void* d_bases, const void* d_scalars, size_t n, void* d_result, void* stream, RlnString* err) {
    INT_OP(DeviceGuard dg(m->device); launch_var_msm_g1(m->ws, (const G1Affine*)d_bases, (const uint8_t*)d_scalars, n, (uint8_t*)d_result, (cudaStream_t)stream); g_launch_count += 20;)
}
int rlnb200_msm_g1(RlnB200Msm_t* m, const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t* result, RlnString* err) {
    INT_OP(DeviceGuard dg(m->device); DevMem raw, db, ds, dr; raw.upload(bases, 64 * n); db.alloc(sizeof(G1Affine) * n); ds.upload(scalars, 32 * n); dr.alloc(64);
           launch_g1_from_bytes(raw.as<uint8_t>(), db.as<G1Affine>(), n, 0);
           launch_var_msm_g1(m->ws, db.as<G1Affine>(), ds.as<uint8_t>(), n, dr.as<uint8_t>(), 0); g_launch_count += 21;
           ZK_CUDA_CHECK(cudaMemcpy(result, dr.p, 64, cudaMemcpyDeviceToHost));)
}

int rlnb200_poseidon_hash(const uint8_t* inputs, int n_inputs, uint8_t* out32, RlnString* err) {
    INT_OP(if (n_inputs < 1 || n_inputs > 3) throw RlnError("Input length must be valid with supported round parameters");
           device_poseidon(inputs, n_inputs, out32);)
}
int rlnb200_poseidon_hash_batch(const uint8_t* inputs, int n_inputs, size_t count, uint8_t* out, RlnString* err) {
    INT_OP(if (n_inputs < 1 || n_inputs > 3) throw RlnError("Input length must be valid with supported round parameters");
           global_init(); DevMem in, res; in.upload(inputs, 32 * (size_t)n_inputs * count); res.alloc(32 * count);
           launch_poseidon_batch(in.as<uint8_t>(), n_inputs, count, res.as<uint8_t>(), 0); g_launch_count++;
           ZK_CUDA_CHECK(cudaMemcpy(out, res.p, 32 * count, cudaMemcpyDeviceToHost));)
}
int rlnb200_hash_pairs(const uint8_t* pairs, size_t n, uint8_t* out, RlnString* err) {
    INT_OP(global_init(); DevMem raw, in, res, outb; raw.upload(pairs, 64 * n); in.alloc(sizeof(Fr) * 2 * n); res.alloc(sizeof(Fr) * n); outb.alloc(32 * n);
           launch_fr_from_bytes(raw.as<uint8_t>(), in.as<Fr>(), 2 * n, 0); launch_hash_pairs(in.as<Fr>(), res.as<Fr>(), n, 0);
           launch_fr_to_bytes(res.as<Fr>(), outb.as<uint8_t>(), n, 0); g_launch_count += 3;
           ZK_CUDA_CHECK(cudaMemcpy(out, outb.p, 32 * n, cudaMemcpyDeviceToHost));)
}
int rlnb200_field_op(int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out, RlnString* err);  // k_selftest.cu

}  // extern "C"
