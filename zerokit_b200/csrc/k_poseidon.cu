// Poseidon / Merkle-tree kernels (sm_100a).
//
// Path covered: utils/src/merkle_tree/full_merkle_tree.rs:197-223
// (set_range), :360-399 (update_hashes, level by level), :288-304 (proof) and
// rln/src/protocol/witness.rs:759-828 (proof_values_from_witness, compute_tree_root).
//
// Layout: the tree is one Fr array in HBM, 1-indexed heap order, so the two children of node p
// are the 64-byte aligned pair nodes[2p], nodes[2p+1]; one thread hashes one parent (4 × 128-bit
// loads, 2 × 128-bit stores, fully coalesced across the warp).  Round constants and the MDS live
// in __constant__ memory: every lane of a warp reads the same constant in the same round.
// Work per node is ≈ 830 Montgomery products against 96 bytes of traffic, so these kernels are
// bound by the INT32 multiply pipe, not by HBM.
#include "device_api.hpp"

namespace zk {

__constant__ PoseidonTables c_pt;
// a second copy in global memory for the cooperative hash: its three lanes read three different constants per round, which the
// constant cache would serialise; through L1 they are one 96-byte request
__device__ PoseidonTables g_pt;

void poseidon_upload_tables(const PoseidonTables& t) {
    ZK_CUDA_CHECK(cudaMemcpyToSymbol(c_pt, &t, sizeof(PoseidonTables)));
    ZK_CUDA_CHECK(cudaMemcpyToSymbol(g_pt, &t, sizeof(PoseidonTables)));
}

__device__ __forceinline__ Fr d_hash1(const Fr& a) {
    Fr st[2] = {Fr::zero(), a};
    return poseidon_permute<2>(st, c_pt.ark2, c_pt.mds2);
}
__device__ __forceinline__ Fr d_hash2(const Fr& a, const Fr& b) {
    Fr st[3] = {Fr::zero(), a, b};
    return poseidon_permute<3>(st, c_pt.ark3, c_pt.mds3);
}
__device__ __forceinline__ Fr d_hash3(const Fr& a, const Fr& b, const Fr& c) {
    Fr st[4] = {Fr::zero(), a, b, c};
    return poseidon_permute<4>(st, c_pt.ark4, c_pt.mds4);
}

__device__ __forceinline__ Fr load_canonical(const uint8_t* p) {  // 32-byte aligned canonical LE → Montgomery
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    u32 c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    u32 m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = FrCfg::p(i);
    while (Fr::raw_cmp(c, m) >= 0) Fr::raw_sub(c, c, m);  // at most 5 iterations for any 256-bit input
    return Fr::from_canonical(c);
}
__device__ __forceinline__ void store_canonical(uint8_t* p, const Fr& v) {
    u32 c[8];
    v.to_canonical(c);
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(c[0], c[1], c[2], c[3]);
    q[1] = make_uint4(c[4], c[5], c[6], c[7]);
}

__global__ void k_fr_from_bytes(const uint8_t* __restrict__ in, Fr* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) st_fp(out + i, load_canonical(in + 32 * i));
}
__global__ void k_fr_to_bytes(const Fr* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) store_canonical(out + 32 * i, ld_fp(in + i));
}
void launch_fr_from_bytes(const uint8_t* d_in, Fr* d_out, size_t n, cudaStream_t s) {
    if (!n) return;
    k_fr_from_bytes<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_in, d_out, n);
}
void launch_fr_to_bytes(const Fr* d_in, uint8_t* d_out, size_t n, cudaStream_t s) {
    if (!n) return;
    k_fr_to_bytes<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_in, d_out, n);
}

// out[i] = H(in[2i], in[2i+1]) — the Merkle level kernel.  `in` and `out` may be levels of the same array.
__global__ void __launch_bounds__(128) k_hash_pairs(const Fr* __restrict__ in, Fr* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr l = ld_fp(in + 2 * i), r = ld_fp(in + 2 * i + 1);
    st_fp(out + i, d_hash2(l, r));
}
void launch_hash_pairs(const Fr* d_in, Fr* d_out, size_t n, cudaStream_t s) {
    if (!n) return;
    k_hash_pairs<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, d_out, n);
}

// zeros[k] = root of an empty subtree of height k (full_merkle_tree.rs:82-115); one thread computes the chain
__global__ void k_zero_chain(Fr* zeros, u32 depth) {
    Fr z = Fr::zero();
    zeros[0] = z;
    for (u32 k = 0; k < depth; k++) {
        z = d_hash2(z, z);
        zeros[k + 1] = z;
    }
}
__global__ void k_fill_levels(Fr* __restrict__ nodes, const Fr* __restrict__ zeros, u32 depth) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // node index 1 .. 2^(depth+1)-1
    size_t total = (size_t)2 << depth;
    if (i == 0 || i >= total) return;
    u32 level = 63 - __clzll((unsigned long long)i);
    st_fp(nodes + i, ld_fp(zeros + (depth - level)));
}
void launch_merkle_fill_empty(Fr* d_nodes, u32 depth, cudaStream_t s) {
    Fr* scratch = nullptr;
    ZK_CUDA_CHECK(cudaMallocAsync((void**)&scratch, sizeof(Fr) * (depth + 1), s));
    k_zero_chain<<<1, 1, 0, s>>>(scratch, depth);
    size_t total = (size_t)2 << depth;
    k_fill_levels<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_nodes, scratch, depth);
    ZK_CUDA_CHECK(cudaFreeAsync(scratch, s));
}

__global__ void k_set_leaves(Fr* __restrict__ nodes, u32 depth, size_t start, const uint8_t* __restrict__ leaves, size_t count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < count) st_fp(nodes + ((size_t)1 << depth) + start + i, load_canonical(leaves + 32 * i));
}
// parents [first, first+count) of one level
__global__ void __launch_bounds__(128) k_merkle_level(Fr* __restrict__ nodes, size_t first, size_t count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    size_t p = first + i;
    Fr l = ld_fp(nodes + 2 * p), r = ld_fp(nodes + 2 * p + 1);
    st_fp(nodes + p, d_hash2(l, r));
}
// Latency-bound levels (the top of the tree, and every level of a single-leaf update): one hash is ≈ 65 rounds of a serial
// dependency chain, 0.26 ms for one thread.  Three lanes share a hash instead — lane k owns state element k, applies its own
// S-box, reads the other two elements with warp shuffles and computes row k of the MDS product — which cuts the chain per
// round from (3 S-boxes + 3 rows) to (1 S-box + 1 row).  A warp carries 10 hashes on lanes 0..29.
__device__ __forceinline__ Fr shfl_fr(const Fr& v, int src) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}
// mine: element k of the initial state (0, left, right); returns the hash on every lane of the group
__device__ __forceinline__ Fr poseidon3_coop(Fr mine, int k, int base) {
    constexpr int RF = PoseidonShape<3>::RF, RP = PoseidonShape<3>::RP;
    const Fr m0 = ldg_fp(&g_pt.mds3[k * 3]), m1 = ldg_fp(&g_pt.mds3[k * 3 + 1]), m2 = ldg_fp(&g_pt.mds3[k * 3 + 2]);
    const int s1 = base + 1 < 32 ? base + 1 : 31, s2 = base + 2 < 32 ? base + 2 : 31;
#pragma unroll 1
    for (int r = 0; r < RF + RP; r++) {
        mine += ldg_fp(&g_pt.ark3[r * 3 + k]);
        const bool full = (r < RF / 2) || (r >= RF / 2 + RP);
        if (full || k == 0) mine = sbox5(mine);
        const Fr st[3] = {shfl_fr(mine, base), shfl_fr(mine, s1), shfl_fr(mine, s2)};
        const Fr row[3] = {m0, m1, m2};
        mine = Fr::dot<3>(st, row);
    }
    return shfl_fr(mine, base);
}
// parents [first, first+count) of one level, 3 lanes per hash
__global__ void __launch_bounds__(128) k_merkle_level_coop(Fr* __restrict__ nodes, size_t first, size_t count) {
    const int lane = threadIdx.x & 31, g = lane / 3, k = lane - 3 * g;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t i = warp * 10 + g;
    const bool live = g < 10 && i < count;
    const size_t p = first + (live ? i : 0);
    Fr mine = Fr::zero();
    if (live && k) mine = ld_fp(nodes + 2 * p + (k - 1));
    const Fr h = poseidon3_coop(mine, k, 3 * g);
    if (live && k == 0) st_fp(nodes + p, h);
}
// the last levels of a range, while a level has at most 10 parents: one warp walks them all in a single launch
__global__ void __launch_bounds__(32) k_merkle_top_coop(Fr* nodes, size_t lo, size_t hi, u32 levels) {
    const int lane = threadIdx.x, g = lane / 3, k = lane - 3 * g;
    for (u32 l = 0; l < levels; l++) {
        lo >>= 1;
        hi >>= 1;
        const bool live = g < 10 && lo + g <= hi;
        const size_t p = live ? lo + g : lo;
        Fr mine = Fr::zero();
        if (live && k) mine = ld_fp(nodes + 2 * p + (k - 1));
        const Fr h = poseidon3_coop(mine, k, 3 * g);
        if (live && k == 0) st_fp(nodes + p, h);
        __threadfence_block();
        __syncwarp();
    }
}
u32 launch_merkle_rehash(Fr* d_nodes, u32 depth, size_t start, size_t count, cudaStream_t s) {
    if (!count) return 0;
    size_t lo = ((size_t)1 << depth) + start, hi = lo + count - 1;
    for (u32 l = 0; l < depth; l++) {
        if (((hi >> 1) - (lo >> 1) + 1) <= 10) {   // and so is every level above
            k_merkle_top_coop<<<1, 32, 0, s>>>(d_nodes, lo, hi, depth - l);
            return l + 1;
        }
        lo >>= 1;
        hi >>= 1;
        size_t n = hi - lo + 1;
        // below ≈ 8 K parents the one-thread-per-hash kernel cannot fill the chip and its 0.26 ms chain is the cost of the level
        if (n <= 8192) k_merkle_level_coop<<<(unsigned)((n + 39) / 40), 128, 0, s>>>(d_nodes, lo, n);
        else k_merkle_level<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_nodes, lo, n);
    }
    return depth;
}
u32 launch_merkle_set_range(Fr* d_nodes, u32 depth, size_t start, const uint8_t* d_leaves_bytes, size_t count, cudaStream_t s) {
    if (!count) return 0;
    k_set_leaves<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d_nodes, depth, start, d_leaves_bytes, count);
    return 1 + launch_merkle_rehash(d_nodes, depth, start, count, s);
}

// one thread per (path, level): sibling gather — 32 B read, 32 B write, pure HBM traffic
__global__ void k_merkle_paths(const Fr* __restrict__ nodes, u32 depth, const u64* __restrict__ idx, size_t n,
                               uint8_t* __restrict__ elems, uint8_t* __restrict__ bits) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n * depth) return;
    size_t path = t / depth;
    u32 lvl = (u32)(t % depth);
    size_t node = (((size_t)1 << depth) + idx[path]) >> lvl;
    store_canonical(elems + 32 * t, ld_fp(nodes + (node ^ 1)));
    bits[t] = (uint8_t)(node & 1);
}
void launch_merkle_paths(const Fr* d_nodes, u32 depth, const u64* d_indices, size_t n, uint8_t* d_elems_bytes, uint8_t* d_bits,
                         cudaStream_t s) {
    if (!n) return;
    size_t t = n * depth;
    k_merkle_paths<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(d_nodes, depth, d_indices, n, d_elems_bytes, d_bits);
}

// proof_values_from_witness (witness.rs:759-804) + compute_tree_root (:807-828); one witness per thread
__global__ void __launch_bounds__(64) k_proof_values(const uint8_t* __restrict__ inputs, InputSlots sl, size_t n, uint8_t* __restrict__ out) {
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint8_t* in = inputs + j * (size_t)sl.n_slots * 32;
    Fr secret = load_canonical(in + 32 * sl.secret);
    Fr limit = load_canonical(in + 32 * sl.limit);
    Fr x = load_canonical(in + 32 * sl.x);
    Fr en = load_canonical(in + 32 * sl.ext_null);
    Fr root = d_hash2(d_hash1(secret), limit);
    for (u32 i = 0; i < sl.depth; i++) {
        Fr e = load_canonical(in + 32 * (sl.path + i));
        Fr b = load_canonical(in + 32 * (sl.index + i));
        root = b.is_zero() ? d_hash2(root, e) : d_hash2(e, root);
    }
    const u32 k = sl.max_out;
    uint8_t* o = out + j * (size_t)(32 * (3 + 2 * k));
    store_canonical(o, root);
    store_canonical(o + 32, en);
    store_canonical(o + 64, x);
    for (u32 i = 0; i < k; i++) {  // witness.rs:773-798: y = (a0 + x·a1)·selector, nullifier = H(a1)·selector
        Fr mid = load_canonical(in + 32 * (sl.message_id + i));
        Fr a1 = d_hash3(secret, en, mid);
        Fr y = secret + x * a1;
        Fr nullifier = d_hash1(a1);
        if (sl.multi && load_canonical(in + 32 * (sl.selector + i)).is_zero()) { y = Fr::zero(); nullifier = Fr::zero(); }
        store_canonical(o + 96 + 32 * i, y);
        store_canonical(o + 96 + 32 * (k + i), nullifier);
    }
}
void launch_proof_values(const uint8_t* d_inputs, InputSlots sl, size_t n, uint8_t* d_out, cudaStream_t s) {
    if (!n) return;
    k_proof_values<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(d_inputs, sl, n, d_out);
}

// The same values read off the witness: y…, root, nullifier…, x, external nullifier are the circuit's public signals, wires 1 … 2k+3
// (rln/src/protocol/proof.rs:863-884 lists them in that order).  For a handful of proofs this replaces k_proof_values, whose one
// thread per proof hashes the whole Merkle path again (6 ms: longer than the rest of a single proof once that is below 7 ms).
__global__ void __launch_bounds__(64) k_values_from_wires(const Fr* __restrict__ vals, const u32* __restrict__ signals, u32 B, u32 k,
                                                         uint8_t* __restrict__ out) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    auto wire = [&](u32 w) { return ld_fp(vals + (size_t)__ldg(signals + w) * B + j); };
    uint8_t* o = out + j * (size_t)(32 * (3 + 2 * k));
    store_canonical(o, wire(k + 1));
    store_canonical(o + 32, wire(2 * k + 3));
    store_canonical(o + 64, wire(2 * k + 2));
    for (u32 i = 0; i < k; i++) {
        store_canonical(o + 96 + 32 * i, wire(1 + i));
        store_canonical(o + 96 + 32 * (k + i), wire(k + 2 + i));
    }
}
void launch_values_from_wires(const Fr* d_vals, const u32* d_signals, u32 B, u32 max_out, uint8_t* d_out, cudaStream_t s) {
    if (!B) return;
    k_values_from_wires<<<(B + 63) / 64, 64, 0, s>>>(d_vals, d_signals, B, max_out, d_out);
}

__global__ void k_poseidon_n(const uint8_t* in, int n, uint8_t* out) {
    Fr r;
    if (n == 1) r = d_hash1(load_canonical(in));
    else if (n == 2) r = d_hash2(load_canonical(in), load_canonical(in + 32));
    else r = d_hash3(load_canonical(in), load_canonical(in + 32), load_canonical(in + 64));
    store_canonical(out, r);
}
void launch_poseidon_n(const uint8_t* d_in_bytes, int n_inputs, uint8_t* d_out_bytes, cudaStream_t s) {
    k_poseidon_n<<<1, 1, 0, s>>>(d_in_bytes, n_inputs, d_out_bytes);
}
// count independent hashes of n inputs each (canonical bytes in and out), one thread per hash
__global__ void __launch_bounds__(128) k_poseidon_batch(const uint8_t* in, int n, size_t count, uint8_t* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint8_t* p = in + 32 * (size_t)n * i;
    Fr r;
    if (n == 1) r = d_hash1(load_canonical(p));
    else if (n == 2) r = d_hash2(load_canonical(p), load_canonical(p + 32));
    else r = d_hash3(load_canonical(p), load_canonical(p + 32), load_canonical(p + 64));
    store_canonical(out + 32 * i, r);
}
void launch_poseidon_batch(const uint8_t* d_in_bytes, int n_inputs, size_t count, uint8_t* d_out_bytes, cudaStream_t s) {
    if (!count) return;
    k_poseidon_batch<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(d_in_bytes, n_inputs, count, d_out_bytes);
    ZK_CUDA_CHECK(cudaGetLastError());
}

}  // namespace zk
