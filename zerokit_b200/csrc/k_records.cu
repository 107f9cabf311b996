// Wire records on the device: the scatter / gather edge of the batch path.
//
//   k_witness_records: n × rln_witness_to_bytes_le records (rln/src/protocol/witness.rs:369-415; single and multi message-id
//                      layouts) → the circuit's input slots (iden3calc.rs:106-181, witness.rs:832-881) + one flag per record
//                      that says whether bytes_le_to_rln_witness / RLNWitnessInput::new_* (witness.rs:78-176, 470-560) would
//                      have refused it.  The host re-parses a flagged record with its own parser to produce the reference's
//                      error text, so the device only needs the verdict, not the message.
//   k_proof_records:   n × (128-byte ark-compressed proof, proof values) → n × rln_proof_to_bytes_le records
//                      (proof.rs:192-236, 413-428).
//
// With these two kernels a batch that arrives as bytes (from the host, or from another GPU over NCCL) never needs per-record
// host work: H2D / ncclRecv → k_witness_records → prover → k_proof_records → D2H / ncclSend.
#include "device_api.hpp"

namespace zk {

static __device__ __forceinline__ bool canonical_fr(const uint8_t* b) {   // value < r, bytes little-endian, unaligned
#pragma unroll
    for (int w = 7; w >= 0; w--) {
        const u32 v = (u32)b[4 * w] | ((u32)b[4 * w + 1] << 8) | ((u32)b[4 * w + 2] << 16) | ((u32)b[4 * w + 3] << 24);
        const u32 p = FrCfg::p(w);
        if (v < p) return true;
        if (v > p) return false;
    }
    return false;
}
static __device__ __forceinline__ int cmp32(const uint8_t* a, const uint8_t* b) {
    for (int i = 31; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
static __device__ __forceinline__ u64 load_u64(const uint8_t* b) {
    u64 v = 0;
#pragma unroll
    for (int i = 7; i >= 0; i--) v = (v << 8) | b[i];
    return v;
}
static __device__ __forceinline__ void copy32(uint8_t* dst, const uint8_t* src) {   // dst is 32-byte aligned, src is not
    u32 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = (u32)src[4 * i] | ((u32)src[4 * i + 1] << 8) | ((u32)src[4 * i + 2] << 16) | ((u32)src[4 * i + 3] << 24);
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
static __device__ __forceinline__ void small32(uint8_t* dst, u32 v) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4(v, 0, 0, 0);
    d[1] = make_uint4(0, 0, 0, 0);
}

// one warp per record: lane l copies the field elements l, l+32, … of the record; lane 0 also checks the framing and the
// witness rules.  rec_len is fixed by (depth, max_out, mode), so a record whose length prefixes differ from the circuit's
// cannot be parsed at its slot and is flagged.
__global__ void __launch_bounds__(128) k_witness_records(const uint8_t* __restrict__ recs, size_t n, RecordLayout L, uint8_t* __restrict__ slots,
                                                        u32* __restrict__ bad) {
    const size_t j = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const u32 lane = threadIdx.x & 31;
    if (j >= n) return;
    const uint8_t* r = recs + j * (size_t)L.rec_len;
    uint8_t* out = slots + j * (size_t)L.sl.n_slots * 32;
    const u32 d = L.sl.depth, k = L.sl.max_out;
    // byte offsets inside the record
    const u32 o_secret = 1, o_limit = 33;
    const u32 o_mid_single = 65;
    const u32 o_plen = L.sl.multi ? 65 : 97, o_path = o_plen + 8, o_ilen = o_path + 32 * d, o_idx = o_ilen + 8, o_x = o_idx + d, o_en = o_x + 32;
    const u32 o_klen = o_en + 32, o_mids = o_klen + 8, o_slen = o_mids + 32 * k, o_sel = o_slen + 8;
    // every slot the record does not define is zero; slot 0 is the constant 1 (iden3calc.rs:177-181)
    for (u32 s = lane; s < L.sl.n_slots; s += 32) small32(out + 32 * s, s == 0 ? 1u : 0u);
    __syncwarp();
    u32 ok = 1;
    // field elements, spread over the lanes: secret, limit, x, external_nullifier, message ids, path elements
    const u32 n_fe = 4 + k + d;
    for (u32 e = lane; e < n_fe; e += 32) {
        const uint8_t* src;
        u32 slot;
        if (e == 0) { src = r + o_secret; slot = L.sl.secret; }
        else if (e == 1) { src = r + o_limit; slot = L.sl.limit; }
        else if (e == 2) { src = r + o_x; slot = L.sl.x; }
        else if (e == 3) { src = r + o_en; slot = L.sl.ext_null; }
        else if (e < 4 + k) { src = r + (L.sl.multi ? o_mids + 32 * (e - 4) : o_mid_single); slot = L.sl.message_id + (e - 4); }
        else { src = r + o_path + 32 * (e - 4 - k); slot = L.sl.path + (e - 4 - k); }
        if (!canonical_fr(src)) ok = 0;
        copy32(out + 32 * slot, src);
    }
    for (u32 i = lane; i < d; i += 32) small32(out + 32 * (L.sl.index + i), r[o_idx + i]);   // Fr::from(u8) (witness.rs:835-839)
    if (L.sl.multi)
        for (u32 i = lane; i < k; i += 32) small32(out + 32 * (L.sl.selector + i), r[o_sel + i] ? 1u : 0u);   // any non-zero byte is true (utils.rs:407-410)
    if (lane == 0) {
        if (r[0] != (L.sl.multi ? 1 : 0)) ok = 0;
        if (load_u64(r + o_plen) != d || load_u64(r + o_ilen) != d) ok = 0;
        bool zero_limit = true;
        for (int i = 0; i < 32; i++) zero_limit = zero_limit && r[o_limit + i] == 0;
        if (zero_limit) ok = 0;                                                               // "User message limit cannot be zero"
        if (!L.sl.multi) {
            if (cmp32(r + o_mid_single, r + o_limit) >= 0) ok = 0;                            // message id within the limit
        } else {
            if (load_u64(r + o_klen) != k || load_u64(r + o_slen) != k) ok = 0;
            bool any = false;
            for (u32 i = 0; i < k; i++) {
                const bool si = r[o_sel + i] != 0;
                any = any || si;
                if (!si) continue;
                if (cmp32(r + o_mids + 32 * i, r + o_limit) >= 0) ok = 0;
                for (u32 q = 0; q < i; q++)
                    if (r[o_sel + q] != 0 && cmp32(r + o_mids + 32 * i, r + o_mids + 32 * q) == 0) ok = 0;   // duplicate message id
            }
            if (!any) ok = 0;
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) bad[j] = ok ? 0u : 1u;
}

// values layout (launch_proof_values): root | external_nullifier | x | ys[k] | nullifiers[k]
__global__ void __launch_bounds__(256) k_proof_records(const uint8_t* __restrict__ proofs, const uint8_t* __restrict__ values, const uint8_t* __restrict__ slots,
                                                      size_t n, RecordLayout L, uint8_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const u32 lane = threadIdx.x & 31;
    if (j >= n) return;
    const u32 k = L.sl.max_out, vs = 32 * (3 + 2 * k);
    uint8_t* o = out + j * (size_t)L.proof_rec_len;
    const uint8_t* p = proofs + 128 * j;
    const uint8_t* v = values + (size_t)vs * j;
    const uint8_t ver = L.sl.multi ? 1 : 0;
    for (u32 i = lane; i < 128; i += 32) o[1 + i] = p[i];
    if (!L.sl.multi) {   // version | proof | version | root | external_nullifier | x | y | nullifier
        for (u32 i = lane; i < vs; i += 32) o[130 + i] = v[i];
        if (lane == 0) { o[0] = ver; o[129] = ver; }
        return;
    }
    // version | proof | version | root | external_nullifier | x | u64 k | ys | u64 k | nullifiers | u64 k | selector bytes
    for (u32 i = lane; i < 96; i += 32) o[130 + i] = v[i];
    const u32 o_ys = 130 + 96 + 8, o_nl = o_ys + 32 * k + 8, o_sl = o_nl + 32 * k + 8;
    for (u32 i = lane; i < 32 * k; i += 32) { o[o_ys + i] = v[96 + i]; o[o_nl + i] = v[96 + 32 * k + i]; }
    if (lane < 8) {
        const uint8_t b = lane == 0 ? (uint8_t)k : 0;   // k ≤ 16
        o[o_ys - 8 + lane] = b; o[o_nl - 8 + lane] = b; o[o_sl - 8 + lane] = b;
    }
    const uint8_t* sl = slots + j * (size_t)L.sl.n_slots * 32;
    for (u32 i = lane; i < k; i += 32) o[o_sl + i] = sl[32 * (L.sl.selector + i)] ? 1 : 0;
    if (lane == 0) { o[0] = ver; o[129] = ver; }
}

void launch_witness_records(const uint8_t* d_records, size_t n, const RecordLayout& L, uint8_t* d_slots, u32* d_bad, cudaStream_t s) {
    if (!n) return;
    k_witness_records<<<(unsigned)((n + 3) / 4), 128, 0, s>>>(d_records, n, L, d_slots, d_bad);
    ZK_CUDA_CHECK(cudaGetLastError());
}
void launch_proof_records(const uint8_t* d_proofs, const uint8_t* d_values, const uint8_t* d_slots, size_t n, const RecordLayout& L, uint8_t* d_out,
                          cudaStream_t s) {
    if (!n) return;
    k_proof_records<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(d_proofs, d_values, d_slots, n, L, d_out);
    ZK_CUDA_CHECK(cudaGetLastError());
}

}  // namespace zk
