// Internal launcher API between the host layer (rln_host.cu) and the kernel translation units.
// Everything here runs on the GPU; there is no host implementation behind any of these calls.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include <cstddef>
#include <vector>

#include "curve.cuh"
#include "poseidon.cuh"
#include "tower.cuh"
#include "vm.cuh"

namespace zk {

#define ZK_CUDA_CHECK(expr)                                                            \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) throw ::zk::CudaError(_e, #expr, __FILE__, __LINE__);   \
    } while (0)

struct CudaError {
    cudaError_t code;
    const char* expr;
    const char* file;
    int line;
    CudaError(cudaError_t c, const char* e, const char* f, int l) : code(c), expr(e), file(f), line(l) {}
};

// ---- k_poseidon.cu -------------------------------------------------------------------------
void poseidon_upload_tables(const PoseidonTables& t);
// canonical 32-byte LE integers (< r assumed; reduced otherwise) → Montgomery Fr, and back
void launch_fr_from_bytes(const uint8_t* d_in, Fr* d_out, size_t n, cudaStream_t s);
void launch_fr_to_bytes(const Fr* d_in, uint8_t* d_out, size_t n, cudaStream_t s);
// out[i] = Poseidon(in[2i], in[2i+1])   (Montgomery in/out)
void launch_hash_pairs(const Fr* d_in, Fr* d_out, size_t n, cudaStream_t s);
// Merkle tree in HBM, 1-indexed heap: root = nodes[1], level l = nodes[2^l .. 2^(l+1)), leaf i = nodes[2^depth + i].
// fill with the empty-tree values (zeros[k] per level)
void launch_merkle_fill_empty(Fr* d_nodes, u32 depth, cudaStream_t s);
// leaves (canonical bytes, on device) → nodes[2^depth+start ..), then rehash the touched ancestors level by level;
// returns the number of kernels launched
u32 launch_merkle_set_range(Fr* d_nodes, u32 depth, size_t start, const uint8_t* d_leaves_bytes, size_t count, cudaStream_t s);
// rehash ancestors of leaf range [start, start+count)
u32 launch_merkle_rehash(Fr* d_nodes, u32 depth, size_t start, size_t count, cudaStream_t s);
// membership paths: for each index, depth sibling values (canonical bytes, leaf→root) and depth index bits
void launch_merkle_paths(const Fr* d_nodes, u32 depth, const u64* d_indices, size_t n, uint8_t* d_elems_bytes, uint8_t* d_bits, cudaStream_t s);
// proof values per witness (rln/src/protocol/witness.rs:759-828): inputs layout = circuit input slots
// (canonical bytes, n × n_slots × 32); out = n × (3 + 2k) × 32 bytes [root, ext_nullifier, x, y_0..y_{k-1},
// nullifier_0..nullifier_{k-1}], k = max_out (1 for the single message-id circuit)
struct InputSlots { u32 secret, limit, message_id, path, index, x, ext_null, depth, n_slots, selector, max_out, multi; };
void launch_proof_values(const uint8_t* d_inputs, InputSlots sl, size_t n, uint8_t* d_out, cudaStream_t s);
// the same record read off an evaluated witness (vals [node][B], wire → node map): the public signals
void launch_values_from_wires(const Fr* d_vals, const u32* d_signals, u32 B, u32 max_out, uint8_t* d_out, cudaStream_t s);
// generic small helpers used by the FFI utilities (single-thread kernels)
void launch_poseidon_n(const uint8_t* d_in_bytes, int n_inputs, uint8_t* d_out_bytes, cudaStream_t s);
// count independent hashes of n_inputs (1…3) canonical values each, one thread per hash
void launch_poseidon_batch(const uint8_t* d_in_bytes, int n_inputs, size_t count, uint8_t* d_out_bytes, cudaStream_t s);

// ---- k_records.cu --------------------------------------------------------------------------
// wire records on the device: rln_witness_to_bytes_le records → input slots (+ a "would be refused" flag per record),
// (compressed proof, proof values) → rln_proof_to_bytes_le records
struct RecordLayout { InputSlots sl; u32 rec_len, proof_rec_len; };
void launch_witness_records(const uint8_t* d_records, size_t n, const RecordLayout& L, uint8_t* d_slots, u32* d_bad, cudaStream_t s);
void launch_proof_records(const uint8_t* d_proofs, const uint8_t* d_values, const uint8_t* d_slots, size_t n, const RecordLayout& L, uint8_t* d_out,
                          cudaStream_t s);

// ---- k_prover.cu ---------------------------------------------------------------------------
struct CircuitDev {
    // witness graph
    u32 n_nodes, n_slots, n_wires;
    const VmInstr* prog;
    const Fr* consts;
    const u32* signals;  // wire → node
    // list schedule of the graph (host_util.hpp vm_build_schedule): bundle b holds up to VM_SLOTS mutually independent nodes,
    // records sched[VM_SLOTS·b ..], whose operands all lie in earlier bundles; one warp per slot evaluates them (k_witness)
    const uint4* sched;   // VmRecord, two uint4 each
    u32 n_bundles;
    u32 n_consts_smem;    // constants k_witness keeps in shared memory (all of them, or 0 when the table is larger than vm_const_smem_max())
    // QAP
    u32 n_constraints, n_instance, domain, log_domain;
    const u32 *a_ptr, *a_col, *b_ptr, *b_col;
    const Fr *a_val, *b_val;
    const Fr* tw_inv;    // ω^{-k}, k < domain/2
    const Fr* tw_fwd;    // ω^{k}
    const Fr* coset;     // position p (bit-reversed order) → g^{rev(p)} / domain
    // tiled transforms (k_ntt_outer / k_ntt_middle): twiddles stored tile-major so that one bulk copy fetches a tile's
    const Fr* tw_tile_dif = nullptr;   // [64][R], R = domain / 64: tile i0, stage with local half lh at i0·R + R − 2·lh: ω^{−(i0 + 64·kj)·R/(2·lh)}
    const Fr* tw_tile_dit = nullptr;   // the same with ω
    const Fr* tw_mid = nullptr;        // [2][64]: ω^{−j·domain/(2h)} then ω^{+j·domain/(2h)} for the stage with half h ≤ 32 at 64 − 2·h
};
// inputs: n × n_slots canonical bytes → vals [n_nodes][B] (Montgomery).  err[j] != 0 if node evaluation failed.
void launch_witness(const CircuitDev& c, const uint8_t* d_inputs, Fr* d_vals, u32 B, u32* d_err, cudaStream_t s);
// k_witness streams the schedule through shared memory in blocks of this many bundles: CircuitDev::n_bundles must be a multiple
// of it (pad with empty records) and the schedule 128-byte aligned
u32 vm_schedule_block_bundles();
u32 vm_const_smem_max();
// wires: B × n_wires canonical 32-byte values (an externally calculated witness) → the same vals layout
void launch_scatter_wires(const CircuitDev& c, const uint8_t* d_wires, Fr* d_vals, u32 B, u32* d_err, cudaStream_t s);
// a,b,c [domain][B]; afterwards abuf holds h = a·b − c on the coset (natural order)
void launch_qap(const CircuitDev& c, const Fr* d_vals, Fr* d_a, Fr* d_b, Fr* d_c, u32 B, cudaStream_t s);
// plain batched NTT for tests: data [n][B] natural order in/out
void launch_ntt_test(Fr* d_data, u32 log_n, u32 B, bool inverse, const Fr* tw, cudaStream_t s);
u32 ntt_launches_per_transform(u32 log_n);
u32 qap_launch_count(const CircuitDev& c, u32 B);

// ---- k_msm_fixed.cu ------------------------------------------------------------------------
struct MsmGroupDev {
    u32 n_bases;         // non-infinity bases, ordered [wires known to a partial witness | the rest]
    u32 n_known;         // length of the known prefix (partial proofs, rln/src/partial_proof.rs:108-179)
    const u32* row;      // scalar row index per base in the source matrix
    const void* table;   // [base][window][2^(c-1)] affine points
    u32 which_src;       // 0: vals matrix, 1: h matrix
};
struct FixedMsmPlan {
    int c, K;            // G1 tables: window bits, windows
    int glv;             // G1 scalars are split k = k₁ + k₂·λ (|kᵢ| < 2^128): K covers 129 bits and every base is visited twice
    int cd, Kd;          // window bits / windows of the δ₁ table (full 254-bit scalars, no split)
    int cd2, Kd2;        // the same for the δ₂ table
    int c2, K2;          // G2 tables (few bases, so a wider window is affordable)
    MsmGroupDev g1[4];   // A, B1, L, H
    MsmGroupDev g2;      // B2
    const G1Affine* delta1_table;  // [K][2^(c-1)] multiples of δ₁ (r·δ₁, s·δ₁, rs·δ₁ in the assembly)
    const G2Affine* delta2_table;  // [K2][2^(c2-1)] multiples of δ₂
    const G1Affine *alpha1_table, *beta1_table;   // the same for α₁ and β₁ (geometry of the δ₁ table): s·α₁, r·β₁ of the folded assembly
};
// builds [base][window][digit] tables from affine bases (Montgomery, no infinities)
void launch_build_table_g1(const G1Affine* d_bases, u32 n, int c, int K, G1Affine* d_table, cudaStream_t s);
void launch_build_table_g2(const G2Affine* d_bases, u32 n, int c, int K, G2Affine* d_table, cudaStream_t s);
// self-test of the GLV split: n canonical scalars → n × 36 bytes (|k₁| 16 B, |k₂| 16 B, sign₁, sign₂, 2 pad)
void launch_glv_split(const uint8_t* d_scalars, size_t n, uint8_t* d_out, cudaStream_t s);
// self-test of the Straus / GLV double multiplication of the proof assembly: n × (P 64 | kp 32 | Q 64 | kq 32) → n × 64 B
void launch_glv_double_mul(const uint8_t* d_in, size_t n, int use_q, uint8_t* d_out, cudaStream_t s);
struct ProverKeyDev {  // fixed points of the proving key (affine, Montgomery)
    G1Affine alpha_g1, beta_g1, delta_g1;
    G2Affine beta_g2, delta_g2;
};
struct MsmTask { u32 group, lo, hi, half; };  // bases [lo, hi) of one group, summed by one thread per proof; half = 1: the k₂ part of a GLV split
struct MsmWorkspace {
    G1XYZZ* part_g1;  // [tasks][B]
    G2XYZZ* part_g2;
    G1XYZZ* sum_g1;   // [4][B]
    G2XYZZ* sum_g2;   // [B]
    const MsmTask *tasks_g1, *tasks_g2;  // device arrays built from msm_make_tasks for this B
    u32 n_tasks_g1, n_tasks_g2;
    cudaEvent_t* ev;  // optional: 6 events recorded around [g1 accum, g1 reduce, g2 accum, g2 reduce, assemble]
    // optional (a handful of full proofs): s·ΣzᵢAᵢ and r·ΣzᵢB₁ᵢ as two more table sums over the scaled witness values fold_s = s·z,
    // fold_r = r·z ([node][B] like vals), so that the assembly needs no variable-base multiplication: partial sums and sums [2][B]
    const Fr *fold_s = nullptr, *fold_r = nullptr;
    G1XYZZ *fold_part = nullptr, *fold_sum = nullptr;
    u32 n_tasks_ab = 0;                                   // tasks of the groups A and B₁ (a prefix of tasks_g1)
    cudaStream_t side = nullptr;                          // optional second stream + two events: the G2 assembly runs beside the G1 assembly
    cudaEvent_t side_fork = nullptr, side_join = nullptr;
};
// MSM phases: all bases (full proof), the known prefix of A/B₁/B₂/L (partial proof), or the unknown suffix plus H (finish)
enum MsmPhase { MSM_FULL = 0, MSM_KNOWN = 1, MSM_UNKNOWN = 2 };
std::vector<MsmTask> msm_make_tasks(const FixedMsmPlan& plan, u32 B, bool g2, int phase);
// out_s[row][j] = s_j·vals[row][j], out_r[row][j] = r_j·vals[row][j]   (rs: B × (r | s), 32 canonical bytes each)
void launch_scale_vals(const Fr* d_vals, const uint8_t* d_rs, u32 n_rows, u32 B, Fr* d_out_s, Fr* d_out_r, cudaStream_t s);
// accumulate + reduce over ws.tasks_* → ws.sum_g1[4][B], ws.sum_g2[B]
void launch_msm_sums(const FixedMsmPlan& plan, const Fr* d_vals, const Fr* d_h, u32 B, MsmWorkspace& ws, cudaStream_t s);
// assembly (partial_proof.rs:226-273) + affine + ark-compressed bytes.  rs: B × 64 canonical bytes (r | s).
// d_partial (optional): B × 320 bytes canonical affine partial_pi_a | partial_rho | partial_pi_b | partial_pi_c that replace
// α₁ / β₁ / β₂ / 0 (finish_partial_proof_with_assignment, partial_proof.rs:182-274).
// proofs_out: B × 128 bytes.  proofs_affine (optional): B × 256 bytes canonical A|B|C
void launch_assemble(const FixedMsmPlan& plan, const ProverKeyDev& pk, u32 B, const uint8_t* d_rs, MsmWorkspace& ws,
                     const uint8_t* d_partial, uint8_t* d_proofs_out, uint8_t* d_proofs_affine, cudaStream_t s);
// partial proof points from the sums of the known prefix: out_affine B × 320 canonical, out_compressed B × 160 ark-compressed
void launch_partial_out(const ProverKeyDev& pk, u32 B, MsmWorkspace& ws, uint8_t* d_out_affine, uint8_t* d_out_compressed, cudaStream_t s);

// ---- k_msm_var.cu --------------------------------------------------------------------------
// variable-base G1 Pippenger MSM (rln/src/partial_proof.rs:98-104 `msm`): bases affine Montgomery (device),
// scalars canonical 32-byte LE (device).  result: 64 bytes canonical affine (x|y), all-zero = infinity.
struct VarMsmWorkspace;
VarMsmWorkspace* var_msm_workspace_create(size_t max_n);
void var_msm_workspace_destroy(VarMsmWorkspace* w);
void launch_var_msm_g1(VarMsmWorkspace* w, const G1Affine* d_bases, const uint8_t* d_scalars, size_t n, uint8_t* d_result,
                       cudaStream_t s);
// canonical x|y bytes → Montgomery affine (flags in the top bits of y's last byte: 0x40 = infinity)
void launch_g1_from_bytes(const uint8_t* d_in, G1Affine* d_out, size_t n, cudaStream_t s);
void launch_g2_from_bytes(const uint8_t* d_in, G2Affine* d_out, size_t n, cudaStream_t s);
// bases[i] = k_i · G for the G1 generator (bench input generation on the device)
void launch_g1_mul_gen(const uint8_t* d_scalars, G1Affine* d_out, size_t n, cudaStream_t s);

// ---- k_verify.cu ---------------------------------------------------------------------------
struct VerifyKeyDev {
    G1Affine alpha_g1;
    G2Affine beta_g2, gamma_g2, delta_g2;
    const G1Affine* gamma_abc;  // device pointer, n_public + 1 entries
    u32 n_public;
    const G1Affine* gamma_tab;  // window tables of gamma_abc[1..n_public]: [n_public][gK][2^(gc−1)] (vk_x without doublings)
    int gc, gK;
    // prepare_verifying_key (ark-groth16): everything that depends only on the key is computed once
    const Fq12* ml_alpha_beta;            // device pointer: Miller value of (α₁, β₂)
    const Fq2 *gamma_lam, *gamma_c;       // device pointers: line coefficients of γ₂ (MILLER_STEPS each)
    const Fq2 *delta_lam, *delta_c;       // … and of δ₂
};
void pairing_upload_tables(const PairingTables& t);
// proofs: n × 128 B ark-compressed; publics: n × n_public × 32 canonical bytes (circuit order);
// ok[j] = 1 valid, 0 invalid, 2 malformed encoding (not on curve / bad flags)
void launch_verify(const VerifyKeyDev& vk, const uint8_t* d_proofs, const uint8_t* d_publics, size_t n, uint8_t* d_ok, cudaStream_t s);
// ---- k_verify_vm.cu ------------------------------------------------------------------------
// the same verification, one CTA per proof, as a host-scheduled program of lane-parallel sums of products (verify_vm.hpp);
// ok codes as launch_verify plus 3 = "not decided here" (a point at infinity, an exceptional addition): re-run with launch_verify
struct VerifyVmDev {
    const u32* code;      // n_levels × REC_WORDS × LANES record words
    const Fq* consts;     // image of the pinned slots [0, n_const)
    u32 n_levels, n_const, n_slots;
    u32 exps[3][8];       // the public exponents of the SP_EXP levels
    long long* trace;     // optional: clock64() of thread 0 at the start of every level (+ one at the end), for the cost model
};
void launch_verify_vm(const VerifyVmDev& prog, const VerifyKeyDev& vk, const uint8_t* d_proofs, const uint8_t* d_publics, size_t n, uint8_t* d_ok,
                      cudaStream_t s);
// decompress n proofs to affine canonical bytes (256 B each: A 64 | B 128 | C 64); ok[j]=0 if malformed
void launch_decompress(const uint8_t* d_proofs, size_t n, uint8_t* d_affine, uint8_t* d_ok, cudaStream_t s);

}  // namespace zk
