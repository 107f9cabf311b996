// Variable-base BN254 G1 multi-scalar multiplication (Pippenger bucket method) for sm_100a.
//
// Path covered: `msm<G>` in rln/src/partial_proof.rs:98-104 → ark-ec 0.5.0
// VariableBaseMSM::msm_bigint (Cargo.lock:106).  Same mathematical object (Σ sᵢ·Pᵢ), so the affine
// result is bit-identical; the schedule is GPU-shaped:
//   1. signed radix-2^c digits per scalar                         (k_digits, 32 B read / scalar)
//   2. counting sort of the point indices by (window, bucket): histogram with one atomic per digit, exclusive scan, scatter
//      through per-bucket cursors (k_digits<COUNT/SCATTER>, k_scan_*: hand-written — the order inside a bucket is irrelevant to
//      a sum, so no stable radix sort is needed and neither keys nor an unsorted copy of the indices ever touch HBM)
//   3. one thread per bucket: sum its points with XYZZ mixed adds  (k_bucket_sum: the hot kernel —
//      128-bit loads of the 64-byte affine bases, ≈ 10 modular products per 64 bytes)
//   4. per window: Σ (b+1)·B_b by segment running sums, warp-shuffle tree over the segments
//   5. Horner over the windows, affine normalisation.
// Algorithmic HBM bytes: 96 per term (32 B scalar + 64 B base).  The kernel is bound by the integer
// multiply pipe (≈ 250 IMAD per byte), see DESIGN.md §roofline.
#include <cstdlib>

#include "device_api.hpp"
#include "glv.cuh"

namespace zk {

struct VarMsmWorkspace {
    size_t max_n = 0;
    u32* vals_out = nullptr;                          // point index | sign << 31 of every non-zero digit, grouped by (window, bucket)
    u32 *bucket_cnt = nullptr, *bucket_start = nullptr, *bucket_end = nullptr;
    u32 *size_bins = nullptr, *order = nullptr;       // buckets ordered by size, largest first (counting sort over clamped sizes)
    u32* scan_tmp = nullptr;                          // block sums of the exclusive scans
    u32 *slice_cnt = nullptr, *slice_off = nullptr;   // per ordered bucket: number of ≤ SLICE-point slices and their first slice id
    G1XYZZ* partial = nullptr;                        // one partial sum per slice
    size_t max_slices = 0;
    G1XYZZ* buckets = nullptr;
    G1XYZZ* seg = nullptr;
    G1XYZZ* win = nullptr;
    size_t max_buckets = 0;
    G1Affine* phi = nullptr;                          // GLV mode only: φ(Pᵢ) = (β·xᵢ, yᵢ), allocated on first use
    size_t phi_cap = 0;
};

// GLV in the variable-base MSM: every scalar is split k = k₁ + k₂·λ with |kᵢ| < 2^128 and the MSM runs over the 2n terms [Pᵢ | φ(Pᵢ)].
// At small n the addition count stays (2¹⁶: 2¹⁷ × 12 windows of c = 11 against 2¹⁶ × 26 of c = 10) while the Horner doublings — one
// thread's dependent chain, the floor of every small MSM — and the bucket reductions halve; at 2²² it would cost 9 windows instead of
// 8 per 128 bits.  Measured on a B200 (profiles/r02a_flags.txt): 2¹⁶ 3.03 → 2.50 ms, 2¹⁸ 3.94 → 3.11, 2²⁰ 6.40 → 5.01, unchanged above.
// RLN_B200_VARMSM_GLV: 1 (default) = for n ≤ 2²⁰, 0 = never, 2 = always.
static int var_msm_glv_mode() {
    static const int m = [] { const char* v = getenv("RLN_B200_VARMSM_GLV"); return v && *v ? atoi(v) : 1; }();
    return m;
}
static bool var_msm_use_glv(size_t n) { return var_msm_glv_mode() == 2 || (var_msm_glv_mode() == 1 && n <= ((size_t)1 << 20)); }
static int glv_windows(int c) { return 129 / c + 1; }   // c·K ≥ 129: the top signed digit and its carry stay below 2^(c−1)

static void msm_params(size_t n, int& c, int& K) {
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    c = lg - 6;
    if (c < 8) c = 8;
    if (c > 16) c = 16;
    K = (255 + c - 1) / c;  // K·c ≥ 255 keeps the top signed digit + carry below 2^(c−1)
}
static const u32 SLICE_MIN = 256, SLICE_MAX = 2048;  // a thread never sums more than `slice` points: heavier buckets are split
static const int SEGS = 2048;  // segments per window in the bucket reduction (16 buckets each at c = 16: 32 K short threads)
static const u32 SIZE_BINS = 4096;   // buckets are ordered by min(size, SIZE_BINS − 1): anything fuller is split into slices anyway
static const u32 SCAN_TILE = 1024;   // elements per CTA of the exclusive scan

VarMsmWorkspace* var_msm_workspace_create(size_t max_n) {
    VarMsmWorkspace* w = new VarMsmWorkspace();
    w->max_n = max_n;
    int c, K;
    size_t max_items = 0, max_b = 0;
    for (size_t n = 1; n <= max_n; n <<= 1) {
        msm_params(n, c, K);
        if (n * K > max_items) max_items = n * K;
        size_t b = (size_t)K << (c - 1);
        if (b > max_b) max_b = b;
        if (var_msm_glv_mode()) {
            msm_params(2 * n, c, K);
            K = glv_windows(c);
            if (2 * n * K > max_items) max_items = 2 * n * K;
            b = (size_t)K << (c - 1);
            if (b > max_b) max_b = b;
        }
    }
    msm_params(max_n, c, K);
    if (max_n * K > max_items) max_items = max_n * K;
    if (((size_t)K << (c - 1)) > max_b) max_b = (size_t)K << (c - 1);
    if (var_msm_glv_mode()) {
        msm_params(2 * max_n, c, K);
        K = glv_windows(c);
        if (2 * max_n * K > max_items) max_items = 2 * max_n * K;
        if (((size_t)K << (c - 1)) > max_b) max_b = (size_t)K << (c - 1);
    }
    w->max_buckets = max_b;
    ZK_CUDA_CHECK(cudaMalloc(&w->vals_out, 4 * max_items));
    ZK_CUDA_CHECK(cudaMalloc(&w->bucket_cnt, 4 * (max_b + 1)));
    ZK_CUDA_CHECK(cudaMalloc(&w->bucket_start, 4 * (max_b + 1)));
    ZK_CUDA_CHECK(cudaMalloc(&w->bucket_end, 4 * (max_b + 1)));
    ZK_CUDA_CHECK(cudaMalloc(&w->buckets, sizeof(G1XYZZ) * max_b));
    ZK_CUDA_CHECK(cudaMalloc(&w->size_bins, 4 * 2 * SIZE_BINS));
    ZK_CUDA_CHECK(cudaMalloc(&w->order, 4 * max_b));
    ZK_CUDA_CHECK(cudaMalloc(&w->slice_cnt, 4 * max_b));
    ZK_CUDA_CHECK(cudaMalloc(&w->slice_off, 4 * max_b));
    ZK_CUDA_CHECK(cudaMalloc(&w->scan_tmp, 4 * 4096));   // two scratch areas of ≤ 1024 tile sums (offsets 0 and 2048)
    w->max_slices = (max_items / SLICE_MIN > 2 * 65536 ? max_items / SLICE_MIN : 2 * 65536) + max_b + 1;   // slices only shrink below SLICE_MIN while there are < 64 K of them
    ZK_CUDA_CHECK(cudaMalloc(&w->partial, sizeof(G1XYZZ) * w->max_slices));
    ZK_CUDA_CHECK(cudaMalloc(&w->seg, sizeof(G1XYZZ) * 32 * SEGS));
    ZK_CUDA_CHECK(cudaMalloc(&w->win, sizeof(G1XYZZ) * 32));
    return w;
}
void var_msm_workspace_destroy(VarMsmWorkspace* w) {
    if (!w) return;
    cudaFree(w->vals_out);
    cudaFree(w->bucket_cnt); cudaFree(w->bucket_start); cudaFree(w->bucket_end); cudaFree(w->buckets);
    cudaFree(w->size_bins); cudaFree(w->order); cudaFree(w->scan_tmp);
    cudaFree(w->slice_cnt); cudaFree(w->slice_off); cudaFree(w->partial); cudaFree(w->seg); cudaFree(w->win);
    cudaFree(w->phi);
    delete w;
}

// Signed radix-2^c digits of every scalar, visited twice: PASS 0 counts the digits per (window, bucket) with one atomic each,
// PASS 1 (after the exclusive scan of the counts) writes the point index | sign << 31 of every non-zero digit to the next free
// slot of its bucket (cursor = a copy of the bucket starts, advanced by atomics; it ends up holding the bucket ends).  The digits
// are recomputed instead of stored: 40 integer instructions against 8 bytes of HBM traffic per digit each way.
// GLV: term i + h·n carries |k_h| of scalar i (h = 0, 1); the sign of the half flips every digit.
template <bool GLV, int PASS>
__global__ void __launch_bounds__(256) k_digits(const uint8_t* __restrict__ scalars, size_t n, int c, int K, u32* __restrict__ cnt_or_cursor,
                                                u32* __restrict__ vals_out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(scalars + 32 * i);
    uint4 a = q[0], b = q[1];
    u32 s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    u32 m[8];
#pragma unroll
    for (int t = 0; t < 8; t++) m[t] = FrCfg::p(t);
    while (Fr::raw_cmp(s, m) >= 0) Fr::raw_sub(s, s, m);  // ark reduces through Fr::into_bigint
    const u32 half = 1u << (c - 1);
    for (int h = 0; h < (GLV ? 2 : 1); h++) {
        u32 kh[8];
        bool neg = false;
        if (GLV) neg = glv::split(s, h, kh);
        else {
#pragma unroll
            for (int t = 0; t < 8; t++) kh[t] = s[t];
        }
        const u32 term = (u32)(i + (size_t)h * n);
        u32 carry = 0;
        for (int k = 0; k < K; k++) {
            const int bit = k * c, w = bit >> 5, sh = bit & 31;
            u32 v = 0;
            if (w < 8) {
                v = kh[w] >> sh;
                if (sh + c > 32 && w + 1 < 8) v |= kh[w + 1] << (32 - sh);
            }
            int d = (int)(v & ((1u << c) - 1)) + (int)carry;
            if (d > (int)half) { d -= (1 << c); carry = 1; } else carry = 0;
            if (d == 0) continue;
            const u32 bucket = (u32)k * half + (u32)((d < 0 ? -d : d) - 1);
            if (PASS == 0) atomicAdd(cnt_or_cursor + bucket, 1u);
            else vals_out[atomicAdd(cnt_or_cursor + bucket, 1u)] = term | (((d < 0) != neg) ? 0x80000000u : 0u);
        }
    }
}

// ---- exclusive scan of u32 (n ≤ SCAN_TILE²): per-tile scan + tile sums, scan of the tile sums by one CTA, add back -------------
__global__ void __launch_bounds__(256) k_scan_tiles(const u32* __restrict__ in, u32 n, u32* __restrict__ out, u32* __restrict__ tile_sum) {
    __shared__ u32 warp_tot[8];
    const u32 base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    u32 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = base + k < n ? in[base + k] : 0u;
    const u32 mine = v[0] + v[1] + v[2] + v[3];
    u32 incl = mine;
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (u32)d) incl += o;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    u32 before = 0;
    for (u32 w = 0; w < wid; w++) before += warp_tot[w];
    u32 run = before + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 255) tile_sum[blockIdx.x] = before + incl;
}
__global__ void __launch_bounds__(1024) k_scan_top(u32* __restrict__ tile_sum, u32 n_tiles) {   // in place, one CTA, n_tiles ≤ 1024
    __shared__ u32 warp_tot[32];
    const u32 t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const u32 mine = t < n_tiles ? tile_sum[t] : 0u;
    u32 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (u32)d) incl += o;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    u32 before = 0;
    for (u32 w = 0; w < wid; w++) before += warp_tot[w];
    if (t < n_tiles) tile_sum[t] = before + incl - mine;
}
__global__ void __launch_bounds__(256) k_scan_add(u32* __restrict__ out, u32 n, const u32* __restrict__ tile_off, u32* __restrict__ copy) {
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const u32 v = out[i] + tile_off[i / SCAN_TILE];
    out[i] = v;
    if (copy) copy[i] = v;   // the scatter cursors start as a copy of the bucket starts
}
static void exclusive_scan(const u32* in, u32 n, u32* out, u32* scratch, u32* copy, cudaStream_t s) {
    const u32 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles > 1024) throw CudaError(cudaErrorInvalidValue, "var msm: scan longer than 2^20 elements", __FILE__, __LINE__);
    k_scan_tiles<<<tiles, 256, 0, s>>>(in, n, out, scratch);
    k_scan_top<<<1, 1024, 0, s>>>(scratch, tiles);
    k_scan_add<<<(n + 255) / 256, 256, 0, s>>>(out, n, scratch, copy);
}

// ---- buckets in size order, fullest first: counting sort over bin(size) = SIZE_BINS − 1 − min(size, SIZE_BINS − 1) ---------------
__global__ void __launch_bounds__(256) k_size_hist(const u32* __restrict__ cnt, u32 n_buckets, u32* __restrict__ bins) {
    __shared__ u32 sh[SIZE_BINS];
    for (u32 t = threadIdx.x; t < SIZE_BINS; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += gridDim.x * blockDim.x) {
        const u32 sz = cnt[b];
        atomicAdd(&sh[SIZE_BINS - 1 - (sz < SIZE_BINS - 1 ? sz : SIZE_BINS - 1)], 1u);
    }
    __syncthreads();
    for (u32 t = threadIdx.x; t < SIZE_BINS; t += blockDim.x)
        if (sh[t]) atomicAdd(bins + t, sh[t]);
}
__global__ void __launch_bounds__(256) k_size_scatter(const u32* __restrict__ cnt, u32 n_buckets, u32* __restrict__ bin_cursor, u32* __restrict__ order) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    const u32 sz = cnt[b];
    order[atomicAdd(bin_cursor + (SIZE_BINS - 1 - (sz < SIZE_BINS - 1 ? sz : SIZE_BINS - 1)), 1u)] = b;
}
// number of slices of the i-th bucket in size order
__global__ void k_slice_counts(const u32* __restrict__ cnt, const u32* __restrict__ order, u32 n_buckets, u32 slice, u32* __restrict__ out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_buckets) out[i] = (cnt[order[i]] + slice - 1) / slice;
}

// φ(P) = (β·x, y); the point at infinity (0, 0) maps to itself
__global__ void k_phi_bases(const G1Affine* __restrict__ bases, size_t n, G1Affine* __restrict__ phi) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p = {ldg_fp(&bases[i].x), ldg_fp(&bases[i].y)};
    p.x = p.x * glv::beta();
    st_fp(&phi[i].x, p.x);
    st_fp(&phi[i].y, p.y);
}

// The hot kernel.  Work item = one slice of ≤ `slice` points of one bucket (slice ≈ twice the mean bucket size, so with uniform
// scalars every bucket is one slice); slices are numbered bucket after bucket in size order (fullest first), so the 32 slices of
// a warp hold (almost) the same number of points — no lanes idling behind a long one — and no thread ever walks a giant bucket
// alone (skewed scalars, or the few huge buckets of a short top window).  A single-slice bucket is written straight to its
// place; slices of split buckets go to `partial` for k_bucket_combine.
template <bool GLV>
__global__ void __launch_bounds__(128) k_bucket_sum(const G1Affine* __restrict__ bases, const u32* __restrict__ vals,
                                                    const u32* __restrict__ start, const u32* __restrict__ end, const u32* __restrict__ order,
                                                    const u32* __restrict__ cnt, const u32* __restrict__ off, u32 n_buckets, u32 slice,
                                                    G1XYZZ* __restrict__ partial, G1XYZZ* __restrict__ buckets,
                                                    const G1Affine* __restrict__ phi, u32 n) {
    // GLV: term index t < n is Pₜ, t ≥ n is φ(P_{t−n})
    auto base_of = [&](u32 t) -> const G1Affine* { return GLV && t >= n ? phi + (t - n) : bases + t; };
    const u32 sl = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 total = off[n_buckets - 1] + cnt[n_buckets - 1];
    if (sl >= total) return;
    // ordered position i with off[i] ≤ sl < off[i] + cnt[i]: the last i with off[i] ≤ sl (empty buckets sort last and own no slice)
    u32 lo_i = 0, hi_i = n_buckets - 1;
    while (lo_i < hi_i) {
        const u32 mid = (lo_i + hi_i + 1) >> 1;
        if (off[mid] <= sl) lo_i = mid; else hi_i = mid - 1;
    }
    const u32 b = order[lo_i];
    const u32 lo = start[b] + (sl - off[lo_i]) * slice;
    const u32 hi = min(end[b], lo + slice);
    G1XYZZ acc = G1XYZZ::infinity();
    u32 v = vals[lo];
    const G1Affine* bp = base_of(v & 0x7fffffffu);
    G1Affine nxt = {ldg_fp(&bp->x), ldg_fp(&bp->y)};
    if (v >> 31) nxt.y = nxt.y.neg();
    for (u32 t = lo; t < hi; t++) {
        G1Affine cur = nxt;
        if (t + 1 < hi) {  // prefetch the next point while the current addition runs
            v = vals[t + 1];
            bp = base_of(v & 0x7fffffffu);
            nxt = {ldg_fp(&bp->x), ldg_fp(&bp->y)};
            if (v >> 31) nxt.y = nxt.y.neg();
        }
        if (!cur.is_inf()) acc.add_affine(cur);
    }
    if (cnt[lo_i] == 1) buckets[b] = acc;
    else partial[sl] = acc;
}

__device__ __forceinline__ G1XYZZ shfl_xor_point(const G1XYZZ& p, int mask);
// split buckets only: bucket = Σ of its slices' partial sums.  L lanes per ordered position — a lane-strided loop and a shuffle
// tree, so even a bucket of millions of points is combined in ≈ slices/32 + 5 additions.  Positions are in size order (fullest
// first), so their slice counts never increase: the kernel finds where they drop to ≤ 16, ≤ 8 and ≤ 4 and gives those stretches
// 16 / 8 / 4 lanes per bucket (2 / 4 / 8 buckets per warp, trees of 4 / 3 / 2 levels).  Small MSMs cut their slices short to fill
// the chip, so nearly every bucket is split in a handful of slices there, next to the few giant buckets of the short top window
// (2^16 terms: 12 K warps of mostly idle lanes with 32 lanes per bucket).  Buckets at the clamp of the size ordering
// (≥ SIZE_BINS − 1 points, unordered among themselves) have more than 16 slices whenever slice ≤ 255; otherwise every position
// keeps 32 lanes.
__device__ __forceinline__ u32 first_at_most(const u32* __restrict__ cnt, u32 n, u32 bound) {   // cnt is non-increasing
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (cnt[mid] <= bound) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__global__ void __launch_bounds__(128) k_bucket_combine(const G1XYZZ* __restrict__ partial, const u32* __restrict__ order,
                                                        const u32* __restrict__ cnt, const u32* __restrict__ off, u32 n_positions,
                                                        u32 slice, G1XYZZ* __restrict__ buckets) {
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    u32 p16 = n_positions, p8 = n_positions, p4 = n_positions;
    if (16 * slice < SIZE_BINS - 1) {                       // the order is exact wherever counts are ≤ 16
        p16 = first_at_most(cnt, n_positions, 16);
        p8 = first_at_most(cnt, n_positions, 8);
        p4 = first_at_most(cnt, n_positions, 4);
    }
    const u32 w16 = (p8 - p16 + 1) / 2, w8 = (p4 - p8 + 3) / 4, w4 = (n_positions - p4 + 7) / 8;
    u32 L, i, end;
    if (warp < p16) { L = 32; i = warp; end = p16; }
    else if (warp < p16 + w16) { L = 16; i = p16 + (warp - p16) * 2 + lane / 16; end = p8; }
    else if (warp < p16 + w16 + w8) { L = 8; i = p8 + (warp - p16 - w16) * 4 + lane / 8; end = p4; }
    else if (warp < p16 + w16 + w8 + w4) { L = 4; i = p4 + (warp - p16 - w16 - w8) * 8 + lane / 4; end = n_positions; }
    else return;                                           // the whole warp: the grid is sized for 32 lanes per position
    const u32 l = lane % L;
    const bool live = i < end;
    const u32 c = live ? cnt[i] : 0, o = live ? off[i] : 0;
    if (!__any_sync(0xffffffffu, c > 1)) return;           // single-slice buckets were written by k_bucket_sum, empty ones by the memset
    G1XYZZ acc = G1XYZZ::infinity();
    if (c > 1)
        for (u32 t = l; t < c; t += L) acc.add(partial[o + t]);
    for (u32 m = 16; m >= 1; m >>= 1) {
        if (m >= L) continue;                              // uniform over the warp
        G1XYZZ other = shfl_xor_point(acc, (int)m);
        acc.add(other);
    }
    if (l == 0 && c > 1) buckets[order[i]] = acc;
}

// segment s of window k covers buckets [s·L, (s+1)·L): W = Σ (b+1)·B_b over the segment
__global__ void __launch_bounds__(64) k_segment_reduce(const G1XYZZ* __restrict__ buckets, u32 half, u32 seg_len, G1XYZZ* __restrict__ seg) {
    const u32 k = blockIdx.y;
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n_seg = half / seg_len;
    if (s >= n_seg) return;
    const G1XYZZ* B = buckets + (size_t)k * half + (size_t)s * seg_len;
    G1XYZZ run = G1XYZZ::infinity(), acc = G1XYZZ::infinity();
    for (int b = (int)seg_len - 1; b >= 0; b--) {
        run.add(B[b]);
        acc.add(run);  // after the loop: acc = Σ (b+1)·B_b (local index), run = Σ B_b
    }
    // shift the local weights by s·seg_len (< 2^16: a short double-and-add, not the 256-bit ladder)
    const u32 off = s * seg_len;
    if (off) {
        G1XYZZ m = G1XYZZ::infinity();
        for (int i = 31 - __clz(off); i >= 0; i--) {
            m = m.dbl();
            if ((off >> i) & 1) m.add(run);
        }
        acc.add(m);
    }
    seg[(size_t)k * n_seg + s] = acc;
}

__device__ __forceinline__ G1XYZZ shfl_xor_point(const G1XYZZ& p, int mask) {
    G1XYZZ r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.X.l[i] = __shfl_xor_sync(0xffffffffu, p.X.l[i], mask);
        r.Y.l[i] = __shfl_xor_sync(0xffffffffu, p.Y.l[i], mask);
        r.ZZ.l[i] = __shfl_xor_sync(0xffffffffu, p.ZZ.l[i], mask);
        r.ZZZ.l[i] = __shfl_xor_sync(0xffffffffu, p.ZZZ.l[i], mask);
    }
    return r;
}
// one block per window: warp-shuffle tree over the segment sums, then across warps through shared memory
__global__ void __launch_bounds__(256) k_window_reduce(const G1XYZZ* __restrict__ seg, u32 n_seg, G1XYZZ* __restrict__ win) {
    __shared__ G1XYZZ sh[8];
    const u32 k = blockIdx.x;
    G1XYZZ acc = G1XYZZ::infinity();
    for (u32 s = threadIdx.x; s < n_seg; s += blockDim.x) acc.add(seg[(size_t)k * n_seg + s]);
    for (int m = 16; m >= 1; m >>= 1) {
        G1XYZZ o = shfl_xor_point(acc, m);
        acc.add(o);
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        G1XYZZ t = sh[0];
        for (u32 w = 1; w < blockDim.x / 32; w++) t.add(sh[w]);
        win[k] = t;
    }
}

// c doublings of an XYZZ point through Jacobian coordinates: (X, Y, ZZ, ZZZ) ↦ (X·ZZ², Y·ZZ³, Z = ZZZ) costs 4 products, a Jacobian
// doubling with a = 0 (dbl-2009-l) 2M + 5S against XYZZ's 6M + 3S, and the way back is ZZ = Z², ZZZ = Z³.
__device__ __forceinline__ Fq shfl_fq(const Fq& v, int src) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}
// The chain runs on ONE WARP whose lanes all hold the same point: the seven products of a doubling have a dependency depth of
// three — {X², Y², Y·Z}, then {(3X²)², (Y²)², (X + Y²)²}, then E·(D − X₃) — so three lanes take one product each per level and
// hand the results round by shuffles; the last level and the additions run redundantly on every lane (no broadcast needed).
// ≈ 2 900 cycles per doubling instead of ≈ 6 100: the 112 … 240 doublings of the final Horner pass are the floor of a small MSM.
__device__ G1XYZZ xyzz_dbl_n_warp(const G1XYZZ& p, int n) {
    if (p.is_inf() || n <= 0) return p;                 // uniform: every lane holds the same point
    const u32 lane = threadIdx.x & 31;
    const Fq zz2 = p.ZZ.sqr();
    Fq X = p.X * zz2, Y = p.Y * (zz2 * p.ZZ), Z = p.ZZZ;
    for (int i = 0; i < n; i++) {
        if (Y.is_zero()) return G1XYZZ::infinity();
        const Fq a1 = lane == 0 ? X : Y;
        const Fq b1 = lane == 0 ? X : (lane == 1 ? Y : Z);
        const Fq r1 = a1 * b1;                          // lane 0: X², lane 1: Y², lane 2: Y·Z
        const Fq A = shfl_fq(r1, 0), Bq = shfl_fq(r1, 1), YZ = shfl_fq(r1, 2);
        const Fq E = A.dbl() + A;
        const Fq xb = X + Bq;
        const Fq a2 = lane == 0 ? E : (lane == 1 ? Bq : xb);
        const Fq r2 = a2.sqr();                         // lane 0: E², lane 1: (Y²)², lane 2: (X + Y²)²
        const Fq F = shfl_fq(r2, 0), C = shfl_fq(r2, 1), T = shfl_fq(r2, 2);
        const Fq D = (T - A - C).dbl();
        const Fq X3 = F - D.dbl();
        const Fq C8 = C.dbl().dbl().dbl();
        Y = E * (D - X3) - C8;
        X = X3;
        Z = YZ.dbl();
    }
    const Fq ZZ = Z.sqr();
    return {X, Y, ZZ, ZZ * Z};
}
// one warp; every lane holds the running point
__global__ void __launch_bounds__(32) k_horner(const G1XYZZ* __restrict__ win, int c, int K, uint8_t* __restrict__ out) {
    G1XYZZ t = win[K - 1];
    for (int k = K - 2; k >= 0; k--) {
        t = xyzz_dbl_n_warp(t, c);
        t.add(win[k]);
    }
    G1Affine a = t.to_affine();
    u32 x[8] = {0, 0, 0, 0, 0, 0, 0, 0}, y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!a.is_inf()) { a.x.to_canonical(x); a.y.to_canonical(y); }
    if (threadIdx.x == 0)
        for (int i = 0; i < 8; i++) { reinterpret_cast<u32*>(out)[i] = x[i]; reinterpret_cast<u32*>(out)[8 + i] = y[i]; }
}

void launch_var_msm_g1(VarMsmWorkspace* w, const G1Affine* d_bases, const uint8_t* d_scalars, size_t n, uint8_t* d_result, cudaStream_t s) {
    if (n == 0) { ZK_CUDA_CHECK(cudaMemsetAsync(d_result, 0, 64, s)); return; }
    if (n > w->max_n) throw CudaError(cudaErrorInvalidValue, "var msm: n exceeds workspace", __FILE__, __LINE__);
    int c, K;
    const bool use_glv = var_msm_use_glv(n);
    const size_t terms = use_glv ? 2 * n : n;
    msm_params(terms, c, K);
    if (use_glv) K = glv_windows(c);
    const u32 half = 1u << (c - 1);
    const size_t items = terms * (size_t)K, n_buckets = (size_t)K * half;
    ZK_CUDA_CHECK(cudaMemsetAsync(w->bucket_cnt, 0, 4 * n_buckets, s));
    ZK_CUDA_CHECK(cudaMemsetAsync(w->size_bins, 0, 4 * SIZE_BINS, s));
    const unsigned dg = (unsigned)((n + 255) / 256);
    if (use_glv) {
        if (w->phi_cap < n) {
            ZK_CUDA_CHECK(cudaStreamSynchronize(s));
            if (w->phi) ZK_CUDA_CHECK(cudaFree(w->phi));
            ZK_CUDA_CHECK(cudaMalloc(&w->phi, sizeof(G1Affine) * n));
            w->phi_cap = n;
        }
        k_phi_bases<<<dg, 256, 0, s>>>(d_bases, n, w->phi);
        k_digits<true, 0><<<dg, 256, 0, s>>>(d_scalars, n, c, K, w->bucket_cnt, nullptr);
    } else {
        k_digits<false, 0><<<dg, 256, 0, s>>>(d_scalars, n, c, K, w->bucket_cnt, nullptr);
    }
    // bucket starts = exclusive scan of the counts; the scatter cursors (bucket_end) start there and finish at the bucket ends
    exclusive_scan(w->bucket_cnt, (u32)n_buckets, w->bucket_start, w->scan_tmp, w->bucket_end, s);
    if (use_glv) k_digits<true, 1><<<dg, 256, 0, s>>>(d_scalars, n, c, K, w->bucket_end, w->vals_out);
    else k_digits<false, 1><<<dg, 256, 0, s>>>(d_scalars, n, c, K, w->bucket_end, w->vals_out);
    // buckets in size order
    u32* bins = w->size_bins;
    u32* bin_start = w->size_bins + SIZE_BINS;
    k_size_hist<<<(unsigned)((n_buckets + 2047) / 2048 < 296 ? (n_buckets + 2047) / 2048 : 296), 256, 0, s>>>(w->bucket_cnt, (u32)n_buckets, bins);
    exclusive_scan(bins, SIZE_BINS, bin_start, w->scan_tmp + 2048, nullptr, s);
    k_size_scatter<<<(unsigned)((n_buckets + 255) / 256), 256, 0, s>>>(w->bucket_cnt, (u32)n_buckets, bin_start, w->order);
    // slice ≈ twice the mean bucket size at large n (one thread per bucket); at small n there are too few buckets to fill the chip
    // (2^16: 12 K buckets of 128 points on 75 K thread slots), so slices shrink until there are ≈ 64 K of them and the split
    // buckets are re-assembled by k_bucket_combine
    u32 slice = SLICE_MIN;
    while (slice < SLICE_MAX && (size_t)slice * n_buckets < 2 * items) slice <<= 1;
    while (slice > 16 && items / slice < 65536) slice >>= 1;
    k_slice_counts<<<(unsigned)((n_buckets + 255) / 256), 256, 0, s>>>(w->bucket_cnt, w->order, (u32)n_buckets, slice, w->slice_cnt);
    exclusive_scan(w->slice_cnt, (u32)n_buckets, w->slice_off, w->scan_tmp, nullptr, s);
    const size_t max_slices = items / slice + n_buckets;   // upper bound known on the host; threads beyond the real total exit
    ZK_CUDA_CHECK(cudaMemsetAsync(w->buckets, 0, sizeof(G1XYZZ) * n_buckets, s));   // all-zero XYZZ = infinity: the empty buckets
    if (use_glv)
        k_bucket_sum<true><<<(unsigned)((max_slices + 127) / 128), 128, 0, s>>>(d_bases, w->vals_out, w->bucket_start, w->bucket_end, w->order,
                                                                                w->slice_cnt, w->slice_off, (u32)n_buckets, slice, w->partial,
                                                                                w->buckets, w->phi, (u32)n);
    else
        k_bucket_sum<false><<<(unsigned)((max_slices + 127) / 128), 128, 0, s>>>(d_bases, w->vals_out, w->bucket_start, w->bucket_end, w->order,
                                                                                 w->slice_cnt, w->slice_off, (u32)n_buckets, slice, w->partial,
                                                                                 w->buckets, nullptr, (u32)n);
    // a split bucket holds more than `slice` points, so at most items / slice ordered positions can be split
    const size_t n_split_max = items / slice < n_buckets ? items / slice : n_buckets;
    if (n_split_max)
        k_bucket_combine<<<(unsigned)((n_split_max * 32 + 127) / 128), 128, 0, s>>>(w->partial, w->order, w->slice_cnt, w->slice_off,
                                                                                    (u32)n_split_max, slice, w->buckets);
    const u32 n_seg = half < (u32)SEGS ? half : (u32)SEGS;
    const u32 seg_len = half / n_seg;
    k_segment_reduce<<<dim3((n_seg + 63) / 64, K), 64, 0, s>>>(w->buckets, half, seg_len, w->seg);
    k_window_reduce<<<K, 256, 0, s>>>(w->seg, n_seg, w->win);
    k_horner<<<1, 32, 0, s>>>(w->win, c, K, d_result);
}

// ------------------------------------------------------------------------------------------- point I/O helpers
__device__ __forceinline__ Fq fq_from_bytes(const uint8_t* p, bool mask_flags) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    u32 c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (mask_flags) c[7] &= 0x3fffffffu;
    return Fq::from_canonical(c);
}
__global__ void k_g1_from_bytes(const uint8_t* __restrict__ in, G1Affine* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = in + 64 * i;
    G1Affine a = G1Affine::infinity();
    if (!(p[63] & 0x40)) a = {fq_from_bytes(p, false), fq_from_bytes(p + 32, true)};
    st_fp(&out[i].x, a.x);
    st_fp(&out[i].y, a.y);
}
__global__ void k_g2_from_bytes(const uint8_t* __restrict__ in, G2Affine* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = in + 128 * i;
    G2Affine a = G2Affine::infinity();
    if (!(p[127] & 0x40)) a = {{fq_from_bytes(p, false), fq_from_bytes(p + 32, false)}, {fq_from_bytes(p + 64, false), fq_from_bytes(p + 96, true)}};
    st_fp(&out[i].x.a, a.x.a); st_fp(&out[i].x.b, a.x.b);
    st_fp(&out[i].y.a, a.y.a); st_fp(&out[i].y.b, a.y.b);
}
void launch_g1_from_bytes(const uint8_t* d_in, G1Affine* d_out, size_t n, cudaStream_t s) {
    if (n) k_g1_from_bytes<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, d_out, n);
}
void launch_g2_from_bytes(const uint8_t* d_in, G2Affine* d_out, size_t n, cudaStream_t s) {
    if (n) k_g2_from_bytes<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, d_out, n);
}

__global__ void __launch_bounds__(64) k_g1_mul_gen(const uint8_t* __restrict__ scalars, G1Affine* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 k[8];
    for (int t = 0; t < 8; t++) k[t] = reinterpret_cast<const u32*>(scalars + 32 * i)[t];
    G1Affine g = {Fq::from_u32(1), Fq::from_u32(2)};
    out[i] = G1XYZZ::from_affine(g).mul(k).to_affine();
}
void launch_g1_mul_gen(const uint8_t* d_scalars, G1Affine* d_out, size_t n, cudaStream_t s) {
    if (n) k_g1_mul_gen<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(d_scalars, d_out, n);
}

}  // namespace zk
