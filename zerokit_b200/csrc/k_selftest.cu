// Self-test / calibration kernels: element-wise field operations through the PTX carry-chain path
// (parity-checked against the oracle by tests/test_gpu_field.py) and a multiply-throughput probe
// used by bench.py to state the INT32-pipe ceiling next to the HBM roofline.
#include "../../include/rln_b200.h"
#include "device_api.hpp"

namespace zk {
extern std::atomic<uint64_t> g_launch_count;

template <class F>
__global__ void k_field_op(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, uint8_t* __restrict__ out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 x[8], y[8], r[8];
    for (int k = 0; k < 8; k++) {
        x[k] = reinterpret_cast<const u32*>(a + 32 * i)[k];
        y[k] = reinterpret_cast<const u32*>(b + 32 * i)[k];
    }
    F fx = F::from_canonical(x), fy = F::from_canonical(y), fr;
    if (op == 0) fr = fx * fy;
    else if (op == 1) fr = fx + fy;
    else if (op == 2) fr = fx - fy;
    else if (op == 3) { F::mul_portable(fr.l, fx.l, fy.l); }   // portable CIOS path on the device, for cross-checking
    else if (op == 4) fr = fx.inv();
    else if (op == 5) fr = fx.sqr();                                           // dedicated squaring schedule
    else if (op == 6) fr = F::sub_prod(fx, fx, fy, fy);                         // x² − y², one reduction
    else if (op == 7) {                                                         // 4-term dot product with lazily negated operands: y² − x²
        const F u[4] = {fx, fy, fx.neg_lazy(), fy.neg_lazy()}, v[4] = {fy, fy, fx, fx};
        fr = F::template dot<4>(u, v);
    } else {                                                                    // 5 terms, all operands at their maximum (p and x)
        const F u[5] = {fx.neg_lazy(), F::zero().neg_lazy(), fy, fx, F::zero().neg_lazy()}, v[5] = {fy, fx, fx, fx, fy};
        fr = F::template dot<5>(u, v);                                          // −xy + 0 + xy + x² + 0
    }
    fr.to_canonical(r);
    for (int k = 0; k < 8; k++) reinterpret_cast<u32*>(out + 32 * i)[k] = r[k];
}

// 4 independent multiply chains per thread, `iters` steps each
template <class F>
__global__ void __launch_bounds__(256) k_mul_throughput(F* __restrict__ data, int iters) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    F a = data[4 * i], b = data[4 * i + 1], c = data[4 * i + 2], d = data[4 * i + 3];
    const F m = a + b;
    for (int t = 0; t < iters; t++) {
        a = a * m;
        b = b * m;
        c = c * m;
        d = d * m;
    }
    data[4 * i] = a; data[4 * i + 1] = b; data[4 * i + 2] = c; data[4 * i + 3] = d;
}
// KIND 1 = dedicated squaring, 2 = two-term dot product (counts as two products), 3 = Fq2 product (counts as three)
template <int KIND>
__global__ void __launch_bounds__(256) k_op_throughput(Fq* __restrict__ data, int iters) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    Fq a = data[4 * i], b = data[4 * i + 1], c = data[4 * i + 2], d = data[4 * i + 3];
    const Fq m = a + b;
    for (int t = 0; t < iters; t++) {
        if (KIND == 1) { a = a.sqr(); b = b.sqr(); c = c.sqr(); d = d.sqr(); }
        else if (KIND == 2) { Fq na = Fq::dot2(a, m, b, c), nc = Fq::dot2(c, m, d, a); b = a; d = c; a = na; c = nc; }
        else { Fq2 x = Fq2{a, b} * Fq2{c, d}; Fq2 y = Fq2{c, d} * Fq2{m, a}; a = x.a; b = x.b; c = y.a; d = y.b; }
    }
    data[4 * i] = a; data[4 * i + 1] = b; data[4 * i + 2] = c; data[4 * i + 3] = d;
}
// pipe probe: mode 0 = wide integer MADs only, 1 = FP64 FMAs only, 2 = both interleaved (do the pipes overlap?)
template <int MODE>
__global__ void __launch_bounds__(256) k_pipe_probe(u64* __restrict__ out, int iters) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u64 a0 = t + 1, a1 = t + 2, a2 = t + 3, a3 = t + 5, a4 = t + 7, a5 = t + 11, a6 = t + 13, a7 = t + 17;
    double d0 = t + 0.5, d1 = t + 1.5, d2 = t + 2.5, d3 = t + 3.5, d4 = t + 4.5, d5 = t + 5.5, d6 = t + 6.5, d7 = t + 7.5;
    const u32 m = 0x9e3779b9u + t, n = 0x85ebca6bu ^ t;
    const double x = 1.0000001, y = 0.9999999;
    for (int i = 0; i < iters; i++) {
        if (MODE != 1) {
            // one multiplicand is the low word of the accumulator itself: a loop-invariant product would be hoisted by ptxas and
            // the loop would measure 64-bit additions, not the multiplier (an earlier version of this probe did exactly that)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a0) : "r"((u32)a0), "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a1) : "r"((u32)a1), "r"(n));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a2) : "r"((u32)a2), "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a3) : "r"((u32)a3), "r"(n));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a4) : "r"((u32)a4), "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a5) : "r"((u32)a5), "r"(n));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a6) : "r"((u32)a6), "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a7) : "r"((u32)a7), "r"(n));
        }
        if (MODE != 0) {
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d0) : "d"(x), "d"(y));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d1) : "d"(x), "d"(y));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d2) : "d"(x), "d"(y));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d3) : "d"(x), "d"(y));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d4) : "d"(x), "d"(y));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d5) : "d"(x), "d"(y));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d6) : "d"(x), "d"(y));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d7) : "d"(x), "d"(y));
        }
    }
    out[t] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ (u64)__double_as_longlong(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7);
}

// ---- batched-affine probe (measured, not estimated) -------------------------------------------------------------------
// Bucket-style accumulation, M independent running sums per thread kept in global memory (as Pippenger buckets are), one random-ish
// point added to each per round.  mode 0: XYZZ mixed additions (what every MSM kernel of this library does: 8M + 2S, no inversion);
// mode 1 / 2: affine additions that share ONE inversion per thread and round (Montgomery's trick over the M denominators):
// 3 products per addition for the trick + 2M + 1S for the addition itself + the inversion / M; mode 1 inverts by Fermat (a^(q−2),
// ≈ 380 products), mode 2 by the binary extended Euclid below.  Probe only: no special cases (the points are distinct by
// construction), results are folded into a checksum so nothing is optimised away.
__device__ void raw_shr1(u32* a, u32 top) {
#pragma unroll
    for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[7] = (a[7] >> 1) | (top << 31);
}
__device__ Fq fq_inv_egcd(const Fq& a) {   // a ≠ 0, Montgomery in and out
    u32 u[8], v[8], x1[8], x2[8], p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = a.l[i]; p[i] = FqCfg::p(i); v[i] = p[i]; x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    auto is_one = [](const u32* w) { u32 t = w[0] ^ 1u; for (int i = 1; i < 8; i++) t |= w[i]; return t == 0; };
    auto halve_mod = [&](u32* x) {   // x/2 mod p
        u32 carry = 0;
        if (x[0] & 1) carry = Fq::raw_add(x, x, p);
        raw_shr1(x, carry);
    };
    while (!is_one(u) && !is_one(v)) {
        while (!(u[0] & 1)) { raw_shr1(u, 0); halve_mod(x1); }
        while (!(v[0] & 1)) { raw_shr1(v, 0); halve_mod(x2); }
        if (Fq::raw_cmp(u, v) >= 0) {
            Fq::raw_sub(u, u, v);
            if (Fq::raw_sub(x1, x1, x2)) Fq::raw_add(x1, x1, p);
        } else {
            Fq::raw_sub(v, v, u);
            if (Fq::raw_sub(x2, x2, x1)) Fq::raw_add(x2, x2, p);
        }
    }
    Fq r;   // raw inverse of a·R is a⁻¹·R⁻¹: one Montgomery product by R³ brings it back to a⁻¹·R
    const u32* w = is_one(u) ? x1 : x2;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = w[i];
    const Fq r3 = Fq::rsquared() * Fq::rsquared();
    return r * r3;
}
__global__ void k_probe_points(G1Affine* __restrict__ pts, u32 n) {   // pts[i] = (i + 2)·G
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const G1Affine g = {Fq::from_u32(1), Fq::from_u32(2)};
    u32 k[8] = {i + 2, 0, 0, 0, 0, 0, 0, 0};
    pts[i] = G1XYZZ::from_affine(g).mul(k).to_affine();
}
__global__ void k_egcd_check(const G1Affine* __restrict__ pts, u32 n, u32* __restrict__ bad) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fq a = pts[i].x * pts[(i * 7 + 3) % n].y + Fq::from_u32(i + 1);
    if (a.is_zero()) return;
    if (!(fq_inv_egcd(a) == a.inv()) || !(fq_inv_egcd(a) * a == Fq::one())) atomicAdd(bad, 1u);
}
template <int MODE>
__global__ void __launch_bounds__(128) k_affine_probe(const G1Affine* __restrict__ pts, u32 n_pts, int M, int rounds, G1XYZZ* __restrict__ acc_x,
                                                      G1Affine* __restrict__ acc_a, Fq* __restrict__ pref, u32* __restrict__ check) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
    // start every running sum at a different point so that no addition meets its own operand
    for (int m = 0; m < M; m++) {
        const G1Affine p0 = pts[(t * 7 + m * 131 + 1) % n_pts];
        if (MODE == 0) acc_x[(size_t)m * T + t] = G1XYZZ::from_affine(p0);
        else acc_a[(size_t)m * T + t] = p0;
    }
    for (int r = 0; r < rounds; r++) {
        if (MODE == 0) {
            for (int m = 0; m < M; m++) {
                G1XYZZ a = acc_x[(size_t)m * T + t];
                a.add_affine(pts[(t * 13 + m * 17 + r * 29 + 5) % n_pts]);
                acc_x[(size_t)m * T + t] = a;
            }
        } else {
            Fq run = Fq::one();
            for (int m = 0; m < M; m++) {   // forward: prefix products of the denominators
                const Fq d = ldg_fp(&pts[(t * 13 + m * 17 + r * 29 + 5) % n_pts].x) - ld_fp(&acc_a[(size_t)m * T + t].x);
                pref[(size_t)m * T + t] = run;
                run = run * (d.is_zero() ? Fq::one() : d);
            }
            Fq inv = MODE == 1 ? run.inv() : fq_inv_egcd(run);
            for (int m = M - 1; m >= 0; m--) {   // backward: individual inverses, then the additions
                const G1Affine q = pts[(t * 13 + m * 17 + r * 29 + 5) % n_pts];
                G1Affine a = acc_a[(size_t)m * T + t];
                Fq d = q.x - a.x;
                if (d.is_zero()) d = Fq::one();
                const Fq dinv = inv * pref[(size_t)m * T + t];
                inv = inv * d;
                const Fq lam = (q.y - a.y) * dinv;
                const Fq x3 = lam.sqr() - a.x - q.x;
                a.y = lam * (a.x - x3) - a.y;
                a.x = x3;
                acc_a[(size_t)m * T + t] = a;
            }
        }
    }
    u32 c = 0;
    for (int m = 0; m < M; m++) c ^= MODE == 0 ? acc_x[(size_t)m * T + t].X.l[0] : acc_a[(size_t)m * T + t].x.l[0];
    check[t] = c;
}
// ---- latency probe: cycles per DEPENDENT product in a lone warp (what the witness VM and every one-thread-per-proof chain pay) -----
// kind 0: mul_ptx (865 cycles measured), 1: mul_portable (CIOS in C: 1 131), 3: sqr_ptx (682), 4: modular addition (76);
// `lanes` live lanes of ONE warp on one SM (the figure does not depend on it)
template <int KIND>
__global__ void k_latency_probe(Fq* __restrict__ data, int iters, long long* __restrict__ cycles) {
    Fq a = data[threadIdx.x], b = data[32 + threadIdx.x];
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        Fq r;
#if ZK_PTX
        if (KIND == 0) Fq::mul_ptx(r.l, a.l, b.l);
        else if (KIND == 1) Fq::mul_portable(r.l, a.l, b.l);
        else if (KIND == 3) Fq::sqr_ptx(r.l, a.l);
        else r = a + b;
#else
        r = a + b;
#endif
        a = r;
    }
    const long long t1 = clock64();
    data[threadIdx.x] = a;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
}  // namespace zk

using namespace zk;
extern "C" {
double rlnb200_latency_probe(int kind, int lanes, int iters) {
    try {
        Fq* d = nullptr;
        long long* c = nullptr;
        ZK_CUDA_CHECK(cudaMalloc(&d, sizeof(Fq) * 64));
        ZK_CUDA_CHECK(cudaMalloc(&c, 8));
        ZK_CUDA_CHECK(cudaMemset(d, 0x11, sizeof(Fq) * 64));
        if (lanes < 1) lanes = 1;
        if (lanes > 32) lanes = 32;
        for (int rep = 0; rep < 2; rep++) {
            switch (kind) {
                case 0: k_latency_probe<0><<<1, lanes>>>(d, iters, c); break;
                case 1: k_latency_probe<1><<<1, lanes>>>(d, iters, c); break;
                case 3: k_latency_probe<3><<<1, lanes>>>(d, iters, c); break;
                default: k_latency_probe<4><<<1, lanes>>>(d, iters, c); break;
            }
        }
        long long h = 0;
        ZK_CUDA_CHECK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost));
        g_launch_count += 2;
        cudaFree(d); cudaFree(c);
        return (double)h / iters;
    } catch (const CudaError&) {
        return -1.0;
    }
}
// additions per second of the batched-affine probe; mode 0 XYZZ, 1 affine + Fermat inversion, 2 affine + binary EGCD; M running sums
// per thread.  mode 3: self-check of the EGCD inversion against Fermat on 4 096 values (returns the number of mismatches).
double rlnb200_affine_batch_probe(int mode, int M, int rounds) {
    try {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const u32 n_pts = 4096;
        G1Affine* pts = nullptr;
        ZK_CUDA_CHECK(cudaMalloc(&pts, sizeof(G1Affine) * n_pts));
        k_probe_points<<<n_pts / 128, 128>>>(pts, n_pts);
        if (mode == 3) {
            u32* bad = nullptr;
            ZK_CUDA_CHECK(cudaMalloc(&bad, 4));
            ZK_CUDA_CHECK(cudaMemset(bad, 0, 4));
            k_egcd_check<<<n_pts / 128, 128>>>(pts, n_pts, bad);
            u32 h = 0;
            ZK_CUDA_CHECK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
            cudaFree(bad); cudaFree(pts);
            return (double)h;
        }
        const u32 T = (u32)sms * 4 * 128;   // 4 CTAs of 128 threads per SM
        G1XYZZ* ax = nullptr; G1Affine* aa = nullptr; Fq* pref = nullptr; u32* check = nullptr;
        ZK_CUDA_CHECK(cudaMalloc(&ax, sizeof(G1XYZZ) * (size_t)M * T));
        ZK_CUDA_CHECK(cudaMalloc(&aa, sizeof(G1Affine) * (size_t)M * T));
        ZK_CUDA_CHECK(cudaMalloc(&pref, sizeof(Fq) * (size_t)M * T));
        ZK_CUDA_CHECK(cudaMalloc(&check, 4 * (size_t)T));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&](int r) {
            if (mode == 0) k_affine_probe<0><<<T / 128, 128>>>(pts, n_pts, M, r, ax, aa, pref, check);
            else if (mode == 1) k_affine_probe<1><<<T / 128, 128>>>(pts, n_pts, M, r, ax, aa, pref, check);
            else k_affine_probe<2><<<T / 128, 128>>>(pts, n_pts, M, r, ax, aa, pref, check);
        };
        run(1);
        cudaEventRecord(e0);
        run(rounds);
        cudaEventRecord(e1);
        ZK_CUDA_CHECK(cudaEventSynchronize(e1));
        g_launch_count += 3;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaFree(ax); cudaFree(aa); cudaFree(pref); cudaFree(check); cudaFree(pts);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return (double)T * M * rounds / (ms * 1e-3);
    } catch (const CudaError&) {
        return -1.0;
    }
}
int rlnb200_field_op(int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out, RlnString* err) {
    try {
        void *da = nullptr, *db = nullptr, *dout = nullptr;
        ZK_CUDA_CHECK(cudaMalloc(&da, 32 * n)); ZK_CUDA_CHECK(cudaMalloc(&db, 32 * n)); ZK_CUDA_CHECK(cudaMalloc(&dout, 32 * n));
        ZK_CUDA_CHECK(cudaMemcpy(da, a, 32 * n, cudaMemcpyHostToDevice));
        ZK_CUDA_CHECK(cudaMemcpy(db, b, 32 * n, cudaMemcpyHostToDevice));
        unsigned g = (unsigned)((n + 127) / 128);
        if (field == 0) k_field_op<Fr><<<g, 128>>>(op, (uint8_t*)da, (uint8_t*)db, n, (uint8_t*)dout);
        else k_field_op<Fq><<<g, 128>>>(op, (uint8_t*)da, (uint8_t*)db, n, (uint8_t*)dout);
        g_launch_count++;
        ZK_CUDA_CHECK(cudaMemcpy(out, dout, 32 * n, cudaMemcpyDeviceToHost));
        cudaFree(da); cudaFree(db); cudaFree(dout);
        return 0;
    } catch (const CudaError& e) {
        if (err) {
            const char* m = cudaGetErrorString(e.code);
            size_t l = strlen(m);
            err->ptr = (uint8_t*)malloc(l + 1); memcpy(err->ptr, m, l + 1); err->len = l; err->cap = l + 1;
        }
        return -1;
    }
}
// per-SM-clock warp instructions of each kind issued per second for the probe above; out[0] = wide MADs/s, out[1] = DFMAs/s
int rlnb200_pipe_probe(int mode, int iters, double out[2]) {
    try {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t threads = (size_t)sms * 2048;
        u64* d = nullptr;
        ZK_CUDA_CHECK(cudaMalloc(&d, 8 * threads));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&](int it) {
            unsigned g = (unsigned)(threads / 256);
            if (mode == 0) k_pipe_probe<0><<<g, 256>>>(d, it);
            else if (mode == 1) k_pipe_probe<1><<<g, 256>>>(d, it);
            else k_pipe_probe<2><<<g, 256>>>(d, it);
        };
        run(64);
        cudaEventRecord(e0);
        run(iters);
        cudaEventRecord(e1);
        ZK_CUDA_CHECK(cudaEventSynchronize(e1));
        g_launch_count += 2;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
        const double ops = (double)threads * 8.0 * iters / (ms * 1e-3);
        out[0] = mode != 1 ? ops : 0.0;
        out[1] = mode != 0 ? ops : 0.0;
        return 0;
    } catch (const CudaError&) {
        return -1;
    }
}
// product-equivalents per second of the squaring / dot-product / Fq2 schedules (kind as in k_op_throughput)
double rlnb200_op_throughput(int kind, int iters) {
    try {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t threads = (size_t)sms * 2048;
        Fq* d = nullptr;
        ZK_CUDA_CHECK(cudaMalloc(&d, sizeof(Fq) * 4 * threads));
        ZK_CUDA_CHECK(cudaMemset(d, 0x11, sizeof(Fq) * 4 * threads));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&](int it) {
            unsigned g = (unsigned)(threads / 256);
            if (kind == 1) k_op_throughput<1><<<g, 256>>>(d, it);
            else if (kind == 2) k_op_throughput<2><<<g, 256>>>(d, it);
            else k_op_throughput<3><<<g, 256>>>(d, it);
        };
        run(8);
        cudaEventRecord(e0);
        run(iters);
        cudaEventRecord(e1);
        ZK_CUDA_CHECK(cudaEventSynchronize(e1));
        g_launch_count += 2;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
        const double per_iter = kind == 1 ? 4.0 : kind == 2 ? 4.0 : 6.0;
        return (double)threads * per_iter * iters / (ms * 1e-3);
    } catch (const CudaError&) {
        return -1.0;
    }
}
// returns Montgomery products per second measured with CUDA events (Fq, all SMs busy), or a negative value
double rlnb200_mul_throughput(int iters) {
    try {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t threads = (size_t)sms * 2048;
        Fq* d = nullptr;
        ZK_CUDA_CHECK(cudaMalloc(&d, sizeof(Fq) * 4 * threads));
        ZK_CUDA_CHECK(cudaMemset(d, 0x11, sizeof(Fq) * 4 * threads));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_mul_throughput<Fq><<<(unsigned)(threads / 256), 256>>>(d, 8);
        cudaEventRecord(e0);
        k_mul_throughput<Fq><<<(unsigned)(threads / 256), 256>>>(d, iters);
        cudaEventRecord(e1);
        ZK_CUDA_CHECK(cudaEventSynchronize(e1));
        g_launch_count += 2;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
        return (double)threads * 4.0 * iters / (ms * 1e-3);
    } catch (const CudaError&) {
        return -1.0;
    }
}
}
