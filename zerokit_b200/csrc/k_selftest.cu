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
}  // namespace zk

using namespace zk;
extern "C" {
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
