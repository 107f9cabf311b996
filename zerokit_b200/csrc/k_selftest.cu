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
    else fr = fx.inv();
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
