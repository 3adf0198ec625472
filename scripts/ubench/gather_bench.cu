// Micro-benchmark: how fast can a B200 gather 4-byte words x[idx[i]]?  (the operation that bounds SpMV on power-law graphs)
// idx: uniform random or R-MAT-like (product of per-bit Bernoulli(0.24)) over n = 2^22 entries; variants: default loads,
// ld.global.cg (L1 bypass), unroll depth, table size, shared-memory table.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

template <int UNROLL, int MODE>
__global__ void __launch_bounds__(256) gather_kernel(const int *__restrict__ idx, const float *__restrict__ x, float *__restrict__ out, long n) {
    long i = (long)blockIdx.x * blockDim.x * UNROLL + threadIdx.x;
    float acc = 0.f;
    const long stride = (long)gridDim.x * blockDim.x * UNROLL;
    for (; i + (UNROLL - 1) * 256 < n; i += stride) {
        int c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) c[u] = idx[i + u * 256];
        float v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (MODE == 0) v[u] = x[c[u]];
            else if (MODE == 1) v[u] = __ldcg(x + c[u]);
            else v[u] = __ldg(x + c[u]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc += v[u];
    }
    if (acc == 123.456f) out[0] = acc;
}

// idx only (no gather): the streaming floor
__global__ void __launch_bounds__(256) stream_kernel(const int4 *__restrict__ idx, float *__restrict__ out, long n4) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int acc = 0;
    for (; i < n4; i += (long)gridDim.x * blockDim.x) { int4 v = idx[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 123456) out[0] = acc;
}

// shared-memory table of TABLE floats, idx reduced mod TABLE
template <int UNROLL>
__global__ void __launch_bounds__(1024) smem_gather_kernel(const int *__restrict__ idx, const float *__restrict__ x, float *__restrict__ out, long n, int table) {
    extern __shared__ float s[];
    for (int t = threadIdx.x; t < table; t += blockDim.x) s[t] = x[t];
    __syncthreads();
    long i = (long)blockIdx.x * blockDim.x * UNROLL + threadIdx.x;
    float acc = 0.f;
    const long stride = (long)gridDim.x * blockDim.x * UNROLL;
    for (; i + (UNROLL - 1) * 1024 < n; i += stride) {
        int c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) c[u] = idx[i + u * 1024];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc += s[(unsigned)c[u] % (unsigned)table];
    }
    if (acc == 123.456f) out[0] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    const long n = 64L << 20;   // 64 Mi gathers (like nnz of R-MAT scale 22)
    const int scale = 22, N = 1 << scale;
    std::vector<int> h_uni(n), h_rmat(n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (long i = 0; i < n; i++) {
        h_uni[i] = (int)(rnd() & (N - 1));
        uint64_t r = rnd(), r2 = rnd();
        int c = 0;
        for (int b = 0; b < scale; b++) {   // bit = 1 with probability ~0.24 (61/256)
            unsigned byte = (b < 8 ? (r >> (8 * b)) : (b < 16 ? (r2 >> (8 * (b - 8))) : (rnd() >> 8))) & 0xff;
            c |= (byte < 61) << b;
        }
        h_rmat[i] = c;
    }
    int *d_idx; float *d_x, *d_out;
    cudaMalloc(&d_idx, n * 4); cudaMalloc(&d_x, (size_t)N * 4 * 16); cudaMalloc(&d_out, 4);
    cudaMemset(d_x, 0, (size_t)N * 4 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 148;
    for (int dist = 0; dist < 2; dist++) {
        cudaMemcpy(d_idx, dist ? h_rmat.data() : h_uni.data(), n * 4, cudaMemcpyHostToDevice);
        const char *dn = dist ? "rmat " : "unif ";
        // stream floor
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); stream_kernel<<<sms * 8, 256>>>((const int4 *)d_idx, d_out, n / 4); cudaEventRecord(e1); cudaEventSynchronize(e1); }
        printf("%s idx stream only: %.1f us (%.0f GB/s)\n", dn, time_ms(e0, e1) * 1e3, n * 4 / time_ms(e0, e1) / 1e6);
#define RUN(U, M, name, blocks)                                                                         \
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); gather_kernel<U, M><<<blocks, 256>>>(d_idx, d_x, d_out, n); cudaEventRecord(e1); cudaEventSynchronize(e1); } \
        printf("%s %-28s unroll %d blocks/SM %2d: %.1f us  (%.1f G gathers/s)\n", dn, name, U, (blocks) / sms, time_ms(e0, e1) * 1e3, n / time_ms(e0, e1) / 1e6);
        RUN(1, 0, "ld.global (L1)", sms * 8)
        RUN(4, 0, "ld.global (L1)", sms * 8)
        RUN(8, 0, "ld.global (L1)", sms * 8)
        RUN(8, 0, "ld.global (L1)", sms * 4)
        RUN(16, 0, "ld.global (L1)", sms * 4)
        RUN(8, 1, "ld.global.cg (L2 only)", sms * 8)
        RUN(16, 1, "ld.global.cg (L2 only)", sms * 4)
        RUN(8, 2, "ld.global.nc", sms * 8)
        for (int table_kb : {32, 64, 128, 192}) {
            int table = table_kb * 256;
            cudaFuncSetAttribute(smem_gather_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, table_kb * 1024);
            for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); smem_gather_kernel<8><<<sms, 1024, table_kb * 1024>>>(d_idx, d_x, d_out, n, table); cudaEventRecord(e1); cudaEventSynchronize(e1); }
            printf("%s smem table %3d KB (idx mod table), 1024 thr/SM: %.1f us (%.1f G gathers/s)\n", dn, table_kb, time_ms(e0, e1) * 1e3, n / time_ms(e0, e1) / 1e6);
        }
    }
    // table-size sweep, uniform: where does the gather rate fall off (L1 / L2 / HBM)?
    cudaMemcpy(d_idx, h_uni.data(), n * 4, cudaMemcpyHostToDevice);
    for (int shift : {8, 6, 4, 2, 0}) {
        // restrict indices to N >> shift entries by masking on the fly is not possible in-kernel: regenerate
        std::vector<int> h(n);
        for (long i = 0; i < n; i++) h[i] = h_uni[i] & ((N >> shift) - 1);
        cudaMemcpy(d_idx, h.data(), n * 4, cudaMemcpyHostToDevice);
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); gather_kernel<8, 0><<<sms * 8, 256>>>(d_idx, d_x, d_out, n); cudaEventRecord(e1); cudaEventSynchronize(e1); }
        printf("unif table %7d KB: %.1f us (%.1f G gathers/s)\n", (N >> shift) * 4 / 1024, time_ms(e0, e1) * 1e3, n / time_ms(e0, e1) / 1e6);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
