// Gather bandwidth of 12 KB float32 rows (the re-rank's access pattern: a warp reads 4 rows at a time, 16 bytes per lane,
// two chunk steps in flight, L1::no_allocate) as a function of the row window the concurrently running CTAs share.
// 483 k row reads over a 50,000 x 3000 matrix per pass, as in one headline batch; the reads are dealt out window by window
// (all CTAs inside one window at a time), every row of a window is asked for ~9.7 times.  Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/l2_window_probe.bin scripts/l2_window_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(128) gather_kernel(const float *m, int ld, const int *rows, int n_items, unsigned *counter, float *out) {
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, chunks = ld >> 2;
    float acc = 0.f;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (int)atomicAdd(counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        const float4 *src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) src[u] = reinterpret_cast<const float4 *>(m + (size_t)rows[item * 16 + warp * 4 + u] * ld);
        for (int c = lane; c < chunks; c += 64) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = ldg_stream_f4(src[u] + c);
            if (c + 32 < chunks) {
#pragma unroll
                for (int u = 0; u < 4; ++u) b[u] = ldg_stream_f4(src[u] + c + 32);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += a[u].x + a[u].w + b[u].y + b[u].z;
        }
    }
    if (acc == 123.456f) out[blockIdx.x] = acc;
}

int main() {
    const int n = 50000, ld = 3000, reads = 483000 / 16 * 16;
    float *m; int *d_rows; unsigned *counter; float *out;
    cudaMalloc(&m, (size_t)n * ld * 4); cudaMemset(m, 0, (size_t)n * ld * 4);
    cudaMalloc(&d_rows, reads * 4); cudaMalloc(&counter, 4); cudaMalloc(&out, 4096 * 4);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::mt19937 rng(1);
    const int window_mb[] = {8, 16, 24, 32, 48, 64, 96, 128, 256, 600};
    for (int wmb : window_mb) {
        int wrows = (int)((size_t)wmb * 1024 * 1024 / ((size_t)ld * 4));
        if (wrows > n) wrows = n;
        const int windows = (n + wrows - 1) / wrows;
        std::vector<int> rows(reads);
        for (int i = 0; i < reads; ++i) {                        // window-major: read i belongs to window i * windows / reads
            const int w = (int)((long long)i * windows / reads);
            const int lo = w * wrows, hi = lo + wrows < n ? lo + wrows : n;
            rows[i] = lo + (int)(rng() % (unsigned)(hi - lo));
        }
        cudaMemcpy(d_rows, rows.data(), reads * 4, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemset(counter, 0, 4);
            cudaEventRecord(e0);
            gather_kernel<<<sms * 8, 128>>>(m, ld, d_rows, reads / 16, counter, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("window %4d MB (%5d rows, %3d windows): %.3f ms for %.2f GB of row reads = %.0f GB/s\n", wmb, wrows, windows, best,
               (double)reads * ld * 4 / 1e9, (double)reads * ld * 4 / 1e9 / (best * 1e-3));
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
