// Issue rate of the legacy warp-level tensor-core path on sm_100a: mma.sync.m16n8k16 (f16 x f16 -> f32) from
// registers, 8 independent accumulator tiles per warp.  Question: could the LSTM recurrence's 256 x 64 mat-vec
// (4 chunks = 4 of 8 columns) run there instead of on the FFMA2 pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) hmma_loop(float *out, int iters, unsigned a0, unsigned b0) {
    unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                "{%0,%1,%2,%3};"
                : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256, 1) ffma2_loop(float *out, int iters, float w0) {
    float2 c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = make_float2(0.f, 0.f);
    float2 h = make_float2(w0, w0 + 1.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = __ffma2_rn(h, make_float2(w0, w0), c[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i].x + c[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int dev_clock_khz = 0;
    cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            hmma_loop<<<148, threads>>>(out, iters, 0x3c003c00u, 0x3c003c00u);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double warps = threads / 32.0, n = warps * iters * 8.0;  // HMMA per SM
        const double clk = ms * 1e-3 * dev_clock_khz * 1e3;
        printf("mma.sync m16n8k16: %4d threads/SM: %.2f cycles per HMMA per SM sub-partition, %.0f MAC/clk/SM (%.3f ms)\n",
               threads, clk / (n / 4.0), n * 2048.0 / clk, ms);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            ffma2_loop<<<148, threads>>>(out, iters, 1.0f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        const double clk2 = ms * 1e-3 * dev_clock_khz * 1e3;
        printf("FFMA2            : %4d threads/SM: %.2f cycles per FFMA2 per SM sub-partition, %.0f MAC/clk/SM\n", threads,
               clk2 / (n / 4.0), n * 64.0 / clk2);
    }
    printf("clock %d kHz; cudaGetLastError: %s\n", dev_clock_khz, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
