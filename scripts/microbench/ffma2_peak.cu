// Micro-benchmark: peak issue rate of FFMA / FFMA2 forms on sm_100a (register-only loops).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float *out, int iters, float seed) {
    float2 acc[16];
    float2 a[4];
    float s[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = make_float2(seed + i, seed - i);
        s[i] = seed * (i + 1);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) {  // FFMA2, pair * broadcast scalar (scalar shared by 4 consecutive)
                    acc[i] = __ffma2_rn(a[i & 3], make_float2(s[r], s[r]), acc[i]);
                } else if (MODE == 1) {  // FFMA2, pair * pair
                    acc[i] = __ffma2_rn(a[i & 3], a[(i + r) & 3], acc[i]);
                } else if (MODE == 2) {  // scalar FFMA x2
                    acc[i].x = fmaf(a[i & 3].x, s[r], acc[i].x);
                    acc[i].y = fmaf(a[i & 3].y, s[r], acc[i].y);
                } else if (MODE == 3) {  // FFMA2, weight pair reused across 16, scalar varies
                    acc[i] = __ffma2_rn(a[r], make_float2(s[i & 3], s[i & 3]), acc[i]);
                }
            }
        }
    }
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
void run(const char *name, int threads, int ctas_per_sm) {
    int sms = 148;
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * threads);
    const int iters = 4096;
    k<MODE><<<sms * ctas_per_sm, threads>>>(out, 16, 1.0001f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * ctas_per_sm, threads>>>(out, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * ctas_per_sm * threads * iters * 64.0 * 2.0;  // FMAs (2 per FFMA2)
    const double per_clk_sm = fma / (ms * 1e-3) / 1.965e9 / sms;
    printf("%-34s thr=%4d x%d  %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM (peak 128)\n", name, threads,
           ctas_per_sm, ms, 2 * fma / (ms * 1e-3) / 1e12, per_clk_sm);
    cudaFree(out);
}

int main() {
    for (int thr : {256, 512, 1024}) {
        run<0>("FFMA2 pair*scalar (scalar reused x16)", thr, 1);
        run<3>("FFMA2 pair(reused x16)*scalar", thr, 1);
        run<1>("FFMA2 pair*pair", thr, 1);
        run<2>("FFMA scalar", thr, 1);
    }
    return 0;
}
