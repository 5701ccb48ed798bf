// Accuracy of 3xTF32 accumulation chains on tcgen05 (K = 640 like merge_conv1), vs fp64 reference.
//   variant 0: one accumulator, hi*hi + lo*hi + hi*lo interleaved (240 accumulate steps)
//   variant 1: main accumulator (hi*hi) + separate accumulator for the two correction products
//   variant 2: as 1, main accumulation split into 4 chains (one per 160-wide K quarter), summed in fp32
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
constexpr int NKB = 20;  // 20 x 32 = 640
__global__ void __launch_bounds__(128, 1) acc_test(const float *a_hi, const float *a_lo, const float *b_hi, const float *b_lo, float *out, int variant) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar; __shared__ uint32_t tbase;
    // smem: per kb stage: A_hi 16K, A_lo 16K, B_hi 8K, B_lo 8K (single stage, reloaded per kb with syncs)
    float *sa_hi = (float *)sm, *sa_lo = (float *)(sm + 16384), *sb_hi = (float *)(sm + 32768), *sb_lo = (float *)(sm + 40960);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tbase;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;
    for (int kb = 0; kb < NKB; ++kb) {
        for (int i = tid; i < 128 * 32; i += 128) { sa_hi[i] = a_hi[kb * 4096 + i]; sa_lo[i] = a_lo[kb * 4096 + i]; }
        for (int i = tid; i < 64 * 32; i += 128) { sb_hi[i] = b_hi[kb * 2048 + i]; sb_lo[i] = b_lo[kb * 2048 + i]; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t main_col = variant == 2 ? (kb / 5) * 64 : 0;  // 4 chains of 5 k-blocks
            const uint32_t small_col = 256;
            for (int k8 = 0; k8 < 4; ++k8) {
                const uint32_t o = k8 * 32;
                const bool first_main = (variant == 2 ? (kb % 5 == 0) : (kb == 0)) && k8 == 0;
                mma(tmem + main_col, desc_sw128(smem_u32(sa_hi) + o), desc_sw128(smem_u32(sb_hi) + o), idesc, first_main ? 0 : 1);
                const uint32_t sc = variant == 0 ? main_col : small_col;
                const bool first_small = variant != 0 && kb == 0 && k8 == 0;
                mma(tmem + sc, desc_sw128(smem_u32(sa_lo) + o), desc_sw128(smem_u32(sb_hi) + o), idesc, first_small ? 0 : 1);
                mma(tmem + sc, desc_sw128(smem_u32(sa_hi) + o), desc_sw128(smem_u32(sb_lo) + o), idesc, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        uint32_t done = 0;
        while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float tot[16];
        for (int i = 0; i < 16; ++i) tot[i] = 0.f;
        const int nacc = variant == 2 ? 5 : (variant == 1 ? 2 : 1);
        for (int a = 0; a < nacc; ++a) {
            const uint32_t col = variant == 2 ? (a < 4 ? a * 64 : 256) : (a == 0 ? 0 : 256);
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(tmem + ((uint32_t)(warp * 32) << 16) + col + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 16; ++i) tot[i] += __uint_as_float(v[i]);
        }
        for (int i = 0; i < 16; ++i) out[row * 64 + c0 + i] = tot[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
static inline int sw_index(int r, int k) { return r * 32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3); }
static void split(float a, float &hi, float &lo) { uint32_t u; memcpy(&u, &a, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&hi, &u, 4); lo = a - hi; }
int main() {
    const int K = 32 * NKB;
    std::vector<float> A(128 * K), B(64 * K);
    srand(7);
    for (auto &v : A) v = (rand() / (float)RAND_MAX) * 2.f - 0.3f;   // not zero-mean: worst case for truncation bias
    for (auto &v : B) v = (rand() / (float)RAND_MAX) * 0.2f - 0.05f;
    std::vector<float> ah(NKB * 4096), al(NKB * 4096), bh(NKB * 2048), bl(NKB * 2048);
    for (int kb = 0; kb < NKB; ++kb) {
        for (int r = 0; r < 128; ++r) for (int k = 0; k < 32; ++k) split(A[r * K + kb * 32 + k], ah[kb * 4096 + sw_index(r, k)], al[kb * 4096 + sw_index(r, k)]);
        for (int r = 0; r < 64; ++r) for (int k = 0; k < 32; ++k) split(B[r * K + kb * 32 + k], bh[kb * 2048 + sw_index(r, k)], bl[kb * 2048 + sw_index(r, k)]);
    }
    float *dah, *dal, *dbh, *dbl, *dout;
    CHECK(cudaMalloc(&dah, ah.size() * 4)); CHECK(cudaMalloc(&dal, al.size() * 4)); CHECK(cudaMalloc(&dbh, bh.size() * 4)); CHECK(cudaMalloc(&dbl, bl.size() * 4)); CHECK(cudaMalloc(&dout, 128 * 64 * 4));
    CHECK(cudaMemcpy(dah, ah.data(), ah.size() * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dal, al.data(), al.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dbh, bh.data(), bh.size() * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dbl, bl.data(), bl.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaFuncSetAttribute(acc_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024));
    // fp32 sequential reference error for comparison
    double e32 = 0, mref = 0;
    std::vector<double> ref(128 * 64);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        double r = 0; float f = 0.f;
        for (int k = 0; k < K; ++k) { r += (double)A[m * K + k] * B[n * K + k]; f = fmaf(A[m * K + k], B[n * K + k], f); }
        ref[m * 64 + n] = r; e32 = fmax(e32, fabs(r - f)); mref = fmax(mref, fabs(r));
    }
    printf("max|ref| = %.3f ; sequential fp32 FMA max err = %.3e (rel %.2e)\n", mref, e32, e32 / mref);
    for (int variant = 0; variant < 3; ++variant) {
        acc_test<<<1, 128, 50 * 1024>>>(dah, dal, dbh, dbl, dout, variant);
        CHECK(cudaDeviceSynchronize());
        std::vector<float> out(128 * 64);
        CHECK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        double e = 0, bias = 0;
        for (int i = 0; i < 128 * 64; ++i) { e = fmax(e, fabs(ref[i] - out[i])); bias += out[i] - ref[i]; }
        printf("variant %d: max err = %.3e (rel %.2e), mean signed err = %.3e\n", variant, e, e / mref, bias / (128 * 64));
    }
    return 0;
}
