// Bring-up: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st) and a
// small N (16 / 32) B tile in shared memory - the shape of the LSTM recurrence (A = W_hh resident).
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
__global__ void __launch_bounds__(128, 1) k(const float *A, const float *b_sw, float *out, int N, long long *cyc) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar; __shared__ uint32_t tbase;
    float *b_s = (float *)sm;  // 2 K-blocks x [N rows][32] swizzled
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2 * N * 32; i += 128) b_s[i] = b_sw[i];
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tbase))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tbase;
    // A: row = tid (TMEM lane), 64 k values -> TMEM columns [64, 128) via two 32-column stores
    {
        const int row = warp * 32 + lane;
        for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(A[row * 64 + h * 32 + i]);
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64 + h * 32;
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
                "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
                "r"(r[30]), "r"(r[31]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        for (int rep = 0; rep < 8; ++rep)  // repeat to time the MMA rate; only the last result matters
        for (int kb = 0; kb < 2; ++kb)
            for (int k8 = 0; k8 < 4; ++k8) {
                const uint32_t a_t = tmem + 64 + kb * 32 + k8 * 8;   // A columns of this K slice
                const uint64_t bd = desc_sw128(smem_u32(b_s) + kb * N * 128 + k8 * 32);
                const uint32_t acc = (kb | k8) ? 1u : 0u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                             ::"r"(tmem), "r"(a_t), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    long long t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[row * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}
static float trunc13(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }
static inline int sw_index(int r, int k) { return r * 32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3); }
int main() {
    for (int N : {16, 32}) {
        std::vector<float> A(128 * 64), B(N * 64);
        srand(3);
        for (auto &v : A) v = rand() / (float)RAND_MAX - 0.5f;
        for (auto &v : B) v = rand() / (float)RAND_MAX - 0.5f;
        std::vector<float> bsw(2 * N * 32);
        for (int kb = 0; kb < 2; ++kb) for (int r = 0; r < N; ++r) for (int k = 0; k < 32; ++k) bsw[kb * N * 32 + sw_index(r, k)] = B[r * 64 + kb * 32 + k];
        float *dA, *dB, *dO; long long *dc;
        CHECK(cudaMalloc(&dA, A.size() * 4)); CHECK(cudaMalloc(&dB, bsw.size() * 4)); CHECK(cudaMalloc(&dO, 128 * N * 4)); CHECK(cudaMalloc(&dc, 8));
        CHECK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dB, bsw.data(), bsw.size() * 4, cudaMemcpyHostToDevice));
        k<<<1, 128, 16384>>>(dA, dB, dO, N, dc);
        CHECK(cudaDeviceSynchronize());
        std::vector<float> out(128 * N); long long cyc;
        CHECK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost)); CHECK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
        double e = 0, mr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double r = 0; for (int kk = 0; kk < 64; ++kk) r += (double)trunc13(A[m * 64 + kk]) * trunc13(B[n * 64 + kk]);
            // the kernel repeats the 8-MMA sequence 8 times, each restarting with accumulate=0 -> result of one pass
            e = fmax(e, fabs(r - out[m * N + n])); mr = fmax(mr, fabs(r));
        }
        printf("A-from-TMEM N=%d: max err %.3e (max ref %.2f) %s ; 64 MMAs took %lld cycles (%.1f per MMA)\n", N, e, mr, e < 1e-4 ? "OK" : "MISMATCH", cyc, cyc / 64.0);
    }
    return 0;
}
