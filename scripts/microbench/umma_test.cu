// Bring-up test for tcgen05.mma kind::tf32 with K-major SWIZZLE_128B operands in shared memory,
// including a row-shifted A descriptor (sliding-window convolution) and the 3xTF32 split.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor, K-major, 128B swizzle (8-row x 128B atoms, SBO = 1024 B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);           // start address, bits 0-13
    d |= (uint64_t)(1) << 16;                         // LBO (ignored for swizzled K-major), bits 16-29
    d |= (uint64_t)(1024 >> 4) << 32;                 // SBO = 1024 B, bits 32-45
    d |= (uint64_t)1 << 46;                           // descriptor version 1 (sm_100)
    d |= (uint64_t)(base_offset & 7) << 49;           // matrix base offset, bits 49-51
    d |= (uint64_t)2 << 61;                           // layout type: SWIZZLE_128B
    return d;
}

// instruction descriptor: D=f32, A=B=tf32, both K-major, M x N
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                // c_format = F32
    d |= 2u << 7;                // a_format = TF32
    d |= 2u << 10;               // b_format = TF32
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__global__ void __launch_bounds__(128, 1)
umma_test(const float *__restrict__ a_sw, const float *__restrict__ b_sw, float *__restrict__ out,
          int a_rows, int shift, int base_off_mode, int n_kblocks) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    // smem: A k-blocks [n_kblocks][a_rows][128B], then B k-blocks [n_kblocks][64][128B]
    float *a_s = reinterpret_cast<float *>(smem_raw);
    const int a_kb_bytes = ((a_rows * 128 + 1023) / 1024) * 1024;
    float *b_s = reinterpret_cast<float *>(smem_raw + (size_t)n_kblocks * a_kb_bytes);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int kb = 0; kb < n_kblocks; ++kb) {
        for (int i = tid; i < a_rows * 32; i += 128) a_s[kb * (a_kb_bytes / 4) + i] = a_sw[kb * a_rows * 32 + i];
        for (int i = tid; i < 64 * 32; i += 128) b_s[kb * 64 * 32 + i] = b_sw[kb * 64 * 32 + i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, 64);
        bool first = true;
        for (int kb = 0; kb < n_kblocks; ++kb) {
            const uint32_t a_addr = smem_u32(a_s) + kb * a_kb_bytes + shift * 128;
            const uint32_t b_addr = smem_u32(b_s) + kb * 64 * 128;
            const uint32_t boff = base_off_mode ? ((a_addr >> 7) & 7) : 0;
            for (int k8 = 0; k8 < 4; ++k8) {
                const uint64_t ad = make_desc_sw128(a_addr + k8 * 32, boff);
                const uint64_t bd = make_desc_sw128(b_addr + k8 * 32, 0);
                const uint32_t acc = first ? 0u : 1u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                    "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                    : "memory");
                first = false;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&bar_mma)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes 32w..32w+31 (= D rows), 64 columns
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[row * 64 + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

// host-side swizzle: element (row r, k) of a [rows][32] K-block -> float index in the 128B-swizzled tile
static inline int sw_index(int r, int k) { const int chunk = (k >> 2) ^ (r & 7); return r * 32 + chunk * 4 + (k & 3); }

int main() {
    const int a_rows = 136, nkb = 2, K = 32 * nkb;
    std::vector<float> A(a_rows * K), Bm(64 * K);
    srand(1);
    for (auto &v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto &v : Bm) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    std::vector<float> a_sw(nkb * a_rows * 32), b_sw(nkb * 64 * 32);
    for (int kb = 0; kb < nkb; ++kb) {
        for (int r = 0; r < a_rows; ++r) for (int k = 0; k < 32; ++k) a_sw[kb * a_rows * 32 + sw_index(r, k)] = A[r * K + kb * 32 + k];
        for (int r = 0; r < 64; ++r) for (int k = 0; k < 32; ++k) b_sw[kb * 64 * 32 + sw_index(r, k)] = Bm[r * K + kb * 32 + k];
    }
    float *da, *db, *dout;
    CHECK(cudaMalloc(&da, a_sw.size() * 4)); CHECK(cudaMalloc(&db, b_sw.size() * 4)); CHECK(cudaMalloc(&dout, 128 * 64 * 4));
    CHECK(cudaMemcpy(da, a_sw.data(), a_sw.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(db, b_sw.data(), b_sw.size() * 4, cudaMemcpyHostToDevice));
    const int a_kb_bytes = ((a_rows * 128 + 1023) / 1024) * 1024;
    const size_t smem = (size_t)nkb * a_kb_bytes + (size_t)nkb * 64 * 128 + 1024;
    CHECK(cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int shift : {0, 3, 5}) for (int mode : {1, 0}) {
        CHECK(cudaMemset(dout, 0, 128 * 64 * 4));
        umma_test<<<1, 128, smem>>>(da, db, dout, a_rows, shift, mode, nkb);
        CHECK(cudaDeviceSynchronize());
        std::vector<float> out(128 * 64);
        CHECK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        double max_err = 0, max_ref = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)tf32_trunc(A[(m + shift) * K + k]) * (double)tf32_trunc(Bm[n * K + k]);
            max_err = fmax(max_err, fabs(ref - out[m * 64 + n]));
            max_ref = fmax(max_ref, fabs(ref));
        }
        printf("shift=%d base_offset_mode=%d  max|err|=%.3e (max|ref|=%.2f)  %s\n", shift, mode, max_err, max_ref, max_err < 1e-4 ? "OK" : "MISMATCH");
    }
    return 0;
}
