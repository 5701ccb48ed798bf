// Bring-up test for tcgen05.mma kind::f16 (fp16 / bf16 operands, fp32 accumulate in TMEM) with
// K-major NO-SWIZZLE ("interleaved") operands: 16-byte K chunks are the slow axis, rows the fast one,
//     byte address(row, k) = base + (k / 8) * LBO + row * 16 + (k % 8) * 2,   SBO = 128 B, LBO = rows * 16 B
// so an 8-row x 16 B core matrix is 128 contiguous bytes and a descriptor whose start address is
// advanced by i * 16 B reads the tile shifted by i rows (sliding-window convolution, no im2col).
// Checks: (1) row-shifted A descriptors, N = 64 / 128 / 256, fp16 and bf16; (2) the N = 64 view of an
// N = 128 B tile; (3) the three-product fp16 split  a*b ~= ah*bh + al*bh + ah*bl  against fp64;
// (4) issue rate of back-to-back MMAs per N.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: LBO = byte distance between the two 16-byte K chunks of one MMA (and between
// consecutive chunks in general), SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100); layout type 0 = no swizzle
    return d;
}
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N, int bf16) {
    uint32_t d = 0;
    d |= 1u << 4;                        // D = f32
    d |= (uint32_t)(bf16 ? 1 : 0) << 7;  // A format: 0 = f16, 1 = bf16
    d |= (uint32_t)(bf16 ? 1 : 0) << 10; // B format
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mbar_wait0(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

// mode 0: D[128 x N] = A(shifted) . B^T over nkc/2 MMAs of K = 16
// mode 1: split test: A tile = [hi tile | lo tile], B tile N = 128 = [hi rows 0..63 ; lo rows 64..127]:
//         cols [0,128) <- ah * [bh ; bl] (N = 128), cols [64,128) += al * bh (N = 64)
// mode 2: timing, `reps` back-to-back MMAs of the same operands
__global__ void __launch_bounds__(128, 1)
umma_f16_test(const uint16_t *__restrict__ a_g, const uint16_t *__restrict__ b_g, float *__restrict__ out,
              int a_rows, int N, int nkc, int shift, int bf16, int swap_ls, int mode, int reps,
              long long *__restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int a_tile = nkc * a_rows * 16;       // bytes of one A tile
    const int n_a_tiles = mode == 1 ? 2 : 1;
    uint8_t *a_s = smem_raw;
    uint8_t *b_s = smem_raw + ((n_a_tiles * a_tile + 127) & ~127);
    const int b_bytes = nkc * N * 16;
    for (int i = tid; i < n_a_tiles * a_tile / 4; i += 128)
        reinterpret_cast<uint32_t *>(a_s)[i] = reinterpret_cast<const uint32_t *>(a_g)[i];
    for (int i = tid; i < b_bytes / 4; i += 128)
        reinterpret_cast<uint32_t *>(b_s)[i] = reinterpret_cast<const uint32_t *>(b_g)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t a_lbo = a_rows * 16, b_lbo = N * 16, sbo = 128;
        auto adesc = [&](uint32_t addr) { return swap_ls ? make_desc_ns(addr, sbo, a_lbo) : make_desc_ns(addr, a_lbo, sbo); };
        auto bdesc = [&](uint32_t addr) { return swap_ls ? make_desc_ns(addr, sbo, b_lbo) : make_desc_ns(addr, b_lbo, sbo); };
        const uint32_t a0 = smem_u32(a_s) + shift * 16, b0 = smem_u32(b_s);
        long long t0 = 0;
        if (mode == 0) {
            const uint32_t idesc = make_idesc_f16(128, N, bf16);
            for (int k = 0; k < nkc / 2; ++k)
                mma_f16(tmem, adesc(a0 + 2 * k * a_lbo), bdesc(b0 + 2 * k * b_lbo), idesc, k ? 1u : 0u);
        } else if (mode == 1) {
            const uint32_t id128 = make_idesc_f16(128, 128, bf16), id64 = make_idesc_f16(128, 64, bf16);
            for (int k = 0; k < nkc / 2; ++k) {
                mma_f16(tmem, adesc(a0 + 2 * k * a_lbo), bdesc(b0 + 2 * k * b_lbo), id128, k ? 1u : 0u);
                mma_f16(tmem + 64, adesc(a0 + a_tile + 2 * k * a_lbo), bdesc(b0 + 2 * k * b_lbo), id64, 1u);
            }
        } else {
            const uint32_t idesc = make_idesc_f16(128, N, bf16);
            t0 = clock64();
            for (int r = 0; r < reps; ++r)
                mma_f16(tmem, adesc(a0 + 2 * (r % (nkc / 2)) * a_lbo), bdesc(b0 + 2 * (r % (nkc / 2)) * b_lbo), idesc, r ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
        if (mode == 2) {
            const long long t1 = clock64();
            mbar_wait0(&bar_mma, 0);
            const long long t2 = clock64();
            cycles[0] = t1 - t0;
            cycles[1] = t2 - t0;
        }
    }
    mbar_wait0(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + (tid & 31);
    const int ncols = mode == 1 ? 128 : N;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[row * 256 + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

static uint16_t to16(float x, int bf16) {
    if (bf16) { __nv_bfloat16 h = __float2bfloat16_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
    __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u;
}
static float from16(uint16_t u, int bf16) {
    if (bf16) { __nv_bfloat16 h; memcpy(&h, &u, 2); return __bfloat162float(h); }
    __half h; memcpy(&h, &u, 2); return __half2float(h);
}
// element (row, k) of a [kc][rows][8] tile
static inline size_t ns_index(int rows, int r, int k) { return ((size_t)(k >> 3) * rows + r) * 8 + (k & 7); }

int main() {
    const int a_rows = 136, K = 64, nkc = K / 8;
    float *dout; long long *dcyc;
    CHECK(cudaMalloc(&dout, 128 * 256 * 4)); CHECK(cudaMalloc(&dcyc, 16));
    CHECK(cudaFuncSetAttribute(umma_f16_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    srand(3);
    std::vector<float> A(a_rows * K), Bm(256 * K);
    for (auto &v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto &v : Bm) v = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
    uint16_t *da, *db;
    CHECK(cudaMalloc(&da, 2 * a_rows * K * 2)); CHECK(cudaMalloc(&db, 256 * K * 2));
    // ---- (1) shifted descriptors, N sweep, both formats, both LBO/SBO assignments ----
    for (int bf16 : {0, 1}) for (int N : {64, 128, 256}) for (int swap_ls : {0}) for (int shift : {0, 1, 3, 4}) {
        std::vector<uint16_t> a_t(a_rows * K), b_t(N * K);
        for (int r = 0; r < a_rows; ++r) for (int k = 0; k < K; ++k) a_t[ns_index(a_rows, r, k)] = to16(A[r * K + k], bf16);
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) b_t[ns_index(N, n, k)] = to16(Bm[n * K + k], bf16);
        CHECK(cudaMemcpy(da, a_t.data(), a_t.size() * 2, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(db, b_t.data(), b_t.size() * 2, cudaMemcpyHostToDevice));
        CHECK(cudaMemset(dout, 0, 128 * 256 * 4));
        umma_f16_test<<<1, 128, 160 * 1024>>>(da, db, dout, a_rows, N, nkc, shift, bf16, swap_ls, 0, 0, dcyc);
        CHECK(cudaDeviceSynchronize());
        std::vector<float> out(128 * 256);
        CHECK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        double max_err = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k)
                ref += (double)from16(to16(A[(m + shift) * K + k], bf16), bf16) * (double)from16(to16(Bm[n * K + k], bf16), bf16);
            max_err = fmax(max_err, fabs(ref - out[m * 256 + n]));
        }
        printf("%s N=%3d swapLS=%d shift=%d  max|err|=%.3e  %s\n", bf16 ? "bf16" : "fp16", N, swap_ls, shift, max_err,
               max_err < 1e-5 ? "OK" : "MISMATCH");
    }
    // ---- (2)+(3) three-product fp16 split against fp64, N = 128 B tile [hi ; lo], N = 64 view ----
    for (int swap_ls : {0}) for (float wscale : {1.f, 4096.f}) {
        // weights (B) scaled by a power of two so that their lo parts stay normal in fp16; activations (A) as is
        std::vector<uint16_t> a_t(2 * a_rows * K), b_t(128 * K);
        for (int r = 0; r < a_rows; ++r) for (int k = 0; k < K; ++k) {
            const float a = A[r * K + k];
            const uint16_t hi = to16(a, 0);
            a_t[ns_index(a_rows, r, k)] = hi;
            a_t[(size_t)a_rows * K + ns_index(a_rows, r, k)] = to16(a - from16(hi, 0), 0);
        }
        for (int n = 0; n < 64; ++n) for (int k = 0; k < K; ++k) {
            const float b = Bm[n * K + k] * wscale;
            const uint16_t hi = to16(b, 0);
            b_t[ns_index(128, n, k)] = hi;
            b_t[ns_index(128, 64 + n, k)] = to16(b - from16(hi, 0), 0);
        }
        CHECK(cudaMemcpy(da, a_t.data(), a_t.size() * 2, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(db, b_t.data(), b_t.size() * 2, cudaMemcpyHostToDevice));
        CHECK(cudaMemset(dout, 0, 128 * 256 * 4));
        umma_f16_test<<<1, 128, 160 * 1024>>>(da, db, dout, a_rows, 128, nkc, 2, 0, swap_ls, 1, 0, dcyc);
        CHECK(cudaDeviceSynchronize());
        std::vector<float> out(128 * 256);
        CHECK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        double max_rel = 0, max_rel_fp32 = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
            double ref = 0, mag = 0; float f32 = 0.f;
            for (int k = 0; k < K; ++k) {
                ref += (double)A[(m + 2) * K + k] * (double)Bm[n * K + k];
                mag += fabs((double)A[(m + 2) * K + k] * (double)Bm[n * K + k]);
                f32 = fmaf(A[(m + 2) * K + k], Bm[n * K + k], f32);
            }
            const double got = ((double)out[m * 256 + n] + (double)out[m * 256 + 64 + n]) / wscale;
            max_rel = fmax(max_rel, fabs(ref - got) / mag);
            max_rel_fp32 = fmax(max_rel_fp32, fabs(ref - f32) / mag);
        }
        printf("3xFP16 split swapLS=%d wscale=%g: max err / sum|terms| = %.3e   (sequential fp32 FMA: %.3e)\n", swap_ls,
               wscale, max_rel, max_rel_fp32);
    }
    // ---- (4) issue rate ----
    for (int N : {8, 16, 64, 128, 256}) {
        std::vector<uint16_t> z((size_t)256 * K, 0);
        CHECK(cudaMemcpy(db, z.data(), z.size() * 2, cudaMemcpyHostToDevice));
        const int reps = 256;
        umma_f16_test<<<1, 128, 160 * 1024>>>(da, db, dout, a_rows, N < 64 ? 64 : N, nkc, 0, 0, 0, 2, reps, dcyc);
        CHECK(cudaDeviceSynchronize());
        if (N < 64) {  // small-N timing needs its own idesc: rerun through mode 2 with N as given
            umma_f16_test<<<1, 128, 160 * 1024>>>(da, db, dout, a_rows, N, nkc, 0, 0, 0, 2, reps, dcyc);
            CHECK(cudaDeviceSynchronize());
        }
        long long c[2];
        CHECK(cudaMemcpy(c, dcyc, 16, cudaMemcpyDeviceToHost));
        printf("timing M128 N=%3d K16 fp16: issue %.1f clk/MMA, complete %.1f clk/MMA over %d MMAs\n", N,
               c[0] / (double)reps, c[1] / (double)reps, reps);
    }
    return 0;
}
