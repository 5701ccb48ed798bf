// How fast does one CTA's tcgen05.mma stream run for a given operand layout?  One thread issues `reps`
// M128 x N x K16 fp16 MMAs (operands are zeros: timing only), cycling over the K slices of a 64-wide tile,
// into 1 / 2 / 4 independent accumulators.  Layouts: K-major without swizzle ([16-byte K chunk][row], what
// rb200_mega.cu uses), and K-major SWIZZLE_32B / 64B / 128B rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// layout: 0 none (LBO = rows * 16, SBO = 128), 6 = 32B, 4 = 64B, 2 = 128B swizzle (row pitch = swizzle width)
__global__ void __launch_bounds__(128, 1) rate(int N, int layout, int nacc, int reps, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sm), b0 = a0 + 32 * 1024;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int width = layout == 2 ? 128 : layout == 4 ? 64 : layout == 6 ? 32 : 0;  // bytes per swizzled row
        const int kslices = width ? width / 32 : 4;                                      // K16 slices per tile
        // descriptors and accumulator addresses are precomputed: the issuing thread's loop is nothing but MMAs
        // (the first version of this benchmark computed them per MMA and measured ITS OWN 223-cycle loop body)
        uint64_t da[4], db[4];
        uint32_t dacc[4];
        for (int ks = 0; ks < 4; ++ks) {
            const int k = ks % kslices;
            if (width == 0) {
                da[ks] = make_desc(a0 + k * 2 * 136 * 16, 136 * 16, 128, 0);
                db[ks] = make_desc(b0 + k * 2 * N * 16, N * 16, 128, 0);
            } else {
                da[ks] = make_desc(a0 + k * 32, 16, 8 * width, layout);
                db[ks] = make_desc(b0 + k * 32, 16, 8 * width, layout);
            }
            dacc[ks] = tmem + (ks % nacc) * (512 / nacc);
        }
        const long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; r += 4) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_f16(dacc[ks], da[ks], db[ks], idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
        const long long t1 = clock64();
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(smem_u32(&bar)), "r"(0)
                         : "memory");
        } while (!done);
        cycles[0] = t1 - t0;
        cycles[1] = clock64() - t0;
    }
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    long long *dc, hc[2];
    cudaMalloc(&dc, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 512;
    const char *names[7] = {"no swizzle", "", "SWIZZLE_128B", "", "SWIZZLE_64B", "", "SWIZZLE_32B"};
    for (int layout : {0, 6, 4, 2})
        for (int N : {64, 128, 256})
            for (int nacc : {1, 2, 4}) {
                if (N * nacc > 512) continue;
                rate<<<1, 128, 100 * 1024>>>(N, layout, nacc, reps, dc);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
                printf("%-13s N=%3d accumulators=%d: issue %.1f, complete %.1f cycles per MMA (ideal %d)%s\n", names[layout], N,
                       nacc, hc[0] / (double)reps, hc[1] / (double)reps, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
