// TEST INFRASTRUCTURE.  Host emulation of one warp of refine_dp_kernel: compiles the PRODUCT's
// per-read algorithm (remora_b200/csrc/rb200_refine_core.cuh) for the CPU and runs it with `n_lanes`
// real threads that meet at a pthread barrier wherever the kernel executes __syncwarp().  With
// n_lanes = 1 it checks the arithmetic/branch logic; with n_lanes = 32, built with
// -fsanitize=thread, it checks that every shared-row access is ordered by a barrier (a missing
// __syncwarp shows up as a data race).  Built by tests/test_refine.py:
//   g++ -O1 -g -std=c++17 -ffp-contract=off [-fsanitize=thread] -shared -fPIC -pthread
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <thread>
#include <vector>

#include "rb200_refine_core.cuh"

namespace {

struct HostCtx {
    int lane;
    int nl;
    pthread_barrier_t *bar;
    int *xchg;  // nl ints shared by the lanes (what the warp shuffles exchange on the GPU)
    int skip_mod = 0, skip_rem = -1, count = 0;  // negative control: drop every sync with count % mod == rem
    void sync() {
        const int c = count++;
        if (skip_mod > 0 && c % skip_mod == skip_rem) return;
        if (nl > 1) pthread_barrier_wait(bar);
    }
    void hard_sync() {
        if (nl > 1) pthread_barrier_wait(bar);
    }
    int scan_max(int v) {
        if (nl == 1) return v;
        xchg[lane] = v;
        hard_sync();
        int m = v;
        for (int i = 0; i < lane; ++i) m = xchg[i] > m ? xchg[i] : m;
        hard_sync();
        return m;
    }
    int bcast_last(int v) {
        if (nl == 1) return v;
        if (lane == nl - 1) xchg[0] = v;
        hard_sync();
        const int out = xchg[0];
        hard_sync();
        return out;
    }
};

}  // namespace

extern "C" int emul_refine_read(const float *sig, const float *levels, const int32_t *st, const int32_t *en,
                                int n_bases, const float *pen, int n_pen, int algo, int max_w, int near_cap,
                                int n_lanes, int32_t *tb, int32_t *path, float *score, int32_t *status) {
    // "near" rows (shared memory on the GPU) hold bands up to near_cap samples, "far" rows (global
    // scratch) the wider ones; rows are read in 16-byte groups
    const size_t ncap = (size_t)((near_cap + 3) & ~3), fcap = (size_t)((max_w + 3) & ~3);
    // + 4 floats: the chain's look-ahead load reads one group past the last row
    std::vector<float> near_buf(rb200::refine::kRowsPerWarp * ncap + 4), far_buf(rb200::refine::kRowsPerWarp * fcap + 4);
    std::vector<int32_t> slot(4), spec(32);
    std::vector<int> xchg(32);
    const rb200::refine::Rows near = rb200::refine::carve_rows(near_buf.data(), ncap);
    const rb200::refine::Rows far = rb200::refine::carve_rows(far_buf.data(), fcap);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, n_lanes);
    auto body = [&](int lane) {
        HostCtx ctx{lane, n_lanes, &bar, xchg.data()};
        if (const char *e = getenv("EMUL_SKIP_SYNC")) {  // "mod,rem" - proves the race detector sees a missing barrier
            sscanf(e, "%d,%d", &ctx.skip_mod, &ctx.skip_rem);
        }
        rb200::refine::refine_read_warp(ctx, sig, levels, st, en, n_bases, pen, n_pen, algo, tb, path, score,
                                        status, near, near_cap, far, slot.data(), spec.data());
    };
    if (n_lanes == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        for (int l = 0; l < n_lanes; ++l) th.emplace_back(body, l);
        for (auto &t : th) t.join();
    }
    pthread_barrier_destroy(&bar);
    return 0;
}

#ifdef EMUL_MAIN
// Standalone form (needed for ThreadSanitizer, which cannot be loaded into a running python):
//   refine_emul <in.bin> <out.bin>
// in : int32 {n_bases, sig_len, n_pen, algo, max_w, n_lanes, near_cap}, float sig[sig_len], float levels[n_bases],
//      int32 st[n_bases], int32 en[n_bases], float pen[n_pen]
// out: int32 path[n_bases+1], float score, int32 status, int32 tb[band_len]
#include <cstdio>
int main(int argc, char **argv) {
    if (argc != 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t h[7];
    if (fread(h, 4, 7, f) != 7) return 2;
    const int n_bases = h[0], sig_len = h[1], n_pen = h[2], algo = h[3], max_w = h[4], n_lanes = h[5];
    const int near_cap = h[6];
    std::vector<float> sig(sig_len), levels(n_bases), pen(n_pen > 0 ? n_pen : 1);
    std::vector<int32_t> st(n_bases), en(n_bases);
    if (fread(sig.data(), 4, sig_len, f) != (size_t)sig_len) return 2;
    if (fread(levels.data(), 4, n_bases, f) != (size_t)n_bases) return 2;
    if (fread(st.data(), 4, n_bases, f) != (size_t)n_bases) return 2;
    if (fread(en.data(), 4, n_bases, f) != (size_t)n_bases) return 2;
    if (n_pen > 0 && fread(pen.data(), 4, n_pen, f) != (size_t)n_pen) return 2;
    fclose(f);
    size_t band_len = 0;
    for (int b = 0; b < n_bases; ++b) band_len += (size_t)(en[b] - st[b]);
    std::vector<int32_t> tb(band_len), path(n_bases + 1);
    float score = 0;
    int32_t status = -1;
    emul_refine_read(sig.data(), levels.data(), st.data(), en.data(), n_bases, pen.data(), n_pen, algo, max_w,
                     near_cap, n_lanes, tb.data(), path.data(), &score, &status);
    f = fopen(argv[2], "wb");
    if (!f) return 2;
    fwrite(path.data(), 4, path.size(), f);
    fwrite(&score, 4, 1, f);
    fwrite(&status, 4, 1, f);
    fwrite(tb.data(), 4, tb.size(), f);
    fclose(f);
    return 0;
}
#endif
