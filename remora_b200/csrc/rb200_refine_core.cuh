// Warp-cooperative banded dynamic programme of the signal-mapping refinement ("next" row 4,
// SURVEY.md 8f).  Restates, for ONE read handled by ONE warp,
//   banded_forward_dp / banded_forward_vit_step / banded_forward_dwell_penalty_step / banded_traceback
//   (reference src/remora/refine_signal_map_core.pyx:118-400)
// with bit-identical scores, traceback and path.
//
// Why a warp per read and not more parallelism: every cell adds a squared error to the minimum of two
// earlier cells and the reference rounds after every add, so the stay chain of a base
// (x[p] = min(move[p], x[p-1]) + bs[p]) cannot be re-associated into a scan without changing bits, and
// base b needs the LAST score of base b-1 (its "invalid" marker, LARGE_SCORE + prev[-1]) before its
// first cell.  What is parallel inside a base, and is spread over the 32 lanes here:
//   phase 0  squared errors bs[p] = (level - signal[p])^2 and move candidates
//            prev[p-1+d] + bs[p] of the band                                      (parallel)
//   phase 1  the un-penalised Viterbi row: one serial chain (FADD + FMNMX per cell), run redundantly
//            by all lanes out of shared memory (broadcast 16-byte reads, next group prefetched), each
//            group of four cells stored by one owner lane                         (serial)
//   phase 2  the short-dwell-penalty row: per cell <= D+1 candidates built from the previous row, the
//            un-penalised row and bs                                              (parallel)
//   phase 3  cells more than D samples past the previous band: pure stay chain    (serial, short)
//   traceback entries go to global memory coalesced; the final traceback is one serial walk.
// Every float operation uses the explicitly rounded add/sub/mul (no fused multiply-add), in the
// reference's order.
//
// The same source compiles for the host (tests/native/refine_emul.cpp: 32 threads + a barrier, run
// under ThreadSanitizer) so that the lane/synchronisation logic is checked on CPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RB_HD __device__ __forceinline__
#define RB_FADD(a, b) __fadd_rn((a), (b))
#define RB_FSUB(a, b) __fsub_rn((a), (b))
#define RB_FMUL(a, b) __fmul_rn((a), (b))
#define RB_INF __int_as_float(0x7f800000)
#else
#include <cmath>
#define RB_HD static inline
// host emulation is compiled with -ffp-contract=off
#define RB_FADD(a, b) ((a) + (b))
#define RB_FSUB(a, b) ((a) - (b))
#define RB_FMUL(a, b) ((a) * (b))
#define RB_INF HUGE_VALF
#endif

namespace rb200 {
namespace refine {

struct alignas(16) Vec4f {
    float a, b, c, d;
};
struct alignas(16) Vec4i {
    int32_t a, b, c, d;
};

constexpr float kLargeScore = 100.0f;  // LARGE_SCORE, refine_signal_map_core.pyx:13
constexpr int kMaxPen = 16;            // longest short-dwell penalty array supported
enum { kAlgoViterbi = 0, kAlgoDwellPenalty = 1 };
enum { kStatusOk = 0, kStatusTracebackLeftBand = 1 };

// Ctx supplies: int lane; int nl (lanes); void sync() (warp barrier with memory ordering).
//
// Per-warp scratch (shared memory, or global for reads with very wide bands), each 16-byte aligned and
// holding >= the widest band rounded up to a multiple of 4:
//   row_a, row_b  penalised scores of the previous / current base (swapped every base)
//   unp, utb      un-penalised Viterbi scores and traceback of the current base
//   bs, mvs       squared errors and move candidates of the current base's band
//   slot          one int (traceback count at the last cell before the stay-only region)
//
// Preconditions (validated by the caller, they are what adjust_seq_band/validate_band establish):
//   st[0] == 0, st and en strictly increasing, en[b] > st[b], st[b] <= en[b-1], n_pen in [1, kMaxPen].
template <class Ctx>
RB_HD void refine_read_warp(Ctx &ctx, const float *__restrict__ sig, const float *__restrict__ levels,
                            const int32_t *__restrict__ st, const int32_t *__restrict__ en, int n_bases,
                            const float *pen, int n_pen, int algo, int32_t *__restrict__ tb,
                            int32_t *__restrict__ path, float *final_score, int32_t *status, float *row_a,
                            float *row_b, float *unp, int32_t *utb, float *bs, float *mvs, int32_t *slot) {
    const int lane = ctx.lane, nl = ctx.nl;
    float *prev = row_a, *cur = row_b;

    // spoofed previous row [0, inf, inf, ...] forces stays through the first base (pyx:365-378)
    int pw = en[0];   // width of the previous row
    int pst = -1;     // its start, so that the first base sees band_start_diff = 1
    for (int i = lane; i < pw; i += nl) prev[i] = (i == 0) ? 0.0f : RB_INF;

    int64_t off = 0;  // offset of the current base in the ragged traceback array
    int cst = 0, cen = en[0];
    float lvl = levels[0];
    for (int b = 0; b < n_bases; ++b) {
        // fetch the next base's band and level early (hidden under this base's chain)
        int nst = 0, nen = 0;
        float nlvl = 0.0f;
        if (b + 1 < n_bases) {
            nst = st[b + 1];
            nen = en[b + 1];
            nlvl = levels[b + 1];
        }
        if (lvl != lvl) lvl = 0.0f;  // NaN levels are zeroed (refine_signal_map.py:829-830)
        const int n = cen - cst;     // band width of this base
        const int d = cst - pst;     // band_start_diff (>= 1)
        const int m = pw;
        const float *s = sig + cst;

        // ---- phase 0: squared errors and move candidates (parallel) ---------------------------------
        // move[p] = prev[p - 1 + d] + bs[p] is the reference's move_score (pyx:292, 303); cells past the
        // previous band have no move: +inf there makes the chain below take the stay branch.
        {
            int mo = m - d;  // previous-row entries left after clipping its first d
            if (mo < 0) mo = 0;
            const int mv = (mo == n) ? n - 1 : mo;  // cells 0..mv have a move candidate (pyx:297-301)
            const float *pm = prev + d - 1;          // pm[p]: previous base, one sample earlier
            for (int p = lane; p < n; p += nl) {
                const float t = RB_FSUB(lvl, s[p]);
                const float e = RB_FMUL(t, t);
                bs[p] = e;
                mvs[p] = (p <= mv) ? RB_FADD(pm[p], e) : RB_INF;
            }
        }
        ctx.sync();

        // ---- phase 1: un-penalised Viterbi row (pyx:256-317) --------------------------------------
        // x[p] = min(move[p], x[p-1] + bs[p]); traceback = samples since the last move.  One serial
        // chain (FADD + FMNMX per cell), run redundantly by every lane from broadcast 16-byte loads;
        // each group of four cells is stored by one owner lane.  Starting from x = +inf, t = -1 makes
        // cell 0 (a forced move, pyx:290-293) the same code as every other cell.  fminf equals the
        // reference's `move < stay ? move : stay` for every non-NaN input.
        // With the Viterbi algorithm this row IS the result: write it straight into cur.
        {
            float *vrow = (algo == kAlgoViterbi) ? cur : unp;
            const Vec4f *__restrict__ e4 = reinterpret_cast<const Vec4f *>(bs);
            const Vec4f *__restrict__ m4 = reinterpret_cast<const Vec4f *>(mvs);
            Vec4f *__restrict__ x4 = reinterpret_cast<Vec4f *>(vrow);
            Vec4i *__restrict__ t4 = reinterpret_cast<Vec4i *>(utb);
            const int groups = (n + 3) >> 2;
            float x = RB_INF;
            int t = -1;
            Vec4f e = e4[0], mvv = m4[0];
            for (int g = 0; g < groups; ++g) {
                // fetch the next group before the dependent chain of this one
                const int gn = (g + 1 < groups) ? g + 1 : g;
                const Vec4f en_ = e4[gn], mn_ = m4[gn];
                Vec4f xo;
                Vec4i to;
                float stay;
                stay = RB_FADD(x, e.a); t = (mvv.a < stay) ? 0 : t + 1; x = fminf(mvv.a, stay); xo.a = x; to.a = t;
                stay = RB_FADD(x, e.b); t = (mvv.b < stay) ? 0 : t + 1; x = fminf(mvv.b, stay); xo.b = x; to.b = t;
                stay = RB_FADD(x, e.c); t = (mvv.c < stay) ? 0 : t + 1; x = fminf(mvv.c, stay); xo.c = x; to.c = t;
                stay = RB_FADD(x, e.d); t = (mvv.d < stay) ? 0 : t + 1; x = fminf(mvv.d, stay); xo.d = x; to.d = t;
                if ((g % nl) == lane) {
                    x4[g] = xo;
                    t4[g] = to;
                }
                e = en_;
                mvv = mn_;
            }
        }
        ctx.sync();

        if (algo == kAlgoViterbi) {
            for (int p = lane; p < n; p += nl) tb[off + p] = utb[p];
        } else {
            // ---- phase 2: short-dwell-penalty row (pyx:150-253), cells with p + d - m < n_pen ---------
            const int pf_raw = m - d + n_pen;  // first stay-only cell
            const int pf = pf_raw < n ? pf_raw : n;
            const float init = RB_FADD(kLargeScore, prev[m - 1]);
            for (int p = lane; p < pf; p += nl) {
                float c = init;
                int t = -1;
                float run = 0.0f;
                for (int k = 0; k < n_pen; ++k) {
                    if (k > p) break;
                    run = RB_FADD(run, bs[p - k]);
                    const int q = p - k - 1 + d;
                    if (q >= m) continue;
                    const float cand = RB_FADD(RB_FADD(prev[q], run), pen[k]);
                    if (cand < c) {
                        c = cand;
                        t = k;
                    }
                }
                if (p >= n_pen) {
                    const float cand = RB_FADD(unp[p - n_pen], run);
                    if (cand < c) {
                        c = cand;
                        t = utb[p - n_pen] + n_pen;
                    }
                }
                cur[p] = c;
                tb[off + p] = t;
                if (p == pf - 1) *slot = t;
            }
            ctx.sync();
            // ---- phase 3: past the previous band by >= n_pen samples only a stay is possible --------
            if (pf < n) {
                float x = cur[pf - 1];
                const int t0 = *slot;
                for (int p = pf; p < n; ++p) {
                    x = RB_FADD(x, bs[p]);
                    if ((p % nl) == lane) {
                        cur[p] = x;
                        tb[off + p] = t0 + (p - pf + 1);
                    }
                }
            }
        }
        ctx.sync();  // cur complete before it is read as prev; bs / unp / utb / slot free for reuse

        // next base
        float *tmp = prev;
        prev = cur;
        cur = tmp;
        pw = n;
        pst = cst;
        off += n;
        cst = nst;
        cen = nen;
        lvl = nlvl;
    }

    // ---- traceback (pyx:118-148): one serial walk from the last base ------------------------------
    if (lane == 0) {
        int32_t bad = kStatusOk;
        *final_score = prev[pw - 1];
        path[0] = 0;
        int nxt = en[n_bases - 1];
        path[n_bases] = nxt;
        int64_t o = off;
        for (int b = n_bases - 1; b > 0; --b) {
            const int bst = st[b], w = en[b] - bst;
            o -= w;
            const int look = nxt - 1;
            int idx = look - bst;
            if (idx < 0 || idx >= w) {  // undefined behaviour in the reference; flagged, not followed
                bad = kStatusTracebackLeftBand;
                idx = idx < 0 ? 0 : w - 1;
            }
            nxt = look - tb[o + idx];
            path[b] = nxt;
        }
        *status = bad;
    }
    ctx.sync();
}

}  // namespace refine
}  // namespace rb200
