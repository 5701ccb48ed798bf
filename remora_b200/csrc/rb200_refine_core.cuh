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
//   phase 1b its traceback counts, re-derived from the stored row (prefix maximum) (parallel)
//   phase 2  the short-dwell-penalty row: per cell <= D+1 candidates built from the previous row, the
//            un-penalised row and bs                                              (parallel)
//   phase 3  cells more than D samples past the previous band: pure stay chain    (serial, short)
//   traceback entries go to global memory coalesced; the final traceback is one serial walk in which
//   the 32 lanes also fetch the entry of the NEXT base for every dwell 0..31 of the current one, so a
//   memory round trip usually resolves two bases.
// Rows live in shared memory for every base whose band fits `near_cap` samples and in a per-warp global
// scratch area for the few wider ones (stalls), so the shared-memory footprint - and with it the number
// of resident warps - does not depend on the widest band of the batch.
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
#define RB_PREFETCH(ptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr))
#else
#include <cmath>
#define RB_HD static inline
// host emulation is compiled with -ffp-contract=off
#define RB_FADD(a, b) ((a) + (b))
#define RB_FSUB(a, b) ((a) - (b))
#define RB_FMUL(a, b) ((a) * (b))
#define RB_INF HUGE_VALF
#define RB_PREFETCH(ptr) ((void)(ptr))
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

// Ctx supplies: int lane; int nl (lanes); void sync() (warp barrier with memory ordering);
// int scan_max(int) (inclusive prefix maximum over the lanes); int bcast_last(int) (value of lane nl-1).
//
// Per-warp scratch rows, each 16-byte aligned and holding >= its capacity rounded up to a multiple of 4;
// the last row (mvs) must be followed by at least 16 readable bytes:
//   row[2]        penalised scores of the previous / current base (alternating)
//   unp, utb      un-penalised Viterbi scores and traceback of the current base
//   bs, mvs       squared errors and move candidates of the current base's band
struct Rows {
    float *row0, *row1;
    float *unp;
    int32_t *utb;
    float *bs;
    float *mvs;
};

RB_HD Rows carve_rows(float *base, size_t cap) {
    Rows r;
    r.row0 = base;
    r.row1 = base + cap;
    r.unp = base + 2 * cap;
    r.utb = reinterpret_cast<int32_t *>(base + 3 * cap);
    r.bs = base + 4 * cap;
    r.mvs = base + 5 * cap;
    return r;
}
constexpr int kRowsPerWarp = 6;
constexpr int kSigRegs = 8;  // samples per lane fetched one base ahead (bands up to 8 x 32 samples)

// One base of the forward pass.  kD > 0: number of short-dwell penalties known at compile time (the
// default table has 3), 0: taken from n_pen.  prev (m entries, band start d samples before this band's)
// may live in either scratch space; everything written here lives in one.
template <int kD, class Ctx>
RB_HD void base_step(Ctx &ctx, const float *prev, int m, int d, float *__restrict__ cur,
                     float *__restrict__ unp, int32_t *__restrict__ utb, float *__restrict__ bs,
                     float *__restrict__ mvs, int32_t *slot, const float *__restrict__ s,
                     const float (&sreg)[kSigRegs], float lvl, int n, const float *pen, int n_pen, int algo,
                     int32_t *__restrict__ tb) {
    const int lane = ctx.lane, nl = ctx.nl;
    // ---- phase 0: squared errors and move candidates (parallel) ---------------------------------
    // move[p] = prev[p - 1 + d] + bs[p] is the reference's move_score (pyx:292, 303); cells past the
    // previous band have no move: +inf there makes the chain below take the stay branch.
    {
        int mo = m - d;  // previous-row entries left after clipping its first d
        if (mo < 0) mo = 0;
        const int mv = (mo == n) ? n - 1 : mo;  // cells 0..mv have a move candidate (pyx:297-301)
        const float *pm = prev + d - 1;          // pm[p]: previous base, one sample earlier
#pragma unroll
        for (int j = 0; j < kSigRegs; ++j) {  // samples fetched into registers during the previous base
            const int p = lane + j * nl;
            if (p < n) {
                const float t = RB_FSUB(lvl, sreg[j]);
                const float e = RB_FMUL(t, t);
                bs[p] = e;
                mvs[p] = (p <= mv) ? RB_FADD(pm[p], e) : RB_INF;
            }
        }
        for (int p = lane + kSigRegs * nl; p < n; p += nl) {
            const float t = RB_FSUB(lvl, s[p]);
            const float e = RB_FMUL(t, t);
            bs[p] = e;
            mvs[p] = (p <= mv) ? RB_FADD(pm[p], e) : RB_INF;
        }
    }
    ctx.sync();

    // ---- phase 1: un-penalised Viterbi row (pyx:256-317) --------------------------------------
    // x[p] = min(move[p], x[p-1] + bs[p]).  One serial chain, FADD + FMNMX per cell and nothing else,
    // run redundantly by every lane from broadcast 16-byte loads (next group fetched ahead); each
    // group of four cells is stored by one owner lane.  Starting from x = +inf makes cell 0 (a forced
    // move, pyx:290-293) the same code as every other cell.  fminf equals the reference's
    // `move < stay ? move : stay` for every non-NaN input.
    // With the Viterbi algorithm this row IS the result: write it straight into cur.
    float *vrow = (algo == kAlgoViterbi) ? cur : unp;
    {
        const Vec4f *__restrict__ e4 = reinterpret_cast<const Vec4f *>(bs);
        const Vec4f *__restrict__ m4 = reinterpret_cast<const Vec4f *>(mvs);
        Vec4f *__restrict__ x4 = reinterpret_cast<Vec4f *>(vrow);
        const int groups = (n + 3) >> 2;
        float x = RB_INF;
        Vec4f e = e4[0], mvv = m4[0];
        // the fetch of group g+1 runs one group past the row end on the last iteration: rows are followed
        // by readable scratch (the next row, or the pad after the last one), so no clamp is needed and
        // the body stays branch-free for the unroller
#pragma unroll 2
        for (int g = 0; g < groups; ++g) {
            const Vec4f en_ = e4[g + 1], mn_ = m4[g + 1];
            Vec4f xo;
            x = fminf(mvv.a, RB_FADD(x, e.a)); xo.a = x;
            x = fminf(mvv.b, RB_FADD(x, e.b)); xo.b = x;
            x = fminf(mvv.c, RB_FADD(x, e.c)); xo.c = x;
            x = fminf(mvv.d, RB_FADD(x, e.d)); xo.d = x;
            if ((g % nl) == lane) x4[g] = xo;
            e = en_;
            mvv = mn_;
        }
    }
    ctx.sync();

    // ---- phase 1b: traceback counts of that row (parallel) -----------------------------------------
    // The reference counts samples since the last move: t = (move < stay) ? 0 : t + 1 (pyx:305-310),
    // with cell 0 a move.  The comparison is re-evaluated per cell from the stored row with the same
    // two operations (stay = x[p-1] + bs[p]; move < stay), and the count is p minus the running
    // maximum of the move positions: a prefix maximum over the lanes plus a carry between blocks.
    {
        int carry = 0;
        for (int base = 0; base < n; base += nl) {
            const int p = base + lane;
            int pos = -1;
            if (p < n) {
                bool mw = true;
                if (p > 0) mw = mvs[p] < RB_FADD(vrow[p - 1], bs[p]);
                pos = mw ? p : -1;
            }
            pos = ctx.scan_max(pos);
            if (pos < carry) pos = carry;
            if (p < n) {
                utb[p] = p - pos;
                if (algo == kAlgoViterbi) tb[p] = p - pos;
            }
            carry = ctx.bcast_last(pos);
        }
    }
    if (algo == kAlgoViterbi) return;
    ctx.sync();

    // ---- phase 2: short-dwell-penalty row (pyx:150-253), cells with p + d - m < n_pen ---------
    const int D = kD > 0 ? kD : n_pen;
    const int pf_raw = m - d + D;  // first stay-only cell
    const int pf = pf_raw < n ? pf_raw : n;
    const float init = RB_FADD(kLargeScore, prev[m - 1]);
    for (int p = lane; p < pf; p += nl) {
        float c = init;
        int t = -1;
        float run = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (k <= p) {  // the reference leaves its loop at the first k > p
                run = RB_FADD(run, bs[p - k]);
                const int q = p - k - 1 + d;
                if (q < m) {
                    const float cand = RB_FADD(RB_FADD(prev[q], run), pen[k]);
                    if (cand < c) {
                        c = cand;
                        t = k;
                    }
                }
            }
        }
        if (p >= D) {
            const float cand = RB_FADD(unp[p - D], run);
            if (cand < c) {
                c = cand;
                t = utb[p - D] + D;
            }
        }
        cur[p] = c;
        tb[p] = t;
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
                tb[p] = t0 + (p - pf + 1);
            }
        }
    }
}

// One read.  near: rows for bases whose band is at most near_cap samples wide (shared memory on the
// GPU); far: rows for wider bands (global scratch; may be all-null when no band of the batch exceeds
// near_cap).  slot / spec: one int and 32 ints of near scratch.
//
// Preconditions (validated by the caller, they are what adjust_seq_band/validate_band establish):
//   st[0] == 0, st and en strictly increasing, en[b] > st[b], st[b] <= en[b-1], n_pen in [1, kMaxPen].
template <class Ctx>
RB_HD void refine_read_warp(Ctx &ctx, const float *__restrict__ sig, const float *__restrict__ levels,
                            const int32_t *__restrict__ st, const int32_t *__restrict__ en, int n_bases,
                            const float *pen, int n_pen, int algo, int32_t *__restrict__ tb,
                            int32_t *__restrict__ path, float *final_score, int32_t *status,
                            const Rows near, int near_cap, const Rows far, int32_t *slot, int32_t *spec) {
    const int lane = ctx.lane, nl = ctx.nl;

    // spoofed previous row [0, inf, inf, ...] forces stays through the first base (pyx:365-378)
    int pw = en[0];   // width of the previous row
    int pst = -1;     // its start, so that the first base sees band_start_diff = 1
    float *prev = (pw <= near_cap) ? near.row1 : far.row1;
    for (int i = lane; i < pw; i += nl) prev[i] = (i == 0) ? 0.0f : RB_INF;
    ctx.sync();

    int64_t off = 0;  // offset of the current base in the ragged traceback array
    int cst = 0, cen = en[0];
    float lvl = levels[0];
    float sreg[kSigRegs];  // this base's first samples, fetched during the previous base
#pragma unroll
    for (int j = 0; j < kSigRegs; ++j) sreg[j] = (lane + j * nl < cen) ? sig[lane + j * nl] : 0.0f;
    for (int b = 0; b < n_bases; ++b) {
        // fetch the next base's band and level early (hidden under this base's chain), and pull the
        // samples its band adds over this one towards the SM
        int nst = 0, nen = 0;
        float nlvl = 0.0f;
        if (b + 1 < n_bases) {
            nst = st[b + 1];
            nen = en[b + 1];
            nlvl = levels[b + 1];
            if (cen + lane * 32 < nen) RB_PREFETCH(sig + cen + lane * 32);
        }
        float snext[kSigRegs];
#pragma unroll
        for (int j = 0; j < kSigRegs; ++j)
            snext[j] = (nst + lane + j * nl < nen) ? sig[nst + lane + j * nl] : 0.0f;
        if (lvl != lvl) lvl = 0.0f;  // NaN levels are zeroed (refine_signal_map.py:829-830)
        const int n = cen - cst;     // band width of this base
        const int d = cst - pst;     // band_start_diff (>= 1)
        float *cur;
        if (n <= near_cap) {
            cur = (b & 1) ? near.row1 : near.row0;
            if (n_pen == 3)
                base_step<3>(ctx, prev, pw, d, cur, near.unp, near.utb, near.bs, near.mvs, slot, sig + cst, sreg,
                             lvl, n, pen, n_pen, algo, tb + off);
            else
                base_step<0>(ctx, prev, pw, d, cur, near.unp, near.utb, near.bs, near.mvs, slot, sig + cst, sreg,
                             lvl, n, pen, n_pen, algo, tb + off);
        } else {
            cur = (b & 1) ? far.row1 : far.row0;
            base_step<0>(ctx, prev, pw, d, cur, far.unp, far.utb, far.bs, far.mvs, slot, sig + cst, sreg, lvl, n,
                         pen, n_pen, algo, tb + off);
        }
        ctx.sync();  // cur complete before it is read as prev; bs / unp / utb / slot free for reuse

        prev = cur;
        pw = n;
        pst = cst;
        off += n;
        cst = nst;
        cen = nen;
        lvl = nlvl;
#pragma unroll
        for (int j = 0; j < kSigRegs; ++j) sreg[j] = snext[j];
    }

    // ---- traceback (pyx:118-148) ----------------------------------------------------------------
    // path[b] = look - tb_b[look - st[b]] with look = path[b+1] - 1: a chain of dependent loads from
    // HBM.  All lanes walk together; while the entry of base b is in flight, lane j fetches the entry
    // base b-1 would need if base b turns out to have dwelt j samples (32 consecutive words), so one
    // round trip usually settles two bases.  Out-of-band lookups (undefined behaviour in the
    // reference) are never speculated: they go through the plain step, which clamps and flags them.
    {
        int32_t bad = kStatusOk;
        int nxt = en[n_bases - 1];
        if (lane == 0) {
            *final_score = prev[pw - 1];
            path[0] = 0;
            path[n_bases] = nxt;
        }
        int64_t o = off;
        int b = n_bases - 1;
        while (b > 0) {
            const int bst = st[b], w = en[b] - bst;
            o -= w;
            const int look = nxt - 1;
            int idx = look - bst;
            if (idx < 0 || idx >= w) {
                bad = kStatusTracebackLeftBand;
                idx = idx < 0 ? 0 : w - 1;
            }
            // speculative fetch for base b-1, assuming base b dwelt `lane` samples
            int cand = -2;  // -2: not usable
            int w2 = 0;
            if (b > 1) {
                const int bst2 = st[b - 1];
                w2 = en[b - 1] - bst2;
                const int idx2 = (look - lane) - 1 - bst2;
                if (idx2 >= 0 && idx2 < w2) cand = tb[o - w2 + idx2];
            }
            const int t = tb[o + idx];
            nxt = look - t;
            if (lane == 0) path[b] = nxt;
            spec[lane] = cand;
            ctx.sync();
            int t2 = -2;
            if (b > 1 && t >= 0 && t < nl) t2 = spec[t];
            ctx.sync();
            if (t2 != -2) {
                nxt = (nxt - 1) - t2;
                if (lane == 0) path[b - 1] = nxt;
                o -= w2;
                b -= 2;
            } else {
                b -= 1;
            }
        }
        if (lane == 0) *status = bad;
    }
    ctx.sync();
}

}  // namespace refine
}  // namespace rb200
