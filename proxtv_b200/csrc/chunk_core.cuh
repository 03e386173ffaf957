// chunk_core.cuh -- per-lane phases of the chunked speculative TV-L1 prox (host+device so a CTA can be emulated on CPU).
//
// One fiber is cut into chunks of CH = 32 samples; one lane owns one chunk.  State per fiber: the staged input y (shared
// memory), one (present, kind-bit-0, kind-bit-1) mask word triple per chunk (shared memory), and a SPARSE value store
// vs (vs[a] = value of the segment that starts at a, meaningful only where a start is recorded) for which the kernel uses
// the fiber's own OUTPUT row in global memory (it lives in L2 for the few microseconds between the scan and the fill),
// so shared memory holds 8 bytes per sample and twice as many fibers are resident per SM.
// Phases, with a CTA barrier between consecutive ones:
//
//   walk own chunk (round 0)  every lane runs the exact scan (taut_scan.cuh) from a COLD START at its chunk's first sample
//              (lane 0: the true start of the fiber) until its first segment start at or beyond the chunk end, recording
//              the starts it creates inside its chunk in the masks and each finished segment's value in vs.
//   round r >= 1  every lane that has not merged yet walks through chunk q + r: at each of its own segment starts (a, kind)
//              it looks the position up in that chunk's CURRENT masks.  Equal (position, kind) means both scans are in the
//              identical renewal state -- the state after a break is a pure function of (position, kind) -- hence identical
//              from there on: the lane splices its starts in front of the matched one and retires.  Otherwise it overrides
//              the chunk's masks (and vs entries) with its own and goes on.  In round r only lane q writes chunk q + r, so
//              rounds are race free; lanes further left arrive later and therefore win, and lane 0 -- the true scan --
//              wins everywhere it passes.  On noise-like data one round suffices (merge distance: median 3, p99.9 < 32
//              samples); a fiber without breaks degrades to one lane's sequential scan, never to a wrong
//              answer.
//   fill       the final masks are the exact segmentation and vs holds the exact segment values: every output sample is
//              f(y[j], vs[start(j)]).  Done one 32-sample window at a time (all reads of vs inside the window, then all
//              writes), with the value of the segment entering each window gathered beforehand, because vs IS the output.
//
// Every number produced is bit-identical to the sequential scan: the same operations in the same order per segment.  The
// scan loop is FLAT: one scan step per iteration for every lane (a break is handled inside the same iteration), so lanes
// of a warp only diverge in their trip counts.  Divisions by the small integer (i - last) use a correctly rounded
// reciprocal table + two FMAs (Markstein): q = a*r, q' = fma(fma(-q, d, a), r, q) == RN(a / d); verified exhaustively
// against IEEE division for all table divisors (tests/test_chunk_emulation.py::test_table_division_is_exact).
#pragma once
#include "taut_scan.cuh"
#include <math.h>

namespace ptv {

constexpr int CH = 32;                      // samples per chunk == bits per mask word
constexpr int RCP_N = 64;                   // reciprocal table covers divisors 0..RCP_N-1 (entry 0 unused)

struct ChunkMasks { uint32_t* P; uint32_t* K0; uint32_t* K1; };     // one word per chunk of one fiber

template <typename T> struct alignas(2 * sizeof(T)) RcpPair { T r, d; };     // (1/d correctly rounded, d)

// out-of-line IEEE division for divisors beyond the table (rare: a segment longer than RCP_N samples is being built);
// must not be inlined or the compiler evaluates it speculatively on the hot path
template <typename T> __host__ __device__ __noinline__ T slow_div(T a, int d) { return a / T(d); }

// exact a / d for integer d >= 1
template <typename T> struct RcpDiv {
    const RcpPair<T>* tbl;
    PTV_HD T operator()(T a, int d) const {
        if (d < RCP_N) {
            const RcpPair<T> e = tbl[d];
            const T q = a * e.r;
            return fma(fma(-q, e.d, a), e.r, q);
        }
        return slow_div<T>(a, d);
    }
};

template <typename T> struct LaneState {
    Scan<T> s;
    int pend_a, pend_k;     // the lane's latest segment start, not yet matched / recorded (pend_a < 0: none)
    bool finished;          // the scan reached the end of the fiber
    bool active;            // still has chunks to walk through
};

PTV_HD int high_bit(uint32_t m) {          // m != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)m);
#else
    return 31 - __builtin_clz(m);
#endif
}

// Walk through chunk c = q + round.  ROUND0: the lane's own chunk, entered with a cold start (no merging possible: the
// chunk's masks are not written yet).  Otherwise a chunk to the right, entered with the lane's scan state and pending
// start.  StV(j, v): store into the sparse value store.  Returns true if the lane is still active afterwards.
// CHT: samples per chunk (32, or 16 for the weighted float64 instantiation, whose second staged row halves residency: twice
// the lanes per fiber then restore the warps per SM); mask words stay 32 bits wide, a 16-sample chunk uses the low half.
template <typename T, bool ROUND0, int CHT = CH, class LdY, class StV, class Lam>
PTV_HD bool walk_chunk(int q, int round, int nchunks, int n, LdY y, StV stv, Lam lam, RcpDiv<T> div, LaneState<T>& st,
                       ChunkMasks m) {
    const int c = q + round;
    uint32_t oP = 0, oK0 = 0, oK1 = 0, P = 0, K0 = 0, K1 = 0;
    const int cb = c * CHT, ce = (cb + CHT < n) ? cb + CHT : n;
    if (ROUND0) {
        st.s.begin(cb, y, lam);            // q == 0: the true start; q > 0: speculative cold start
        st.finished = false; st.pend_a = -1; st.pend_k = K_NONE;
    } else {
        if (!st.active) return false;
        if (c >= nchunks) { st.active = false; return false; }
        if (st.finished || st.pend_a >= ce) {      // nothing of this lane's scan starts inside this chunk: it is covered
            m.P[c] = 0; m.K0[c] = 0; m.K1[c] = 0;  // by the lane's open segment -> override with "no starts"
            st.active = (c + 1 < nchunks);
            return st.active;
        }
        oP = m.P[c]; oK0 = m.K0[c]; oK1 = m.K1[c];
        // the pending start lies inside this chunk (pend_a >= cb holds: the previous round consumed everything before)
        const int bit = st.pend_a - cb, kk = st.pend_k - 1;
        if (((oP >> bit) & 1u) && (int)((oK0 >> bit) & 1u) == (kk & 1) && (int)((oK1 >> bit) & 1u) == (kk >> 1)) {
            const uint32_t keep = ~0u << bit;      // merged right away: drop the chunk's starts that lie inside the lane's
            m.P[c] = oP & keep; m.K0[c] = oK0 & keep; m.K1[c] = oK1 & keep;      // open segment, keep the rest
            st.active = false;
            return false;
        }
        P = 1u << bit; K0 = (uint32_t)(kk & 1) << bit; K1 = (uint32_t)(kk >> 1) << bit;
    }
    bool merged = false; int mbit = 0;
    // hot-path guard folded into one compare: i < stop, stop = min(n - 1, last + RCP_N), refreshed whenever `last` moves
    int stop = (st.s.last + RCP_N < n - 1) ? st.s.last + RCP_N : n - 1;
    for (;;) {
        int k, f, a; T v;
        Scan<T>& s = st.s;
        const int i = s.i;
        if (i < stop) {
            const int d = i - s.last;
            // ---- hot path: regular step (taut_scan.cuh Scan::step, i < n-1 branch), straight-line, break handled in place.
            //      Both touch updates are computed unconditionally with the exact table division and selected. ----
            const T yi = y(i);
            const T li = lam(i);
            const RcpPair<T> e = div.tbl[d];
            const T hlo = s.hlo + (s.lo - yi);
            const T hhi = s.hhi + (s.hi - yi);
            const bool cbk = li < hlo;
            const bool fbk = !cbk && (-li > hhi);
            if (!(cbk | fbk)) {
                const T nh = li - hhi, nl = -li - hlo;
                const T qh0 = nh * e.r, ql0 = nl * e.r;
                const T qh = fma(fma(-qh0, e.d, nh), e.r, qh0);      // == nh / d, correctly rounded
                const T ql = fma(fma(-ql0, e.d, nl), e.r, ql0);      // == nl / d
                const bool thi = hhi >= li, tlo = hlo <= -li;
                s.hi = thi ? s.hi + qh : s.hi;  s.hhi = thi ? li : hhi;   s.bhi = thi ? i : s.bhi;
                s.lo = tlo ? s.lo + ql : s.lo;  s.hlo = tlo ? -li : hlo;  s.blo = tlo ? i : s.blo;
                s.i = i + 1;
                continue;
            }
            a = (cbk ? s.blo : s.bhi) + 1;
            f = s.last + 1;
            v = cbk ? s.lo : s.hi;
            const T yp = y(a);
            if (!Lam::weighted) {
                const T l2 = T(2) * li, nl2 = T(2) * (-li);
                s.lo = cbk ? yp : nl2 + yp;
                s.hi = cbk ? l2 + yp : yp;
                s.hhi = li; s.hlo = -li;
            } else {
                const T lp = lam(a - 1), lq = lam(a);
                if (cbk) { s.lo = yp + lp - lq; s.hi = yp + lp + lq; }
                else     { s.hi = yp - lp + lq; s.lo = yp - lp - lq; }
                s.hhi = lq; s.hlo = -lq;
            }
            s.last = a - 1; s.blo = s.bhi = a; s.i = a + 1;
            stop = (a - 1 + RCP_N < n - 1) ? a - 1 + RCP_N : n - 1;
            // the finished segment [f, a-1] has value v; a new one starts at a, kind CEIL (cbk) or FLOOR (kind bit 0 set)
            stv(f, v);
            if (a >= ce) { st.pend_a = a; st.pend_k = cbk ? K_CEIL : K_FLOOR; break; }
            const uint32_t mb = 1u << (a - cb);
            if (!ROUND0 && (oP & mb) && !(oK1 & mb) && (((oK0 & mb) != 0u) == fbk)) { merged = true; mbit = a - cb; break; }
            P |= mb; K0 |= fbk ? mb : 0u;
            continue;
        } else if (i < n) {                        // closing sample, or a segment longer than the table: rare, generic code
            int l;
            k = s.step(n, y, lam, f, l, v);
            stop = (s.last + RCP_N < n - 1) ? s.last + RCP_N : n - 1;
            if (k == K_NONE) continue;
            a = l + 1;
        } else {                                   // the fiber ended: value of the last open segment
            stv(s.last + 1, s.lo);
            st.finished = true; st.pend_a = -1;
            break;
        }
        // ---- a segment [f, a-1] with value v was finished and a new one starts at a with kind k ----
        stv(f, v);
        if (a >= ce) { st.pend_a = a; st.pend_k = k; break; }
        const int bit = a - cb, kk = k - 1;
        if (!ROUND0 && ((oP >> bit) & 1u) && (int)((oK0 >> bit) & 1u) == (kk & 1) && (int)((oK1 >> bit) & 1u) == (kk >> 1)) {
            merged = true; mbit = bit; break;
        }
        const uint32_t mb = 1u << bit;
        P |= mb; K0 |= (uint32_t)(kk & 1) << bit; K1 |= (uint32_t)(kk >> 1) << bit;
    }
    if (merged) {
        const uint32_t keep = ~0u << mbit;
        m.P[c] = P | (oP & keep); m.K0[c] = K0 | (oK0 & keep); m.K1[c] = K1 | (oK1 & keep);
        st.active = false;
    } else {
        m.P[c] = P; m.K0[c] = K0; m.K1[c] = K1;
        st.active = (c + 1 < nchunks);
    }
    return st.active;
}

// Output forms (what is written for input sample yin and prox value x).  The two DR_ROWS forms additionally read the
// arrays the pass was staged from (A = Y, B = s, C = t) at the sample's own position and fuse the second half of a
// Douglas-Rachford iteration into the pass (src/TV2Dopt.cpp:515-520, :419, :422 resp. :429-430), same operation order.
enum OutOp { OUT_X = 0, OUT_REFLECT = 1, OUT_DIFF = 2, OUT_DR_ROWS = 3, OUT_DR_ROWS_FINAL = 4,
             OUT_DRW_ROWS = 5, OUT_DRW_ROWS_FINAL = 6 };     // weighted Douglas-Rachford (src/TV2DWopt.cpp): opposite signs
template <typename T> PTV_HD T apply_out(int op, T yin, T x) {
    if (op == OUT_X) return x;
    T d = yin - x;                                            // DR_proxDiff           (src/TV2Dopt.cpp:545-546)
    if (op == OUT_REFLECT) return T(2) * d - yin;             // s = 2 s - t           (:411)
    return d;                                                 // final projection      (:427)
}

template <typename T> PTV_HD T apply_out_ex(int op, T yin, T x, const T* A, const T* B, const T* C, long long g) {
    if (op < OUT_DR_ROWS) return apply_out<T>(op, yin, x);
    const T Yv = A[g], sv = B[g];
    T tb = Yv - (yin - x);                                    // output[idx] = ref[idx] - (in - prox)      (:520, :546)
    if (op == OUT_DR_ROWS_FINAL) return tb - sv;              // s = tb - s                                (:430)
    tb = T(2) * tb - sv;                                      // tb = 2 tb - s                             (:419)
    return T(0.5) * (C[g] + tb);                              // t = 0.5 (t + tb)                          (:422)
}

// apply_out_ex plus the weighted Douglas-Rachford forms.  Only the sequential kernel and the fused scatter use this one: the
// two extra cases cost the chunked kernels 2-7 % (measured) through a fatter fill loop, and they never see those forms.
template <typename T> PTV_HD T apply_out_any(int op, T yin, T x, const T* A, const T* B, const T* C, long long g) {
    if (op < OUT_DRW_ROWS) return apply_out_ex<T>(op, yin, x, A, B, C, g);
    const T Yv = A[g], sv = B[g];
    T tw = (yin - x) - Yv;                                    // output[idx] = out - ref             (src/TV2DWopt.cpp:218)
    if (op == OUT_DRW_ROWS_FINAL) return -sv - tw;            // s = -s - tb                          (:125)
    tw = T(-2) * tw - sv;                                     // tb = -2 tb - s                       (:115)
    return T(0.5) * (C[g] + tw);                              // t = 0.5 (t + tb)                     (:118)
}

// carry[c] = start of the segment that covers sample c*CH when the chunk's own bit 0 is not set: the last recorded start
// before the chunk (sample 0 always starts a segment).  Lane-parallel form: each lane looks left for a non-empty chunk.
template <int CHT = CH> PTV_HD int carry_of(int c, ChunkMasks m) {
    int c2 = c - 1;
    while (c2 >= 0 && m.P[c2] == 0) c2--;
    return (c2 >= 0) ? c2 * CHT + high_bit(m.P[c2]) : 0;
}

// start of the segment covering sample j = c*CH + b, or -1 if that segment starts before the chunk (use the carry)
template <int CHT = CH> PTV_HD int seg_start_in_chunk(int c, int b, ChunkMasks m) {
    const uint32_t w = m.P[c] & (0xffffffffu >> (31 - b));
    return w ? c * CHT + high_bit(w) : -1;
}

}  // namespace ptv
