// chunk_core.cuh -- per-lane phases of the chunked speculative TV-L1 prox (host+device so a CTA can be emulated on CPU).
//
// One fiber is cut into chunks of CH = 32 samples; one lane owns one chunk.  Shared state per fiber: the staged input y,
// a SPARSE value array vs (vs[a] = value of the segment that starts at a, valid only where a start is recorded), and one
// (present, kind-bit-0, kind-bit-1) mask word triple per chunk.  Phases, with a CTA barrier between consecutive ones:
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
//              f(y[j], vs[start(j)]), computed cooperatively with coalesced stores.
//
// Every number produced is bit-identical to the sequential scan: the same operations in the same order per segment.  The
// scan loop is FLAT: one scan step per iteration for every lane (breaks are handled by predicated code in the same
// iteration), so lanes of a warp only diverge in their trip counts.  Divisions by the small integer (i - last) use a
// correctly rounded reciprocal table + two FMAs (Markstein): q = a*r, q' = fma(fma(-q, d, a), r, q) == RN(a / d).
#pragma once
#include "taut_scan.cuh"
#include <math.h>

namespace ptv {

constexpr int CH = 32;                      // samples per chunk == bits per mask word
constexpr int RCP_N = 64;                   // reciprocal table covers divisors 0..RCP_N-1 (entry 0 unused)

struct ChunkMasks { uint32_t* P; uint32_t* K0; uint32_t* K1; };     // one word per chunk of one fiber

// exact a / d for integer d >= 1
template <typename T> struct RcpDiv {
    const T* tbl;
    PTV_HD T operator()(T a, int d) const {
        if (d < RCP_N) {
            const T r = tbl[d], dd = T(d);
            const T q = a * r;
            return fma(fma(-q, dd, a), r, q);
        }
        return a / T(d);
    }
};

template <typename T> struct LaneState {
    Scan<T> s;
    int pend_a, pend_k;     // the lane's latest segment start, not yet matched / recorded (pend_a < 0: none)
    bool finished;          // the scan reached the end of the fiber
    bool active;            // still has chunks to walk through
};

PTV_HD int low_bit(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
PTV_HD int high_bit(uint32_t m) {          // m != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)m);
#else
    return 31 - __builtin_clz(m);
#endif
}

// One regular scan step (requires s.i < n - 1), written so that it compiles to straight-line predicated code.
// Returns K_NONE / K_CEIL / K_FLOOR; on a break the finished segment is [f, s.last] (s.last already updated) with value v.
template <typename T, class LdY, class Lam>
PTV_HD int fast_step(Scan<T>& s, LdY y, Lam lam, RcpDiv<T> div, int& f, T& v) {
    const int i = s.i;
    const T yi = y(i);
    const T li = lam(i);
    const T hlo = s.hlo + (s.lo - yi);
    const T hhi = s.hhi + (s.hi - yi);
    const bool cb = li < hlo;
    const bool fb = !cb && (-li > hhi);
    if (cb | fb) {
        const int p = (cb ? s.blo : s.bhi) + 1;
        f = s.last + 1;
        v = cb ? s.lo : s.hi;
        const T yp = y(p);
        if (!Lam::weighted) {
            const T l2 = T(2) * li, nl2 = T(2) * (-li);
            s.lo = cb ? yp : nl2 + yp;
            s.hi = cb ? l2 + yp : yp;
            s.hhi = li; s.hlo = -li;
        } else {
            const T lp = lam(p - 1), lq = lam(p);
            if (cb) { s.lo = yp + lp - lq; s.hi = yp + lp + lq; }
            else    { s.hi = yp - lp + lq; s.lo = yp - lp - lq; }
            s.hhi = lq; s.hlo = -lq;
        }
        s.last = p - 1; s.blo = s.bhi = p; s.i = p + 1;
        return cb ? K_CEIL : K_FLOOR;
    }
    const int d = i - s.last;
    T hh = hhi, hl = hlo;
    if (hhi >= li)  { s.hi = s.hi + div(li - hhi, d);  hh = li;  s.bhi = i; }
    if (hlo <= -li) { s.lo = s.lo + div(-li - hlo, d); hl = -li; s.blo = i; }
    s.hhi = hh; s.hlo = hl; s.i = i + 1;
    return K_NONE;
}

// Walk through chunk c.  round == 0: the lane's own chunk, entered with a cold start (no merging possible: the chunk's
// masks are not written yet).  round >= 1: a chunk to the right, entered with the lane's scan state and pending start.
// StV(j, v): store into the sparse value array.  Returns true if the lane is still active afterwards.
template <typename T, class LdY, class StV, class Lam>
PTV_HD bool walk_chunk(int q, int round, int nchunks, int n, LdY y, StV stv, Lam lam, RcpDiv<T> div, LaneState<T>& st,
                       ChunkMasks m) {
    const int c = q + round;
    if (round > 0) {
        if (!st.active) return false;
        if (c >= nchunks) { st.active = false; return false; }
        if (st.finished) {                 // the lane's scan ended further left: it passed over this chunk without a start
            m.P[c] = 0; m.K0[c] = 0; m.K1[c] = 0;
            st.active = (c + 1 < nchunks);
            return st.active;
        }
    }
    const int cb = c * CH, ce = (cb + CH < n) ? cb + CH : n;
    uint32_t oP = 0, oK0 = 0, oK1 = 0;
    if (round > 0) { oP = m.P[c]; oK0 = m.K0[c]; oK1 = m.K1[c]; }
    else {
        st.s.begin(cb, y, lam);            // q == 0: the true start; q > 0: speculative cold start
        st.finished = false; st.pend_a = -1; st.pend_k = K_NONE;
    }
    uint32_t P = 0, K0 = 0, K1 = 0;
    bool merged = false; int mbit = 0;
    // a start found in an earlier phase that lies inside this chunk is handled first, then one scan step per iteration
    bool have = (st.pend_a >= 0);
    for (;;) {
        if (have) {
            if (st.pend_a >= ce) break;                                   // beyond this chunk: keep it pending
            const int bit = st.pend_a - cb, kk = st.pend_k - 1;
            if (((oP >> bit) & 1u) && (int)((oK0 >> bit) & 1u) == (kk & 1) && (int)((oK1 >> bit) & 1u) == (kk >> 1)) {
                merged = true; mbit = bit; break;
            }
            P |= 1u << bit; K0 |= (uint32_t)(kk & 1) << bit; K1 |= (uint32_t)(kk >> 1) << bit;
            have = false;
        }
        int k, f; T v;
        if (st.s.i < n - 1) k = fast_step<T>(st.s, y, lam, div, f, v);
        else if (st.s.i == n - 1) { int l; k = st.s.step(n, y, lam, f, l, v); }     // closing sample: rare, generic code
        else {                                                                     // the fiber ended: last open segment
            stv(st.s.last + 1, st.s.lo);
            st.finished = true; st.pend_a = -1;
            break;
        }
        if (k != K_NONE) { stv(f, v); st.pend_a = st.s.last + 1; st.pend_k = k; have = true; }
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

// Output forms (what is written for input sample yin and prox value x).
enum OutOp { OUT_X = 0, OUT_REFLECT = 1, OUT_DIFF = 2 };
template <typename T> PTV_HD T apply_out(int op, T yin, T x) {
    if (op == OUT_X) return x;
    T d = yin - x;                                            // DR_proxDiff           (src/TV2Dopt.cpp:545-546)
    if (op == OUT_REFLECT) return T(2) * d - yin;             // s = 2 s - t           (:411)
    return d;                                                 // final projection      (:427)
}

// carry[c] = position of the last segment start strictly before chunk c's first sample that is <= that sample's segment,
// i.e. the start of the segment covering sample c*CH when the chunk's own bit 0 is not set.  Sequential reference
// implementation (the kernel computes the same thing with a warp scan).
PTV_HD void fill_carry_seq(int nchunks, ChunkMasks m, int* carry) {
    int lastpos = 0;                                         // sample 0 always starts a segment
    for (int c = 0; c < nchunks; c++) {
        carry[c] = lastpos;
        if (m.P[c]) lastpos = c * CH + high_bit(m.P[c]);
    }
}

// start position of the segment covering sample j
PTV_HD int seg_start_of(int j, ChunkMasks m, const int* carry) {
    const int c = j >> 5, b = j & 31;
    const uint32_t w = m.P[c] & (0xffffffffu >> (31 - b));
    return w ? (c << 5) + high_bit(w) : carry[c];
}

}  // namespace ptv
