// taut_scan.cuh -- device-side primitives of the exact 1D TV-L1 prox (linearized taut-string scan).
//
// Behavioural reference (what this computes, not how):
//   unweighted  src/TVL1opt_hybridtautstring.cpp:56-235 == src/TVL1opt.cpp:359-564 (linearizedTautString_TV1)
//   weighted    src/TVL1Wopt.cpp:364-567 (tautString_TV1_Weighted)
// The arithmetic (operation order, IEEE division, no FMA contraction: this TU is built with -fmad=false) is kept
// identical to those routines so that the scan's break decisions and segment values are bit-identical to the
// reference's linearized taut-string on the same input.
//
// B200 formulation.  The sequential scan is re-expressed as a *renewal process*: every finished segment leaves the scan
// in a state that is a pure function of (start position, start kind) -- see Start kinds below.  That is what lets many
// lanes scan chunks of one fiber speculatively and stitch them exactly (kernels_chunked.cu), and what lets segment
// values be recomputed independently per segment (replay()).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PTV_HD __host__ __device__ __forceinline__

namespace ptv {

// Kind of a segment start (how the scan state at that position was created).
enum : int {
    K_NONE = 0,
    K_CEIL = 1,      // previous segment ended by a ceiling violation  (hybridtautstring.cpp:93-111)
    K_FLOOR = 2,     // previous segment ended by a floor violation    (:119-137)
    K_ENDCEIL = 3,   // ceiling violation detected at the last sample  (:176-193)
    K_ENDFLOOR = 4,  // floor violation detected at the last sample    (:201-218)
    K_BEGIN = 5      // start of the fiber (:76-80) -- or a speculative cold start, which uses the same state
};

template <typename T> struct Eps { };
template <> struct Eps<double> { static PTV_HD double v() { return 1e-10; } };   // general.h:64
template <> struct Eps<float>  { static PTV_HD float  v() { return 1e-10f; } };

// Per-edge weight access: uniform lambda or a per-fiber array lam[0..n-2] (strided like the fiber itself).
template <typename T> struct UniformLam {
    T lam;
    static constexpr bool weighted = false;
    PTV_HD T operator()(int) const { return lam; }
};
template <typename T, class Ld> struct ArrayLam {
    Ld ld;      // ld(i) -> lam[i]
    static constexpr bool weighted = true;
    PTV_HD T operator()(int i) const { return ld(i); }
};

template <typename T> struct Scan {
    T lo, hi;        // slopes (candidate segment values) of the lower / upper string
    T hlo, hhi;      // their heights over the tube centre
    int blo, bhi;    // last touch positions
    int last;        // last position of the previous (finished) segment
    int i;           // next sample to process

    // State at the beginning of a fiber (or of a speculative cold start at position p).
    template <class LdY, class Lam>
    PTV_HD void begin(int p, LdY y, Lam lam) {
        T l0 = lam(p);
        T y0 = y(p);
        hlo = hhi = T(0);
        lo = -l0 + y0;
        hi = l0 + y0;
        last = p - 1;
        blo = bhi = p;
        i = p;
    }

    // Renewal state: a segment of the given kind starts at position p (0 < p < n).
    template <class LdY, class Lam>
    PTV_HD void renew(int p, int kind, int n, LdY y, Lam lam) {
        T yp = y(p);
        last = p - 1;
        blo = bhi = p;
        if (!Lam::weighted) {
            T l = lam(0), l2 = T(2) * l, nl2 = T(2) * (-l);
            if (kind == K_CEIL || kind == K_ENDCEIL) { lo = yp; hi = l2 + yp; }
            else                                     { hi = yp; lo = nl2 + yp; }
            if (kind == K_CEIL || kind == K_FLOOR) { hhi = l; hlo = -l; i = p + 1; }
            else if (kind == K_ENDCEIL)            { hhi = hlo = -l; i = p; }
            else                                   { hhi = hlo = l;  i = p; }
        } else {
            T lp = lam(p - 1);
            if (kind == K_CEIL || kind == K_FLOOR) {
                T li = lam(p);
                if (kind == K_CEIL) { lo = yp + lp - li; hi = yp + lp + li; }
                else                { hi = yp - lp + li; lo = yp - lp - li; }
                hhi = li; hlo = -li; i = p + 1;
            } else {
                T li = (p == n - 1) ? T(0) : lam(p);
                if (kind == K_ENDCEIL) { lo = yp + lp - li; hi = yp + lp + li; hhi = hlo = -lp; }
                else                   { hi = yp - lp + li; lo = yp - lp - li; hhi = hlo = lp; }
                i = p;
            }
        }
    }

    // One scan step at sample i (requires i < n).  Returns K_NONE if the scan simply advanced (or finished: i == n
    // afterwards), otherwise the kind of the segment start that was just created; in that case the finished segment
    // is [seg_first, seg_last] with value seg_val, and the state has been renewed at position seg_last + 1.
    template <class LdY, class Lam>
    PTV_HD int step(int n, LdY y, Lam lam, int& seg_first, int& seg_last, T& seg_val) {
        const T yi = y(i);
        if (i < n - 1) {
            const T li = lam(i);
            hlo += lo - yi;
            if (li < hlo) {
                seg_first = last + 1; seg_last = blo; seg_val = lo;
                renew(blo + 1, K_CEIL, n, y, lam);
                return K_CEIL;
            }
            hhi += hi - yi;
            if (-li > hhi) {
                seg_first = last + 1; seg_last = bhi; seg_val = hi;
                renew(bhi + 1, K_FLOOR, n, y, lam);
                return K_FLOOR;
            }
            if (hhi >= li)  { hi += (li - hhi) / T(i - last);  hhi = li;  bhi = i; }
            if (hlo <= -li) { lo += (-li - hlo) / T(i - last); hlo = -li; blo = i; }
            i++;
            return K_NONE;
        }
        // last sample: the tube closes on the centre line
        hlo += lo - yi;
        if (hlo > Eps<T>::v()) {
            seg_first = last + 1; seg_last = blo; seg_val = lo;
            renew(blo + 1, K_ENDCEIL, n, y, lam);
            return K_ENDCEIL;
        }
        hhi += hi - yi;
        if (hhi < -Eps<T>::v()) {
            seg_first = last + 1; seg_last = bhi; seg_val = hi;
            renew(bhi + 1, K_ENDFLOOR, n, y, lam);
            return K_ENDFLOOR;
        }
        if (hlo <= T(0)) lo += (-hlo) / T(i - last);
        i++;   // == n: finished; the open segment [last+1, n-1] has value lo
        return K_NONE;
    }
};

}  // namespace ptv
