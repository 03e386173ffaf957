// emu_chunked.cu -- TEST INFRASTRUCTURE: runs the per-lane phases of proxtv_b200/csrc/chunk_core.cuh on the CPU, lane by
// lane and phase by phase, exactly as one CTA of kernels_chunked.cu executes them (lanes interact only across barriers,
// so a sequential sweep over lanes per phase is an exact emulation; the one-writer-per-chunk-per-round discipline that
// makes this true is asserted, not assumed).  The output array doubles as the sparse value store, as in the kernel.
// Not part of the product; built by tests/test_chunk_emulation.py with nvcc (host code only).
#include "../../proxtv_b200/csrc/chunk_core.cuh"
#include <vector>
#include <string.h>

using namespace ptv;

template <typename T> struct PtrLd { const T* p; PTV_HD T operator()(int i) const { return p[i]; } };
// value store that records which chunk each round writes, to assert the write discipline
template <typename T, int CHT> struct ChkSt {
    T* p; int* owner_round; int* owner_lane; int* cur_round; int* cur_lane; int* bad;
    PTV_HD void operator()(int j, T v) const {
        int c = j / CHT;
        if (owner_round[c] == *cur_round && owner_lane[c] != *cur_lane) *bad = 1;     // two lanes wrote one chunk in a round
        owner_round[c] = *cur_round; owner_lane[c] = *cur_lane;
        p[j] = v;
    }
};

template <typename T, int CHT = CH>
static int emu(const T* yin, int n, T lam, const T* lamv, T* x, int out_op, int* rounds_out) {
    if (n <= 0) return 0;
    const int nchunks = (n + CHT - 1) / CHT;
    std::vector<T> ys(yin, yin + n), ws(n, T(0)), cval(nchunks);
    std::vector<RcpPair<T>> rcp(RCP_N);
    for (int d = 0; d < RCP_N; d++) { rcp[d].r = d ? T(1) / T(d) : T(0); rcp[d].d = T(d); }
    if (lamv) memcpy(ws.data(), lamv, sizeof(T) * (size_t)(n - 1));
    for (int j = 0; j < n; j++) x[j] = T(-12345);                 // the output row doubles as the sparse value store
    std::vector<uint32_t> P(nchunks), K0(nchunks), K1(nchunks);
    ChunkMasks m{P.data(), K0.data(), K1.data()};
    std::vector<LaneState<T>> st(nchunks);
    std::vector<int> orr(nchunks, -1), orl(nchunks, -1);
    int cur_round = 0, cur_lane = 0, bad = 0;
    PtrLd<T> y{ys.data()};
    ChkSt<T, CHT> stv{x, orr.data(), orl.data(), &cur_round, &cur_lane, &bad};
    RcpDiv<T> div{rcp.data()};
    auto run = [&](auto lamf) -> int {
        int r = 0;
        for (;; r++) {
            cur_round = r;
            std::vector<uint32_t> sP(P), sK0(K0), sK1(K1);
            bool any = false;
            for (int q = 0; q < nchunks; q++) {
                cur_lane = q;
                int c = q + r;          // the chunk lane q reads/writes this round must still hold its pre-round masks
                if (r > 0 && st[q].active && c < nchunks && (P[c] != sP[c] || K0[c] != sK0[c] || K1[c] != sK1[c])) return -1;
                if (r == 0) { st[q].active = false; st[q].finished = false; st[q].pend_a = -1; st[q].pend_k = K_NONE;
                              any |= walk_chunk<T, true, CHT>(q, 0, nchunks, n, y, stv, lamf, div, st[q], m); }
                else any |= walk_chunk<T, false, CHT>(q, r, nchunks, n, y, stv, lamf, div, st[q], m);
            }
            if (bad) return -2;
            if (!any) break;
        }
        if (rounds_out) *rounds_out = r;
        for (int c = 0; c < nchunks; c++) cval[c] = x[carry_of<CHT>(c, m)];      // gather phase (barrier after it)
        if (out_op == -1) {
            // SPARSE result (kernels_chunked.cu, Mk / Cv) expanded the way the fused scatter does it (transpose.cu, EXPAND): per
            // 32-sample window one start mask (two 16-bit halves when CHT == 16) and the value entering the window; the segment
            // values stay in x at their start positions.  Expansion goes to a separate array first: x is its own source.
            const int nwin = (n + 31) / 32;
            std::vector<T> dense(n);
            for (int w = 0; w < nwin; w++) {
                const uint32_t mk = (CHT == 32) ? P[w] : (P[2 * w] | ((2 * w + 1 < nchunks) ? (P[2 * w + 1] << 16) : 0u));
                const T cv = cval[(CHT == 32) ? w : 2 * w];
                for (int b = 0; b < 32 && w * 32 + b < n; b++) {
                    const uint32_t ww = mk & (0xffffffffu >> (31 - b));
                    dense[w * 32 + b] = ww ? x[w * 32 + high_bit(ww)] : cv;
                }
            }
            for (int j = 0; j < n; j++) x[j] = dense[j];
            return 0;
        }
        for (int c = 0; c < nchunks; c++) {                                        // fill, one window at a time
            T v[CHT];
            for (int b = 0; b < CHT && c * CHT + b < n; b++) {                     // all reads of the window ...
                int sa = seg_start_in_chunk<CHT>(c, b, m);
                if (sa >= 0 && sa / CHT != c) return -3;                           // ... stay inside the window
                v[b] = sa >= 0 ? x[sa] : cval[c];
            }
            for (int b = 0; b < CHT && c * CHT + b < n; b++) x[c * CHT + b] = apply_out<T>(out_op, ys[c * CHT + b], v[b]);
        }
        return 0;
    };
    if (lamv) return run(ArrayLam<T, PtrLd<T>>{PtrLd<T>{ws.data()}});
    return run(UniformLam<T>{lam});
}

extern "C" int emu_chunked_f64(const double* y, int n, double lam, const double* lamv, double* x, int out_op, int* rounds) {
    return emu<double>(y, n, lam, lamv, x, out_op, rounds);
}
extern "C" int emu_chunked_f64_ch16(const double* y, int n, double lam, const double* lamv, double* x, int out_op, int* rounds) {
    return emu<double, 16>(y, n, lam, lamv, x, out_op, rounds);
}
extern "C" int emu_chunked_f32(const float* y, int n, float lam, const float* lamv, float* x, int out_op, int* rounds) {
    return emu<float>(y, n, lam, lamv, x, out_op, rounds);
}
// exactness of the reciprocal-table division against IEEE division: returns the number of mismatches
extern "C" long emu_check_table_division(long reps_per_divisor, unsigned long long seed) {
    std::vector<RcpPair<double>> rd(RCP_N); std::vector<RcpPair<float>> rf(RCP_N);
    for (int d = 0; d < RCP_N; d++) { rd[d].r = d ? 1.0 / d : 0.0; rd[d].d = d; rf[d].r = d ? 1.0f / d : 0.0f; rf[d].d = (float)d; }
    RcpDiv<double> dd{rd.data()}; RcpDiv<float> df{rf.data()};
    unsigned long long s = seed ? seed : 88172645463325252ULL; long bad = 0;
    for (int d = 1; d < RCP_N; d++)
        for (long k = 0; k < reps_per_divisor; k++) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            double a; unsigned long long mant = (s & 0xFFFFFFFFFFFFFULL) | ((unsigned long long)(1023 - 40 + (s >> 58) * 2) << 52);
            memcpy(&a, &mant, 8); if (s & (1ULL << 57)) a = -a;
            if (k % 3 == 1) { a = (double)(long long)(s >> 24) * d; if (k & 4) a = nextafter(a, (k & 8) ? 1e300 : -1e300); }
            if (dd(a, d) != a / d) bad++;
            float af = (float)a;
            if (k % 3 == 1) af = (float)((s >> 44) * (unsigned long long)d);
            if (df(af, d) != af / (float)d) bad++;
        }
    return bad;
}
