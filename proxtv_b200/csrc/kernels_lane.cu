// kernels_lane.cu -- the lane-per-fiber streaming TV-L1 prox kernel for sm_100a (algorithm and shared code: lane_core.cuh).
//
// One warp = 32 adjacent fibers x one chunk of rows.  Rows enter the warp's circular shared-memory window as TMA tiles
// (cp.async.bulk.tensor, one mbarrier per tile slot; SASS: UTMALDG), the 32 lanes scan their own window column (bank-conflict
// free), and finished rows leave through the drain, which also applies the fused Douglas-Rachford arithmetic:
//     STRIDED layout (fibers adjacent in memory: every dimension but the first of a column-major array)
//         a window row IS one contiguous 32-sample line of the array: the tile lands in place, rows are drained with one
//         coalesced store each -- no transposed copy of anything (this replaces the gather / scatter kernels of transpose.cu)
//     CONTIG layout (fibers contiguous: the first dimension)
//         a TMA box of 16 rows x 32 fibers lands fiber-major with the 128-byte hardware swizzle and is transposed into the window
//         by the warp; finished boxes are transposed back and leave with a TMA store (UTMASTG)
// No CTA-wide barrier anywhere: warps are independent (a CTA is just NW of them sharing the reciprocal table).
#include "ptv_internal.h"
#include "lane_core.cuh"
#include <cuda.h>
#include <stdio.h>

namespace ptvl {

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(s32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- kernel arguments
enum LaneOp { LOP_PLAIN = 0, LOP_DR_B = 1, LOP_DR_B_FINAL = 2 };

template <typename T> struct LaneArgs {
    CUtensorMap tmA, tmB, tmC;      // loads (B, C only for the fused Douglas-Rachford forms)
    CUtensorMap tmX;                // store (CONTIG layout)
    const T* A; const T* B; const T* C; T* X;
    long long inc;                  // fiber stride (STRIDED: fibers per slab; CONTIG: 1)
    long long per_slab;             // fibers per slab (STRIDED: inc; CONTIG: nf)
    int slabs, gps;                 // groups of 32 fibers per slab
    ChunkPlan plan;
    T lam;
    int* rec;                       // [3][nchunks][slabs*gps*32] chunk records (in, out, overflow)
    int* group_count;               // [slabs*gps] finished-task counters (self-resetting)
    unsigned long long* stats;      // [0] repair scans, [1] retired lanes
    long long ntasks;
};

// fused arithmetic of a pass: what a lane scans (in) and what it writes for prox value x (out)
template <typename T, int OP> struct PassOp {
    static __device__ __forceinline__ T in(T a, T b, T c) {
        if (OP == LOP_PLAIN) return a;
        const T d = c - b;                                   // t - prox_cols(t)              (src/TV2Dopt.cpp:545-546)
        if (OP == LOP_DR_B) return a - (T(2) * d - c);       // Y - s, s = 2 (t - x) - t      (:411, :515)
        return a - d;                                        // final: s = t - x              (:427)
    }
    static __device__ __forceinline__ T out(T x, T b, T c) {
        if (OP == LOP_DR_B) return (c - b) + x;              // t' = 0.5 (t + s + 2 prox_rows(Y - s)) = (t - x_cols) + x_rows   (:419-422)
        return x;
    }
};

// ---------------------------------------------------------------- the scan loop, device form
// Same arithmetic and the same decisions as Lane<T>::run (lane_core.cuh, the form the CPU emulation executes), with positions
// carried pre-multiplied by the window row pitch so that a window access is ONE logic op + the shared-memory instruction:
//     address(row) = (((row * ROWB) & (W * ROWB - 1)) | lane offset) + window base
template <typename T> struct SmemIO { };
template <> struct SmemIO<double> {
    static __device__ __forceinline__ double ld(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
    static __device__ __forceinline__ void st(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
};
template <> struct SmemIO<float> {
    static __device__ __forceinline__ float ld(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
    static __device__ __forceinline__ void st(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
};
__device__ __forceinline__ void opaque(double& v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void opaque(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// value + flag byte, both under one predicate (kept as predicated stores: the compiler would otherwise branch around them)
__device__ __forceinline__ void sts_pred(uint32_t va, double v, uint32_t fa, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.shared.f64 [%0], %1;\n\t@q st.shared.u8 [%2], 1;\n\t}"
                 ::"r"(va), "d"(v), "r"(fa), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ void sts_pred(uint32_t va, float v, uint32_t fa, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.shared.f32 [%0], %1;\n\t@q st.shared.u8 [%2], 1;\n\t}"
                 ::"r"(va), "f"(v), "r"(fa), "r"((uint32_t)p) : "memory");
}

template <typename T, int W> struct DevWin {
    static constexpr int ROWB = LANES * (int)sizeof(T);          // bytes per window row
    static constexpr uint32_t MASK = (uint32_t)(W * ROWB - 1) & ~(uint32_t)(ROWB - 1);
    static constexpr int SH = (ROWB == 256) ? 5 : 5;             // (k * ROWB) >> SH == k * sizeof(T): 256 >> 5 = 8, 128 >> 5 = 4
    static constexpr int FSH = (ROWB == 256) ? 8 : 7;            // scaled position -> row
    uint32_t wbase;    // shared address of the warp's window (a kernel constant when the CTA is a single warp: folds into the access)
    uint32_t lane8;    // byte offset of this lane's column inside a window row
    uint32_t flg;      // shared address of this lane's flag bytes
    uint32_t rcp;      // shared address of the reciprocal table
    __device__ __forceinline__ uint32_t at(int pos) const {
        uint32_t o;                                    // (pos & MASK) | lane8 in one LOP3 (the two fields never overlap)
        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(o) : "r"((uint32_t)pos), "r"(MASK), "r"(lane8));
        return o + wbase;
    }
};

template <bool PH1, typename T, int W>
__device__ __forceinline__ void run_dev(Lane<T>& L, const DevWin<T, W> dw, const TaskGeom& g, T lam2, int niter) {
    using DW = DevWin<T, W>;
    constexpr int ROWB = DW::ROWB;
    T Z = L.Z, lo = L.lo, hi = L.hi;
    int ia = L.i * ROWB, la = L.last * ROWB, bloa = L.blo * ROWB, bhia = L.bhi * ROWB;
    int kind = L.kind, lprev = L.lprev * ROWB, kprev = L.kprev, in_ = L.in_rec;
    const int cea = g.ce * ROWB, csa = g.cs * ROWB;
    T nlam2 = -lam2;
    opaque(nlam2);                                     // keep -2 lam in a register (else it is re-negated every iteration)
#pragma unroll 1
    for (int it = 0; it < niter; it++) {
        const T y = SmemIO<T>::ld(dw.at(ia));
        const int ka = ia - la;
        const T r = SmemIO<T>::ld(dw.rcp + ((uint32_t)ka >> DW::SH));
        Z += y;
        const T cl = Z * r, ch = (Z + lam2) * r;
        const bool first = (ka == ROWB);
        const bool can = !first & (la < cea);
        const bool cbk = can & (lo > ch);
        const bool fbk = can & !cbk & (hi < cl);
        const bool brk = cbk | fbk;
        const int ea = cbk ? bloa : bhia;
        const T v = cbk ? lo : hi;
        const int fa = la + ROWB;
        // emit: the finished segment's value goes to its first (owned) row, plus the flag byte -- two predicated stores
        bool em = brk; int fea = fa;
        if (PH1) {
            em = brk & (ea >= csa);
            fea = fa > csa ? fa : csa;
            in_ = (em & (in_ == REC_NONE)) ? rec_pack(fa / ROWB, kind) : in_;
        }
        const uint32_t va = dw.at(fea);
        const uint32_t fla = dw.flg + (((uint32_t)fea >> DW::FSH) & (uint32_t)(W - 1));
        sts_pred(va, v, fla, em);
        const bool tlo = first | (cl >= lo), thi = first | (ch <= hi);
        lo = tlo ? cl : lo; bloa = tlo ? ia : bloa;
        hi = thi ? ch : hi; bhia = thi ? ia : bhia;
        lprev = brk ? la : lprev; kprev = brk ? kind : kprev;
        kind = cbk ? (int)LK_CEIL : (fbk ? (int)LK_FLOOR : kind);
        Z = cbk ? T(0) : (fbk ? nlam2 : Z);
        ia = (brk ? ea : ia) + ROWB;
        la = brk ? ea : la;
    }
    L.Z = Z; L.lo = lo; L.hi = hi; L.i = ia / ROWB; L.last = la / ROWB; L.blo = bloa / ROWB; L.bhi = bhia / ROWB;
    L.kind = kind; L.lprev = lprev / ROWB; L.kprev = kprev; L.in_rec = in_;
}

template <typename T, int W> struct DevEnv {
    Lane<T> L; int lane; DevWin<T, W> dw;
    template <class F> __device__ __forceinline__ void each(F f) { f(L, lane); }
    template <class F> __device__ __forceinline__ int rmin(F f) { return __reduce_min_sync(0xffffffffu, f(L, lane)); }
    template <class F> __device__ __forceinline__ int rmax(F f) { return __reduce_max_sync(0xffffffffu, f(L, lane)); }
    template <class F> __device__ __forceinline__ bool any(F f) { return __any_sync(0xffffffffu, f(L, lane)); }
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ void scan(Lane<T>& l, const Window<T, W>&, int, const TaskGeom& g, const T*, T lam2, bool ph1, int niter) {
        if (ph1) run_dev<true, T, W>(l, dw, g, lam2, niter); else run_dev<false, T, W>(l, dw, g, lam2, niter);
    }
};

// ---------------------------------------------------------------- STRIDED layout: feed and drain
template <typename T, int W, int RT, int OP> struct FeedStrided {
    static constexpr int R = RT;
    static constexpr int NST = 2;                       // staging tiles for the B / C operands of the fused forms
    static constexpr int MAXQ = (OP == LOP_PLAIN) ? W / RT : NST;
    static constexpr int NBAR = W / RT;
    const LaneArgs<T>* a; T* win; T* stB; T* stC; uint64_t* bar; int x0, z, q0, lane;
    __device__ __forceinline__ void request(DevEnv<T, W>&, int row0) {
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();                        // the slot's last generic-proxy accesses precede the async write
            const int q = row0 / R - q0;
            uint64_t* b = bar + (q % NBAR);
            mbar_expect_tx(b, (uint32_t)(R * LANES * sizeof(T) * (OP == LOP_PLAIN ? 1 : 3)));
            tma_load_3d(win + ((row0 & (W - 1)) << 5), &a->tmA, x0, row0, z, b);
            if (OP != LOP_PLAIN) {
                tma_load_3d(stB + (q % NST) * R * LANES, &a->tmB, x0, row0, z, b);
                tma_load_3d(stC + (q % NST) * R * LANES, &a->tmC, x0, row0, z, b);
            }
        }
    }
    __device__ __forceinline__ bool landed(DevEnv<T, W>&, int row0, bool block) {
        const int q = row0 / R - q0;
        uint64_t* b = bar + (q % NBAR);
        const uint32_t parity = (uint32_t)((q / NBAR) & 1);
        // the decision must be warp-uniform (it steers the task's control flow), and every lane needs the acquire of its own
        // successful wait before it reads the tile: all lanes poll, the vote decides
        bool ok = __all_sync(0xffffffffu, mbar_try(b, parity));
        if (!ok && !block) return false;
        while (!ok) ok = __all_sync(0xffffffffu, mbar_try(b, parity));
        if (OP != LOP_PLAIN) {
            const T* sb = stB + (q % NST) * R * LANES; const T* sc = stC + (q % NST) * R * LANES;
            T* wr = win + ((row0 & (W - 1)) << 5);
#pragma unroll
            for (int r = 0; r < R; r++) wr[r * LANES + lane] = PassOp<T, OP>::in(wr[r * LANES + lane], sb[r * LANES + lane], sc[r * LANES + lane]);
        }
        __syncwarp();
        return true;
    }
};

template <typename T, int W, int OP> struct DrainStrided {
    const T* __restrict__ B; const T* __restrict__ C; T* __restrict__ X; long long gbase, stride;     // gbase includes the lane
    __device__ __forceinline__ void rows8(const Window<T, W>&, int r0, int cnt, int, const T* xs, bool valid) {
        if (!valid) return;
        const long long g0 = gbase + (long long)r0 * stride;
        if (OP == LOP_DR_B) {
            T b[8], c[8];
#pragma unroll
            for (int u = 0; u < 8; u++) if (u < cnt) { b[u] = __ldg(B + g0 + u * stride); c[u] = __ldg(C + g0 + u * stride); }
#pragma unroll
            for (int u = 0; u < 8; u++) if (u < cnt) X[g0 + u * stride] = PassOp<T, OP>::out(xs[u], b[u], c[u]);
        } else {
#pragma unroll
            for (int u = 0; u < 8; u++) if (u < cnt) X[g0 + u * stride] = xs[u];
        }
    }
    __device__ __forceinline__ void flush(DevEnv<T, W>&, const Window<T, W>&, int, bool) {}
    __device__ __forceinline__ int hold(int, int) const { return 0x3fffffff; }
};

// ---------------------------------------------------------------- the kernel
template <typename T, int W, int RT, int OP> struct LaneSmem {
    static constexpr int NST = 2;
    static constexpr size_t win_bytes = (size_t)W * LANES * sizeof(T);
    static constexpr size_t stg_bytes = (OP == LOP_PLAIN) ? 0 : (size_t)2 * NST * RT * LANES * sizeof(T);
    static constexpr size_t flg_bytes = ((size_t)LANES * (W + 8) + 127) / 128 * 128;
    static constexpr size_t bar_bytes = 128;            // W / RT <= 16 barriers
    static constexpr size_t per_warp = win_bytes + stg_bytes + flg_bytes + bar_bytes;
    static constexpr size_t rcp_bytes = ((W + 2) * sizeof(T) + 127) / 128 * 128;
};

// repair: exact sequential continuation in global memory (rare; kept out of line so it does not cost the scan registers)
template <typename T, int OP>
__device__ __noinline__ int repair_fiber(const LaneArgs<T>* a, long long fiber, long long gbase, long long stride, long long nfp) {
    const ChunkPlan pl = a->plan;
    const int* rec = a->rec;
    const long long cstride = nfp;
    const T* A = a->A; const T* B = a->B; const T* C = a->C; T* X = a->X;
    return verify_repair_fiber<T>(pl, a->lam,
        [&](int c) { return rec[(0 * (long long)pl.nchunks + c) * cstride + fiber]; },
        [&](int c) { return rec[(1 * (long long)pl.nchunks + c) * cstride + fiber]; },
        [&](int c) { return rec[(2 * (long long)pl.nchunks + c) * cstride + fiber]; },
        [&](int r) { const long long g = gbase + (long long)r * stride;
                     return OP == LOP_PLAIN ? A[g] : PassOp<T, OP>::in(A[g], B[g], C[g]); },
        [&](int r, T v) { const long long g = gbase + (long long)r * stride;
                          X[g] = OP == LOP_DR_B ? PassOp<T, OP>::out(v, B[g], C[g]) : v; });
}

template <typename T, int W, int RT, int TITER, int OP, int NW>
__global__ void __launch_bounds__(NW * 32) k_lane_strided(const __grid_constant__ LaneArgs<T> a) {
    using SM = LaneSmem<T, W, RT, OP>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* rcp = reinterpret_cast<T*>(smem);
    unsigned char* wb = smem + SM::rcp_bytes + (size_t)warp * SM::per_warp;
    T* win = reinterpret_cast<T*>(wb);
    T* stB = reinterpret_cast<T*>(wb + SM::win_bytes);
    T* stC = stB + SM::NST * RT * LANES;
    uint8_t* flg = wb + SM::win_bytes + SM::stg_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(wb + SM::win_bytes + SM::stg_bytes + SM::flg_bytes);

    for (int k = threadIdx.x; k < W + 2; k += NW * 32) rcp[k] = k ? T(1) / T(k) : T(0);
    const long long task = (long long)blockIdx.x * NW + warp;
    const bool has_task = task < a.ntasks;
    if (has_task) {
        if (lane < W / RT) mbar_init(bar + lane, 1);
        for (int k = lane; k < (int)(SM::flg_bytes / 4); k += 32) reinterpret_cast<uint32_t*>(flg)[k] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                                    // the only CTA barrier: reciprocal table ready
    if (!has_task) return;

    const ChunkPlan pl = a.plan;
    const long long group = task / pl.nchunks;
    const int chunk = (int)(task - group * pl.nchunks);
    const int z = (int)(group / a.gps);
    const int x0 = (int)(group - (long long)z * a.gps) * LANES;
    const TaskGeom g = pl.geom(chunk);
    const bool valid = (long long)x0 + lane < a.per_slab;
    const long long gbase = (long long)z * a.inc * pl.n + x0 + lane;        // element (row 0) of this lane's fiber
    const long long nfp = (long long)a.slabs * a.gps * LANES;
    const long long fiber = group * LANES + lane;

    DevEnv<T, W> env; env.lane = lane; env.L.init(g, a.lam, valid);
    env.dw.wbase = s32(win); env.dw.lane8 = lane * (uint32_t)sizeof(T); env.dw.flg = s32(flg) + lane * (uint32_t)Window<T, W>::FP; env.dw.rcp = s32(rcp);
    Window<T, W> w{win, flg};
    FeedStrided<T, W, RT, OP> feed{&a, win, stB, stC, bar, x0, z, g.p0 / RT, lane};
    DrainStrided<T, W, OP> drain{a.B, a.C, a.X, gbase, a.inc};
    warp_task<T, W, TITER>(env, feed, drain, w, rcp, g, a.lam, TITER + RT, (TaskStats*)nullptr);

    // ---- chunk records; the last warp of the fiber group to finish verifies (and repairs) its 32 fibers ----
    int* rec = a.rec;
    rec[(0 * (long long)pl.nchunks + chunk) * nfp + fiber] = env.L.in_rec;
    rec[(1 * (long long)pl.nchunks + chunk) * nfp + fiber] = env.L.out_rec;
    rec[(2 * (long long)pl.nchunks + chunk) * nfp + fiber] = env.L.retired ? env.L.ovf_rec : REC_NONE;
    const bool any_retired = __any_sync(0xffffffffu, env.L.retired);
    if (pl.nchunks == 1 && !any_retired) return;
    __threadfence();
    __syncwarp();
    int old = 0;
    if (lane == 0) old = atomicAdd(a.group_count + group, 1);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != pl.nchunks - 1) return;
    __threadfence();
    if (lane == 0) a.group_count[group] = 0;
    // quick test first: every entry record equals the predecessor's exit record and nobody retired
    bool bad = false;
    if (valid) {
        for (int c = 0; c < pl.nchunks; c++) {
            const int ri = __ldcg(rec + (0 * (long long)pl.nchunks + c) * nfp + fiber);
            const int ro = c > 0 ? __ldcg(rec + (1 * (long long)pl.nchunks + c - 1) * nfp + fiber) : ri;
            const int rv = __ldcg(rec + (2 * (long long)pl.nchunks + c) * nfp + fiber);
            bad |= (c > 0 && ri != ro) || rv != REC_NONE;
        }
    }
    if (bad) {
        const int n = repair_fiber<T, OP>(&a, fiber, gbase, a.inc, nfp);
        if (n) atomicAdd(a.stats, (unsigned long long)n);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 3D view (d0 contiguous) of `base`; box (b0, b1, 1)
template <typename T>
static bool make_map(CUtensorMap* m, const T* base, long long d0, long long d1, long long d2, int b0, int b1, CUtensorMapSwizzle sw) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)d0 * sizeof(T), (cuuint64_t)d0 * (cuuint64_t)d1 * sizeof(T)};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return fn(m, dt, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct LaneTuning { int clen, halo, variant; };
static LaneTuning g_tune = {0, 32, 0};
void lane_set_tuning(int clen, int halo, int variant) { g_tune.clen = clen; g_tune.halo = halo; g_tune.variant = variant; }

long long lane_scratch_bytes(long long nf, int len) {
    // records for the finest chunking the launcher may choose (chunks of >= 64 rows) + group counters + stats
    const long long nfp = (nf + 31) / 32 * 32 + 32 * 64;
    const long long maxchunks = len / 64 + 2;
    return 3 * maxchunks * nfp * 4 + (nfp / 32 + 128) * 4 + 64;
}

// launch one variant; with nchunks <= 0 only report how many warp tasks the device can hold at once
template <typename T, int W, int RT, int TITER, int OP, int NW>
static cudaError_t launch_strided_v(LaneArgs<T>& a, cudaStream_t st, int* slots) {
    using SM = LaneSmem<T, W, RT, OP>;
    auto kern = k_lane_strided<T, W, RT, TITER, OP, NW>;
    const size_t smem = SM::rcp_bytes + (size_t)NW * SM::per_warp;
    static int s_slots = 0;                      // per instantiation: resident warps on the current device
    if (!s_slots) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
        if (e != cudaSuccess) return e;
        s_slots = sms * per_sm * NW;
    }
    if (slots) { *slots = s_slots; return cudaSuccess; }
    const unsigned blocks = (unsigned)((a.ntasks + NW - 1) / NW);
    kern<<<blocks, NW * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <typename T, int OP>
static cudaError_t launch_strided_variant(int variant, LaneArgs<T>& a, cudaStream_t st, int* slots) {
    switch (variant) {
        case 1: return launch_strided_v<T, 32, 8, 8, OP, 4>(a, st, slots);
        case 2: return launch_strided_v<T, 64, 8, 8, OP, 4>(a, st, slots);
        case 3: return launch_strided_v<T, 64, 8, 16, OP, 2>(a, st, slots);
        case 4: return launch_strided_v<T, 64, 8, 16, OP, 1>(a, st, slots);
        case 5: return launch_strided_v<T, 32, 8, 8, OP, 1>(a, st, slots);
        case 6: return launch_strided_v<T, 32, 8, 16, OP, 1>(a, st, slots);
        default: return launch_strided_v<T, 64, 8, 16, OP, 4>(a, st, slots);
    }
}

// x = prox over fibers with element stride inc > 1 (fibers adjacent in memory).  scratch: lane_scratch_bytes(), zero-initialised
// once (the group counters reset themselves).  Returns cudaErrorInvalidConfiguration when the shape does not suit TMA tiling.
template <typename T>
cudaError_t lane_prox_strided(int op, const T* A, const T* B, const T* C, T* X, long long nf, int len, long long inc, T lam, void* scratch,
                              cudaStream_t st) {
    if (inc <= 1 || nf % inc != 0) return cudaErrorInvalidConfiguration;
    if ((inc * sizeof(T)) % 16 != 0 || ((uintptr_t)A & 15) || (B && ((uintptr_t)B & 15)) || (C && ((uintptr_t)C & 15)))
        return cudaErrorInvalidConfiguration;
    if (len < 2 || !(lam > T(0))) return cudaErrorInvalidConfiguration;
    LaneArgs<T> a;
    memset(&a, 0, sizeof(a));
    const long long slabs = nf / inc;
    if (slabs > 0x7fffffff) return cudaErrorInvalidConfiguration;
    const int RT = 8;
    if (!make_map<T>(&a.tmA, A, inc, len, slabs, LANES, RT, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidConfiguration;
    if (op != LOP_PLAIN) {
        if (!make_map<T>(&a.tmB, B, inc, len, slabs, LANES, RT, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidConfiguration;
        if (!make_map<T>(&a.tmC, C, inc, len, slabs, LANES, RT, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidConfiguration;
    }
    a.A = A; a.B = B; a.C = C; a.X = X; a.inc = inc; a.per_slab = inc; a.slabs = (int)slabs; a.gps = (int)((inc + LANES - 1) / LANES);
    a.lam = lam;
    // chunking: whole fibers when there are enough of them to fill the machine; else as many chunks (multiples of 16 rows) as
    // there are resident warp slots, so that the whole pass is ONE wave of warp tasks
    const long long groups = slabs * a.gps;
    int slots = 0;
    cudaError_t e = op == LOP_PLAIN ? launch_strided_variant<T, LOP_PLAIN>(g_tune.variant, a, st, &slots)
                  : op == LOP_DR_B ? launch_strided_variant<T, LOP_DR_B>(g_tune.variant, a, st, &slots)
                                   : launch_strided_variant<T, LOP_DR_B_FINAL>(g_tune.variant, a, st, &slots);
    if (e != cudaSuccess) return e;
    int clen = g_tune.clen, halo = g_tune.halo;
    if (clen <= 0) {
        long long c = groups >= slots ? 1 : slots / groups;
        clen = (int)((len + c - 1) / c);
        if (clen < 64) clen = 64;
    }
    clen = (clen + 15) / 16 * 16; halo = (halo + 7) / 8 * 8;
    a.plan.n = len; a.plan.halo = halo;
    if (clen >= len) { a.plan.clen = len; a.plan.nchunks = 1; } else { a.plan.clen = clen; a.plan.nchunks = (len + clen - 1) / clen; }
    if (a.plan.nchunks > len / 64 + 2) return cudaErrorInvalidConfiguration;
    // scratch layout (fixed places, whatever the chunking): [stats 64 B][group counters, zero between launches][records]
    a.stats = (unsigned long long*)scratch;
    a.group_count = (int*)((char*)scratch + 64);
    a.rec = a.group_count + ((groups + 63) / 64) * 64;
    a.ntasks = groups * a.plan.nchunks;
    return op == LOP_PLAIN ? launch_strided_variant<T, LOP_PLAIN>(g_tune.variant, a, st, nullptr)
         : op == LOP_DR_B ? launch_strided_variant<T, LOP_DR_B>(g_tune.variant, a, st, nullptr)
                          : launch_strided_variant<T, LOP_DR_B_FINAL>(g_tune.variant, a, st, nullptr);
}

// per-device scratch for the records (grow-only, zero-initialised: the group counters must start at 0 and reset themselves)
struct LaneScratch { void* p = nullptr; size_t cap = 0; };
static LaneScratch g_scr[64];
void* lane_scratch(long long nf, int len) {
    int d = 0; if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return nullptr;
    const size_t need = (size_t)lane_scratch_bytes(nf, len);
    LaneScratch& s = g_scr[d];
    if (need > s.cap) {
        if (s.p) { cudaDeviceSynchronize(); cudaFree(s.p); s.p = nullptr; s.cap = 0; }
        if (cudaMalloc(&s.p, need) != cudaSuccess) { cudaGetLastError(); s.p = nullptr; return nullptr; }
        cudaMemset(s.p, 0, need); s.cap = need;
    }
    return s.p;
}
unsigned long long lane_read_stats(int reset) {
    int d = 0; if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64 || !g_scr[d].p) return 0;
    unsigned long long v = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&v, g_scr[d].p, sizeof(v), cudaMemcpyDeviceToHost);
    if (reset) cudaMemset(g_scr[d].p, 0, sizeof(v));
    return v;
}

template cudaError_t lane_prox_strided<double>(int, const double*, const double*, const double*, double*, long long, int, long long, double,
                                               void*, cudaStream_t);
template cudaError_t lane_prox_strided<float>(int, const float*, const float*, const float*, float*, long long, int, long long, float, void*,
                                              cudaStream_t);

}  // namespace ptvl
