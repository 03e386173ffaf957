// kernels_chunked.cu -- chunked speculative TV-L1 prox: many lanes per fiber, exact stitching (phases in chunk_core.cuh).
//
// k_prox_chunked_contig: contiguous fibers (element stride 1).  A CTA stages `fpb` whole fibers in shared memory
// (chunk rows padded by 16 B so lanes that walk their chunks in lock-step do not all hit one bank), runs
// walk-own-chunk -> rounds -> fill with one lane per 32-sample chunk, and streams the fused result back with coalesced
// stores.  HBM traffic per launch: every input array read once, the output written once.
#include "ptv_internal.h"
#include "chunk_core.cuh"

namespace ptv {

template <typename T> struct PadCfg { static constexpr int PADE = 16 / (int)sizeof(T); };   // 16 bytes of padding per chunk row

template <typename T> struct SmemLd {
    const T* p;
    PTV_HD T operator()(int j) const { return p[j + (j >> 5) * PadCfg<T>::PADE]; }
};
template <typename T> struct SmemSt {
    T* p;
    PTV_HD void operator()(int j, T v) const { p[j + (j >> 5) * PadCfg<T>::PADE] = v; }
};

template <typename T, bool WEIGHTED>
__global__ void k_prox_chunked_contig(const T* __restrict__ A, const T* __restrict__ B, int in_op, T* __restrict__ X, int out_op,
                                      long long nf, int n, T lam, const T* __restrict__ lamv, int lpf, int fpb, int npad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* ys = reinterpret_cast<T*>(smem_raw);                       // staged input            [fpb][npad]
    T* vs = ys + (size_t)fpb * npad;                              // sparse segment values   [fpb][npad]
    T* wsm = vs + (size_t)fpb * npad;                             // per-edge weights        [fpb][npad] (weighted only)
    T* rcp = wsm + (WEIGHTED ? (size_t)fpb * npad : 0);           // reciprocal table        [RCP_N]
    uint32_t* mk = reinterpret_cast<uint32_t*>(rcp + RCP_N);      // masks P, K0, K1         [3][fpb][lpf]
    int* carry = reinterpret_cast<int*>(mk + (size_t)3 * fpb * lpf);   // [fpb][lpf]
    constexpr int PADE = PadCfg<T>::PADE;
    const int tid = threadIdx.x;
    const long long f0 = (long long)blockIdx.x * fpb;
    const int nfib = (int)((nf - f0) < fpb ? (nf - f0) : fpb);

    // ---- stage the fibers: coalesced reads, input op applied on the fly ----
    if (tid < RCP_N) rcp[tid] = tid ? T(1) / T(tid) : T(0);
    for (int fb = 0; fb < nfib; fb++) {
        const long long base = (f0 + fb) * (long long)n;
        T* row = ys + (size_t)fb * npad;
        for (int j = tid; j < n; j += blockDim.x) {
            T a = A[base + j];
            if (in_op == IN_A_MINUS_B) a = a - B[base + j];
            else if (in_op == IN_A_PLUS_B) a = a + B[base + j];
            row[j + (j >> 5) * PADE] = a;
        }
        if (WEIGHTED) {
            const long long wbase = (f0 + fb) * (long long)(n - 1);
            T* wrow = wsm + (size_t)fb * npad;
            for (int j = tid; j < n; j += blockDim.x) wrow[j + (j >> 5) * PADE] = (j < n - 1) ? lamv[wbase + j] : T(0);
        }
    }
    __syncthreads();

    const int nchunks = (n + CH - 1) / CH;
    const int fb = tid / lpf, q = tid - fb * lpf;
    const bool lane_ok = fb < nfib && q < nchunks;
    const int fbc = lane_ok ? fb : 0;
    ChunkMasks m{mk + (size_t)fbc * lpf, mk + (size_t)(fpb + fbc) * lpf, mk + (size_t)(2 * fpb + fbc) * lpf};
    SmemLd<T> y{ys + (size_t)fbc * npad};
    SmemSt<T> stv{vs + (size_t)fbc * npad};
    RcpDiv<T> div{rcp};
    LaneState<T> st;
    st.active = false; st.finished = false; st.pend_a = -1; st.pend_k = K_NONE;

    auto phases = [&](auto lamf) {
        bool act = lane_ok ? walk_chunk<T>(q, 0, nchunks, n, y, stv, lamf, div, st, m) : false;
        for (int r = 1; __syncthreads_or(act ? 1 : 0); r++)
            act = lane_ok ? walk_chunk<T>(q, r, nchunks, n, y, stv, lamf, div, st, m) : false;
    };
    if (WEIGHTED) phases(ArrayLam<T, SmemLd<T>>{SmemLd<T>{wsm + (size_t)fbc * npad}});
    else phases(UniformLam<T>{lam});

    // ---- carry: start of the segment that covers each chunk's first sample ----
    int* cr = carry + (size_t)fbc * lpf;
    if (lane_ok) {
        int c2 = q - 1;
        while (c2 >= 0 && m.P[c2] == 0) c2--;
        cr[q] = (c2 >= 0) ? c2 * CH + high_bit(m.P[c2]) : 0;
    }
    __syncthreads();

    // ---- fill + stream the result back: out[j] = f(y[j], value of the segment covering j), coalesced ----
    for (int fb2 = 0; fb2 < nfib; fb2++) {
        const long long base = (f0 + fb2) * (long long)n;
        const T* yrow = ys + (size_t)fb2 * npad;
        const T* vrow = vs + (size_t)fb2 * npad;
        ChunkMasks m2{mk + (size_t)fb2 * lpf, nullptr, nullptr};
        const int* cr2 = carry + (size_t)fb2 * lpf;
        for (int j = tid; j < n; j += blockDim.x) {
            const int a = seg_start_of(j, m2, cr2);
            const T x = vrow[a + (a >> 5) * PADE];
            X[base + j] = apply_out<T>(out_op, yrow[j + (j >> 5) * PADE], x);
        }
    }
}

// Returns cudaErrorInvalidConfiguration if the fibers do not fit in shared memory (caller falls back to the sequential kernel).
template <typename T>
cudaError_t prox_fibers_chunked_contig(const T* A, const T* B, InOp op, T* X, int out_op, FiberGeom g, T lam, const T* lamv,
                                       cudaStream_t st) {
    if (g.inc != 1) return cudaErrorInvalidConfiguration;
    if (g.nf <= 0 || g.len <= 0) return cudaSuccess;
    const int n = g.len;
    const int lpf = (n + CH - 1) / CH;
    if (lpf > 1024) return cudaErrorInvalidConfiguration;
    const int npad = n + (lpf + 1) * PadCfg<T>::PADE;
    const size_t per_fiber = (size_t)npad * sizeof(T) * (lamv ? 3 : 2) + (size_t)lpf * 16;
    int fpb = 128 / lpf; if (fpb < 1) fpb = 1;
    if ((long long)fpb > g.nf) fpb = (int)g.nf;
    const size_t smem = per_fiber * fpb + RCP_N * sizeof(T) + 16;
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    int threads = ((fpb * lpf + 31) / 32) * 32;
    if (threads < RCP_N) threads = RCP_N;
    unsigned blocks = (unsigned)((g.nf + fpb - 1) / fpb);
    auto kern = lamv ? k_prox_chunked_contig<T, true> : k_prox_chunked_contig<T, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<blocks, threads, smem, st>>>(A, B, (int)op, X, out_op, g.nf, n, lam, lamv, lpf, fpb, npad);
    return cudaGetLastError();
}

template cudaError_t prox_fibers_chunked_contig<double>(const double*, const double*, InOp, double*, int, FiberGeom, double,
                                                        const double*, cudaStream_t);
template cudaError_t prox_fibers_chunked_contig<float>(const float*, const float*, InOp, float*, int, FiberGeom, float,
                                                       const float*, cudaStream_t);

}  // namespace ptv
