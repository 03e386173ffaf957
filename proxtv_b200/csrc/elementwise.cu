// elementwise.cu -- streaming helpers of the outer loops (unfused forms; the chunked kernels fuse most of these).
//
// Arithmetic is written in the reference's operation order so that, given identical prox outputs, every intermediate is
// bit-identical to the reference's (src/TV2Dopt.cpp:390-430, :212-277; src/TVNDopt.cpp:212-227).  The only exception is
// the two sums (image mean, stop criterion), which are tree reductions with a fixed, deterministic order instead of the
// reference's serial / OpenMP-reduction order.
#include "ptv_internal.h"
#include <stdint.h>

namespace ptv {

static inline unsigned grid_for(long long n, int threads, int per_thread = 4) {
    long long b = (n + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
    if (b < 1) b = 1;
    if (b > 148LL * 32) b = 148LL * 32;
    return (unsigned)b;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sm[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sm[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (w == 0) for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    return v;   // valid in thread 0
}

// ---- per-image 2*mean (DR initialisation, src/TV2Dopt.cpp:390-395) ----
template <typename T>
__global__ void k_image_partial(const T* __restrict__ Y, long long per_image, double* __restrict__ partial) {
    const T* img = Y + (long long)blockIdx.y * per_image;
    double acc = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < per_image; q += (long long)gridDim.x * blockDim.x)
        acc += (double)img[q];
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
template <typename T>
__global__ void k_image_fill(T* __restrict__ t, long long per_image, const double* __restrict__ partial, int nblk) {
    __shared__ double mean2;
    double acc = 0;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) acc += partial[(long long)blockIdx.y * nblk + q];
    acc = block_sum(acc);
    if (threadIdx.x == 0) mean2 = 2 * acc / (double)per_image;
    __syncthreads();
    T v = (T)mean2;
    T* img = t + (long long)blockIdx.y * per_image;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < per_image; q += (long long)gridDim.x * blockDim.x)
        img[q] = v;
}
template <typename T>
cudaError_t ew_image_means_x2(const T* Y, long long per_image, int batch, T* t, double* scratch, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 2 * ((batch + 0) > 0 ? 1 : 0), st);
    int nblk = (int)grid_for(per_image, 256, 8);
    if (nblk > REDUCE_BLOCKS) nblk = REDUCE_BLOCKS;
    // scratch holds REDUCE_BLOCKS doubles: split them between the images of one launch
    int share = REDUCE_BLOCKS / (batch < REDUCE_BLOCKS ? batch : REDUCE_BLOCKS);
    if (nblk > share) nblk = share < 1 ? 1 : share;
    int per_launch = REDUCE_BLOCKS / nblk; if (per_launch < 1) per_launch = 1;
    for (int b0 = 0; b0 < batch; b0 += per_launch) {
        int nb = batch - b0 < per_launch ? batch - b0 : per_launch;
        dim3 g(nblk, nb);
        k_image_partial<T><<<g, 256, 0, st>>>(Y + (long long)b0 * per_image, per_image, scratch);
        k_image_fill<T><<<g, 256, 0, st>>>(t + (long long)b0 * per_image, per_image, scratch, nblk);
    }
    return cudaGetLastError();
}

// ---- Douglas-Rachford elementwise steps ----
template <typename T> __global__ void k_dr_reflect_cols(const T* __restrict__ t, const T* __restrict__ x, T* __restrict__ s, long long n) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        T d = t[q] - x[q];                 // DR_proxDiff: in - prox(in)            (:545-546)
        s[q] = T(2) * d - t[q];            // reflection                            (:411)
    }
}
template <typename T> __global__ void k_dr_combine_rows(const T* __restrict__ Y, const T* __restrict__ s, const T* __restrict__ x,
                                                       T* __restrict__ t, long long n) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        T in = Y[q] - s[q];                // :515
        T tb = Y[q] - (in - x[q]);         // :520 with :546
        tb = T(2) * tb - s[q];             // :419
        t[q] = T(0.5) * (t[q] + tb);       // :422
    }
}
template <typename T> __global__ void k_dr_final_cols(const T* __restrict__ t, const T* __restrict__ x, T* __restrict__ s, long long n) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        s[q] = t[q] - x[q];                // :427
}
template <typename T> __global__ void k_dr_final_rows(const T* __restrict__ Y, const T* __restrict__ s, const T* __restrict__ x,
                                                     T* __restrict__ out, long long n) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        T in = Y[q] - s[q];
        T tb = Y[q] - (in - x[q]);         // :429
        out[q] = tb - s[q];                // :430
    }
}
template <typename T> __global__ void k_dual_update(T* __restrict__ p, const T* __restrict__ a, const T* __restrict__ b, long long n) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        p[q] += a[q] - b[q];               // src/TV2Dopt.cpp:213,:263
}

template <typename T> cudaError_t ew_dr_reflect_cols(const T* t, const T* x, T* s, long long n, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 1, st);
    k_dr_reflect_cols<T><<<grid_for(n, 256), 256, 0, st>>>(t, x, s, n); return cudaGetLastError(); }
template <typename T> cudaError_t ew_dr_combine_rows(const T* Y, const T* s, const T* x, T* t, long long n, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 1, st);
    k_dr_combine_rows<T><<<grid_for(n, 256), 256, 0, st>>>(Y, s, x, t, n); return cudaGetLastError(); }
template <typename T> cudaError_t ew_dr_final_cols(const T* t, const T* x, T* s, long long n, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 1, st);
    k_dr_final_cols<T><<<grid_for(n, 256), 256, 0, st>>>(t, x, s, n); return cudaGetLastError(); }
template <typename T> cudaError_t ew_dr_final_rows(const T* Y, const T* s, const T* x, T* out, long long n, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 1, st);
    k_dr_final_rows<T><<<grid_for(n, 256), 256, 0, st>>>(Y, s, x, out, n); return cudaGetLastError(); }
template <typename T> cudaError_t ew_dual_update(T* p, const T* a, const T* b, long long n, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 1, st);
    k_dual_update<T><<<grid_for(n, 256), 256, 0, st>>>(p, a, b, n); return cudaGetLastError(); }

// ---- first DR pass on a constant image: s = 2 (t - x1[image][k]) - t, k = position of the sample along its axis-0 fiber ----
template <typename T> __global__ void k_dr_reflect_bcast(const T* __restrict__ t, const T* __restrict__ x1, T* __restrict__ s, long long n,
                                                        long long per_image, int len, long long inc) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const long long b = q / per_image, r = q - b * per_image;
        const int k = (int)((r / inc) % len);
        const T d = t[q] - x1[b * len + k];
        s[q] = T(2) * d - t[q];
    }
}
template <typename T> cudaError_t ew_dr_reflect_bcast(const T* t, const T* x1, T* s, long long n, long long per_image, int len,
                                                      long long inc, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 1, st);
    k_dr_reflect_bcast<T><<<grid_for(n, 256), 256, 0, st>>>(t, x1, s, n, per_image, len, inc); return cudaGetLastError(); }

// ---- stop criterion: mean |a - b|  (src/TV2Dopt.cpp:273-277) ----
template <typename T> __global__ void k_absdiff_partial(const T* __restrict__ a, const T* __restrict__ b, long long n, double* __restrict__ partial) {
    double acc = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        acc += fabs((double)a[q] - (double)b[q]);
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void k_final_mean(const double* __restrict__ partial, int nblk, long long n, double* __restrict__ result) {
    double acc = 0;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) acc += partial[q];
    acc = block_sum(acc);
    if (threadIdx.x == 0) *result = acc / (double)n;
}
template <typename T> cudaError_t ew_mean_abs_diff(const T* a, const T* b, long long n, double* scratch, double* result, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 2, st);
    int nblk = (int)grid_for(n, 256, 8); if (nblk > REDUCE_BLOCKS) nblk = REDUCE_BLOCKS;
    k_absdiff_partial<T><<<nblk, 256, 0, st>>>(a, b, n, scratch);
    k_final_mean<<<1, 256, 0, st>>>(scratch, nblk, n, result);
    return cudaGetLastError();
}

// ---- parallel proximal Dykstra combine (src/TVNDopt.cpp:212-227): x = sum_i p_i/k ; z_i += x - p_i ; stop = mean|x - x_old| ----
template <typename T> __global__ void k_pd_combine(T* const* __restrict__ p, T* const* __restrict__ z, int k, T* __restrict__ x,
                                                  long long n, double* __restrict__ partial) {
    double acc = 0;
    const T kk = (T)k;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        T xo = x[q], xn = T(0);
        for (int i = 0; i < k; i++) xn += p[i][q] / kk;
        for (int i = 0; i < k; i++) z[i][q] += xn - p[i][q];
        x[q] = xn;
        acc += fabs((double)xn - (double)xo);
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
// Same arithmetic, K known at compile time, the 2K array pointers passed by value and 16-byte vector accesses: the generic kernel
// above chases 2K pointers through memory and moves 4-8 bytes per access (measured 1474 us for 3 x 67M f32, 3.2x the HBM time).
template <typename T, int K> struct PdPtrs { T* p[K]; T* z[K]; };
template <typename T, int K>
__global__ void __launch_bounds__(256) k_pd_combine_vec(PdPtrs<T, K> a, T* __restrict__ x, long long n, double* __restrict__ partial) {
    constexpr int V = 16 / (int)sizeof(T);
    struct alignas(16) Vec { T v[V]; };
    double acc = 0;
    const T kk = (T)K;
    const long long nv = n / V;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nv; q += (long long)gridDim.x * blockDim.x) {
        Vec xo = reinterpret_cast<const Vec*>(x)[q], pv[K], zv[K], xn;
#pragma unroll
        for (int i = 0; i < K; i++) { pv[i] = reinterpret_cast<const Vec*>(a.p[i])[q]; zv[i] = reinterpret_cast<const Vec*>(a.z[i])[q]; }
#pragma unroll
        for (int e = 0; e < V; e++) {
            T s = T(0);
#pragma unroll
            for (int i = 0; i < K; i++) s += pv[i].v[e] / kk;
#pragma unroll
            for (int i = 0; i < K; i++) zv[i].v[e] += s - pv[i].v[e];
            xn.v[e] = s;
            acc += fabs((double)s - (double)xo.v[e]);
        }
#pragma unroll
        for (int i = 0; i < K; i++) reinterpret_cast<Vec*>(a.z[i])[q] = zv[i];
        reinterpret_cast<Vec*>(x)[q] = xn;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)                       // the < V trailing elements
        for (long long q = nv * V; q < n; q++) {
            T xo = x[q], s = T(0);
            for (int i = 0; i < K; i++) s += a.p[i][q] / kk;
            for (int i = 0; i < K; i++) a.z[i][q] += s - a.p[i][q];
            x[q] = s; acc += fabs((double)s - (double)xo);
        }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
template <typename T, int K>
static void launch_pd_vec(T* const* hp, T* const* hz, T* x, long long n, double* scratch, int nblk, cudaStream_t st) {
    PdPtrs<T, K> a;
    for (int i = 0; i < K; i++) { a.p[i] = hp[i]; a.z[i] = hz[i]; }
    k_pd_combine_vec<T, K><<<nblk, 256, 0, st>>>(a, x, n, scratch);
}

// dp, dz: device copies of the pointer arrays (generic kernel); hp, hz: the same pointers on the host (vector kernel, K <= 4)
template <typename T> cudaError_t ew_pd_combine(T* const* dp, T* const* dz, T* const* hp, T* const* hz, int k, T* x, long long n,
                                                double* scratch, double* result, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 2, st);
    bool aligned = (((uintptr_t)x) & 15) == 0;
    for (int i = 0; i < k && aligned; i++) aligned = ((((uintptr_t)hp[i]) | ((uintptr_t)hz[i])) & 15) == 0;
    int nblk;
    if (aligned && k >= 1 && k <= 4) {
        nblk = (int)grid_for(n, 256, 16 / (int)sizeof(T) * 2); if (nblk > REDUCE_BLOCKS) nblk = REDUCE_BLOCKS;
        switch (k) {
            case 1: launch_pd_vec<T, 1>(hp, hz, x, n, scratch, nblk, st); break;
            case 2: launch_pd_vec<T, 2>(hp, hz, x, n, scratch, nblk, st); break;
            case 3: launch_pd_vec<T, 3>(hp, hz, x, n, scratch, nblk, st); break;
            default: launch_pd_vec<T, 4>(hp, hz, x, n, scratch, nblk, st); break;
        }
    } else {
        nblk = (int)grid_for(n, 256, 4); if (nblk > REDUCE_BLOCKS) nblk = REDUCE_BLOCKS;
        k_pd_combine<T><<<nblk, 256, 0, st>>>(dp, dz, k, x, n, scratch);
    }
    k_final_mean<<<1, 256, 0, st>>>(scratch, nblk, n, result);
    return cudaGetLastError();
}

// ---- parallel Douglas-Rachford combine (src/TVNDopt.cpp:456-475): q = sum_i p_i/k ; x = sum_i z_i/k (both accumulated term by
//      term) ; z_i += 2 q - x - p_i ; stop = mean|x - x_old| ----
template <typename T> __global__ void k_pdr_combine(T* const* __restrict__ p, T* const* __restrict__ z, int k, T* __restrict__ x,
                                                   long long n, double* __restrict__ partial) {
    double acc = 0;
    const T kk = (T)k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        T xo = x[e], xn = T(0), q = T(0);
        for (int i = 0; i < k; i++) { q += p[i][e] / kk; xn += z[i][e] / kk; }
        for (int i = 0; i < k; i++) z[i][e] += T(2) * q - xn - p[i][e];
        x[e] = xn;
        acc += fabs((double)xn - (double)xo);
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
template <typename T, int K>
__global__ void __launch_bounds__(256) k_pdr_combine_vec(PdPtrs<T, K> a, T* __restrict__ x, long long n, double* __restrict__ partial) {
    constexpr int V = 16 / (int)sizeof(T);
    struct alignas(16) Vec { T v[V]; };
    double acc = 0;
    const T kk = (T)K;
    const long long nv = n / V;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nv; e += (long long)gridDim.x * blockDim.x) {
        Vec xo = reinterpret_cast<const Vec*>(x)[e], pv[K], zv[K], xn;
#pragma unroll
        for (int i = 0; i < K; i++) { pv[i] = reinterpret_cast<const Vec*>(a.p[i])[e]; zv[i] = reinterpret_cast<const Vec*>(a.z[i])[e]; }
#pragma unroll
        for (int c = 0; c < V; c++) {
            T q = T(0), s = T(0);
#pragma unroll
            for (int i = 0; i < K; i++) { q += pv[i].v[c] / kk; s += zv[i].v[c] / kk; }
#pragma unroll
            for (int i = 0; i < K; i++) zv[i].v[c] += T(2) * q - s - pv[i].v[c];
            xn.v[c] = s;
            acc += fabs((double)s - (double)xo.v[c]);
        }
#pragma unroll
        for (int i = 0; i < K; i++) reinterpret_cast<Vec*>(a.z[i])[e] = zv[i];
        reinterpret_cast<Vec*>(x)[e] = xn;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)                       // the < V trailing elements
        for (long long e = nv * V; e < n; e++) {
            T xo = x[e], s = T(0), q = T(0);
            for (int i = 0; i < K; i++) { q += a.p[i][e] / kk; s += a.z[i][e] / kk; }
            for (int i = 0; i < K; i++) a.z[i][e] += T(2) * q - s - a.p[i][e];
            x[e] = s; acc += fabs((double)s - (double)xo);
        }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
template <typename T, int K>
static void launch_pdr_vec(T* const* hp, T* const* hz, T* x, long long n, double* scratch, int nblk, cudaStream_t st) {
    PdPtrs<T, K> a;
    for (int i = 0; i < K; i++) { a.p[i] = hp[i]; a.z[i] = hz[i]; }
    k_pdr_combine_vec<T, K><<<nblk, 256, 0, st>>>(a, x, n, scratch);
}
template <typename T> cudaError_t ew_pdr_combine(T* const* dp, T* const* dz, T* const* hp, T* const* hz, int k, T* x, long long n,
                                                 double* scratch, double* result, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 2, st);
    bool aligned = (((uintptr_t)x) & 15) == 0;
    for (int i = 0; i < k && aligned; i++) aligned = ((((uintptr_t)hp[i]) | ((uintptr_t)hz[i])) & 15) == 0;
    int nblk;
    if (aligned && k >= 1 && k <= 4) {
        nblk = (int)grid_for(n, 256, 16 / (int)sizeof(T) * 2); if (nblk > REDUCE_BLOCKS) nblk = REDUCE_BLOCKS;
        switch (k) {
            case 1: launch_pdr_vec<T, 1>(hp, hz, x, n, scratch, nblk, st); break;
            case 2: launch_pdr_vec<T, 2>(hp, hz, x, n, scratch, nblk, st); break;
            case 3: launch_pdr_vec<T, 3>(hp, hz, x, n, scratch, nblk, st); break;
            default: launch_pdr_vec<T, 4>(hp, hz, x, n, scratch, nblk, st); break;
        }
    } else {
        nblk = (int)grid_for(n, 256, 4); if (nblk > REDUCE_BLOCKS) nblk = REDUCE_BLOCKS;
        k_pdr_combine<T><<<nblk, 256, 0, st>>>(dp, dz, k, x, n, scratch);
    }
    k_final_mean<<<1, 256, 0, st>>>(scratch, nblk, n, result);
    return cudaGetLastError();
}
// x = y / k (src/TVNDopt.cpp:367)
template <typename T> __global__ void k_div_scalar(const T* __restrict__ y, T* __restrict__ x, long long n, T kk) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) x[e] = y[e] / kk;
}
template <typename T> cudaError_t ew_div_scalar(const T* y, T* x, long long n, T k, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 1, st);
    int nblk = (int)grid_for(n, 256, 8); if (nblk > REDUCE_BLOCKS) nblk = REDUCE_BLOCKS;
    k_div_scalar<T><<<nblk, 256, 0, st>>>(y, x, n, k);
    return cudaGetLastError();
}

// first Douglas-Rachford half-iteration on the constant start image t (2 * mean): the column prox of a constant is that constant, so
// d = t - x_cols and u = Y - (2 d - t) need no scan; written with the very expressions the lane drain uses (PassOp<LOP_DRA>)
template <typename T> __global__ void k_dr_first(const T* Y, const T* __restrict__ t, T* __restrict__ U, T* D, long long n) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const T c = t[e], d = c - c, y = Y[e];            // Y may be D (in place): read before the write
        D[e] = d; U[e] = y - (T(2) * d - c);
    }
}
template <typename T> cudaError_t ew_dr_first(const T* Y, const T* t, T* U, T* D, long long n, cudaStream_t st) {
    KernelSpan span(KC_ELEMENTWISE, 1, st);
    int nblk = (int)grid_for(n, 256, 8); if (nblk > REDUCE_BLOCKS) nblk = REDUCE_BLOCKS;
    k_dr_first<T><<<nblk, 256, 0, st>>>(Y, t, U, D, n);
    return cudaGetLastError();
}

// mean |y[e + 1] - y[e]| over n consecutive elements: one block, fixed summation order (the lane-engine suitability test, solver.cu)
template <typename T> __global__ void k_mean_abs_step(const T* __restrict__ y, long long n, double* __restrict__ out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (long long e = threadIdx.x; e + 1 < n; e += blockDim.x) s += fabs((double)y[e + 1] - (double)y[e]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) { if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k]; __syncthreads(); }
    if (threadIdx.x == 0) out[0] = n > 1 ? sh[0] / (double)(n - 1) : 0.0;
}
template <typename T> cudaError_t ew_mean_abs_step(const T* y, long long n, double* out, cudaStream_t st) {
    k_mean_abs_step<T><<<1, 256, 0, st>>>(y, n, out);
    return cudaGetLastError();
}

#define INST(T) \
    template cudaError_t ew_mean_abs_step<T>(const T*, long long, double*, cudaStream_t); \
    template cudaError_t ew_dr_first<T>(const T*, const T*, T*, T*, long long, cudaStream_t); \
    template cudaError_t ew_image_means_x2<T>(const T*, long long, int, T*, double*, cudaStream_t); \
    template cudaError_t ew_dr_reflect_cols<T>(const T*, const T*, T*, long long, cudaStream_t); \
    template cudaError_t ew_dr_combine_rows<T>(const T*, const T*, const T*, T*, long long, cudaStream_t); \
    template cudaError_t ew_dr_final_cols<T>(const T*, const T*, T*, long long, cudaStream_t); \
    template cudaError_t ew_dr_final_rows<T>(const T*, const T*, const T*, T*, long long, cudaStream_t); \
    template cudaError_t ew_dual_update<T>(T*, const T*, const T*, long long, cudaStream_t); \
    template cudaError_t ew_dr_reflect_bcast<T>(const T*, const T*, T*, long long, long long, int, long long, cudaStream_t); \
    template cudaError_t ew_mean_abs_diff<T>(const T*, const T*, long long, double*, double*, cudaStream_t); \
    template cudaError_t ew_pd_combine<T>(T* const*, T* const*, T* const*, T* const*, int, T*, long long, double*, double*, cudaStream_t); \
    template cudaError_t ew_pdr_combine<T>(T* const*, T* const*, T* const*, T* const*, int, T*, long long, double*, double*, cudaStream_t); \
    template cudaError_t ew_div_scalar<T>(const T*, T*, long long, T, cudaStream_t);
INST(double)
INST(float)

}  // namespace ptv
