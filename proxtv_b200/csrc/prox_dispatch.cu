// prox_dispatch.cu -- chooses the kernel family for one batched 1D prox over the fibers of an array.
//
//   contiguous fibers (inc == 1), len >= 64, fits shared memory -> chunked speculative kernel (kernels_chunked.cu)
//   strided fibers with scratch space                          -> gather (fused input op) + chunked kernel + scatter
//   everything else (tiny fibers, weighted strided, no scratch, ENGINE_SEQ) -> sequential lane-per-fiber kernel
#include "ptv_internal.h"

namespace ptv {

template <typename T>
cudaError_t prox_fibers_seq(const T* A, const T* B, InOp op, T* X, int out_op, FiberGeom g, T lam, const T* lamv, const int* list,
                            long long nlist, cudaStream_t st);
template <typename T>
cudaError_t prox_fibers_chunked_contig(const T* A, const T* B, InOp op, T* X, int out_op, FiberGeom g, T lam, const T* lamv,
                                       cudaStream_t st);
template <typename T> cudaError_t gather_fibers(const T* A, const T* B, InOp op, T* out, FiberGeom g, cudaStream_t st);
template <typename T> cudaError_t scatter_fibers(const T* in, T* X, FiberGeom g, cudaStream_t st);

template <typename T>
cudaError_t prox_fibers(const T* A, const T* B, InOp op, T* X, int out_op, FiberGeom g, T lam, const T* lamv, Engine eng,
                        T* scratch, cudaStream_t st) {
    if (g.nf <= 0 || g.len <= 0) return cudaSuccess;
    if (eng != ENGINE_SEQ && g.len >= 2 * 32) {
        if (g.inc == 1) {
            KernelSpan span(KC_PROX_CONTIG, 1, st);
            cudaError_t e = prox_fibers_chunked_contig<T>(A, B, op, X, out_op, g, lam, lamv, st);
            if (e != cudaErrorInvalidConfiguration) return e;
            cudaGetLastError();
        } else if (scratch && !lamv && g.nf % g.inc == 0) {
            const long long n = g.nf * (long long)g.len;
            T* t1 = scratch; T* t2 = scratch + n;
            const FiberGeom gc{g.nf, g.len, 1};
            cudaError_t e;
            { KernelSpan span(KC_ELEMENTWISE, 1, st); e = gather_fibers<T>(A, B, op, t1, g, st); }
            if (e != cudaSuccess) return e;
            { KernelSpan span(KC_PROX_STRIDED, 1, st); e = prox_fibers_chunked_contig<T>(t1, nullptr, IN_A, t2, out_op, gc, lam, nullptr, st); }
            if (e == cudaSuccess) { KernelSpan span(KC_ELEMENTWISE, 1, st); return scatter_fibers<T>(t2, X, g, st); }
            if (e != cudaErrorInvalidConfiguration) return e;
            cudaGetLastError();
        }
    }
    KernelSpan span(g.inc == 1 ? KC_PROX_CONTIG : KC_PROX_STRIDED, 1, st);
    return prox_fibers_seq<T>(A, B, op, X, out_op, g, lam, lamv, nullptr, 0, st);
}

template cudaError_t prox_fibers<double>(const double*, const double*, InOp, double*, int, FiberGeom, double, const double*, Engine, double*, cudaStream_t);
template cudaError_t prox_fibers<float>(const float*, const float*, InOp, float*, int, FiberGeom, float, const float*, Engine, float*, cudaStream_t);

}  // namespace ptv
