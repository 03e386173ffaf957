// prox_dispatch.cu -- chooses the kernel family for one batched 1D prox over the fibers of an array.
#include "ptv_internal.h"

namespace ptv {

template <typename T>
cudaError_t prox_fibers_seq(const T* A, const T* B, InOp op, T* X, FiberGeom g, T lam, const T* lamv, const int* list,
                            long long nlist, cudaStream_t st);

template <typename T>
cudaError_t prox_fibers(const T* A, const T* B, InOp op, T* X, FiberGeom g, T lam, const T* lamv, Engine eng, cudaStream_t st) {
    (void)eng;
    KernelSpan span(g.inc == 1 ? KC_PROX_CONTIG : KC_PROX_STRIDED, 1, st);
    return prox_fibers_seq<T>(A, B, op, X, g, lam, lamv, nullptr, 0, st);
}

template cudaError_t prox_fibers<double>(const double*, const double*, InOp, double*, FiberGeom, double, const double*, Engine, cudaStream_t);
template cudaError_t prox_fibers<float>(const float*, const float*, InOp, float*, FiberGeom, float, const float*, Engine, cudaStream_t);

}  // namespace ptv
