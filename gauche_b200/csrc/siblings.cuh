// siblings.cuh — the arithmetic of GAUCHE's other twelve fingerprint kernels, one expression per kernel in the reference's
// evaluation order (gauche/kernels/fingerprint_kernels/<name>_kernel.py; cited per line).  Shared by the tensor-core
// epilogue (gram_siblings.cu: `which` is a template constant there, the switch folds away) and the dense real-valued
// path (siblings.cu).  Translation units that include this are compiled with -fmad=false: `n * dot + eps` and
// `(dot + eps)^2` must not be contracted into FMAs.
//   dot   <x1_i, x2_j>
//   ra    main row term: |x1_i|_1 — or 2*|x1_i|_1 + 2*|x2_i|_1 for rogers_tanimoto / sokal_sneath, which add x2_norm
//         UNTRANSPOSED in the reference (row i uses |x2_i|: needs N == M or M == 1)
//   rb    extra row term: d = common zeros of the LAST rows (faith / intersection / rand / rogers_tanimoto; one value per
//         row for batched inputs) or the Braun-Blanquet denominator max(x1_norm[-1], x2_norm[-1])
//   n2    |x2_j|_1
#pragma once
#include "common.cuh"

namespace gk {

enum Sibling {
  SIB_DICE = 0, SIB_BRAUN_BLANQUET, SIB_FAITH, SIB_FORBES, SIB_INNER_PRODUCT, SIB_INTERSECTION, SIB_OTSUKA,
  SIB_RAND, SIB_ROGERS_TANIMOTO, SIB_RUSSELL_RAO, SIB_SOGENFREI, SIB_SOKAL_SNEATH, SIB_COUNT
};

// python-side scalars of the reference, evaluated in double and cast to the tensor dtype like torch does
template <typename T>
struct SiblingConsts {
  T eps, nfeat, nfeat2eps, nfeateps;
  __host__ __device__ static SiblingConsts make(double nfeat, double eps) {
    return SiblingConsts{(T)eps, (T)nfeat, (T)(2.0 * nfeat + eps), (T)(nfeat + eps)};
  }
};

template <typename T>
__device__ __forceinline__ T sibling_eval(int which, T dot, T ra, T rb, T n2, const SiblingConsts<T>& k) {
  switch (which) {
    case SIB_DICE: return ((T)2 * dot + k.eps) / ((ra + n2) + k.eps);                      // dice_kernel.py:37-39
    case SIB_BRAUN_BLANQUET: return (dot + k.eps) / (rb + k.eps);                          // braun_blanquet_kernel.py:37-40
    case SIB_FAITH: return (((T)2 * dot + rb) + k.eps) / k.nfeat2eps;                      // faith_kernel.py:35-38
    case SIB_FORBES: return (k.nfeat * dot + k.eps) / ((ra + n2) + k.eps);                 // forbes_kernel.py:41-43
    case SIB_INNER_PRODUCT: return dot + k.eps;                                            // inner_product_kernel.py:36
    case SIB_INTERSECTION: return (dot + rb) + k.eps;                                      // intersection_kernel.py:41
    case SIB_OTSUKA: return (dot + k.eps) / sqrt((ra + n2) + k.eps);                       // otsuka_kernel.py:39-41
    case SIB_RAND: return ((dot + rb) + k.eps) / k.nfeateps;                               // rand_kernel.py:38
    case SIB_ROGERS_TANIMOTO:                                                              // rogers_tanimoto_kernel.py:40-42
      return ((dot + rb) + k.eps) / (((ra - (T)3 * dot) + rb) + k.eps);
    case SIB_RUSSELL_RAO: return (dot + k.eps) / k.nfeateps;                               // russell_rao_kernel.py:37
    case SIB_SOGENFREI: { const T t = dot + k.eps; return (t * t) / ((ra + n2) + k.eps); }  // sogenfrei_kernel.py:39-41
    case SIB_SOKAL_SNEATH: return (dot + k.eps) / ((ra - (T)3 * dot) + k.eps);             // sokal_sneath_kernel.py:39-41
    default: return (T)0;
  }
}

}  // namespace gk
