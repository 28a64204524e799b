// siblings.cu — epilogues of GAUCHE's other twelve fingerprint kernels over the exact integer statistics
// the tensor-core Gram kernel already produces.
//
// Every sibling is f(dot[i,j], |x1_i|_1, |x2_j|_1, d, n): gauche/kernels/fingerprint_kernels/
// {dice,braun_blanquet,faith,forbes,inner_product,intersection,otsuka,rand,rogers_tanimoto,russell_rao,
// sogenfrei,sokal_sneath}_kernel.py.  The float operations below are written in the reference's evaluation
// order and compute dtype so integer-valued fingerprints give bit-identical results, including the
// reference's shape quirks, which are reproduced rather than "fixed":
//   * braun_blanquet uses max(x1_norm[-1], x2_norm[-1]) — the LAST rows' norms — as a scalar denominator;
//   * faith / intersection / rand / rogers_tanimoto take d (common zeros) from the LAST rows only;
//   * rogers_tanimoto / sokal_sneath add x2_norm untransposed, i.e. row i uses |x2_i| (needs N == M or M == 1).
#include "common.cuh"

namespace gk {

enum Sibling {
  SIB_DICE = 0, SIB_BRAUN_BLANQUET, SIB_FAITH, SIB_FORBES, SIB_INNER_PRODUCT, SIB_INTERSECTION, SIB_OTSUKA,
  SIB_RAND, SIB_ROGERS_TANIMOTO, SIB_RUSSELL_RAO, SIB_SOGENFREI, SIB_SOKAL_SNEATH, SIB_COUNT
};

template <typename T>
__device__ __forceinline__ T sibling_value(int which, T dot, T n1, T n2j, T n2i, T nlast1, T nlast2, T d, T nfeat2eps,
                                           T nfeateps, T nfeat, T eps) {
  switch (which) {
    case SIB_DICE: return ((T)2 * dot + eps) / ((n1 + n2j) + eps);                       // dice_kernel.py:37-39
    case SIB_BRAUN_BLANQUET: return (dot + eps) / (fmax(nlast1, nlast2) + eps);            // braun_blanquet_kernel.py:37-40
    case SIB_FAITH: return (((T)2 * dot + d) + eps) / nfeat2eps;                           // faith_kernel.py:35-38
    case SIB_FORBES: return (nfeat * dot + eps) / ((n1 + n2j) + eps);                      // forbes_kernel.py:41-43
    case SIB_INNER_PRODUCT: return dot + eps;                                             // inner_product_kernel.py:36
    case SIB_INTERSECTION: return (dot + d) + eps;                                        // intersection_kernel.py:41
    case SIB_OTSUKA: return (dot + eps) / sqrt((n1 + n2j) + eps);                          // otsuka_kernel.py:39-41
    case SIB_RAND: return ((dot + d) + eps) / nfeateps;                                    // rand_kernel.py:38
    case SIB_ROGERS_TANIMOTO:                                                             // rogers_tanimoto_kernel.py:40-42
      return ((dot + d) + eps) / (((((T)2 * n1 + (T)2 * n2i) - (T)3 * dot) + d) + eps);
    case SIB_RUSSELL_RAO: return (dot + eps) / nfeateps;                                   // russell_rao_kernel.py:37
    case SIB_SOGENFREI: { const T t = dot + eps; return (t * t) / ((n1 + n2j) + eps); }    // sogenfrei_kernel.py:39-41
    case SIB_SOKAL_SNEATH: return (dot + eps) / ((((T)2 * n1 + (T)2 * n2i) - (T)3 * dot) + eps);  // sokal_sneath_kernel.py:39-41
    default: return (T)0;
  }
}

template <typename T, typename OutT>
__global__ void __launch_bounds__(256)
sibling_epilogue_kernel(int which, const int32_t* __restrict__ dot, int64_t ldd, int64_t n, int64_t m,
                        const int32_t* __restrict__ l1a, const int32_t* __restrict__ l1b,
                        const int64_t* __restrict__ dcommon, double nfeat, double eps, OutT* __restrict__ out,
                        int64_t ldo) {
  const T e = (T)eps;
  const T d = dcommon ? (T)(*dcommon) : (T)0;
  const T nlast1 = (T)l1a[n - 1], nlast2 = (T)l1b[m - 1];
  const T nf = (T)nfeat;
  const T nfeat2eps = (T)(2.0 * nfeat + eps);   // python-side scalar arithmetic in double, then cast to the tensor dtype
  const T nfeateps = (T)(nfeat + eps);
  const int64_t total = n * m;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / m, j = idx - i * m;
    const T n2i = (T)l1b[m == 1 ? 0 : i];   // untransposed x2_norm broadcast (N == M or M == 1)
    T q = sibling_value<T>(which, (T)dot[i * ldd + j], (T)l1a[i], (T)l1b[j], n2i, nlast1, nlast2, d, nfeat2eps,
                           nfeateps, nf, e);
    OutT o = (OutT)q;                         // similarity.to(double) where the reference does it
    out[i * ldo + j] = o < (OutT)0 ? (OutT)0 : o;   // clamp_min_(0)
  }
}

}  // namespace gk

extern "C" int gk_sibling_epilogue(int which, const int32_t* dot, int64_t ldd, int64_t n, int64_t m,
                                   const int32_t* l1a, const int32_t* l1b, const int64_t* dcommon, int64_t nfeat,
                                   int compute_dtype, void* out, int64_t ldo, int out_dtype, double eps,
                                   void* stream) {
  using namespace gk;
  const char* name = "gk_sibling_epilogue";
  GK_CHECK_ARG(which >= 0 && which < SIB_COUNT, GK_EINVAL, "%s: unknown kernel id %d", name, which);
  GK_CHECK_ARG(n >= 0 && m >= 0, GK_EINVAL, "%s: negative size", name);
  if (n == 0 || m == 0) return 0;
  GK_CHECK_ARG(dot && l1a && l1b && out && ldd >= m && ldo >= m, GK_EINVAL, "%s: bad arguments", name);
  if (which == SIB_ROGERS_TANIMOTO || which == SIB_SOKAL_SNEATH)
    GK_CHECK_ARG(n == m || m == 1, GK_EINVAL,
                 "%s: the reference adds x2_norm untransposed; this only broadcasts for N == M or M == 1", name);
  int64_t blocks = (n * m + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  const double nf = (double)nfeat;
#define GK_SIB_LAUNCH(T, OutT)                                                                                  \
  sibling_epilogue_kernel<T, OutT><<<(unsigned)blocks, 256, 0, s>>>(which, dot, ldd, n, m, l1a, l1b, dcommon, nf, \
                                                                     eps, (OutT*)out, ldo)
  if (compute_dtype == GK_F32 && out_dtype == GK_F32) GK_SIB_LAUNCH(float, float);
  else if (compute_dtype == GK_F32 && out_dtype == GK_F64) GK_SIB_LAUNCH(float, double);
  else if (compute_dtype == GK_F64 && out_dtype == GK_F64) GK_SIB_LAUNCH(double, double);
  else return set_err(GK_EINVAL, "%s: unsupported compute/out dtype combination (%d, %d)", name, compute_dtype, out_dtype);
#undef GK_SIB_LAUNCH
  return cuda_err(cudaGetLastError(), name);
}
