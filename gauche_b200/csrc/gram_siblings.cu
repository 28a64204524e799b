// gram_siblings.cu — GAUCHE's other twelve fingerprint kernels as EPILOGUES of the tcgen05 Gram kernel (gram_tc.cuh):
// the exact intersection counts / dot products never leave the SM as integers, the similarity tile is written once.
//
// Every sibling is f(dot[i,j], row terms of i, column term of j): gauche/kernels/fingerprint_kernels/
// {dice,braun_blanquet,faith,forbes,inner_product,intersection,otsuka,rand,rogers_tanimoto,russell_rao,sogenfrei,
// sokal_sneath}_kernel.py.  The float operations are written in the reference's evaluation order and compute dtype, so
// integer-valued fingerprints give bit-identical results (this file is compiled with -fmad=false: `n * dot + eps` and
// `(dot + eps)^2` must not contract).  The reference's shape quirks are DATA here, not code: the caller passes
//   row_a[i]  the main per-row integer: |x1_i|_1, or |x1_i|_1 + |x2_i|_1 for rogers_tanimoto / sokal_sneath (the reference
//             adds x2_norm UNTRANSPOSED; 2*n1 + 2*n2 is an exact integer in float, so (T)(2 * (n1 + n2)) reproduces it)
//   row_b[i]  the extra per-row integer: d = common zeros of the LAST rows (faith / intersection / rand /
//             rogers_tanimoto; a per-row vector for batched inputs) or the Braun-Blanquet denominator
//             max(x1_norm[-1], x2_norm[-1]); may be NULL (zeros)
//   col[j]    |x2_j|_1
#include "gram_tc.cuh"
#include "siblings.cuh"

namespace gk {
namespace gtc {

template <typename T>
struct SibRow { T a, b; };

template <int kWhich, typename T, typename Out>
struct SiblingEpi {
  using TermT = T;
  using RowT = SibRow<T>;
  using OutT = Out;
  static constexpr bool kPacked = false;
  static constexpr bool kRowdot = false;
  static constexpr bool kSparse = false;
  static constexpr bool kRow2 = true;
  const int32_t* row2;                 // row_b, may be NULL
  SiblingConsts<T> k;
  __device__ __forceinline__ T col_term(int v) const { return (T)v; }
  __device__ __forceinline__ RowT row_term2(int a, int b) const {
    // rogers_tanimoto / sokal_sneath: row_a = |x1_i| + |x2_i|; 2*n1 + 2*n2 is an exact integer in float
    const bool doubled = kWhich == SIB_ROGERS_TANIMOTO || kWhich == SIB_SOKAL_SNEATH;
    return RowT{doubled ? (T)(2 * a) : (T)a, (T)b};
  }
  template <bool kF32Acc>
  __device__ __forceinline__ Out one(uint32_t raw, RowT r, T n2) const {
    const T dot = (T)(kF32Acc ? __float2int_rn(__uint_as_float(raw)) : (int)raw);
    const Out o = (Out)sibling_eval<T>(kWhich, dot, r.a, r.b, n2, k);   // kWhich is a constant: the switch folds away
    return o < (Out)0 ? (Out)0 : o;                                    // .to(double) where the reference does it; clamp_min_(0)
  }
};

struct SibArgs {
  const void *a, *b;
  int64_t lda, ldb, n, m, k;
  const int32_t *row_a, *row_b, *col;
  void* out;
  int64_t ldo;
  double nfeat, eps;
  cudaStream_t s;
};

template <int kWhich, typename T, typename Out>
static int launch_one(int operand_kind, const SibArgs& x) {
  SiblingEpi<kWhich, T, Out> epi;
  epi.row2 = x.row_b;
  epi.k = SiblingConsts<T>::make(x.nfeat, x.eps);
  const CUtensorMapDataType dt = sizeof(Out) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (operand_kind == 1)   // packed e2m1 bits
    return launch<1, KindMxf4, false>(x.a, x.lda, x.row_a, x.n, x.b, x.ldb, x.col, x.m, x.k, x.out, x.ldo, dt, epi, x.s);
  return launch<2, KindI8, false>(x.a, x.lda, x.row_a, x.n, x.b, x.ldb, x.col, x.m, x.k, x.out, x.ldo, dt, epi, x.s);   // u8 counts
}

template <int kWhich>
static int launch_dtypes(int operand_kind, int compute_dtype, int out_dtype, const SibArgs& x) {
  // kernels whose reference ends with `.to(dtype=torch.double)` compute in the input dtype and store doubles
  constexpr bool kToDouble = !(kWhich == SIB_DICE || kWhich == SIB_FAITH || kWhich == SIB_INNER_PRODUCT);
  if (compute_dtype == GK_F64 && out_dtype == GK_F64) return launch_one<kWhich, double, double>(operand_kind, x);
  if constexpr (kToDouble) {
    if (compute_dtype == GK_F32 && out_dtype == GK_F64) return launch_one<kWhich, float, double>(operand_kind, x);
  } else {
    if (compute_dtype == GK_F32 && out_dtype == GK_F32) return launch_one<kWhich, float, float>(operand_kind, x);
  }
  return set_err(GK_EINVAL, "gk_sibling_gram: kernel %d does not produce (compute %d, out %d)", kWhich, compute_dtype, out_dtype);
}

}  // namespace gtc
}  // namespace gk

extern "C" int gk_sibling_gram(int which, int operand_kind, const void* a, int64_t lda, const int32_t* row_a,
                               const int32_t* row_b, int64_t n, const void* b, int64_t ldb, const int32_t* col,
                               int64_t m, int64_t k_bytes, int64_t nfeat, int compute_dtype, void* out, int64_t ldo,
                               int out_dtype, double eps, void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_sibling_gram";
  GK_CHECK_ARG(which >= 0 && which < SIB_COUNT, GK_EINVAL, "%s: unknown kernel id %d", name, which);
  GK_CHECK_ARG(operand_kind == 0 || operand_kind == 1, GK_EINVAL, "%s: operand_kind must be 0 (u8) or 1 (packed e2m1)", name);
  if (n == 0 || m == 0) return 0;
  if (int e = check_args(name, a, lda, row_a, n, b, ldb, col, m, k_bytes, BK, out, ldo, out_dtype)) return e;
  if (operand_kind == 1)
    GK_CHECK_ARG(k_bytes * 2 < (1LL << 24), GK_ERANGE, "%s: fingerprint too long for exact f32 accumulation", name);
  const SibArgs x{a, b, lda, ldb, n, m, k_bytes, row_a, row_b, col, out, ldo, (double)nfeat, eps, (cudaStream_t)stream};
  switch (which) {
    case SIB_DICE: return launch_dtypes<SIB_DICE>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_BRAUN_BLANQUET: return launch_dtypes<SIB_BRAUN_BLANQUET>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_FAITH: return launch_dtypes<SIB_FAITH>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_FORBES: return launch_dtypes<SIB_FORBES>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_INNER_PRODUCT: return launch_dtypes<SIB_INNER_PRODUCT>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_INTERSECTION: return launch_dtypes<SIB_INTERSECTION>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_OTSUKA: return launch_dtypes<SIB_OTSUKA>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_RAND: return launch_dtypes<SIB_RAND>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_ROGERS_TANIMOTO: return launch_dtypes<SIB_ROGERS_TANIMOTO>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_RUSSELL_RAO: return launch_dtypes<SIB_RUSSELL_RAO>(operand_kind, compute_dtype, out_dtype, x);
    case SIB_SOGENFREI: return launch_dtypes<SIB_SOGENFREI>(operand_kind, compute_dtype, out_dtype, x);
    default: return launch_dtypes<SIB_SOKAL_SNEATH>(operand_kind, compute_dtype, out_dtype, x);
  }
}
