// screen.cu — caller-side helpers of the cross-kernel used for candidate screening
// (protocol of gauche/gp.py:169-226: K(X*,X) feeds the posterior mean K·alpha; the
// BO notebook then ranks candidates by acquisition score).
//
//   gk_rowdot_f32 : scores = bias + K @ alpha, one CTA per Gram row, float4 loads.
//   gk_topk_f32   : exact, deterministic top-k by 8-pass radix select on the unique
//                   64-bit key (ordered(score) << 32 | ~index), then an in-smem bitonic
//                   sort of the k survivors.  Ties go to the smaller index (== stable
//                   descending sort), NaN ranks above +inf like torch.topk.
#include "common.cuh"

namespace gk {

// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ K, int64_t ldk, int64_t n, int64_t m,
              const float* __restrict__ alpha, float bias, float* __restrict__ scores,
              bool vec_ok) {
  __shared__ float red[8];
  for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
    const float* __restrict__ row = K + r * ldk;
    float acc = 0.f;
    if (vec_ok) {
      const int64_t m4 = m >> 2;
      for (int64_t v = threadIdx.x; v < m4; v += blockDim.x) {
        const float4 kv = __ldg(reinterpret_cast<const float4*>(row) + v);
        const float4 av = __ldg(reinterpret_cast<const float4*>(alpha) + v);
        acc = fmaf(kv.x, av.x, acc);
        acc = fmaf(kv.y, av.y, acc);
        acc = fmaf(kv.z, av.z, acc);
        acc = fmaf(kv.w, av.w, acc);
      }
      for (int64_t j = (m4 << 2) + threadIdx.x; j < m; j += blockDim.x)
        acc = fmaf(row[j], alpha[j], acc);
    } else {
      for (int64_t j = threadIdx.x; j < m; j += blockDim.x) acc = fmaf(row[j], alpha[j], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (threadIdx.x == 0) scores[r] = bias + t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------
struct TopkState {
  unsigned long long prefix;  // bits decided so far (high to low)
  long long k_rem;            // how many keys >= the final threshold are still to be found
  unsigned int count;         // append cursor of the select pass
  unsigned int pad;
};

__device__ __forceinline__ unsigned long long make_key(float s, int64_t i) {
  const uint32_t u = __float_as_uint(s);
  const uint32_t f = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)f << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
}
__device__ __forceinline__ float key_score(unsigned long long key) {
  const uint32_t f = (uint32_t)(key >> 32);
  const uint32_t u = (f & 0x80000000u) ? (f & 0x7fffffffu) : ~f;
  return __uint_as_float(u);
}
__device__ __forceinline__ int64_t key_index(unsigned long long key) {
  return (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffu));
}

__global__ void __launch_bounds__(256)
topk_hist_kernel(const float* __restrict__ scores, int64_t n, const TopkState* __restrict__ st,
                 unsigned int* __restrict__ hist, int pass) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = st->prefix;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = make_key(scores[i], i);
    const bool match = (pass == 0) || ((key >> (shift + 8)) == (prefix >> (shift + 8)));
    if (match) atomicAdd(&sh[(unsigned)(key >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[pass * 256 + threadIdx.x], sh[threadIdx.x]);
}

__global__ void topk_init_kernel(TopkState* st, long long k) {
  st->prefix = 0ull;
  st->k_rem = k;
  st->count = 0u;
  st->pad = 0u;
}

__global__ void topk_pick_kernel(TopkState* st, const unsigned int* __restrict__ hist, int pass) {
  // single thread: 256 bins
  if (threadIdx.x != 0) return;
  const int shift = 56 - 8 * pass;
  long long k_rem = st->k_rem;
  long long above = 0;
  int d = 255;
  for (; d > 0; --d) {
    const long long c = hist[pass * 256 + d];
    if (above + c >= k_rem) break;
    above += c;
  }
  st->prefix |= ((unsigned long long)d) << shift;
  st->k_rem = k_rem - above;
}

__global__ void __launch_bounds__(256)
topk_select_kernel(const float* __restrict__ scores, int64_t n, TopkState* st,
                   unsigned long long* __restrict__ cand, int k) {
  const unsigned long long thr = st->prefix;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = make_key(scores[i], i);
    if (key >= thr) {
      const unsigned int pos = atomicAdd(&st->count, 1u);
      if (pos < (unsigned)k) cand[pos] = key;
    }
  }
}

__global__ void __launch_bounds__(1024)
topk_sort_kernel(const unsigned long long* __restrict__ cand, int k, int kpad, int64_t index_base,
                 float* __restrict__ out_vals, int64_t* __restrict__ out_idx) {
  extern __shared__ unsigned long long skeys[];
  for (int i = threadIdx.x; i < kpad; i += blockDim.x) skeys[i] = (i < k) ? cand[i] : 0ull;
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (kpad >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = skeys[lo], b = skeys[hi];
        if ((a < b) == desc) {
          skeys[lo] = b;
          skeys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    out_vals[i] = key_score(skeys[i]);
    out_idx[i] = index_base + key_index(skeys[i]);
  }
}

constexpr int kTopkMaxK = 16384;
static int64_t topk_ws_bytes(int k) {
  return 256 /*state, padded*/ + 8 * 256 * 4 /*hist*/ + (int64_t)k * 8 /*candidates*/;
}

}  // namespace gk

extern "C" int gk_rowdot_f32(const float* K, int64_t ldk, int64_t n, int64_t m, const float* alpha,
                             float bias, float* scores, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(n >= 0 && m >= 0, GK_EINVAL, "gk_rowdot_f32: negative size");
  if (n == 0) return 0;
  GK_CHECK_ARG(K && alpha && scores, GK_EINVAL, "gk_rowdot_f32: null pointer");
  GK_CHECK_ARG(ldk >= m, GK_EINVAL, "gk_rowdot_f32: ldk < m");
  const bool vec_ok = ((uintptr_t)K % 16 == 0) && ((uintptr_t)alpha % 16 == 0) && (ldk % 4 == 0);
  int64_t blocks = n;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  rowdot_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(K, ldk, n, m, alpha, bias,
                                                                     scores, vec_ok);
  return cuda_err(cudaGetLastError(), "rowdot_kernel launch");
}

extern "C" int64_t gk_topk_workspace(int64_t n, int k) {
  (void)n;
  if (k <= 0 || k > gk::kTopkMaxK) return -1;
  return gk::topk_ws_bytes(k);
}

extern "C" int gk_topk_f32(const float* scores, int64_t n, int k, int64_t index_base,
                           float* out_vals, int64_t* out_idx, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(k > 0 && k <= kTopkMaxK, GK_ERANGE, "gk_topk_f32: k must be in [1,%d]", kTopkMaxK);
  GK_CHECK_ARG(n >= k, GK_EINVAL, "gk_topk_f32: n (%lld) < k (%d)", (long long)n, k);
  GK_CHECK_ARG(n <= 0xffffffffLL, GK_ERANGE, "gk_topk_f32: n exceeds 2^32-1");
  GK_CHECK_ARG(scores && out_vals && out_idx && workspace, GK_EINVAL, "gk_topk_f32: null pointer");
  GK_CHECK_ARG(workspace_bytes >= topk_ws_bytes(k), GK_EINVAL, "gk_topk_f32: workspace too small");
  GK_CHECK_ARG((uintptr_t)workspace % 8 == 0, GK_EALIGN, "gk_topk_f32: workspace must be 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  TopkState* st = (TopkState*)ws;
  unsigned int* hist = (unsigned int*)(ws + 256);
  unsigned long long* cand = (unsigned long long*)(ws + 256 + 8 * 256 * 4);
  GK_CUDA(cudaMemsetAsync(hist, 0, 8 * 256 * 4, s));
  topk_init_kernel<<<1, 1, 0, s>>>(st, (long long)k);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  for (int pass = 0; pass < 8; ++pass) {
    topk_hist_kernel<<<(unsigned)blocks, 256, 0, s>>>(scores, n, st, hist, pass);
    topk_pick_kernel<<<1, 32, 0, s>>>(st, hist, pass);
  }
  topk_select_kernel<<<(unsigned)blocks, 256, 0, s>>>(scores, n, st, cand, k);
  int kpad = 2;
  while (kpad < k) kpad <<= 1;
  const size_t smem = (size_t)kpad * sizeof(unsigned long long);
  if (smem > 48 * 1024) {
    GK_CUDA(cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  }
  topk_sort_kernel<<<1, 1024, smem, s>>>(cand, k, kpad, index_base, out_vals, out_idx);
  return cuda_err(cudaGetLastError(), "gk_topk_f32 launch");
}
