// pack.cu — dense fingerprint matrix -> packed bits + u8 operand + exact integer row norms
// + classification flags, one pass, HBM-bound.
//
// Replaces the row-norm reductions of the reference (tanimoto_kernel.py:37-38,
// minmax_kernel.py:33-34) and produces the operands of the integer Gram kernels.
//
// Mapping: one warp per row.  Per iteration the warp covers 128 consecutive elements,
// lane l owning elements [4l, 4l+4): one vector load in, one 32-bit store of four u8
// out, and a 4-bit nibble that three xor-shuffles merge into the four 32-bit bit-words
// of the 128-element chunk (bit i of word w <-> element 32w+i).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"

namespace gk {

template <typename T>
__device__ __forceinline__ double to_double(T v) { return (double)v; }
template <>
__device__ __forceinline__ double to_double<__half>(__half v) { return (double)__half2float(v); }
template <>
__device__ __forceinline__ double to_double<__nv_bfloat16>(__nv_bfloat16 v) {
  return (double)__bfloat162float(v);
}

template <typename T>
struct Vec4 { T v[4]; };

template <typename T>
__device__ __forceinline__ void load4(const T* __restrict__ row, int64_t e, int64_t D, bool vec_ok,
                                      Vec4<T>& out) {
  if (vec_ok && e + 4 <= D) {
    if constexpr (sizeof(T) == 4) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(row + e));
      *reinterpret_cast<uint4*>(out.v) = raw;
    } else if constexpr (sizeof(T) == 8) {
      const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(row + e));
      const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(row + e + 2));
      reinterpret_cast<uint4*>(out.v)[0] = r0;
      reinterpret_cast<uint4*>(out.v)[1] = r1;
    } else if constexpr (sizeof(T) == 2) {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(row + e));
      *reinterpret_cast<uint2*>(out.v) = raw;
    } else {
      const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(row + e));
      *reinterpret_cast<uint32_t*>(out.v) = raw;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) out.v[i] = (e + i < D) ? row[e + i] : T(0);
  }
}

// unconditional vector load of four consecutive elements (16-byte aligned rows, e % 4 == 0, e + 4 <= D)
template <typename T>
__device__ __forceinline__ void load4_vec(const T* __restrict__ row, int64_t e, Vec4<T>& out) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<uint4*>(out.v) = __ldg(reinterpret_cast<const uint4*>(row + e));
  } else if constexpr (sizeof(T) == 8) {
    reinterpret_cast<uint4*>(out.v)[0] = __ldg(reinterpret_cast<const uint4*>(row + e));
    reinterpret_cast<uint4*>(out.v)[1] = __ldg(reinterpret_cast<const uint4*>(row + e + 2));
  } else if constexpr (sizeof(T) == 2) {
    *reinterpret_cast<uint2*>(out.v) = __ldg(reinterpret_cast<const uint2*>(row + e));
  } else {
    *reinterpret_cast<uint32_t*>(out.v) = __ldg(reinterpret_cast<const uint32_t*>(row + e));
  }
}

// kFast: rows are 16-byte aligned and D % 4 == 0, so every lane's four elements lie completely inside or completely
// outside the row: the loads of a batch of kU chunks are predicated vector loads with no branches between them and are
// all in flight before the first one is consumed (the branchy generic loader serialised them: one 512-byte request per
// warp at a time, 1.5 TB/s on fp32 input).
template <typename T, bool kFast>
__global__ void __launch_bounds__(256)
pack_classify_kernel(const T* __restrict__ x, int64_t rows, int64_t D, int64_t ld,
                     uint32_t* __restrict__ bits, int64_t ld_bits, uint8_t* __restrict__ q8,
                     int64_t ld_q8, int32_t* __restrict__ rowsum, int32_t* __restrict__ rowsq,
                     int32_t* __restrict__ flags) {
  constexpr int kU = sizeof(T) >= 8 ? 4 : 8;      // chunks (128 elements each) per batch
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int local_flags = 0;
  int local_max = 0;

  for (int64_t r = warp0; r < rows; r += nwarps) {
    const T* __restrict__ row = x + r * ld;
    int sum = 0, sq = 0;
    for (int64_t base0 = 0; base0 < ld_q8; base0 += 128 * kU) {
      Vec4<T> in[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t e = base0 + 128 * u + 4 * lane;
        if constexpr (kFast) {
#pragma unroll
          for (int i = 0; i < 4; ++i) in[u].v[i] = T(0);
          if (e < D) load4_vec<T>(row, e, in[u]);
        } else {
          if (e < D) {
            load4<T>(row, e, D, false, in[u]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) in[u].v[i] = T(0);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t base = base0 + 128 * u;
        if (base >= ld_q8) break;                 // warp-uniform
        const int64_t e = base + 4 * lane;
        uint32_t packed = 0, nib = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // classification in the narrowest arithmetic that is exact for T: fp32 inputs (the common case: 8 KB per
          // molecule to stream) stay in fp32, small integers in int; only 64-bit inputs need double
          bool is_bin, is_u8, nz;
          uint32_t q;
          if constexpr (sizeof(T) <= 4 && !std::is_integral<T>::value) {         // float, half, bfloat16
            float v;
            if constexpr (std::is_same<T, float>::value) v = in[u].v[i];
            else v = (float)to_double(in[u].v[i]);
            is_bin = (v == 0.0f) || (v == 1.0f);
            is_u8 = (v >= 0.0f) && (v <= 255.0f) && (v == truncf(v));
            nz = v != 0.0f;
            q = is_u8 ? (uint32_t)v : 0u;
          } else if constexpr (std::is_integral<T>::value && sizeof(T) <= 4) {    // u8 / bool / i8 / i16 / i32
            const int v = (int)in[u].v[i];
            is_bin = (v == 0) || (v == 1);
            is_u8 = (v >= 0) && (v <= 255);
            nz = v != 0;
            q = is_u8 ? (uint32_t)v : 0u;
          } else {                                                               // double, int64
            const double v = to_double(in[u].v[i]);
            is_bin = (v == 0.0) || (v == 1.0);
            is_u8 = (v >= 0.0) && (v <= 255.0) && (v == floor(v));
            nz = v != 0.0;
            q = is_u8 ? (uint32_t)v : 0u;
          }
          if (!is_bin) local_flags |= GK_FLAG_NONBINARY;
          if (!is_u8) local_flags |= GK_FLAG_NONU8;
          packed |= q << (8 * i);
          nib |= (nz ? 1u : 0u) << i;
          sum += (int)q;
          sq += (int)(q * q);
          local_max = max(local_max, (int)q);
        }
        if (q8) *reinterpret_cast<uint32_t*>(q8 + r * ld_q8 + e) = packed;   // bit fingerprints never need the u8 copy
        uint32_t w = nib << (4 * (lane & 7));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if ((lane & 7) == 0) bits[r * ld_bits + (base >> 5) + (lane >> 3)] = w;
      }
    }
    // per-lane partial sums fit int32 (D <= 2^20); the row totals are formed in 64 bits: a count row whose
    // squared norm does not fit int32 is classified real-valued and takes the dense path (any D) instead of failing
    long long sq64 = sq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq64 += __shfl_xor_sync(0xffffffffu, sq64, o);
    }
    if (sq64 > 0x7fffffffLL) local_flags |= GK_FLAG_NONBINARY | GK_FLAG_NONU8;
    if (lane == 0) {
      rowsum[r] = sum;
      rowsq[r] = (int32_t)sq64;
    }
  }
  local_flags |= __shfl_xor_sync(0xffffffffu, local_flags, 16);
  local_flags |= __shfl_xor_sync(0xffffffffu, local_flags, 8);
  local_flags |= __shfl_xor_sync(0xffffffffu, local_flags, 4);
  local_flags |= __shfl_xor_sync(0xffffffffu, local_flags, 2);
  local_flags |= __shfl_xor_sync(0xffffffffu, local_flags, 1);
  if (lane == 0 && local_flags) atomicOr(flags, local_flags);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if (lane == 0 && local_max > 0) atomicMax(flags + 1, local_max);   // largest u8 value seen (thermometer depth)
}

template <typename T>
static int launch_pack(const void* x, int64_t rows, int64_t D, int64_t ld, uint32_t* bits,
                       int64_t ld_bits, uint8_t* q8, int64_t ld_q8, int32_t* rowsum,
                       int32_t* rowsq, int32_t* flags, cudaStream_t stream) {
  const int warps_per_block = 8;
  int64_t blocks = (rows + warps_per_block - 1) / warps_per_block;
  const int64_t cap = (int64_t)sm_count() * 32;  // persistent-ish grid, 8 CTAs x 8 warps per SM x 4
  if (blocks > cap) blocks = cap;
  // vector loads need 16-byte aligned rows; the branch-free batch loader additionally whole 4-element groups
  const bool fast = (((uintptr_t)x) % 16 == 0) && ((ld * (int64_t)sizeof(T)) % 16 == 0) && (D % 4 == 0);
  if (fast)
    pack_classify_kernel<T, true><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(
        (const T*)x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags);
  else
    pack_classify_kernel<T, false><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(
        (const T*)x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags);
  return cuda_err(cudaGetLastError(), "pack_classify_kernel launch");
}

// bits -> packed e2m1 nibbles (1.0 = 0b0010, 0.0 = 0b0000), the operand format of the kind::mxf4
// Gram kernel.  Nibble k of the output row <-> bit k of the input row (any fixed order works: both
// operands of the dot product use the same one).  One thread per 32-bit word -> one 16-byte store.
__global__ void __launch_bounds__(256)
expand_bits_e2m1_kernel(const uint32_t* __restrict__ bits, int64_t ld_bits, int64_t rows, int64_t words_in,
                        uint8_t* __restrict__ q4, int64_t ld_q4, int64_t words_out) {
  const int64_t total = rows * words_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / words_out, w = i - r * words_out;
    const uint32_t word = (w < words_in) ? __ldg(bits + r * ld_bits + w) : 0u;
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint32_t t = (word >> (8 * b)) & 0xFFu;          // 8 bits -> 8 nibbles
      t = (t | (t << 12)) & 0x000F000Fu;
      t = (t | (t << 6)) & 0x03030303u;
      t = (t | (t << 3)) & 0x11111111u;
      o[b] = t << 1;                                   // nibble value 0b0010 == e2m1 1.0
    }
    *reinterpret_cast<uint4*>(q4 + r * ld_q4 + w * 16) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// bits -> the bit order the hybrid Gram kernel expands from (gram_tc.cuh, kBX = 3).  Inside every 256-bit k-block (eight
// words) bit 8j+i of word 2s+h moves to bit 4i+s of word 4h+j (h < 2, j < 4, s < 4, i < 8): the kernel's five-op expansion
// — nibble i of the word it writes to K segment s, chunk h, position j is bit 4i+s of word 4h+j — then yields the natural
// element order of the e2m1 operand (gk_expand_bits_e2m1) on the other side of the MMA.  One thread per output word.
__global__ void __launch_bounds__(256)
permute_bits_kblock_kernel(const uint32_t* __restrict__ bits, int64_t ld_bits, int64_t rows, int64_t words_in,
                           uint32_t* __restrict__ out, int64_t ld_out, int64_t words_out) {
  const int64_t total = rows * words_out;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / words_out, w = idx - r * words_out;
    const int64_t kb = w >> 3;
    const int h = (int)(w & 7) >> 2, j = (int)(w & 3);
    uint32_t o = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int64_t wi = kb * 8 + 2 * s + h;
      uint32_t t = (wi < words_in) ? (__ldg(bits + r * ld_bits + wi) >> (8 * j)) & 0xFFu : 0u;   // 8 bits -> every 4th bit
      t = (t | (t << 12)) & 0x000F000Fu;
      t = (t | (t << 6)) & 0x03030303u;
      t = (t | (t << 3)) & 0x11111111u;
      o |= t << s;
    }
    out[r * ld_out + w] = o;
  }
}

// bits -> u8 0/1 operand rows (lazy operand of the int8 kernels when the source was packed bits)
__global__ void __launch_bounds__(256)
expand_bits_u8_kernel(const uint32_t* __restrict__ bits, int64_t ld_bits, int64_t rows, int64_t words,
                      uint8_t* __restrict__ q8, int64_t ld_q8) {
  const int64_t total = rows * words * 2;   // one thread per 16 bits -> one 16-byte store
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (words * 2), h = i - r * (words * 2);
    const uint32_t half = (__ldg(bits + r * ld_bits + (h >> 1)) >> (16 * (h & 1))) & 0xFFFFu;
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint32_t t = (half >> (4 * b)) & 0xFu;           // 4 bits -> 4 bytes
      t = (t | (t << 14)) & 0x00030003u;
      t = (t | (t << 7)) & 0x01010101u;
      o[b] = t;
    }
    *reinterpret_cast<uint4*>(q8 + r * ld_q8 + h * 16) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// per-row popcount of packed bit rows (the exact norms of a library delivered as packed bits)
__global__ void __launch_bounds__(256)
bits_popcount_kernel(const uint32_t* __restrict__ bits, int64_t ld_bits, int64_t rows, int64_t words,
                     int32_t* __restrict__ popcnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp0; r < rows; r += nwarps) {
    int c = 0;
    for (int64_t w = lane; w < words; w += 32) c += __popc(__ldg(bits + r * ld_bits + w));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) popcnt[r] = c;
  }
}

// u8 counts -> T thermometer planes of packed e2m1 nibbles: plane t (1..T) holds [x >= t] as 1.0/0.0.
// sum_c min(a_c, b_c) = sum_t <plane_t(a), plane_t(b)>, so MinMax runs on the FP4 tensor-core Gram
// kernel with K = T * Kp.  One thread per (row, plane, 8 input bytes) -> one 4-byte store.
__global__ void __launch_bounds__(256)
expand_thermometer_e2m1_kernel(const uint8_t* __restrict__ q8, int64_t ld_q8, int64_t rows, int64_t kp, int T,
                               uint8_t* __restrict__ q4t, int64_t ld_q4t, int64_t plane_bytes) {
  const int64_t chunks = plane_bytes / 4;           // 8 elements (4 output bytes) per chunk
  const int64_t total = rows * T * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i % chunks;
    const int64_t rt = i / chunks;
    const int t = (int)(rt % T) + 1;
    const int64_t r = rt / T;
    uint32_t o = 0;
    if (c * 8 < kp) {
      const uint2 in = __ldg(reinterpret_cast<const uint2*>(q8 + r * ld_q8 + c * 8));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if ((int)((in.x >> (8 * e)) & 0xFFu) >= t) o |= 0x2u << (4 * e);
        if ((int)((in.y >> (8 * e)) & 0xFFu) >= t) o |= 0x2u << (4 * (e + 4));
      }
    }
    *reinterpret_cast<uint32_t*>(q4t + r * ld_q4t + (int64_t)(t - 1) * plane_bytes + c * 4) = o;
  }
}

}  // namespace gk

extern "C" int gk_expand_thermometer_e2m1(const uint8_t* q8, int64_t ld_q8, int64_t rows, int64_t kp, int planes,
                                          uint8_t* q4t, int64_t ld_q4t, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(rows >= 0 && kp > 0 && planes >= 1 && planes <= 255, GK_EINVAL, "gk_expand_thermometer_e2m1: bad size");
  if (rows == 0) return 0;
  GK_CHECK_ARG(q8 && q4t, GK_EINVAL, "gk_expand_thermometer_e2m1: null pointer");
  GK_CHECK_ARG(kp % 128 == 0 && ld_q8 >= kp && ld_q8 % 8 == 0 && ((uintptr_t)q8) % 8 == 0, GK_EALIGN,
               "gk_expand_thermometer_e2m1: q8 must be 8-byte aligned with kp a multiple of 128");
  const int64_t plane_bytes = (kp + 255) / 256 * 128;
  GK_CHECK_ARG(ld_q4t >= plane_bytes * planes && ld_q4t % 128 == 0 && ((uintptr_t)q4t) % 16 == 0, GK_EALIGN,
               "gk_expand_thermometer_e2m1: ld_q4t must be a multiple of 128 and >= planes * plane bytes");
  const int64_t total = rows * planes * (plane_bytes / 4);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  expand_thermometer_e2m1_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(q8, ld_q8, rows, kp, planes, q4t,
                                                                                    ld_q4t, plane_bytes);
  return cuda_err(cudaGetLastError(), "expand_thermometer_e2m1_kernel launch");
}

extern "C" int gk_expand_bits_u8(const uint32_t* bits, int64_t ld_bits, int64_t rows, int64_t words,
                                 uint8_t* q8, int64_t ld_q8, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(rows >= 0 && words >= 0, GK_EINVAL, "gk_expand_bits_u8: negative size");
  if (rows == 0 || words == 0) return 0;
  GK_CHECK_ARG(bits && q8, GK_EINVAL, "gk_expand_bits_u8: null pointer");
  GK_CHECK_ARG(ld_bits >= words && ld_q8 >= words * 32, GK_EINVAL, "gk_expand_bits_u8: leading dimension too small");
  GK_CHECK_ARG(((uintptr_t)q8) % 16 == 0 && ld_q8 % 16 == 0, GK_EALIGN, "gk_expand_bits_u8: q8 must be 16-byte aligned");
  int64_t blocks = (rows * words * 2 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  expand_bits_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bits, ld_bits, rows, words, q8, ld_q8);
  return cuda_err(cudaGetLastError(), "expand_bits_u8_kernel launch");
}

extern "C" int gk_bits_popcount(const uint32_t* bits, int64_t ld_bits, int64_t rows, int64_t words,
                                int32_t* popcnt, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(rows >= 0 && words >= 0, GK_EINVAL, "gk_bits_popcount: negative size");
  if (rows == 0) return 0;
  GK_CHECK_ARG(bits && popcnt, GK_EINVAL, "gk_bits_popcount: null pointer");
  GK_CHECK_ARG(ld_bits >= words, GK_EINVAL, "gk_bits_popcount: ld_bits < words");
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  bits_popcount_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bits, ld_bits, rows, words, popcnt);
  return cuda_err(cudaGetLastError(), "bits_popcount_kernel launch");
}

extern "C" int gk_expand_bits_e2m1(const uint32_t* bits, int64_t ld_bits, int64_t rows, int64_t words,
                                   uint8_t* q4, int64_t ld_q4, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(rows >= 0 && words >= 0, GK_EINVAL, "gk_expand_bits_e2m1: negative size");
  if (rows == 0) return 0;
  GK_CHECK_ARG(bits && q4, GK_EINVAL, "gk_expand_bits_e2m1: null pointer");
  GK_CHECK_ARG(ld_bits >= words, GK_EINVAL, "gk_expand_bits_e2m1: ld_bits < words");
  GK_CHECK_ARG(ld_q4 % 128 == 0 && ld_q4 >= words * 16 && ld_q4 > 0, GK_EALIGN,
               "gk_expand_bits_e2m1: ld_q4 must be a positive multiple of 128 bytes and >= 16*words");
  GK_CHECK_ARG(((uintptr_t)q4) % 16 == 0, GK_EALIGN, "gk_expand_bits_e2m1: q4 must be 16-byte aligned");
  const int64_t words_out = ld_q4 / 16;
  int64_t blocks = (rows * words_out + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  expand_bits_e2m1_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bits, ld_bits, rows, words, q4,
                                                                             ld_q4, words_out);
  return cuda_err(cudaGetLastError(), "expand_bits_e2m1_kernel launch");
}

extern "C" int gk_permute_bits_kblock(const uint32_t* bits, int64_t ld_bits, int64_t rows, int64_t words,
                                      uint32_t* out, int64_t ld_out, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(rows >= 0 && words >= 0, GK_EINVAL, "gk_permute_bits_kblock: negative size");
  if (rows == 0) return 0;
  GK_CHECK_ARG(bits && out, GK_EINVAL, "gk_permute_bits_kblock: null pointer");
  GK_CHECK_ARG(ld_bits >= words, GK_EINVAL, "gk_permute_bits_kblock: ld_bits < words");
  GK_CHECK_ARG(ld_out % 8 == 0 && ld_out >= words && ld_out > 0, GK_EALIGN,
               "gk_permute_bits_kblock: ld_out must be a positive multiple of 8 words (whole k-blocks) and >= words");
  int64_t blocks = (rows * ld_out + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  permute_bits_kblock_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bits, ld_bits, rows, words, out, ld_out,
                                                                                ld_out);
  return cuda_err(cudaGetLastError(), "permute_bits_kblock_kernel launch");
}

extern "C" int gk_pack_classify(const void* x, int dtype, int64_t rows, int64_t D, int64_t ld,
                                uint32_t* bits, int64_t ld_bits, uint8_t* q8, int64_t ld_q8,
                                int32_t* rowsum, int32_t* rowsq, int32_t* flags, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(rows >= 0 && D >= 0, GK_EINVAL, "gk_pack_classify: negative size");
  if (rows == 0) return 0;
  GK_CHECK_ARG(x && bits && rowsum && rowsq && flags, GK_EINVAL, "gk_pack_classify: null pointer");
  GK_CHECK_ARG(ld >= D, GK_EINVAL, "gk_pack_classify: ld (%lld) < D (%lld)", (long long)ld,
               (long long)D);
  GK_CHECK_ARG(ld_q8 >= D && ld_q8 % 128 == 0 && ld_q8 > 0, GK_EALIGN,
               "gk_pack_classify: ld_q8 (%lld) must be a positive multiple of 128 and >= D",
               (long long)ld_q8);
  GK_CHECK_ARG(ld_bits * 32 == ld_q8, GK_EALIGN, "gk_pack_classify: ld_bits must equal ld_q8/32");
  GK_CHECK_ARG(((uintptr_t)q8) % 4 == 0, GK_EALIGN, "gk_pack_classify: q8 must be 4-byte aligned");
  // per-lane int32 partial sums of squared u8 values (D/32 elements each)
  GK_CHECK_ARG(D <= (1LL << 20), GK_ERANGE, "gk_pack_classify: D (%lld) > 2^20", (long long)D);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case GK_F32: return launch_pack<float>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    case GK_F64: return launch_pack<double>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    case GK_I64: return launch_pack<int64_t>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    case GK_I32: return launch_pack<int32_t>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    case GK_I16: return launch_pack<int16_t>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    case GK_I8: return launch_pack<int8_t>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    case GK_U8:
    case GK_BOOL: return launch_pack<uint8_t>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    case GK_F16: return launch_pack<__half>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    case GK_BF16: return launch_pack<__nv_bfloat16>(x, rows, D, ld, bits, ld_bits, q8, ld_q8, rowsum, rowsq, flags, s);
    default: return set_err(GK_EINVAL, "gk_pack_classify: unsupported dtype %d", dtype);
  }
}
