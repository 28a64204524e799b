// host_pack.cpp — host-side bit packing of dense 0/1 fingerprint blocks (caller side of the path: the candidate
// library of gauche/gp.py:169-226 lives in host memory as dense float tensors, 8 KB per 2048-bit molecule).
//
// A dense fp32 block crosses PCIe at ~50 GB/s (2.7 ms per 16 384 x 2048 block on this pool's B200 boxes); packing
// it to 256 B per molecule on the host first runs at ~120 GB/s on the box's 16 cores (1.1 ms) and leaves 4 MB to
// copy.  This is input preparation only: no similarity arithmetic happens on the host, and a block that is not
// exactly 0/1 is reported back so the caller ships it dense.
//
// Threads: a small persistent pool whose workers SLEEP between calls (condition variable).  OpenMP's worker threads
// spin after a parallel region and starved the CUDA driver threads of the same process (the GPU legs measured right
// after a pack call ran 20-30 % slower), so no OpenMP here.
//
// Output format = the packed-bit wire format of gk_bits_popcount / gk_expand_bits_*: bit i of word w is element
// 32*w + i (np.packbits(bitorder="little")), rows padded with zero words up to ld_words.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <pthread.h>
#include <thread>
#include <type_traits>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/gauche_b200.h"

namespace {

template <typename T>
inline int pack_row_scalar(const T* p, int64_t c0, int64_t d, uint32_t* o, int64_t w0, int* cnt) {
  int bad = 0;
  for (int64_t w = w0; w * 32 < d; ++w) {
    uint32_t word = 0;
    const int64_t hi = (w * 32 + 32 < d) ? w * 32 + 32 : d;
    for (int64_t c = w * 32; c < hi; ++c) {
      const T v = p[c];
      if (v == (T)1) word |= 1u << (c - w * 32);
      else if (!(v == (T)0)) bad = 1;      // NaN, counts, fractions: not a bit fingerprint
    }
    o[w] = word;
    *cnt += __builtin_popcount(word);
  }
  (void)c0;
  return bad;
}

#if defined(__x86_64__)
// software prefetch of the fp32 stream (read once): GK_PF_MODE 0 none (hardware prefetcher only), 1 one line 512 B ahead,
// 2 both lines of the iteration 2 KB ahead
#ifndef GK_PF_MODE
#define GK_PF_MODE 0
#endif
#if GK_PF_MODE == 2
#define GK_PREFETCH_F32(q) do { _mm_prefetch(reinterpret_cast<const char*>((q) + 512), _MM_HINT_NTA); _mm_prefetch(reinterpret_cast<const char*>((q) + 528), _MM_HINT_NTA); } while (0)
#elif GK_PF_MODE == 1
#define GK_PREFETCH_F32(q) _mm_prefetch(reinterpret_cast<const char*>((q) + 128), _MM_HINT_NTA)
#else
#define GK_PREFETCH_F32(q) do { } while (0)
#endif
// 32 floats -> one word; any element whose bit pattern is neither +0.0 nor 1.0 sets `bad` (-0.0 counts as bad:
// the block is then simply shipped dense)
__attribute__((target("avx2"))) int pack_row_f32_avx2(const float* p, int64_t d, uint32_t* o, int* cnt) {
  const __m256i one = _mm256_set1_epi32(0x3F800000);
  __m256i badv = _mm256_setzero_si256();
  const int64_t full = d / 32;
  int c = 0;
  for (int64_t w = 0; w < full; ++w) {
    GK_PREFETCH_F32(p + w * 32);
    const __m256i v0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + w * 32));
    const __m256i v1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + w * 32 + 8));
    const __m256i v2 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + w * 32 + 16));
    const __m256i v3 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + w * 32 + 24));
    const __m256i e0 = _mm256_cmpeq_epi32(v0, one), e1 = _mm256_cmpeq_epi32(v1, one);
    const __m256i e2 = _mm256_cmpeq_epi32(v2, one), e3 = _mm256_cmpeq_epi32(v3, one);
    badv = _mm256_or_si256(badv, _mm256_or_si256(_mm256_or_si256(_mm256_andnot_si256(e0, v0), _mm256_andnot_si256(e1, v1)),
                                                 _mm256_or_si256(_mm256_andnot_si256(e2, v2), _mm256_andnot_si256(e3, v3))));
    const uint32_t word = (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(e0)) |
                          ((uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(e1)) << 8) |
                          ((uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(e2)) << 16) |
                          ((uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(e3)) << 24);
    o[w] = word;
    c += __builtin_popcount(word);
  }
  int bad = !_mm256_testz_si256(badv, badv);
  if (full * 32 < d) bad |= pack_row_scalar<float>(p, 0, d, o, full, &c);
  *cnt = c;
  return bad;
}

// 32 doubles -> one word
__attribute__((target("avx2"))) int pack_row_f64_avx2(const double* p, int64_t d, uint32_t* o, int* cnt) {
  const __m256i one = _mm256_set1_epi64x(0x3FF0000000000000LL);
  __m256i badv = _mm256_setzero_si256();
  const int64_t full = d / 32;
  int c = 0;
  for (int64_t w = 0; w < full; ++w) {
    uint32_t word = 0;
#pragma GCC unroll 8
    for (int k = 0; k < 8; ++k) {
      const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + w * 32 + k * 4));
      const __m256i e = _mm256_cmpeq_epi64(v, one);
      badv = _mm256_or_si256(badv, _mm256_andnot_si256(e, v));
      word |= (uint32_t)_mm256_movemask_pd(_mm256_castsi256_pd(e)) << (4 * k);
    }
    o[w] = word;
    c += __builtin_popcount(word);
  }
  int bad = !_mm256_testz_si256(badv, badv);
  if (full * 32 < d) bad |= pack_row_scalar<double>(p, 0, d, o, full, &c);
  *cnt = c;
  return bad;
}

// 32 bytes -> one word
__attribute__((target("avx2"))) int pack_row_u8_avx2(const uint8_t* p, int64_t d, uint32_t* o, int* cnt) {
  const __m256i one = _mm256_set1_epi8(1);
  __m256i badv = _mm256_setzero_si256();
  const int64_t full = d / 32;
  int c = 0;
  for (int64_t w = 0; w < full; ++w) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + w * 32));
    const __m256i e = _mm256_cmpeq_epi8(v, one);
    badv = _mm256_or_si256(badv, _mm256_andnot_si256(e, v));
    const uint32_t word = (uint32_t)_mm256_movemask_epi8(e);
    o[w] = word;
    c += __builtin_popcount(word);
  }
  int bad = !_mm256_testz_si256(badv, badv);
  if (full * 32 < d) bad |= pack_row_scalar<uint8_t>(p, 0, d, o, full, &c);
  *cnt = c;
  return bad;
}

// AVX-512 twins: compare-to-mask gives 16 (fp32) / 8 (fp64) / 64 (u8) result bits per instruction, no movemask
// shuffles; "bad" = some element that is neither the bit pattern of +0 nor of 1.  Two cache lines ahead are
// prefetched: the pack streams 8 KB per fp32 molecule exactly once and is bound by host DRAM bandwidth.
__attribute__((target("avx512f,avx512bw"))) int pack_row_f32_avx512(const float* p, int64_t d, uint32_t* o, int* cnt) {
  const __m512i one = _mm512_set1_epi32(0x3F800000);
  const int64_t full = d / 32;
  unsigned bad_any = 0;
  int c = 0;
  for (int64_t w = 0; w < full; ++w) {
    GK_PREFETCH_F32(p + w * 32);
    const __m512i v0 = _mm512_loadu_si512(p + w * 32), v1 = _mm512_loadu_si512(p + w * 32 + 16);
    const __mmask16 e0 = _mm512_cmpeq_epi32_mask(v0, one), e1 = _mm512_cmpeq_epi32_mask(v1, one);
    bad_any |= (unsigned)(_mm512_test_epi32_mask(v0, v0) & ~e0) | (unsigned)(_mm512_test_epi32_mask(v1, v1) & ~e1);
    const uint32_t word = (uint32_t)e0 | ((uint32_t)e1 << 16);
    o[w] = word;
    c += __builtin_popcount(word);
  }
  int bad = (bad_any & 0xFFFFu) != 0;
  if (full * 32 < d) bad |= pack_row_scalar<float>(p, 0, d, o, full, &c);
  *cnt = c;
  return bad;
}

__attribute__((target("avx512f,avx512bw"))) int pack_row_f64_avx512(const double* p, int64_t d, uint32_t* o, int* cnt) {
  const __m512i one = _mm512_set1_epi64(0x3FF0000000000000LL);
  const int64_t full = d / 32;
  unsigned bad_any = 0;
  int c = 0;
  for (int64_t w = 0; w < full; ++w) {
    uint32_t word = 0;
#pragma GCC unroll 4
    for (int k = 0; k < 4; ++k) {
      _mm_prefetch(reinterpret_cast<const char*>(p + w * 32 + k * 8 + 64), _MM_HINT_NTA);
      const __m512i v = _mm512_loadu_si512(p + w * 32 + k * 8);
      const __mmask8 e = _mm512_cmpeq_epi64_mask(v, one);
      bad_any |= (unsigned)(_mm512_test_epi64_mask(v, v) & ~e) & 0xFFu;
      word |= (uint32_t)e << (8 * k);
    }
    o[w] = word;
    c += __builtin_popcount(word);
  }
  int bad = bad_any != 0;
  if (full * 32 < d) bad |= pack_row_scalar<double>(p, 0, d, o, full, &c);
  *cnt = c;
  return bad;
}

__attribute__((target("avx512f,avx512bw"))) int pack_row_u8_avx512(const uint8_t* p, int64_t d, uint32_t* o, int* cnt) {
  const __m512i one = _mm512_set1_epi8(1);
  const int64_t full = d / 64;            // 64 bytes -> two words
  unsigned long long bad_any = 0;
  int c = 0;
  for (int64_t q = 0; q < full; ++q) {
    const __m512i v = _mm512_loadu_si512(p + q * 64);
    const __mmask64 e = _mm512_cmpeq_epi8_mask(v, one);
    bad_any |= (unsigned long long)(_mm512_test_epi8_mask(v, v) & ~e);
    o[2 * q] = (uint32_t)e;
    o[2 * q + 1] = (uint32_t)((unsigned long long)e >> 32);
    c += __builtin_popcountll((unsigned long long)e);
  }
  int bad = bad_any != 0;
  if (full * 64 < d) bad |= pack_row_scalar<uint8_t>(p, 0, d, o, full * 2, &c);
  *cnt = c;
  return bad;
}
#endif

template <typename T>
int pack_row(const T* p, int64_t d, uint32_t* o, int* cnt, bool avx2, bool avx512) {
#if defined(__x86_64__)
  if (avx512) {
    if (std::is_same<T, float>::value) return pack_row_f32_avx512(reinterpret_cast<const float*>(p), d, o, cnt);
    if (std::is_same<T, double>::value) return pack_row_f64_avx512(reinterpret_cast<const double*>(p), d, o, cnt);
    if (std::is_same<T, uint8_t>::value) return pack_row_u8_avx512(reinterpret_cast<const uint8_t*>(p), d, o, cnt);
  }
  if (avx2) {
    if (std::is_same<T, float>::value) return pack_row_f32_avx2(reinterpret_cast<const float*>(p), d, o, cnt);
    if (std::is_same<T, double>::value) return pack_row_f64_avx2(reinterpret_cast<const double*>(p), d, o, cnt);
    if (std::is_same<T, uint8_t>::value) return pack_row_u8_avx2(reinterpret_cast<const uint8_t*>(p), d, o, cnt);
  }
#endif
  (void)avx2;
  (void)avx512;
  *cnt = 0;
  return pack_row_scalar<T>(p, 0, d, o, 0, cnt);
}

// ---- persistent worker pool: run(fn, parts) calls fn(part) for part in [0, parts) on sleeping workers + the caller
class Pool {
 public:
  // leaked on purpose: detached workers may still sleep on its condition variable when the process exits; a forked
  // child (no threads survive fork) starts with a fresh pool
  static Pool& get() {
    static std::once_flag once;
    std::call_once(once, [] { pthread_atfork(nullptr, nullptr, [] { instance() = nullptr; }); });
    Pool*& p = instance();
    if (!p) p = new Pool;
    return *p;
  }
  void run(const std::function<void(int)>& fn, int parts) {
    std::lock_guard<std::mutex> serial(call_mu_);          // one parallel region at a time
    if (parts <= 1) {
      for (int p = 0; p < parts; ++p) fn(p);
      return;
    }
    grow(parts - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      parts_ = parts;
      next_ = 1;                // part 0 runs on the calling thread
      pending_.store(parts - 1, std::memory_order_relaxed);
      epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    fn(0);
    // the other parts finish within microseconds of this one: spin briefly before sleeping
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(200);
    while (pending_.load(std::memory_order_acquire) != 0 && std::chrono::steady_clock::now() < t_end) cpu_relax();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0; });
    fn_ = nullptr;
  }

 private:
  Pool() = default;
  static Pool*& instance() {
    static Pool* p = nullptr;
    return p;
  }
  void grow(int n) {
    while (n_workers_ < n) {
      std::thread([this] { loop(); }).detach();
      ++n_workers_;
    }
  }
  static void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  // Workers stay hot for spin_us() after a job (back-to-back blocks of a screening run find them awake: a futex
  // wake-up of 15 sleeping threads costs ~0.5 ms on the GPU boxes), then sleep on the condition variable so an idle
  // pool costs nothing.
  // GAUCHE_B200_HOST_SPIN_US overrides the spin window (0 = always sleep)
  static int spin_us() {
    static const int v = [] { const char* e = getenv("GAUCHE_B200_HOST_SPIN_US"); return e ? atoi(e) : 500; }();
    return v;
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us());
      while (epoch_.load(std::memory_order_acquire) == seen && std::chrono::steady_clock::now() < t_end) cpu_relax();
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
      seen = epoch_.load(std::memory_order_acquire);
      while (fn_ && next_ < parts_) {
        const int part = next_++;
        const std::function<void(int)>* fn = fn_;
        lk.unlock();
        (*fn)(part);
        lk.lock();
        if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) done_cv_.notify_one();
      }
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  int n_workers_ = 0;
  const std::function<void(int)>* fn_ = nullptr;
  int parts_ = 0, next_ = 0;
  std::atomic<int> pending_{0};
  std::atomic<uint64_t> epoch_{0};
};

template <typename T>
int pack_block(const T* x, int64_t rows, int64_t d, int64_t ld, uint32_t* bits, int64_t ld_words, int32_t* popcnt,
               int threads) {
#if defined(__x86_64__)
  const bool avx2 = __builtin_cpu_supports("avx2");
  const bool avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
#else
  const bool avx2 = false, avx512 = false;
#endif
  const int64_t words = (d + 31) / 32;
  std::vector<int> bad(threads, 0);
  Pool::get().run(
      [&](int part) {
        const int64_t r0 = rows * part / threads, r1 = rows * (part + 1) / threads;
        int b = 0;
        for (int64_t r = r0; r < r1; ++r) {
          uint32_t* o = bits + r * ld_words;
          int cnt = 0;
          b |= pack_row<T>(x + r * ld, d, o, &cnt, avx2, avx512);
          if (ld_words > words) memset(o + words, 0, (size_t)(ld_words - words) * sizeof(uint32_t));
          popcnt[r] = cnt;
        }
        bad[part] = b;
      },
      threads);
  int bad_any = 0;
  for (int b : bad) bad_any |= b;
  return bad_any;
}

}  // namespace

extern "C" int gk_host_pack_bits(const void* x, int dtype, int64_t rows, int64_t d, int64_t ld, uint32_t* bits,
                                 int64_t ld_words, int32_t* popcnt, int32_t* not_binary, int threads) {
  if (rows < 0 || d <= 0 || ld < d || ld_words * 32 < d || !bits || !popcnt || !not_binary || (rows > 0 && !x))
    return GK_EINVAL;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  if (threads > 256) threads = 256;
  if (threads > rows) threads = rows > 0 ? (int)rows : 1;
  int bad;
  switch (dtype) {
    case GK_F32: bad = pack_block<float>(static_cast<const float*>(x), rows, d, ld, bits, ld_words, popcnt, threads); break;
    case GK_F64: bad = pack_block<double>(static_cast<const double*>(x), rows, d, ld, bits, ld_words, popcnt, threads); break;
    case GK_U8: bad = pack_block<uint8_t>(static_cast<const uint8_t*>(x), rows, d, ld, bits, ld_words, popcnt, threads); break;
    default: return GK_EINVAL;
  }
  *not_binary = bad ? 1 : 0;
  return 0;
}
