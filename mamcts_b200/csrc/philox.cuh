// philox.cuh - counter-based Philox4x32-10 (Salmon et al., SC'11) and the engine's action-draw
// stream layout. Device-only; the CPU checker has its own independent copy.
#pragma once
#include <stdint.h>

namespace mamcts {

struct Philox4 {
  uint32_t w[4];
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0);
    const uint32_t lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2);
    const uint32_t lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 out;
  out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
  return out;
}

// 16-bit lane h (0..7) of a block. h is a compile-time constant at every call site.
__device__ __forceinline__ uint32_t philox_half(const Philox4& b, int h) {
  return (h & 1) ? (b.w[h >> 1] >> 16) : (b.w[h >> 1] & 0xFFFFu);
}

// Rejection branch of philox_bounded: taken with probability (65536 mod n) / 65536 (1e-4 for n = 7),
// so it is kept out of line - every call site would otherwise carry its own ten Philox rounds.
static __device__ __noinline__ uint32_t philox_bounded_retry(uint32_t n, uint32_t t, uint32_t pos, uint32_t stream_lo,
                                                      uint32_t stream_hi, uint32_t k0, uint32_t k1) {
  uint32_t r = 0, m, l;
  do {
    const Philox4 b = philox4x32_10(pos, stream_lo, stream_hi, 0x10000u + (r >> 3), k0, k1);
    const uint32_t hh = r & 7u;
    const uint32_t word = (hh >> 1) == 0 ? b.w[0] : (hh >> 1) == 1 ? b.w[1] : (hh >> 1) == 2 ? b.w[2] : b.w[3];
    const uint32_t x = (hh & 1u) ? (word >> 16) : (word & 0xFFFFu);
    m = x * n;
    l = m & 0xFFFFu;
    ++r;
  } while (l < t);
  return m;
}

// Exactly-uniform draws in [0, n) from 16-bit words (Lemire's multiply-shift with rejection): a word x
// gives m = x * n, the draw is m >> 16 unless (m & 0xFFFF) < 65536 mod n, in which case the r-th retry
// reads 16-bit lane (r & 7) of philox({pos, lo, hi, 0x10000 + (r >> 3)}).
//
// The N draws of one step at once: the multiply-shift of every draw runs branch-free,
// the rejection test of all of them is folded into one mask and a single (almost never taken) branch
// per step handles the retries - a third of the instructions of N
// separate draws (the per-draw branch and its reconvergence point dominated the rollout kernel).
// x[i] in: the 16-bit word of draw i; out: the action in [0, n[i]). pos0 + i = stream position of draw i.
template <int N>
__device__ __forceinline__ void philox_bounded_all(uint32_t (&x)[N], const uint32_t (&n)[N], uint32_t pos0,
                                                   uint32_t stream_lo, uint32_t stream_hi, uint32_t k0,
                                                   uint32_t k1) {
  // the pre-filter of the rejection test, once for the step: a draw can only be rejected if the low half of its
  // product is below 65536 mod n < n, so the smallest low half against the largest n decides whether anything has
  // to be looked at (one AND + one MIN per draw instead of AND + compare + merge)
  uint32_t m[N];
  uint32_t low_min = 0xFFFFFFFFu, n_max = 0u;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    m[i] = x[i] * n[i];
    const uint32_t low = m[i] & 0xFFFFu;
    low_min = low < low_min ? low : low_min;
    n_max = n[i] > n_max ? n[i] : n_max;
  }
  if (low_min < n_max) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if ((m[i] & 0xFFFFu) < n[i]) {
        const uint32_t t = 65536u % n[i];
        if ((m[i] & 0xFFFFu) < t) m[i] = philox_bounded_retry(n[i], t, pos0 + i, stream_lo, stream_hi, k0, k1);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = m[i] >> 16;
}

}  // namespace mamcts
