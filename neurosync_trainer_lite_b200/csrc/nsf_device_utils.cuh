// Small device helpers shared by the kernel translation units.
#ifndef NSF_DEVICE_UTILS_CUH_
#define NSF_DEVICE_UTILS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace nsf {

// Largest c in [0, n) with off[c] <= g (off is a non-decreasing prefix sum, off[0] == 0, g < off[n]).
// Batches are usually clips of (nearly) equal length, so the proportional guess c = g n / off[n] is
// probed first - two independent loads instead of log2(n) dependent ones (14 for the 10 000-clip
// batches) - and only a miss falls back to the bisection, on the side of the guess that holds g.
__device__ __forceinline__ int find_segment(const int64_t* __restrict__ off, int n, int64_t g) {
  int lo = 0, hi = n;  // invariant: off[lo] <= g < off[hi]
  if (n > 2) {
    const int64_t total = __ldg(off + n);
    int c = total > 0 ? static_cast<int>(static_cast<double>(g) * static_cast<double>(n) / static_cast<double>(total)) : 0;
    c = c < 0 ? 0 : (c > n - 1 ? n - 1 : c);
    const int64_t a = __ldg(off + c), e = __ldg(off + c + 1);
    if (a <= g) {
      if (g < e) return c;
      lo = c + 1;          // off[c + 1] <= g
    } else {
      hi = c;              // g < off[c]
    }
  }
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= g) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// order-preserving float -> uint32 key (so atomicMax works for negative dB values); 0 = "-inf"
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

}  // namespace nsf
#endif  // NSF_DEVICE_UTILS_CUH_
