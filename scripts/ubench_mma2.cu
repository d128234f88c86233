// Microbenchmark 2: what the legacy HMMA pipe of a B200 sustains when the operands are NOT the same registers for
// every instruction (scripts/ubench_mma.cu reuses one A and one B fragment, so the operand-reuse cache feeds the pipe).
//   v0  same A, same B for all MMAs                     (the 553 TFLOP/s figure)
//   v1  A rotates over 4 fragments, B over 8            (register-file operand traffic of a real kernel)
//   v2  v1 + per 6 MMAs one ldmatrix.x4 pair and four 32-bit shared loads that REPLACE fragments (the autocorrelation
//       kernel's mix: operands arrive from shared memory)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mma2 ubench_mma2.cu
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int kVariant>
__global__ void k(float* out, int iters) {
  __shared__ __align__(16) uint32_t sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0x3c003c00u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t a[4][4], b[8][2];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) a[i][j] = 0x3c003c00u + (kVariant ? (i << 8) : 0);
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 2; ++j) b[i][j] = 0x3c003c00u + (kVariant ? (i << 4) : 0);
  float d[6][4] = {};
  const uint32_t saddr = static_cast<uint32_t>(__cvta_generic_to_shared(sm)) + 16u * lane;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int blk = 0; blk < 4; ++blk) {
      if (kVariant == 2) {
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(a[(blk + 1) & 3][0]), "=r"(a[(blk + 1) & 3][1]), "=r"(a[(blk + 1) & 3][2]), "=r"(a[(blk + 1) & 3][3])
                     : "r"(saddr + 512u * blk));
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(a[(blk + 3) & 3][0]), "=r"(a[(blk + 3) & 3][1]), "=r"(a[(blk + 3) & 3][2]), "=r"(a[(blk + 3) & 3][3])
                     : "r"(saddr + 512u * blk + 2048u));
        b[(2 * blk + 2) & 7][0] = sm[lane + 64 * blk]; b[(2 * blk + 2) & 7][1] = sm[lane + 64 * blk + 32];
        b[(2 * blk + 3) & 7][0] = sm[lane + 64 * blk + 1024]; b[(2 * blk + 3) & 7][1] = sm[lane + 64 * blk + 1056];
      }
      const int A0 = kVariant ? blk & 3 : 0, A1 = kVariant ? (blk + 2) & 3 : 0;
      const int B0 = kVariant ? (2 * blk) & 7 : 0, B1 = kVariant ? (2 * blk + 1) & 7 : 0, B2 = kVariant ? (2 * blk + 4) & 7 : 0,
                B3 = kVariant ? (2 * blk + 5) & 7 : 0;
      mma(d[0], a[A0], b[B1][0], b[B1][1]);
      mma(d[1], a[A0], b[B3][0], b[B3][1]);
      mma(d[2], a[A0], b[B0][0], b[B0][1]);
      mma(d[3], a[A0], b[B2][0], b[B2][1]);
      mma(d[4], a[A1], b[B0][0], b[B0][1]);
      mma(d[5], a[A1], b[B2][0], b[B2][1]);
    }
  }
  float s = 0;
  for (int k2 = 0; k2 < 6; ++k2) for (int j = 0; j < 4; ++j) s += d[k2][j];
  if (s == 12345.f) out[0] = s;
}
template <typename F> float time_it(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 4);
  const int iters = 20000;
  for (int w : {4, 8}) {          // warps per block, two blocks per SM: 8 / 16 warps per SM
    const int blk = 32 * w;
    const double fl = 148.0 * 2 * w * iters * 24 * (16.0 * 8 * 16 * 2);
    float ms = time_it([&] { k<0><<<148 * 2, blk>>>(out, iters); });
    printf("v0 same operands           %2d warps/SM: %.1f TFLOP/s\n", 2 * w, fl / ms / 1e9);
    ms = time_it([&] { k<1><<<148 * 2, blk>>>(out, iters); });
    printf("v1 rotating register frags %2d warps/SM: %.1f TFLOP/s\n", 2 * w, fl / ms / 1e9);
    ms = time_it([&] { k<2><<<148 * 2, blk>>>(out, iters); });
    printf("v2 + ldmatrix / LDS feed   %2d warps/SM: %.1f TFLOP/s\n", 2 * w, fl / ms / 1e9);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
