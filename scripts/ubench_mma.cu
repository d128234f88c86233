// Microbenchmark: legacy mma.sync throughput on sm_100a (fp16 m16n8k16, tf32 m16n8k8) and FFMA peak.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
__global__ void k_hmma(float* out, int iters) {
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
  float d[8][4] = {};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[k][0]), "+f"(d[k][1]), "+f"(d[k][2]), "+f"(d[k][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int k = 0; k < 8; ++k) for (int j = 0; j < 4; ++j) s += d[k][j];
  if (s == 12345.f) out[0] = s;
}
__global__ void k_tf32(float* out, int iters) {
  uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
  float d[8][4] = {};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[k][0]), "+f"(d[k][1]), "+f"(d[k][2]), "+f"(d[k][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int k = 0; k < 8; ++k) for (int j = 0; j < 4; ++j) s += d[k][j];
  if (s == 12345.f) out[0] = s;
}
__global__ void k_ffma(float* out, int iters, float x, float y) {
  float d[16];
  for (int k = 0; k < 16; ++k) d[k] = threadIdx.x * 0.001f + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = fmaf(d[k], x, y);
  }
  float s = 0; for (int k = 0; k < 16; ++k) s += d[k];
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
  const int iters = 20000, grid = 148 * 4, block = 256;
  for (int w : {4, 8, 16}) {
    int blk = 32 * w;
    float ms = time_it([&] { k_hmma<<<148 * 2, blk>>>(out, iters); });
    double fl = 148.0 * 2 * w * iters * 8 * (16.0 * 8 * 16 * 2);
    printf("hmma m16n8k16 f16  warps/blk=%d x2 blk/SM: %.1f TFLOP/s (%.3f ms)\n", w, fl / ms / 1e9, ms);
    ms = time_it([&] { k_tf32<<<148 * 2, blk>>>(out, iters); });
    fl = 148.0 * 2 * w * iters * 8 * (16.0 * 8 * 8 * 2);
    printf("mma  m16n8k8  tf32 warps/blk=%d x2 blk/SM: %.1f TFLOP/s (%.3f ms)\n", w, fl / ms / 1e9, ms);
  }
  float ms = time_it([&] { k_ffma<<<grid, block>>>(out, iters, 1.0001f, 0.5f); });
  double fl = (double)grid * block * iters * 16 * 2;
  printf("ffma: %.1f TFLOP/s (%.3f ms)\n", fl / ms / 1e9, ms);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
