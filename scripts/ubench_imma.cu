// Microbenchmark: legacy integer mma.sync (m16n8k32 s8 x s8 -> s32) against fp16 m16n8k16 on sm_100a.
// Question it answers: does an int8-digit formulation of the autocorrelation (exact integer arithmetic,
// K = 32 per instruction) buy tensor throughput over the split-fp16 one on a B200?
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
__global__ void k_imma(float* out, int iters) {
  uint32_t a[4] = {0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u}, b[2] = {0x01010101u, 0x01010101u};
  int d[8][4] = {};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(d[k][0]), "+r"(d[k][1]), "+r"(d[k][2]), "+r"(d[k][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  int s = 0; for (int k = 0; k < 8; ++k) for (int j = 0; j < 4; ++j) s += d[k][j];
  if (s == 12345) out[0] = s;
}
__global__ void k_imma16(float* out, int iters) {
  uint32_t a[2] = {0x01010101u, 0x01010101u}, b[1] = {0x01010101u};
  int d[8][4] = {};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+r"(d[k][0]), "+r"(d[k][1]), "+r"(d[k][2]), "+r"(d[k][3])
                   : "r"(a[0]), "r"(a[1]), "r"(b[0]));
  }
  int s = 0; for (int k = 0; k < 8; ++k) for (int j = 0; j < 4; ++j) s += d[k][j];
  if (s == 12345) out[0] = s;
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
  for (int w : {4, 8, 16}) {
    int blk = 32 * w;
    float ms = time_it([&] { k_hmma<<<148 * 2, blk>>>(out, iters); });
    double n = 148.0 * 2 * w * iters * 8;
    printf("hmma m16n8k16 f16 warps/blk=%d x2 blk/SM: %.1f TFLOP/s  %.2f instr/clk/SM (%.3f ms)\n", w, n * 4096 / ms / 1e9, n / 148 / (ms * 1.965e6), ms);
    ms = time_it([&] { k_imma<<<148 * 2, blk>>>(out, iters); });
    printf("imma m16n8k32 s8  warps/blk=%d x2 blk/SM: %.1f TOP/s    %.2f instr/clk/SM (%.3f ms)\n", w, n * 8192 / ms / 1e9, n / 148 / (ms * 1.965e6), ms);
    ms = time_it([&] { k_imma16<<<148 * 2, blk>>>(out, iters); });
    printf("imma m16n8k16 s8  warps/blk=%d x2 blk/SM: %.1f TOP/s    %.2f instr/clk/SM (%.3f ms)\n", w, n * 4096 / ms / 1e9, n / 148 / (ms * 1.965e6), ms);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
