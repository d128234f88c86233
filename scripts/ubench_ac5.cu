// Microbenchmark: the autocorrelation kernel's MMA loops (am_mma5: five tiles per K-block; am_mma: six MMAs) run
// back to back on a resident frame buffer, WITHOUT the staging half of the kernel, at 4 / 8 / 16 warps per SM.
// Answers: how much of the legacy tensor pipe can the loop itself fill, and with how many warps per scheduler?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I neurosync_trainer_lite_b200/csrc \
//        -o scripts/ubench_ac5 scripts/ubench_ac5.cu
#include "../neurosync_trainer_lite_b200/csrc/nsf_autocorr_mma.cu"
#include <cstdio>
namespace nsf { namespace {
template <bool kFive>
__global__ void __launch_bounds__(256, 2) k_loop(int F, int reps, float* out) {
  extern __shared__ __align__(16) __half s_am[];
  const AmGeom geo = am_geom(F);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t region = kExtraFront + 4 * static_cast<size_t>(geo.len);
  __half* copies = s_am + warp * region + kExtraFront;
  for (int i = lane; i < static_cast<int>(region); i += 32) (copies - kExtraFront)[i] = __float2half(0.001f * ((i * 7 + warp) % 13));
  __syncthreads();
  float acc = 0.0f;
  for (int r = 0; r < reps; ++r) {
    float val[kVals];
    if (kFive) am_mma5(copies, geo, lane, val); else am_mma(copies, geo, lane, val);
    acc += val[1] + val[5];
  }
  if (acc == 12345.0f) out[0] = acc;
}
}}
int main() {
  using namespace nsf;
  float* out; cudaMalloc(&out, 4);
  cudaFuncSetAttribute(k_loop<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_loop<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int F : {1470, 266}) {
    const AmGeom geo = am_geom(F);
    for (int warps : {2, 4, 8}) {            // per block, two blocks per SM
      const size_t smem = warps * (kExtraFront + 4 * (size_t)geo.len) * 2;
      const int reps = F > 1000 ? 200 : 1000;
      for (int five = 1; five >= 0; --five) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int it = 0; it < 2; ++it) {
          cudaEventRecord(e0);
          if (five) k_loop<true><<<296, warps * 32, smem>>>(F, reps, out); else k_loop<false><<<296, warps * 32, smem>>>(F, reps, out);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double frames = 296.0 * warps * reps;
        const double hmma = five ? (5.0 * geo.nblk - 12) : 6.0 * geo.nblk;
        printf("F=%d %s loop, %2d warps/SM: %.3f us per frame and SM-warp-slot, %.3f HMMA/clk/SM (pipe peak 0.465), "
               "C2-equivalent %.3f ms for 216060 frames  [%s]\n", F, five ? "five" : "six ", 2 * warps,
               ms * 1e3 / reps, frames * hmma / 148 / (ms * 1.965e6), 216060.0 / frames * ms, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
