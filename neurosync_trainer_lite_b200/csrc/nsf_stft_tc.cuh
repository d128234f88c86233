// tcgen05 / TMEM / TMA implementation of the STFT power spectrum ("DFT as GEMM"), sm_100a only.
#ifndef NSF_STFT_TC_CUH_
#define NSF_STFT_TC_CUH_

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "nsf.h"
#include "nsf_internal.h"
#include "nsf_kernels.cuh"

namespace nsf {

constexpr int kTcBScaleExp = 14;  // DFT matrix entries (|.| <= 1) are scaled by 2^14 before the split

struct StftTcHostBlob {
  std::vector<char> bytes;  // fp16 B^T planes: [(chain*2 + part)*2 + hl][np_ld][kp]
  int kp = 0, np_ld = 0, planes = 0;
};

struct StftTcTables {
  const void* bt = nullptr;  // device copy of the blob
  int kp = 0, np_ld = 0, chains = 0;
  int np[2] = {0, 0};
  int col_off[2] = {0, 0};   // column of chain c inside a power row (chain-major layout)
  CUtensorMap map_b;        // box of 128 bin rows (single-CTA kernels)
  CUtensorMap map_b_half;   // box of 64 bin rows (CTA-pair kernel: each CTA loads half of a tile)
  bool ready = false;
};

void build_stft_tc_blob(const Plan& p, StftTcHostBlob* blob);
nsf_status bind_stft_tc_tables(const Plan& p, const StftTcHostBlob& blob, const void* dev_ptr,
                               StftTcTables* out);
// bytes of the folded fp16 operand workspace for `frames` hop-frames
size_t stft_tc_operand_bytes(const Plan& p, int64_t frames);
int launch_stft_tc_fold(cudaStream_t s, const StftTcTables& tc, const DeviceTables& t,
                        const BatchView& b, const float* y, void* operands);
// persistent fused kernel: STFT GEMM -> |.|^2 -> mel -> dB (+ per-clip dB max); writes db [T][n_mels]
int launch_stft_tc_mel(cudaStream_t s, const StftTcTables& tc, const DeviceTables& t, const BatchView& b,
                       void* operands, float* db, uint32_t* dbmax_key);
int launch_stft_tc_gemm(cudaStream_t s, const StftTcTables& tc, const DeviceTables& t,
                        const BatchView& b, void* operands, float* power);

}  // namespace nsf
#endif  // NSF_STFT_TC_CUH_
