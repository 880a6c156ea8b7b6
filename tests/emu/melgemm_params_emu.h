// Host-visible parameter block of the tensor-core mel projection (dl_speech_enhancement_b200/csrc/melgemm.cuh),
// restated for the emulator build, which cannot include <cuda.h>.  The emulator only rejects the call.
#pragma once
namespace spl {
constexpr int kGemmMaxN = 128;
struct MelGemmParams {
  long long rows;
  int n_mels, n_pad;
  int frames;
  int kblocks;
  float eps, log_scale;
  float* out;
};
}  // namespace spl
