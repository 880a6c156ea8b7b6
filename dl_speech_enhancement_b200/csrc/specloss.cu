// libspecloss.so -- CUDA (sm_100a) backend of the C ABI in include/specloss.h.
// Kernels: specloss_kernels.cuh.  Argument checking / parameter assembly: specloss_host.inl.
#include "../../include/specloss.h"
#include "specloss_kernels.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define SPL_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) return fail(SPL_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

template <int NFFT, int KIND, bool GRAD, int WIN_T>
int spl_launch_transform(const spl::TransformParams& p, int n_mels, void* stream) {
  using SL = spl::SmemLayout<NFFT, KIND, GRAD>;
  const size_t smem = (size_t)SL::words_per_warp(p.ring_n, n_mels) * 4 * spl::kWarpsPerCta;
  auto kern = spl::transform_kernel<NFFT, KIND, GRAD, WIN_T>;
  if (smem > 227 * 1024) return fail(SPL_E_INVALID, "shared memory %zu B exceeds 227 KB (win/hop too large)", smem);
  // per (instantiation, device): opt in to > 48 KB dynamic shared memory and ask how many CTAs fit an SM
  static thread_local size_t configured[64] = {0};
  static thread_local int ctas_per_sm[64] = {0};
  static thread_local int sm_count[64] = {0};
  int dev = 0;
  SPL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SPL_E_INVALID, "device ordinal %d out of range", dev);
  if (smem != configured[dev]) {
    SPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[dev], kern, spl::kWarpsPerCta * 32, smem));
    SPL_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    configured[dev] = smem;
  }
  // persistent-style launch: at most one resident wave, warps stride over the chunks
  const long long groups = (long long)p.B * p.n_chunks;
  const long long need = (groups + spl::kWarpsPerCta - 1) / spl::kWarpsPerCta;
  const long long wave = (long long)sm_count[dev] * (ctas_per_sm[dev] > 0 ? ctas_per_sm[dev] : 1);
  const unsigned grid = (unsigned)(need < wave ? need : wave);
  kern<<<grid, spl::kWarpsPerCta * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_reduce(const spl::ReduceParams& rp, void* stream) {
  spl::reduce_kernel<<<rp.n_sums, 256, 0, static_cast<cudaStream_t>(stream)>>>(rp);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_finalize(const spl::FinalizeParams& fp, void* stream) {
  spl::finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(fp);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_combine(const spl::CombineParams& cp, void* stream) {
  const long long total = (long long)cp.B * cp.T;
  const unsigned grid = (unsigned)((total + 255) / 256);
  spl::combine_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(cp);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

}  // namespace

#include "specloss_host.inl"
