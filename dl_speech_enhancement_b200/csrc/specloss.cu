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

constexpr size_t kMaxSmem = 227 * 1024;

template <int NFFT, int KIND, bool GRAD, int WIN_T>
int spl_launch_transform(const spl::TransformParams& p, int n_mels, void* stream) {
  using SL = spl::SmemLayout<NFFT, KIND, GRAD>;
  constexpr int L = spl::FftGeom<NFFT>::L;
  const spl::CtaTables ct = spl::cta_tables(NFFT, p.win, KIND, L, p.mel_rounds, p.mel_entry_rows);
  const size_t table_bytes = (size_t)ct.total * 4;
  const size_t warp_bytes = (size_t)SL::words_per_warp(p.ring_n, n_mels) * 4;
  if (table_bytes + warp_bytes > kMaxSmem)
    return fail(SPL_E_INVALID, "shared memory %zu B (tables) + %zu B (one warp) exceeds 227 KB", table_bytes, warp_bytes);
  auto kern = spl::transform_kernel<NFFT, KIND, GRAD, WIN_T>;
  // per (instantiation, device): opt in to the full 227 KB of dynamic shared memory
  static thread_local bool configured[64] = {false};
  static thread_local int sm_count[64] = {0};
  int dev = 0;
  SPL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SPL_E_INVALID, "device ordinal %d out of range", dev);
  if (!configured[dev]) {
    SPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    SPL_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    configured[dev] = true;
  }
  // One persistent CTA per SM.  Warps per CTA: as many as registers (launch bounds) and shared memory
  // allow, but no more than needed to give every SM work; warps then stride over the chunks.
  const long long chunks = (long long)p.B * p.n_chunks;
  int wpc = spl::MaxWarps<NFFT>::value;
  const int by_smem = (int)((kMaxSmem - table_bytes) / warp_bytes);
  if (by_smem < wpc) wpc = by_smem;
  const long long spread = (chunks + sm_count[dev] - 1) / sm_count[dev];
  if (spread < wpc) wpc = (int)(spread < 1 ? 1 : spread);
  const long long need = (chunks + wpc - 1) / wpc;
  const unsigned grid = (unsigned)(need < sm_count[dev] ? need : sm_count[dev]);
  const size_t smem = table_bytes + warp_bytes * wpc;
  kern<<<grid, wpc * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
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

int spl_launch_reduce_finalize(const spl::ReduceFinalizeParams& rf, void* stream) {
  spl::reduce_finalize_kernel<<<rf.r.n_sums, 256, 0, static_cast<cudaStream_t>(stream)>>>(rf);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_combine(const spl::CombineParams& cp, void* stream) {
  const long long total = (long long)cp.B * ((cp.T + 3) / 4);
  const unsigned grid = (unsigned)((total + 127) / 128);
  spl::combine_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(cp);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

}  // namespace

#include "specloss_host.inl"
