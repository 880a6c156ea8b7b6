// libspecloss.so -- CUDA (sm_100a) backend of the C ABI in include/specloss.h.
// Kernels: specloss_kernels.cuh.  Argument checking / parameter assembly: specloss_host.inl.
#include "../../include/specloss.h"
#include "specloss_kernels.cuh"
#include "melgemm.cuh"
#include "melpower.cuh"
#include "transform_eo.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
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

int device_sm_count(int* sms) {
  static thread_local int cached[64] = {0};
  int dev = 0;
  SPL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SPL_E_INVALID, "device ordinal %d out of range", dev);
  if (!cached[dev]) SPL_CUDA(cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev));
  *sms = cached[dev];
  return SPL_OK;
}

// One persistent CTA per SM.  Warps per CTA: as many as registers (launch bounds) and shared memory allow, but
// no more than needed to give every SM work; the warps then stride over the work items.
int spl_launch_shape(int n_fft, size_t table_bytes, size_t warp_bytes, long long items, int* grid, int* wpc) {
  if (table_bytes + warp_bytes > kMaxSmem)
    return fail(SPL_E_INVALID, "shared memory %zu B (tables) + %zu B (one warp) exceeds 227 KB", table_bytes, warp_bytes);
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc) return rc;
  int w = n_fft == 2048 ? spl::MaxWarps<2048>::value : spl::MaxWarps<1024>::value;
  if (n_fft == 2048 && w > SPL_EO_WARPS) w = SPL_EO_WARPS;      // (variant builds only: SPL_EO_WARPS < 12 also caps the 64-point kernels)
  static const int cap2048 = [] { const char* e = std::getenv("SPECLOSS_WARPS_2048"); return e ? std::atoi(e) : 0; }();
  static const int cap_small = [] { const char* e = std::getenv("SPECLOSS_WARPS_SMALL"); return e ? std::atoi(e) : 0; }();
  const int cap = n_fft == 2048 ? cap2048 : cap_small;          // tuning knobs (profiles/): fewer resident warps per SM
  if (cap > 0 && cap < w) w = cap;
  const int by_smem = (int)((kMaxSmem - table_bytes) / warp_bytes);
  if (by_smem < w) w = by_smem;
  const long long spread = (items + sms - 1) / sms;
  if (spread < w) w = (int)(spread < 1 ? 1 : spread);
  // Small launches: the fewest warps per CTA that still finish in the same number of rounds (17.4 frames per SM take two
  // rounds with 12 warps and with 9; nine warps per round contend less -- profiles/r3p_coresidency.txt: mel 45.2 -> 42.8 us
  // at configs[1] with 10).  Irrelevant for large launches (w stays at its maximum).
  {
    const long long rounds = (spread + w - 1) / w;
    const int balanced = (int)((spread + rounds - 1) / rounds);
    if (balanced >= 1 && balanced < w) w = balanced;
  }
  const long long need = (items + w - 1) / w;
  *grid = (int)(need < sms ? need : sms);
  *wpc = w;
  return SPL_OK;
}

// per (kernel instantiation, device): opt in to the full 227 KB of dynamic shared memory.  `configured` must be a
// static of the calling launch template, i.e. one flag array per kernel.
template <typename K>
int opt_in_smem(K kern, bool* configured) {
  int dev = 0;
  SPL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SPL_E_INVALID, "device ordinal %d out of range", dev);
  if (!configured[dev]) {
    SPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    configured[dev] = true;
  }
  return SPL_OK;
}

// Side streams for the concurrent transforms of one spl_forward() call: created once per (thread, device), never
// destroyed.  Fork/join is event based, so it is also legal while `stream` is being captured into a CUDA graph.
// SPECLOSS_SERIAL=1 keeps everything on the caller's stream.
struct SideStreams {
  cudaStream_t s[SPL_MAX_TRANSFORMS];
  cudaEvent_t fork, join[SPL_MAX_TRANSFORMS];
  bool ready;
};

int side_streams(SideStreams** out) {
  static thread_local SideStreams per_dev[64];
  int dev = 0;
  SPL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SPL_E_INVALID, "device ordinal %d out of range", dev);
  SideStreams& ss = per_dev[dev];
  if (!ss.ready) {
    SPL_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
    for (int i = 1; i < SPL_MAX_TRANSFORMS; ++i) {
      SPL_CUDA(cudaStreamCreateWithFlags(&ss.s[i], cudaStreamNonBlocking));
      SPL_CUDA(cudaEventCreateWithFlags(&ss.join[i], cudaEventDisableTiming));
    }
    ss.ready = true;
  }
  *out = &ss;
  return SPL_OK;
}

bool serial_mode() {
  static const bool v = [] { const char* e = std::getenv("SPECLOSS_SERIAL"); return e && e[0] == '1'; }();
  return v;
}

int spl_fork(void* stream, int n, void** streams) {
  for (int i = 0; i < n; ++i) streams[i] = stream;
  if (n < 2 || serial_mode()) return SPL_OK;
  SideStreams* ss = nullptr;
  int rc = side_streams(&ss);
  if (rc) return rc;
  SPL_CUDA(cudaEventRecord(ss->fork, static_cast<cudaStream_t>(stream)));
  for (int i = 1; i < n; ++i) {
    SPL_CUDA(cudaStreamWaitEvent(ss->s[i], ss->fork, 0));
    streams[i] = ss->s[i];
  }
  return SPL_OK;
}

int spl_join(void* stream, int n, void** streams) {
  if (n < 2 || serial_mode()) return SPL_OK;
  SideStreams* ss = nullptr;
  int rc = side_streams(&ss);
  if (rc) return rc;
  for (int i = 1; i < n; ++i) {
    if (streams[i] == stream) continue;
    SPL_CUDA(cudaEventRecord(ss->join[i], static_cast<cudaStream_t>(streams[i])));
    SPL_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), ss->join[i], 0));
  }
  return SPL_OK;
}

template <int NFFT, int KIND, bool GRAD, int WIN_T, bool RING = false>
int spl_launch_transform(const spl::TransformParams& p, int grid, int wpc, size_t smem, void* stream) {
  auto kern = spl::transform_kernel<NFFT, KIND, GRAD, WIN_T, RING>;
  static thread_local bool configured[64] = {false};
  int rc = opt_in_smem(kern, configured);
  if (rc) return rc;
  kern<<<grid, wpc * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

template <int KIND, bool GRAD, int WIN_T>
int spl_launch_transform_eo(const spl::TransformParams& p, const float2* twiddle_eo, const void* mel_entries_eo, int grid, int wpc,
                            size_t smem, void* stream) {
  auto kern = spl::transform_eo_kernel<KIND, GRAD, WIN_T>;
  static thread_local bool configured[64] = {false};
  int rc = opt_in_smem(kern, configured);
  if (rc) return rc;
  kern<<<grid, wpc * 32, smem, static_cast<cudaStream_t>(stream)>>>(p, twiddle_eo, mel_entries_eo);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

template <int NFFT>
int spl_launch_spec(const spl::SpecParams& p, int grid, int wpc, size_t smem, void* stream) {
  auto kern = spl::spec_kernel<NFFT>;
  static thread_local bool configured[64] = {false};
  int rc = opt_in_smem(kern, configured);
  if (rc) return rc;
  kern<<<grid, wpc * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

template <int NFFT, int KIND>
int spl_launch_specgrad(const spl::SpecGradParams& q, int grid, int wpc, size_t smem, void* stream) {
  auto kern = spl::specgrad_kernel<NFFT, KIND>;
  static thread_local bool configured[64] = {false};
  int rc = opt_in_smem(kern, configured);
  if (rc) return rc;
  kern<<<grid, wpc * 32, smem, static_cast<cudaStream_t>(stream)>>>(q);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

// ---- tensor-core mel projection: TMA descriptors through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tiled_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    SPL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !fn) return fail(SPL_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return SPL_OK;
}

// (rows, ld) fp32 row-major, boxes of box_rows x 32 elements, 128-byte swizzle, zero fill outside
int make_map(EncodeTiledFn enc, CUtensorMap* map, const float* base, long long rows, int ld, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SPL_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SPL_OK;
}

int spl_launch_mel_gemm(const float* amp_hi, const float* amp_lo, const float* w_hi, const float* w_lo, int ld,
                        const spl::MelGemmParams& p, void* stream) {
  EncodeTiledFn enc = nullptr;
  int rc = encode_tiled_fn(&enc);
  if (rc) return rc;
  CUtensorMap m_ah, m_al, m_wh, m_wl;
  if ((rc = make_map(enc, &m_ah, amp_hi, p.rows, ld, spl::kGemmBM))) return rc;
  if ((rc = make_map(enc, &m_al, amp_lo, p.rows, ld, spl::kGemmBM))) return rc;
  if ((rc = make_map(enc, &m_wh, w_hi, p.n_pad, ld, p.n_pad))) return rc;
  if ((rc = make_map(enc, &m_wl, w_lo, p.n_pad, ld, p.n_pad))) return rc;
  const size_t smem = spl::mel_gemm_smem_bytes(p.n_pad);
  // this kernel also has static shared memory (barriers): opt in to exactly what the largest tile set needs
  static thread_local bool configured[64] = {false};
  int dev = 0;
  SPL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SPL_E_INVALID, "device ordinal %d out of range", dev);
  if (!configured[dev]) {
    SPL_CUDA(cudaFuncSetAttribute(spl::mel_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)spl::mel_gemm_smem_bytes(spl::kGemmMaxN)));
    configured[dev] = true;
  }
  const unsigned grid = (unsigned)((p.rows + spl::kGemmBM - 1) / spl::kGemmBM);
  spl::mel_gemm_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(m_ah, m_al, m_wh, m_wl, p);
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

int spl_shape_dims(long long items, int* grid, int* wpc) {
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc) return rc;
  const long long need = (items + 7) / 8, cap = (long long)sms * 8;       // 8 warps per CTA, 8 CTAs per SM
  *grid = (int)(need < 1 ? 1 : (need < cap ? need : cap));
  *wpc = 8;
  return SPL_OK;
}

int spl_launch_shape_forward(const spl::ShapeParams& p, int grid, int wpc, void* stream) {
  // 16-byte loads when blocks, rows and base pointers are 16-byte aligned
  const bool vec = p.block > 0 && p.block % 4 == 0 && p.T % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y)) & 15) == 0;
  if (vec && p.block <= 128) spl::shape_forward_vec_kernel<1><<<grid, wpc * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else if (vec)              spl::shape_forward_vec_kernel<2><<<grid, wpc * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else                       spl::shape_forward_kernel<<<grid, wpc * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_shape_backward(const spl::ShapeParams& p, int grid, int wpc, void* stream) {
  static thread_local bool configured[64] = {false};
  int rc = opt_in_smem(spl::shape_backward_kernel, configured);
  if (rc) return rc;
  spl::shape_backward_kernel<<<grid, wpc * 32, (size_t)wpc * spl::kShapeSpan * 4, static_cast<cudaStream_t>(stream)>>>(p);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_shape_finalize(const spl::ShapeFinalizeParams& fp, void* stream) {
  spl::shape_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(fp);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_melpow(const spl::MelPowParams& p, int grid, int wpc, size_t smem, void* stream) {
  if (smem > kMaxSmem) return fail(SPL_E_INVALID, "power-mel metric: %zu B of shared memory exceed 227 KB", smem);
  static thread_local bool configured[64] = {false};
  int rc = opt_in_smem(spl::melpow_kernel, configured);
  if (rc) return rc;
  spl::melpow_kernel<<<grid, wpc * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_reduce_exchange(const spl::ExchangeParams& ep, void* stream) {
  spl::reduce_exchange_finalize_kernel<<<ep.r.n_sums, 256, 0, static_cast<cudaStream_t>(stream)>>>(ep);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_mag_sums(const spl::MagLossParams& p, int grid, int wpc, void* stream) {
  spl::mag_sums_kernel<<<grid, wpc * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_mag_backward(const spl::MagLossParams& p, void* stream) {
  int grid = 0, wpc = 0;
  int rc = spl_shape_dims((p.n + 2047) / 2048, &grid, &wpc);      // persistent: at most 8 CTAs of 8 warps per SM
  if (rc) return rc;
  spl::mag_backward_kernel<<<grid, wpc * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  SPL_CUDA(cudaGetLastError());
  return SPL_OK;
}

int spl_launch_mag_finalize(const spl::MagFinalizeParams& fp, void* stream) {
  spl::mag_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(fp);
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
