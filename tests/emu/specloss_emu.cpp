// CPU SIMT emulation backend of the C ABI in include/specloss.h -- TEST INFRASTRUCTURE ONLY.
// Runs the very same device code (specloss_kernels.cuh) and host logic (specloss_host.inl) as
// libspecloss.so, with every CUDA thread of a warp played by a host thread.  The product never
// loads this library; tests hand it to the host-side engine explicitly.
#define SPECLOSS_EMU 1
#include "../../include/specloss.h"
#include "../../dl_speech_enhancement_b200/csrc/specloss_kernels.cuh"
#include "../../dl_speech_enhancement_b200/csrc/melpower.cuh"
#include "../../dl_speech_enhancement_b200/csrc/transform_eo.cuh"
#include "melgemm_params_emu.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

thread_local EmuWarp* emu_warp = nullptr;
thread_local int emu_lane = 0;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

template <typename F>
void run_warp(F&& body) {     // body(lane), 32 lanes in lock step
  EmuWarp w;
  std::vector<std::thread> th;
  th.reserve(32);
  for (int lane = 0; lane < 32; ++lane)
    th.emplace_back([&, lane] {
      emu_warp = &w;
      emu_lane = lane;
      body(lane);
      w.bar.arrive_and_drop();
    });
  for (auto& t : th) t.join();
}

// emulated device: 5 "SMs", 3 warps per CTA (odd on purpose; fewer warps than frames exercises the stride loop)
int spl_launch_shape(int, size_t, size_t, long long items, int* grid, int* wpc) {
  const int w = 3;
  const long long need = (items + w - 1) / w;
  *grid = (int)std::min<long long>(need, 5);
  *wpc = w;
  return SPL_OK;
}

int spl_fork(void* stream, int n, void** streams) {
  for (int i = 0; i < n; ++i) streams[i] = stream;
  return SPL_OK;
}
int spl_join(void*, int, void**) { return SPL_OK; }

template <typename Load, typename Body>
void run_grid(int grid, int wpc, size_t smem_bytes, Load&& load, Body&& body) {
  const size_t words = smem_bytes / 4;
  std::vector<std::thread> blocks;
  const int par = std::max(1u, std::thread::hardware_concurrency());
  for (int b0 = 0; b0 < grid; b0 += par) {
    blocks.clear();
    for (int block = b0; block < std::min(grid, b0 + par); ++block)
      blocks.emplace_back([&, block] {
        std::vector<float> smem(words, -12345.0f);   // poison: uninitialised reads show up as garbage
        for (int tid = 0; tid < wpc * 32; ++tid) load(smem.data(), tid);
        for (int warp = 0; warp < wpc; ++warp)
          run_warp([&](int lane) { body(smem.data(), block, warp * 32 + lane); });
      });
    for (auto& t : blocks) t.join();
  }
}

template <int NFFT, int KIND, bool GRAD, int WIN_T, bool RING = false>
int spl_launch_transform(const spl::TransformParams& p, int grid, int wpc, size_t smem, void*) {
  run_grid(grid, wpc, smem,
           [&](float* sm, int tid) { spl::cta_load_tables<NFFT, KIND>(p, sm, tid, wpc * 32); },
           [&](float* sm, int block, int tid) { spl::transform_body<NFFT, KIND, GRAD, WIN_T, RING>(p, sm, block, tid, grid, wpc); });
  return SPL_OK;
}

template <int KIND, bool GRAD, int WIN_T>
int spl_launch_transform_eo(const spl::TransformParams& p, const float2* twiddle_eo, const void* mel_entries_eo, int grid, int wpc,
                            size_t smem, void*) {
  run_grid(grid, wpc, smem,
           [&](float* sm, int tid) { spl::cta_load_tables_eo<KIND>(p, twiddle_eo, mel_entries_eo, sm, tid, wpc * 32); },
           [&](float* sm, int block, int tid) { spl::transform_eo_body<KIND, GRAD, WIN_T>(p, sm, block, tid, grid, wpc); });
  return SPL_OK;
}

template <int NFFT>
int spl_launch_spec(const spl::SpecParams& p, int grid, int wpc, size_t smem, void*) {
  const spl::CtaTables ct = spl::cta_tables(NFFT, p.win, spl::kKindStft, 0, 0);
  run_grid(grid, wpc, smem,
           [&](float* sm, int tid) { spl::cta_load_fft_tables<NFFT>(ct, p.twiddle, p.window, p.win, sm, tid, wpc * 32); },
           [&](float* sm, int block, int tid) { spl::spec_body<NFFT>(p, sm, block, tid, grid, wpc); });
  return SPL_OK;
}

template <int NFFT, int KIND>
int spl_launch_specgrad(const spl::SpecGradParams& q, int grid, int wpc, size_t smem, void*) {
  run_grid(grid, wpc, smem,
           [&](float* sm, int tid) { spl::cta_load_tables<NFFT, KIND>(q.t, sm, tid, wpc * 32); },
           [&](float* sm, int block, int tid) { spl::specgrad_body<NFFT, KIND>(q, sm, block, tid, grid, wpc); });
  return SPL_OK;
}

// the tensor-core GEMM has no emulation: tcgen05 / TMA are GPU-only (covered by the -m gpu tests)
int spl_launch_mel_gemm(const float*, const float*, const float*, const float*, int, const spl::MelGemmParams&, void*) {
  return fail(SPL_E_INVALID, "spl_mel_project needs a GPU (tcgen05)");
}

int spl_launch_reduce(const spl::ReduceParams& rp, void*) {
  for (int block = 0; block < rp.n_sums; ++block) {
    double sh[32];
    run_warp([&](int lane) { spl::reduce_body(rp, sh, block, lane, 32); });
  }
  return SPL_OK;
}

int spl_launch_finalize(const spl::FinalizeParams& fp, void*) {
  spl::finalize_body(fp);
  return SPL_OK;
}

int spl_launch_reduce_finalize(const spl::ReduceFinalizeParams& rf, void*) {
  spl_launch_reduce(rf.r, nullptr);
  spl::finalize_body(rf.f);
  return SPL_OK;
}

int spl_shape_dims(long long items, int* grid, int* wpc) {
  *wpc = 3;
  *grid = (int)std::max<long long>(1, std::min<long long>((items + 2) / 3, 5));
  return SPL_OK;
}

int spl_launch_shape_forward(const spl::ShapeParams& p, int grid, int wpc, void*) {
  const bool vec = p.block > 0 && p.block % 4 == 0 && p.T % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y)) & 15) == 0;
  run_grid(grid, wpc, (size_t)wpc * spl::kShapeMaxBlocks * 16, [](float*, int) {},
           [&](float* sm, int block, int tid) {
             if (vec && p.block <= 128) spl::shape_forward_vec_body<1>(p, sm, block, tid, grid, wpc);
             else if (vec)              spl::shape_forward_vec_body<2>(p, sm, block, tid, grid, wpc);
             else                       spl::shape_forward_body(p, sm, block, tid, grid, wpc);
           });
  return SPL_OK;
}

int spl_launch_shape_backward(const spl::ShapeParams& p, int grid, int wpc, void*) {
  run_grid(grid, wpc, (size_t)wpc * spl::kShapeSpan * 4, [](float*, int) {},
           [&](float* sm, int block, int tid) { spl::shape_backward_body(p, sm, block, tid, grid, wpc); });
  return SPL_OK;
}

int spl_launch_shape_finalize(const spl::ShapeFinalizeParams& fp, void*) {
  spl::shape_finalize_body(fp);
  return SPL_OK;
}

int spl_launch_melpow(const spl::MelPowParams& p, int grid, int wpc, size_t smem, void*) {
  run_grid(grid, wpc, smem, [&](float* sm, int tid) { spl::melpow_load_tables(p, sm, tid, wpc * 32); },
           [&](float* sm, int block, int tid) { spl::melpow_body(p, sm, block, tid, grid, wpc); });
  return SPL_OK;
}

// "ranks" of the emulated exchange are host threads of one process calling in concurrently (tests/test_distributed_gloo.py)
int spl_launch_reduce_exchange(const spl::ExchangeParams& ep, void*) {
  spl_launch_reduce(ep.r, nullptr);
  run_warp([&](int lane) { spl::exchange_body(ep, lane); });
  return SPL_OK;
}

int spl_launch_mag_sums(const spl::MagLossParams& p, int grid, int wpc, void*) {
  run_grid(grid, wpc, 0, [](float*, int) {},
           [&](float*, int block, int tid) { spl::mag_sums_body(p, block, tid, grid, wpc); });
  return SPL_OK;
}

int spl_launch_mag_backward(const spl::MagLossParams& p, void*) {
  for (long long t = 0; t < 96; ++t) spl::mag_backward_body(p, t, 96);       // 96 "threads" stride over the elements
  return SPL_OK;
}

int spl_launch_mag_finalize(const spl::MagFinalizeParams& fp, void*) {
  spl::mag_finalize_body(fp);
  return SPL_OK;
}

int spl_launch_combine(const spl::CombineParams& cp, void*) {
  const long long total = (long long)cp.B * ((cp.T + 3) / 4);
  for (long long g = 0; g < total; ++g) spl::combine_body(cp, g);
  return SPL_OK;
}

}  // namespace

#include "../../dl_speech_enhancement_b200/csrc/specloss_host.inl"
