// Host-side logic of the C ABI (argument checks, geometry, kernel-parameter assembly), shared by
// the CUDA backend (specloss.cu) and by the CPU SIMT emulator used in the test-suite
// (tests/emu/specloss_emu.cpp).  The includer provides, before including this file:
//   int fail(int code, const char* fmt, ...);
//   int spl_launch_shape(int n_fft, size_t table_bytes, size_t warp_bytes, long long items, int* grid, int* wpc);
//       -- CTAs and warps per CTA of a transform launch over `items` warp work items (the reduction needs the
//          same numbers: one row of partial sums per warp)
//   template <int NFFT, int KIND, bool GRAD, int WIN_T> int spl_launch_transform(const spl::TransformParams&, int grid, int wpc, size_t smem, void* stream);
//   template <int KIND, bool GRAD, int WIN_T> int spl_launch_transform_eo(const spl::TransformParams&, const float2* twiddle_eo,
//                                                                         const void* mel_entries_eo, int grid, int wpc, size_t smem, void* stream);
//   template <int NFFT> int spl_launch_spec(const spl::SpecParams&, int grid, int wpc, size_t smem, void* stream);
//   template <int NFFT, int KIND> int spl_launch_specgrad(const spl::SpecGradParams&, int grid, int wpc, size_t smem, void* stream);
//   int spl_fork(void* stream, int n, void** streams);   -- streams[0] = stream, streams[1..n) run concurrently after
//   int spl_join(void* stream, int n, void** streams);      everything queued on `stream` so far; join = stream waits for all
//   int spl_launch_mel_gemm(const float* amp_hi, const float* amp_lo, const float* w_hi, const float* w_lo, int ld,
//                           const spl::MelGemmParams&, void* stream);
//   int spl_launch_reduce(const spl::ReduceParams&, void* stream);
//   int spl_launch_finalize(const spl::FinalizeParams&, void* stream);
//   int spl_launch_reduce_finalize(const spl::ReduceFinalizeParams&, void* stream);
//   int spl_launch_combine(const spl::CombineParams&, void* stream);
//   int spl_launch_reduce_exchange(const spl::ExchangeParams&, void* stream);
//   int spl_launch_mag_sums(const spl::MagLossParams&, int grid, int wpc, void* stream);      -- grid/wpc from spl_shape_dims
//   int spl_launch_mag_backward(const spl::MagLossParams&, void* stream);
//   int spl_launch_mag_finalize(const spl::MagFinalizeParams&, void* stream);
//   int spl_shape_dims(long long items, int* grid, int* wpc);      -- CTAs x warps of the shape-loss kernels
//   int spl_launch_shape_forward(const spl::ShapeParams&, int grid, int wpc, void* stream);
//   int spl_launch_shape_backward(const spl::ShapeParams&, int grid, int wpc, void* stream);
//   int spl_launch_shape_finalize(const spl::ShapeFinalizeParams&, void* stream);
//   int spl_launch_melpow(const spl::MelPowParams&, int grid, int wpc, size_t smem, void* stream);   -- grid/wpc from spl_shape_dims
namespace {

constexpr long long kMaxPartialRows = 1024 * 32;      // upper bound on grid * warps per CTA on any device

bool supported_nfft(int n) { return n == 512 || n == 1024 || n == 2048; }

int frames_in_flight(int n_fft) { return n_fft == 512 ? 2 : 1; }

// 2048-point losses on the 32 x 32 geometry (transform_eo.cuh) when the caller supplied its tables.  Measured on B200
// (profiles/README.md r3h): the mel loss gains 29 % (3 half-size FFTs instead of 4, 12 resident warps instead of 8), the
// STFT loss loses 3 % -- so by default only the mel loss takes this route.  SPECLOSS_EO_2048 = 0 (never) | mel (default) |
// 1 (mel and STFT), for A/B measurements.
int eo_mode_from_env() {
  const char* e = std::getenv("SPECLOSS_EO_2048");
  if (!e || !e[0] || e[0] == 'm') return 1;
  return e[0] == '0' ? 0 : 2;
}
int eo_mode() {
#ifdef SPECLOSS_EMU
  return eo_mode_from_env();          // the test-suite switches routes between calls
#else
  static const int v = eo_mode_from_env();
  return v;
#endif
}
bool use_eo(const spl_transform* t) {
  if (t->n_fft != 2048 || !t->twiddle_eo) return false;
  if (t->kind == SPL_KIND_MEL) return t->mel_entries_eo && eo_mode() >= 1;
  return eo_mode() >= 2;
}

int check_transform(const spl_transform* t, int B, int T) {
  if (!t) return fail(SPL_E_INVALID, "null transform");
  if (t->kind != SPL_KIND_STFT && t->kind != SPL_KIND_MEL) return fail(SPL_E_INVALID, "kind %d unknown", t->kind);
  if (!supported_nfft(t->n_fft)) return fail(SPL_E_INVALID, "n_fft %d not in {512,1024,2048}", t->n_fft);
  if (t->win < 1 || t->win > t->n_fft) return fail(SPL_E_INVALID, "win %d must be in [1, n_fft=%d]", t->win, t->n_fft);
  if (t->hop < 1) return fail(SPL_E_INVALID, "hop %d < 1", t->hop);      // hop > win is legal (torch.stft): frames with gaps
  if (B < 1) return fail(SPL_E_INVALID, "batch %d < 1", B);
  if (T <= t->n_fft / 2) return fail(SPL_E_INVALID, "reflect padding needs T > n_fft/2 (T=%d, n_fft=%d)", T, t->n_fft);
  if ((long long)B * (1 + T / t->hop) > 0x7fffffffLL) return fail(SPL_E_INVALID, "B * frames exceeds 2^31");
  if (t->kind == SPL_KIND_MEL && (t->n_mels < 2 || t->n_mels > 512))
    return fail(SPL_E_INVALID, "n_mels %d must be in [2, 512]", t->n_mels);
  return SPL_OK;
}

int warp_words(int n_fft, int kind, int n_mels) {
  if (kind == SPL_KIND_STFT)
    return n_fft == 512 ? spl::SmemLayout<512, spl::kKindStft>::words_per_warp(0)
         : n_fft == 1024 ? spl::SmemLayout<1024, spl::kKindStft>::words_per_warp(0)
                         : spl::SmemLayout<2048, spl::kKindStft>::words_per_warp(0);
  return n_fft == 512 ? spl::SmemLayout<512, spl::kKindMel>::words_per_warp(n_mels)
       : n_fft == 1024 ? spl::SmemLayout<1024, spl::kKindMel>::words_per_warp(n_mels)
                       : spl::SmemLayout<2048, spl::kKindMel>::words_per_warp(n_mels);
}

// ---- launch plan of one transform over (B, T): geometry, resident warps and -- STFT loss with gradient -- how many
// consecutive frames a warp overlap-adds in shared memory before the gradient leaves the SM (run_frames) -----------
struct LaunchPlan {
  int grid, wpc;
  int run_frames, runs_per_utt, run_len;
  long long items;               // warp work items
  size_t table_bytes, warp_bytes;
};

// SPECLOSS_RUN_FRAMES: unset / 1 = one gradient slot per frame (the default), k >= 2 = runs of k frames overlap-added in
// shared memory, 0 = choose by size.  Measured on B200 (profiles/r3j_ring.txt, 256 x 4 s): runs of 16 cut the gather from
// 335 to 127 us per STFT resolution and the workspace 3.6x, but the transform kernels pay for the ring traffic (1024: +174 us,
// 512: +224 us; 2048 loses 4 of its 12 resident warps to the ring and doubles) -- break-even at best, so it stays an
// option for memory-bound deployments rather than the default.
int run_frames_from_env() {
  const char* e = std::getenv("SPECLOSS_RUN_FRAMES");
  return e ? std::atoi(e) : 1;
}
int run_frames_request() {
#ifdef SPECLOSS_EMU
  return run_frames_from_env();       // the test-suite switches between calls
#else
  static const int v = run_frames_from_env();
  return v;
#endif
}

// Device-independent part of the plan (spl_geometry_of() must work without a GPU): shared-memory footprint and the run
// length.  The run rule is written for the one device this library targets (B200: 148 SMs).
constexpr int kNominalSms = 148;

void plan_static(const spl_transform* t, int B, int T, bool grad, LaunchPlan* lp) {
  const int fpw = frames_in_flight(t->n_fft);
  const int n_frames = 1 + T / t->hop;
  const bool eo = use_eo(t);
  size_t table_bytes, warp_bytes;
  if (eo) {
    table_bytes = (size_t)spl::cta_tables_eo(t->win, t->kind, t->mel_rounds, t->mel_entry_rows).total * 4;
    warp_bytes = (size_t)spl::eo_words_per_warp(t->kind, t->n_mels) * 4;
  } else {
    table_bytes = (size_t)spl::cta_tables(t->n_fft, t->win, t->kind, t->mel_rounds, t->mel_entry_rows).total * 4;
    warp_bytes = (size_t)warp_words(t->n_fft, t->kind, t->n_mels) * 4;
  }
  lp->grid = lp->wpc = 0;
  lp->run_frames = 1; lp->runs_per_utt = n_frames; lp->run_len = t->win;
  lp->table_bytes = table_bytes; lp->warp_bytes = warp_bytes;
  lp->items = ((long long)B * n_frames + fpw - 1) / fpw;
  if (!grad || t->kind != SPL_KIND_STFT || eo || t->hop >= t->win) return;     // no overlap: nothing for a ring to add up
  // Runs: the gradient of a run of m frames is (m - 1) hop + win taps instead of m win.  Worth it when the per-frame
  // slots no longer fit the L2 (they then make a round trip through HBM) and only while every resident warp still gets
  // several runs; the ring costs win * 8 bytes of shared memory per frame in flight.
  const size_t ring_bytes = (size_t)fpw * 8 * ((t->win + 1) & ~1);
  const int want = run_frames_request();
  if (want == 1 || table_bytes + warp_bytes + ring_bytes > 227 * 1024) return;
  const int reg_warps = t->n_fft == 2048 ? 12 : 16;                              // MaxWarps<NFFT> of the kernels
  const int smem_warps = (int)((227 * 1024 - table_bytes) / (warp_bytes + ring_bytes));
  const long long dev_warps = (long long)kNominalSms * (smem_warps < reg_warps ? smem_warps : reg_warps);
  const bool big = (long long)B * n_frames * t->win * 8 > (48ll << 20);          // slots of this transform vs the L2
  const int cand[4] = {16, 8, 4, 2};
  for (int c = 0; c < 4; ++c) {
    const int m = want >= 2 ? want : cand[c];
    if (m > n_frames) { if (want >= 2) break; continue; }
    const int runs = (n_frames + m - 1) / m;
    const long long items = ((long long)B * runs + fpw - 1) / fpw;
    if (want >= 2 || (big && items >= 4 * dev_warps)) {
      lp->run_frames = m; lp->runs_per_utt = runs; lp->run_len = (m - 1) * t->hop + t->win;
      lp->items = items; lp->warp_bytes = warp_bytes + ring_bytes;
      return;
    }
  }
}

int plan_transform(const spl_transform* t, int B, int T, bool grad, LaunchPlan* lp) {
  plan_static(t, B, T, grad, lp);
  return spl_launch_shape(t->n_fft, lp->table_bytes, lp->warp_bytes, lp->items, &lp->grid, &lp->wpc);
}

void geometry(const spl_transform* t, int B, int T, spl_geometry* g) {
  std::memset(g, 0, sizeof(*g));
  g->n_frames = 1 + T / t->hop;
  g->n_bins = t->n_fft / 2 + 1;
  g->n_sums = t->kind == SPL_KIND_STFT ? 3 : 1;
  LaunchPlan fwd, grd;
  plan_static(t, B, T, false, &fwd);
  plan_static(t, B, T, true, &grd);
  // one row per warp of the launch: grid * wpc < items + wpc (the last CTA may hold idle warps, which still write a
  // row of zeros), wpc <= 32; the forward-only launch (frames) has at least as many items as the gradient launch (runs)
  const long long items = fwd.items > grd.items ? fwd.items : grd.items;
  g->partial_count = (int64_t)((items < kMaxPartialRows ? items : kMaxPartialRows) + 32) * g->n_sums;
  g->gframe_bytes = (int64_t)B * grd.runs_per_utt * grd.run_len * (t->kind == SPL_KIND_STFT ? 8 : 4);
  g->smem_table_bytes = (int64_t)grd.table_bytes;
  g->smem_warp_bytes = (int64_t)grd.warp_bytes;
}

// CTAs x warps of the launch of transform t over (B, T); grad = the launch that also writes the gradient
int shape_of(const spl_transform* t, int B, int T, bool grad, int* grid, int* wpc) {
  LaunchPlan lp;
  int rc = plan_transform(t, B, T, grad, &lp);
  if (rc) return rc;
  *grid = lp.grid; *wpc = lp.wpc;
  if ((long long)lp.grid * lp.wpc > kMaxPartialRows + 32)
    return fail(SPL_E_INVALID, "launch of %d x %d warps exceeds the partial-sum rows", lp.grid, lp.wpc);
  return SPL_OK;
}

template <int NFFT, int WIN_T>
int launch_win(const spl::TransformParams& p, int kind, bool grad, int grid, int wpc, size_t smem, void* s) {
  if (kind == SPL_KIND_STFT && grad && p.run_frames > 1)         // overlap-add ring: its own instantiation
    return spl_launch_transform<NFFT, spl::kKindStft, true, WIN_T, true>(p, grid, wpc, smem, s);
  if (kind == SPL_KIND_STFT)
    return grad ? spl_launch_transform<NFFT, spl::kKindStft, true, WIN_T>(p, grid, wpc, smem, s)
                : spl_launch_transform<NFFT, spl::kKindStft, false, WIN_T>(p, grid, wpc, smem, s);
  return grad ? spl_launch_transform<NFFT, spl::kKindMel, true, WIN_T>(p, grid, wpc, smem, s)
              : spl_launch_transform<NFFT, spl::kKindMel, false, WIN_T>(p, grid, wpc, smem, s);
}

// Window lengths of the shipped configurations get kernels with the window support known at compile
// time (zero taps pruned); every other (n_fft, win) pair runs the generic kernel of that n_fft.
int launch_any(const spl::TransformParams& p, int n_fft, int kind, bool grad, int grid, int wpc, size_t smem, void* s) {
  if (n_fft == 1024) return p.win == 600 ? launch_win<1024, 600>(p, kind, grad, grid, wpc, smem, s) : launch_win<1024, 0>(p, kind, grad, grid, wpc, smem, s);
  if (n_fft == 512) return p.win == 240 ? launch_win<512, 240>(p, kind, grad, grid, wpc, smem, s) : launch_win<512, 0>(p, kind, grad, grid, wpc, smem, s);
  if (p.win == 1200) return launch_win<2048, 1200>(p, kind, grad, grid, wpc, smem, s);
  if (p.win == 2048) return launch_win<2048, 2048>(p, kind, grad, grid, wpc, smem, s);
  return launch_win<2048, 0>(p, kind, grad, grid, wpc, smem, s);
}

template <int WIN_T>
int launch_eo_win(const spl::TransformParams& p, const spl_transform* t, bool grad, int grid, int wpc, size_t smem, void* s) {
  const float2* tw = reinterpret_cast<const float2*>(t->twiddle_eo);
  if (t->kind == SPL_KIND_STFT)
    return grad ? spl_launch_transform_eo<spl::kKindStft, true, WIN_T>(p, tw, nullptr, grid, wpc, smem, s)
                : spl_launch_transform_eo<spl::kKindStft, false, WIN_T>(p, tw, nullptr, grid, wpc, smem, s);
  return grad ? spl_launch_transform_eo<spl::kKindMel, true, WIN_T>(p, tw, t->mel_entries_eo, grid, wpc, smem, s)
              : spl_launch_transform_eo<spl::kKindMel, false, WIN_T>(p, tw, t->mel_entries_eo, grid, wpc, smem, s);
}

int launch_eo(const spl::TransformParams& p, const spl_transform* t, bool grad, int grid, int wpc, size_t smem, void* s) {
  if (p.win == 1200) return launch_eo_win<1200>(p, t, grad, grid, wpc, smem, s);
  if (p.win == 2048) return launch_eo_win<2048>(p, t, grad, grid, wpc, smem, s);
  return launch_eo_win<0>(p, t, grad, grid, wpc, smem, s);
}

}  // namespace

extern "C" {

int32_t spl_fill_twiddle_eo(float* host_out) {
  if (!host_out) return fail(SPL_E_INVALID, "spl_fill_twiddle_eo: null output");
  const double two_pi = 6.283185307179586476925286766559;
  for (int k2 = 0; k2 < 32; ++k2)
    for (int n1 = 0; n1 < 32; ++n1) {
      const double th = two_pi * (double)((n1 * k2) % 1024) / 1024.0;
      host_out[2 * (k2 * 32 + n1)] = (float)std::cos(th);
      host_out[2 * (k2 * 32 + n1) + 1] = (float)(-std::sin(th));
    }
  float* wk = host_out + 2 * 1024;
  for (int k = 0; k < spl::eo::WK; ++k) {
    const double th = two_pi * (double)k / 2048.0;
    wk[2 * k] = k <= 512 ? (float)std::cos(th) : 0.f;
    wk[2 * k + 1] = k <= 512 ? (float)(-std::sin(th)) : 0.f;
  }
  return SPL_OK;
}

int32_t spl_abi_version(void) { return SPL_ABI_VERSION; }

const char* spl_last_error(void) { return g_err; }

int32_t spl_fill_twiddle(int32_t n_fft, float* host_out) {
  if (!supported_nfft(n_fft) || !host_out) return fail(SPL_E_INVALID, "spl_fill_twiddle: n_fft %d unsupported or null output", n_fft);
  const int L = n_fft == 512 ? 16 : 32, R = n_fft / L;
  const double two_pi = 6.283185307179586476925286766559;
  for (int k2 = 0; k2 < R; ++k2)
    for (int n1 = 0; n1 < L; ++n1) {
      const int e = (n1 * k2) % n_fft;
      const double th = two_pi * (double)e / (double)n_fft;
      host_out[2 * (k2 * L + n1)] = (float)std::cos(th);
      host_out[2 * (k2 * L + n1) + 1] = (float)(-std::sin(th));
    }
  return SPL_OK;
}

int32_t spl_geometry_of(const spl_transform* t, int32_t B, int32_t T, spl_geometry* out) {
  if (!out) return fail(SPL_E_INVALID, "null geometry output");
  int rc = check_transform(t, B, T);
  if (rc) return rc;
  geometry(t, B, T, out);
  return SPL_OK;
}

int32_t spl_forward(const spl_transform* ts, int32_t n, const float* x, const float* y,
                    int32_t B, int32_t T, void* stream) {
  if (n < 1 || n > SPL_MAX_TRANSFORMS) return fail(SPL_E_INVALID, "n=%d transforms not in [1,%d]", n, SPL_MAX_TRANSFORMS);
  if (!x || !y) return fail(SPL_E_INVALID, "null input");
  for (int r = 0; r < n; ++r) {
    const spl_transform* t = ts + r;
    int rc = check_transform(t, B, T);
    if (rc) return rc;
    if (!t->window || !t->twiddle || !t->partials) return fail(SPL_E_INVALID, "transform %d: null window/twiddle/partials", r);
    if (t->kind == SPL_KIND_MEL && (!t->mel_tasks || !t->mel_entries || t->mel_rounds < 1 || t->mel_entry_rows < 1 || !t->bin_tab))
      return fail(SPL_E_INVALID, "transform %d: null mel table", r);
  }
  // The transforms are independent: each runs on its own stream, most expensive first, so that the CTAs of the
  // next kernel take over the SMs one by one as the CTAs of the previous one retire (every kernel is one
  // persistent CTA per SM; back to back on ONE stream each of them would drain completely before the next starts).
  int order[SPL_MAX_TRANSFORMS];
  double cost[SPL_MAX_TRANSFORMS];
  for (int r = 0; r < n; ++r) {
    order[r] = r;
    cost[r] = (double)(1 + T / ts[r].hop) * ts[r].n_fft * (ts[r].kind == SPL_KIND_MEL ? 1.4 : 1.0);
  }
  for (int i = 1; i < n; ++i)
    for (int j = i; j > 0 && cost[order[j]] > cost[order[j - 1]]; --j) { const int tmp = order[j]; order[j] = order[j - 1]; order[j - 1] = tmp; }
  void* streams[SPL_MAX_TRANSFORMS];
  int rc = spl_fork(stream, n, streams);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    const spl_transform* t = ts + order[i];
    LaunchPlan lp;
    const bool grad = t->gframes != nullptr;
    rc = plan_transform(t, B, T, grad, &lp);
    if (rc) break;
    const int grid = lp.grid, wpc = lp.wpc;
    spl::TransformParams p;
    std::memset(&p, 0, sizeof(p));
    p.x = x; p.y = y; p.B = B; p.T = T;
    p.hop = t->hop; p.win = t->win; p.left = (t->n_fft - t->win) / 2;
    p.n_frames = 1 + T / t->hop;
    p.eps = t->eps; p.window = t->window; p.twiddle = reinterpret_cast<const float2*>(t->twiddle);
    p.partials = t->partials; p.gframes = t->gframes;
    p.run_frames = lp.run_frames; p.runs_per_utt = lp.runs_per_utt; p.run_len = lp.run_len;
    p.n_mels = t->kind == SPL_KIND_MEL ? t->n_mels : 0;
    p.inv_ln_base = t->inv_ln_base;
    p.mel_tasks = t->mel_tasks; p.mel_entries = t->mel_entries;
    p.mel_rounds = t->kind == SPL_KIND_MEL ? t->mel_rounds : 0;
    p.mel_entry_rows = t->kind == SPL_KIND_MEL ? t->mel_entry_rows : 0;
    p.bin_tab = t->bin_tab;
    const size_t smem = lp.table_bytes + lp.warp_bytes * wpc;
    rc = use_eo(t) ? launch_eo(p, t, t->gframes != nullptr, grid, wpc, smem, streams[i])
                   : launch_any(p, t->n_fft, t->kind, t->gframes != nullptr, grid, wpc, smem, streams[i]);
    if (rc) break;
  }
  const int rc_join = spl_join(stream, n, streams);      // always re-join, also after a failed launch
  return rc ? rc : rc_join;
}

int32_t spl_spectrogram(const float* x, int32_t B, int32_t T, int32_t n_fft, int32_t hop, int32_t win,
                        const float* window, const float* twiddle, float eps, float* out, float* out_lo, int32_t ld,
                        void* stream) {
  spl_transform t;
  std::memset(&t, 0, sizeof(t));
  t.kind = SPL_KIND_STFT; t.n_fft = n_fft; t.hop = hop; t.win = win;
  if (!supported_nfft(n_fft)) return fail(SPL_E_INVALID, "n_fft %d not in {512,1024,2048}", n_fft);
  if (win < 1 || win > n_fft) return fail(SPL_E_INVALID, "win %d must be in [1, n_fft=%d]", win, n_fft);
  if (hop < 1) return fail(SPL_E_INVALID, "hop %d < 1", hop);
  if (B < 1) return fail(SPL_E_INVALID, "batch %d < 1", B);
  if (T <= n_fft / 2) return fail(SPL_E_INVALID, "reflect padding needs T > n_fft/2 (T=%d, n_fft=%d)", T, n_fft);
  if (!x || !window || !twiddle || !out) return fail(SPL_E_INVALID, "spl_spectrogram: null pointer");
  if (ld < n_fft / 2 + 1) return fail(SPL_E_INVALID, "ld %d < n_fft/2+1", ld);
  const int n_frames = 1 + T / hop, n_pairs = (n_frames + 1) / 2;
  if ((long long)B * n_frames > 0x7fffffffLL) return fail(SPL_E_INVALID, "B * frames exceeds 2^31");
  const spl::CtaTables ct = spl::cta_tables(n_fft, win, spl::kKindStft, 0, 0);
  const size_t table_bytes = (size_t)ct.total * 4, warp_bytes = (size_t)warp_words(n_fft, SPL_KIND_STFT, 0) * 4;
  const int fpw = frames_in_flight(n_fft);
  int grid = 0, wpc = 0;
  int rc = spl_launch_shape(n_fft, table_bytes, warp_bytes, ((long long)B * n_pairs + fpw - 1) / fpw, &grid, &wpc);
  if (rc) return rc;
  spl::SpecParams p;
  std::memset(&p, 0, sizeof(p));
  p.x = x; p.B = B; p.T = T; p.hop = hop; p.win = win; p.left = (n_fft - win) / 2; p.n_frames = n_frames;
  p.n_pairs = n_pairs; p.eps = eps; p.window = window; p.twiddle = reinterpret_cast<const float2*>(twiddle);
  p.out = out; p.out_lo = out_lo; p.ld = ld;
  const size_t smem = table_bytes + warp_bytes * wpc;
  if (n_fft == 512) return spl_launch_spec<512>(p, grid, wpc, smem, stream);
  if (n_fft == 1024) return spl_launch_spec<1024>(p, grid, wpc, smem, stream);
  return spl_launch_spec<2048>(p, grid, wpc, smem, stream);
}

int32_t spl_spectrogram_backward(const spl_transform* t, const float* x, int32_t B, int32_t T,
                                 const float* g, int32_t ld, float* dx, void* stream) {
  if (!t) return fail(SPL_E_INVALID, "null transform");
  if (t->kind != SPL_KIND_STFT && t->kind != SPL_KIND_MEL) return fail(SPL_E_INVALID, "kind %d unknown", t->kind);
  const int n_fft = t->n_fft, hop = t->hop, win = t->win;
  if (!supported_nfft(n_fft)) return fail(SPL_E_INVALID, "n_fft %d not in {512,1024,2048}", n_fft);
  if (win < 1 || win > n_fft) return fail(SPL_E_INVALID, "win %d must be in [1, n_fft=%d]", win, n_fft);
  if (hop < 1) return fail(SPL_E_INVALID, "hop %d < 1", hop);
  if (B < 1) return fail(SPL_E_INVALID, "batch %d < 1", B);
  if (T <= n_fft / 2) return fail(SPL_E_INVALID, "reflect padding needs T > n_fft/2 (T=%d, n_fft=%d)", T, n_fft);
  if (!x || !g || !dx || !t->window || !t->twiddle || !t->gframes)
    return fail(SPL_E_INVALID, "spl_spectrogram_backward: null pointer (x, g, dx, window, twiddle or gframes workspace)");
  const bool mel = t->kind == SPL_KIND_MEL;
  if (mel) {
    if (t->n_mels < 2 || t->n_mels > 512) return fail(SPL_E_INVALID, "n_mels %d must be in [2, 512]", t->n_mels);
    if (!t->mel_tasks || !t->mel_entries || t->mel_rounds < 1 || t->mel_entry_rows < 1 || !t->bin_tab)
      return fail(SPL_E_INVALID, "spl_spectrogram_backward: null mel table");
  } else if (ld < n_fft / 2 + 1) {
    return fail(SPL_E_INVALID, "ld %d < n_fft/2+1", ld);
  }
  const int n_frames = 1 + T / hop, n_pairs = (n_frames + 1) / 2;
  if ((long long)B * n_frames > 0x7fffffffLL) return fail(SPL_E_INVALID, "B * frames exceeds 2^31");
  const spl::CtaTables ct = spl::cta_tables(n_fft, win, t->kind, t->mel_rounds, t->mel_entry_rows);
  const size_t table_bytes = (size_t)ct.total * 4, warp_bytes = (size_t)warp_words(n_fft, t->kind, t->n_mels) * 4;
  const int fpw = frames_in_flight(n_fft);
  int grid = 0, wpc = 0;
  int rc = spl_launch_shape(n_fft, table_bytes, warp_bytes, ((long long)B * n_pairs + fpw - 1) / fpw, &grid, &wpc);
  if (rc) return rc;
  spl::SpecGradParams q;
  std::memset(&q, 0, sizeof(q));
  spl::TransformParams& p = q.t;
  p.x = x; p.B = B; p.T = T; p.hop = hop; p.win = win; p.left = (n_fft - win) / 2; p.n_frames = n_frames;
  p.eps = t->eps; p.window = t->window; p.twiddle = reinterpret_cast<const float2*>(t->twiddle);
  p.gframes = t->gframes;
  p.n_mels = mel ? t->n_mels : 0;
  p.inv_ln_base = t->inv_ln_base;
  p.mel_tasks = t->mel_tasks; p.mel_entries = t->mel_entries;
  p.mel_rounds = mel ? t->mel_rounds : 0;
  p.mel_entry_rows = mel ? t->mel_entry_rows : 0;
  p.bin_tab = t->bin_tab;
  q.n_pairs = n_pairs; q.g = g; q.ld = ld;
  const size_t smem = table_bytes + warp_bytes * wpc;
  if (n_fft == 512) rc = mel ? spl_launch_specgrad<512, spl::kKindMel>(q, grid, wpc, smem, stream) : spl_launch_specgrad<512, spl::kKindStft>(q, grid, wpc, smem, stream);
  else if (n_fft == 1024) rc = mel ? spl_launch_specgrad<1024, spl::kKindMel>(q, grid, wpc, smem, stream) : spl_launch_specgrad<1024, spl::kKindStft>(q, grid, wpc, smem, stream);
  else rc = mel ? spl_launch_specgrad<2048, spl::kKindMel>(q, grid, wpc, smem, stream) : spl_launch_specgrad<2048, spl::kKindStft>(q, grid, wpc, smem, stream);
  if (rc) return rc;
  spl::CombineParams cp;
  std::memset(&cp, 0, sizeof(cp));
  cp.n = 1;
  spl::CombineEntry& e = cp.e[0];
  e.frames = t->gframes; e.kind = SPL_KIND_MEL /* one float per tap */; e.half = n_fft / 2; e.hop = hop; e.win = win;
  e.left = (n_fft - win) / 2; e.n_frames = n_frames;
  cp.dx = dx; cp.B = B; cp.T = T; cp.unit = 1;
  return spl_launch_combine(cp, stream);
}

int32_t spl_mel_project(const float* amp_hi, const float* amp_lo, int64_t rows, int32_t ld,
                        const float* w_hi, const float* w_lo, int32_t n_mels, int32_t n_pad, int32_t frames,
                        float eps, float log_scale, float* out, void* stream) {
  if (!amp_hi || !amp_lo || !w_hi || !w_lo || !out) return fail(SPL_E_INVALID, "spl_mel_project: null pointer");
  if (rows < 1 || rows > 0x7fffff00LL) return fail(SPL_E_INVALID, "rows %lld out of range", (long long)rows);
  if (frames < 1 || rows % frames) return fail(SPL_E_INVALID, "rows %lld is not a multiple of frames %d", (long long)rows, frames);
  if (ld < 32 || ld % 32) return fail(SPL_E_INVALID, "ld %d must be a positive multiple of 32", ld);
  if (n_mels < 1 || n_pad < n_mels || n_pad % 16 || n_pad > spl::kGemmMaxN)
    return fail(SPL_E_INVALID, "n_mels %d / n_pad %d: n_pad must be a multiple of 16 in [n_mels, %d]", n_mels, n_pad, spl::kGemmMaxN);
  if (((uintptr_t)amp_hi | (uintptr_t)amp_lo | (uintptr_t)w_hi | (uintptr_t)w_lo) & 15)
    return fail(SPL_E_INVALID, "spl_mel_project: operands must be 16-byte aligned");
  spl::MelGemmParams p;
  std::memset(&p, 0, sizeof(p));
  p.rows = rows; p.n_mels = n_mels; p.n_pad = n_pad; p.frames = frames; p.kblocks = ld / 32;
  p.eps = eps; p.log_scale = log_scale; p.out = out;
  return spl_launch_mel_gemm(amp_hi, amp_lo, w_hi, w_lo, ld, p, stream);
}

static int build_reduce(const spl_transform* ts, int n, int B, int T, double* sums, spl::ReduceParams* rp) {
  std::memset(rp, 0, sizeof(*rp));
  int k = 0;
  for (int r = 0; r < n; ++r) {
    int rc = check_transform(ts + r, B, T);
    if (rc) return rc;
    if (!ts[r].partials) return fail(SPL_E_INVALID, "transform %d: null partials", r);
    int grid = 0, wpc = 0;
    rc = shape_of(ts + r, B, T, ts[r].gframes != nullptr, &grid, &wpc);
    if (rc) return rc;
    const int n_sums = ts[r].kind == SPL_KIND_STFT ? 3 : 1;
    for (int j = 0; j < n_sums; ++j) {
      if (k >= spl::kMaxSums) return fail(SPL_E_INVALID, "too many sums");
      rp->base[k] = ts[r].partials + j;
      rp->stride[k] = n_sums;
      rp->count[k] = grid * wpc;
      ++k;
    }
  }
  rp->n_sums = k;
  rp->out = sums;
  return SPL_OK;
}

static int build_finalize(const spl_transform* ts, int n, const double* sums, int64_t B_global, int T,
                          float* sc, float* mag, float* mel, float* coefs, spl::FinalizeParams* fp) {
  if (B_global < 1) return fail(SPL_E_INVALID, "B_global %lld < 1", (long long)B_global);
  std::memset(fp, 0, sizeof(*fp));
  fp->n = n;
  int ofs = 0;
  for (int r = 0; r < n; ++r) {
    int rc = check_transform(ts + r, 1, T);
    if (rc) return rc;
    const double frames = 1 + T / ts[r].hop;
    fp->kind[r] = ts[r].kind;
    fp->sum_ofs[r] = ofs;
    if (ts[r].kind == SPL_KIND_STFT) { fp->count[r] = (double)B_global * frames * (ts[r].n_fft / 2 + 1); ofs += 3; }
    else { fp->count[r] = (double)B_global * frames * ts[r].n_mels; ofs += 1; }
  }
  fp->sums = sums; fp->sc = sc; fp->mag = mag; fp->mel = mel; fp->coefs = coefs;
  return SPL_OK;
}

int32_t spl_reduce(const spl_transform* ts, int32_t n, int32_t B, int32_t T, double* sums, void* stream) {
  if (n < 1 || n > SPL_MAX_TRANSFORMS || !sums) return fail(SPL_E_INVALID, "spl_reduce: bad n or null sums");
  spl::ReduceParams rp;
  int rc = build_reduce(ts, n, B, T, sums, &rp);
  if (rc) return rc;
  return spl_launch_reduce(rp, stream);
}

int32_t spl_finalize(const spl_transform* ts, int32_t n, const double* sums, int64_t B_global, int32_t T,
                     float* sc, float* mag, float* mel, float* coefs, void* stream) {
  if (n < 1 || n > SPL_MAX_TRANSFORMS || !sums || !coefs) return fail(SPL_E_INVALID, "spl_finalize: bad n or null sums/coefs");
  spl::FinalizeParams fp;
  int rc = build_finalize(ts, n, sums, B_global, T, sc, mag, mel, coefs, &fp);
  if (rc) return rc;
  return spl_launch_finalize(fp, stream);
}

int32_t spl_reduce_finalize(const spl_transform* ts, int32_t n, int32_t B, int32_t T, double* sums,
                            float* sc, float* mag, float* mel, float* coefs, uint32_t* counter, void* stream) {
  if (n < 1 || n > SPL_MAX_TRANSFORMS || !sums || !coefs || !counter)
    return fail(SPL_E_INVALID, "spl_reduce_finalize: bad n or null sums/coefs/counter");
  spl::ReduceFinalizeParams rf;
  int rc = build_reduce(ts, n, B, T, sums, &rf.r);
  if (rc) return rc;
  rc = build_finalize(ts, n, sums, B, T, sc, mag, mel, coefs, &rf.f);
  if (rc) return rc;
  rf.counter = counter;
  return spl_launch_reduce_finalize(rf, stream);
}

int64_t spl_exchange_buffer_bytes(void) { return (int64_t)spl::kExchangeBytes; }

int32_t spl_reduce_exchange_finalize(const spl_transform* ts, int32_t n, int32_t B, int32_t T, int64_t B_global,
                                     double* sums_local, double* sums_global, int32_t rank, int32_t world,
                                     void* const* peer_bufs, uint32_t* state, int64_t timeout_ns, uint32_t* error_flag,
                                     float* sc, float* mag, float* mel, float* coefs, void* stream) {
  if (n < 1 || n > SPL_MAX_TRANSFORMS || !sums_local || !sums_global || !coefs || !state || !peer_bufs)
    return fail(SPL_E_INVALID, "spl_reduce_exchange_finalize: bad n or null pointer");
  if (world < 1 || world > spl::kMaxRanks || rank < 0 || rank >= world)
    return fail(SPL_E_INVALID, "rank %d / world %d outside [0, %d]", rank, world, spl::kMaxRanks);
  spl::ExchangeParams ep;
  std::memset(&ep, 0, sizeof(ep));
  int rc = build_reduce(ts, n, B, T, sums_local, &ep.r);
  if (rc) return rc;
  if (ep.r.n_sums > spl::kExchangeSums) return fail(SPL_E_INVALID, "%d sums exceed the exchange slot", ep.r.n_sums);
  rc = build_finalize(ts, n, sums_global, B_global, T, sc, mag, mel, coefs, &ep.f);
  if (rc) return rc;
  ep.gsums = sums_global; ep.state = state; ep.rank = rank; ep.world = world;
  ep.timeout_ns = timeout_ns; ep.error_flag = error_flag;
  for (int r = 0; r < world; ++r) {
    if (!peer_bufs[r]) return fail(SPL_E_INVALID, "null symmetric buffer of rank %d", r);
    ep.peers[r] = peer_bufs[r];
  }
  return spl_launch_reduce_exchange(ep, stream);
}

int32_t spl_backward(const spl_transform* ts, int32_t n, int32_t B, int32_t T, const float* coefs,
                     const float* g_sc, const float* g_mag, const float* g_mel, float* dx, void* stream) {
  if (n < 1 || n > SPL_MAX_TRANSFORMS || !coefs || !dx) return fail(SPL_E_INVALID, "spl_backward: bad n or null coefs/dx");
  spl::CombineParams cp;
  std::memset(&cp, 0, sizeof(cp));
  cp.n = n;
  for (int r = 0; r < n; ++r) {
    int rc = check_transform(ts + r, B, T);
    if (rc) return rc;
    if (!ts[r].gframes) return fail(SPL_E_INVALID, "transform %d: forward ran without gradient workspace", r);
    spl::CombineEntry& e = cp.e[r];
    LaunchPlan lp;
    plan_static(ts + r, B, T, true, &lp);
    // runs look to the gather like long frames: hop = run_frames * hop, win = run_len, one "frame" per run
    e.frames = ts[r].gframes; e.kind = ts[r].kind; e.half = ts[r].n_fft / 2;
    e.hop = ts[r].hop * lp.run_frames; e.win = lp.run_len;
    e.left = (ts[r].n_fft - ts[r].win) / 2; e.n_frames = lp.runs_per_utt;
    e.planar = (ts[r].kind == SPL_KIND_STFT && use_eo(ts + r)) ? 1 : 0;
  }
  cp.coefs = coefs; cp.g_sc = g_sc; cp.g_mag = g_mag; cp.g_mel = g_mel; cp.dx = dx; cp.B = B; cp.T = T;
  return spl_launch_combine(cp, stream);
}

// ---- one call per direction (the unchanged-trainer path is host-bound at small batches: trainerGAN.py:214-241 launches
// eagerly) -- the transforms are a per-recipe TEMPLATE (tables filled in, partials / gframes NULL); the per-call workspace
// is one buffer carved up by byte offsets the caller computed once from spl_geometry_of().
static int place_transforms(const spl_transform* ts, int n, void* ws, const int64_t* off_partials, const int64_t* off_gframes,
                            spl_transform* out) {
  if (n < 1 || n > SPL_MAX_TRANSFORMS) return fail(SPL_E_INVALID, "n=%d transforms not in [1,%d]", n, SPL_MAX_TRANSFORMS);
  if (!ts || !ws) return fail(SPL_E_INVALID, "null transforms / workspace");
  char* base = static_cast<char*>(ws);
  for (int r = 0; r < n; ++r) {
    out[r] = ts[r];
    out[r].partials = off_partials ? reinterpret_cast<double*>(base + off_partials[r]) : nullptr;
    out[r].gframes = off_gframes ? static_cast<void*>(base + off_gframes[r]) : nullptr;
  }
  return SPL_OK;
}

int32_t spl_loss_forward(const spl_transform* ts, int32_t n, const float* x, const float* y, int32_t B, int32_t T,
                         void* ws, const int64_t* off_partials, const int64_t* off_gframes, int64_t off_sums,
                         int64_t off_coefs, float* sc, float* mag, float* mel, uint32_t* counter, void* stream) {
  if (!off_partials) return fail(SPL_E_INVALID, "spl_loss_forward: null partial-sum offsets");
  spl_transform local[SPL_MAX_TRANSFORMS];
  int rc = place_transforms(ts, n, ws, off_partials, off_gframes, local);
  if (rc) return rc;
  rc = spl_forward(local, n, x, y, B, T, stream);
  if (rc) return rc;
  char* base = static_cast<char*>(ws);
  return spl_reduce_finalize(local, n, B, T, reinterpret_cast<double*>(base + off_sums), sc, mag, mel,
                             reinterpret_cast<float*>(base + off_coefs), counter, stream);
}

int32_t spl_loss_backward(const spl_transform* ts, int32_t n, int32_t B, int32_t T, void* ws, const int64_t* off_gframes,
                          int64_t off_coefs, const float* g_sc, const float* g_mag, const float* g_mel, float* dx, void* stream) {
  if (!off_gframes) return fail(SPL_E_INVALID, "spl_loss_backward: forward ran without gradient workspace");
  spl_transform local[SPL_MAX_TRANSFORMS];
  int rc = place_transforms(ts, n, ws, nullptr, off_gframes, local);
  if (rc) return rc;
  return spl_backward(local, n, B, T, reinterpret_cast<const float*>(static_cast<char*>(ws) + off_coefs), g_sc, g_mag, g_mel, dx, stream);
}

// ---- waveform shape loss (losses/waveform_loss.py) ---------------------------------------------------------------
static int shape_params(int rows, int T, const int32_t* winlens, int n, spl::ShapeParams* p, long long* records) {
  if (rows < 1 || T < 1) return fail(SPL_E_INVALID, "shape loss: rows %d / T %d must be positive", rows, T);
  if (T >= (1 << 29)) return fail(SPL_E_INVALID, "shape loss: T %d must be below 2^29", T);
  if (!winlens || n < 1 || n > spl::kShapeMaxWin) return fail(SPL_E_INVALID, "shape loss: %d window lengths not in [1,%d]", n, spl::kShapeMaxWin);
  std::memset(p, 0, sizeof(*p));
  p->rows = rows; p->T = T; p->n = n;
  long long ofs = 0;
  for (int r = 0; r < n; ++r) {
    if (winlens[r] < 1 || winlens[r] > T)
      return fail(SPL_E_INVALID, "shape loss: window length %d must be in [1, T=%d] (MaxPool1d raises likewise)", winlens[r], T);
    p->win[r] = winlens[r];
    p->rec_ofs[r] = ofs;
    ofs += (long long)rows * (T / winlens[r]);
  }
  *records = ofs;
  return SPL_OK;
}

static long long shape_items(int rows, int T, int span = spl::kShapeSpan) { return (long long)rows * ((T + span - 1) / span); }

static long long gcd_ll(long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; }

// Forward launch plan.  One-pass path: block = gcd of the window lengths when it is in [32, 256] and the least common multiple
// fits a span; the span is then a multiple of the lcm (windows never straddle spans) holding <= kShapeMaxBlocks
// blocks, shrunk (in lcm steps) until the device has `want` warp work items.  Otherwise: pass per window length, span
// halved from kShapeSpan down to 128 samples.
static void shape_forward_plan(int rows, int T, const int32_t* winlens, int n, long long want, int* span, int* block) {
  long long g = 0, l = 1;
  for (int r = 0; r < n; ++r) {
    g = gcd_ll(g, winlens[r]);
    l = l / gcd_ll(l, winlens[r]) * winlens[r];
    if (l > spl::kShapeSpan) l = spl::kShapeSpan + 1;
  }
  if (g >= 32 && g <= 256 && l <= spl::kShapeSpan && l / g <= spl::kShapeMaxBlocks) {
    long long k = spl::kShapeSpan / l;
    if (k * (l / g) > spl::kShapeMaxBlocks) k = spl::kShapeMaxBlocks / (l / g);
    while (k > 1 && shape_items(rows, T, (int)(k * l)) < want) --k;
    *span = (int)(k * l);
    *block = (int)g;
    return;
  }
  int sp = spl::kShapeSpan;
  while (sp > 128 && shape_items(rows, T, sp) < want) sp >>= 1;
  *span = sp;
  *block = 0;
}

// grid, warps per CTA, span and block of the forward launch (geometry and forward must agree: one partial row per warp)
static int shape_forward_dims(int rows, int T, const int32_t* winlens, int n, int* grid, int* wpc, int* span, int* block) {
  int g0 = 0, w0 = 0;
  int rc = spl_shape_dims(1LL << 40, &g0, &w0);          // the device's full complement of warps
  if (rc) return rc;
  shape_forward_plan(rows, T, winlens, n, (long long)g0 * w0, span, block);
  return spl_shape_dims(shape_items(rows, T, *span), grid, wpc);
}

int32_t spl_shape_geometry(int32_t rows, int32_t T, const int32_t* winlens, int32_t n, int64_t* record_count,
                           int64_t* partial_count) {
  spl::ShapeParams p;
  long long recs = 0;
  int rc = shape_params(rows, T, winlens, n, &p, &recs);
  if (rc) return rc;
  int grid = 0, wpc = 0, span = 0, block = 0;
  rc = shape_forward_dims(rows, T, winlens, n, &grid, &wpc, &span, &block);
  if (rc) return rc;
  if (record_count) *record_count = recs;
  if (partial_count) *partial_count = (int64_t)grid * wpc * n;
  return SPL_OK;
}

int32_t spl_shape_forward(const float* x, const float* y, int32_t rows, int32_t T, const int32_t* winlens, int32_t n,
                          int32_t* records, double* partials, double* sums, void* stream) {
  if (!x || !y || !records || !partials || !sums) return fail(SPL_E_INVALID, "spl_shape_forward: null pointer");
  spl::ShapeParams p;
  long long recs = 0;
  int rc = shape_params(rows, T, winlens, n, &p, &recs);
  if (rc) return rc;
  int grid = 0, wpc = 0, span = 0, block = 0;
  rc = shape_forward_dims(rows, T, winlens, n, &grid, &wpc, &span, &block);
  if (rc) return rc;
  p.x = x; p.y = y; p.records = records; p.partials = partials; p.span = span; p.block = block;
  rc = spl_launch_shape_forward(p, grid, wpc, stream);
  if (rc) return rc;
  spl::ReduceParams rp;
  std::memset(&rp, 0, sizeof(rp));
  rp.n_sums = n;
  for (int r = 0; r < n; ++r) { rp.base[r] = partials + r; rp.stride[r] = n; rp.count[r] = grid * wpc; }
  rp.out = sums;
  return spl_launch_reduce(rp, stream);
}

int32_t spl_shape_finalize(const double* sums, int64_t rows_global, int32_t T, const int32_t* winlens, int32_t n,
                           float* loss, void* stream) {
  if (!sums || !loss) return fail(SPL_E_INVALID, "spl_shape_finalize: null pointer");
  if (rows_global < 1) return fail(SPL_E_INVALID, "rows_global %lld < 1", (long long)rows_global);
  spl::ShapeParams p;
  long long recs = 0;
  int rc = shape_params(1, T, winlens, n, &p, &recs);
  if (rc) return rc;
  spl::ShapeFinalizeParams fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.n = n;
  for (int r = 0; r < n; ++r) fp.count[r] = (double)rows_global * (T / winlens[r]);
  fp.sums = sums; fp.loss = loss;
  return spl_launch_shape_finalize(fp, stream);
}

int32_t spl_shape_backward(const int32_t* records, int32_t rows, int64_t rows_global, int32_t T, const int32_t* winlens,
                           int32_t n, const float* g, float* dx, void* stream) {
  if (!records || !g || !dx) return fail(SPL_E_INVALID, "spl_shape_backward: null pointer");
  if (rows_global < 1) return fail(SPL_E_INVALID, "rows_global %lld < 1", (long long)rows_global);
  spl::ShapeParams p;
  long long recs = 0;
  int rc = shape_params(rows, T, winlens, n, &p, &recs);
  if (rc) return rc;
  int grid = 0, wpc = 0;
  rc = spl_shape_dims(shape_items(rows, T), &grid, &wpc);
  if (rc) return rc;
  p.records = const_cast<int32_t*>(records); p.g = g; p.dx = dx;
  for (int r = 0; r < n; ++r) p.coef[r] = (float)(1.0 / ((double)n * (double)rows_global * (double)(T / winlens[r])));
  return spl_launch_shape_backward(p, grid, wpc, stream);
}

// ---- losses on explicit magnitude tensors (stft_loss.py:38-77) ------------------------------------------------------
static int mag_dims(long long n, int* grid, int* wpc) { return spl_shape_dims((n + 4095) / 4096, grid, wpc); }

int32_t spl_mag_loss_geometry(int64_t n, int64_t* partial_count) {
  if (n < 1 || !partial_count) return fail(SPL_E_INVALID, "spl_mag_loss_geometry: n %lld < 1 or null output", (long long)n);
  int grid = 0, wpc = 0;
  int rc = mag_dims(n, &grid, &wpc);
  if (rc) return rc;
  *partial_count = (int64_t)grid * wpc * 3;
  return SPL_OK;
}

int32_t spl_mag_loss_forward(const float* x_mag, const float* y_mag, int64_t n, double* partials, double* sums,
                             float* sc, float* mag, void* stream) {
  if (!x_mag || !y_mag || !partials || !sums) return fail(SPL_E_INVALID, "spl_mag_loss_forward: null pointer");
  if (n < 1) return fail(SPL_E_INVALID, "spl_mag_loss_forward: n %lld < 1", (long long)n);
  int grid = 0, wpc = 0;
  int rc = mag_dims(n, &grid, &wpc);
  if (rc) return rc;
  spl::MagLossParams p;
  std::memset(&p, 0, sizeof(p));
  p.x = x_mag; p.y = y_mag; p.n = n; p.partials = partials;
  p.vec = (((uintptr_t)x_mag | (uintptr_t)y_mag) & 15) == 0;
  rc = spl_launch_mag_sums(p, grid, wpc, stream);
  if (rc) return rc;
  spl::ReduceParams rp;
  std::memset(&rp, 0, sizeof(rp));
  rp.n_sums = 3;
  for (int j = 0; j < 3; ++j) { rp.base[j] = partials + j; rp.stride[j] = 3; rp.count[j] = grid * wpc; }
  rp.out = sums;
  rc = spl_launch_reduce(rp, stream);
  if (rc) return rc;
  spl::MagFinalizeParams fp;
  fp.sums = sums; fp.n = (double)n; fp.sc = sc; fp.mag = mag;
  return spl_launch_mag_finalize(fp, stream);
}

int32_t spl_mag_loss_backward(const float* x_mag, const float* y_mag, int64_t n, const double* sums,
                              const float* g_sc, const float* g_mag, float* gx, float* gy, void* stream) {
  if (!x_mag || !y_mag || !sums || (!gx && !gy)) return fail(SPL_E_INVALID, "spl_mag_loss_backward: null pointer");
  if (n < 1) return fail(SPL_E_INVALID, "spl_mag_loss_backward: n %lld < 1", (long long)n);
  spl::MagLossParams p;
  std::memset(&p, 0, sizeof(p));
  p.x = x_mag; p.y = y_mag; p.n = n; p.sums = sums; p.g_sc = g_sc; p.g_mag = g_mag; p.gx = gx; p.gy = gy;
  p.vec = (((uintptr_t)x_mag | (uintptr_t)y_mag | (uintptr_t)gx | (uintptr_t)gy) & 15) == 0;
  return spl_launch_mag_backward(p, stream);
}

// ---- power-mel L1 metric, n_fft = 400 (mel_spectrogram.py:36-44) -----------------------------------------------------
static int melpow_check(int rows, int T, int n_fft, int hop) {
  if (n_fft != spl::kMp) return fail(SPL_E_INVALID, "power-mel metric: n_fft %d is not %d (torchaudio MelSpectrogram default)", n_fft, spl::kMp);
  if (rows < 1) return fail(SPL_E_INVALID, "rows %d < 1", rows);
  if (hop < 1) return fail(SPL_E_INVALID, "hop %d < 1", hop);
  if (T <= n_fft / 2) return fail(SPL_E_INVALID, "reflect padding needs T > n_fft/2 (T=%d, n_fft=%d)", T, n_fft);
  if ((long long)rows * (1 + T / hop) > 0x7fffffffLL) return fail(SPL_E_INVALID, "rows * frames exceeds 2^31");
  return SPL_OK;
}

int32_t spl_melpow_geometry(int32_t rows, int32_t T, int32_t n_fft, int32_t hop, int64_t* partial_count) {
  if (!partial_count) return fail(SPL_E_INVALID, "spl_melpow_geometry: null output");
  int rc = melpow_check(rows, T, n_fft, hop);
  if (rc) return rc;
  int grid = 0, wpc = 0;
  rc = spl_shape_dims(((long long)rows * (1 + T / hop) + 1) / 2, &grid, &wpc);      // a warp takes two frames at a time
  if (rc) return rc;
  *partial_count = (int64_t)grid * wpc;
  return SPL_OK;
}

int32_t spl_melpow_l1(const float* x, const float* y, int32_t rows, int32_t T, int32_t n_fft, int32_t hop,
                      const float* window, const float* twiddle, int32_t n_mels, int32_t nnz,
                      const int32_t* mel_ptr, const int32_t* mel_ent, double* partials, double* sum, float* loss,
                      float* mel_x, float* mel_y, void* stream) {
  int rc = melpow_check(rows, T, n_fft, hop);
  if (rc) return rc;
  if (!x || !y || !window || !twiddle || !mel_ptr || !mel_ent || !partials || !sum || !loss)
    return fail(SPL_E_INVALID, "spl_melpow_l1: null pointer");
  if (n_mels < 1 || n_mels > spl::kMpMaxMels) return fail(SPL_E_INVALID, "n_mels %d must be in [1, %d]", n_mels, spl::kMpMaxMels);
  if (nnz < 1 || nnz > spl::kMpMaxNnz) return fail(SPL_E_INVALID, "filterbank non-zeros %d must be in [1, %d]", nnz, spl::kMpMaxNnz);
  const int n_frames = 1 + T / hop;
  int grid = 0, wpc = 0;
  rc = spl_shape_dims(((long long)rows * n_frames + 1) / 2, &grid, &wpc);
  if (rc) return rc;
  spl::MelPowParams p;
  std::memset(&p, 0, sizeof(p));
  p.x = x; p.y = y; p.rows = rows; p.T = T; p.hop = hop; p.n_frames = n_frames;
  p.window = window; p.twiddle = reinterpret_cast<const float2*>(twiddle);
  p.n_mels = n_mels; p.nnz = nnz; p.mel_ptr = mel_ptr; p.mel_ent = reinterpret_cast<const int2*>(mel_ent);
  p.partials = partials; p.mel_x = mel_x; p.mel_y = mel_y;
  const spl::MelPowSmem sm = spl::melpow_smem(n_mels, nnz);
  rc = spl_launch_melpow(p, grid, wpc, ((size_t)sm.total + (size_t)sm.per_warp * wpc) * 4, stream);
  if (rc) return rc;
  spl::ReduceParams rp;
  std::memset(&rp, 0, sizeof(rp));
  rp.n_sums = 1; rp.base[0] = partials; rp.stride[0] = 1; rp.count[0] = grid * wpc; rp.out = sum;
  rc = spl_launch_reduce(rp, stream);
  if (rc) return rc;
  spl::ShapeFinalizeParams fp;                       // loss = sum / count: the one-term case of the shape-loss finalize
  std::memset(&fp, 0, sizeof(fp));
  fp.n = 1; fp.count[0] = (double)rows * n_mels * n_frames; fp.sums = sum; fp.loss = loss;
  return spl_launch_shape_finalize(fp, stream);
}

}  // extern "C"
