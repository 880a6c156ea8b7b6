// Host-side logic of the C ABI (argument checks, geometry, kernel-parameter assembly), shared by
// the CUDA backend (specloss.cu) and by the CPU SIMT emulator used in the test-suite
// (tests/emu/specloss_emu.cpp).  The includer provides, before including this file:
//   int fail(int code, const char* fmt, ...);
//   template <int NFFT, int KIND, bool GRAD, int WIN_T> int spl_launch_transform(const spl::TransformParams&, int n_mels, void* stream);
//   int spl_launch_reduce(const spl::ReduceParams&, void* stream);
//   int spl_launch_finalize(const spl::FinalizeParams&, void* stream);
//   int spl_launch_reduce_finalize(const spl::ReduceFinalizeParams&, void* stream);
//   int spl_launch_combine(const spl::CombineParams&, void* stream);
namespace {

bool supported_nfft(int n) { return n == 512 || n == 1024 || n == 2048; }

int frames_in_flight(int n_fft) { return n_fft == 512 ? 2 : 1; }

// overlap-add ring per warp; none when every chunk is a single frame (the frame is its own slot)
int ring_entries(const spl_transform* t) {
  const int fpw = frames_in_flight(t->n_fft);
  if (fpw == 1 && t->frames_per_chunk == 1) return 0;
  return t->win + (fpw - 1) * t->hop;
}

int check_transform(const spl_transform* t, int B, int T) {
  if (!t) return fail(SPL_E_INVALID, "null transform");
  if (t->kind != SPL_KIND_STFT && t->kind != SPL_KIND_MEL) return fail(SPL_E_INVALID, "kind %d unknown", t->kind);
  if (!supported_nfft(t->n_fft)) return fail(SPL_E_INVALID, "n_fft %d not in {512,1024,2048}", t->n_fft);
  if (t->win < 1 || t->win > t->n_fft) return fail(SPL_E_INVALID, "win %d must be in [1, n_fft=%d]", t->win, t->n_fft);
  if (t->hop < 1 || t->hop > t->win) return fail(SPL_E_INVALID, "hop %d must be in [1, win=%d]", t->hop, t->win);
  if (t->frames_per_chunk < 1) return fail(SPL_E_INVALID, "frames_per_chunk %d < 1", t->frames_per_chunk);
  if (B < 1) return fail(SPL_E_INVALID, "batch %d < 1", B);
  if (T <= t->n_fft / 2) return fail(SPL_E_INVALID, "reflect padding needs T > n_fft/2 (T=%d, n_fft=%d)", T, t->n_fft);
  if (t->kind == SPL_KIND_MEL && (t->n_mels < 2 || t->n_mels > 512))
    return fail(SPL_E_INVALID, "n_mels %d must be in [2, 512]", t->n_mels);
  return SPL_OK;
}

void geometry(const spl_transform* t, int B, int T, spl_geometry* g) {
  g->n_frames = 1 + T / t->hop;
  g->n_bins = t->n_fft / 2 + 1;
  g->n_chunks = (g->n_frames + t->frames_per_chunk - 1) / t->frames_per_chunk;
  g->span = (t->frames_per_chunk - 1) * t->hop + t->win;
  g->n_sums = t->kind == SPL_KIND_STFT ? 3 : 1;
  g->partial_count = (int64_t)B * g->n_chunks * g->n_sums;
  g->gchunk_bytes = (int64_t)B * g->n_chunks * g->span * (t->kind == SPL_KIND_STFT ? 8 : 4);
  const int ring_n = ring_entries(t);
  int words = 0;
#define SPL_WORDS(N)                                                                                        \
  words = t->kind == SPL_KIND_STFT ? spl::SmemLayout<N, spl::kKindStft, true>::words_per_warp(ring_n, 0)    \
                                   : spl::SmemLayout<N, spl::kKindMel, true>::words_per_warp(ring_n, t->n_mels)
  if (t->n_fft == 512) { SPL_WORDS(512); } else if (t->n_fft == 1024) { SPL_WORDS(1024); } else { SPL_WORDS(2048); }
#undef SPL_WORDS
  const int lanes = t->n_fft == 512 ? 16 : 32;
  const spl::CtaTables ct = spl::cta_tables(t->n_fft, t->win, t->kind, lanes, t->mel_rounds, t->mel_entry_rows);
  g->smem_table_bytes = (int64_t)ct.total * 4;
  g->smem_warp_bytes = (int64_t)words * 4;
}

template <int NFFT, int WIN_T>
int launch_win(const spl::TransformParams& p, int kind, bool grad, int n_mels, void* s) {
  if (kind == SPL_KIND_STFT)
    return grad ? spl_launch_transform<NFFT, spl::kKindStft, true, WIN_T>(p, n_mels, s)
                : spl_launch_transform<NFFT, spl::kKindStft, false, WIN_T>(p, n_mels, s);
  return grad ? spl_launch_transform<NFFT, spl::kKindMel, true, WIN_T>(p, n_mels, s)
              : spl_launch_transform<NFFT, spl::kKindMel, false, WIN_T>(p, n_mels, s);
}

// Window lengths of the shipped configurations get kernels with the window support known at compile
// time (zero taps pruned); every other (n_fft, win) pair runs the generic kernel of that n_fft.
int launch_any(const spl::TransformParams& p, int n_fft, int kind, bool grad, int n_mels, void* s) {
  if (n_fft == 1024) return p.win == 600 ? launch_win<1024, 600>(p, kind, grad, n_mels, s) : launch_win<1024, 0>(p, kind, grad, n_mels, s);
  if (n_fft == 512) return p.win == 240 ? launch_win<512, 240>(p, kind, grad, n_mels, s) : launch_win<512, 0>(p, kind, grad, n_mels, s);
  if (p.win == 1200) return launch_win<2048, 1200>(p, kind, grad, n_mels, s);
  if (p.win == 2048) return launch_win<2048, 2048>(p, kind, grad, n_mels, s);
  return launch_win<2048, 0>(p, kind, grad, n_mels, s);
}

}  // namespace

extern "C" {

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
    if (frames_in_flight(t->n_fft) == 2 && (t->frames_per_chunk & 1))
      return fail(SPL_E_INVALID, "transform %d: frames_per_chunk must be even for n_fft=512", r);
    if (t->kind == SPL_KIND_MEL && (!t->mel_tasks || !t->mel_entries || t->mel_rounds < 1 || t->mel_entry_rows < 1 || !t->bin_tab))
      return fail(SPL_E_INVALID, "transform %d: null mel table", r);
    spl_geometry g;
    geometry(t, B, T, &g);
    spl::TransformParams p;
    std::memset(&p, 0, sizeof(p));
    p.x = x; p.y = y; p.B = B; p.T = T;
    p.hop = t->hop; p.win = t->win; p.left = (t->n_fft - t->win) / 2;
    p.n_frames = g.n_frames; p.m = t->frames_per_chunk; p.n_chunks = g.n_chunks; p.span = g.span;
    p.ring_n = ring_entries(t);
    p.eps = t->eps; p.window = t->window; p.twiddle = reinterpret_cast<const float2*>(t->twiddle);
    p.partials = t->partials; p.gchunks = t->gchunks;
    p.n_mels = t->kind == SPL_KIND_MEL ? t->n_mels : 0;
    p.inv_ln_base = t->inv_ln_base;
    p.mel_tasks = t->mel_tasks; p.mel_entries = t->mel_entries;
    p.mel_rounds = t->kind == SPL_KIND_MEL ? t->mel_rounds : 0;
    p.mel_entry_rows = t->kind == SPL_KIND_MEL ? t->mel_entry_rows : 0;
    p.bin_tab = t->bin_tab;
    const bool grad = t->gchunks != nullptr;
    void* s = stream;
    rc = launch_any(p, t->n_fft, t->kind, grad, p.n_mels, s);
    if (rc) return rc;
  }
  return SPL_OK;
}

static int build_reduce(const spl_transform* ts, int n, int B, int T, double* sums, spl::ReduceParams* rp) {
  std::memset(rp, 0, sizeof(*rp));
  int k = 0;
  for (int r = 0; r < n; ++r) {
    int rc = check_transform(ts + r, B, T);
    if (rc) return rc;
    if (!ts[r].partials) return fail(SPL_E_INVALID, "transform %d: null partials", r);
    spl_geometry g;
    geometry(ts + r, B, T, &g);
    for (int j = 0; j < g.n_sums; ++j) {
      if (k >= 16) return fail(SPL_E_INVALID, "too many sums");
      rp->base[k] = ts[r].partials + j;
      rp->stride[k] = g.n_sums;
      rp->count[k] = (int)((int64_t)B * g.n_chunks);
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

int32_t spl_backward(const spl_transform* ts, int32_t n, int32_t B, int32_t T, const float* coefs,
                     const float* g_sc, const float* g_mag, const float* g_mel, float* dx, void* stream) {
  if (n < 1 || n > SPL_MAX_TRANSFORMS || !coefs || !dx) return fail(SPL_E_INVALID, "spl_backward: bad n or null coefs/dx");
  spl::CombineParams cp;
  std::memset(&cp, 0, sizeof(cp));
  cp.n = n;
  for (int r = 0; r < n; ++r) {
    int rc = check_transform(ts + r, B, T);
    if (rc) return rc;
    if (!ts[r].gchunks) return fail(SPL_E_INVALID, "transform %d: forward ran without gradient workspace", r);
    spl_geometry g;
    geometry(ts + r, B, T, &g);
    spl::CombineEntry& e = cp.e[r];
    e.chunks = ts[r].gchunks; e.kind = ts[r].kind; e.half = ts[r].n_fft / 2; e.hop = ts[r].hop; e.win = ts[r].win;
    e.left = (ts[r].n_fft - ts[r].win) / 2; e.m = ts[r].frames_per_chunk; e.n_chunks = g.n_chunks;
    e.span = g.span; e.n_frames = g.n_frames;
  }
  cp.coefs = coefs; cp.g_sc = g_sc; cp.g_mag = g_mag; cp.g_mel = g_mel; cp.dx = dx; cp.B = B; cp.T = T;
  return spl_launch_combine(cp, stream);
}

}  // extern "C"
