// Device code of the B200 spectral-loss path (sm_100a): fused framing + windowing + FFT +
// loss partial sums + adjoint FFT + overlap-add, one warp per STFT frame.
//
// Replaces, on the GPU, what the reference computes with torch.stft + ~95 ATen ops per
// resolution (reference: losses/stft_loss.py:19-117, losses/mel_loss.py:74-94,151-154).
// Math spec: SURVEY.md appendix A; design: DESIGN.md.
//
// The same source is compiled by g++ against tests/emu/cuda_emu.h (SPECLOSS_EMU) so that the
// index maps, the epilogue and the overlap-add logic can be executed lane-accurately on a CPU
// in the `-m "not gpu"` test-suite.  That build is a test artefact, never a product fallback.
#pragma once

#ifdef SPECLOSS_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <string.h>

#define SPL_DEVICE __device__ __forceinline__
#include "fft_codelets.cuh"

// MUFU.RSQ / MUFU.LG2 without the denormal fix-up code: their arguments are clamped at >= 1e-10 first.
#ifdef SPECLOSS_EMU
static inline float spl_fast_rsqrt(float x) { return 1.0f / std::sqrt(x); }
static inline float spl_fast_log2(float x) { return std::log2(x); }
#else
__device__ __forceinline__ float spl_fast_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float spl_fast_log2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#endif

#ifndef SPL_MAX_WARPS_SMALL
#define SPL_MAX_WARPS_SMALL 16
#endif

namespace spl {

constexpr int kKindStft = 0;
constexpr int kKindMel = 1;

// ---------------------------------------------------------------------------------------------
// FFT geometry: N = L * R.  A group of L lanes owns one frame; every lane holds R complex points.
//   pass A: in-lane R-point DFT over n2   (n = n1 + L*n2, n1 = lane in group)
//   twiddle W_N^(n1*k2), transpose through shared memory
//   pass B: in-lane L-point DFTs over n1  (R/L rows k2 per lane), output bin k = k2 + R*k1
// ---------------------------------------------------------------------------------------------
template <int NFFT> struct FftGeom;
template <> struct FftGeom<512>  { static constexpr int L = 16, R = 32; };
template <> struct FftGeom<1024> { static constexpr int L = 32, R = 32; };
template <> struct FftGeom<2048> { static constexpr int L = 32, R = 64; };

template <int P> struct Dft;
template <> struct Dft<16> { static SPL_DEVICE void run(float2 (&v)[16]) { fft16(v); } };
template <> struct Dft<32> { static SPL_DEVICE void run(float2 (&v)[32]) { fft32(v); } };
template <> struct Dft<64> { static SPL_DEVICE void run(float2 (&v)[64]) { fft64(v); } };

// One transform (= one STFT resolution or one mel resolution) over a batch of utterances.
struct TransformParams {
  const float* x;        // prediction  (B, T)
  const float* y;        // target      (B, T)
  int B, T;
  int hop, win, left;    // left = (N - win) / 2 : first non-zero tap of the centred window
  int n_frames;          // 1 + T / hop
  int m;                 // frames per chunk (one warp walks one chunk)
  int n_chunks;          // chunks per utterance
  int span;              // (m - 1) * hop + win : gradient slot length per chunk
  int ring_n;            // ring buffer entries per warp: win + (32/L - 1) * hop
  float eps;
  const float* window;   // win taps
  const float2* twiddle; // [R][L] : W_N^(n1*k2) at [k2 * L + n1]
  double* partials;      // [B * n_chunks][n_sums]
  void* gchunks;         // [B * n_chunks][span] float2 (stft: u=sc part, v=log-mag part) | float (mel)
  // mel only
  int n_mels;
  float inv_ln_base;     // 1 / ln(log_base)  (1 for natural log)
  const void* mel_tasks;      // int4[mel_rounds][L]: {row | group << 12 | iters << 20, first entry row, -, -}
  const void* mel_entries;    // int2[mel_entry_rows][L]: {slot offset of the bin's amplitudes, weight bits}
  int mel_rounds;
  int mel_entry_rows;         // sum of iters over the rounds
  const void* bin_tab;        // int4[K]: {m0, bits(W[k,m0]), bits(W[k,m0+1]), 0}: bin k feeds rows m0, m0+1 only
};

// a * w on the packed fp32 pipe: w.x * (a.x, a.y) + w.y * (-a.y, a.x)  (FMUL2 + FFMA2; the broadcasts
// and the swap/negate of `a` are operand modifiers)
SPL_DEVICE float2 cmul(float2 a, float2 w) {
  return __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}

SPL_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Constant tables of a transform, staged once per CTA in shared memory (one persistent CTA per SM):
// at 200+ KB of shared memory per SM the L1 keeps only ~28 KB, and twiddles + window + mel tables
// (30-64 KB) would otherwise be re-fetched from L2 at every frame (r1c profile: long-scoreboard 3.7/issue).
struct CtaTables {
  int tw, win, tasks, entries, bintab, total;     // word offsets
};
SPL_DEVICE int align4(int n) { return (n + 3) & ~3; }
static __host__ __device__ inline CtaTables cta_tables(int n_fft, int win, int kind, int lanes, int mel_rounds,
                                                       int mel_entry_rows) {
  CtaTables t;
  int o = 0;
  t.tw = o;      o += 2 * n_fft;
  t.win = o;     o += (win + 3) & ~3;
  t.tasks = o;   o += kind == kKindMel ? 4 * mel_rounds * lanes : 0;
  t.entries = o; o += kind == kKindMel ? ((2 * mel_entry_rows * lanes + 3) & ~3) : 0;
  t.bintab = o;  o += kind == kKindMel ? 4 * (n_fft / 2 + 1) : 0;
  t.total = o;
  return t;
}

// shared memory carve-up per warp (in 4-byte words); the host side uses the same function
template <int NFFT, int KIND, bool GRAD>
struct SmemLayout {
  using G = FftGeom<NFFT>;
  static constexpr int FPW = 32 / G::L;                      // frames in flight per warp
  static constexpr int BUF_F2 = G::R * (G::L + 1);           // float2 per frame slot: R rows of L (+1 pad)
  static __host__ __device__ int words_per_warp(int ring_n, int n_mels) {
    int w = FPW * BUF_F2 * 2;
    if (GRAD) w += (KIND == kKindStft ? 2 : 1) * ((ring_n + 3) & ~3);
    if (KIND == kKindMel) w += FPW * 2 * ((n_mels + 3) & ~3);
    return (w + 3) & ~3;
  }
};

// Spectrum / time-sample layout inside a frame slot: element k = k2 + R*k1 lives in row k2, column k1.
// Pass B works on whole rows in place (one owner lane per row, no hazards); lanes that walk
// k = l + L*i (or N - k) touch a different bank each: the row pitch L+1 is odd in float2 units.
template <int NFFT>
SPL_DEVICE int pos(int k) {
  using G = FftGeom<NFFT>;
  return (k & (G::R - 1)) * (G::L + 1) + (k / G::R);
}

// ---------------------------------------------------------------------------------------------
// The one FFT core of a kernel.  In: this lane's R points of the sequence, element n = l + L*n2 in
// (re[n2], im[n2]).  Out: the forward DFT in the frame slot, element k at buf[pos(k)].
// The inverse (un-normalised, e^{+i..}) is the same code on swapped components: feed (im, re),
// read back (.y, .x).
// ---------------------------------------------------------------------------------------------
template <int NFFT>
SPL_DEVICE void fft_core(float2 (&v)[FftGeom<NFFT>::R], float2* buf, const float2* tw, int l) {
  using G = FftGeom<NFFT>;
  constexpr int L = G::L, R = G::R, RPL = R / L;
  // [region: fft_core pass A (in-lane R-point DFT)]
  Dft<R>::run(v);
  // [region: fft_core twiddle + transposed store]
#pragma unroll
  for (int k2 = 0; k2 < R; ++k2) buf[k2 * (L + 1) + l] = k2 > 0 ? cmul(v[k2], tw[k2 * L + l]) : v[0];
  __syncwarp();
  // [region: fft_core pass B (row load, L-point DFT, row store)]
#pragma unroll 1
  for (int j = 0; j < RPL; ++j) {
    float2* row = buf + (j * L + l) * (L + 1);
    float2 b[L];
#pragma unroll
    for (int n1 = 0; n1 < L; ++n1) b[n1] = row[n1];
    Dft<L>::run(b);
#pragma unroll
    for (int k1 = 0; k1 < L; ++k1) row[k1] = b[k1];
  }
  __syncwarp();
}

// [region: misc helpers]
// reflect index without edge repeat (torch.stft center=True, pad_mode="reflect")
SPL_DEVICE int reflect(int s, int T) {
  s = s < 0 ? -s : s;
  return s >= T ? 2 * (T - 1) - s : s;
}

SPL_DEVICE float bits_to_float(int b) {
#ifdef SPECLOSS_EMU
  float f; memcpy(&f, &b, 4); return f;
#else
  return __int_as_float(b);
#endif
}

// ---------------------------------------------------------------------------------------------
// The transform kernel body.  One warp walks chunks of `m` consecutive frames of one utterance
// (grid-stride over chunks).  WIN_T > 0: window length known at compile time (shipped configs),
// which prunes the zero taps out of the load, the first butterflies and the overlap-add.
// ---------------------------------------------------------------------------------------------
// One mirror pair (k, N-k) of the packed spectrum Z = FFT(x + i y) of an STFT-loss frame:
//   2X[k] = Z[k] + conj Z[N-k],  2Y[k] = -i (Z[k] - conj Z[N-k]);
// loss terms from the doubled spectra (powers x4, magnitudes x2; the chunk's sums are rescaled once), and --
// GRAD -- the un-scaled gradient spectra written back in place:
//   H[k] = w/2 (alpha + i beta) (2X),  H[N-k] = w/2 (alpha + i beta) conj(2X),  w = 1/2 (1 when k mirrors itself)
//   alpha = gate (Ax - Ay)/Ax  (spectral convergence),  beta = gate sign(Ax - Ay)/Ax^2  (log magnitude).
// EQ: prediction and target frames are bit-identical, Y := X, every difference term is exactly zero.
template <bool GRAD, bool EQ>
SPL_DEVICE void stft_pair(float2* qa, float2* qb, bool self, float eps4, float& s1, float& s2, float& s3) {
  const float2 a = *qa, bm = *qb;
  const float2 x2 = __fadd2_rn(a, make_float2(bm.x, -bm.y));
  const float px = fmaf(x2.x, x2.x, x2.y * x2.y);
  const float pxc = fmaxf(px, eps4);
  if (EQ) {
    s2 += pxc;
    if (GRAD) {
      *qa = make_float2(0.f, 0.f);
      if (!self) *qb = make_float2(0.f, 0.f);
    }
    return;
  }
  const float2 y2 = __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));
  const float py = fmaf(y2.x, y2.x, y2.y * y2.y);
  const float pyc = fmaxf(py, eps4);
  const float rx = spl_fast_rsqrt(pxc), ry = spl_fast_rsqrt(pyc);
  const float ax = __fmul_rn(pxc, rx), ay = __fmul_rn(pyc, ry);       // 2 Ax, 2 Ay
  const float d = __fsub_rn(ay, ax);                                  // never contracted: 0 when pxc == pyc
  s1 = fmaf(d, d, s1);
  s2 += pyc;
  s3 += fabsf(spl_fast_log2(pyc * rx * rx));
  if (GRAD) {
    const float wq = self ? 0.5f : 0.25f;
    const float rxg = px >= eps4 ? rx : 0.f;                          // clamp gate of the reference
    const float sgn = fminf(fmaxf((pxc - pyc) * 1e25f, -1.f), 1.f);   // exact sign, 0 when equal
    const float gr = -wq * d * rxg;
    const float gi = 4.f * wq * sgn * rx * rxg;
    const float2 g2 = make_float2(gi, gi), r2 = make_float2(gr, gr);
    *qa = __ffma2_rn(make_float2(-x2.y, x2.x), g2, __fmul2_rn(x2, r2));
    if (!self) *qb = __ffma2_rn(make_float2(x2.y, x2.x), g2, __fmul2_rn(make_float2(x2.x, -x2.y), r2));
  }
}

// all pairs of one lane: rows l + L*j, columns 0 .. L/2-1 (column 0 of row 0 mirrors itself)
template <int NFFT, bool GRAD, bool EQ>
SPL_DEVICE void stft_epilogue(float2* buf, int l, float eps4, float& s1, float& s2, float& s3) {
  using G = FftGeom<NFFT>;
  constexpr int L = G::L, R = G::R;
#pragma unroll
  for (int j = 0; j < R / L; ++j) {
    const int row = l + L * j;
    float2* pa = buf + row * (L + 1);
    float2* pb = (row == 0) ? buf + L : buf + (R - row) * (L + 1) + (L - 1);      // mirror of column c is pb[-c]
    stft_pair<GRAD, EQ>(pa, row == 0 ? pa : pb, row == 0, eps4, s1, s2, s3);
#pragma unroll (EQ ? 1 : (L == 32 ? 5 : 7))
    for (int c = 1; c < L / 2; ++c) stft_pair<GRAD, EQ>(pa + c, pb - c, false, eps4, s1, s2, s3);
  }
  if (l == 0) stft_pair<GRAD, EQ>(buf + L / 2, buf + L / 2, true, eps4, s1, s2, s3);   // bin N/2
}

// mel, pass 1 on one mirror pair: keep 2X[k] in the bin's own slot, park (Ax, Ay) in `amp_slot`.
template <bool EQ>
SPL_DEVICE void mel_pair_amp(float2* qa, float2* qb, float2* x_slot, float2* amp_slot, float eps4) {
  const float2 a = *qa, bm = *qb;
  const float2 x2 = __fadd2_rn(a, make_float2(bm.x, -bm.y));
  const float2 y2 = EQ ? x2 : __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));
  const float pxc = fmaxf(fmaf(x2.x, x2.x, x2.y * x2.y), eps4);
  const float pyc = EQ ? pxc : fmaxf(fmaf(y2.x, y2.x, y2.y * y2.y), eps4);
  const float ax = 0.5f * pxc * spl_fast_rsqrt(pxc);                       // sqrt(max(|X|^2, eps))
  const float ay = EQ ? ax : 0.5f * pyc * spl_fast_rsqrt(pyc);
  *x_slot = x2;
  *amp_slot = make_float2(ax, ay);
}

// mel, pass 3 on one mirror pair: gA[k] = gM[m0] W[k,m0] + gM[m0+1] W[k,m0+1];  H[k] = w gA gate / Ax * X
SPL_DEVICE void mel_pair_grad(float2* qa, float2* qb, const float2* x_slot, bool self, int4 bt, const float2* msum,
                              float eps4, bool active) {
  const float2 x2 = *x_slot;
  const float ga = fmaf(msum[bt.x].x, bits_to_float(bt.y), msum[bt.x + 1].x * bits_to_float(bt.z));
  const float px = fmaf(x2.x, x2.x, x2.y * x2.y);                          // 4 |X|^2
  // w gA / Ax * X = w gA * 2 rsqrt(px) * (2X) / 2
  const float g = (active && px >= eps4) ? (self ? 1.f : 0.5f) * ga * spl_fast_rsqrt(px) : 0.f;
  const float2 hk = __fmul2_rn(x2, make_float2(g, g));
  *qa = hk;
  if (!self) *qb = make_float2(hk.x, -hk.y);
}

// CTA prologue: every thread copies its share of the constant tables into shared memory.
template <int NFFT, int KIND>
SPL_DEVICE void cta_load_tables(const TransformParams& p, float* smem, int tid, int nthreads) {
  constexpr int L = FftGeom<NFFT>::L;
  const CtaTables ct = cta_tables(NFFT, p.win, KIND, L, p.mel_rounds, p.mel_entry_rows);
  const float* tw = reinterpret_cast<const float*>(p.twiddle);
  for (int i = tid; i < 2 * NFFT; i += nthreads) smem[ct.tw + i] = __ldg(&tw[i]);
  for (int i = tid; i < p.win; i += nthreads) smem[ct.win + i] = __ldg(&p.window[i]);
  if (KIND == kKindMel) {
    int* ism = reinterpret_cast<int*>(smem);
    const int* a = reinterpret_cast<const int*>(p.mel_tasks);
    const int* b = reinterpret_cast<const int*>(p.mel_entries);
    const int* c = reinterpret_cast<const int*>(p.bin_tab);
    for (int i = tid; i < 4 * p.mel_rounds * L; i += nthreads) ism[ct.tasks + i] = __ldg(&a[i]);
    for (int i = tid; i < 2 * p.mel_entry_rows * L; i += nthreads) ism[ct.entries + i] = __ldg(&b[i]);
    for (int i = tid; i < 4 * (NFFT / 2 + 1); i += nthreads) ism[ct.bintab + i] = __ldg(&c[i]);
  }
}

template <int NFFT, int KIND, bool GRAD, int WIN_T>
SPL_DEVICE void transform_body(const TransformParams& p, float* smem, int block, int tid, int grid, int wpc) {
  using G = FftGeom<NFFT>;
  using SL = SmemLayout<NFFT, KIND, GRAD>;
  constexpr int L = G::L, R = G::R, FPW = SL::FPW, HALF = NFFT / 2;
  const int warp = tid >> 5, lane = tid & 31;
  const int l = lane & (L - 1), h = lane / L;
  const int win = WIN_T > 0 ? WIN_T : p.win;
  const int left = WIN_T > 0 ? (NFFT - WIN_T) / 2 : p.left;

  const CtaTables ct = cta_tables(NFFT, p.win, KIND, L, p.mel_rounds, p.mel_entry_rows);
  const float2* tw = reinterpret_cast<const float2*>(smem + ct.tw);
  const float* wtab = smem + ct.win;
  const int4* mel_tasks = reinterpret_cast<const int4*>(smem + ct.tasks);
  const int2* mel_entries = reinterpret_cast<const int2*>(smem + ct.entries);
  const int4* bin_tab = reinterpret_cast<const int4*>(smem + ct.bintab);

  float* wsm = smem + ct.total + (size_t)warp * SL::words_per_warp(p.ring_n, p.n_mels);
  float2* buf = reinterpret_cast<float2*>(wsm) + h * SL::BUF_F2;     // this frame slot
  float* ring_base = wsm + FPW * SL::BUF_F2 * 2;
  float2* ring2 = reinterpret_cast<float2*>(ring_base);              // stft: (u, v)
  float* ring1 = ring_base;                                          // mel : u
  float2* msum = reinterpret_cast<float2*>(ring_base + (GRAD ? (KIND == kKindStft ? 2 : 1) * align4(p.ring_n) : 0)) +
                 h * align4(p.n_mels);              // mel: per-row (Mx, My), then (gM, -)
  const bool no_ring = (FPW == 1) && (p.m == 1);    // one frame per chunk: windowed frame goes straight to its slot
  const int total_chunks = p.B * p.n_chunks;

  // [region: chunk loop setup]
  for (int chunk_id = block * wpc + warp; chunk_id < total_chunks; chunk_id += grid * wpc) {
    const int b = chunk_id / p.n_chunks, c = chunk_id - b * p.n_chunks;
    const int t0 = c * p.m;
    const int m_c = min(p.m, p.n_frames - t0);
    const float* __restrict__ xb = p.x + (size_t)b * p.T;
    const float* __restrict__ yb = p.y + (size_t)b * p.T;

    if (GRAD && !no_ring) {
      if (KIND == kKindStft) for (int i = lane; i < p.ring_n; i += 32) ring2[i] = make_float2(0.f, 0.f);
      else                   for (int i = lane; i < p.ring_n; i += 32) ring1[i] = 0.f;
      __syncwarp();
    }
    float2* out2 = reinterpret_cast<float2*>(p.gchunks) + (size_t)chunk_id * p.span;
    float* out1 = reinterpret_cast<float*>(p.gchunks) + (size_t)chunk_id * p.span;
    const int span_c = (m_c - 1) * p.hop + win;
    int flushed = 0;
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;   // stft: 4*S1, 4*S2, S3/(0.5 ln 2) ; mel: s1 = S4

    for (int step = 0; step * FPW < m_c; ++step) {
      const int jc = step * FPW + h;          // frame index inside the chunk
      const bool active = jc < m_c;
      const int t = t0 + jc;
      bool frame_equal = false;
#pragma unroll 1
      for (int job = 0; job < (GRAD ? 2 : 1); ++job) {
        float2 v[R];
        if (job == 0) {
          // [region: A load taps + window + equality vote]
          // ---- A. taps of frame t: reflect-pad, window, pack z = x*w + i*y*w ---------------------
          const int s0 = t * p.hop - HALF;                         // sample index of tap n = 0
          const bool interior = (s0 + left >= 0) && (s0 + left + win <= p.T);
          bool same = true;
          if (active && interior) {
            // fast path: no reflection; every address is a compile-time offset from three pointers
            const float* __restrict__ xp = xb + s0 + l;
            const float* __restrict__ yp = yb + s0 + l;
            const float* wp = wtab + (l - left);
#pragma unroll
            for (int n2 = 0; n2 < R; ++n2) {
              const int lo = L * n2 - left;                        // tap index of lane 0
              if (WIN_T > 0 && (lo + L - 1 < 0 || lo >= WIN_T)) { v[n2] = make_float2(0.f, 0.f); continue; }
              const bool all_lanes = WIN_T > 0 && lo >= 0 && lo + L - 1 < WIN_T;
              float2 xy = make_float2(0.f, 0.f);
              if (all_lanes || (lo + l >= 0 && lo + l < win)) {
                const float w = wp[L * n2];
                xy = __fmul2_rn(make_float2(__ldg(xp + L * n2), __ldg(yp + L * n2)), make_float2(w, w));
              }
              same = same && (xy.x == xy.y);
              v[n2] = xy;
            }
          } else {
#pragma unroll
            for (int n2 = 0; n2 < R; ++n2) {
              const int lo = L * n2 - left;
              if (WIN_T > 0 && (lo + L - 1 < 0 || lo >= WIN_T)) { v[n2] = make_float2(0.f, 0.f); continue; }
              const int tap = lo + l;
              float xv = 0.f, yv = 0.f;
              if (active && tap >= 0 && tap < win) {
                const int sidx = reflect(s0 + L * n2 + l, p.T);
                const float w = wtab[tap];
                xv = __ldg(&xb[sidx]) * w;
                yv = __ldg(&yb[sidx]) * w;
              }
              same = same && (xv == yv);
              v[n2] = make_float2(xv, yv);
            }
          }
          // A frame whose prediction and target taps are bit-identical must contribute exactly zero
          // (the reference returns sc = mag = mel = 0 and a zero gradient for x == y); the packed FFT
          // would leave ~1e-7 of rounding asymmetry between X and Y, so such frames reuse X for Y.
          const unsigned eq_bits = __ballot_sync(0xffffffffu, same);
          const unsigned grp_mask = (L == 32) ? 0xffffffffu : (((1u << (L & 31)) - 1u) << (h * L));
          frame_equal = (eq_bits & grp_mask) == grp_mask;
        } else {
          // [region: D job-1 load of H]
          // ---- D. gradient spectrum H (slot layout) -> inverse DFT via swapped components ------
#pragma unroll
          for (int n2 = 0; n2 < R; ++n2) {
            const float2 hk = buf[(l + L * (n2 % (R / L))) * (L + 1) + n2 / (R / L)];   // element n = l + L*n2
            v[n2] = make_float2(hk.y, hk.x);
          }
          __syncwarp();
        }
        fft_core<NFFT>(v, buf, tw, l);
        if (job == 1) {
          // [region: E window + overlap-add]
          // ---- E. window, overlap-add into the ring (slot holds (imag, real) = (v, u) swapped) --
          const int n2_lo = left / L, n2_hi = (left + win - 1) / L;      // register slots with live taps
          if (no_ring) {
            if (active) {
#pragma unroll 4
              for (int n2 = n2_lo; n2 <= n2_hi; ++n2) {
                const int tap = L * n2 - left + l;
                if (tap >= 0 && tap < win) {
                  const float w = wtab[tap];
                  const float2 g = buf[(l + L * (n2 % (R / L))) * (L + 1) + n2 / (R / L)];
                  if (KIND == kKindStft) out2[tap] = __fmul2_rn(make_float2(g.y, g.x), make_float2(w, w));
                  else                   out1[tap] = g.y * w;
                }
              }
            }
            __syncwarp();
            continue;
          }
#pragma unroll 1
          for (int hh = 0; hh < FPW; ++hh) {
            if (h == hh && active) {
              const int base = (jc * p.hop) % p.ring_n;
#pragma unroll 4
              for (int n2 = n2_lo; n2 <= n2_hi; ++n2) {
                const int tap = L * n2 - left + l;
                if (tap >= 0 && tap < win) {
                  const float w = wtab[tap];
                  const float2 g = buf[(l + L * (n2 % (R / L))) * (L + 1) + n2 / (R / L)];
                  int idx = base + tap;
                  idx -= (idx >= p.ring_n) ? p.ring_n : 0;
                  if (KIND == kKindStft) ring2[idx] = __ffma2_rn(make_float2(g.y, g.x), make_float2(w, w), ring2[idx]);
                  else                   ring1[idx] = fmaf(g.y, w, ring1[idx]);
                }
              }
            }
            __syncwarp();
          }
          continue;
        }
        // [region: C epilogue]
        // ---- C. job 0 epilogue ------------------------------------------------------------------
        // Bin k = row + R*col sits at buf[row*(L+1) + col].  A lane owns rows l + L*j; for col < L/2 the
        // bin is k < N/2 and its mirror N-k is (R - row, L-1-col) -- or (0, L-col) inside row 0.  Bins 0
        // and N/2 (row 0, cols 0 and L/2) mirror themselves; lane 0 handles N/2 as the extra pair.
        if (KIND == kKindStft) {
          const float eps4 = 4.f * p.eps;
          if (active) {
            if (frame_equal) stft_epilogue<NFFT, GRAD, true>(buf, l, eps4, s1, s2, s3);
            else             stft_epilogue<NFFT, GRAD, false>(buf, l, eps4, s1, s2, s3);
          } else if (GRAD) {
            for (int i = l; i < R * (L + 1); i += L) buf[i] = make_float2(0.f, 0.f);
          }
        } else {
          // ---- mel: amplitudes -> banded projection -> log-mel L1 -> gradient spectrum ------------
          // pass 1: X stays in the bin's own slot; (Ax, Ay) are parked in the mirror slot; bin 0 parks
          // them in the pad slot of row 0, bin N/2 keeps them in place and moves X to the pad slot of row 1.
          constexpr int EX0 = L, EX1 = (L + 1) + L;
          const float eps4 = 4.f * p.eps;
#pragma unroll
          for (int j = 0; j < R / L; ++j) {
            const int row = l + L * j;
            float2* pa = buf + row * (L + 1);
            float2* pb = (row == 0) ? buf + L : buf + (R - row) * (L + 1) + (L - 1);    // mirror of column c is pb[-c]
            if (frame_equal) {
              mel_pair_amp<true>(pa, row == 0 ? pa : pb, pa, row == 0 ? buf + EX0 : pb, eps4);
#pragma unroll 1
              for (int c = 1; c < L / 2; ++c) mel_pair_amp<true>(pa + c, pb - c, pa + c, pb - c, eps4);
            } else {
              mel_pair_amp<false>(pa, row == 0 ? pa : pb, pa, row == 0 ? buf + EX0 : pb, eps4);
#pragma unroll (L == 32 ? 5 : 7)
              for (int c = 1; c < L / 2; ++c) mel_pair_amp<false>(pa + c, pb - c, pa + c, pb - c, eps4);
            }
          }
          if (l == 0) {                                        // bin N/2: amplitudes stay, X moves to the pad slot
            if (frame_equal) mel_pair_amp<true>(buf + L / 2, buf + L / 2, buf + EX1, buf + L / 2, eps4);
            else             mel_pair_amp<false>(buf + L / 2, buf + L / 2, buf + EX1, buf + L / 2, eps4);
          }
          __syncwarp();
          // pass 2: balanced projection.  Every mel row is summed by a group of 1..L lanes walking a
          // host-built table of (amplitude slot, weight) entries, then reduced with shuffles.
          for (int r = 0; r < p.mel_rounds; ++r) {
            const int4 tk = mel_tasks[r * L + l];
            const int row = tk.x & 0xfff, grp = (tk.x >> 12) & 0xff, iters = tk.x >> 20;
            const int2* en = mel_entries + tk.y * L + l;
            float mx = 0.f, my = 0.f;
#pragma unroll 4
            for (int s = 0; s < iters; ++s) {
              const int2 e = en[s * L];
              const float2 amp = buf[e.x];
              const float w = bits_to_float(e.y);
              mx = fmaf(amp.x, w, mx);
              my = fmaf(amp.y, w, my);
            }
#pragma unroll
            for (int o = 1; o < L; o <<= 1) {
              const float tx = __shfl_xor_sync(0xffffffffu, mx, o), ty = __shfl_xor_sync(0xffffffffu, my, o);
              if (o < grp) { mx += tx; my += ty; }
            }
            if (row != 0xfff && (l & (grp - 1)) == 0) msum[row] = make_float2(mx, my);
          }
          __syncwarp();
          for (int row = l; row < p.n_mels; row += L) {
            const float2 mm = msum[row];
            const float mxc = fmaxf(mm.x, p.eps), myc = fmaxf(mm.y, p.eps);
            const float dl = (mxc == myc) ? 0.f : (logf(mxc) - logf(myc)) * p.inv_ln_base;
            if (active) s1 += fabsf(dl);
            if (GRAD) {
              const float sgn = (dl > 0.f) ? 1.f : ((dl < 0.f) ? -1.f : 0.f);
              msum[row].x = (mm.x >= p.eps) ? sgn * p.inv_ln_base / mxc : 0.f;     // gM[row]
            }
          }
          __syncwarp();
          if (GRAD) {
            // pass 3: gA[k] = sum_m gM[m] W[k,m] (<= 2 adjacent rows), H[k] = w gA gate / Ax * X
            const float eps4g = 4.f * p.eps;
#pragma unroll
            for (int j = 0; j < R / L; ++j) {
              const int row = l + L * j;
              float2* pa = buf + row * (L + 1);
              float2* pb = (row == 0) ? buf + L : buf + (R - row) * (L + 1) + (L - 1);
              mel_pair_grad(pa, row == 0 ? pa : pb, pa, row == 0, bin_tab[row], msum, eps4g, active);
#pragma unroll (L == 32 ? 5 : 7)
              for (int c = 1; c < L / 2; ++c)
                mel_pair_grad(pa + c, pb - c, pa + c, false, bin_tab[row + R * c], msum, eps4g, active);
            }
            if (l == 0) mel_pair_grad(buf + L / 2, buf + L / 2, buf + EX1, true, bin_tab[NFFT / 2], msum, eps4g, active);
          }
        }
        __syncwarp();
      }  // job
      if (GRAD && !no_ring) {
        // [region: F ring flush]
        // ---- F. flush the ring entries no later frame of this chunk touches ----------------------
        const int done = min(m_c, (step + 1) * FPW);
        const int limit = (done == m_c) ? span_c : done * p.hop;
        int idx = (flushed + lane) % p.ring_n;                  // one division per flush, then wrap by compare
        const int wrap = 32 % p.ring_n;
        for (int q = flushed + lane; q < limit; q += 32) {
          if (KIND == kKindStft) { out2[q] = ring2[idx]; ring2[idx] = make_float2(0.f, 0.f); }
          else                   { out1[q] = ring1[idx]; ring1[idx] = 0.f; }
          idx += wrap;
          idx -= (idx >= p.ring_n) ? p.ring_n : 0;
        }
        flushed = limit;
        __syncwarp();
      }
    }  // step

    // [region: partial sums]
    // ---- partial sums of this chunk (one writer per slot: deterministic) ------------------------
    s1 = warp_sum(s1);
    if (KIND == kKindStft) { s2 = warp_sum(s2); s3 = warp_sum(s3); }
    if (lane == 0) {
      if (KIND == kKindStft) {
        double* o = p.partials + (size_t)chunk_id * 3;
        o[0] = 0.25 * (double)s1; o[1] = 0.25 * (double)s2; o[2] = 0.34657359027997264 * (double)s3;  // 0.5 ln 2
      } else {
        p.partials[chunk_id] = (double)s1;
      }
    }
  }  // chunk
}

// ---------------------------------------------------------------------------------------------
// Explicit magnitude spectrogram, forward only: out[b, t, k] = sqrt(max(|STFT(x)[b, t, k]|^2, eps)),
// laid out (B, F, ld >= K) -- the tensor stft() returns (losses/stft_loss.py:19-35) and the A operand of
// the mel projection GEMM.  Frames t and t+1 of the SAME signal share one complex FFT (real / imaginary
// slot), so the work per frame is half of a naive real transform.
// ---------------------------------------------------------------------------------------------
struct SpecParams {
  const float* x;        // (B, T)
  int B, T;
  int hop, win, left, n_frames;
  int n_pairs;           // frame pairs per utterance = (n_frames + 1) / 2
  float eps;
  const float* window;
  const float2* twiddle;
  float* out;            // (B, n_frames, ld)
  int ld;
};

template <int NFFT>
SPL_DEVICE void spec_load_tables(const SpecParams& p, float* smem, int tid, int nthreads) {
  constexpr int L = FftGeom<NFFT>::L;
  const CtaTables ct = cta_tables(NFFT, p.win, kKindStft, L, 0, 0);
  const float* tw = reinterpret_cast<const float*>(p.twiddle);
  for (int i = tid; i < 2 * NFFT; i += nthreads) smem[ct.tw + i] = __ldg(&tw[i]);
  for (int i = tid; i < p.win; i += nthreads) smem[ct.win + i] = __ldg(&p.window[i]);
}

template <int NFFT>
SPL_DEVICE void spec_body(const SpecParams& p, float* smem, int block, int tid, int grid, int wpc) {
  using G = FftGeom<NFFT>;
  using SL = SmemLayout<NFFT, kKindStft, false>;
  constexpr int L = G::L, R = G::R, FPW = SL::FPW, HALF = NFFT / 2;
  const int warp = tid >> 5, lane = tid & 31;
  const int l = lane & (L - 1), h = lane / L;
  const CtaTables ct = cta_tables(NFFT, p.win, kKindStft, L, 0, 0);
  const float2* tw = reinterpret_cast<const float2*>(smem + ct.tw);
  const float* wtab = smem + ct.win;
  float2* buf = reinterpret_cast<float2*>(smem + ct.total + (size_t)warp * SL::words_per_warp(0, 0)) + h * SL::BUF_F2;
  const int total = p.B * p.n_pairs;                 // work items: (utterance, frame pair)
  const float eps4 = 4.f * p.eps;
  const int steps = (total + grid * wpc * FPW - 1) / (grid * wpc * FPW);
  for (int it = 0; it < steps; ++it) {
    const int item = (it * grid * wpc + block * wpc + warp) * FPW + h;
    const bool active = item < total;
    const int b = active ? item / p.n_pairs : 0, pr = active ? item - b * p.n_pairs : 0;
    const int t = 2 * pr;
    const bool second = active && (t + 1 < p.n_frames);
    const float* __restrict__ xb = p.x + (size_t)b * p.T;
    float2 v[R];
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) {
      const int tap = L * n2 - p.left + l;
      float a0 = 0.f, a1 = 0.f;
      if (active && tap >= 0 && tap < p.win) {
        const float w = wtab[tap];
        const int s = t * p.hop + L * n2 + l - HALF;
        a0 = __ldg(&xb[reflect(s, p.T)]) * w;
        if (second) a1 = __ldg(&xb[reflect(s + p.hop, p.T)]) * w;
      }
      v[n2] = make_float2(a0, a1);
    }
    fft_core<NFFT>(v, buf, tw, l);
    if (active) {
      float* o0 = p.out + ((size_t)b * p.n_frames + t) * p.ld;
      float* o1 = o0 + p.ld;
#pragma unroll
      for (int j = 0; j < R / L; ++j) {
        const int row = l + L * j;
        const float2* pa = buf + row * (L + 1);
        const float2* pb = (row == 0) ? buf + L : buf + (R - row) * (L + 1) + (L - 1);
#pragma unroll 4
        for (int c = 0; c <= L / 2; ++c) {
          if (c == L / 2 && row != 0) break;                    // bin N/2 lives in row 0 only
          const bool self = (row == 0) && (c == 0 || c == L / 2);
          const float2 a = pa[c], bm = self ? a : pb[-c];
          const float2 x2 = __fadd2_rn(a, make_float2(bm.x, -bm.y));                           // 2 X_t[k]
          const float2 y2 = __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));       // 2 X_{t+1}[k]
          const float p0 = fmaxf(fmaf(x2.x, x2.x, x2.y * x2.y), eps4);
          const float p1 = fmaxf(fmaf(y2.x, y2.x, y2.y * y2.y), eps4);
          const int k = row + R * c;
          o0[k] = 0.5f * p0 * spl_fast_rsqrt(p0);
          if (second) o1[k] = 0.5f * p1 * spl_fast_rsqrt(p1);
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// deterministic reduction of the per-chunk partial sums: one CTA per output sum
// ---------------------------------------------------------------------------------------------
struct ReduceParams {
  int n_sums;
  const double* base[16];   // first element of the column
  int stride[16];           // doubles between consecutive items
  int count[16];            // items
  double* out;              // [n_sums]
};

SPL_DEVICE void reduce_body(const ReduceParams& p, double* sh, int block, int tid, int nthreads) {
  const double* src = p.base[block];
  const int stride = p.stride[block], count = p.count[block];
  double acc = 0.0;
  for (int i = tid; i < count; i += nthreads) acc += src[(size_t)i * stride];
  sh[tid] = acc;
  __syncthreads();
  for (int o = nthreads >> 1; o > 0; o >>= 1) {
    if (tid < o) sh[tid] += sh[tid + o];
    __syncthreads();
  }
  if (tid == 0) p.out[block] = sh[0];
}

// ---------------------------------------------------------------------------------------------
// losses and backward coefficients from the (all-reduced) sums.  One thread.
// ---------------------------------------------------------------------------------------------
struct FinalizeParams {
  int n;                    // transforms
  int kind[8];
  double count[8];          // global element count of the transform's mean
  int sum_ofs[8];           // offset of the transform's sums in `sums`
  const double* sums;
  float* sc;                // may be null when no STFT transform is present
  float* mag;
  float* mel;
  float* coefs;             // [2 * n] : stft (sc, mag) ; mel (mel, 0)
};

SPL_DEVICE void finalize_body(const FinalizeParams& p) {
  int n_stft = 0, n_mel = 0;
  for (int r = 0; r < p.n; ++r) (p.kind[r] == kKindStft ? n_stft : n_mel) += 1;
  double sc = 0.0, mag = 0.0, mel = 0.0;
  for (int r = 0; r < p.n; ++r) {
    const double* s = p.sums + p.sum_ofs[r];
    if (p.kind[r] == kKindStft) {
      const double d = sqrt(__ldcg(&s[0])), ny = sqrt(__ldcg(&s[1]));
      sc += d / ny;
      mag += __ldcg(&s[2]) / p.count[r];
      p.coefs[2 * r] = (d > 0.0) ? (float)(1.0 / (n_stft * d * ny)) : 0.f;
      p.coefs[2 * r + 1] = (float)(1.0 / (n_stft * p.count[r]));
    } else {
      mel += __ldcg(&s[0]) / p.count[r];
      p.coefs[2 * r] = (float)(1.0 / (n_mel * p.count[r]));
      p.coefs[2 * r + 1] = 0.f;
    }
  }
  if (p.sc && n_stft) *p.sc = (float)(sc / n_stft);
  if (p.mag && n_stft) *p.mag = (float)(mag / n_stft);
  if (p.mel && n_mel) *p.mel = (float)(mel / n_mel);
}

// ---------------------------------------------------------------------------------------------
// backward: dx[b, i] = sum over transforms of coef * (overlap-added frame gradients), gathered
// from the per-chunk slots (no atomics: every slot has one writer, every dx sample one reader)
// and folded over the reflect-padding margins.
// ---------------------------------------------------------------------------------------------
struct CombineEntry {
  const void* chunks;
  int kind, half, hop, win, left, m, n_chunks, span, n_frames;
};
struct CombineParams {
  int n;
  CombineEntry e[8];
  const float* coefs;    // from finalize
  const float* g_sc;     // upstream gradients (device scalars); null => 0
  const float* g_mag;
  const float* g_mel;
  float* dx;             // (B, T)
  int B, T;
};

// overlap-added value at one padded position (used for the reflect-fold margins)
SPL_DEVICE float gather_padded(const CombineEntry& e, int b, int ppos, float cu, float cv) {
  const int q_abs = ppos - e.left;
  if (q_abs < 0 || q_abs >= (e.n_frames - 1) * e.hop + e.win) return 0.f;
  const int mh = e.m * e.hop;
  int c = min(q_abs / mh, e.n_chunks - 1);
  float acc = 0.f;
  for (; c >= 0; --c) {
    const int q = q_abs - c * mh;
    if (q >= e.span) break;
    const int m_c = min(e.m, e.n_frames - c * e.m);
    if (q < (m_c - 1) * e.hop + e.win) {
      const size_t o = ((size_t)b * e.n_chunks + c) * e.span + q;
      if (e.kind == kKindStft) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(e.chunks) + o);
        acc = fmaf(cu, v.x, fmaf(cv, v.y, acc));
      } else {
        acc = fmaf(cu, __ldg(reinterpret_cast<const float*>(e.chunks) + o), acc);
      }
    }
  }
  return acc;
}

// same for four consecutive padded positions ppos0 .. ppos0+3: one chunk walk, 16-byte loads when the
// four samples sit inside one slot and the address is aligned (always the case for the shipped configs)
SPL_DEVICE void gather_padded4(const CombineEntry& e, int b, int ppos0, float cu, float cv, float (&acc)[4]) {
  const int q0 = ppos0 - e.left;
  if (q0 + 3 < 0 || q0 >= (e.n_frames - 1) * e.hop + e.win) return;
  const int mh = e.m * e.hop;
  int c = min((q0 + 3) / mh, e.n_chunks - 1);
  for (; c >= 0; --c) {
    const int q = q0 - c * mh;
    if (q >= e.span) break;
    const int m_c = min(e.m, e.n_frames - c * e.m);
    const int lim = (m_c - 1) * e.hop + e.win;
    const size_t o = ((size_t)b * e.n_chunks + c) * e.span + q;     // meaningful for q >= 0 only
    if (e.kind == kKindStft) {
      const float2* src = reinterpret_cast<const float2*>(e.chunks);
      if (q >= 0 && q + 3 < lim && (o & 1) == 0) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + o));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + o + 2));
        acc[0] = fmaf(cu, v0.x, fmaf(cv, v0.y, acc[0]));
        acc[1] = fmaf(cu, v0.z, fmaf(cv, v0.w, acc[1]));
        acc[2] = fmaf(cu, v1.x, fmaf(cv, v1.y, acc[2]));
        acc[3] = fmaf(cu, v1.z, fmaf(cv, v1.w, acc[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (q + j >= 0 && q + j < lim) {
            const float2 v = __ldg(src + (o + j));
            acc[j] = fmaf(cu, v.x, fmaf(cv, v.y, acc[j]));
          }
      }
    } else {
      const float* src = reinterpret_cast<const float*>(e.chunks);
      if (q >= 0 && q + 3 < lim && (o & 3) == 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + o));
        acc[0] = fmaf(cu, v.x, acc[0]); acc[1] = fmaf(cu, v.y, acc[1]);
        acc[2] = fmaf(cu, v.z, acc[2]); acc[3] = fmaf(cu, v.w, acc[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (q + j >= 0 && q + j < lim) acc[j] = fmaf(cu, __ldg(src + (o + j)), acc[j]);
      }
    }
  }
}

// one thread = four consecutive output samples of one utterance
SPL_DEVICE void combine_body(const CombineParams& p, long long gid) {
  const int quads = (p.T + 3) >> 2;
  if (gid >= (long long)p.B * quads) return;
  const int b = (int)(gid / quads), i0 = 4 * (int)(gid - (long long)b * quads);
  const float gsc = p.g_sc ? *p.g_sc : 0.f, gmag = p.g_mag ? *p.g_mag : 0.f, gmel = p.g_mel ? *p.g_mel : 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < p.n; ++r) {
    const CombineEntry& e = p.e[r];
    float cu, cv;
    if (e.kind == kKindStft) { cu = gsc * p.coefs[2 * r]; cv = gmag * p.coefs[2 * r + 1]; }
    else { cu = gmel * p.coefs[2 * r]; cv = 0.f; }
    const int P = e.half;
    gather_padded4(e, b, P + i0, cu, cv, acc);
    if (i0 <= P || i0 + 3 >= p.T - 1 - P) {          // reflect-fold margins
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j;
        if (i >= 1 && i <= P) acc[j] += gather_padded(e, b, P - i, cu, cv);
        if (i >= p.T - 1 - P && i <= p.T - 2) acc[j] += gather_padded(e, b, P + 2 * (p.T - 1) - i, cu, cv);
      }
    }
  }
  float* out = p.dx + (size_t)b * p.T + i0;
  if ((p.T & 3) == 0) {
    *reinterpret_cast<float4*>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j < p.T) out[j] = acc[j];
  }
}

// ---------------------------------------------------------------------------------------------
// reduce + finalize in one launch (single-GPU path): the last CTA to finish its column computes
// the losses.  `counter` must be zero before the first use; atomicInc wraps it back to zero.
// ---------------------------------------------------------------------------------------------
struct ReduceFinalizeParams {
  ReduceParams r;
  FinalizeParams f;
  unsigned* counter;
};

#ifndef SPECLOSS_EMU
// One persistent CTA per SM: up to 12 warps (168 registers) for the 64-point-per-lane kernels, 16 (128
// registers) for the others; the host picks the actual warp count from the shared-memory budget.
template <int NFFT> struct MaxWarps { static constexpr int value = NFFT == 2048 ? 12 : SPL_MAX_WARPS_SMALL; };

template <int NFFT, int KIND, bool GRAD, int WIN_T>
__global__ void __launch_bounds__(MaxWarps<NFFT>::value * 32, 1) transform_kernel(const TransformParams p) {
  extern __shared__ __align__(16) float smem_dyn[];
  cta_load_tables<NFFT, KIND>(p, smem_dyn, threadIdx.x, blockDim.x);
  __syncthreads();
  transform_body<NFFT, KIND, GRAD, WIN_T>(p, smem_dyn, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
template <int NFFT>
__global__ void __launch_bounds__(MaxWarps<NFFT>::value * 32, 1) spec_kernel(const SpecParams p) {
  extern __shared__ __align__(16) float smem_dyn[];
  spec_load_tables<NFFT>(p, smem_dyn, threadIdx.x, blockDim.x);
  __syncthreads();
  spec_body<NFFT>(p, smem_dyn, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
__global__ void __launch_bounds__(256) reduce_kernel(const ReduceParams p) {
  __shared__ double sh[256];
  reduce_body(p, sh, blockIdx.x, threadIdx.x, 256);
}
__global__ void finalize_kernel(const FinalizeParams p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) finalize_body(p);
}
__global__ void __launch_bounds__(256) reduce_finalize_kernel(const ReduceFinalizeParams p) {
  __shared__ double sh[256];
  __shared__ bool last;
  reduce_body(p.r, sh, blockIdx.x, threadIdx.x, 256);
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicInc(p.counter, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    finalize_body(p.f);
  }
}
__global__ void __launch_bounds__(128) combine_kernel(const CombineParams p) {
  combine_body(p, (long long)blockIdx.x * 128 + threadIdx.x);
}
#endif

}  // namespace spl
