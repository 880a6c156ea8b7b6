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
// Code-size experiments, all measured slower on B200 (gpurun_out/r1k_variants.txt) because ptxas starts spilling
// the tap registers: 2048-point kernels with rolled row loops (rows parked in the slot), and the two R-point
// codelet sites merged into a 2-iteration loop.
#ifndef SPL_PARK_2048
#define SPL_PARK_2048 0
#endif
#ifndef SPL_PHASE_LOOP
#define SPL_PHASE_LOOP 0
#endif
// Third experiment, also slower (gpurun_out/r1k_sync.txt: stft-2048 206 -> 232 us at 32 x 4 s): a CTA barrier once
// per frame round in the 2048-point kernels, meant to keep the warps of a CTA in the same region of the ~100 KB
// instruction stream (shared instruction-cache lines); lock-stepped warps contend for the same pipe instead.
#ifndef SPL_SYNC_2048
#define SPL_SYNC_2048 0
#endif

namespace spl {

constexpr int kKindStft = 0;
constexpr int kKindMel = 1;
constexpr int kPhaseUnroll = SPL_PHASE_LOOP ? 1 : 2;   // 1: the two phases of a frame share one copy of the R-point codelet

// ---------------------------------------------------------------------------------------------
// FFT geometry: N = L * R.  A group of L lanes owns one frame; every lane holds R complex points.
//   forward : in-lane R-point DFTs over n2 (n = n1 + L*n2, n1 = lane), twiddle W_N^(n1*k2), transpose through
//             shared memory, in-lane L-point DFTs over n1 -> bin k = k2 + R*k1, row k2 owned by lane k2 % L,
//             LEFT IN REGISTERS;
//   inverse : the transposed algorithm on component-swapped data: L-point DFTs over k1 (registers), the same
//             twiddle, transpose, R-point DFTs over k2 -> sample n = n1 + L*n2 back in lane n1.
// The frame slot in shared memory is R rows of L (+1 pad) float2; its only uses are the two transposes and the
// exchange of mirror halves (columns >= L/2 of every row) between the lanes that own rows k2 and R - k2.
// ---------------------------------------------------------------------------------------------
template <int NFFT> struct FftGeom;
template <> struct FftGeom<512>  { static constexpr int L = 16, R = 32; };
template <> struct FftGeom<1024> { static constexpr int L = 32, R = 32; };
template <> struct FftGeom<2048> { static constexpr int L = 32, R = 64; };

template <int NFFT> struct Geo {
  static constexpr int L = FftGeom<NFFT>::L, R = FftGeom<NFFT>::R;
  static constexpr int PITCH = L + 1;       // float2 per row: odd, so row-wise and column-wise accesses are conflict-free
  static constexpr int RPL = R / L;         // rows per lane in the L-point passes
  static constexpr int FPW = 32 / L;        // frames in flight per warp
  static constexpr int HL = L / 2;          // columns kept in registers per row (bins below N/2)
  static constexpr int SLOT_F2 = R * PITCH; // float2 per frame slot
  // Rows per lane held in registers between the passes.  2048-point transforms (two 32-column rows per lane) keep
  // ONE row at a time and park the other in its own slot positions: their row loops stay rolled, which halves the
  // code (an unrolled kernel is ~100 KB of SASS against a 32 KB instruction cache: no_instruction stalls 1.2-2.0
  // per issue, profiles r1j).
  static constexpr bool PARK = SPL_PARK_2048 && NFFT == 2048;
  static constexpr int AROWS = PARK ? 1 : RPL;
};

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
  float eps;
  const float* window;   // win taps
  const float2* twiddle; // [R][L] : W_N^(n1*k2) at [k2 * L + n1]
  double* partials;      // [grid * warps per CTA][n_sums] : one row per warp
  void* gframes;         // [B * n_frames][win] float2 (stft: u = sc part, v = log-mag part) | float (mel);
                         // run_frames > 1: [B * runs_per_utt][run_len] float2, the frames of a run overlap-added
  // Overlap-add in shared memory before the gradient leaves the SM (STFT loss, large batches): a warp takes a RUN of
  // run_frames consecutive frames of one utterance, accumulates their windowed gradients in a ring of `win` taps and
  // writes every sample of the run once: (run_frames - 1) * hop + win taps per run instead of run_frames * win.
  int run_frames;        // 1: one gradient slot per frame (no ring)
  int runs_per_utt;      // ceil(n_frames / run_frames)
  int run_len;           // (run_frames - 1) * hop + win
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

SPL_DEVICE double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Constant tables of a transform, staged once per CTA in shared memory (one persistent CTA per SM): at 200+ KB
// of shared memory per SM the L1 keeps only ~28 KB, and twiddles + window + mel tables (30-64 KB) would
// otherwise be re-fetched from L2 at every frame.  The twiddle rows are padded to PITCH so that the forward
// pass (lane = n1, row fixed) and the inverse pass (lane = row, n1 fixed) both read without bank conflicts.
struct CtaTables {
  int tw, win, tasks, entries, bintab, total;     // word offsets
};
SPL_DEVICE int align4(int n) { return (n + 3) & ~3; }
static __host__ __device__ inline CtaTables cta_tables(int n_fft, int win, int kind, int mel_rounds, int mel_entry_rows) {
  const int lanes = n_fft == 512 ? 16 : 32, rows = n_fft / lanes;
  CtaTables t;
  int o = 0;
  t.tw = o;      o += (2 * rows * (lanes + 1) + 3) & ~3;
  t.win = o;     o += (win + 3) & ~3;
  t.tasks = o;   o += kind == kKindMel ? 4 * mel_rounds * lanes : 0;
  t.entries = o; o += kind == kKindMel ? ((2 * mel_entry_rows * lanes + 3) & ~3) : 0;
  t.bintab = o;  o += kind == kKindMel ? 4 * (n_fft / 2 + 1) : 0;
  t.total = o;
  return t;
}

// shared memory per warp (in 4-byte words); the host side uses the same function
template <int NFFT, int KIND>
struct SmemLayout {
  using G = Geo<NFFT>;
  // ring_taps: window taps of the overlap-add ring (0 = none), one ring of float2 per frame in flight
  static __host__ __device__ int words_per_warp(int n_mels, int ring_taps = 0) {
    int w = G::FPW * G::SLOT_F2 * 2;
    if (KIND == kKindMel) w += G::FPW * 2 * ((n_mels + 3) & ~3);
    w = (w + 3) & ~3;
    return w + G::FPW * 2 * ((ring_taps + 1) & ~1);
  }
};

// reflect index without edge repeat (torch.stft center=True, pad_mode="reflect")
SPL_DEVICE int reflect(int s, int T) {
  s = s < 0 ? -s : s;
  return s >= T ? 2 * (T - 1) - s : s;
}

// Opaque copy: stops the compiler from hoisting everything derived from the value out of the enclosing loop
// (with rolled row loops, hoisted per-lane addresses otherwise stay live across the tap loads and force spills).
template <typename T>
SPL_DEVICE T launder(T v) {
#ifndef SPECLOSS_EMU
  asm volatile("" : "+r"(v));
#endif
  return v;
}
template <typename T>
SPL_DEVICE T* launder_ptr(T* v) {
#ifndef SPECLOSS_EMU
  asm volatile("" : "+l"(v));
#endif
  return v;
}

SPL_DEVICE unsigned float_bits(float f) {
#ifdef SPECLOSS_EMU
  unsigned u; memcpy(&u, &f, 4); return u;
#else
  return __float_as_uint(f);
#endif
}

SPL_DEVICE double __longlong_as_double_nan() {
#ifdef SPECLOSS_EMU
  return std::nan("");
#else
  return __longlong_as_double(0x7ff8000000000000LL);
#endif
}

SPL_DEVICE float bits_to_float(int b) {
#ifdef SPECLOSS_EMU
  float f; memcpy(&f, &b, 4); return f;
#else
  return __int_as_float(b);
#endif
}

// CTA prologue: every thread copies its share of the constant tables into shared memory -- with cp.async, so that all
// the copies of a thread are in flight together (as register-staged 4-byte loads the prologue cost one L2 round trip
// per element and thread: 16 % of the stall samples of the mel kernel at configs[1], profiles r1v).
#ifdef SPECLOSS_EMU
static inline void cp_async4(void* d, const void* s) { memcpy(d, s, 4); }
static inline void cp_async8(void* d, const void* s) { memcpy(d, s, 8); }
static inline void cp_async16(void* d, const void* s) { memcpy(d, s, 16); }
static inline void cp_async_wait_all() {}
#else
__device__ __forceinline__ void cp_async4(void* d, const void* s) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void cp_async8(void* d, const void* s) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void cp_async16(void* d, const void* s) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

// n 4-byte words, global -> shared (dst 16-byte aligned by construction of CtaTables): 16-byte chunks when the source
// is 16-byte aligned, words otherwise and for the tail
SPL_DEVICE void cta_copy_words(float* dst, const void* src, int n, int tid, int nthreads) {
  const char* s8 = reinterpret_cast<const char*>(src);
  const int n16 = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? n >> 2 : 0;
  for (int i = tid; i < n16; i += nthreads) cp_async16(dst + 4 * i, s8 + 16 * i);
  for (int i = 4 * n16 + tid; i < n; i += nthreads) cp_async4(dst + i, s8 + 4 * i);
}

template <int NFFT>
SPL_DEVICE void cta_load_fft_tables(const CtaTables& ct, const float2* twiddle, const float* window, int win,
                                    float* smem, int tid, int nthreads) {
  using G = Geo<NFFT>;
  float2* tw = reinterpret_cast<float2*>(smem + ct.tw);
  for (int i = tid; i < NFFT; i += nthreads) cp_async8(&tw[(i / G::L) * G::PITCH + (i % G::L)], &twiddle[i]);
  cta_copy_words(smem + ct.win, window, win, tid, nthreads);
}

// ends with the wait for this thread's copies; the caller's __syncthreads() publishes the tables
template <int NFFT, int KIND>
SPL_DEVICE void cta_load_tables(const TransformParams& p, float* smem, int tid, int nthreads) {
  constexpr int L = Geo<NFFT>::L;
  const CtaTables ct = cta_tables(NFFT, p.win, KIND, p.mel_rounds, p.mel_entry_rows);
  cta_load_fft_tables<NFFT>(ct, p.twiddle, p.window, p.win, smem, tid, nthreads);
  if (KIND == kKindMel) {
    cta_copy_words(smem + ct.tasks, p.mel_tasks, 4 * p.mel_rounds * L, tid, nthreads);
    cta_copy_words(smem + ct.entries, p.mel_entries, 2 * p.mel_entry_rows * L, tid, nthreads);
    cta_copy_words(smem + ct.bintab, p.bin_tab, 4 * (NFFT / 2 + 1), tid, nthreads);
  }
  cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------
// FFT passes
// ---------------------------------------------------------------------------------------------
// [region: fwd twiddle + store]
// Forward, after the in-lane R-point DFT of this lane's points (element n = l + L*n2 in v[n2]): twiddle, column l
// of the slot.  Ends with the warp barrier that makes the rows readable.
template <int NFFT>
SPL_DEVICE void fwd_store_cols(const float2 (&v)[Geo<NFFT>::R], float2* S, const float2* tw, int l) {
  using G = Geo<NFFT>;
#pragma unroll
  for (int k2 = 0; k2 < G::R; ++k2) S[k2 * G::PITCH + l] = k2 > 0 ? cmul(v[k2], tw[k2 * G::PITCH + l]) : v[0];
  __syncwarp();
}

// [region: fwd pass B]
// Forward, second half, for the rows l + L*j of this lane: L-point DFT of the row.  Columns >= L/2 go back to the
// row in place, where the lane owning the mirror row R - row picks them up; columns < L/2 (bins below N/2) stay in
// registers (A) -- or, PARK, go back to the row as well.  Ends with the warp barrier that publishes them.
template <int NFFT>
SPL_DEVICE void fwd_pass_b(float2 (&A)[Geo<NFFT>::AROWS][Geo<NFFT>::HL], float2* S, int l) {
  using G = Geo<NFFT>;
#pragma unroll(G::PARK ? 1 : G::RPL)
  for (int j = 0; j < G::RPL; ++j) {
    float2* row = S + (l + G::L * j) * G::PITCH;
    float2 b[G::L];
#pragma unroll
    for (int n1 = 0; n1 < G::L; ++n1) b[n1] = row[n1];
    Dft<G::L>::run(b);
#pragma unroll
    for (int k1 = 0; k1 < G::HL; ++k1) {
      if (G::PARK) row[k1] = b[k1];
      else         A[j][k1] = b[k1];
    }
#pragma unroll
    for (int k1 = G::HL; k1 < G::L; ++k1) row[k1] = b[k1];
  }
  __syncwarp();
}

// pointer to the mirror of column 0 of `row`: the mirror of (row, k1) is pb[-k1].  Row 0 mirrors into itself
// ((0, k1) <-> (0, L - k1)); its column 0 (bin 0) has no partner: pb[0] is then the pad word of row 0.
template <int NFFT>
SPL_DEVICE float2* mirror_ptr(float2* S, int row) {
  using G = Geo<NFFT>;
  return row == 0 ? S + G::L : S + ((G::R - row) & (G::R - 1)) * G::PITCH + (G::L - 1);
}

// [region: inv pass B]
// Inverse, first half: gradient spectrum rows (columns < L/2 in A -- PARK: in the slot --, columns >= L/2 in the
// slot) -> swapped components -> L-point DFT -> twiddle -> back to the row.  Ends with a warp barrier.
template <int NFFT>
SPL_DEVICE void inv_pass_b(float2 (&A)[Geo<NFFT>::AROWS][Geo<NFFT>::HL], float2* S, const float2* tw, int l) {
  using G = Geo<NFFT>;
#pragma unroll(G::PARK ? 1 : G::RPL)
  for (int j = 0; j < G::RPL; ++j) {
    const int r = l + G::L * j;
    float2* row = S + r * G::PITCH;
    const float2* twr = tw + r * G::PITCH;
    float2 b[G::L];
#pragma unroll
    for (int k1 = 0; k1 < G::HL; ++k1) {
      const float2 h = G::PARK ? row[k1] : A[G::PARK ? 0 : j][k1];
      b[k1] = make_float2(h.y, h.x);
    }
#pragma unroll
    for (int k1 = G::HL; k1 < G::L; ++k1) { const float2 h = row[k1]; b[k1] = make_float2(h.y, h.x); }
    Dft<G::L>::run(b);
    row[0] = b[0];
#pragma unroll
    for (int n1 = 1; n1 < G::L; ++n1) row[n1] = cmul(b[n1], twr[n1]);
  }
  __syncwarp();
}

// Register slots whose taps enter the level statistic of equalise_pair(): four per lane, spread over the slots the window
// covers completely (compile-time for the shipped windows).  With L lanes that is 4 L taps per frame, each lane's share a
// uniformly strided subsample; the estimator (mean binade of the lane maxima, applied to both signals alike) needs no more,
// and every tap would cost 2 R FMNMX per frame instead of 8.
template <int NFFT, int WIN_T>
SPL_DEVICE constexpr bool level_slot(int n2) {
  constexpr int L = Geo<NFFT>::L, R = Geo<NFFT>::R;
  constexpr int left = WIN_T > 0 ? (NFFT - WIN_T) / 2 : 0;
  constexpr int first = WIN_T > 0 ? (left + L - 1) / L : 0;                 // first slot whose L taps all lie under the window
  constexpr int last = WIN_T > 0 ? (left + WIN_T) / L - 1 : R - 1;          // last such slot
  constexpr int span = last - first + 1;
  return n2 == first + span / 8 || n2 == first + (3 * span) / 8 || n2 == first + (5 * span) / 8 || n2 == first + (7 * span) / 8;
}

// [region: tap load]
// Taps of frame t of utterance rows xb / yb: reflect-pad, window, pack z = x*w + i*y*w into v[n2] (element
// n = l + L*n2).  Returns whether every tap of this lane has x*w == y*w bit for bit; amax = (max |x w|, max |y w|) over
// four sampled taps of this lane (level_slot; FMNMX: not on the FMA pipe), the input of equalise_pair() below.
template <int NFFT, int WIN_T>
SPL_DEVICE bool load_taps(float2 (&v)[Geo<NFFT>::R], const float* __restrict__ xb, const float* __restrict__ yb, int T,
                          int s0, int win, int left, const float* wtab, int l, bool active, float2& amax) {
  using G = Geo<NFFT>;
  constexpr int L = G::L, R = G::R;
  const bool interior = (s0 + left >= 0) && (s0 + left + win <= T);
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
      if (level_slot<NFFT, WIN_T>(n2)) amax = make_float2(fmaxf(amax.x, fabsf(xy.x)), fmaxf(amax.y, fabsf(xy.y)));
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
        const int sidx = reflect(s0 + L * n2 + l, T);
        const float w = wtab[tap];
        xv = __ldg(&xb[sidx]) * w;
        yv = __ldg(&yb[sidx]) * w;
      }
      same = same && (xv == yv);
      if (level_slot<NFFT, WIN_T>(n2)) amax = make_float2(fmaxf(amax.x, fabsf(xv)), fmaxf(amax.y, fabsf(yv)));
      v[n2] = make_float2(xv, yv);
    }
  }
  return same;
}

// Prediction and target share one complex FFT (z = x w + i y w), whose rounding error in EVERY bin is eps * rms(Z): a
// prediction much weaker than the target (an untrained decoder: |X^| << |Y|) would inherit the target's rounding noise, and
// the log-magnitude gradient ~ 1/|X^| amplifies it (measured: gradient 5e-5 ... 2e-2 from fp64 where the reference's fp32 is
// at 1e-6, profiles/r4g_*).  So the prediction's taps are multiplied by the power of two that brings the frame's level of x w to
// the level of y w -- exact in fp32 -- and the unpacked spectrum 2 X^ is multiplied by the inverse power of two,
// again exact: the packed transform then sees two signals of equal level and each spectrum carries the rounding error of
// its own level, as in a transform of its own.  Bit-identical frames have equal maxima: no scaling, and the exact-zero
// property of loss(x, x) is untouched.  Returns 1 / s; v[].x is scaled in place.  Cost: ~3 % of a transform kernel.
// One binade of bias towards the prediction: the gradient is taken w.r.t. the prediction, whose PHASE at weak bins enters
// through X^ / |X^| and 1 / |X^|, while the target enters through magnitudes only -- so the prediction may be the louder
// of the two in the shared transform (measured on DC-dominated decoder outputs whose rms is half the target's: 8.0e-5 from
// fp64 without the bias, 4.7e-5 with it = the reference's fp32).  Equal levels still give shift = 1: untouched.
#ifndef SPL_EQ_BIAS
#define SPL_EQ_BIAS 1
#endif
template <int L, int R>
SPL_DEVICE float equalise_pair(float2 (&v)[R], float2 amax, unsigned grp_mask) {
  // Level of a signal in this frame = the mean binade, over the lanes, of each lane's largest sampled tap (a lane holds every
  // L-th tap): one integer redux.sync per signal.  Unlike the frame maximum it is not fooled by a single spike (an untrained
  // HiFiGAN decoder emits them: max / rms = 34 on the vocoder trainer's tensors), unlike an energy sum it costs nothing on
  // the FMA pipe.  A lane without any signal (digital silence) switches the equalisation off.
  const unsigned ex = float_bits(amax.x) >> 23, ey = float_bits(amax.y) >> 23;      // biased exponents; 0: zero / denormal
  const unsigned quiet = __reduce_min_sync(grp_mask, ex < ey ? ex : ey);
  const int sx = (int)__reduce_add_sync(grp_mask, ex), sy = (int)__reduce_add_sync(grp_mask, ey);
  constexpr int LOG_L = L == 32 ? 5 : 4;
  static_assert((1 << LOG_L) == L, "lanes per frame");
  int shift = quiet == 0 ? 0 : (sy - sx + L / 2 + SPL_EQ_BIAS * L) >> LOG_L;   // arithmetic shift: floor((d + L/2) / L) = round(d / L)
  // Levels within a factor of 4 stay as they are: that is every frame once training has brought the prediction near the
  // target, where the roundings of X^ and Y out of ONE transform are correlated and partly cancel in A_y - A_x (measured:
  // equalising by a single binade there doubles the gradient's distance from fp64, 2.8e-5 -> 6.1e-5).
  shift = (shift > -2 && shift < 2) ? 0 : shift;
  shift = shift < -96 ? -96 : (shift > 96 ? 96 : shift);
  if (shift != 0) {
    const float sc = bits_to_float((127 + shift) << 23);
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) v[n2].x *= sc;
  }
  return bits_to_float((127 - shift) << 23);
}

// ---------------------------------------------------------------------------------------------
// Epilogues on mirror pairs (k, N-k) of the packed spectrum Z = FFT(x + i y):
//   2X[k] = Z[k] + conj Z[N-k],  2Y[k] = -i (Z[k] - conj Z[N-k]).
// ---------------------------------------------------------------------------------------------
// [region: stft epilogue]
// STFT loss terms from the doubled spectra (powers x4, magnitudes x2; the sums are rescaled once at the end), and --
// GRAD -- the un-scaled gradient spectra:
//   H[k] = w/2 (alpha + i beta) (2X),  H[N-k] = w/2 (alpha + i beta) conj(2X),  w = 1/2 (1 when k mirrors itself)
//   alpha = gate (Ax - Ay)/Ax  (spectral convergence),  beta = gate sign(Ax - Ay)/Ax^2  (log magnitude).
// eq: prediction and target frames are bit-identical; Y := X then makes every difference term exactly zero
// (d = 0, lg2(p) - lg2(p) = 0, sign(0) = 0).
template <bool GRAD>
SPL_DEVICE void stft_pair(float2 a, float2 bm, bool self, bool eq, float eps4, float inv_s, float& s1, float& s2, float& s3,
                          float2& ha, float2& hb) {
  const float2 x2 = __fmul2_rn(__fadd2_rn(a, make_float2(bm.x, -bm.y)), make_float2(inv_s, inv_s));   // exact: power of two
  float2 y2 = __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));
  y2 = eq ? x2 : y2;
  const float px = fmaf(x2.x, x2.x, x2.y * x2.y);
  const float py = fmaf(y2.x, y2.x, y2.y * y2.y);
  const float pxc = fmaxf(px, eps4), pyc = fmaxf(py, eps4);
  const float rx = spl_fast_rsqrt(pxc), ry = spl_fast_rsqrt(pyc);
  const float ax = __fmul_rn(pxc, rx), ay = __fmul_rn(pyc, ry);       // 2 Ax, 2 Ay
  const float d = __fsub_rn(ay, ax);                                  // never contracted: 0 when pxc == pyc
  s1 = fmaf(d, d, s1);
  s2 += pyc;
  s3 += fabsf(__fsub_rn(spl_fast_log2(pyc), spl_fast_log2(pxc)));
  if (GRAD) {
    const float wq = self ? 0.5f : 0.25f;
    const float rxg = px >= eps4 ? rx : 0.f;                          // clamp gate of the reference
    const float sgn = fminf(fmaxf((pxc - pyc) * 1e25f, -1.f), 1.f);   // exact sign, 0 when equal
    const float gr = -wq * d * rxg;
    const float gi = 4.f * wq * sgn * rx * rxg;
    const float2 g2 = make_float2(gi, gi), r2 = make_float2(gr, gr);
    ha = __ffma2_rn(make_float2(-x2.y, x2.x), g2, __fmul2_rn(x2, r2));
    hb = __ffma2_rn(make_float2(x2.y, x2.x), g2, __fmul2_rn(make_float2(x2.x, -x2.y), r2));
  }
}

// all pairs of one lane.  In: columns < L/2 of this lane's rows in A (PARK: in the slot), mirror halves in the slot.
// Out (GRAD): H in the same places.
template <int NFFT, bool GRAD>
SPL_DEVICE void stft_epilogue(float2 (&A)[Geo<NFFT>::AROWS][Geo<NFFT>::HL], float2* S, int l, bool eq, float eps4, float inv_s,
                              float& s1, float& s2, float& s3) {
  using G = Geo<NFFT>;
#pragma unroll(G::PARK ? 1 : G::RPL)
  for (int j = 0; j < G::RPL; ++j) {
    const int row = l + G::L * j;
    float2* pb = mirror_ptr<NFFT>(S, row);
    float2* arow = S + row * G::PITCH;
#pragma unroll
    for (int k1 = 0; k1 < G::HL; ++k1) {
      const float2 a = G::PARK ? arow[k1] : A[G::PARK ? 0 : j][k1];
      float2 bm = pb[-k1];
      bool self = false;
      if (k1 == 0) {                                 // bin 0 (row 0) mirrors itself
        self = row == 0;
        bm = self ? a : bm;
      }
      float2 ha, hb;
      stft_pair<GRAD>(a, bm, self, eq, eps4, inv_s, s1, s2, s3, ha, hb);
      if (GRAD) {
        if (G::PARK) arow[k1] = ha;
        else         A[G::PARK ? 0 : j][k1] = ha;
        pb[-k1] = hb;
      }
    }
  }
  if (l == 0) {                                      // bin N/2 = (row 0, column L/2) mirrors itself
    const float2 a = S[G::HL];
    float2 ha, hb;
    stft_pair<GRAD>(a, a, true, eq, eps4, inv_s, s1, s2, s3, ha, hb);
    if (GRAD) S[G::HL] = ha;
  }
}

// [region: mel epilogue]
// mel, pass 1 on one mirror pair: 2X[k] and the amplitudes (Ax, Ay) = sqrt(max(|.|^2, eps)); eq: Y := X
SPL_DEVICE void mel_pair_amp(float2 a, float2 bm, bool eq, float eps4, float2& x2, float2& amp, float inv_s = 1.f) {
  x2 = __fmul2_rn(__fadd2_rn(a, make_float2(bm.x, -bm.y)), make_float2(inv_s, inv_s));      // un-scale (equalise_pair)
  float2 y2 = __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));
  y2 = eq ? x2 : y2;
  const float pxc = fmaxf(fmaf(x2.x, x2.x, x2.y * x2.y), eps4);
  const float pyc = fmaxf(fmaf(y2.x, y2.x, y2.y * y2.y), eps4);
  amp = __fmul2_rn(make_float2(0.5f * pxc, 0.5f * pyc), make_float2(spl_fast_rsqrt(pxc), spl_fast_rsqrt(pyc)));
}

// mel, pass 3 on one bin: gA[k] = gM[m0] W[k,m0] + gM[m0+1] W[k,m0+1];  H[k] = w gA gate / Ax * X
SPL_DEVICE float2 mel_bin_grad(float2 x2, bool self, int4 bt, const float2* msum, float eps4, bool active) {
  const float ga = fmaf(msum[bt.x].x, bits_to_float(bt.y), msum[bt.x + 1].x * bits_to_float(bt.z));
  const float px = fmaf(x2.x, x2.x, x2.y * x2.y);                          // 4 |X|^2
  // w gA / Ax * X = w gA * 2 rsqrt(px) * (2X) / 2
  const float g = (active && px >= eps4) ? (self ? 1.f : 0.5f) * ga * spl_fast_rsqrt(px) : 0.f;
  return __fmul2_rn(x2, make_float2(g, g));
}

// mel, pass 2: balanced banded projection of the amplitude pairs parked in the frame slot.  Every mel row is summed
// by a group of 1..L lanes walking a host-built table of (amplitude slot, weight) entries, then reduced with
// shuffles; msum[row] receives the pair of mel energies.  Ends with the warp barrier that publishes msum.
template <int L>
SPL_DEVICE void mel_project_pairs(const float2* S, float2* msum, int l, int mel_rounds, const int4* mel_tasks,
                                  const int2* mel_entries) {
  for (int r = 0; r < mel_rounds; ++r) {
    const int4 tk = mel_tasks[r * L + l];
    const int row = tk.x & 0xfff, grp = (tk.x >> 12) & 0xff, iters = tk.x >> 20;
    const int2* en = mel_entries + tk.y * L + l;
    float mx = 0.f, my = 0.f;
#pragma unroll 4
    for (int s = 0; s < iters; ++s) {
      const int2 e = en[s * L];
      const float2 amp = S[e.x];
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
}

// mel epilogue of one frame: amplitudes -> banded projection -> log-mel L1 -> (GRAD) gradient spectrum.
// In: Z columns < L/2 in A (PARK: in the slot), mirror halves in the slot.  Out (GRAD): H in the same places.
// The amplitudes of bin k <= N/2 are parked at the bin's own slot position; 2X[k] waits for pass 3 in A (PARK: in
// the mirror position, whose Z[N-k] is consumed by then).
template <int NFFT, bool GRAD>
SPL_DEVICE void mel_epilogue(float2 (&A)[Geo<NFFT>::AROWS][Geo<NFFT>::HL], float2* S, float2* msum, int l, bool eq,
                             bool active, const TransformParams& p, const int4* mel_tasks, const int2* mel_entries,
                             const int4* bin_tab, float inv_s, float& s1) {
  using G = Geo<NFFT>;
  constexpr int L = G::L, R = G::R;
  const float eps4 = 4.f * p.eps;
  float2 xh = make_float2(0.f, 0.f);                   // 2X[N/2], lane 0
  // pass 1
#pragma unroll(G::PARK ? 1 : G::RPL)
  for (int j = 0; j < G::RPL; ++j) {
    const int row = l + L * j;
    float2* pb = mirror_ptr<NFFT>(S, row);
    float2* arow = S + row * G::PITCH;
#pragma unroll
    for (int k1 = 0; k1 < G::HL; ++k1) {
      const float2 a = G::PARK ? arow[k1] : A[G::PARK ? 0 : j][k1];
      float2 bm = pb[-k1];
      if (k1 == 0) bm = (row == 0) ? a : bm;
      float2 x2, amp;
      mel_pair_amp(a, bm, eq, eps4, x2, amp, inv_s);
      if (G::PARK) pb[-k1] = x2;
      else         A[G::PARK ? 0 : j][k1] = x2;
      arow[k1] = amp;
    }
  }
  if (l == 0) {
    const float2 a = S[G::HL];
    float2 amp;
    mel_pair_amp(a, a, eq, eps4, xh, amp, inv_s);
    S[G::HL] = amp;
  }
  __syncwarp();
  mel_project_pairs<L>(S, msum, l, p.mel_rounds, mel_tasks, mel_entries);
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
    // pass 3: gA[k] = sum_m gM[m] W[k,m] (<= 2 adjacent rows), H[k] = w gA gate / Ax * X, H[N-k] = conj H[k]
#pragma unroll(G::PARK ? 1 : G::RPL)
    for (int j = 0; j < G::RPL; ++j) {
      const int row = l + L * j;
      float2* pb = mirror_ptr<NFFT>(S, row);
      float2* arow = S + row * G::PITCH;
      const int4* bt = bin_tab + row;
#pragma unroll
      for (int k1 = 0; k1 < G::HL; ++k1) {
        const bool self = k1 == 0 && row == 0;
        const float2 x2 = G::PARK ? pb[-k1] : A[G::PARK ? 0 : j][k1];
        const float2 hk = mel_bin_grad(x2, self, bt[R * k1], msum, eps4, active);
        if (G::PARK) arow[k1] = hk;
        else         A[G::PARK ? 0 : j][k1] = hk;
        pb[-k1] = make_float2(hk.x, -hk.y);
      }
    }
    if (l == 0) S[G::HL] = mel_bin_grad(xh, true, bin_tab[NFFT / 2], msum, eps4, active);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// The transform kernel body.  A group of L lanes takes one frame at a time (warps stride over the frames of the
// whole batch); everything between the waveform loads and the per-frame gradient store stays on chip.
// WIN_T > 0: window length known at compile time (shipped configs), which prunes the zero taps out of the
// load, the first butterflies of the forward transform and the last ones of the inverse.
// ---------------------------------------------------------------------------------------------
template <int NFFT, int KIND, bool GRAD, int WIN_T, bool RING = false>
SPL_DEVICE void transform_body(const TransformParams& p, float* smem, int block, int tid, int grid, int wpc) {
  using G = Geo<NFFT>;
  using SL = SmemLayout<NFFT, KIND>;
  constexpr int L = G::L, R = G::R, FPW = G::FPW, HALF = NFFT / 2;
  const int warp = tid >> 5, lane = tid & 31;
  const int l = lane & (L - 1), h = lane / L;
  const int win = WIN_T > 0 ? WIN_T : p.win;
  const int left = WIN_T > 0 ? (NFFT - WIN_T) / 2 : p.left;

  const CtaTables ct = cta_tables(NFFT, p.win, KIND, p.mel_rounds, p.mel_entry_rows);
  const float2* tw = reinterpret_cast<const float2*>(smem + ct.tw);
  const float* wtab = smem + ct.win;
  const int4* mel_tasks = reinterpret_cast<const int4*>(smem + ct.tasks);
  const int2* mel_entries = reinterpret_cast<const int2*>(smem + ct.entries);
  const int4* bin_tab = reinterpret_cast<const int4*>(smem + ct.bintab);

  // RING (a separate instantiation, so that the default kernels carry none of its code): runs of p.run_frames frames
  const int m = (RING && GRAD && KIND == kKindStft) ? p.run_frames : 1;       // frames per run (1: one gradient slot per frame)
  const bool ring_on = RING && m > 1;
  const int ring_taps = ring_on ? win : 0;
  float* wsm = smem + ct.total + (size_t)warp * SL::words_per_warp(p.n_mels, ring_taps);
  float2* S = reinterpret_cast<float2*>(wsm) + h * G::SLOT_F2;                       // this group's frame slot
  float2* msum = reinterpret_cast<float2*>(wsm + FPW * G::SLOT_F2 * 2) + h * align4(p.n_mels);   // mel: (Mx, My), then (gM, -)
  float2* ring = reinterpret_cast<float2*>(wsm + SL::words_per_warp(p.n_mels, 0)) + h * ((ring_taps + 1) & ~1);

  const int per_utt = ring_on ? p.runs_per_utt : p.n_frames;     // work items per utterance: runs or frames
  const int total = p.B * per_utt;
  const unsigned grp_mask = (L == 32) ? 0xffffffffu : (((1u << (L & 31)) - 1u) << (h * L));
  double d1 = 0.0, d2 = 0.0, d3 = 0.0;      // this lane's share of the sums (stft: 4*S1, 4*S2, S3/(0.5 ln 2); mel: S4)
  if (ring_on) {
    for (int i = l; i < win; i += L) ring[i] = make_float2(0.f, 0.f);      // every flush leaves its taps at zero again
    __syncwarp();
  }

  // [region: frame loop]
  constexpr bool CTA_SYNC = SPL_SYNC_2048 && NFFT == 2048;
  const int stride = grid * wpc * FPW;
  const int rounds = (total + stride - 1) / stride;
  for (int round = 0; round < rounds; ++round) {
    const int base = round * stride + (block * wpc + warp) * FPW;
    if (CTA_SYNC) __syncthreads();
    if (base >= total) continue;
    const int item = base + h;
    const bool item_ok = item < total;
    const int b = item_ok ? item / per_utt : 0;
    const int t0 = item_ok ? (item - b * per_utt) * m : 0;
    int start = 0;                 // ring position of tap 0 of the current frame
    int done = 0;                  // frames of this run added to the ring so far
    for (int j = 0; j < m; ++j) {
    const int t = t0 + j;
    const bool active = item_ok && t < p.n_frames;
    if (FPW == 1 && !active) break;                    // warp-uniform: the last run of an utterance is shorter
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    // Two phases share ONE copy of the R-point codelet: 0 = forward transform + epilogue (+ first half of the
    // inverse), 1 = second half of the inverse + windowed store.
#pragma unroll(kPhaseUnroll)
    for (int phase = 0; phase < (GRAD ? 2 : 1); ++phase) {
      float2 v[R];              // defined afresh by each phase: nothing is carried around the loop in registers
      bool frame_equal = false;
      float inv_s = 1.f;
      if (phase == 0) {
        float2 amax = make_float2(0.f, 0.f);
        const bool same = load_taps<NFFT, WIN_T>(v, p.x + (size_t)b * p.T, p.y + (size_t)b * p.T, p.T,
                                                 t * p.hop - HALF, win, left, wtab, l, active, amax);
        inv_s = equalise_pair<L, R>(v, amax, grp_mask);
        // A frame whose prediction and target taps are bit-identical must contribute exactly zero (the reference
        // returns sc = mag = mel = 0 and a zero gradient for x == y); the packed FFT would leave ~1e-7 of rounding
        // asymmetry between X and Y, so such frames reuse X for Y.
        const unsigned eq_bits = __ballot_sync(0xffffffffu, same);
        frame_equal = (eq_bits & grp_mask) == grp_mask;
      } else {
        // [region: inv column load]
#pragma unroll
        for (int m2 = 0; m2 < R; ++m2) v[m2] = S[m2 * G::PITCH + l];
      }
      Dft<R>::run(v);
      if (phase == 0) {
        fwd_store_cols<NFFT>(v, S, tw, l);
        float2 A[G::AROWS][G::HL];
        // PARK: the row passes below are rolled loops; fresh opaque copies of the lane index and the slot pointer
        // keep their address arithmetic inside this frame instead of live across the tap loads
        const int lr = G::PARK ? launder(l) : l;
        float2* Sr = G::PARK ? launder_ptr(S) : S;
        fwd_pass_b<NFFT>(A, Sr, lr);
        if (KIND == kKindStft) {
          stft_epilogue<NFFT, GRAD>(A, Sr, lr, frame_equal, 4.f * p.eps, inv_s, s1, s2, s3);
          __syncwarp();    // mirror halves: all reads (and the H written back over them) done before the slot is reused
        } else {
          mel_epilogue<NFFT, GRAD>(A, Sr, msum, lr, frame_equal, active, p, mel_tasks, mel_entries, bin_tab, inv_s, s1);
        }
        if (active) { d1 += (double)s1; d2 += (double)s2; d3 += (double)s3; }
        if (GRAD) inv_pass_b<NFFT>(A, Sr, tw, lr);
      } else if (!ring_on) {
        if (active) {
        // [region: window + store]
        // v[n2] = sample n = l + L*n2 of the two real gradient sequences with swapped components: (.y, .x) = (u, v).
        // Windowed frame gradient -> this frame's slot in HBM (coalesced; one writer per element).
        const int fidx = b * p.n_frames + t;
        float2* out2 = reinterpret_cast<float2*>(p.gframes) + (size_t)fidx * win;
        float* out1 = reinterpret_cast<float*>(p.gframes) + (size_t)fidx * win;
        const float* wp = wtab + (l - left);
#pragma unroll
        for (int n2 = 0; n2 < R; ++n2) {
          const int lo = L * n2 - left;
          if (WIN_T > 0 && (lo + L - 1 < 0 || lo >= WIN_T)) continue;
          const bool all_lanes = WIN_T > 0 && lo >= 0 && lo + L - 1 < WIN_T;
          if (all_lanes || (lo + l >= 0 && lo + l < win)) {
            const float w = wp[L * n2];
            if (KIND == kKindStft) out2[lo + l] = __fmul2_rn(make_float2(v[n2].y, v[n2].x), make_float2(w, w));
            else                   out1[lo + l] = v[n2].y * w;
          }
        }
        }
      } else {
        // [region: window + ring]
        // Overlap-add into the ring: tap n of frame j of the run sits at ring position (j * hop + n) mod win.
        if (active) {
          const float* wp = wtab + (l - left);
#pragma unroll
          for (int n2 = 0; n2 < R; ++n2) {
            const int lo = L * n2 - left;
            if (WIN_T > 0 && (lo + L - 1 < 0 || lo >= WIN_T)) continue;
            const bool all_lanes = WIN_T > 0 && lo >= 0 && lo + L - 1 < WIN_T;
            if (all_lanes || (lo + l >= 0 && lo + l < win)) {
              const float w = wp[L * n2];
              int idx = start + lo + l;
              idx = idx >= win ? idx - win : idx;
              ring[idx] = __ffma2_rn(make_float2(v[n2].y, v[n2].x), make_float2(w, w), ring[idx]);
            }
          }
        }
        __syncwarp();
        // the first `hop` taps of this frame are final: no later frame of the run reaches them
        if (active) {
          float2* out2 = reinterpret_cast<float2*>(p.gframes) + (size_t)item * p.run_len + (size_t)done * p.hop;
          for (int i = l; i < p.hop; i += L) {
            int idx = start + i;
            idx = idx >= win ? idx - win : idx;
            out2[i] = ring[idx];
            ring[idx] = make_float2(0.f, 0.f);
          }
          start += p.hop;
          start = start >= win ? start - win : start;
          done += 1;
        }
      }
    }
    }
    if (ring_on) {
      // end of the run: the remaining win - hop taps of its last frame, then zeros up to run_len (a short last run)
      __syncwarp();
      if (item_ok && done > 0) {
        float2* out2 = reinterpret_cast<float2*>(p.gframes) + (size_t)item * p.run_len + (size_t)done * p.hop;
        const int rest = win - p.hop;
        for (int i = l; i < rest; i += L) {
          int idx = start + i;
          idx = idx >= win ? idx - win : idx;
          out2[i] = ring[idx];
          ring[idx] = make_float2(0.f, 0.f);
        }
        for (int i = done * p.hop + rest + l; i < p.run_len; i += L) out2[i - done * p.hop] = make_float2(0.f, 0.f);
      }
      __syncwarp();
    }
  }

  // [region: partial sums]
  // one row of partial sums per warp (fixed frame -> warp assignment and fixed order: deterministic)
  d1 = warp_sum(d1);
  if (KIND == kKindStft) { d2 = warp_sum(d2); d3 = warp_sum(d3); }
  if (lane == 0) {
    const int prow = block * wpc + warp;
    if (KIND == kKindStft) {
      double* o = p.partials + (size_t)prow * 3;
      o[0] = 0.25 * d1; o[1] = 0.25 * d2; o[2] = 0.34657359027997264 * d3;     // 0.5 ln 2
    } else {
      p.partials[prow] = d1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Explicit magnitude spectrogram, forward only: out[b, t, k] = sqrt(max(|STFT(x)[b, t, k]|^2, eps)),
// laid out (B, F, ld >= K) -- the tensor stft() returns (losses/stft_loss.py:19-35) and the A operand of
// the mel projection GEMM (mel_loss.py:88-91).  Frames t and t+1 of the SAME signal share one complex FFT (real /
// imaginary part), so the work per frame is half of a naive real transform.
// ---------------------------------------------------------------------------------------------
struct SpecParams {
  const float* x;        // (B, T)
  int B, T;
  int hop, win, left, n_frames;
  int n_pairs;           // frame pairs per utterance = (n_frames + 1) / 2
  float eps;
  const float* window;
  const float2* twiddle;
  float* out;            // (B, n_frames, ld); pad columns [n_fft/2+1, ld) are written as zeros
  float* out_lo;         // null, or the second half of the TF32 split: out = tf32(A) exactly, out_lo = A - out
  int ld;
};

// A = hi + lo with hi exactly representable in TF32 (low 13 mantissa bits cleared); lo = A - hi is exact in fp32
SPL_DEVICE void spec_store(float* o, float* o_lo, int k, float a) {
  if (o_lo) {
    int bits;
    memcpy(&bits, &a, 4);
    bits &= ~0x1fff;
    float hi;
    memcpy(&hi, &bits, 4);
    o[k] = hi;
    o_lo[k] = a - hi;
  } else {
    o[k] = a;
  }
}

// sqrt(p4 / 4) for p4 = max(4 |X|^2, 4 eps); eps = 0 (plain |X|, torchaudio power=1) must give exactly 0 at p4 = 0, not
// 0 * rsqrt(0) = NaN
SPL_DEVICE float half_sqrt(float p4) { return p4 >= 1.17549435e-38f ? 0.5f * p4 * spl_fast_rsqrt(p4) : 0.f; }

// [region: spectrogram]
template <int NFFT>
SPL_DEVICE void spec_body(const SpecParams& p, float* smem, int block, int tid, int grid, int wpc) {
  using G = Geo<NFFT>;
  using SL = SmemLayout<NFFT, kKindStft>;
  constexpr int L = G::L, R = G::R, FPW = G::FPW, HALF = NFFT / 2;
  const int warp = tid >> 5, lane = tid & 31;
  const int l = lane & (L - 1), h = lane / L;
  const CtaTables ct = cta_tables(NFFT, p.win, kKindStft, 0, 0);
  const float2* tw = reinterpret_cast<const float2*>(smem + ct.tw);
  const float* wtab = smem + ct.win;
  float2* S = reinterpret_cast<float2*>(smem + ct.total + (size_t)warp * SL::words_per_warp(0)) + h * G::SLOT_F2;
  const int total = p.B * p.n_pairs;                 // work items: (utterance, frame pair)
  const float eps4 = 4.f * p.eps;
  for (int base = (block * wpc + warp) * FPW; base < total; base += grid * wpc * FPW) {
    const int item = base + h;
    const bool active = item < total;
    const int b = active ? item / p.n_pairs : 0, pr = active ? item - b * p.n_pairs : 0;
    const int t = 2 * pr;
    const bool second = active && (t + 1 < p.n_frames);
    const float* __restrict__ xb = p.x + (size_t)b * p.T;
    float2 A[G::AROWS][G::HL];
    {
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
      Dft<R>::run(v);
      fwd_store_cols<NFFT>(v, S, tw, l);
    }
    fwd_pass_b<NFFT>(A, S, l);
    const size_t orow = ((size_t)b * p.n_frames + t) * p.ld;
    float* o0 = p.out + orow;
    float* o1 = o0 + p.ld;
    float* l0 = p.out_lo ? p.out_lo + orow : nullptr;
    float* l1 = p.out_lo ? l0 + p.ld : nullptr;
#pragma unroll(G::PARK ? 1 : G::RPL)
    for (int j = 0; j < G::RPL; ++j) {
      const int row = l + L * j;
      const float2* pb = mirror_ptr<NFFT>(S, row);
      const float2* arow = S + row * G::PITCH;
#pragma unroll
      for (int k1 = 0; k1 < G::HL; ++k1) {
        const float2 a = G::PARK ? arow[k1] : A[G::PARK ? 0 : j][k1];
        float2 bm = pb[-k1];
        if (k1 == 0) bm = (row == 0) ? a : bm;
        const float2 x2 = __fadd2_rn(a, make_float2(bm.x, -bm.y));                           // 2 X_t[k]
        const float2 y2 = __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));       // 2 X_{t+1}[k]
        const float p0 = fmaxf(fmaf(x2.x, x2.x, x2.y * x2.y), eps4);
        const float p1 = fmaxf(fmaf(y2.x, y2.x, y2.y * y2.y), eps4);
        const int k = row + R * k1;
        if (active) spec_store(o0, l0, k, half_sqrt(p0));
        if (second) spec_store(o1, l1, k, half_sqrt(p1));
      }
    }
    for (int k = NFFT / 2 + 1 + l; k < p.ld; k += L) {               // pad columns of the GEMM operand
      if (active) { o0[k] = 0.f; if (l0) l0[k] = 0.f; }
      if (second) { o1[k] = 0.f; if (l1) l1[k] = 0.f; }
    }
    if (l == 0) {                                                    // bin N/2
      const float2 a = S[G::HL];
      const float p0 = fmaxf(4.f * a.x * a.x, eps4), p1 = fmaxf(4.f * a.y * a.y, eps4);
      if (active) spec_store(o0, l0, NFFT / 2, half_sqrt(p0));
      if (second) spec_store(o1, l1, NFFT / 2, half_sqrt(p1));
    }
    __syncwarp();       // the slot's mirror halves are read above; the next pass-A store must wait for every lane
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of the explicit spectrograms: what autograd derives for stft() (losses/stft_loss.py:19-35) and for
// MelSpectrogram.forward (losses/mel_loss.py:74-94) given the upstream gradient of their OUTPUT tensor
// (SURVEY appendix A.2 steps 3-6).  The spectra are recomputed (frames t and t+1 of the same signal share one
// complex FFT, as in spec_body), multiplied by the upstream gradient and sent back through the adjoint transform;
// one complex inverse FFT returns the two real frame gradients.  The windowed frame gradients go to one slot per
// frame (float [B * F][win]); combine_kernel (unit coefficients) overlap-adds and folds them into dx.
//   kStft: g = dL/dA          (B, F, ld)       gX = g [|X|^2 >= eps] X / |X|
//   kMel : g = dL/d log-mel   (B, n_mels, F)   gM = g [M >= eps] / (M ln b),  gA = gM W^T (banded),  gX as above
// ---------------------------------------------------------------------------------------------
struct SpecGradParams {
  TransformParams t;     // x, B, T, hop, win, left, n_frames, eps, window, twiddle, gframes, mel tables (y, partials unused)
  int n_pairs;           // frame pairs per utterance = (n_frames + 1) / 2
  const float* g;        // upstream gradient
  int ld;                // kStft: floats per frame row of g
};

// One mirror pair (k, N-k) of the packed spectrum Z = FFT(f_t + i f_{t+1}):
//   H[k] = w (c0 2X_t[k] + i c1 2X_{t+1}[k]),  H[N-k] = w (c0 conj(2X_t[k]) + i c1 conj(2X_{t+1}[k])),
//   c = gA gate / sqrt(4 |X|^2),  w = 1/2 (1 when k mirrors itself).
SPL_DEVICE void specgrad_pair(float2 a, float2 bm, bool self, float eps4, float g0, float g1, float2& ha, float2& hb) {
  const float2 x2 = __fadd2_rn(a, make_float2(bm.x, -bm.y));
  const float2 y2 = __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));
  const float p0 = fmaf(x2.x, x2.x, x2.y * x2.y);
  const float p1 = fmaf(y2.x, y2.x, y2.y * y2.y);
  const float w = self ? 1.f : 0.5f;
  const float c0 = p0 >= eps4 ? w * g0 * spl_fast_rsqrt(p0) : 0.f;      // clamp gate of the reference
  const float c1 = p1 >= eps4 ? w * g1 * spl_fast_rsqrt(p1) : 0.f;
  const float2 u = __fmul2_rn(x2, make_float2(c0, c0));
  const float2 v = __fmul2_rn(y2, make_float2(c1, c1));
  ha = make_float2(u.x - v.y, u.y + v.x);
  hb = make_float2(u.x + v.y, v.x - u.y);
}

// [region: spectrogram backward]
template <int NFFT, int KIND>
SPL_DEVICE void specgrad_body(const SpecGradParams& q, float* smem, int block, int tid, int grid, int wpc) {
  using G = Geo<NFFT>;
  using SL = SmemLayout<NFFT, KIND>;
  static_assert(!G::PARK, "the spectrogram backward keeps its rows in registers");
  constexpr int L = G::L, R = G::R, FPW = G::FPW, HALF = NFFT / 2;
  const TransformParams& p = q.t;
  const int warp = tid >> 5, lane = tid & 31;
  const int l = lane & (L - 1), h = lane / L;
  const CtaTables ct = cta_tables(NFFT, p.win, KIND, p.mel_rounds, p.mel_entry_rows);
  const float2* tw = reinterpret_cast<const float2*>(smem + ct.tw);
  const float* wtab = smem + ct.win;
  const int4* mel_tasks = reinterpret_cast<const int4*>(smem + ct.tasks);
  const int2* mel_entries = reinterpret_cast<const int2*>(smem + ct.entries);
  const int4* bin_tab = reinterpret_cast<const int4*>(smem + ct.bintab);
  float* wsm = smem + ct.total + (size_t)warp * SL::words_per_warp(p.n_mels);
  float2* S = reinterpret_cast<float2*>(wsm) + h * G::SLOT_F2;
  float2* msum = reinterpret_cast<float2*>(wsm + FPW * G::SLOT_F2 * 2) + h * align4(p.n_mels);
  const int total = p.B * q.n_pairs;                 // work items: (utterance, frame pair)
  // clamp gate; never below the smallest normal, so that eps = 0 (plain |X|) gets d|X|/dX := 0 at X = 0 like torch.abs
  const float eps4 = fmaxf(4.f * p.eps, 1.17549435e-38f);
  for (int base = (block * wpc + warp) * FPW; base < total; base += grid * wpc * FPW) {
    const int item = base + h;
    const bool active = item < total;
    const int b = active ? item / q.n_pairs : 0, pr = active ? item - b * q.n_pairs : 0;
    const int t = 2 * pr;
    const bool second = active && (t + 1 < p.n_frames);
    const float* __restrict__ xb = p.x + (size_t)b * p.T;
    float2 A[G::AROWS][G::HL];
    {
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
      Dft<R>::run(v);
      fwd_store_cols<NFFT>(v, S, tw, l);
    }
    fwd_pass_b<NFFT>(A, S, l);
    if (KIND == kKindStft) {
      const float* __restrict__ g0 = q.g + ((size_t)b * p.n_frames + t) * q.ld;
      const float* __restrict__ g1 = g0 + q.ld;
#pragma unroll
      for (int j = 0; j < G::RPL; ++j) {
        const int row = l + L * j;
        float2* pb = mirror_ptr<NFFT>(S, row);
#pragma unroll
        for (int k1 = 0; k1 < G::HL; ++k1) {
          const float2 a = A[j][k1];
          const bool self = k1 == 0 && row == 0;
          float2 bm = pb[-k1];
          if (k1 == 0) bm = self ? a : bm;
          const int k = row + R * k1;
          float2 ha, hb;
          specgrad_pair(a, bm, self, eps4, active ? __ldg(g0 + k) : 0.f, second ? __ldg(g1 + k) : 0.f, ha, hb);
          A[j][k1] = ha;
          pb[-k1] = hb;
        }
      }
      if (l == 0) {                                                    // bin N/2
        const float2 a = S[G::HL];
        float2 ha, hb;
        specgrad_pair(a, a, true, eps4, active ? __ldg(g0 + HALF) : 0.f, second ? __ldg(g1 + HALF) : 0.f, ha, hb);
        S[G::HL] = ha;
      }
    } else {
      // pass 1: amplitude pairs (A_t, A_{t+1}) parked at the bins' own slot positions; Z stays where it is
      float2 ah = make_float2(0.f, 0.f);                               // Z[N/2], lane 0
#pragma unroll
      for (int j = 0; j < G::RPL; ++j) {
        const int row = l + L * j;
        const float2* pb = mirror_ptr<NFFT>(S, row);
        float2* arow = S + row * G::PITCH;
#pragma unroll
        for (int k1 = 0; k1 < G::HL; ++k1) {
          const float2 a = A[j][k1];
          float2 bm = pb[-k1];
          if (k1 == 0) bm = (row == 0) ? a : bm;
          float2 x2, amp;
          mel_pair_amp(a, bm, false, eps4, x2, amp);
          arow[k1] = amp;
        }
      }
      if (l == 0) {
        ah = S[G::HL];
        float2 x2, amp;
        mel_pair_amp(ah, ah, false, eps4, x2, amp);
        S[G::HL] = amp;
      }
      __syncwarp();
      mel_project_pairs<L>(S, msum, l, p.mel_rounds, mel_tasks, mel_entries);
      // gM = g [M >= eps] / (M ln b) for both frames
      for (int row = l; row < p.n_mels; row += L) {
        const float2 mm = msum[row];
        const float* gp = q.g + ((size_t)b * p.n_mels + row) * p.n_frames + t;
        const float u0 = active ? __ldg(gp) : 0.f, u1 = second ? __ldg(gp + 1) : 0.f;
        msum[row] = make_float2(mm.x >= p.eps ? u0 * p.inv_ln_base / mm.x : 0.f,
                                mm.y >= p.eps ? u1 * p.inv_ln_base / mm.y : 0.f);
      }
      __syncwarp();
      // pass 3: gA[k] = sum_m gM[m] W[k, m] (<= 2 adjacent rows) for both frames -> H
#pragma unroll
      for (int j = 0; j < G::RPL; ++j) {
        const int row = l + L * j;
        float2* pb = mirror_ptr<NFFT>(S, row);
        const int4* btr = bin_tab + row;
#pragma unroll
        for (int k1 = 0; k1 < G::HL; ++k1) {
          const float2 a = A[j][k1];
          const bool self = k1 == 0 && row == 0;
          float2 bm = pb[-k1];
          if (k1 == 0) bm = self ? a : bm;
          const int4 bt = btr[R * k1];
          const float2 m0 = msum[bt.x], m1 = msum[bt.x + 1];
          const float w0 = bits_to_float(bt.y), w1 = bits_to_float(bt.z);
          float2 ha, hb;
          specgrad_pair(a, bm, self, eps4, fmaf(m0.x, w0, m1.x * w1), fmaf(m0.y, w0, m1.y * w1), ha, hb);
          A[j][k1] = ha;
          pb[-k1] = hb;
        }
      }
      if (l == 0) {
        const int4 bt = bin_tab[HALF];
        const float2 m0 = msum[bt.x], m1 = msum[bt.x + 1];
        const float w0 = bits_to_float(bt.y), w1 = bits_to_float(bt.z);
        float2 ha, hb;
        specgrad_pair(ah, ah, true, eps4, fmaf(m0.x, w0, m1.x * w1), fmaf(m0.y, w0, m1.y * w1), ha, hb);
        S[G::HL] = ha;
      }
    }
    __syncwarp();          // H complete in the slot (mirror halves, bin N/2) before the row passes read it
    inv_pass_b<NFFT>(A, S, tw, l);
    {
      float2 v[R];
#pragma unroll
      for (int m2 = 0; m2 < R; ++m2) v[m2] = S[m2 * G::PITCH + l];
      Dft<R>::run(v);
      // v[n2] = sample n = l + L*n2 with swapped components: .y = frame t, .x = frame t+1
      float* o0 = p.gframes ? reinterpret_cast<float*>(p.gframes) + ((size_t)b * p.n_frames + t) * p.win : nullptr;
      float* o1 = o0 + p.win;
#pragma unroll
      for (int n2 = 0; n2 < R; ++n2) {
        const int tap = L * n2 - p.left + l;
        if (tap >= 0 && tap < p.win) {
          const float w = wtab[tap];
          if (active) o0[tap] = v[n2].y * w;
          if (second) o1[tap] = v[n2].x * w;
        }
      }
    }
    __syncwarp();          // column reads done before the next item's pass-A store
  }
}

// ---------------------------------------------------------------------------------------------
// Waveform shape loss (losses/waveform_loss.py:15-75, the third switch of TrainerGAN._metric_loss,
// trainer/trainerGAN.py:235-239): mean over window lengths of L1(maxpool_w |y_hat|, maxpool_w |y|), MaxPool1d(w) =
// disjoint windows [j w, (j+1) w), j < T / w.  HBM-bound: the forward reads both signals ONCE (a warp walks a span of
// the row through every window length while the span is in L1) and leaves, per window, a 4-byte record
// {argmax of |y_hat| (first maximum, as max_pool1d) << 2 | sign(pool(y_hat) - pool(y)) sign(y_hat[argmax]) + 1};
// the backward writes dx once from the records.  Algorithmic bytes: 8 per sample forward + 4 backward.
// ---------------------------------------------------------------------------------------------
constexpr int kShapeMaxWin = 8;
constexpr int kShapeSpan = 2048;          // samples per warp work item
constexpr int kShapeMaxBlocks = 64;       // one-pass forward: blocks per span (block >= 32 samples)

struct ShapeParams {
  const float* x;        // prediction (rows, T)
  const float* y;        // target     (rows, T)
  int rows, T;
  int n;                 // window lengths
  int span;              // forward: samples per warp work item (the host shrinks it until every SM has work)
  int block;             // forward: common divisor of the window lengths for the one-pass path, 0 = pass per length
  int win[kShapeMaxWin];
  long long rec_ofs[kShapeMaxWin];   // first record of window length r; records of r are [rows][T / win[r]]
  int* records;
  double* partials;      // forward: [grid * warps per CTA][n], one row per warp
  // backward
  float coef[kShapeMaxWin];          // 1 / (n * rows_global * (T / win[r]))
  const float* g;        // upstream gradient (device scalar)
  float* dx;             // (rows, T)
};

// windows of length w that START inside [s0, s1): j in [ceil(s0 / w), ceil(s1 / w)), clipped to T / w
SPL_DEVICE void shape_window_range(int s0, int s1, int w, int T, int& j0, int& j1) {
  j0 = (s0 + w - 1) / w;
  j1 = min((s1 + w - 1) / w, T / w);
}

// [region: shape forward]
// Per-window reduction of one warp, lanes across the samples [base, base + w): three redux.sync instructions -- the
// maximum of |x| as an unsigned (non-negative floats order like their bit patterns), the smallest index that attains
// it (max_pool1d routes the gradient to the first maximum) and the maximum of |y|.  Returns (max |x|, max |y|) to
// every lane; the lane owning the argmax gets own = true with its sample in sx and its index in ix.
template <int NS>
SPL_DEVICE void shape_load(const float* __restrict__ xr, const float* __restrict__ yr, int base, int w, int lane,
                           float (&xv)[NS], float (&yv)[NS]) {
#pragma unroll
  for (int u = 0; u < NS; ++u) {
    xv[u] = 0.f;
    yv[u] = 0.f;
    if (32 * u < w) {                                      // warp-uniform: unused slots cost one branch
      const int i = lane + 32 * u;
      if (i < w) { xv[u] = __ldg(xr + base + i); yv[u] = __ldg(yr + base + i); }
    }
  }
}

// running (max |x|, first index, sample there) and max |y| of this lane; mx starts at -1 so that the first sample
// always takes over
template <int NS>
SPL_DEVICE void shape_fold(const float (&xv)[NS], const float (&yv)[NS], int base, int w, int lane, float& mx, float& my,
                           float& sx, int& ix) {
#pragma unroll
  for (int u = 0; u < NS; ++u) {                           // ascending i: strict > keeps the first maximum per lane
    if (32 * u < w) {
      const int i = lane + 32 * u;
      const float ax = fabsf(xv[u]);
      if (i < w && ax > mx) { mx = ax; ix = base + i; sx = xv[u]; }
      my = fmaxf(my, fabsf(yv[u]));
    }
  }
}

SPL_DEVICE void shape_finish(float mx, float my, int ix, float& mxo, float& myo, bool& own) {
  const unsigned mine = float_bits(fmaxf(mx, 0.f));
  const unsigned mxu = __reduce_max_sync(0xffffffffu, mine);
  const unsigned myu = __reduce_max_sync(0xffffffffu, float_bits(my));
  const int first = (int)__reduce_min_sync(0xffffffffu, (ix != 0x7fffffff && mine == mxu) ? (unsigned)ix : 0x7fffffffu);
  own = ix == first;
  mxo = bits_to_float((int)mxu);
  myo = bits_to_float((int)myu);
}

// any window length: chunks of 256 samples, the 16 loads of a lane issued before the first compare
SPL_DEVICE void shape_scan(const float* __restrict__ xr, const float* __restrict__ yr, int base, int w, int lane,
                           float& mxo, float& myo, float& sx, int& ix, bool& own) {
  float mx = -1.f, my = 0.f;
  sx = 0.f;
  ix = 0x7fffffff;
  for (int c = 0; c < w; c += 256) {
    float xv[8], yv[8];
    shape_load<8>(xr, yr, base + c, w - c, lane, xv, yv);
    shape_fold<8>(xv, yv, base + c, w - c, lane, mx, my, sx, ix);
  }
  shape_finish(mx, my, ix, mxo, myo, own);
}

SPL_DEVICE int shape_record(int ix, float d, float sx) {
  const int sd = (d > 0.f) - (d < 0.f), ss = (sx > 0.f) - (sx < 0.f);
  return (ix << 2) | (sd * ss + 1);
}

// one-pass path, second half: every window of every length that starts in the span [s0, s1) is combined from its
// blocks (blk[b] = {max |x|, max |y|, x at argmax, bits(argmax)}), one window per lane
SPL_DEVICE void shape_window_pass(const ShapeParams& p, const float4* blk, int row, int s0, int s1, int lane,
                                  double (&acc)[kShapeMaxWin]) {
  const int g = p.block;
#pragma unroll
  for (int r = 0; r < kShapeMaxWin; ++r) {
    if (r >= p.n) continue;                                // fully unrolled: acc[] stays in registers
    const int w = p.win[r], nw = p.T / w, m = w / g;
    const int j0 = s0 / w, j1 = min((s1 + w - 1) / w, nw);            // s0 is a multiple of w
    int* rec = p.records + p.rec_ofs[r] + (long long)row * nw;
    float sum = 0.f;
    for (int j = j0 + lane; j < j1; j += 32) {
      const int b0 = (j * w - s0) / g;
      float4 best = blk[b0];
      float my = best.y;
      for (int q = 1; q < m; ++q) {                        // ascending blocks: strict > keeps the first maximum
        const float4 c = blk[b0 + q];
        if (c.x > best.x) best = c;
        my = fmaxf(my, c.y);
      }
      const float d = best.x - my;
      sum += fabsf(d);
      rec[j] = shape_record((int)float_bits(best.w), d, best.z);
    }
    acc[r] += (double)sum;
  }
}

// one row of partial sums per warp (fixed span -> warp -> lane assignment: deterministic)
SPL_DEVICE void shape_write_partials(const ShapeParams& p, double (&acc)[kShapeMaxWin], int prow, int lane) {
#pragma unroll
  for (int r = 0; r < kShapeMaxWin; ++r)
    if (r < p.n) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    double* o = p.partials + (size_t)prow * p.n;
#pragma unroll
    for (int r = 0; r < kShapeMaxWin; ++r)
      if (r < p.n) o[r] = acc[r];
  }
}

// [region: shape forward vec]
// One-pass path when everything is a multiple of 4 samples (T, block): 16-byte loads, NV float4 per lane per block and
// signal, the loads of block b+1 in flight while block b is reduced -- the kernel needs ~45 KB of loads in flight per
// SM to keep HBM busy, which scalar loads consumed block by block do not provide.
template <int NV>
SPL_DEVICE void shape_load4(const float* __restrict__ xr, const float* __restrict__ yr, int base, int g4, int lane,
                            float4 (&xv)[NV], float4 (&yv)[NV]) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int q = lane + 32 * v;
    xv[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    yv[v] = xv[v];
    if (q < g4) {
      xv[v] = __ldg(reinterpret_cast<const float4*>(xr + base) + q);
      yv[v] = __ldg(reinterpret_cast<const float4*>(yr + base) + q);
    }
  }
}

template <int NV>
SPL_DEVICE void shape_forward_vec_body(const ShapeParams& p, float* smem, int block, int tid, int grid, int wpc) {
  const int warp = tid >> 5, lane = tid & 31;
  const int span = p.span, g = p.block, g4 = g >> 2;
  const int spans = (p.T + span - 1) / span;
  const long long items = (long long)p.rows * spans;
  float4* blk = reinterpret_cast<float4*>(smem) + warp * kShapeMaxBlocks;
  double acc[kShapeMaxWin];
#pragma unroll
  for (int r = 0; r < kShapeMaxWin; ++r) acc[r] = 0.0;
  for (long long it = (long long)block * wpc + warp; it < items; it += (long long)grid * wpc) {
    const int row = (int)(it / spans), s0 = (int)(it - (long long)row * spans) * span;
    const int s1 = min(s0 + span, p.T);
    const float* __restrict__ xr = p.x + (size_t)row * p.T;
    const float* __restrict__ yr = p.y + (size_t)row * p.T;
    const int nb = (s1 - s0) / g;
    float4 xa[NV], ya[NV], xb[NV], yb[NV];
    if (nb > 0) shape_load4<NV>(xr, yr, s0, g4, lane, xa, ya);
    for (int bi = 0; bi < nb; ++bi) {
      if (bi + 1 < nb) shape_load4<NV>(xr, yr, s0 + (bi + 1) * g, g4, lane, xb, yb);
      float mx = -1.f, my = 0.f, sx = 0.f;
      int ix = 0x7fffffff;
#pragma unroll
      for (int v = 0; v < NV; ++v) {                       // ascending index: strict > keeps the first maximum per lane
        const int q = lane + 32 * v;
        if (q < g4) {
          const int i0 = s0 + bi * g + 4 * q;
          const float xs[4] = {xa[v].x, xa[v].y, xa[v].z, xa[v].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float ax = fabsf(xs[c]);
            if (ax > mx) { mx = ax; ix = i0 + c; sx = xs[c]; }
          }
          my = fmaxf(fmaxf(my, fmaxf(fabsf(ya[v].x), fabsf(ya[v].y))), fmaxf(fabsf(ya[v].z), fabsf(ya[v].w)));
        }
      }
      float mxo, myo;
      bool own;
      shape_finish(mx, my, ix, mxo, myo, own);
      if (own) blk[bi] = make_float4(mxo, myo, sx, bits_to_float(ix));
#pragma unroll
      for (int v = 0; v < NV; ++v) { xa[v] = xb[v]; ya[v] = yb[v]; }
    }
    __syncwarp();
    shape_window_pass(p, blk, row, s0, s1, lane, acc);
    __syncwarp();
  }
  shape_write_partials(p, acc, block * wpc + warp, lane);
}

// A warp takes the windows that start inside its span.
//   p.block > 0 (every window length is a multiple of p.block in [32, 256] and the span a multiple of every window length):
//     ONE pass over the samples in blocks of p.block -- per block (max |x|, its first index, the sample there, max |y|)
//     into the warp's shared-memory scratch -- then every window of every length is combined from its blocks, one
//     window per lane.  Each sample is loaded once.
//   p.block == 0 (window lengths without a usable common divisor): one pass per window length, lanes across the
//     samples of a window (the span stays in L1 between the passes).
SPL_DEVICE void shape_forward_body(const ShapeParams& p, float* smem, int block, int tid, int grid, int wpc) {
  const int warp = tid >> 5, lane = tid & 31;
  const int span = p.span;
  const int spans = (p.T + span - 1) / span;
  const long long items = (long long)p.rows * spans;
  float4* blk = reinterpret_cast<float4*>(smem) + warp * kShapeMaxBlocks;     // {max|x|, max|y|, x at argmax, bits(argmax)}
  double acc[kShapeMaxWin];
#pragma unroll
  for (int r = 0; r < kShapeMaxWin; ++r) acc[r] = 0.0;
  for (long long it = (long long)block * wpc + warp; it < items; it += (long long)grid * wpc) {
    const int row = (int)(it / spans), s0 = (int)(it - (long long)row * spans) * span;
    const int s1 = min(s0 + span, p.T);
    const float* __restrict__ xr = p.x + (size_t)row * p.T;
    const float* __restrict__ yr = p.y + (size_t)row * p.T;
    if (p.block > 0) {
      const int g = p.block, nb = (s1 - s0) / g;           // whole blocks inside the span (windows never use a partial one)
      for (int bi = 0; bi < nb; ++bi) {
        float mx, my, sx;
        int ix;
        bool own;
        shape_scan(xr, yr, s0 + bi * g, g, lane, mx, my, sx, ix, own);
        if (own) blk[bi] = make_float4(mx, my, sx, bits_to_float(ix));
      }
      __syncwarp();
      shape_window_pass(p, blk, row, s0, s1, lane, acc);
      __syncwarp();
    } else {
#pragma unroll
      for (int r = 0; r < kShapeMaxWin; ++r) {
        if (r >= p.n) continue;
        const int w = p.win[r], nw = p.T / w;
        int j0, j1;
        shape_window_range(s0, s1, w, p.T, j0, j1);
        int* rec = p.records + p.rec_ofs[r] + (long long)row * nw;
        float sum = 0.f;
        for (int j = j0; j < j1; ++j) {
          float mx, my, sx;
          int ix;
          bool own;
          shape_scan(xr, yr, j * w, w, lane, mx, my, sx, ix, own);
          const float d = mx - my;
          if (lane == 0) sum += fabsf(d);
          if (own) rec[j] = shape_record(ix, d, sx);
        }
        acc[r] += (double)sum;
      }
    }
  }
  shape_write_partials(p, acc, block * wpc + warp, lane);
}

// [region: shape backward]
// one warp per span: the span of dx is assembled in the warp's shared-memory tile (window lengths one after the
// other: within one length every sample is hit at most once, so there are no conflicts and no atomics) and written
// out with coalesced 16-byte stores.
SPL_DEVICE void shape_backward_body(const ShapeParams& p, float* smem, int block, int tid, int grid, int wpc) {
  const int warp = tid >> 5, lane = tid & 31;
  float* tile = smem + warp * kShapeSpan;
  const int spans = (p.T + kShapeSpan - 1) / kShapeSpan;
  const long long items = (long long)p.rows * spans;
  const float g = *p.g;
  for (long long it = (long long)block * wpc + warp; it < items; it += (long long)grid * wpc) {
    const int row = (int)(it / spans), s0 = (int)(it - (long long)row * spans) * kShapeSpan;
    const int s1 = min(s0 + kShapeSpan, p.T);
    for (int i = lane; i < kShapeSpan; i += 32) tile[i] = 0.f;
    __syncwarp();
    for (int r = 0; r < p.n; ++r) {
      const int w = p.win[r], nw = p.T / w;
      const int j0 = s0 / w, j1 = min((s1 + w - 1) / w, nw);          // windows that OVERLAP the span
      const float c = g * p.coef[r];
      for (int j = j0 + lane; j < j1; j += 32) {
        const int rec = __ldg(p.records + p.rec_ofs[r] + (long long)row * nw + j);
        const int ix = rec >> 2, sg = (rec & 3) - 1;
        if (ix >= s0 && ix < s1) tile[ix - s0] += c * (float)sg;
      }
      __syncwarp();
    }
    float* out = p.dx + (size_t)row * p.T + s0;
    if ((p.T & 3) == 0) {
      for (int i = 4 * lane; i < s1 - s0; i += 128)
        *reinterpret_cast<float4*>(out + i) = make_float4(tile[i], tile[i + 1], tile[i + 2], tile[i + 3]);
    } else {
      for (int i = lane; i < s1 - s0; i += 32) out[i] = tile[i];
    }
    __syncwarp();
  }
}

// loss = mean over window lengths of S_r / (rows_global * (T / w_r)); one thread
struct ShapeFinalizeParams {
  int n;
  double count[kShapeMaxWin];
  const double* sums;
  float* loss;
};
SPL_DEVICE void shape_finalize_body(const ShapeFinalizeParams& p) {
  double l = 0.0;
  for (int r = 0; r < p.n; ++r) l += __ldcg(&p.sums[r]) / p.count[r];
  *p.loss = (float)(l / p.n);
}

// ---------------------------------------------------------------------------------------------
// Losses on EXPLICIT magnitude tensors: SpectralConvergenceLoss.forward(x_mag, y_mag) = ||y - x||_F / ||y||_F
// (losses/stft_loss.py:38-56) and LogSTFTMagnitudeLoss.forward = mean |ln y - ln x| (stft_loss.py:59-77), for callers
// that compose them with stft() themselves (STFTLoss.forward does, stft_loss.py:112-116).  HBM-bound streaming:
// forward reads both tensors once (8 B per element), backward reads both and writes one or two gradients.
// ---------------------------------------------------------------------------------------------
struct MagLossParams {
  const float* x;          // x_mag, n elements
  const float* y;          // y_mag
  long long n;
  int vec;                 // 1: both tensors (and the gradients) are 16-byte aligned -> float4 body + scalar tail
  double* partials;        // forward: [grid * warps per CTA][3]  (S1 = sum (y-x)^2, S2 = sum y^2, S3 = sum |ln y - ln x|)
  // backward
  const double* sums;      // [6]: S1, S2, S3 and the three backward coefficients
  const float* g_sc;       // upstream gradients (device scalars, null = 0)
  const float* g_mag;
  float* gx;               // d/dx_mag (null = skip)
  float* gy;               // d/dy_mag (null = skip)
};

// [region: mag loss]
SPL_DEVICE void mag_accumulate(float xv, float yv, float& s1, float& s2, float& s3) {
  const float d = yv - xv;
  s1 = fmaf(d, d, s1);
  s2 = fmaf(yv, yv, s2);
  s3 += fabsf(spl_fast_log2(yv) - spl_fast_log2(xv));               // x ln 2 once, at the end
}

// 16-byte loads over the aligned body (vec: n / 4 float4 per tensor), scalar loads over the tail
SPL_DEVICE void mag_sums_body(const MagLossParams& p, int block, int tid, int grid, int wpc) {
  const int warp = tid >> 5, lane = tid & 31;
  const long long stride = (long long)grid * wpc * 32, first = ((long long)block * wpc + warp) * 32 + lane;
  double d1 = 0.0, d2 = 0.0, d3 = 0.0;
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int pending = 0;
  const long long nv = p.vec ? p.n >> 2 : 0;
  const float4* x4 = reinterpret_cast<const float4*>(p.x);
  const float4* y4 = reinterpret_cast<const float4*>(p.y);
  for (long long i = first; i < nv; i += stride) {
    const float4 xv = __ldg(x4 + i), yv = __ldg(y4 + i);
    mag_accumulate(xv.x, yv.x, s1, s2, s3);
    mag_accumulate(xv.y, yv.y, s1, s2, s3);
    mag_accumulate(xv.z, yv.z, s1, s2, s3);
    mag_accumulate(xv.w, yv.w, s1, s2, s3);
    if (++pending == 16) { d1 += (double)s1; d2 += (double)s2; d3 += (double)s3; s1 = s2 = s3 = 0.f; pending = 0; }
  }
  for (long long i = 4 * nv + first; i < p.n; i += stride) mag_accumulate(__ldg(p.x + i), __ldg(p.y + i), s1, s2, s3);
  d1 = warp_sum(d1 + (double)s1);
  d2 = warp_sum(d2 + (double)s2);
  d3 = warp_sum(d3 + (double)s3);
  if (lane == 0) {
    double* o = p.partials + (size_t)(block * wpc + warp) * 3;
    o[0] = d1; o[1] = d2; o[2] = 0.69314718055994531 * d3;
  }
}

// gradients of  g_sc * sqrt(S1)/sqrt(S2) + g_mag * S3/n  (what autograd derives: torch.norm backward gives 0 at D = 0,
// sign(0) = 0).  sums[3..5] = {1/(D Ny) (0 at D = 0), D/Ny^3, 1/n} from the finalize step.  For positive magnitudes
// sign(ln x - ln y) = sign(x - y): no logarithm on this path.
SPL_DEVICE void mag_grad(float xv, float yv, float a, float b, float c, float& gx, float& gy) {
  const float d = xv - yv;
  const float sg = (d > 0.f) ? c : ((d < 0.f) ? -c : 0.f);
  gx = fmaf(a, d, sg / xv);
  gy = -fmaf(a, d, fmaf(b, yv, sg / yv));
}

// persistent grid-stride loop (one float4 per thread per CTA would spend its time launching 200 k tiny CTAs)
SPL_DEVICE void mag_backward_body(const MagLossParams& p, long long first, long long stride) {
  const float gsc = p.g_sc ? *p.g_sc : 0.f, gmag = p.g_mag ? *p.g_mag : 0.f;
  const float a = gsc * (float)__ldcg(&p.sums[3]), b = gsc * (float)__ldcg(&p.sums[4]), c = gmag * (float)__ldcg(&p.sums[5]);
  const long long nv = p.vec ? p.n >> 2 : 0;
  const float4* x4 = reinterpret_cast<const float4*>(p.x);
  const float4* y4 = reinterpret_cast<const float4*>(p.y);
  for (long long i = first; i < nv; i += 2 * stride) {         // two independent float4 pairs in flight per thread
    const long long i2 = i + stride;
    const bool two = i2 < nv;
    const float4 xa = __ldg(x4 + i), ya = __ldg(y4 + i);
    const float4 xb = two ? __ldg(x4 + i2) : xa, yb = two ? __ldg(y4 + i2) : ya;
    float4 gx, gy;
    mag_grad(xa.x, ya.x, a, b, c, gx.x, gy.x);
    mag_grad(xa.y, ya.y, a, b, c, gx.y, gy.y);
    mag_grad(xa.z, ya.z, a, b, c, gx.z, gy.z);
    mag_grad(xa.w, ya.w, a, b, c, gx.w, gy.w);
    if (p.gx) reinterpret_cast<float4*>(p.gx)[i] = gx;
    if (p.gy) reinterpret_cast<float4*>(p.gy)[i] = gy;
    if (two) {
      mag_grad(xb.x, yb.x, a, b, c, gx.x, gy.x);
      mag_grad(xb.y, yb.y, a, b, c, gx.y, gy.y);
      mag_grad(xb.z, yb.z, a, b, c, gx.z, gy.z);
      mag_grad(xb.w, yb.w, a, b, c, gx.w, gy.w);
      if (p.gx) reinterpret_cast<float4*>(p.gx)[i2] = gx;
      if (p.gy) reinterpret_cast<float4*>(p.gy)[i2] = gy;
    }
  }
  for (long long i = 4 * nv + first; i < p.n; i += stride) {   // scalar tail (or everything when not 16-byte aligned)
    float gx, gy;
    mag_grad(__ldg(p.x + i), __ldg(p.y + i), a, b, c, gx, gy);
    if (p.gx) p.gx[i] = gx;
    if (p.gy) p.gy[i] = gy;
  }
}

struct MagFinalizeParams {
  double* sums;            // [6]: in S1, S2, S3; out [3..5] = backward coefficients 1/(D Ny), D/Ny^3, 1/n
  double n;
  float* sc;               // null = skip
  float* mag;
};
SPL_DEVICE void mag_finalize_body(const MagFinalizeParams& p) {
  const double S2 = __ldcg(&p.sums[1]);
  const double D = sqrt(__ldcg(&p.sums[0])), Ny = sqrt(S2);
  if (p.sc) *p.sc = (float)(D / Ny);
  if (p.mag) *p.mag = (float)(__ldcg(&p.sums[2]) / p.n);
  p.sums[3] = D > 0.0 ? 1.0 / (D * Ny) : 0.0;
  p.sums[4] = D / (Ny * S2);
  p.sums[5] = 1.0 / p.n;
}

// ---------------------------------------------------------------------------------------------
// deterministic reduction of the per-warp partial sums: one CTA per output sum
// ---------------------------------------------------------------------------------------------
constexpr int kMaxSums = 24;   // 3 sums x SPL_MAX_TRANSFORMS resolutions
struct ReduceParams {
  int n_sums;
  const double* base[kMaxSums];   // first element of the column
  int stride[kMaxSums];           // doubles between consecutive items
  int count[kMaxSums];            // items
  double* out;                    // [n_sums]
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
// Multi-GPU: reduce -> exchange over NVLink peer memory -> finalize in ONE kernel (the single exchange step
// of the sharded loss).  Replaces spl_reduce + ncclAllReduce(10 doubles) + spl_finalize: for an 80-byte payload the
// collective is pure latency (two launches + NCCL's own protocol), so the last CTA of the reduction pushes this rank's
// sums straight into every peer's symmetric buffer (peer-mapped stores), raises a flag there, waits for the peers'
// flags in its own buffer and adds the contributions in rank order -- every rank gets bit-identical global sums.
//   symmetric buffer (one per rank, peer-mapped):  double slots[2][kMaxRanks][kExchangeSums];  unsigned flags[2][kMaxRanks];
// Calls are numbered by an epoch kept in device memory (CUDA-graph replay safe); slot/flag sets alternate with the
// epoch's parity: a rank can be at most one call ahead of the slowest peer, so the set being written is never the set a
// peer still reads.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
constexpr int kExchangeSums = kMaxSums;
constexpr size_t kExchangeBytes = 2 * kMaxRanks * kExchangeSums * sizeof(double) + 2 * kMaxRanks * sizeof(unsigned);

struct ExchangeParams {
  ReduceParams r;          // r.out = this rank's sums (local)
  FinalizeParams f;        // f.sums = gsums
  double* gsums;           // local: receives the global sums
  unsigned* state;         // local: [0] CTA ticket (self-resetting), [1] epoch of the last completed call
  unsigned* error_flag;    // null, or a word the HOST can read without synchronising (mapped pinned memory): receives the
                           // epoch of a call whose peers did not arrive in time
  long long timeout_ns;    // how long to wait for the peers' flags; <= 0: wait for ever, like a collective
  int rank, world;
  void* peers[kMaxRanks];  // symmetric buffers of all ranks, peers[rank] = own
};

#ifdef SPECLOSS_EMU
static inline void st_release_sys(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned ld_acquire_sys(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline double ld_volatile_f64(const double* p) { return *reinterpret_cast<const volatile double*>(p); }
static inline void threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void st_relaxed_sys(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
static inline long long spin_clock() { static thread_local long long c = 0; return c += 64; }     // "ns" 
#else
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void threadfence_system() { __threadfence_system(); }
__device__ __forceinline__ long long spin_clock() {      // nanoseconds (globaltimer: independent of the SM clock)
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif

// [region: exchange]
// one warp (the first warp of the last CTA of the reduction); `lane` in [0, 32)
SPL_DEVICE void exchange_body(const ExchangeParams& p, int lane) {
  const int n = p.r.n_sums, world = p.world;
  const unsigned epoch = __ldcg(&p.state[1]) + 1u;
  const int parity = (int)(epoch & 1u);
  // publish this rank's sums in every rank's buffer (own included)
  for (int t = lane; t < world * n; t += 32) {
    const int peer = t / n, j = t - peer * n;
    double* slots = reinterpret_cast<double*>(p.peers[peer]);
    slots[(parity * kMaxRanks + p.rank) * kExchangeSums + j] = __ldcg(&p.r.out[j]);
  }
  threadfence_system();
  __syncwarp();
  if (lane < world) {
    unsigned* flags = reinterpret_cast<unsigned*>(reinterpret_cast<double*>(p.peers[lane]) + 2 * kMaxRanks * kExchangeSums);
    st_release_sys(&flags[parity * kMaxRanks + p.rank], epoch);
  }
  // Wait for every rank's contribution to this call.  Like a collective this waits as long as it takes by default
  // (rank skew of seconds is routine in training: checkpoints, validation, data stalls).  With a finite timeout_ns a
  // missing peer does not hang the stream: the call's epoch goes to the host-visible error flag (the host raises on its
  // next call) and the sums are poisoned with NaN so the step cannot be used by accident.
  bool ok = true;
  if (lane < world) {
    const unsigned* flags = reinterpret_cast<const unsigned*>(reinterpret_cast<const double*>(p.peers[p.rank]) + 2 * kMaxRanks * kExchangeSums);
    const long long t0 = spin_clock();
    while (ld_acquire_sys(&flags[parity * kMaxRanks + lane]) != epoch) {
      if (p.timeout_ns > 0 && spin_clock() - t0 > p.timeout_ns) { ok = false; break; }
    }
  }
  threadfence_system();
  ok = __ballot_sync(0xffffffffu, !ok) == 0u;
  if (!ok && lane == 0 && p.error_flag) st_relaxed_sys(p.error_flag, epoch);
  if (lane < n) {
    const double* slots = reinterpret_cast<const double*>(p.peers[p.rank]);
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += ld_volatile_f64(&slots[(parity * kMaxRanks + r) * kExchangeSums + lane]);   // rank order
    p.gsums[lane] = ok ? acc : acc * __longlong_as_double_nan();
  }
  __syncwarp();
  if (lane == 0) {
    p.state[1] = epoch;
    __threadfence();
    finalize_body(p.f);
  }
}

// ---------------------------------------------------------------------------------------------
// backward: dx[b, i] = sum over transforms of coef * (overlap-added frame gradients), gathered
// from the per-frame slots (no atomics: every slot has one writer, every dx sample one reader)
// and folded over the reflect-padding margins.
// ---------------------------------------------------------------------------------------------
struct CombineEntry {
  const void* frames;    // [B * n_frames][win] float2 (u, v) (stft) | float (mel) | planar stft: [B * n_frames][2][win] float
  int kind, half, hop, win, left, n_frames;
  int planar;            // stft written by the even/odd 2048-point kernel: a u plane and a v plane per frame
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
  int unit;              // 1: every entry is taken with coefficient 1 (spectrogram backward; coefs / g_* unused)
};

// overlap-added value at one padded position (used for the reflect-fold margins): frames t with
// 0 <= q - t*hop < win, q = position relative to the first live tap of frame 0
SPL_DEVICE float gather_padded(const CombineEntry& e, int b, int ppos, float cu, float cv) {
  const int q = ppos - e.left;
  if (q < 0) return 0.f;
  float acc = 0.f;
  for (int t = min(q / e.hop, e.n_frames - 1); t >= 0; --t) {
    const int tap = q - t * e.hop;
    if (tap >= e.win) break;
    const size_t o = ((size_t)b * e.n_frames + t) * e.win + tap;
    if (e.kind == kKindStft && e.planar) {
      const float* src = reinterpret_cast<const float*>(e.frames) + (((size_t)b * e.n_frames + t) * 2) * e.win + tap;
      acc = fmaf(cu, __ldg(src), fmaf(cv, __ldg(src + e.win), acc));
    } else if (e.kind == kKindStft) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(e.frames) + o);
      acc = fmaf(cu, v.x, fmaf(cv, v.y, acc));
    } else {
      acc = fmaf(cu, __ldg(reinterpret_cast<const float*>(e.frames) + o), acc);
    }
  }
  return acc;
}

// same for four consecutive padded positions ppos0 .. ppos0+3: one walk over the covering frames, 16-byte loads
// when the four taps sit inside the frame and the address is aligned
SPL_DEVICE void gather_padded4(const CombineEntry& e, int b, int ppos0, float cu, float cv, float (&acc)[4]) {
  const int q0 = ppos0 - e.left;
  if (q0 + 3 < 0) return;
  for (int t = min((q0 + 3) / e.hop, e.n_frames - 1); t >= 0; --t) {
    const int tap = q0 - t * e.hop;                 // tap of the first of the four positions (may be negative)
    if (tap >= e.win) break;
    const size_t o = ((size_t)b * e.n_frames + t) * e.win + tap;     // meaningful for tap >= 0 only
    if (e.kind == kKindStft && e.planar) {
      const size_t ou = (((size_t)b * e.n_frames + t) * 2) * e.win + tap;
      const float* su = reinterpret_cast<const float*>(e.frames) + ou;
      const float* sv = su + e.win;
      if (tap >= 0 && tap + 3 < e.win && ((ou | (size_t)e.win) & 3) == 0) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(su));
        const float4 v = __ldg(reinterpret_cast<const float4*>(sv));
        acc[0] = fmaf(cu, u.x, fmaf(cv, v.x, acc[0]));
        acc[1] = fmaf(cu, u.y, fmaf(cv, v.y, acc[1]));
        acc[2] = fmaf(cu, u.z, fmaf(cv, v.z, acc[2]));
        acc[3] = fmaf(cu, u.w, fmaf(cv, v.w, acc[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (tap + j >= 0 && tap + j < e.win) acc[j] = fmaf(cu, __ldg(su + j), fmaf(cv, __ldg(sv + j), acc[j]));
      }
    } else if (e.kind == kKindStft) {
      const float2* src = reinterpret_cast<const float2*>(e.frames);
      if (tap >= 0 && tap + 3 < e.win && (o & 1) == 0) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + o));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + o + 2));
        acc[0] = fmaf(cu, v0.x, fmaf(cv, v0.y, acc[0]));
        acc[1] = fmaf(cu, v0.z, fmaf(cv, v0.w, acc[1]));
        acc[2] = fmaf(cu, v1.x, fmaf(cv, v1.y, acc[2]));
        acc[3] = fmaf(cu, v1.z, fmaf(cv, v1.w, acc[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (tap + j >= 0 && tap + j < e.win) {
            const float2 v = __ldg(src + (o + j));
            acc[j] = fmaf(cu, v.x, fmaf(cv, v.y, acc[j]));
          }
      }
    } else {
      const float* src = reinterpret_cast<const float*>(e.frames);
      if (tap >= 0 && tap + 3 < e.win && (o & 3) == 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + o));
        acc[0] = fmaf(cu, v.x, acc[0]); acc[1] = fmaf(cu, v.y, acc[1]);
        acc[2] = fmaf(cu, v.z, acc[2]); acc[3] = fmaf(cu, v.w, acc[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (tap + j >= 0 && tap + j < e.win) acc[j] = fmaf(cu, __ldg(src + (o + j)), acc[j]);
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
    if (p.unit) { cu = 1.f; cv = 0.f; }
    else if (e.kind == kKindStft) { cu = gsc * p.coefs[2 * r]; cv = gmag * p.coefs[2 * r + 1]; }
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

template <int NFFT, int KIND, bool GRAD, int WIN_T, bool RING = false>
__global__ void __launch_bounds__(MaxWarps<NFFT>::value * 32, 1) transform_kernel(const TransformParams p) {
  extern __shared__ __align__(16) float smem_dyn[];
  cta_load_tables<NFFT, KIND>(p, smem_dyn, threadIdx.x, blockDim.x);
  __syncthreads();
  transform_body<NFFT, KIND, GRAD, WIN_T, RING>(p, smem_dyn, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
template <int NFFT>
__global__ void __launch_bounds__(MaxWarps<NFFT>::value * 32, 1) spec_kernel(const SpecParams p) {
  extern __shared__ __align__(16) float smem_dyn[];
  cta_load_fft_tables<NFFT>(cta_tables(NFFT, p.win, kKindStft, 0, 0), p.twiddle, p.window, p.win, smem_dyn, threadIdx.x,
                            blockDim.x);
  cp_async_wait_all();
  __syncthreads();
  spec_body<NFFT>(p, smem_dyn, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
template <int NFFT, int KIND>
__global__ void __launch_bounds__(MaxWarps<NFFT>::value * 32, 1) specgrad_kernel(const SpecGradParams q) {
  extern __shared__ __align__(16) float smem_dyn[];
  cta_load_tables<NFFT, KIND>(q.t, smem_dyn, threadIdx.x, blockDim.x);
  __syncthreads();
  specgrad_body<NFFT, KIND>(q, smem_dyn, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
template <int NV>
__global__ void __launch_bounds__(256) shape_forward_vec_kernel(const ShapeParams p) {
  __shared__ __align__(16) float scratch[8 * kShapeMaxBlocks * 4];          // 8 warps per CTA (spl_shape_dims)
  shape_forward_vec_body<NV>(p, scratch, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
__global__ void __launch_bounds__(256) shape_forward_kernel(const ShapeParams p) {
  __shared__ __align__(16) float scratch[8 * kShapeMaxBlocks * 4];          // 8 warps per CTA (spl_shape_dims)
  shape_forward_body(p, scratch, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
__global__ void __launch_bounds__(256) shape_backward_kernel(const ShapeParams p) {
  extern __shared__ __align__(16) float smem_dyn[];
  shape_backward_body(p, smem_dyn, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
__global__ void shape_finalize_kernel(const ShapeFinalizeParams p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) shape_finalize_body(p);
}
__global__ void __launch_bounds__(256) mag_sums_kernel(const MagLossParams p) {
  mag_sums_body(p, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
__global__ void __launch_bounds__(256) mag_backward_kernel(const MagLossParams p) {
  mag_backward_body(p, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}
__global__ void mag_finalize_kernel(const MagFinalizeParams p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) mag_finalize_body(p);
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
__global__ void __launch_bounds__(256) reduce_exchange_finalize_kernel(const ExchangeParams p) {
  __shared__ double sh[256];
  __shared__ bool last;
  reduce_body(p.r, sh, blockIdx.x, threadIdx.x, 256);
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicInc(&p.state[0], gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    exchange_body(p, threadIdx.x);
  }
}
__global__ void __launch_bounds__(128) combine_kernel(const CombineParams p) {
  combine_body(p, (long long)blockIdx.x * 128 + threadIdx.x);
}
#endif

}  // namespace spl
