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

#define SPL_DEVICE __device__ __forceinline__
#include "fft_codelets.cuh"

#ifdef SPECLOSS_EMU
#define SPL_FAST_LOGF(x) logf(x)
#else
#define SPL_FAST_LOGF(x) __logf(x)     // MUFU.LG2 path; abs error ~1e-7 on the log-magnitude terms
#endif

namespace spl {

constexpr int kWarpsPerCta = 4;
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
template <> struct Dft<16> { static SPL_DEVICE void run(float (&re)[16], float (&im)[16]) { fft16(re, im); } };
template <> struct Dft<32> { static SPL_DEVICE void run(float (&re)[32], float (&im)[32]) { fft32(re, im); } };
template <> struct Dft<64> { static SPL_DEVICE void run(float (&re)[64], float (&im)[64]) { fft64(re, im); } };

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
  const int* mel_row_start;   // first bin with non-zero weight, per mel row
  const int* mel_row_len;     // number of consecutive bins stored for the row
  const int* mel_row_ptr;     // offset of the row's weights in mel_row_val
  const float* mel_row_val;
  const int* bin_m0;          // per bin: the two (adjacent) mel rows it feeds are m0, m0 + 1
  const float* bin_w0;
  const float* bin_w1;
};

SPL_DEVICE float2 cmul(float2 a, float2 w) {            // a * w
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
SPL_DEVICE float2 cmul_conj(float2 a, float2 w) {       // a * conj(w)
  return make_float2(fmaf(a.x, w.x, a.y * w.y), fmaf(a.y, w.x, -a.x * w.y));
}

SPL_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

SPL_DEVICE int align4(int n) { return (n + 3) & ~3; }

// shared memory carve-up per warp (in 4-byte words); must match spl_smem_bytes() on the host side
template <int NFFT, int KIND, bool GRAD>
struct SmemLayout {
  using G = FftGeom<NFFT>;
  static constexpr int FPW = 32 / G::L;                      // frames in flight per warp
  static constexpr int BUF_F2 = G::R * (G::L + 1);           // float2 per frame slot (>= NFFT + 2)
  static __host__ __device__ int words_per_warp(int ring_n, int n_mels) {
    int w = FPW * BUF_F2 * 2;
    if (GRAD) w += (KIND == kKindStft ? 2 : 1) * ((ring_n + 3) & ~3);
    if (KIND == kKindMel) w += FPW * ((n_mels + 3) & ~3);
    return (w + 3) & ~3;
  }
};

// ---------------------------------------------------------------------------------------------
// forward DFT of the frame held in (re, im) [time layout]  ->  natural-order spectrum in `nat`
// ---------------------------------------------------------------------------------------------
template <int NFFT>
SPL_DEVICE void forward_to_natural(float (&re)[FftGeom<NFFT>::R], float (&im)[FftGeom<NFFT>::R],
                                   float2* buf, const float2* __restrict__ tw, int l) {
  using G = FftGeom<NFFT>;
  constexpr int L = G::L, R = G::R, RPL = R / L;
  Dft<R>::run(re, im);
#pragma unroll
  for (int k2 = 0; k2 < R; ++k2) {
    float2 v = make_float2(re[k2], im[k2]);
    if (k2 > 0) v = cmul(v, __ldg(&tw[k2 * L + l]));
    buf[k2 * (L + 1) + l] = v;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
#pragma unroll
    for (int n1 = 0; n1 < L; ++n1) {
      const float2 v = buf[(j * L + l) * (L + 1) + n1];
      re[j * L + n1] = v.x;
      im[j * L + n1] = v.y;
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < RPL; ++j)
    Dft<L>::run(reinterpret_cast<float(&)[L]>(re[j * L]), reinterpret_cast<float(&)[L]>(im[j * L]));
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
#pragma unroll
    for (int k1 = 0; k1 < L; ++k1) buf[(j * L + l) + R * k1] = make_float2(re[j * L + k1], im[j * L + k1]);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// natural-order spectrum H in `nat`  ->  un-normalised inverse DFT (kernel e^{+i...}) in time layout
// ---------------------------------------------------------------------------------------------
template <int NFFT>
SPL_DEVICE void natural_to_time(float (&re)[FftGeom<NFFT>::R], float (&im)[FftGeom<NFFT>::R],
                                float2* buf, const float2* __restrict__ tw, int l) {
  using G = FftGeom<NFFT>;
  constexpr int L = G::L, R = G::R, RPL = R / L;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
#pragma unroll
    for (int k1 = 0; k1 < L; ++k1) {
      const float2 v = buf[(j * L + l) + R * k1];
      re[j * L + k1] = v.x;
      im[j * L + k1] = v.y;
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < RPL; ++j)   // inverse = forward codelet on swapped components
    Dft<L>::run(reinterpret_cast<float(&)[L]>(im[j * L]), reinterpret_cast<float(&)[L]>(re[j * L]));
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
#pragma unroll
    for (int n1 = 0; n1 < L; ++n1) buf[(j * L + l) * (L + 1) + n1] = make_float2(re[j * L + n1], im[j * L + n1]);
  }
  __syncwarp();
#pragma unroll
  for (int k2 = 0; k2 < R; ++k2) {
    float2 v = buf[k2 * (L + 1) + l];
    if (k2 > 0) v = cmul_conj(v, __ldg(&tw[k2 * L + l]));
    re[k2] = v.x;
    im[k2] = v.y;
  }
  __syncwarp();
  Dft<R>::run(im, re);
}

// reflect index without edge repeat (torch.stft center=True, pad_mode="reflect")
SPL_DEVICE int reflect(int s, int T) {
  s = s < 0 ? -s : s;
  return s >= T ? 2 * (T - 1) - s : s;
}

// ---------------------------------------------------------------------------------------------
// The transform kernel body.  One warp = one chunk of `m` consecutive frames of one utterance.
// ---------------------------------------------------------------------------------------------
template <int NFFT, int KIND, bool GRAD>
SPL_DEVICE void transform_body(const TransformParams& p, float* smem, int block, int tid) {
  using G = FftGeom<NFFT>;
  using SL = SmemLayout<NFFT, KIND, GRAD>;
  constexpr int L = G::L, R = G::R, FPW = SL::FPW, HALF = NFFT / 2, NPAIR = NFFT / (2 * L);
  const int warp = tid >> 5, lane = tid & 31;
  const int l = lane & (L - 1), h = lane / L;

  const int chunk_id = block * kWarpsPerCta + warp;
  if (chunk_id >= p.B * p.n_chunks) return;
  const int b = chunk_id / p.n_chunks, c = chunk_id - b * p.n_chunks;
  const int t0 = c * p.m;
  const int m_c = min(p.m, p.n_frames - t0);

  float* wsm = smem + (size_t)warp * SL::words_per_warp(p.ring_n, p.n_mels);
  float2* buf = reinterpret_cast<float2*>(wsm) + h * SL::BUF_F2;     // this frame slot's exchange buffer
  float* ring_base = wsm + FPW * SL::BUF_F2 * 2;
  float2* ring2 = reinterpret_cast<float2*>(ring_base);              // stft: (u, v)
  float* ring1 = ring_base;                                          // mel : u
  float* gm_s = ring_base + (GRAD ? (KIND == kKindStft ? 2 : 1) * align4(p.ring_n) : 0) + h * align4(p.n_mels);

  const float* xb = p.x + (size_t)b * p.T;
  const float* yb = p.y + (size_t)b * p.T;
  const float2* __restrict__ tw = p.twiddle;

  if (GRAD) {
    if (KIND == kKindStft) for (int i = lane; i < p.ring_n; i += 32) ring2[i] = make_float2(0.f, 0.f);
    else                   for (int i = lane; i < p.ring_n; i += 32) ring1[i] = 0.f;
    __syncwarp();
  }
  float2* out2 = reinterpret_cast<float2*>(p.gchunks) + (size_t)chunk_id * p.span;
  float* out1 = reinterpret_cast<float*>(p.gchunks) + (size_t)chunk_id * p.span;
  const int span_c = (m_c - 1) * p.hop + p.win;
  int flushed = 0;

  float s1 = 0.f, s2 = 0.f, s3 = 0.f;   // stft: S1, S2, S3 ; mel: s1 = S4

  float re[R], im[R];
  for (int step = 0; step * FPW < m_c; ++step) {
    const int jc = step * FPW + h;          // frame index inside the chunk
    const bool active = jc < m_c;
    const int t = t0 + jc;
    // ---- A. load taps, reflect-pad, window; pack z = x*w + i*y*w ------------------------------
    bool same = true;
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) {
      const int n = l + L * n2;
      const int tap = n - p.left;
      float xv = 0.f, yv = 0.f;
      if (active && tap >= 0 && tap < p.win) {
        const int s = reflect(t * p.hop + n - HALF, p.T);
        const float w = __ldg(&p.window[tap]);
        xv = __ldg(&xb[s]) * w;
        yv = __ldg(&yb[s]) * w;
      }
      same = same && (xv == yv);
      re[n2] = xv;
      im[n2] = yv;
    }
    // A frame whose prediction and target taps are bit-identical must contribute exactly zero (the
    // reference returns sc = mag = mel = 0 and a zero gradient for x == y); the packed FFT would
    // leave ~1e-7 of rounding asymmetry between X and Y, so such frames reuse X for Y below.
    const unsigned eq_bits = __ballot_sync(0xffffffffu, same);
    const unsigned grp_mask = (L == 32) ? 0xffffffffu : (((1u << (L & 31)) - 1u) << (h * L));
    const bool frame_equal = (eq_bits & grp_mask) == grp_mask;
    // ---- B. FFT, natural-order Z in buf -------------------------------------------------------
    forward_to_natural<NFFT>(re, im, buf, tw, l);

    const float act = active ? 1.f : 0.f;
    if (KIND == kKindStft) {
      // ---- C. separate X, Y from Z = FFT(x + i y); loss terms; gradient spectrum H -------------
#pragma unroll 4
      for (int i = 0; i <= NPAIR; ++i) {
        const bool extra = (i == NPAIR);            // bin N/2, lane 0 of the group only
        if (extra && l != 0) break;
        const int k = extra ? HALF : l + L * i;
        const int km = (NFFT - k) & (NFFT - 1);
        const float2 a = buf[k], bm = buf[km];
        const float xr = 0.5f * (a.x + bm.x), xi = 0.5f * (a.y - bm.y);   // X[k]
        const float yr = frame_equal ? xr : 0.5f * (a.y + bm.y);          // Y[k]
        const float yi = frame_equal ? xi : 0.5f * (bm.x - a.x);
        const float px = fmaf(xr, xr, xi * xi), py = fmaf(yr, yr, yi * yi);
        const float pxc = fmaxf(px, p.eps), pyc = fmaxf(py, p.eps);
        const float rx = rsqrtf(pxc), ry = rsqrtf(pyc);
        const float ax = pxc * rx, ay = pyc * ry;
        // pxc == pyc must give an exact zero: `ay - ax` alone is contracted into an FMA by nvcc and
        // would leave the rounding error of one product behind
        const float d = (pxc == pyc) ? 0.f : ay - ax;
        s1 = fmaf(act * d, d, s1);
        s2 = fmaf(act, pyc, s2);
        const float lr = (pxc == pyc) ? 0.f : 0.5f * fabsf(SPL_FAST_LOGF(pyc * rx * rx));
        s3 = fmaf(act, lr, s3);
        if (GRAD) {
          const float gate = (px >= p.eps) ? 1.f : 0.f;
          const float sgn = (pxc > pyc) ? 1.f : ((pxc < pyc) ? -1.f : 0.f);
          // gX = alpha * X (spectral convergence, un-scaled) and beta * X (log magnitude, un-scaled)
          const float alpha = -gate * d * rx;
          const float beta = gate * sgn * rx * rx;
          const bool self_mirror = (k == km);
          const float wgt = self_mirror ? 1.f : 0.5f;   // Hermitian extension halves interior bins
          const float gr = wgt * alpha, gi = wgt * beta;
          // H[k] = (gr + i gi) * X ,  H[N-k] = (gr + i gi) * conj(X)
          const float2 ha = make_float2(fmaf(gr, xr, -gi * xi), fmaf(gr, xi, gi * xr));
          const float2 hb = make_float2(fmaf(gr, xr, gi * xi), fmaf(gi, xr, -gr * xi));
          buf[k] = ha;
          if (!self_mirror) buf[km] = hb;
        }
      }
    } else {
      // ---- C'. mel: amplitudes -> banded projection -> log-mel L1 -> gradient spectrum ---------
      // pass 1: X stays in buf[k]; (Ax, Ay) parked in buf[N-k]; bin 0 parks in buf[N]; raw Z[N/2] in buf[N+1]
#pragma unroll 4
      for (int i = 0; i <= NPAIR; ++i) {
        const bool extra = (i == NPAIR);
        if (extra && l != 0) break;
        const int k = extra ? HALF : l + L * i;
        const int km = (NFFT - k) & (NFFT - 1);
        const float2 a = buf[k], bm = buf[km];
        const float xr = 0.5f * (a.x + bm.x), xi = 0.5f * (a.y - bm.y);
        const float yr = frame_equal ? xr : 0.5f * (a.y + bm.y);
        const float yi = frame_equal ? xi : 0.5f * (bm.x - a.x);
        const float pxc = fmaxf(fmaf(xr, xr, xi * xi), p.eps), pyc = fmaxf(fmaf(yr, yr, yi * yi), p.eps);
        const float2 amp = make_float2(pxc * rsqrtf(pxc), pyc * rsqrtf(pyc));
        if (k == 0) { buf[NFFT] = amp; buf[0] = make_float2(xr, xi); }
        else if (extra) { buf[NFFT + 1] = make_float2(xr, xi); buf[HALF] = amp; }
        else { buf[k] = make_float2(xr, xi); buf[km] = amp; }
      }
      __syncwarp();
      // pass 2: one lane per mel row
      for (int mrow = l; mrow < p.n_mels; mrow += L) {
        const int k0 = __ldg(&p.mel_row_start[mrow]), len = __ldg(&p.mel_row_len[mrow]);
        const float* __restrict__ wv = p.mel_row_val + __ldg(&p.mel_row_ptr[mrow]);
        float mx = 0.f, my = 0.f;
        for (int s = 0; s < len; ++s) {
          const int k = k0 + s;
          const float2 amp = buf[k == 0 ? NFFT : NFFT - k];
          const float w = __ldg(&wv[s]);
          mx = fmaf(amp.x, w, mx);
          my = fmaf(amp.y, w, my);
        }
        const float mxc = fmaxf(mx, p.eps), myc = fmaxf(my, p.eps);
        const float dl = (logf(mxc) - logf(myc)) * p.inv_ln_base;
        s1 = fmaf(act, fabsf(dl), s1);
        if (GRAD) {
          const float sgn = (dl > 0.f) ? 1.f : ((dl < 0.f) ? -1.f : 0.f);
          gm_s[mrow] = (mx >= p.eps) ? sgn * p.inv_ln_base / mxc : 0.f;
        }
      }
      __syncwarp();
      if (GRAD) {
        // pass 3: gA[k] = sum_m gM[m] W[k,m] (<= 2 terms), H[k] = 1/2 gA gate / Ax * X
#pragma unroll 4
        for (int i = 0; i <= NPAIR; ++i) {
          const bool extra = (i == NPAIR);
          if (extra && l != 0) break;
          const int k = extra ? HALF : l + L * i;
          const int km = (NFFT - k) & (NFFT - 1);
          const float2 xk = extra ? buf[NFFT + 1] : buf[k];
          const float2 amp = (k == 0) ? buf[NFFT] : buf[km];
          const int m0 = __ldg(&p.bin_m0[k]);
          const float ga = fmaf(gm_s[m0], __ldg(&p.bin_w0[k]), gm_s[m0 + 1] * __ldg(&p.bin_w1[k]));
          const float px = fmaf(xk.x, xk.x, xk.y * xk.y);
          const bool self_mirror = (k == km);
          const float g = (px >= p.eps) ? (self_mirror ? 1.f : 0.5f) * ga / amp.x : 0.f;
          buf[k] = make_float2(g * xk.x, g * xk.y);
          if (!self_mirror) buf[km] = make_float2(g * xk.x, -g * xk.y);
        }
      }
    }
    if (!GRAD) { __syncwarp(); continue; }
    __syncwarp();
    // ---- D. adjoint of the one-sided rFFT = inverse DFT of the Hermitian-extended H --------------
    natural_to_time<NFFT>(re, im, buf, tw, l);
    // ---- E. window, overlap-add into the ring, flush the finished `hop` samples ------------------
#pragma unroll 1
    for (int hh = 0; hh < FPW; ++hh) {
      if (h == hh && active) {
        const int base = (jc * p.hop) % p.ring_n;
#pragma unroll
        for (int n2 = 0; n2 < R; ++n2) {
          const int tap = l + L * n2 - p.left;
          if (tap >= 0 && tap < p.win) {
            const float w = __ldg(&p.window[tap]);
            int idx = base + tap;
            idx -= (idx >= p.ring_n) ? p.ring_n : 0;
            if (KIND == kKindStft) {
              float2 r = ring2[idx];
              r.x = fmaf(re[n2], w, r.x);
              r.y = fmaf(im[n2], w, r.y);
              ring2[idx] = r;
            } else {
              ring1[idx] = fmaf(re[n2], w, ring1[idx]);
            }
          }
        }
      }
      __syncwarp();
    }
    {
      const int done = min(m_c, (step + 1) * FPW);             // frames accumulated so far
      const int limit = (done == m_c) ? span_c : done * p.hop;  // positions no later frame touches
      for (int q = flushed + lane; q < limit; q += 32) {
        const int idx = q % p.ring_n;
        if (KIND == kKindStft) { out2[q] = ring2[idx]; ring2[idx] = make_float2(0.f, 0.f); }
        else                   { out1[q] = ring1[idx]; ring1[idx] = 0.f; }
      }
      flushed = limit;
      __syncwarp();
    }
  }

  // ---- partial sums of this chunk -------------------------------------------------------------
  s1 = warp_sum(s1);
  if (KIND == kKindStft) { s2 = warp_sum(s2); s3 = warp_sum(s3); }
  if (lane == 0) {
    if (KIND == kKindStft) {
      double* o = p.partials + (size_t)chunk_id * 3;
      o[0] = (double)s1; o[1] = (double)s2; o[2] = (double)s3;
    } else {
      p.partials[chunk_id] = (double)s1;
    }
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
      const double d = sqrt(s[0]), ny = sqrt(s[1]);
      sc += d / ny;
      mag += s[2] / p.count[r];
      p.coefs[2 * r] = (d > 0.0) ? (float)(1.0 / (n_stft * d * ny)) : 0.f;
      p.coefs[2 * r + 1] = (float)(1.0 / (n_stft * p.count[r]));
    } else {
      mel += s[0] / p.count[r];
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

SPL_DEVICE void combine_body(const CombineParams& p, long long gid) {
  if (gid >= (long long)p.B * p.T) return;
  const int b = (int)(gid / p.T), i = (int)(gid - (long long)b * p.T);
  const float gsc = p.g_sc ? *p.g_sc : 0.f, gmag = p.g_mag ? *p.g_mag : 0.f, gmel = p.g_mel ? *p.g_mel : 0.f;
  float acc = 0.f;
  for (int r = 0; r < p.n; ++r) {
    const CombineEntry& e = p.e[r];
    float cu, cv;
    if (e.kind == kKindStft) { cu = gsc * p.coefs[2 * r]; cv = gmag * p.coefs[2 * r + 1]; }
    else { cu = gmel * p.coefs[2 * r]; cv = 0.f; }
    const int P = e.half;
    acc += gather_padded(e, b, P + i, cu, cv);
    if (i >= 1 && i <= P) acc += gather_padded(e, b, P - i, cu, cv);
    if (i >= p.T - 1 - P && i <= p.T - 2) acc += gather_padded(e, b, P + 2 * (p.T - 1) - i, cu, cv);
  }
  p.dx[gid] = acc;
}

#ifndef SPECLOSS_EMU
template <int NFFT, int KIND, bool GRAD>
__global__ void __launch_bounds__(kWarpsPerCta * 32) transform_kernel(const TransformParams p) {
  extern __shared__ __align__(16) float smem_dyn[];
  transform_body<NFFT, KIND, GRAD>(p, smem_dyn, blockIdx.x, threadIdx.x);
}
__global__ void __launch_bounds__(256) reduce_kernel(const ReduceParams p) {
  __shared__ double sh[256];
  reduce_body(p, sh, blockIdx.x, threadIdx.x, 256);
}
__global__ void finalize_kernel(const FinalizeParams p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) finalize_body(p);
}
__global__ void __launch_bounds__(256) combine_kernel(const CombineParams p) {
  combine_body(p, (long long)blockIdx.x * 256 + threadIdx.x);
}
#endif

}  // namespace spl
