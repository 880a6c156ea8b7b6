// Power-mel L1 metric: `Mel_L1(pred, target)` of the reference's evaluation scripts
// (/root/reference/mel_spectrogram.py:36-44, duplicated at sandbox.py:183-191):
//     mel = torchaudio.transforms.MelSpectrogram(48000)        # n_fft 400, hop 200, 128 HTK mels, power 2, no log
//     Mel_L1 = nn.L1Loss()(mel(pred), mel(target))
// Forward only (an evaluation metric).  The 400-point transform is not a power of two, so this path has its own
// kernel: N = 400 = 25 x 16 as a two-pass Cooley-Tukey with the generated in-register codelets fft16 / fft25,
//   pass A: lane n1 < 25 holds samples n = n1 + 25 n2 (n2 < 16): fft16 over n2, twiddle W_400^(n1 k2), column n1 of the slot
//   pass B: 16 lanes per frame (two frames per warp), each reads row k2 (25 values over n1): fft25 -> bins k = k2 + 16 k1,
//           written back in natural order.
// Prediction and target share ONE complex FFT (z = w (x + i y)), the pair (k, 400 - k) gives both spectra; the power
// spectra (|X|^2, |Y|^2) of bins 0..200 are parked in shared memory and every lane sums four of the 128 banded mel
// rows from a CSR table; sum |Mx - My| goes to one fp64 partial per warp.  Optionally the two (rows, n_mels, frames)
// mel spectrograms themselves are written out (parity tests compare the tensors, not only the scalar).
// Same emulation contract as specloss_kernels.cuh: *_body functions run under tests/emu on the CPU.
#pragma once

#include "specloss_kernels.cuh"

namespace spl {

constexpr int kMp = 400;           // transform length
constexpr int kMpL = 25;           // lanes of pass A / points per lane in pass B
constexpr int kMpR = 16;           // points per lane in pass A / lanes of pass B
constexpr int kMpBins = kMp / 2 + 1;
constexpr int kMpMaxMels = 256;
constexpr int kMpMaxNnz = 1024;    // non-zeros of the filterbank (HTK 128 x 201: 347)

struct MelPowParams {
  const float* x;          // prediction (rows, T)
  const float* y;          // target     (rows, T)
  int rows, T, hop, n_frames;
  const float* window;     // 400 taps
  const float2* twiddle;   // [16][25]: W_400^(n1 * k2) at [k2 * 25 + n1]
  int n_mels, nnz;
  const int* mel_ptr;      // [n_mels + 1] CSR row offsets into mel_ent
  const int2* mel_ent;     // [nnz] {bin, weight bits}
  double* partials;        // [grid * warps per CTA]
  float* mel_x;            // null, or (rows, n_mels, n_frames): MelSpectrogram(pred)
  float* mel_y;            // null, or the same for the target
};

// shared memory (4-byte words): CTA tables, then per warp two frame slots (400 float2 each) and the power pairs (201 float2)
struct MelPowSmem {
  int tw, win, ptr, ent, total, per_warp;
};
static __host__ __device__ inline MelPowSmem melpow_smem(int n_mels, int nnz) {
  MelPowSmem s;
  int o = 0;
  s.tw = o;  o += 2 * kMp;
  s.win = o; o += kMp;
  s.ptr = o; o += (n_mels + 1 + 3) & ~3;
  s.ent = o; o += (2 * nnz + 3) & ~3;
  s.total = o;
  s.per_warp = 2 * (2 * kMp) + ((2 * kMpBins + 3) & ~3);          // two frame slots (a warp takes two frames at a time)
  return s;
}

SPL_DEVICE void melpow_load_tables(const MelPowParams& p, float* smem, int tid, int nthreads) {
  const MelPowSmem sm = melpow_smem(p.n_mels, p.nnz);
  cta_copy_words(smem + sm.tw, p.twiddle, 2 * kMp, tid, nthreads);
  cta_copy_words(smem + sm.win, p.window, kMp, tid, nthreads);
  cta_copy_words(smem + sm.ptr, p.mel_ptr, p.n_mels + 1, tid, nthreads);
  cta_copy_words(smem + sm.ent, p.mel_ent, 2 * p.nnz, tid, nthreads);
  cp_async_wait_all();
}

// [region: melpow]
// One warp takes TWO frames at a time (consecutive frames of a row: their taps overlap by half, the second read of every tap
// comes from L1): pass A runs once per frame on 25 lanes, pass B ONCE for both -- lanes 0-15 transform the 16 rows of the first
// frame, lanes 16-31 those of the second (round 2: the 25-point pass used to run per frame on half a warp) -- then power
// spectra, projection and L1 per frame on all 32 lanes.
SPL_DEVICE void melpow_body(const MelPowParams& p, float* smem, int block, int tid, int grid, int wpc) {
  const int warp = tid >> 5, lane = tid & 31;
  const MelPowSmem sm = melpow_smem(p.n_mels, p.nnz);
  const float2* tw = reinterpret_cast<const float2*>(smem + sm.tw);
  const float* wtab = smem + sm.win;
  const int* mptr = reinterpret_cast<const int*>(smem + sm.ptr);
  const int2* ment = reinterpret_cast<const int2*>(smem + sm.ent);
  float2* S0 = reinterpret_cast<float2*>(smem + sm.total + (size_t)warp * sm.per_warp);   // 2 x ([16][25], then natural order [400])
  float2* P = S0 + 2 * kMp;                                                                // [201] (|X|^2, |Y|^2)
  const long long total = (long long)p.rows * p.n_frames;
  const long long pairs = (total + 1) / 2;
  double acc = 0.0;
  for (long long pair = (long long)block * wpc + warp; pair < pairs; pair += (long long)grid * wpc) {
    bool eq0 = true, eq1 = true;
    // pass A: 25 lanes, 16 points each, one frame after the other
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
      const long long item = 2 * pair + f;
      const bool live = item < total;
      const int row = live ? (int)(item / p.n_frames) : 0, t = live ? (int)(item - (long long)row * p.n_frames) : 0;
      const float* __restrict__ xb = p.x + (size_t)row * p.T;
      const float* __restrict__ yb = p.y + (size_t)row * p.T;
      const int s0 = t * p.hop - kMp / 2;                    // reflect padding of n_fft / 2 on both sides (center=True)
      float2* S = S0 + f * kMp;
      bool same = true;
      if (live && lane < kMpL) {
        float2 v[kMpR];
        const bool interior = s0 >= 0 && s0 + kMp <= p.T;
#pragma unroll
        for (int n2 = 0; n2 < kMpR; ++n2) {
          const int n = lane + kMpL * n2;
          const int sidx = interior ? s0 + n : reflect(s0 + n, p.T);
          const float w = wtab[n];
          const float xv = __ldg(xb + sidx) * w, yv = __ldg(yb + sidx) * w;
          same = same && (xv == yv);
          v[n2] = make_float2(xv, yv);
        }
        fft16(v);
#pragma unroll
        for (int k2 = 0; k2 < kMpR; ++k2) S[k2 * kMpL + lane] = k2 > 0 ? cmul(v[k2], tw[k2 * kMpL + lane]) : v[0];
      }
      // identical prediction / target frames must give exactly |Mx - My| = 0 (the packed FFT leaves ~1e-7 of asymmetry)
      const bool e = __ballot_sync(0xffffffffu, same) == 0xffffffffu;
      if (f == 0) eq0 = e; else eq1 = e;
    }
    __syncwarp();
    // pass B for both frames at once: lane & 15 = row k2 of frame lane >> 4 -> Z[k2 + 16 k1], back in natural order
    {
      float2* S = S0 + (lane >> 4) * kMp;
      const int k2 = lane & (kMpR - 1);
      float2 b[kMpL];
#pragma unroll
      for (int n1 = 0; n1 < kMpL; ++n1) b[n1] = S[k2 * kMpL + n1];
      fft25(b);
      __syncwarp();                                          // every row is read before the slots are overwritten
#pragma unroll
      for (int k1 = 0; k1 < kMpL; ++k1) S[k2 + kMpR * k1] = b[k1];
      __syncwarp();
    }
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
      const long long item = 2 * pair + f;
      if (item >= total) break;                              // warp-uniform
      const int row = (int)(item / p.n_frames), t = (int)(item - (long long)row * p.n_frames);
      const float2* S = S0 + f * kMp;
      // power spectra of bins 0..200 from the mirror pairs: 2X = Z[k] + conj Z[N-k], 2Y = -i (Z[k] - conj Z[N-k])
      for (int k = lane; k < kMpBins; k += 32) {
        const float2 a = S[k], bm = S[k == 0 ? 0 : kMp - k];
        const float2 x2 = __fadd2_rn(a, make_float2(bm.x, -bm.y));
        float2 y2 = __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));
        y2 = (f == 0 ? eq0 : eq1) ? x2 : y2;
        P[k] = make_float2(0.25f * fmaf(x2.x, x2.x, x2.y * x2.y), 0.25f * fmaf(y2.x, y2.x, y2.y * y2.y));
      }
      __syncwarp();
      // banded mel projection + L1
      float s = 0.f;
      for (int m = lane; m < p.n_mels; m += 32) {
        float mx = 0.f, my = 0.f;
        for (int e = mptr[m]; e < mptr[m + 1]; ++e) {
          const int2 en = ment[e];
          const float2 pw = P[en.x];
          const float w = bits_to_float(en.y);
          mx = fmaf(pw.x, w, mx);
          my = fmaf(pw.y, w, my);
        }
        s += fabsf(mx - my);
        if (p.mel_x) p.mel_x[((size_t)row * p.n_mels + m) * p.n_frames + t] = mx;
        if (p.mel_y) p.mel_y[((size_t)row * p.n_mels + m) * p.n_frames + t] = my;
      }
      acc += (double)s;
      __syncwarp();                                          // P is reused by the next frame, the slots by the next pair
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) p.partials[block * wpc + warp] = acc;
}

#ifndef SPECLOSS_EMU
__global__ void __launch_bounds__(256) melpow_kernel(const MelPowParams p) {
  extern __shared__ __align__(16) float smem_dyn[];
  melpow_load_tables(p, smem_dyn, threadIdx.x, blockDim.x);
  __syncthreads();
  melpow_body(p, smem_dyn, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
#endif

}  // namespace spl
