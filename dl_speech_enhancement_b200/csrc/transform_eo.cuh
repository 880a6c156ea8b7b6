// 2048-point loss kernels on the 32 x 32 geometry ("even/odd" route).
//
// The 64-point-per-lane kernels of specloss_kernels.cuh (2048 = 32 lanes x 64 points, prediction and target packed into one
// complex FFT) are ~100 KB of straight-line SASS and stall on instruction fetch (profiles/r2p: issue active 30-39 %,
// no_instruction 1.35-1.76 per issue).  Here every REAL 2048-point transform is ONE 1024-point complex FFT of
// z[m] = s[2m] + i s[2m+1] plus a twiddle pass on the mirror pairs (k, 1024 - k):
//     2E = Z[k] + conj Z[M-k],  2O = -i (Z[k] - conj Z[M-k]),  2X[k] = 2E + W^k 2O,  2X[M-k] = conj(2E - W^k 2O),  W = e^{-2 pi i/2048}
// (the pair k = 0 yields bins 0 and 1024, k = 512 is its own mirror: X[512] = conj Z[512]) -- the same 32-point codelets,
// slot layout and mirror exchange as the 1024-point kernel, run once per signal by a ROLLED two-iteration loop, so that
// prediction and target go through the very same instructions (bit-identical inputs give bit-identical spectra: the
// exact-zero property of loss(x, x) needs no special case).  The unpacked spectra of both signals wait in registers
// (2 x 32 complex per lane), so a warp needs ONE 8.4 KB frame slot instead of 16.9 KB: 12 resident warps per SM also for
// the mel kernel, whose tables limited the 64-point-per-lane version to 8.
// The adjoint runs the mirrored pre-pass: with H the Hermitian gradient spectrum (H[j] = G[j]/2 inside, G at j = 0, 1024),
//     S = H[k] + conj H[M-k],  D = H[k] - conj H[M-k],  T = i conj(W^k) D,  C[k] = S + T,  C[M-k] = conj(S - T),
// and ONE 1024-point complex inverse returns c[m] = g[2m] + i g[2m+1].  STFT needs two real gradients (spectral
// convergence u, log magnitude v): two inverses, stored as two planes per frame; mel needs one -- 3 half-size FFTs per
// frame instead of 2 full-size ones (-32 % flops).  Math checked against numpy in gen/notes: DESIGN.md section 4.
#pragma once

#include "specloss_kernels.cuh"

namespace spl {

// tuning knobs (profiles/README.md r3h): resident warps per SM (register budget = 65536 / (32 * warps)) and whether the
// two-signal loop is rolled
#ifndef SPL_EO_WARPS
#define SPL_EO_WARPS 12
#endif
#ifndef SPL_EO_UNROLL
#define SPL_EO_UNROLL 2          // measured (profiles/README.md r3h): rolled = loop-carried copies of 64 registers + spills
#endif
#ifndef SPL_EO_UNROLL_Q
#define SPL_EO_UNROLL_Q 1
#endif

namespace eo {
using G = Geo<1024>;
constexpr int M = 1024, L = 32, R = 32, HL = 16, PITCH = G::PITCH, SLOT = G::SLOT_F2, HALF = 1024;
constexpr int UNROLL = SPL_EO_UNROLL, UNROLL_Q = SPL_EO_UNROLL_Q;
constexpr int WK = 516;            // W_2048^k for k = 0 .. 512, padded to a multiple of 4 float2
}  // namespace eo

struct CtaTablesEo {
  int tw, wk, win, tasks, entries, bintab, total;     // word offsets
};
static __host__ __device__ inline CtaTablesEo cta_tables_eo(int win, int kind, int mel_rounds, int mel_entry_rows) {
  CtaTablesEo t;
  int o = 0;
  t.tw = o;      o += (2 * eo::R * eo::PITCH + 3) & ~3;
  t.wk = o;      o += 2 * eo::WK;
  t.win = o;     o += (win + 3) & ~3;
  t.tasks = o;   o += kind == kKindMel ? 4 * mel_rounds * eo::L : 0;
  t.entries = o; o += kind == kKindMel ? ((2 * mel_entry_rows * eo::L + 3) & ~3) : 0;
  t.bintab = o;  o += kind == kKindMel ? 4 * (2048 / 2 + 1) : 0;
  t.total = o;
  return t;
}

static __host__ __device__ inline int eo_words_per_warp(int kind, int n_mels) {
  int w = eo::SLOT * 2;                                           // ONE frame slot: the spectra wait in registers
  if (kind == kKindMel) w += 2 * ((n_mels + 3) & ~3);
  return (w + 3) & ~3;
}

// twiddle_eo (global): [2 * 1024 floats: W_1024^(n1 k2) at [k2 * 32 + n1]] [2 * 516 floats: W_2048^k, k <= 512, zero padded]
template <int KIND>
SPL_DEVICE void cta_load_tables_eo(const TransformParams& p, const float2* twiddle_eo, const void* mel_entries_eo, float* smem,
                                   int tid, int nthreads) {
  const CtaTablesEo ct = cta_tables_eo(p.win, KIND, p.mel_rounds, p.mel_entry_rows);
  float2* tw = reinterpret_cast<float2*>(smem + ct.tw);
  for (int i = tid; i < eo::M; i += nthreads) cp_async8(&tw[(i / eo::L) * eo::PITCH + (i % eo::L)], &twiddle_eo[i]);
  cta_copy_words(smem + ct.wk, twiddle_eo + eo::M, 2 * eo::WK, tid, nthreads);
  cta_copy_words(smem + ct.win, p.window, p.win, tid, nthreads);
  if (KIND == kKindMel) {
    cta_copy_words(smem + ct.tasks, p.mel_tasks, 4 * p.mel_rounds * eo::L, tid, nthreads);
    cta_copy_words(smem + ct.entries, mel_entries_eo, 2 * p.mel_entry_rows * eo::L, tid, nthreads);
    cta_copy_words(smem + ct.bintab, p.bin_tab, 4 * (2048 / 2 + 1), tid, nthreads);
  }
  cp_async_wait_all();
}

// [region: eo tap load]
// z[m] = w[2m] s[2m] + i w[2m+1] s[2m+1] for m = l + 32 n2 -> v[n2]; frame sample j = 2m sits at sb[rho(s0 + j)].
template <int WIN_T>
SPL_DEVICE void eo_load_taps(float2 (&v)[eo::R], const float* __restrict__ sb, int T, int s0, int win, int left,
                             const float* wtab, int l) {
  const bool interior = (s0 + left >= 0) && (s0 + left + win <= T);
  // 8-byte loads: the frame start, the window offset and the window length must keep the (even, odd) pairs aligned
  const bool vec = interior && ((reinterpret_cast<uintptr_t>(sb + s0) & 7) == 0) && (((left | win) & 1) == 0);
  if (vec) {
    const float2* __restrict__ sp = reinterpret_cast<const float2*>(sb + s0) + l;
    const float2* wp = reinterpret_cast<const float2*>(wtab + (2 * l - left));          // left even: 8-byte aligned
#pragma unroll
    for (int n2 = 0; n2 < eo::R; ++n2) {
      const int lo = 64 * n2 - left;                         // tap of lane 0's even sample
      if (WIN_T > 0 && (lo + 63 < 0 || lo >= WIN_T)) { v[n2] = make_float2(0.f, 0.f); continue; }
      const bool all_lanes = WIN_T > 0 && lo >= 0 && lo + 63 < WIN_T;
      float2 z = make_float2(0.f, 0.f);
      if (all_lanes || (lo + 2 * l >= 0 && lo + 2 * l < win)) z = __fmul2_rn(__ldg(sp + 32 * n2), wp[32 * n2]);
      v[n2] = z;
    }
  } else {
#pragma unroll
    for (int n2 = 0; n2 < eo::R; ++n2) {
      const int lo = 64 * n2 - left;
      if (WIN_T > 0 && (lo + 63 < 0 || lo >= WIN_T)) { v[n2] = make_float2(0.f, 0.f); continue; }
      const int tap = lo + 2 * l, j = 64 * n2 + 2 * l;
      float a = 0.f, b = 0.f;
      if (tap >= 0 && tap < win) a = __ldg(&sb[reflect(s0 + j, T)]) * wtab[tap];
      if (tap + 1 >= 0 && tap + 1 < win) b = __ldg(&sb[reflect(s0 + j + 1, T)]) * wtab[tap + 1];
      v[n2] = make_float2(a, b);
    }
  }
}

// doubled spectra of bins k and M - k from the mirror pair of Z (a = Z[k], bm = Z[M-k]) and wk = W_2048^k
SPL_DEVICE void eo_unpack(float2 a, float2 bm, float2 wk, float2& xa, float2& xb) {
  const float2 e2 = __fadd2_rn(a, make_float2(bm.x, -bm.y));
  const float2 o2 = __fadd2_rn(make_float2(a.y, -a.x), make_float2(bm.y, bm.x));
  const float2 t = cmul(o2, wk);
  xa = __fadd2_rn(e2, t);
  xb = make_float2(e2.x - t.x, t.y - e2.y);
}

// adjoint pre-pass: C[k], C[M-k] from H[k] = ha, H[M-k] = hb
SPL_DEVICE void eo_pack(float2 ha, float2 hb, float2 wk, float2& ca, float2& cb) {
  const float2 s = make_float2(ha.x + hb.x, ha.y - hb.y);
  const float2 d = make_float2(ha.x - hb.x, ha.y + hb.y);
  // T = i conj(wk) d
  const float2 t = make_float2(fmaf(wk.y, d.x, -wk.x * d.y), fmaf(wk.x, d.x, wk.y * d.y));
  ca = __fadd2_rn(s, t);
  cb = make_float2(s.x - t.x, t.y - s.y);
}

// STFT loss terms of ONE bin from the doubled spectra x2 = 2 X^[j], y2 = 2 Y[j] (same arithmetic as stft_pair), and -- GRAD
// -- the two un-scaled gradient spectra H_u[j] = hu, H_v[j] = hv (wq = 1/4 inside, 1/2 at j = 0 and j = N/2).
template <bool GRAD>
SPL_DEVICE void eo_stft_bin(float2 x2, float2 y2, float wq, float eps4, float& s1, float& s2, float& s3, float2& hu, float2& hv) {
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
    const float rxg = px >= eps4 ? rx : 0.f;                          // clamp gate of the reference
    const float sgn = fminf(fmaxf((pxc - pyc) * 1e25f, -1.f), 1.f);   // exact sign, 0 when equal
    const float gr = -wq * d * rxg;
    const float gi = 4.f * wq * sgn * rx * rxg;
    hu = __fmul2_rn(x2, make_float2(gr, gr));
    hv = __fmul2_rn(x2, make_float2(gi, gi));
  }
}

// [region: eo stft epilogue]
// In: the doubled spectra of this lane's 16 mirror pairs in registers -- XA[k1] = 2 X^[k], XB[k1] = 2 X^[M-k], YA / YB the
// same for the target, k = l + 32 k1 -- and (lane 0) xh / yh = 2 X^[512], 2 Y[512].  The frame slot is free.
// Out (GRAD): C_u as the inverse transform expects it (columns < 16 in A, mirror halves + bin 512 in the slot);
// C_v waits in XA / XB / xh for the second inverse.
template <bool GRAD>
SPL_DEVICE void eo_stft_epilogue(float2 (&A)[1][eo::HL], float2 (&XA)[eo::HL], float2 (&XB)[eo::HL], const float2 (&YA)[eo::HL],
                                 const float2 (&YB)[eo::HL], float2& xh, float2 yh, float2* S, const float2* wkt, int l,
                                 float eps4, float& s1, float& s2, float& s3) {
  float2* pb = mirror_ptr<1024>(S, l);
#pragma unroll
  for (int k1 = 0; k1 < eo::HL; ++k1) {
    const float wq = (k1 == 0 && l == 0) ? 0.5f : 0.25f;     // the pair k = 0 holds bins 0 and 1024 (weight 1, not 1/2)
    float2 hua, hva, hub, hvb;
    eo_stft_bin<GRAD>(XA[k1], YA[k1], wq, eps4, s1, s2, s3, hua, hva);
    eo_stft_bin<GRAD>(XB[k1], YB[k1], wq, eps4, s1, s2, s3, hub, hvb);
    if (GRAD) {
      const float2 wk = wkt[l + eo::R * k1];
      float2 ca, cb;
      eo_pack(hua, hub, wk, ca, cb);
      A[0][k1] = ca;
      pb[-k1] = cb;
      eo_pack(hva, hvb, wk, XA[k1], XB[k1]);
    }
  }
  if (l == 0) {                                      // k = 512 = (row 0, column 16): C[512] = 2 conj H[512]
    float2 hu, hv;
    eo_stft_bin<GRAD>(xh, yh, 0.25f, eps4, s1, s2, s3, hu, hv);
    if (GRAD) {
      S[eo::HL] = make_float2(2.f * hu.x, -2.f * hu.y);
      xh = make_float2(2.f * hv.x, -2.f * hv.y);
    }
  }
}

// [region: eo mel epilogue]
// In: as above.  The amplitude pair (Ax, Ay) of bin j <= 1024 is parked at the natural position of j in the slot: row
// j % 32, column j / 32 (bin 1024: the pad word of row 0).  Out (GRAD): C in A / the slot.
template <bool GRAD>
SPL_DEVICE void eo_mel_epilogue(float2 (&A)[1][eo::HL], const float2 (&XA)[eo::HL], const float2 (&XB)[eo::HL],
                                const float2 (&YA)[eo::HL], const float2 (&YB)[eo::HL], float2 xh, float2 yh, float2* S,
                                float2* msum, const float2* wkt, int l, const TransformParams& p, const int4* mel_tasks,
                                const int2* mel_entries, const int4* bin_tab, float& s1) {
  const float eps4 = 4.f * p.eps;
  float2* pb = mirror_ptr<1024>(S, l);
  float2* arow = S + l * eo::PITCH;
  // pass 1: amplitudes
#pragma unroll
  for (int k1 = 0; k1 < eo::HL; ++k1) {
    const float2 xa = XA[k1], xb = XB[k1], ya = YA[k1], yb = YB[k1];
    const float pxa = fmaxf(fmaf(xa.x, xa.x, xa.y * xa.y), eps4), pya = fmaxf(fmaf(ya.x, ya.x, ya.y * ya.y), eps4);
    const float pxb = fmaxf(fmaf(xb.x, xb.x, xb.y * xb.y), eps4), pyb = fmaxf(fmaf(yb.x, yb.x, yb.y * yb.y), eps4);
    arow[k1] = __fmul2_rn(make_float2(0.5f * pxa, 0.5f * pya), make_float2(spl_fast_rsqrt(pxa), spl_fast_rsqrt(pya)));
    pb[-k1] = __fmul2_rn(make_float2(0.5f * pxb, 0.5f * pyb), make_float2(spl_fast_rsqrt(pxb), spl_fast_rsqrt(pyb)));
  }
  if (l == 0) {
    const float px = fmaxf(fmaf(xh.x, xh.x, xh.y * xh.y), eps4), py = fmaxf(fmaf(yh.x, yh.x, yh.y * yh.y), eps4);
    S[eo::HL] = __fmul2_rn(make_float2(0.5f * px, 0.5f * py), make_float2(spl_fast_rsqrt(px), spl_fast_rsqrt(py)));
  }
  __syncwarp();
  mel_project_pairs<eo::L>(S, msum, l, p.mel_rounds, mel_tasks, mel_entries);
  for (int row = l; row < p.n_mels; row += eo::L) {
    const float2 mm = msum[row];
    const float mxc = fmaxf(mm.x, p.eps), myc = fmaxf(mm.y, p.eps);
    const float dl = (mxc == myc) ? 0.f : (logf(mxc) - logf(myc)) * p.inv_ln_base;
    s1 += fabsf(dl);
    if (GRAD) {
      const float sgn = (dl > 0.f) ? 1.f : ((dl < 0.f) ? -1.f : 0.f);
      msum[row].x = (mm.x >= p.eps) ? sgn * p.inv_ln_base / mxc : 0.f;     // gM[row]
    }
  }
  __syncwarp();
  if (GRAD) {
    // pass 3: H[j] = w gA[j] gate / Ax * X^[j] for j = k and j = M - k, then the adjoint pre-pass
#pragma unroll
    for (int k1 = 0; k1 < eo::HL; ++k1) {
      const int k = l + eo::R * k1;
      const bool self = k1 == 0 && l == 0;
      const float2 ha = mel_bin_grad(XA[k1], self, bin_tab[k], msum, eps4, true);
      const float2 hb = mel_bin_grad(XB[k1], self, bin_tab[eo::M - k], msum, eps4, true);
      float2 ca, cb;
      eo_pack(ha, hb, wkt[k], ca, cb);
      A[0][k1] = ca;
      pb[-k1] = cb;
    }
    if (l == 0) {
      const float2 h = mel_bin_grad(xh, false, bin_tab[eo::M / 2], msum, eps4, true);
      S[eo::HL] = make_float2(2.f * h.x, -2.f * h.y);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// The 2048-point transform kernel body on the 32 x 32 geometry: one warp = one frame at a time, ONE frame slot per warp
// (the spectra of both signals wait in registers, 2 x 32 complex values per lane).
// gframes: [B * n_frames][NQ][win] floats, NQ = 2 planes (u, v) for STFT, 1 for mel.
// ---------------------------------------------------------------------------------------------
template <int KIND, bool GRAD, int WIN_T>
SPL_DEVICE void transform_eo_body(const TransformParams& p, float* smem, int block, int tid, int grid, int wpc) {
  constexpr int NQ = KIND == kKindStft ? 2 : 1;
  const int warp = tid >> 5, l = tid & 31;
  const int win = WIN_T > 0 ? WIN_T : p.win;
  const int left = WIN_T > 0 ? (2048 - WIN_T) / 2 : p.left;

  const CtaTablesEo ct = cta_tables_eo(p.win, KIND, p.mel_rounds, p.mel_entry_rows);
  const float2* tw = reinterpret_cast<const float2*>(smem + ct.tw);
  const float2* wkt = reinterpret_cast<const float2*>(smem + ct.wk);
  const float* wtab = smem + ct.win;
  const int4* mel_tasks = reinterpret_cast<const int4*>(smem + ct.tasks);
  const int2* mel_entries = reinterpret_cast<const int2*>(smem + ct.entries);
  const int4* bin_tab = reinterpret_cast<const int4*>(smem + ct.bintab);

  float* wsm = smem + ct.total + (size_t)warp * eo_words_per_warp(KIND, p.n_mels);
  float2* S = reinterpret_cast<float2*>(wsm);
  float2* msum = reinterpret_cast<float2*>(wsm + eo::SLOT * 2);

  const int total = p.B * p.n_frames;
  double d1 = 0.0, d2 = 0.0, d3 = 0.0;
  const bool vec_store = ((left | win) & 1) == 0;
  float2* pb = mirror_ptr<1024>(S, l);

  // [region: eo frame loop]
  for (int item = block * wpc + warp; item < total; item += grid * wpc) {
    const int b = item / p.n_frames, t = item - b * p.n_frames;
    const int s0 = t * p.hop - eo::HALF;
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    float2 XA[eo::HL], XB[eo::HL], YA[eo::HL], YB[eo::HL];
    float2 xh = make_float2(0.f, 0.f), yh = make_float2(0.f, 0.f);       // 2 X^[512], 2 Y[512] (lane 0)
#pragma unroll
    for (int k1 = 0; k1 < eo::HL; ++k1) YA[k1] = YB[k1] = make_float2(0.f, 0.f);
    float2 A[1][eo::HL];
    // forward + unpack: the same instructions for both signals (rolled on purpose)
#pragma unroll(eo::UNROLL)
    for (int s = 0; s < 2; ++s) {
#pragma unroll
      for (int k1 = 0; k1 < eo::HL; ++k1) { XA[k1] = YA[k1]; XB[k1] = YB[k1]; }
      xh = yh;
      const float* sb = (s == 0 ? p.x : p.y) + (size_t)b * p.T;
      {
        float2 v[eo::R];
        eo_load_taps<WIN_T>(v, sb, p.T, s0, win, left, wtab, l);
        Dft<32>::run(v);
        fwd_store_cols<1024>(v, S, tw, l);
      }
      fwd_pass_b<1024>(A, S, l);
#pragma unroll
      for (int k1 = 0; k1 < eo::HL; ++k1) {
        float2 zm = pb[-k1];
        if (k1 == 0) zm = (l == 0) ? A[0][0] : zm;           // k = 0 (row 0) mirrors itself: it yields bins 0 and 1024
        eo_unpack(A[0][k1], zm, wkt[l + eo::R * k1], YA[k1], YB[k1]);
      }
      if (l == 0) { const float2 z = S[eo::HL]; yh = make_float2(2.f * z.x, -2.f * z.y); }      // X[512] = conj Z[512]
      __syncwarp();          // mirror halves read: the slot is free for the next transform / the epilogue
    }
    if (KIND == kKindStft) {
      eo_stft_epilogue<GRAD>(A, XA, XB, YA, YB, xh, yh, S, wkt, l, 4.f * p.eps, s1, s2, s3);
      __syncwarp();          // C_u complete in the slot before the rows are read
    } else {
      eo_mel_epilogue<GRAD>(A, XA, XB, YA, YB, xh, yh, S, msum, wkt, l, p, mel_tasks, mel_entries, bin_tab, s1);
    }
    d1 += (double)s1; d2 += (double)s2; d3 += (double)s3;
    if (GRAD) {
      float* out = reinterpret_cast<float*>(p.gframes) + (size_t)item * NQ * win;
#pragma unroll(eo::UNROLL_Q)
      for (int q = 0; q < NQ; ++q) {
        inv_pass_b<1024>(A, S, tw, l);
        float2 v[eo::R];
#pragma unroll
        for (int m2 = 0; m2 < eo::R; ++m2) v[m2] = S[m2 * eo::PITCH + l];
        __syncwarp();        // column reads done: the slot can take C_v / the next frame
        if (NQ == 2 && q == 0) {
          // C_v -> where the inverse transform expects it
#pragma unroll
          for (int k1 = 0; k1 < eo::HL; ++k1) { A[0][k1] = XA[k1]; pb[-k1] = XB[k1]; }
          if (l == 0) S[eo::HL] = xh;
        }
        Dft<32>::run(v);
        // [region: eo window + store]
        // v[n2] = c[m] with swapped components, m = l + 32 n2: (.y, .x) = (g[2m], g[2m+1])
        float* oq = out + q * win;
#pragma unroll
        for (int n2 = 0; n2 < eo::R; ++n2) {
          const int lo = 64 * n2 - left;
          if (WIN_T > 0 && (lo + 63 < 0 || lo >= WIN_T)) continue;
          const bool all_lanes = WIN_T > 0 && lo >= 0 && lo + 63 < WIN_T;
          const int tap = lo + 2 * l;
          if (vec_store) {
            if (all_lanes || (tap >= 0 && tap < win)) {
              const float2 w = *reinterpret_cast<const float2*>(wtab + tap);
              *reinterpret_cast<float2*>(oq + tap) = __fmul2_rn(make_float2(v[n2].y, v[n2].x), w);
            }
          } else {
            if (tap >= 0 && tap < win) oq[tap] = v[n2].y * wtab[tap];
            if (tap + 1 >= 0 && tap + 1 < win) oq[tap + 1] = v[n2].x * wtab[tap + 1];
          }
        }
        __syncwarp();        // (q == 0) C_v published before its rows are read
      }
    }
  }

  d1 = warp_sum(d1);
  if (KIND == kKindStft) { d2 = warp_sum(d2); d3 = warp_sum(d3); }
  if (l == 0) {
    const int prow = block * wpc + warp;
    if (KIND == kKindStft) {
      double* o = p.partials + (size_t)prow * 3;
      o[0] = 0.25 * d1; o[1] = 0.25 * d2; o[2] = 0.34657359027997264 * d3;     // 0.5 ln 2
    } else {
      p.partials[prow] = d1;
    }
  }
}

#ifndef SPECLOSS_EMU
template <int KIND, bool GRAD, int WIN_T>
__global__ void __launch_bounds__(SPL_EO_WARPS * 32, 1)
transform_eo_kernel(const TransformParams p, const float2* twiddle_eo, const void* mel_entries_eo) {
  extern __shared__ __align__(16) float smem_dyn[];
  cta_load_tables_eo<KIND>(p, twiddle_eo, mel_entries_eo, smem_dyn, threadIdx.x, blockDim.x);
  __syncthreads();
  transform_eo_body<KIND, GRAD, WIN_T>(p, smem_dyn, blockIdx.x, threadIdx.x, gridDim.x, blockDim.x >> 5);
}
#endif

}  // namespace spl
