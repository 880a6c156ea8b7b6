/* specloss.h -- C ABI of the B200 (sm_100a) spectral-loss library (libspecloss.so).
 *
 * The reference (s194584/dl-speech-enhancement) has no native layer and no FFI: its boundary
 * for this path is the Python nn.Module contract
 *     MultiResolutionSTFTLoss.forward(x, y) -> (sc_loss, mag_loss)   losses/stft_loss.py:146-170
 *     MultiMelSpectrogramLoss.forward(y_hat, y) -> mel_loss          losses/mel_loss.py:140-156
 * called from trainer/trainerGAN.py:220,227 and train_denoise.py:139.  The entry points below are
 * what a ctypes binding behind those two forward() methods (and their autograd backward) needs;
 * INTEGRATION.md shows that binding.  Each function cites the reference code it replaces.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller; the library
 *     never allocates, frees or synchronises, and launches on the stream it is given;
 *   - return value 0 = ok, negative = error (SPL_E_*); spl_last_error() gives the message of the
 *     last failing call on the calling thread;
 *   - fp32 data, fp64 partial sums; n_fft in {512, 1024, 2048}; hop <= win <= n_fft.
 */
#ifndef SPECLOSS_H
#define SPECLOSS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPL_ABI_VERSION 15

#define SPL_OK 0
#define SPL_E_INVALID (-1)   /* bad argument / outside the supported envelope */
#define SPL_E_CUDA (-2)      /* CUDA runtime error (message carries cudaGetErrorString) */

#define SPL_KIND_STFT 0      /* spectral convergence + log-magnitude L1 (stft_loss.py:80-117) */
#define SPL_KIND_MEL 1       /* log-mel L1 (mel_loss.py:74-94,153)                            */
#define SPL_MAX_TRANSFORMS 8

/* One resolution of one loss.  Plain data; build it once per module and reuse it. */
typedef struct spl_transform {
  int32_t kind;              /* SPL_KIND_* */
  int32_t n_fft, hop, win;   /* torch.stft(x, n_fft, hop, win, window), center=True, reflect pad */
  float eps;                 /* clamp on |X|^2: 1e-7 (stft_loss.py:19) / 1e-10 (mel_loss.py:35); mel reuses it on the mel energies */
  const float* window;       /* device, `win` taps (the module's registered buffer) */
  const float* twiddle;      /* device, 2*n_fft floats written by spl_fill_twiddle() */
  /* --- mel only (kind == SPL_KIND_MEL); NULL / 0 otherwise --- */
  int32_t n_mels;
  float inv_ln_base;         /* 1/ln(log_base); 1 for log_base=None (mel_loss.py:63-71) */
  const int32_t* mel_tasks;     /* device [mel_rounds * L * 4] (L = 16 for n_fft 512, else 32): per (round, lane)
                                   {row | group<<12 | iters<<20, first entry row of the round, 0, 0}; row 0xfff = idle */
  const int32_t* mel_entries;   /* device [sum(iters) * L * 2]: per (entry row, lane) {slot offset of the bin's
                                   amplitudes inside the frame slot, bits of melmat[k, m]}; padding = {0, 0.0f} */
  int32_t mel_rounds;
  int32_t mel_entry_rows;       /* sum of iters over the rounds */
  const int32_t* bin_tab;       /* device [(n_fft/2+1) * 4]: {m0, bits(melmat[k,m0]), bits(melmat[k,m0+1]), 0} */
  /* --- n_fft == 2048 only: tables of the even/odd kernels (32 x 32 geometry, csrc/transform_eo.cuh); NULL selects the
   *     64-point-per-lane kernels --- */
  const float* twiddle_eo;      /* device, 2*1024 + 2*516 floats written by spl_fill_twiddle_eo() */
  const int32_t* mel_entries_eo;/* device, same shape as mel_entries with the amplitude slot offsets of the 32 x 32 geometry:
                                   bin j < 1024 at (j % 32) * 33 + j / 32, bin 1024 at 32 (mel only) */
  /* --- per-call workspace, sized by spl_geometry() --- */
  double* partials;          /* device [partial_count]: one row of n_sums per warp of the launch */
  void* gframes;             /* device [gframe_bytes]: the windowed, un-scaled gradient of every frame
                                ([B * n_frames][win] float2 for STFT, float for mel); NULL = forward only (torch.no_grad) */
} spl_transform;

typedef struct spl_geometry {
  int32_t n_frames;          /* 1 + T / hop */
  int32_t n_bins;            /* n_fft / 2 + 1 */
  int32_t n_sums;            /* 3 (S1, S2, S3) for STFT, 1 (S4) for mel */
  int32_t reserved;
  int64_t partial_count;     /* doubles in `partials` (upper bound over devices) */
  int64_t gframe_bytes;      /* bytes in `gframes` */
  int64_t smem_table_bytes;  /* shared memory per CTA for the constant tables (twiddle, window, mel tables) */
  int64_t smem_warp_bytes;   /* shared memory per warp (frame slot(s), mel scratch) */
} spl_geometry;

int32_t spl_abi_version(void);
const char* spl_last_error(void);

/* Host-side table: W_N^(n1*k2) rounded from fp64, layout [k2][n1] (2*n_fft floats, host memory). */
int32_t spl_fill_twiddle(int32_t n_fft, float* host_out);

/* Host-side tables of the even/odd 2048-point kernels: W_1024^(n1*k2) at [k2*32 + n1] (2*1024 floats) followed by
 * W_2048^k for k = 0..512 zero-padded to 516 entries (2*516 floats); all rounded from fp64. */
int32_t spl_fill_twiddle_eo(float* host_out);

/* Sizes of the per-call workspace for batch (B, T).  Host only. */
int32_t spl_geometry_of(const spl_transform* t, int32_t B, int32_t T, spl_geometry* out);

/* Forward of every transform in ts[0..n): replaces stft()/STFTLoss.forward (stft_loss.py:19-35,
 * 100-117) and MelSpectrogram.forward (mel_loss.py:74-94) for x and y at once.  Writes per-chunk
 * partial sums and, when gframes != NULL, the un-scaled waveform-gradient pieces of dL/dx. */
int32_t spl_forward(const spl_transform* ts, int32_t n, const float* x, const float* y,
                    int32_t B, int32_t T, void* stream);

/* Deterministic reduction of the partial sums into sums[sum of n_sums] (device doubles), in the
 * order of ts.  Multi-GPU: all-reduce `sums` (SUM) between this call and spl_finalize(). */
int32_t spl_reduce(const spl_transform* ts, int32_t n, int32_t B, int32_t T, double* sums, void* stream);

/* Losses from the sums: replaces SpectralConvergenceLoss/LogSTFTMagnitudeLoss (stft_loss.py:56,77),
 * the means over resolutions (stft_loss.py:161-168, mel_loss.py:151-154) and F.l1_loss (mel_loss.py:153).
 * B_global is the batch over all ranks.  sc/mag/mel are separate device scalars (NULL to skip);
 * coefs[2*n] receives the backward coefficients consumed by spl_backward(). */
int32_t spl_finalize(const spl_transform* ts, int32_t n, const double* sums, int64_t B_global, int32_t T,
                     float* sc, float* mag, float* mel, float* coefs, void* stream);

/* spl_reduce + spl_finalize in ONE launch for the unsharded case (B_global == B): the last CTA to finish
 * its column computes the losses.  `counter` is a device uint32 that must be zero before the first call
 * and is left at zero by every call (it may be shared by successive calls on one stream). */
int32_t spl_reduce_finalize(const spl_transform* ts, int32_t n, int32_t B, int32_t T, double* sums,
                            float* sc, float* mag, float* mel, float* coefs, uint32_t* counter, void* stream);

/* Sharded batch (one process per GPU), the single exchange step without a collective library:
 * spl_reduce + all-reduce(SUM) of the sums + spl_finalize in ONE launch.  The last CTA of the reduction stores this rank's
 * sums into every rank's symmetric buffer (peer-mapped device memory: NVLink stores), raises a flag, waits for the peers'
 * flags in its own buffer and adds the contributions in rank order, so every rank finalizes bit-identical global losses.
 *   peer_bufs : HOST array of `world` (<= 8) device pointers, peer_bufs[r] = rank r's buffer of
 *               spl_exchange_buffer_bytes() bytes as mapped in THIS process, zero before the first call; one buffer set
 *               per sequence of calls that all ranks issue in the same order
 *   state     : local device uint32[2], zero before the first call (CTA ticket, call epoch)
 *   sums_local / sums_global : local device doubles [sum of n_sums] (this rank's sums / the global sums)
 *   timeout_ns: <= 0 waits for the peers as long as it takes, like a collective (the default of the Python host);
 *               > 0: a peer that has not arrived after timeout_ns nanoseconds (device globaltimer) ends the wait: the
 *               call's epoch (>= 1) is stored to *error_flag and the losses are NaN, so the step cannot be used silently
 *   error_flag: NULL, or a uint32 the host can read WITHOUT synchronising (pinned mapped host memory, or device memory
 *               the host copies back), zero before the first call */
int64_t spl_exchange_buffer_bytes(void);
int32_t spl_reduce_exchange_finalize(const spl_transform* ts, int32_t n, int32_t B, int32_t T, int64_t B_global,
                                     double* sums_local, double* sums_global, int32_t rank, int32_t world,
                                     void* const* peer_bufs, uint32_t* state, int64_t timeout_ns, uint32_t* error_flag,
                                     float* sc, float* mag, float* mel, float* coefs, void* stream);

/* Backward: dx (B, T) = g_sc * dsc/dx + g_mag * dmag/dx + g_mel * dmel/dx -- what autograd derives for
 * the reference modules.  g_* are device scalars (NULL = 0). */
int32_t spl_backward(const spl_transform* ts, int32_t n, int32_t B, int32_t T, const float* coefs,
                     const float* g_sc, const float* g_mag, const float* g_mel, float* dx, void* stream);

/* One call per direction for the unsharded case (what a trainer that launches eagerly pays per step is host time):
 * spl_loss_forward = spl_forward + spl_reduce_finalize, spl_loss_backward = spl_backward.  `ts` is a per-recipe TEMPLATE
 * (tables filled in; partials / gframes ignored); the per-call workspace `ws` (device) is carved up by BYTE offsets the
 * caller computed once from spl_geometry_of(): off_partials[n], off_gframes[n] (NULL = forward only, torch.no_grad),
 * off_sums (doubles), off_coefs (2*n floats). */
int32_t spl_loss_forward(const spl_transform* ts, int32_t n, const float* x, const float* y, int32_t B, int32_t T,
                         void* ws, const int64_t* off_partials, const int64_t* off_gframes, int64_t off_sums,
                         int64_t off_coefs, float* sc, float* mag, float* mel, uint32_t* counter, void* stream);
int32_t spl_loss_backward(const spl_transform* ts, int32_t n, int32_t B, int32_t T, void* ws, const int64_t* off_gframes,
                          int64_t off_coefs, const float* g_sc, const float* g_mag, const float* g_mel, float* dx, void* stream);

/* Explicit magnitude spectrogram out[b, t, k] = sqrt(max(|STFT(x)[b, t, k]|^2, eps)), (B, 1 + T/hop, ld) with
 * ld >= n_fft/2 + 1 floats per frame (pad columns are zeroed): the tensor stft() returns (stft_loss.py:19-35) and the
 * operand of the mel projection in MelSpectrogram.forward (mel_loss.py:88-91).  Backward: spl_spectrogram_backward().  window: device, `win`
 * taps; twiddle: device, 2*n_fft floats from spl_fill_twiddle().  out_lo: NULL, or a second (B, F, ld) buffer that
 * receives A - tf32(A) while `out` receives tf32(A) -- the operand split spl_mel_project() consumes. */
int32_t spl_spectrogram(const float* x, int32_t B, int32_t T, int32_t n_fft, int32_t hop, int32_t win,
                        const float* window, const float* twiddle, float eps, float* out, float* out_lo, int32_t ld,
                        void* stream);

/* Backward of the explicit spectrograms: what autograd derives for stft() (stft_loss.py:19-35: torch.stft -> power ->
 * clamp -> sqrt) and for MelSpectrogram.forward (mel_loss.py:88-94: ... -> matmul(melmat) -> clamp -> log -> transpose)
 * given the gradient of their output tensor.  The spectra are recomputed from x (two frames per complex FFT).
 *   t->kind == SPL_KIND_STFT: g = dL/d stft(x),             (B, 1 + T/hop, ld) with ld >= n_fft/2+1 floats per frame
 *   t->kind == SPL_KIND_MEL : g = dL/d MelSpectrogram(x),   (B, n_mels, 1 + T/hop); ld is ignored; t carries the mel
 *                             tables, n_mels, inv_ln_base and eps (used for both clamps, mel_loss.py:90,92)
 * t->gframes: device workspace of B * (1 + T/hop) * win floats; t->partials is not used.  dx (B, T) is overwritten. */
int32_t spl_spectrogram_backward(const spl_transform* t, const float* x, int32_t B, int32_t T,
                                 const float* g, int32_t ld, float* dx, void* stream);

/* Mel projection + clamp + log as a tensor-core GEMM (tcgen05.mma kind::tf32, 3xTF32 operand split, TMA-fed):
 *     out[b, m, t] = log_scale * ln(max(sum_k A[b*frames + t, k] * W[m, k], eps))
 * i.e. log_b(clamp(matmul(x_amp, melmat), eps)).transpose(1, 2) of MelSpectrogram.forward (mel_loss.py:91-94).
 * amp_hi / amp_lo: device (rows, ld) from spl_spectrogram(out, out_lo); w_hi / w_lo: device (n_pad, ld) = melmat^T
 * split the same way on the host, zero padded; n_pad = n_mels rounded up to 16 (<= 128); ld a multiple of 32.
 * rows = B * frames.  All device pointers 16-byte aligned. */
int32_t spl_mel_project(const float* amp_hi, const float* amp_lo, int64_t rows, int32_t ld,
                        const float* w_hi, const float* w_lo, int32_t n_mels, int32_t n_pad, int32_t frames,
                        float eps, float log_scale, float* out, void* stream);

/* ---- Waveform shape loss: WaveformShapeLoss / MultiWindowShapeLoss (losses/waveform_loss.py:15-75), the third metric
 * loss of TrainerGAN._metric_loss (trainer/trainerGAN.py:235-239):
 *     loss = mean over r of mean |maxpool_{w_r}(|x|) - maxpool_{w_r}(|y|)|,   MaxPool1d(w) = disjoint windows, T / w of them.
 * x, y: device (rows, T) fp32 (rows = B * C).  winlens: HOST array of n <= 8 window lengths, 1 <= w <= T.
 * Forward leaves one 4-byte record per window (argmax of |x| and the sign of the gradient) for the backward. */
#define SPL_SHAPE_MAX_WINDOWS 8

/* Workspace sizes: records (int32) and partials (double). */
int32_t spl_shape_geometry(int32_t rows, int32_t T, const int32_t* winlens, int32_t n, int64_t* record_count,
                           int64_t* partial_count);

/* Per-window maxima and records, then sums[r] = sum over (row, window) of |pool(x) - pool(y)| (device doubles, fixed
 * order).  Multi-GPU: all-reduce sums (SUM) before spl_shape_finalize(). */
int32_t spl_shape_forward(const float* x, const float* y, int32_t rows, int32_t T, const int32_t* winlens, int32_t n,
                          int32_t* records, double* partials, double* sums, void* stream);

/* loss = (1/n) sum_r sums[r] / (rows_global * (T / w_r)): L1Loss means (waveform_loss.py:21,37) and the mean over the
 * window lengths (waveform_loss.py:70-73). */
int32_t spl_shape_finalize(const double* sums, int64_t rows_global, int32_t T, const int32_t* winlens, int32_t n,
                           float* loss, void* stream);

/* dx (rows, T) = g * dloss/dx: the gradient autograd derives (L1 sign -> max_pool1d argmax routing -> abs sign).
 * g: device scalar.  dx is overwritten. */
int32_t spl_shape_backward(const int32_t* records, int32_t rows, int64_t rows_global, int32_t T, const int32_t* winlens,
                           int32_t n, const float* g, float* dx, void* stream);

/* ---- Losses on explicit magnitude tensors, for callers that compose them with stft() themselves (STFTLoss.forward,
 * stft_loss.py:112-116): SpectralConvergenceLoss.forward(x_mag, y_mag) = ||y - x||_F / ||y||_F (stft_loss.py:38-56) and
 * LogSTFTMagnitudeLoss.forward = mean |ln y - ln x| (stft_loss.py:59-77).  x_mag, y_mag: device, n fp32 elements. */
int32_t spl_mag_loss_geometry(int64_t n, int64_t* partial_count);

/* sums: device doubles [6]; [0..2] = {sum (y-x)^2, sum y^2, sum |ln y - ln x|} (fixed order), [3..5] = coefficients the
 * backward consumes.  sc / mag: device scalars, NULL to skip.  Magnitudes must be positive (stft() output is). */
int32_t spl_mag_loss_forward(const float* x_mag, const float* y_mag, int64_t n, double* partials, double* sums,
                             float* sc, float* mag, void* stream);

/* gx / gy (n fp32 each, NULL to skip one) = g_sc * d sc + g_mag * d mag w.r.t. x_mag / y_mag; g_sc, g_mag: device scalars
 * (NULL = 0). */
int32_t spl_mag_loss_backward(const float* x_mag, const float* y_mag, int64_t n, const double* sums,
                              const float* g_sc, const float* g_mag, float* gx, float* gy, void* stream);

/* Power-mel L1 evaluation metric `Mel_L1(pred, target)` (mel_spectrogram.py:36-44, sandbox.py:183-191):
 * nn.L1Loss()(M(pred), M(target)) with M = torchaudio.transforms.MelSpectrogram(48000), i.e. n_fft = win = 400 (periodic
 * Hann), hop 200, reflect-centred, |STFT|^2, 128 HTK mel filters without normalisation, no log.  Forward only.
 *   x, y      : device (rows, T)
 *   n_fft     : must be 400 (= 25 x 16: the transform of this kernel); hop any >= 1
 *   window    : device, 400 taps;  twiddle: device, 800 floats: W_400^(n1*k2) = exp(-2 pi i n1 k2 / 400) at [k2*25 + n1], k2 < 16, n1 < 25
 *   mel_ptr / mel_ent : device CSR of the (n_mels x 201) filterbank: row m owns entries [mel_ptr[m], mel_ptr[m+1]) of
 *               mel_ent = {bin, bits of the weight} pairs (nnz of them)
 *   partials  : device doubles, spl_melpow_geometry() of them;  sum: device double[1];  loss: device float[1]
 *   mel_x / mel_y : NULL, or device (rows, n_mels, 1 + T/hop) outputs M(pred) / M(target) */
int32_t spl_melpow_geometry(int32_t rows, int32_t T, int32_t n_fft, int32_t hop, int64_t* partial_count);
int32_t spl_melpow_l1(const float* x, const float* y, int32_t rows, int32_t T, int32_t n_fft, int32_t hop,
                      const float* window, const float* twiddle, int32_t n_mels, int32_t nnz,
                      const int32_t* mel_ptr, const int32_t* mel_ent, double* partials, double* sum, float* loss,
                      float* mel_x, float* mel_y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPECLOSS_H */
