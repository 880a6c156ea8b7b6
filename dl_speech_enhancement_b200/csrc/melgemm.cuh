// Mel filterbank projection as a tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) fed by TMA, accumulator
// in TMEM, fused clamp + log epilogue.  Replaces, on the GPU,
//     x_mel = clamp(matmul(x_amp, melmat), eps);  log(x_mel).transpose(1, 2)      (losses/mel_loss.py:91-94)
// for the explicit MelSpectrogram.forward() tensor.  (Inside the fused loss kernels the projection is banded and
// runs on CUDA cores from registers; DESIGN.md section 6 compares the two.)
//
//   D[r, m] = sum_k A[r, k] * W[m, k]        A = amplitudes (rows = B*F frames, K = n_fft/2+1 padded to ld)
//                                            W = melmat^T   (n_pad >= n_mels rows, K-major like A)
// fp32 accuracy out of TF32 tensor cores by operand splitting (3xTF32): A = A_hi + A_lo, W = W_hi + W_lo with the
// *_hi parts exactly representable in TF32; D = A_hi W_hi + A_lo W_hi + A_hi W_lo (the lo*lo term is < 2^-20 relative).
// The splits are made by the producers (spec_kernel writes A_hi / A_lo, the host splits the constant W), so every
// operand tile is a plain TMA load into a 128-byte-swizzled K-major tile.
//
// One CTA = one 128-row tile of A, 4 warps: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// all four warps = epilogue (TMEM -> registers -> log -> (B, n_mels, F) stores, coalesced along F).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace spl {

constexpr int kGemmBM = 128;        // rows per CTA = UMMA M
constexpr int kGemmBK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kGemmUK = 8;          // K of one tcgen05.mma.kind::tf32
constexpr int kGemmStages = 3;
constexpr int kGemmTmemCols = 128;  // accumulator columns allocated (power of two >= n_pad)
constexpr int kGemmMaxN = 128;

struct MelGemmParams {
  long long rows;        // B * F
  int n_mels, n_pad;     // valid / padded (multiple of 16) output columns
  int frames;            // F: rows per utterance
  int kblocks;           // ld / 32
  float eps, log_scale;  // out = log_scale * ln(max(D, eps))
  float* out;            // (B, n_mels, F)
};

__host__ __device__ inline size_t mel_gemm_stage_bytes(int n_pad) { return 2 * (size_t)kGemmBM * 128 + 2 * (size_t)n_pad * 128; }
__host__ __device__ inline size_t mel_gemm_smem_bytes(int n_pad) { return kGemmStages * mel_gemm_stage_bytes(n_pad) + 1024; }

#ifdef __CUDACC__
namespace gemm_detail {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a pipeline bug must trap (and surface as a CUDA error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (= 1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.  Address fields are in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {     // arrives on `bar` when all MMAs issued so far are done
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace gemm_detail

__global__ void __launch_bounds__(128, 1)
mel_gemm_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
                const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                const MelGemmParams p) {
  using namespace gemm_detail;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kGemmStages], bar_empty[kGemmStages], bar_accum;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;       // swizzled tiles need 1024-byte alignment
  const uint32_t a_bytes = kGemmBM * 128, w_bytes = (uint32_t)p.n_pad * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * w_bytes;
  const long long m0 = (long long)blockIdx.x * kGemmBM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGemmStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {            // one warp allocates the accumulator columns and later frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kGemmTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0) {
    // ---- TMA producer: per k-block A_hi, A_lo (128 x 32 fp32) and W_hi, W_lo (n_pad x 32 fp32) ----
    if (lane == 0) {
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int s = kb % kGemmStages;
        const uint32_t ph = (uint32_t)(kb / kGemmStages) & 1u;
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);                 // fresh barrier: parity 1 passes at once
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, stage_bytes);
        const uint32_t st = tiles + (uint32_t)s * stage_bytes;
        tma_load_2d(st, &tm_ah, full, kb * kGemmBK, (int)m0);
        tma_load_2d(st + a_bytes, &tm_al, full, kb * kGemmBK, (int)m0);
        tma_load_2d(st + 2 * a_bytes, &tm_wh, full, kb * kGemmBK, 0);
        tma_load_2d(st + 2 * a_bytes + w_bytes, &tm_wl, full, kb * kGemmBK, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer: 3 x (BK / 8) tcgen05.mma per k-block, then release the stage ----
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kGemmBM, p.n_pad);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int s = kb % kGemmStages;
        const uint32_t ph = (uint32_t)(kb / kGemmStages) & 1u;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = tiles + (uint32_t)s * stage_bytes;
        const uint64_t d_ah = umma_desc_sw128(st), d_al = umma_desc_sw128(st + a_bytes);
        const uint64_t d_wh = umma_desc_sw128(st + 2 * a_bytes), d_wl = umma_desc_sw128(st + 2 * a_bytes + w_bytes);
#pragma unroll
        for (int k = 0; k < kGemmBK / kGemmUK; ++k) {
          const uint64_t adv = (uint64_t)((k * kGemmUK * 4) >> 4);   // 32 bytes per K step, inside the swizzle row
          umma_tf32(tmem_d, d_ah + adv, d_wh + adv, idesc, (kb | k) ? 1u : 0u);
          umma_tf32(tmem_d, d_al + adv, d_wh + adv, idesc, 1u);
          umma_tf32(tmem_d, d_ah + adv, d_wl + adv, idesc, 1u);
        }
        umma_commit(smem_u32(&bar_empty[s]));
      }
      umma_commit(smem_u32(&bar_accum));
    }
    __syncwarp();
  }

  // ---- epilogue: TMEM lane = tile row (warp w owns lanes 32w .. 32w+31), column = mel ----
  mbar_wait(smem_u32(&bar_accum), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const long long r = m0 + warp * 32 + lane;
  const bool valid = r < p.rows;
  const long long b = valid ? r / p.frames : 0;
  const int t = valid ? (int)(r - b * p.frames) : 0;
  float* outp = p.out + (size_t)b * p.n_mels * p.frames + t;
  for (int c = 0; c < p.n_pad; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    if (valid) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c + i < p.n_mels) outp[(size_t)(c + i) * p.frames] = p.log_scale * logf(fmaxf(__uint_as_float(v[i]), p.eps));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kGemmTmemCols) : "memory");
}
#endif  // __CUDACC__

}  // namespace spl
