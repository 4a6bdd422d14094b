// lseg_b200 — shared pieces of the fused multi-head self-attention kernels (mhsa2.cuh, mhsa3.cuh): parameters, the
// scalar softmax chunk helpers, the host-side descriptor. The round-1 single-stream kernel that used to live here
// (described below for the record of the design's evolution) was removed once mhsa2 / mhsa3 superseded it.
//
// softmax( (Q K^T) * dh^-0.5 [+ causal mask] ) V for head_dim 64, reading the packed QKV rows
// produced by the QKV GEMM ([B, N, 3*D] fp16; q | k | v thirds, head h at columns h*64) and
// writing [B*N, D] fp16. Restates timm Attention (corroborated in-tree at
// modules/models/lseg_vit.py:26-39) and CLIP's causal nn.MultiheadAttention (Appendix A.2) without
// ever materialising the [B, heads, N, N] score tensor.
//
// One CTA = one 128-row query tile of one (image, head); 2 CTAs co-reside per SM. 10 warps:
//   warp 0      : TMA producer (Q once; K and V tiles of 128 keys through separate 2-deep rings — a K
//                 slot is released as soon as its S MMA retires, a V slot after its PV MMA) + TMEM alloc
//   warp 1      : MMA issuer:  S_j = Q K_j^T (128x128x64, both K-major)      -> TMEM cols [0,128)
//                              O  += P_j V_j (128x64x128, A = P K-major smem, B = V MN-major smem)
//                                                                          -> TMEM cols [128,192)
//                 S_{j+1} is issued as soon as the softmax warps have read S_j for the last time.
//   warps 2..9  : softmax, thread <-> query row (tcgen05.ld 32x32b); the two warps that share a TMEM lane
//                 quarter split the 128 keys of a tile in halves (64 score registers each) and exchange
//                 their partial row maxima through smem + a 64-thread named barrier. The profile of the
//                 4-warp version was latency-bound (IPC 0.4, MUFU 36 % busy): 16 softmax warps per SM instead
//                 of 8 hide the dependent-issue and TMEM/barrier latencies. P is written to smem as fp16 in the
//                 128B-swizzled K-major layout, O stays in TMEM across key tiles. The running max is LAZY:
//                 the exponent offset m_ref only moves when a tile's row max exceeds it by more than
//                 2^kMhsaTau, in which case l and the TMEM-resident O row are rescaled (rare after tile 0).
#pragma once
#include "common.cuh"

namespace lseg {

constexpr int kMhsaThreads = 320;
constexpr int kMhsaTile = 128;
constexpr int kMhsaDh = 64;
constexpr int kMhsaTileBytes = kMhsaTile * kMhsaDh * 2;  // 16 KB
// Q | K x2 | V x2 | P (2 sub-tiles) | 128 B mbarriers | 512 B partial-max exchange | pad
constexpr int kMhsaSmemBytes = kMhsaTileBytes * (1 + 2 + 2 + 2) + 768;
constexpr float kMhsaTau = 8.0f;  // log2 headroom before the exponent offset is moved

struct MhsaParams {
  CUtensorMap tma_qkv;  // 3-D {3*D, N, B} fp16, box {64, 128, 1}
  CUtensorMap tma_t64;  // same tensor, box {64, 64, 1} (mhsa2.cuh: 64-key tiles, Q as two boxes)
  __half* out;          // [B*N, D]
  int n_tokens;
  int heads;
  int D;
  int causal;
  float scale_log2e;    // dh^-0.5 * log2(e)
  unsigned long long* trace;  // debug timeline buffer (mhsa2 TRACE instantiation only), else nullptr
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// exp2(s*c - m) for one 32-column chunk -> fp16 pairs + fp32 partial row sum. MASK: apply key bounds.
// POLY of every 8 exponentials are evaluated on the FMA pipe (exp2_poly<3>, relative error 7.5e-5 — below the
// fp16 rounding of P) instead of the MUFU pipe, which is the bound of this kernel.
template <bool MASK, int POLY = 0>
__device__ __forceinline__ float mhsa_exp_chunk(const uint32_t (&s)[32], float c, float m, int kv_base, int n_tokens,
                                                int kv_limit, __half2 (&ph)[16]) {
  float sum0 = 0.f, sum1 = 0.f;  // two chains: the row sum must not serialise behind the MUFU results
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float x0 = fmaf(__uint_as_float(s[i]), c, -m), x1 = fmaf(__uint_as_float(s[i + 1]), c, -m);
    float p0 = ((i & 7) < POLY) ? exp2_poly<3>(fmaxf(x0, -126.f)) : ex2_approx(x0);
    float p1 = (((i + 1) & 7) < POLY) ? exp2_poly<3>(fmaxf(x1, -126.f)) : ex2_approx(x1);
    if (MASK) {
      const int kv = kv_base + i;
      if (!(kv < n_tokens && kv <= kv_limit)) p0 = 0.f;
      if (!(kv + 1 < n_tokens && kv + 1 <= kv_limit)) p1 = 0.f;
    }
    ph[i >> 1] = __floats2half2_rn(p0, p1);
    sum0 += p0;  // fp32 row sum (the fp16 rounding of P is zero-mean noise at 2^-12)
    sum1 += p1;
  }
  return sum0 + sum1;
}

template <bool MASK>
__device__ __forceinline__ float mhsa_max_chunk(const uint32_t (&s)[32], int kv_base, int n_tokens, int kv_limit) {
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four chains: the reduction is latency-bound
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float a = __uint_as_float(s[i + 2 * u]), b = __uint_as_float(s[i + 2 * u + 1]);
      if (MASK) {
        const int kv = kv_base + i + 2 * u;
        if (!(kv < n_tokens && kv <= kv_limit)) a = -INFINITY;
        if (!(kv + 1 < n_tokens && kv + 1 <= kv_limit)) b = -INFINITY;
      }
      m[u] = fmaxf(m[u], fmaxf(a, b));
    }
  }
  return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}

struct MhsaDesc {
  const __half* qkv;  // [B, N, 3*D]
  __half* out;        // [B*N, D]
  int B, N, heads;    // D = heads * 64
  int causal;
};

}  // namespace lseg
