// lseg_b200 — fused multi-head self-attention (flash-style) on tcgen05 for sm_100a.
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

// SPIN: barrier waits re-poll without a suspend hint (lower wake-up latency, more issue slots burnt);
// A/B switch LSEG_MHSA_SPIN=1.
template <bool SPIN>
__global__ void __launch_bounds__(kMhsaThreads, 2) mhsa_kernel(const __grid_constant__ MhsaParams p) {
  auto wait_bar = [](uint64_t* bar, uint32_t parity, int tag) {
    if (SPIN)
      mbar_wait_spin(bar, parity, tag);
    else
      mbar_wait(bar, parity, tag);
  };
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kMhsaTileBytes;      // 2 stages
  uint8_t* sV = smem + 3 * kMhsaTileBytes;  // 2 stages
  uint8_t* sP = smem + 5 * kMhsaTileBytes;  // 128 x 128 fp16 = two 128x64 swizzled sub-tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kMhsaTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* o_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  __half* xch = reinterpret_cast<__half*>(smem + 7 * kMhsaTileBytes + 128);  // [2 halves][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int b = blockIdx.y / p.heads;
  const int h = blockIdx.y % p.heads;
  const int q0 = q_tile * kMhsaTile;

  if ((smem_u32(smem) & 1023u) != 0) {  // layout contract of the swizzled tiles
    if (threadIdx.x == 0) atomicCAS(&g_watchdog[0], 0, 99);
    return;
  }

  int kv_end = p.n_tokens;
  if (p.causal) kv_end = min(p.n_tokens, q0 + kMhsaTile);
  const int nkv = (kv_end + kMhsaTile - 1) / kMhsaTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);  // one elected arrival per softmax warp (after __syncwarp)
    mbar_init(p_full, 8);
    mbar_init(o_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + 128;

  if (warp < 2) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_expect_tx(q_full, kMhsaTileBytes);
        tma_load_3d(sQ, &p.tma_qkv, q_full, h * kMhsaDh, q0, b);
        for (int j = 0; j < nkv; ++j) {
          const int s = j & 1;
          const uint32_t par = ((j >> 1) & 1) ^ 1;
          wait_bar(&k_empty[s], par, 11);
          mbar_expect_tx(&k_full[s], kMhsaTileBytes);
          tma_load_3d(sK + s * kMhsaTileBytes, &p.tma_qkv, &k_full[s], p.D + h * kMhsaDh, j * kMhsaTile, b);
          wait_bar(&v_empty[s], par, 12);
          mbar_expect_tx(&v_full[s], kMhsaTileBytes);
          tma_load_3d(sV + s * kMhsaTileBytes, &p.tma_qkv, &v_full[s], 2 * p.D + h * kMhsaDh, j * kMhsaTile, b);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        constexpr uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);  // Q K^T : A, B K-major
        constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 1);   // P V   : A K-major, B MN-major
        const uint32_t q_base = smem_u32(sQ);
        const uint32_t p_base = smem_u32(sP);
        auto issue_s = [&](int j) {
          const uint32_t k_base = smem_u32(sK + (j & 1) * kMhsaTileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tS, umma_desc_sw128(q_base + k * 32, 1024, 0), umma_desc_sw128(k_base + k * 32, 1024, 0),
                        idesc_s, k != 0);
          umma_commit(s_full);
          umma_commit(&k_empty[j & 1]);
        };
        wait_bar(q_full, 0, 13);
        wait_bar(&k_full[0], 0, 14);
        tc_fence_after();
        issue_s(0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) {
            wait_bar(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1, 15);
            wait_bar(s_free, j & 1, 16);  // S_j now lives in the softmax warps' registers
            tc_fence_after();
            issue_s(j + 1);
          }
          wait_bar(&v_full[j & 1], (j >> 1) & 1, 17);
          wait_bar(p_full, j & 1, 18);
          tc_fence_after();
          const uint32_t v_base = smem_u32(sV + (j & 1) * kMhsaTileBytes);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_ss(tO, umma_desc_sw128(p_base + (k >> 2) * kMhsaTileBytes + (k & 3) * 32, 1024, 0),
                        umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, (j | k) != 0);
          umma_commit(o_done);
          umma_commit(&v_empty[j & 1]);
        }
      }
    }
  } else {
    // ===================== softmax warps 2..9 =====================
    const int quarter = warp & 3;        // TMEM lane quarter (warps w and w+4 share it)
    const int hf = (warp - 2) >> 2;      // which 64 keys of the tile / which 32 output columns at the end
    const int r = quarter * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.scale_log2e;
    const int kv_limit = p.causal ? q : 0x7fffffff;
    float m_ref = -INFINITY;  // exponent offset currently baked into l and O (identical in both warps of a row)
    float l_run = 0.f;        // this warp's share of the row sum
    uint8_t* p_row = sP + hf * kMhsaTileBytes + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < nkv; ++j) {
      const int kv0 = j * kMhsaTile + hf * 64;
      const bool need_mask = (j * kMhsaTile + kMhsaTile > p.n_tokens) || (p.causal && (j * kMhsaTile + kMhsaTile - 1 > q0));
      wait_bar(s_full, j & 1, 19);
      tc_fence_after();
      // pass 1: row max over this warp's 64 keys (the scores are re-read from TMEM in pass 2: holding 64 of
      // them across the exchange does not fit the 2-CTA/SM register budget), then exchange with the sibling
      // warp (same rows, other 64 keys)
      float pm;
      {
        uint32_t s0[32], s1[32];
        __syncwarp();
        tmem_ld32(tS + lane_off + hf * 64, s0);
        tmem_ld32(tS + lane_off + hf * 64 + 32, s1);
        tmem_ld_wait();
        pm = need_mask ? fmaxf(mhsa_max_chunk<true>(s0, kv0, p.n_tokens, kv_limit),
                               mhsa_max_chunk<true>(s1, kv0 + 32, p.n_tokens, kv_limit))
                       : fmaxf(mhsa_max_chunk<false>(s0, kv0, p.n_tokens, kv_limit),
                               mhsa_max_chunk<false>(s1, kv0 + 32, p.n_tokens, kv_limit));
      }
      // both warps must use the SAME offset: exchange fp16-rounded values and round the own one too (the
      // offset only has to be shared and within 2^tau of the true max, not exact)
      const __half pm_h = __float2half_rn(pm * c);
      xch[hf * 128 + r] = pm_h;
      named_bar_sync(1 + quarter, 64);
      const float mx = fmaxf(__half2float(pm_h), __half2float(xch[(hf ^ 1) * 128 + r]));
      named_bar_sync(1 + quarter, 64);  // the slot may be rewritten only after the sibling has read it
      const bool move = mx > m_ref + kMhsaTau;  // also true on the first tile (m_ref = -inf)
      const bool any_move = __any_sync(0xffffffffu, move);
      float factor = 1.f;
      if (any_move) {
        const float m_new = move ? mx : m_ref;
        factor = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new);
        l_run *= factor;
        m_ref = m_new;
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      // P buffer free and O quiescent once PV_{j-1} has retired
      if (j > 0) {
        wait_bar(o_done, (j - 1) & 1, 20);
        if (any_move && hf == 0) {  // rescale the TMEM-resident output row (warp-collective; factor = 1 if unmoved)
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t o[16];
            __syncwarp();
            tmem_ld16(tO + lane_off + cc * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st16(tO + lane_off + cc * 16, o);
          }
          tmem_st_wait();
        }
      }
      // p = exp2(s*c - m_ref) -> fp16 -> swizzled smem: this warp owns sub-tile hf (64 keys = 8 16-byte slots per row)
      float l_tile = 0.f;
      {
        uint32_t s0[32];
        __syncwarp();
        tmem_ld32(tS + lane_off + hf * 64, s0);
        tmem_ld_wait();
        __half2 ph[16];
        l_tile += need_mask ? mhsa_exp_chunk<true>(s0, c, m_use, kv0, p.n_tokens, kv_limit, ph)
                            : mhsa_exp_chunk<false>(s0, c, m_use, kv0, p.n_tokens, kv_limit, ph);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          *reinterpret_cast<uint4*>(p_row + ((t ^ sw) * 16)) = *reinterpret_cast<uint4*>(&ph[4 * t]);
      }
      {
        uint32_t s1[32];
        __syncwarp();
        tmem_ld32(tS + lane_off + hf * 64 + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);  // S_j fully consumed: the MMA warp may issue the next tile's scores
        __half2 ph[16];
        l_tile += need_mask ? mhsa_exp_chunk<true>(s1, c, m_use, kv0 + 32, p.n_tokens, kv_limit, ph)
                            : mhsa_exp_chunk<false>(s1, c, m_use, kv0 + 32, p.n_tokens, kv_limit, ph);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          *reinterpret_cast<uint4*>(p_row + (((4 + t) ^ sw) * 16)) = *reinterpret_cast<uint4*>(&ph[4 * t]);
      }
      l_run += l_tile;
      fence_proxy_async_smem();  // every writer makes its P stores visible to the async (UMMA) proxy ...
      tc_fence_before();
      __syncwarp();              // ... before the warp's single elected arrival
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l. The P buffer is dead once the last PV has retired: reuse it to add up the two l shares.
    wait_bar(o_done, (nkv - 1) & 1, 25);
    tc_fence_after();
    float* lx = reinterpret_cast<float*>(sP);
    lx[hf * 128 + r] = l_run;
    named_bar_sync(1 + quarter, 64);
    const float inv = 1.0f / (l_run + lx[(hf ^ 1) * 128 + r]);
    __half* op = p.out + (static_cast<long long>(b) * p.n_tokens + q) * p.D + h * kMhsaDh + hf * 32;
    uint32_t o[32];
    __syncwarp();
    tmem_ld32(tO + lane_off + hf * 32, o);
    tmem_ld_wait();
    if (q < p.n_tokens) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        __half2 hh[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          hh[i] = __floats2half2_rn(__uint_as_float(o[t * 8 + 2 * i]) * inv, __uint_as_float(o[t * 8 + 2 * i + 1]) * inv);
        reinterpret_cast<uint4*>(op)[t] = *reinterpret_cast<uint4*>(hh);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

struct MhsaDesc {
  const __half* qkv;  // [B, N, 3*D]
  __half* out;        // [B*N, D]
  int B, N, heads;    // D = heads * 64
  int causal;
};

}  // namespace lseg
