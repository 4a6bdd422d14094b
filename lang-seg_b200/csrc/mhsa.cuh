// lseg_b200 — fused multi-head self-attention (flash-style) on tcgen05 for sm_100a.
//
// softmax( (Q K^T) * dh^-0.5 [+ causal mask] ) V for head_dim 64, reading the packed QKV rows
// produced by the QKV GEMM ([B, N, 3*D] fp16; q | k | v thirds, head h at columns h*64) and
// writing [B*N, D] fp16. Restates timm Attention (corroborated in-tree at
// modules/models/lseg_vit.py:26-39) and CLIP's causal nn.MultiheadAttention (Appendix A.2) without
// ever materialising the [B, heads, N, N] score tensor.
//
// One CTA = one 128-row query tile of one (image, head); 2 CTAs co-reside per SM so one CTA's
// softmax overlaps the other's MMAs.
//   warp 0      : TMA producer (Q once, then K/V tiles of 128 keys through a 2-deep ring) + TMEM alloc
//   warp 1      : MMA issuer:  S = Q K^T (128x128x64, both K-major)  ->  TMEM cols [0,128)
//                              O_j = P V (128x64x128, A = P K-major in smem, B = V MN-major)
//                                   ->  TMEM cols [128,192) / [192,256) alternating
//   warps 2..5  : softmax: thread <-> query row (tcgen05.ld 32x32b), online max/sum in fp32,
//                 P written to smem as fp16 in the 128B-swizzled K-major layout the MMA expects,
//                 running output kept in registers: o = o*alpha + O_j.
#pragma once
#include "common.cuh"

namespace lseg {

constexpr int kMhsaThreads = 192;
constexpr int kMhsaTile = 128;
constexpr int kMhsaDh = 64;
constexpr int kMhsaTileBytes = kMhsaTile * kMhsaDh * 2;  // 16 KB
constexpr int kMhsaSmemBytes = kMhsaTileBytes * (1 + 2 + 2 + 2) + 256;

struct MhsaParams {
  CUtensorMap tma_qkv;  // 3-D {3*D, N, B} fp16, box {64, 128, 1}
  __half* out;          // [B*N, D]
  int n_tokens;
  int heads;
  int D;
  int causal;
  float scale_log2e;    // dh^-0.5 * log2(e)
};

__global__ void __launch_bounds__(kMhsaThreads, 2) mhsa_kernel(const __grid_constant__ MhsaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kMhsaTileBytes;       // 2 stages
  uint8_t* sV = smem + 3 * kMhsaTileBytes;   // 2 stages
  uint8_t* sP = smem + 5 * kMhsaTileBytes;   // 128 x 128 fp16 = two 128x64 swizzled sub-tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kMhsaTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

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

  // number of key tiles this query tile attends to
  int kv_end = p.n_tokens;
  if (p.causal) kv_end = min(p.n_tokens, q0 + kMhsaTile);
  const int nkv = (kv_end + kMhsaTile - 1) / kMhsaTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&o_full[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kMhsaTileBytes);
      tma_load_3d(sQ, &p.tma_qkv, q_full, h * kMhsaDh, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1, 11);
        mbar_expect_tx(&kv_full[s], 2 * kMhsaTileBytes);
        tma_load_3d(sK + s * kMhsaTileBytes, &p.tma_qkv, &kv_full[s], p.D + h * kMhsaDh, j * kMhsaTile, b);
        tma_load_3d(sV + s * kMhsaTileBytes, &p.tma_qkv, &kv_full[s], 2 * p.D + h * kMhsaDh, j * kMhsaTile, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);  // Q K^T : A, B K-major
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 1);   // P V   : A K-major, B MN-major
      const uint32_t q_base = smem_u32(sQ);
      const uint32_t p_base = smem_u32(sP);
      mbar_wait(q_full, 0, 12);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1, 13);
        tc_fence_after();
        const uint32_t k_base = smem_u32(sK + s * kMhsaTileBytes);
        const uint32_t v_base = smem_u32(sV + s * kMhsaTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tS, umma_desc_sw128(q_base + k * 32, 1024, 0), umma_desc_sw128(k_base + k * 32, 1024, 0),
                      idesc_s, k != 0);
        umma_commit(s_full);
        mbar_wait(p_full, j & 1, 14);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ss(tO + (j & 1) * 64, umma_desc_sw128(p_base + (k >> 2) * kMhsaTileBytes + (k & 3) * 32, 1024, 0),
                      umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, k != 0);
        umma_commit(&o_full[j & 1]);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // ===================== softmax warps =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    float m_run = -INFINITY;
    float l_run = 0.f;
    float alpha_prev = 0.f;
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    const float c = p.scale_log2e;
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < nkv; ++j) {
      const int kv0 = j * kMhsaTile;
      const bool need_mask = (kv0 + kMhsaTile > p.n_tokens) || (p.causal && (kv0 + kMhsaTile - 1 > q0));
      mbar_wait(s_full, j & 1, 15);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tS + lane_off + cc * 32, v);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kv = kv0 + cc * 32 + i;
            const bool ok = (kv < p.n_tokens) && (!p.causal || kv <= q);
            mx = fmaxf(mx, ok ? __uint_as_float(v[i]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx * c);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2_approx(m_run - m_use);  // m_run = -inf -> 0
      // P buffer is free once PV_{j-1} has retired
      if (j > 0) mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1, 16);
      // pass 2: p = exp2(s*c - m), row sum, fp16 P -> swizzled smem
      float l_tile = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tS + lane_off + cc * 32, v);
        tmem_ld_wait();
        __half2 ph[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), c, -m_use));
          float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -m_use));
          if (need_mask) {
            const int kv = kv0 + cc * 32 + i;
            if (!((kv < p.n_tokens) && (!p.causal || kv <= q))) p0 = 0.f;
            if (!((kv + 1 < p.n_tokens) && (!p.causal || kv + 1 <= q))) p1 = 0.f;
          }
          const __half2 hh = __floats2half2_rn(p0, p1);
          ph[i >> 1] = hh;
          const float2 back = __half22float2(hh);  // sum what the MMA will actually multiply
          l_tile += back.x + back.y;
        }
        // columns [cc*32, cc*32+32) -> sub-tile (cc>>1), 16-byte chunks ((cc&1)*4 + t) ^ (r&7)
        uint8_t* sub = p_row + (cc >> 1) * kMhsaTileBytes;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int chunk = ((cc & 1) * 4 + t) ^ sw;
          *reinterpret_cast<uint4*>(sub + chunk * 16) = *reinterpret_cast<uint4*>(&ph[4 * t]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      // fold in O_{j-1} while PV_j runs
      if (j > 0) {
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(tO + ((j - 1) & 1) * 64 + lane_off + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[cc * 32 + i] = fmaf(o_acc[cc * 32 + i], alpha_prev, __uint_as_float(v[i]));
        }
      }
      alpha_prev = alpha;
    }
    // last tile's O
    {
      const int j = nkv - 1;
      mbar_wait(&o_full[j & 1], (j >> 1) & 1, 17);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tO + (j & 1) * 64 + lane_off + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[cc * 32 + i] = fmaf(o_acc[cc * 32 + i], alpha_prev, __uint_as_float(v[i]));
      }
    }
    if (q < p.n_tokens) {
      const float inv = 1.0f / l_run;
      __half* op = p.out + (static_cast<long long>(b) * p.n_tokens + q) * p.D + h * kMhsaDh;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        __half2 hh[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hh[i] = __floats2half2_rn(o_acc[t * 8 + 2 * i] * inv, o_acc[t * 8 + 2 * i + 1] * inv);
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
