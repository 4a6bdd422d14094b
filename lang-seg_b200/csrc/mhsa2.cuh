// lseg_b200 — fused multi-head self-attention, two-stream version (the one the engine runs).
//
// Same contract as mhsa.cuh (softmax(Q K^T * dh^-0.5 [+ causal]) V, head_dim 64, packed [B, N, 3*D] fp16 in,
// [B*N, D] fp16 out; restates timm Attention — SURVEY.md Appendix A.1, modules/models/lseg_vit.py:26-39 — and
// CLIP's causal nn.MultiheadAttention, Appendix A.2). What changed against mhsa.cuh, and why (profiles/r01_mhsa*):
// that kernel ran the MUFU pipe at 45 % and the tensor pipe at 22 % because all eight softmax warps of a CTA
// stalled together (S single-buffered, sibling-warp max exchange, PV hand-off), so an SM only ever had two
// independent instruction streams. Here a CTA runs TWO independent online-softmax streams over 64-key tiles:
//   stream A = even key tiles, stream B = odd key tiles; each has its own S buffer and its own O accumulator in
//   TMEM, its own P buffer in smem and its own (m, l) row state in registers, and four softmax warps
//   (thread <-> query row). No exchange between streams until the end, where the two partial results are merged
//   (O = O_A 2^(mA-M) + O_B 2^(mB-M), same for l). With 2 CTAs/SM an SM has four independent streams.
//   * S is read from TMEM once into registers and released immediately, so the next S MMA of the stream runs
//     under the exponentials of the current tile.
//   * the last key tile only computes ceil16(valid keys) columns (S MMA with N = 16.., PV with K = 16..):
//     901 tokens = 14 tiles of 64 + 5 keys instead of 8 x 128 (11 % fewer exponentials and MMAs).
//   * row groups of 32 that lie entirely beyond the sequence (the last query tile has 5 valid rows of 128)
//     skip the softmax arithmetic: their P rows are garbage, which only reaches their own (unstored) O rows.
//
// Warps: 0 TMA producer (Q; K and V 64-key tiles through 4-deep rings) + TMEM alloc; 1 MMA issuer;
//        2..5 softmax stream A; 6..9 softmax stream B (warp & 3 = TMEM lane quarter).
// TMEM (256 columns, 2 CTAs/SM): S_A [0,64) S_B [64,128) O_A [128,192) O_B [192,256).
#pragma once
#include "common.cuh"
#include "mhsa.cuh"

namespace lseg {

constexpr int kM2Threads = 320;
constexpr int kM2KT = 64;                      // keys per tile
constexpr int kM2KvBytes = kM2KT * kMhsaDh * 2;  // 8 KB
constexpr int kM2Stages = 4;
constexpr int kM2QBytes = 128 * kMhsaDh * 2;   // 16 KB
constexpr int kM2PBytes = 128 * kM2KT * 2;     // 16 KB per stream
// Q | K ring | V ring | P_A P_B | mbarriers + tmem slot
constexpr int kM2SmemBytes = kM2QBytes + 2 * kM2Stages * kM2KvBytes + 2 * kM2PBytes + 1024;

template <bool SPIN>
__global__ void __launch_bounds__(kM2Threads, 2) mhsa2_kernel(const __grid_constant__ MhsaParams p) {
  auto wait_bar = [](uint64_t* bar, uint32_t parity, int tag) {
    if (SPIN)
      mbar_wait_spin(bar, parity, tag);
    else
      mbar_wait(bar, parity, tag);
  };
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kM2QBytes;
  uint8_t* sV = sK + kM2Stages * kM2KvBytes;
  uint8_t* sP = sV + kM2Stages * kM2KvBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kM2PBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [4]
  uint64_t* k_empty = bars + 5;   // [4]
  uint64_t* v_full = bars + 9;    // [4]
  uint64_t* v_empty = bars + 13;  // [4]
  uint64_t* s_full = bars + 17;   // [2] per stream
  uint64_t* s_free = bars + 19;   // [2]
  uint64_t* p_full = bars + 21;   // [2]
  uint64_t* o_done = bars + 23;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int b = blockIdx.y / p.heads;
  const int h = blockIdx.y % p.heads;
  const int q0 = q_tile * 128;

  if ((smem_u32(smem) & 1023u) != 0) {  // layout contract of the swizzled tiles
    if (threadIdx.x == 0) atomicCAS(&g_watchdog[0], 0, 99);
    return;
  }

  const int kv_end = p.causal ? min(p.n_tokens, q0 + 128) : p.n_tokens;
  const int nkt = (kv_end + kM2KT - 1) / kM2KT;
  // columns the S MMA produces / the PV MMA consumes for tile j: valid keys rounded up to the UMMA granule
  auto tile_cols = [&](int j) { return min(kM2KT, ((kv_end - j * kM2KT) + 15) & ~15); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_t64);
    mbar_init(q_full, 1);
    for (int i = 0; i < kM2Stages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], 4);  // one elected arrival per softmax warp of the stream
      mbar_init(&p_full[s], 4);
      mbar_init(&o_done[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, kM2QBytes);
      tma_load_3d(sQ, &p.tma_t64, q_full, h * kMhsaDh, q0, b);
      tma_load_3d(sQ + kM2QBytes / 2, &p.tma_t64, q_full, h * kMhsaDh, q0 + 64, b);
      for (int j = 0; j < nkt; ++j) {
        const int slot = j & (kM2Stages - 1);
        const uint32_t par = ((j / kM2Stages) & 1) ^ 1;
        wait_bar(&k_empty[slot], par, 11);
        mbar_expect_tx(&k_full[slot], kM2KvBytes);
        tma_load_3d(sK + slot * kM2KvBytes, &p.tma_t64, &k_full[slot], p.D + h * kMhsaDh, j * kM2KT, b);
        wait_bar(&v_empty[slot], par, 12);
        mbar_expect_tx(&v_full[slot], kM2KvBytes);
        tma_load_3d(sV + slot * kM2KvBytes, &p.tma_t64, &v_full[slot], 2 * p.D + h * kMhsaDh, j * kM2KT, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 1);  // P V : A K-major, B (V) MN-major
      const uint32_t q_base = smem_u32(sQ);
      auto issue_s = [&](int j) {  // S_j = Q K_j^T, 128 x cols x 64
        const int s = j & 1, t = j >> 1, slot = j & (kM2Stages - 1);
        wait_bar(&k_full[slot], (j / kM2Stages) & 1, 14);
        if (t > 0) wait_bar(&s_free[s], (t - 1) & 1, 16);  // S_{j-2} now lives in the softmax warps' registers
        tc_fence_after();
        const uint32_t idesc_s = umma_idesc_f16(128, tile_cols(j), 0, 0);
        const uint32_t k_base = smem_u32(sK + slot * kM2KvBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_base + s * 64, umma_desc_sw128(q_base + k * 32, 1024, 0),
                      umma_desc_sw128(k_base + k * 32, 1024, 0), idesc_s, k != 0);
        umma_commit(&s_full[s]);
        umma_commit(&k_empty[slot]);
      };
      auto issue_pv = [&](int j) {  // O_s += P_j V_j, 128 x 64 x cols
        const int s = j & 1, t = j >> 1, slot = j & (kM2Stages - 1);
        wait_bar(&v_full[slot], (j / kM2Stages) & 1, 17);
        wait_bar(&p_full[s], t & 1, 18);
        tc_fence_after();
        const uint32_t p_base = smem_u32(sP + s * kM2PBytes);
        const uint32_t v_base = smem_u32(sV + slot * kM2KvBytes);
        const int ksteps = tile_cols(j) >> 4;
        for (int k = 0; k < ksteps; ++k)
          umma_f16_ss(tmem_base + 128 + s * 64, umma_desc_sw128(p_base + k * 32, 1024, 0),
                      umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, (t | k) != 0);
        umma_commit(&o_done[s]);
        umma_commit(&v_empty[slot]);
      };
      wait_bar(q_full, 0, 13);
      issue_s(0);
      if (nkt > 1) issue_s(1);
      for (int j = 0; j < nkt; j += 2) {
        if (j + 2 < nkt) issue_s(j + 2);
        if (j + 3 < nkt) issue_s(j + 3);
        issue_pv(j);
        if (j + 1 < nkt) issue_pv(j + 1);
      }
    }
  } else {
    // ===================== softmax warps: stream s, TMEM lane quarter =====================
    const int s = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int q = q0 + r;
    const bool row_active = q0 + quarter * 32 < p.n_tokens;  // warp-uniform
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + s * 64;
    const uint32_t tO = tmem_base + lane_off + 128 + s * 64;
    const float c = p.scale_log2e;
    const int kv_limit = p.causal ? q : 0x7fffffff;
    const int n_s = (nkt - s + 1) >> 1;  // tiles j = s, s + 2, ...
    float m_ref = -INFINITY;  // exponent offset baked into l_run and O_s
    float l_run = 0.f;
    uint8_t* p_row = sP + s * kM2PBytes + r * 128;
    const int sw = r & 7;

    for (int t = 0; t < n_s; ++t) {
      const int j = 2 * t + s;
      const int kv0 = j * kM2KT;
      const int nc = tile_cols(j);
      const bool need_mask = (kv0 + kM2KT > kv_end) || (p.causal && (kv0 + kM2KT - 1 > q0));
      wait_bar(&s_full[s], t & 1, 19);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      if (row_active) {
        __syncwarp();
        tmem_ld32(tS, s0);
        if (nc > 32) tmem_ld32(tS + 32, s1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[s]);  // scores are in registers: the stream's next S MMA may start
      bool any_move = false;
      float factor = 1.f, m_use = 0.f;
      if (row_active) {
        float pm = need_mask ? mhsa_max_chunk<true>(s0, kv0, p.n_tokens, kv_limit)
                             : mhsa_max_chunk<false>(s0, kv0, p.n_tokens, kv_limit);
        if (nc > 32)
          pm = fmaxf(pm, need_mask ? mhsa_max_chunk<true>(s1, kv0 + 32, p.n_tokens, kv_limit)
                                   : mhsa_max_chunk<false>(s1, kv0 + 32, p.n_tokens, kv_limit));
        const float mx = pm * c;
        const bool move = mx > m_ref + kMhsaTau;  // lazy offset: also true on the stream's first unmasked tile
        any_move = __any_sync(0xffffffffu, move);
        if (any_move) {
          const float m_new = move ? mx : m_ref;
          factor = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new);
          l_run *= factor;
          m_ref = m_new;
        }
        m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      }
      // P_s buffer free and O_s quiescent once the stream's previous PV has retired
      if (t > 0) {
        wait_bar(&o_done[s], (t - 1) & 1, 20);
        if (any_move) {  // rescale the TMEM-resident output row (warp-collective; factor = 1 for unmoved rows)
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t o[16];
            __syncwarp();
            tmem_ld16(tO + cc * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st16(tO + cc * 16, o);
          }
          tmem_st_wait();
        }
      }
      if (row_active) {
        // p = exp2(s*c - m_ref) -> fp16 -> 128B-swizzled K-major smem row (8 16-byte slots = 64 keys)
        float l_tile;
        {
          __half2 ph[16];
          l_tile = need_mask ? mhsa_exp_chunk<true>(s0, c, m_use, kv0, p.n_tokens, kv_limit, ph)
                             : mhsa_exp_chunk<false>(s0, c, m_use, kv0, p.n_tokens, kv_limit, ph);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(p_row + ((i ^ sw) * 16)) = *reinterpret_cast<uint4*>(&ph[4 * i]);
        }
        if (nc > 32) {
          __half2 ph[16];
          l_tile += need_mask ? mhsa_exp_chunk<true>(s1, c, m_use, kv0 + 32, p.n_tokens, kv_limit, ph)
                              : mhsa_exp_chunk<false>(s1, c, m_use, kv0 + 32, p.n_tokens, kv_limit, ph);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(p_row + (((4 + i) ^ sw) * 16)) = *reinterpret_cast<uint4*>(&ph[4 * i]);
        }
        l_run += l_tile;
      }
      fence_proxy_async_smem();  // P stores visible to the async (UMMA) proxy ...
      tc_fence_before();
      __syncwarp();              // ... before the warp's single elected arrival
      if (lane == 0) mbar_arrive(&p_full[s]);
    }
    // ---- merge the two streams and write O / l; each warp emits 32 of the 64 output columns ----
    if (n_s > 0) wait_bar(&o_done[s], (n_s - 1) & 1, 25);
    if (row_active) {
      // the stream's P buffer is dead once its last PV has retired: use it to publish (m, l)
      reinterpret_cast<float2*>(sP + s * kM2PBytes)[r] = make_float2(m_ref, l_run);
      tc_fence_before();
      named_bar_sync(1 + quarter, 64);
      tc_fence_after();
      const float2 oth = reinterpret_cast<const float2*>(sP + (s ^ 1) * kM2PBytes)[r];
      const float M = fmaxf(m_ref, oth.x);
      const float fa = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - M);
      const float fo = (oth.x == -INFINITY) ? 0.f : ex2_approx(oth.x - M);
      const float inv = 1.0f / (l_run * fa + oth.y * fo);
      const int n_o = (nkt - (s ^ 1) + 1) >> 1;
      const uint32_t tOa = tmem_base + lane_off + 128 + s * 64 + s * 32;        // own stream, this warp's columns
      const uint32_t tOb = tmem_base + lane_off + 128 + (s ^ 1) * 64 + s * 32;  // other stream, same columns
      uint32_t oa[32], ob[32];
      __syncwarp();
      if (n_s > 0) tmem_ld32(tOa, oa);
      if (n_o > 0) tmem_ld32(tOb, ob);
      tmem_ld_wait();
      const float wa = (n_s > 0) ? fa * inv : 0.f, wb = (n_o > 0) ? fo * inv : 0.f;
      if (q < p.n_tokens) {
        __half* op = p.out + (static_cast<long long>(b) * p.n_tokens + q) * p.D + h * kMhsaDh + s * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __half2 hh[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = g * 8 + 2 * i;
            const float x0 = (n_s > 0 ? __uint_as_float(oa[e]) * wa : 0.f) + (n_o > 0 ? __uint_as_float(ob[e]) * wb : 0.f);
            const float x1 =
                (n_s > 0 ? __uint_as_float(oa[e + 1]) * wa : 0.f) + (n_o > 0 ? __uint_as_float(ob[e + 1]) * wb : 0.f);
            hh[i] = __floats2half2_rn(x0, x1);
          }
          reinterpret_cast<uint4*>(op)[g] = *reinterpret_cast<uint4*>(hh);
        }
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

}  // namespace lseg
