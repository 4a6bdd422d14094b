// lseg_b200 — fused multi-head self-attention, round-2 kernel with P in tensor memory (the one the engine runs).
//
// Same contract and two-stream online softmax as mhsa2.cuh / mhsa3.cuh (softmax(Q K^T * dh^-0.5 [+ causal]) V, head_dim 64,
// packed [B, N, 3*D] fp16 in, [B*N, D] fp16 out; restates timm Attention — SURVEY.md Appendix A.1,
// modules/models/lseg_vit.py:26-39).
//
// Why this kernel exists (profiles/r02_mhsa_analysis.md): the round-1 kernel was bound by SHARED-MEMORY BANDWIDTH, not
// by MUFU, issue slots or hand-off latency — which is why cutting each of those in mhsa3.cuh bought nothing. Per 64-key
// tile and stream the SM moved 80 KB through shared memory (128 B/clk): the S MMA reads Q 16 KB + K 8 KB, the softmax
// warps STORE P 16 KB, the PV MMA reads P 16 KB + V 8 KB, TMA writes K and V 16 KB; four streams per SM -> 320 KB =
// 2500 clk of a measured 2800-clk tile period. Here the probabilities never touch shared memory:
//   * P is written with tcgen05.st into the TMEM columns of the stream's own S tile (fp16 pairs: 64 keys = 32 columns,
//     over the 64 fp32 columns of S, once both halves of S are in registers) and the PV MMA takes its A operand from
//     TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc — "TS" form): 48 KB per tile and stream instead of 80, no
//     fence.proxy.async, and a PV MMA that is no longer bound by operand bandwidth (128x64x16 from smem: 6 KB per
//     instruction = 48 clk; with A in TMEM only V's 2 KB come from smem and the MMA runs at its 32-clk math rate).
//   * one MMA-issuing warp per stream with a fixed order: S_0, then per tile PV_j followed by S_{j+2}. tcgen05.mma
//     instructions of one thread execute in issue order, so S_{j+2} may overwrite the S/P columns as soon as PV_j has
//     been ISSUED; no "S consumed" barrier is needed, and since commits also complete in order the softmax warps'
//     wait for PV_{j-2} (O quiescent before a rescale) is already satisfied when S_j arrives: ONE blocking wait per tile.
//   * PACK / POLYQ as in mhsa3.cuh (packed FFMA2 / FADD2 arithmetic; POLYQ of 4 score pairs on the FMA-pipe exp2).
// Warps: 0 TMA producer + TMEM alloc; 1 MMA stream A; 2 MMA stream B; 3 idle; 4..7 softmax A; 8..11 softmax B
//        (warp & 3 = TMEM lane quarter). TMEM (256 columns, 2 CTAs/SM): S_A/P_A [0,64) S_B/P_B [64,128) O_A [128,192)
//        O_B [192,256). Shared memory: Q 16 KB | K ring 4 x 8 KB | V ring 4 x 8 KB | 2 KB merge scratch | barriers.
#pragma once
#include "common.cuh"
#include "mhsa.cuh"
#include "mhsa2.cuh"
#include "mhsa3.cuh"

namespace lseg {

constexpr int kM4ScratchBytes = 2 * 128 * 8;  // (m, l) exchange of the two streams at the end
constexpr int kM4SmemBytes = kM2QBytes + 2 * kM2Stages * kM2KvBytes + kM4ScratchBytes + 1024;

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M x 16, fp16 pairs in 8 consecutive 32-bit columns, row = lane) comes
// from tensor memory.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// PACK: packed arithmetic on unmasked chunks (masked chunks — the tail key tile, the causal diagonal — always take the
// scalar path of mhsa.cuh). POLYQ: see above (only with PACK).
template <bool PACK, int POLYQ>
__global__ void __launch_bounds__(kM3Threads, 2) mhsa4_kernel(const __grid_constant__ MhsaParams p) {
  auto wait_bar = [](uint64_t* bar, uint32_t parity, int tag) { mbar_wait_inl(bar, parity, tag); };
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kM2QBytes;
  uint8_t* sV = sK + kM2Stages * kM2KvBytes;
  uint8_t* sX = sV + kM2Stages * kM2KvBytes;  // merge scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + kM4ScratchBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [4]
  uint64_t* k_empty = bars + 5;   // [4]
  uint64_t* v_full = bars + 9;    // [4]
  uint64_t* v_empty = bars + 13;  // [4]
  uint64_t* s_full = bars + 17;   // [2] per stream
  uint64_t* p_full = bars + 21;   // [2]
  uint64_t* o_done = bars + 23;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = warp_idx_sync();
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
      mbar_init(&p_full[s], 4);
      mbar_init(&o_done[s], 1);
    }
    mbar_fence_init();
  }
  griddep_launch_dependents();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // the QKV GEMM must have completed before the first TMA load / output store

  if (warp < 4) {
    setmaxnreg_dec<kM3CtrlRegs>();
    if (warp == 0) {
      // ===================== TMA producer (converged warp, elected issuing lane) =====================
      const bool leader = elect_one_sync();
      if (leader) {
        mbar_expect_tx(q_full, kM2QBytes);
        tma_load_3d(sQ, &p.tma_t64, q_full, h * kMhsaDh, q0, b);
        tma_load_3d(sQ + kM2QBytes / 2, &p.tma_t64, q_full, h * kMhsaDh, q0 + 64, b);
      }
      __syncwarp();
      for (int j = 0; j < nkt; ++j) {
        const int slot = j & (kM2Stages - 1);
        const uint32_t par = ((j / kM2Stages) & 1) ^ 1;
        wait_bar(&k_empty[slot], par, 11);
        if (leader) {
          mbar_expect_tx(&k_full[slot], kM2KvBytes);
          tma_load_3d(sK + slot * kM2KvBytes, &p.tma_t64, &k_full[slot], p.D + h * kMhsaDh, j * kM2KT, b);
        }
        __syncwarp();
        wait_bar(&v_empty[slot], par, 12);
        if (leader) {
          mbar_expect_tx(&v_full[slot], kM2KvBytes);
          tma_load_3d(sV + slot * kM2KvBytes, &p.tma_t64, &v_full[slot], 2 * p.D + h * kMhsaDh, j * kM2KT, b);
        }
        __syncwarp();
      }
    } else if (warp < 3) {
      // ===================== MMA issuer of stream s (converged warp, elected issuing lane) =====================
      const int s = warp - 1;
      const bool leader = elect_one_sync();
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 1);  // P V : A K-major, B (V) MN-major
      const uint32_t q_base = smem_u32(sQ);
      const uint32_t tS = tmem_base + s * 64, tO = tmem_base + 128 + s * 64;
      const uint32_t tP = tS;  // fp16 P over the first 32 columns of the stream's S tile
      auto issue_s = [&](int j) {  // S_j = Q K_j^T, 128 x cols x 64
        const int slot = j & (kM2Stages - 1);
        wait_bar(&k_full[slot], (j / kM2Stages) & 1, 14);
        tc_fence_after();
        const uint32_t idesc_s = umma_idesc_f16(128, tile_cols(j), 0, 0);
        const uint32_t k_base = smem_u32(sK + slot * kM2KvBytes);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tS, umma_desc_sw128(q_base + k * 32, 1024, 0), umma_desc_sw128(k_base + k * 32, 1024, 0), idesc_s,
                        k != 0);
          umma_commit(&s_full[s]);
          umma_commit(&k_empty[slot]);
        }
        __syncwarp();
      };
      const int n_s = (nkt - s + 1) >> 1;  // tiles j = s, s + 2, ...
      if (n_s > 0) {
        wait_bar(q_full, 0, 13);
        issue_s(s);
        for (int t = 0; t < n_s; ++t) {
          const int j = 2 * t + s;
          const int slot = j & (kM2Stages - 1);
          wait_bar(&v_full[slot], (j / kM2Stages) & 1, 17);
          wait_bar(&p_full[s], t & 1, 18);
          tc_fence_after();
          const uint32_t v_base = smem_u32(sV + slot * kM2KvBytes);
          const int ksteps = tile_cols(j) >> 4;
          if (leader) {
            if (ksteps == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(tO, tP + k * 8, umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, (t | k) != 0);
            } else {
              for (int k = 0; k < ksteps; ++k)
                umma_f16_ts(tO, tP + k * 8, umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, (t | k) != 0);
            }
            umma_commit(&o_done[s]);
            umma_commit(&v_empty[slot]);
          }
          __syncwarp();
          // in issue order behind PV_j: the next S of this stream may overwrite the S / P columns
          if (t + 1 < n_s) issue_s(j + 2);
        }
      }
    }
  } else {
    setmaxnreg_inc<kM3SoftmaxRegs>();
    // ===================== softmax warps: stream s, TMEM lane quarter =====================
    const int s = (warp - 4) >> 2;
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
    const uint32_t tP = tS;  // this row's fp16 probabilities: 32 columns over the S tile

    for (int t = 0; t < n_s; ++t) {
      const int j = 2 * t + s;
      const int kv0 = j * kM2KT;
      const int nc = tile_cols(j);
      const bool need_mask = (kv0 + kM2KT > kv_end) || (p.causal && (kv0 + kM2KT - 1 > q0));
      wait_bar(&s_full[s], t & 1, 19);
      tc_fence_after();
      // (single pass over S, 32 columns at a time, lazy running offset: see mhsa2.cuh)
      auto rescale_o = [&](float factor) {  // warp-collective; factor = 1 for rows whose offset did not move
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 8; ++cc) {
          uint32_t o[8];
          __syncwarp();
          tmem_ld8(tO + cc * 8, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st8(tO + cc * 8, o);
        }
        tmem_st_wait();
      };
      auto move_offset = [&](float pm, float& factor) -> bool {  // returns the warp-uniform "some row moved"
        const float mx = pm * c;
        const bool move = mx > m_ref + kMhsaTau;  // also true on the stream's first unmasked chunk (m_ref = -inf)
        const bool any = __any_sync(0xffffffffu, move);
        factor = 1.f;
        if (any) {
          const float m_new = move ? mx : m_ref;
          factor = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new);
          m_ref = m_new;
        }
        return any;
      };
      auto chunk_max = [&](const uint32_t (&sc)[32], int kvb) -> float {
        if (need_mask) return mhsa_max_chunk<true>(sc, kvb, p.n_tokens, kv_limit);
        if (PACK) return mhsa3_max_chunk(sc);
        return mhsa_max_chunk<false>(sc, kvb, p.n_tokens, kv_limit);
      };
      auto chunk_exp = [&](const uint32_t (&sc)[32], float m_use, int kvb, __half2 (&ph)[16]) -> float {
        if (need_mask) return mhsa_exp_chunk<true, 0>(sc, c, m_use, kvb, p.n_tokens, kv_limit, ph);
        if (PACK) return mhsa3_exp_chunk<POLYQ>(sc, c, m_use, ph);
        return mhsa_exp_chunk<false, 0>(sc, c, m_use, kvb, p.n_tokens, kv_limit, ph);
      };
      __half2 ph[16];
      uint32_t sc1[32];
      float l_tile = 0.f, factor0 = 1.f;
      bool any0 = false;
      const bool second = row_active && nc > 32;  // warp-uniform
      if (row_active) {
        uint32_t sc0[32];
        __syncwarp();
        tmem_ld32(tS, sc0);
        tmem_ld_wait();
        const float pm = chunk_max(sc0, kv0);
        any0 = move_offset(pm, factor0);
        if (any0) l_run *= factor0;
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
        l_tile = chunk_exp(sc0, m_use, kv0, ph);
        __syncwarp();
        tmem_ld32(tS + 32, sc1);  // unconditional (columns beyond a short tail tile are stale but allocated)
        tmem_ld_wait();
      }
      // Both halves of S_j are in registers: its columns may now receive P_j. The O rescale needs the stream's previous
      // PV retired; its commit completes before the commit of S_j (in-order), so this wait never blocks.
      if (t > 0) {
        wait_bar(&o_done[s], (t - 1) & 1, 20);
        if (any0) rescale_o(factor0);
      }
      if (row_active) {
        // fp16 P row -> tensor memory: keys 0..31 of the tile = 16 packed columns
        __syncwarp();
        tmem_st16(tP, reinterpret_cast<const uint32_t(&)[16]>(ph));
        if (second) {
          const float pm = chunk_max(sc1, kv0 + 32);
          float factor1;
          if (move_offset(pm, factor1)) {  // rare: the first half of this tile was exponentiated against the old offset
            l_run *= factor1;
            l_tile *= factor1;
            if (t > 0) rescale_o(factor1);
            uint32_t pw[16];  // re-scale the first half of P in place
            tmem_st_wait();
            __syncwarp();
            tmem_ld16(tP, pw);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float2 f2 = __half22float2(*reinterpret_cast<__half2*>(&pw[u]));
              const __half2 h2 = __floats2half2_rn(f2.x * factor1, f2.y * factor1);
              pw[u] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            __syncwarp();
            tmem_st16(tP, pw);
          }
          const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
          l_tile += chunk_exp(sc1, m_use, kv0 + 32, ph);
          __syncwarp();
          tmem_st16(tP + 16, reinterpret_cast<const uint32_t(&)[16]>(ph));
        }
        l_run += l_tile;
        tmem_st_wait();
      }
      tc_fence_before();         // P (tcgen05.st) ordered before the hand-off ...
      __syncwarp();              // ... and before the warp's single elected arrival
      if (lane == 0) mbar_arrive(&p_full[s]);
    }
    // ---- merge the two streams and write O / l; each warp emits 32 of the 64 output columns ----
    if (n_s > 0) wait_bar(&o_done[s], (n_s - 1) & 1, 25);
    if (row_active) {
      reinterpret_cast<float2*>(sX + s * (kM4ScratchBytes / 2))[r] = make_float2(m_ref, l_run);  // publish (m, l)
      tc_fence_before();
      named_bar_sync(1 + quarter, 64);
      tc_fence_after();
      const float2 oth = reinterpret_cast<const float2*>(sX + (s ^ 1) * (kM4ScratchBytes / 2))[r];
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
