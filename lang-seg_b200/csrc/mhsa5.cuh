// lseg_b200 — fused multi-head self-attention, two-query-tile kernel (FlashAttention-4 style layout).
//
// Contract as mhsa2.cuh (softmax(Q K^T * dh^-0.5) V, head_dim 64, packed [B, N, 3*D] fp16 in, [B*N, D] fp16 out; restates
// timm Attention — SURVEY.md Appendix A.1, modules/models/lseg_vit.py:26-39), NON-causal only (the CLIP text tower has its
// own kernel, text_attn.cuh).
//
// profiles/r02_mhsa_analysis.md: the 64-key, four-streams-per-SM kernels are bound by the length of each stream's
// S -> softmax -> PV chain (two MMA hand-offs per 64 keys on a tensor pipe shared by four streams, N = 64 UMMAs at 66 % of
// the array rate, K/V re-read 8x from L2). This kernel changes the layout instead of the details:
//   * ONE CTA per SM owns TWO 128-row query tiles of one (image, head) and walks the keys in tiles of 128: every K / V
//     tile is loaded once for 256 queries; S = Q K^T is a 128x128x64 MMA (N = 128: full array rate), half the hand-offs
//     per key.
//   * all 512 TMEM columns: S_0 [0,128) S_1 [128,256) | O_0 [256,320) O_1 [320,384) | P_0 [384,448) P_1 [448,512).
//     P (fp16 pairs, 128 keys = 64 columns) has its OWN region, so it never touches shared memory (tcgen05.st; the PV MMA
//     takes its A operand from tensor memory: 128x64x16 at the 32-clk math rate instead of 48 clk of smem operand
//     traffic) AND the next S of a query tile can be issued as soon as the current one is in registers.
//   * the two query tiles ping-pong: while one tile's softmax warpgroup waits for its S or PV, the other one owns the
//     MUFU / FMA pipes; each tile has its own MMA-issuing warp with a fixed order (S_0; per key tile: S_{j+1} once S_j is
//     in registers, then PV_j once P_j is stored), blocking waits only.
//   * softmax arithmetic as mhsa3.cuh: lazy running offset checked per 32-column chunk (O / P rescale only when a row
//     maximum moves by more than 2^8), packed FFMA2 / FADD2, POLYQ of 4 score pairs on the FMA-pipe exp2 polynomial.
// Warps: 0 TMA producer + TMEM alloc; 1 MMA tile 0; 2 MMA tile 1; 3 idle; 4..7 softmax tile 0; 8..11 softmax tile 1
//        (warp & 3 = TMEM lane quarter, thread <-> query row).
// Shared memory: Q_0 Q_1 2 x 16 KB | K ring 4 x 16 KB | V ring 4 x 16 KB | barriers.
#pragma once
#include "common.cuh"
#include "mhsa.cuh"
#include "mhsa3.cuh"
#include "mhsa4.cuh"

namespace lseg {

constexpr int kM5Threads = 384;
constexpr int kM5KT = 128;                        // keys per tile
constexpr int kM5KvBytes = kM5KT * kMhsaDh * 2;   // 16 KB
constexpr int kM5Stages = 4;
constexpr int kM5QBytes = 128 * kMhsaDh * 2;      // 16 KB per query tile
constexpr int kM5SmemBytes = 2 * kM5QBytes + 2 * kM5Stages * kM5KvBytes + 1024;

template <bool PACK, int POLYQ>
__global__ void __launch_bounds__(kM5Threads, 1) mhsa5_kernel(const __grid_constant__ MhsaParams p) {
  auto wait_bar = [](uint64_t* bar, uint32_t parity, int tag) { mbar_wait_inl(bar, parity, tag); };
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                   // [2][16 KB]
  uint8_t* sK = sQ + 2 * kM5QBytes;                     // [stages][16 KB]
  uint8_t* sV = sK + kM5Stages * kM5KvBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kM5Stages * kM5KvBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                          // [4]
  uint64_t* k_empty = bars + 1 + kM5Stages;             // [4]
  uint64_t* v_full = bars + 1 + 2 * kM5Stages;          // [4]
  uint64_t* v_empty = bars + 1 + 3 * kM5Stages;         // [4]
  uint64_t* s_full = bars + 1 + 4 * kM5Stages;          // [2] per query tile
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_full + 4;
  uint64_t* o_done = s_full + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y / p.heads;
  const int h = blockIdx.y % p.heads;
  const int q_base_row = blockIdx.x * 256;
  const bool tile1_valid = q_base_row + 128 < p.n_tokens;  // the second query tile may lie entirely beyond the sequence
  const int n_tiles_q = tile1_valid ? 2 : 1;

  if ((smem_u32(smem) & 1023u) != 0) {  // layout contract of the swizzled tiles
    if (threadIdx.x == 0) atomicCAS(&g_watchdog[0], 0, 99);
    return;
  }

  const int nkt = (p.n_tokens + kM5KT - 1) / kM5KT;
  // columns the S MMA produces / keys the PV MMA consumes for key tile j: valid keys rounded up to the UMMA granule
  auto tile_cols = [&](int j) { return min(kM5KT, ((p.n_tokens - j * kM5KT) + 15) & ~15); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_t64);
    mbar_init(q_full, 1);
    for (int i = 0; i < kM5Stages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], n_tiles_q);  // one tcgen05.commit per query tile that reads the stage
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], n_tiles_q);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);  // one elected arrival per softmax warp of the tile
      mbar_init(&p_full[t], 4);
      mbar_init(&o_done[t], 1);
    }
    mbar_fence_init();
  }
  griddep_launch_dependents();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // the QKV GEMM must have completed before the first TMA load / output store

  if (warp == 0) {
    // ===================== TMA producer (converged warp, elected issuing lane) =====================
    const bool leader = elect_one_sync();
    if (leader) {
      mbar_expect_tx(q_full, n_tiles_q * kM5QBytes);
      for (int t = 0; t < n_tiles_q; ++t) {
        tma_load_3d(sQ + t * kM5QBytes, &p.tma_t64, q_full, h * kMhsaDh, q_base_row + t * 128, b);
        tma_load_3d(sQ + t * kM5QBytes + kM5QBytes / 2, &p.tma_t64, q_full, h * kMhsaDh, q_base_row + t * 128 + 64, b);
      }
    }
    __syncwarp();
    for (int j = 0; j < nkt; ++j) {
      const int slot = j & (kM5Stages - 1);
      const uint32_t par = ((j / kM5Stages) & 1) ^ 1;
      wait_bar(&k_empty[slot], par, 11);
      if (leader) {
        mbar_expect_tx(&k_full[slot], kM5KvBytes);
        tma_load_3d(sK + slot * kM5KvBytes, &p.tma_t64, &k_full[slot], p.D + h * kMhsaDh, j * kM5KT, b);
        tma_load_3d(sK + slot * kM5KvBytes + kM5KvBytes / 2, &p.tma_t64, &k_full[slot], p.D + h * kMhsaDh, j * kM5KT + 64, b);
      }
      __syncwarp();
      wait_bar(&v_empty[slot], par, 12);
      if (leader) {
        mbar_expect_tx(&v_full[slot], kM5KvBytes);
        tma_load_3d(sV + slot * kM5KvBytes, &p.tma_t64, &v_full[slot], 2 * p.D + h * kMhsaDh, j * kM5KT, b);
        tma_load_3d(sV + slot * kM5KvBytes + kM5KvBytes / 2, &p.tma_t64, &v_full[slot], 2 * p.D + h * kMhsaDh, j * kM5KT + 64,
                    b);
      }
      __syncwarp();
    }
  } else if (warp == 1 || (warp == 2 && tile1_valid)) {
    // ===================== MMA issuer of query tile t (converged warp, elected issuing lane) =====================
    const int t = warp - 1;
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 1);  // P V : A (TMEM) K-major, B (V) MN-major
    const uint32_t q_base = smem_u32(sQ + t * kM5QBytes);
    const uint32_t tS = tmem_base + t * 128, tO = tmem_base + 256 + t * 64, tP = tmem_base + 384 + t * 64;
    auto issue_s = [&](int j) {  // S_j = Q_t K_j^T, 128 x cols x 64
      const int slot = j & (kM5Stages - 1);
      wait_bar(&k_full[slot], (j / kM5Stages) & 1, 14);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_f16(128, tile_cols(j), 0, 0);
      const uint32_t k_base = smem_u32(sK + slot * kM5KvBytes);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tS, umma_desc_sw128(q_base + k * 32, 1024, 0), umma_desc_sw128(k_base + k * 32, 1024, 0), idesc_s,
                      k != 0);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[slot]);
      }
      __syncwarp();
    };
    wait_bar(q_full, 0, 13);
    issue_s(0);
    for (int j = 0; j < nkt; ++j) {
      if (j + 1 < nkt) {
        wait_bar(&s_free[t], j & 1, 16);  // S_j lives in the softmax warps' registers now
        issue_s(j + 1);
      }
      const int slot = j & (kM5Stages - 1);
      wait_bar(&v_full[slot], (j / kM5Stages) & 1, 17);
      wait_bar(&p_full[t], j & 1, 18);
      tc_fence_after();
      const uint32_t v_base = smem_u32(sV + slot * kM5KvBytes);
      const int ksteps = tile_cols(j) >> 4;
      if (leader) {
        if (ksteps == 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_ts(tO, tP + k * 8, umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, (j | k) != 0);
        } else {
          for (int k = 0; k < ksteps; ++k)
            umma_f16_ts(tO, tP + k * 8, umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, (j | k) != 0);
        }
        umma_commit(&o_done[t]);
        umma_commit(&v_empty[slot]);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && (warp < 8 || tile1_valid)) {
    // ===================== softmax warps: query tile t, TMEM lane quarter =====================
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int q0 = q_base_row + t * 128;
    const int q = q0 + r;
    const bool row_active = q0 + quarter * 32 < p.n_tokens;  // warp-uniform
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128;
    const uint32_t tO = tmem_base + lane_off + 256 + t * 64;
    const uint32_t tP = tmem_base + lane_off + 384 + t * 64;
    const float c = p.scale_log2e;
    float m_ref = -INFINITY;  // exponent offset baked into l_run, O_t and the stored P
    float l_run = 0.f;

    for (int j = 0; j < nkt; ++j) {
      const int kv0 = j * kM5KT;
      const int nc = tile_cols(j);
      const bool need_mask = kv0 + kM5KT > p.n_tokens;
      wait_bar(&s_full[t], j & 1, 19);
      tc_fence_after();
      bool o_ready = (j == 0);  // PV_{j-1} retired: O_t quiescent and P_t free (waited for once, lazily)
      auto need_o = [&]() {
        if (!o_ready) {
          wait_bar(&o_done[t], (j - 1) & 1, 20);
          tc_fence_after();
          o_ready = true;
        }
      };
      auto rescale_o = [&](float factor) {  // warp-collective; factor = 1 for rows whose offset did not move
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
      auto rescale_p = [&](int chunks, float factor) {  // the tile's already stored P chunks (16 columns each)
        tmem_st_wait();
#pragma unroll 1
        for (int cc = 0; cc < chunks; ++cc) {
          uint32_t pw[16];
          __syncwarp();
          tmem_ld16(tP + cc * 16, pw);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float2 f2 = __half22float2(*reinterpret_cast<__half2*>(&pw[u]));
            const __half2 h2 = __floats2half2_rn(f2.x * factor, f2.y * factor);
            pw[u] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          tmem_st16(tP + cc * 16, pw);
        }
      };
      float l_tile = 0.f;
      // one 32-column chunk: lazy offset check, exponentials, P chunk -> TMEM
      auto do_chunk = [&](const uint32_t (&sc)[32], int chunk) {
        const int kvb = kv0 + chunk * 32;
        const float pm = need_mask ? mhsa_max_chunk<true>(sc, kvb, p.n_tokens, 0x7fffffff)
                                   : (PACK ? mhsa3_max_chunk(sc) : mhsa_max_chunk<false>(sc, kvb, p.n_tokens, 0x7fffffff));
        const float mx = pm * c;
        const bool move = mx > m_ref + kMhsaTau;  // also true on the tile's first chunk of the first key tile
        const bool any = __any_sync(0xffffffffu, move);
        if (any) {
          const float m_new = move ? mx : m_ref;
          const float factor = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new);
          m_ref = m_new;
          l_run *= factor;
          l_tile *= factor;
          if (j > 0) {
            need_o();
            rescale_o(factor);
          }
          if (chunk > 0) rescale_p(chunk, factor);
        }
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
        __half2 ph[16];
        l_tile += need_mask ? mhsa_exp_chunk<true, 0>(sc, c, m_use, kvb, p.n_tokens, 0x7fffffff, ph)
                            : (PACK ? mhsa3_exp_chunk<POLYQ>(sc, c, m_use, ph)
                                    : mhsa_exp_chunk<false, 0>(sc, c, m_use, kvb, p.n_tokens, 0x7fffffff, ph));
        need_o();  // P_t is read by PV_{j-1} until it retires
        __syncwarp();
        tmem_st16(tP + chunk * 16, reinterpret_cast<const uint32_t(&)[16]>(ph));
      };
      // the tile in two halves of 64 columns; S is released to the MMA warp once the second half is in registers
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        const bool have = row_active && (hf * 64 < nc);        // warp-uniform
        const bool have2 = row_active && (hf * 64 + 32 < nc);  // second chunk of the half
        uint32_t sc0[32], sc1[32];
        if (have) {
          __syncwarp();
          tmem_ld32(tS + hf * 64, sc0);
          if (have2) tmem_ld32(tS + hf * 64 + 32, sc1);
          tmem_ld_wait();
        }
        if (hf == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);  // last read of S_j: the tile's next S MMA may start
        }
        if (have) do_chunk(sc0, hf * 2);
        if (have2) do_chunk(sc1, hf * 2 + 1);
      }
      if (row_active) {
        l_run += l_tile;
        tmem_st_wait();
      }
      tc_fence_before();  // P (tcgen05.st) ordered before the hand-off ...
      __syncwarp();       // ... and before the warp's single elected arrival
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // ---- O / l -> out ----
    wait_bar(&o_done[t], (nkt - 1) & 1, 25);
    tc_fence_after();
    if (row_active) {
      const float inv = 1.0f / l_run;
      uint32_t oa[32], ob[32];
      __syncwarp();
      tmem_ld32(tO, oa);
      tmem_ld32(tO + 32, ob);
      tmem_ld_wait();
      if (q < p.n_tokens) {
        __half* op = p.out + (static_cast<long long>(b) * p.n_tokens + q) * p.D + h * kMhsaDh;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          __half2 hh[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = (g & 3) * 8 + 2 * i;
            const float x0 = __uint_as_float(g < 4 ? oa[e] : ob[e]) * inv;
            const float x1 = __uint_as_float(g < 4 ? oa[e + 1] : ob[e + 1]) * inv;
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
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lseg
