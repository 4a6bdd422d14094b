// lseg_b200 — fused multi-head self-attention, round-2 kernel (the one the engine runs for the ViT trunk).
//
// Same contract and the same two-stream online softmax as mhsa2.cuh (softmax(Q K^T * dh^-0.5 [+ causal]) V, head_dim
// 64, packed [B, N, 3*D] fp16 in, [B*N, D] fp16 out; restates timm Attention — SURVEY.md Appendix A.1,
// modules/models/lseg_vit.py:26-39). What changed against mhsa2.cuh and why (profiles/r01_mhsa_full.md: tensor pipe
// 24 %, MUFU ~50 %, a 64-key tile of one stream took ~3000 clk of which ~950 clk were the softmax warps WAITING for an
// MMA hand-off):
//   * one MMA-issuing warp PER STREAM, each walking its stream's fixed issue order (S_0, then per tile: S_{j+2} as soon
//     as S_j has been read and K_{j+2} has landed, then P_j V_j) with BLOCKING mbarrier waits. mhsa2 multiplexed both
//     streams from one warp by polling four barriers with test_wait (~150 clk each): a hand-off was noticed half a
//     polling round (~300 clk) late, twice per tile.
//   * 12 warps instead of 10: the launch bound (384 threads x 2 CTAs/SM) is 80 registers per thread, and by role
//     (setmaxnreg) the control warpgroup (TMA producer, 2 MMA warps) drops to 40 while the two softmax warpgroups may
//     rise to 96. With that head-room ptxas allocates the softmax code without a single spill (and in 78 registers);
//     under a flat 80-register cap it spills 12-16 bytes in two instantiations, which is ruinous here (the shared-memory
//     carve-out leaves almost no L1 for local memory).
//   * PACK: the softmax arithmetic uses the sm_100 packed-fp32 instructions (FFMA2 / FADD2: two lanes of fp32 per
//     issue slot) and the 3-input FMNMX3 for the row maximum: 6 issue slots per pair of scores instead of 9. The
//     softmax warps are issue-bound as much as MUFU-bound (4 softmax warps per SM sub-partition share one issue port
//     and one 4-lane MUFU), so
//   * POLYQ of every 4 score pairs take their 2^x from a degree-3 polynomial on the FMA pipe (Cody-Waite split,
//     exponent spliced in with one LEA; relative error 7.5e-5, below the fp16 rounding of P) instead of MUFU.EX2.
//     In mhsa2 this was slower (the extra instructions cost more issue slots than the MUFU time they saved); with the
//     packed forms the balance tips.
// Warps: 0 TMA producer + TMEM alloc; 1 MMA stream A; 2 MMA stream B; 3 idle; 4..7 softmax A; 8..11 softmax B
//        (warp & 3 = TMEM lane quarter). TMEM (256 columns, 2 CTAs/SM): S_A [0,64) S_B [64,128) O_A [128,192) O_B [192,256).
#pragma once
#include "common.cuh"
#include "mhsa.cuh"
#include "mhsa2.cuh"

namespace lseg {

constexpr int kM3Threads = 384;
constexpr int kM3CtrlRegs = 40;     // pool arithmetic per CTA: 128 x (80 - 40) = 5120 released >= 256 x (96 - 80) = 4096 claimed
constexpr int kM3SoftmaxRegs = 96;

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// row maximum of one unmasked 32-column chunk: 16 FMNMX3 in four chains + 2 to combine
__device__ __forceinline__ float mhsa3_max_chunk(const uint32_t (&s)[32]) {
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
#pragma unroll
    for (int u = 0; u < 4; ++u) m[u] = fmax3(m[u], __uint_as_float(s[i + 2 * u]), __uint_as_float(s[i + 2 * u + 1]));
  }
  return fmaxf(fmax3(m[0], m[1], m[2]), m[3]);
}

// 2^x for a pair on the FMA pipe: exp2_poly<3> (common.cuh) in packed form. x <= ~8 by construction of the lazy
// offset; the lower clamp keeps the spliced exponent in range (results below 2^-126 would be flushed anyway).
__device__ __forceinline__ float2 exp2_poly3_x2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 r = __fadd2_rn(x, magic);
  const float2 t = __fadd2_rn(r, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(t, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(f, make_float2(5.517166885e-02f, 5.517166885e-02f), make_float2(2.426111221e-01f, 2.426111221e-01f));
  p = __ffma2_rn(p, f, make_float2(6.932609855e-01f, 6.932609855e-01f));
  p = __ffma2_rn(p, f, make_float2(9.999280736e-01f, 9.999280736e-01f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23)));
}

// exp2(s*c - m) for one UNMASKED 32-column chunk -> fp16 pairs + fp32 partial row sum, packed arithmetic.
// POLYQ of every 4 pairs use the polynomial.
template <int POLYQ>
__device__ __forceinline__ float mhsa3_exp_chunk(const uint32_t (&s)[32], float c, float m, __half2 (&ph)[16]) {
  float2 sum = make_float2(0.f, 0.f);
  const float2 c2 = make_float2(c, c), nm2 = make_float2(-m, -m);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), c2, nm2);
    float2 p;
    if ((i & 3) < POLYQ) {
      p = exp2_poly3_x2(x);
    } else {
      p.x = ex2_approx(x.x);
      p.y = ex2_approx(x.y);
    }
    ph[i] = __floats2half2_rn(p.x, p.y);
    sum = __fadd2_rn(sum, p);
  }
  return sum.x + sum.y;
}

// PACK: packed arithmetic on unmasked chunks (masked chunks — the tail key tile, the causal diagonal — always take the
// scalar path of mhsa.cuh). POLYQ: see above (only with PACK).
template <bool PACK, int POLYQ>
__global__ void __launch_bounds__(kM3Threads, 2) mhsa3_kernel(const __grid_constant__ MhsaParams p) {
  auto wait_bar = [](uint64_t* bar, uint32_t parity, int tag) { mbar_wait_inl(bar, parity, tag); };
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
      mbar_init(&s_free[s], 4);  // one elected arrival per softmax warp of the stream
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
      const uint32_t p_base = smem_u32(sP + s * kM2PBytes);
      const uint32_t tS = tmem_base + s * 64, tO = tmem_base + 128 + s * 64;
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
          if (t + 1 < n_s) {
            wait_bar(&s_free[s], t & 1, 16);  // S_j lives in the softmax warps' registers now
            issue_s(j + 2);
          }
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
                umma_f16_ss(tO, umma_desc_sw128(p_base + k * 32, 1024, 0), umma_desc_sw128(v_base + k * 2048, 1024, 8192),
                            idesc_o, (t | k) != 0);
            } else {
              for (int k = 0; k < ksteps; ++k)
                umma_f16_ss(tO, umma_desc_sw128(p_base + k * 32, 1024, 0), umma_desc_sw128(v_base + k * 2048, 1024, 8192),
                            idesc_o, (t | k) != 0);
            }
            umma_commit(&o_done[s]);
            umma_commit(&v_empty[slot]);
          }
          __syncwarp();
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
    uint8_t* p_row = sP + s * kM2PBytes + r * 128;
    const int sw = r & 7;

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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[s]);  // last read of S_j: the stream's next S MMA may start
      // only the P store and the O rescale need the stream's previous PV retired (P buffer free, O quiescent)
      if (t > 0) {
        wait_bar(&o_done[s], (t - 1) & 1, 20);
        if (any0) rescale_o(factor0);
      }
      if (row_active) {
        // fp16 P row -> 128B-swizzled K-major smem (8 16-byte slots = 64 keys)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(p_row + ((i ^ sw) * 16)) = *reinterpret_cast<uint4*>(&ph[4 * i]);
        if (second) {
          const float pm = chunk_max(sc1, kv0 + 32);
          float factor1;
          if (move_offset(pm, factor1)) {  // rare: the first half of this tile was exponentiated against the old offset
            l_run *= factor1;
            l_tile *= factor1;
            if (t > 0) rescale_o(factor1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4* slot = reinterpret_cast<uint4*>(p_row + ((i ^ sw) * 16));
              uint4 qv = *slot;
              __half2* hh = reinterpret_cast<__half2*>(&qv);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 f2 = __half22float2(hh[u]);
                hh[u] = __floats2half2_rn(f2.x * factor1, f2.y * factor1);
              }
              *slot = qv;
            }
          }
          const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
          l_tile += chunk_exp(sc1, m_use, kv0 + 32, ph);
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
