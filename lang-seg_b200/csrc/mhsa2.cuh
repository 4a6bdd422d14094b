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
//   * S is read from TMEM once, 32 columns at a time; the second half is fetched (which releases S for the
//     stream's next S MMA) before the wait for the previous PV, so both hand-offs hide under exponentials.
//   * the running offset is lazy and checked per 32-column chunk on the registers that feed the exponentials:
//     no separate max pass (it cost a second trip to TMEM per tile).
//   * the last key tile only computes ceil16(valid keys) columns (S MMA with N = 16.., PV with K = 16..):
//     901 tokens = 14 tiles of 64 + 5 keys instead of 8 x 128 (11 % fewer exponentials and MMAs).
//   * row groups of 32 that lie entirely beyond the sequence (the last query tile has 5 valid rows of 128)
//     skip the softmax arithmetic: their P rows are garbage, which only reaches their own (unstored) O rows.
//
// Warps: 0 TMA producer (Q; K and V 64-key tiles through 4-deep rings) + TMEM alloc; 1 MMA issuer: one warp
//        multiplexes both streams by polling the hand-off barriers, PV first (both run converged with an elected
//        issuing lane: from a divergent single lane a 64-row TMA box cost 300-500 clk to issue, a UMMA 130);
//        2..5 softmax stream A;
//        6..9 softmax stream B (warp & 3 = TMEM lane quarter).
// TMEM (256 columns, 2 CTAs/SM): S_A [0,64) S_B [64,128) O_A [128,192) O_B [192,256).
#pragma once
#include "common.cuh"
#include "mhsa.cuh"

namespace lseg {

constexpr int kM2Threads = 320;
constexpr int kMhsaPolyDefault = 0;  // of every 8 exponentials on the FMA pipe (LSEG_MHSA_POLY=2|3|4: measured slower)
constexpr int kM2KT = 64;                      // keys per tile
constexpr int kM2KvBytes = kM2KT * kMhsaDh * 2;  // 8 KB
constexpr int kM2Stages = 4;
constexpr int kM2QBytes = 128 * kMhsaDh * 2;   // 16 KB
constexpr int kM2PBytes = 128 * kM2KT * 2;     // 16 KB per stream
// Q | K ring | V ring | P_A P_B | mbarriers + tmem slot
constexpr int kM2SmemBytes = kM2QBytes + 2 * kM2Stages * kM2KvBytes + 2 * kM2PBytes + 1024;

// POLY: how many of every 8 exponentials run on the FMA pipe (mhsa_exp_chunk).
// TRACE: debug instantiation that stamps clock64() at the hand-off points of 16 sampled CTAs into p.trace
// ([16 CTAs][10 warps][256] of (clock << 8 | tag)); tools/mhsa_trace.py prints the timeline.
template <int POLY, bool TRACE = false>
__global__ void __launch_bounds__(kM2Threads, 2) mhsa2_kernel(const __grid_constant__ MhsaParams p) {
  auto wait_bar = [](uint64_t* bar, uint32_t parity, int tag) {
    mbar_wait_inl(bar, parity, tag);
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

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int b = blockIdx.y / p.heads;
  const int h = blockIdx.y % p.heads;
  const int q0 = q_tile * 128;
  int trace_n = 0;
  unsigned long long* trace_w = nullptr;
  if (TRACE) {
    const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
    const int slot = (cta_lin % 128 == 0) ? cta_lin / 128 : ((cta_lin % 128 == 7) ? 8 + cta_lin / 128 : -1);
    if (slot >= 0 && slot < 16 && lane == 0) trace_w = p.trace + (slot * 10 + warp) * 256;
  }
  auto TR = [&](int tag) {
    if (TRACE && trace_w && trace_n < 254) trace_w[trace_n++] = (static_cast<unsigned long long>(clock64()) << 8) | tag;
  };

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
  TR(0);

  if (warp == 0) {
    // ===================== TMA producer (converged warp, elected issuing lane) =====================
    {
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
        TR(20);
        wait_bar(&v_empty[slot], par, 12);
        if (leader) {
          mbar_expect_tx(&v_full[slot], kM2KvBytes);
          tma_load_3d(sV + slot * kM2KvBytes, &p.tma_t64, &v_full[slot], 2 * p.D + h * kMhsaDh, j * kM2KT, b);
        }
        __syncwarp();
        TR(21);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      const bool leader = elect_one_sync();  // the one lane that issues (and commits) every MMA of this CTA
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 1);  // P V : A K-major, B (V) MN-major
      const uint32_t q_base = smem_u32(sQ);
      auto issue_s = [&](int j) {  // S_j = Q K_j^T, 128 x cols x 64 (operands known to be ready)
        const int s = j & 1, slot = j & (kM2Stages - 1);
        const uint32_t idesc_s = umma_idesc_f16(128, tile_cols(j), 0, 0);
        const uint32_t k_base = smem_u32(sK + slot * kM2KvBytes);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + s * 64, umma_desc_sw128(q_base + k * 32, 1024, 0),
                        umma_desc_sw128(k_base + k * 32, 1024, 0), idesc_s, k != 0);
          umma_commit(&s_full[s]);
          umma_commit(&k_empty[slot]);
        }
        __syncwarp();
        TR(11);
      };
      auto issue_pv = [&](int j) {  // O_s += P_j V_j, 128 x 64 x cols
        const int s = j & 1, t = j >> 1, slot = j & (kM2Stages - 1);
        const uint32_t p_base = smem_u32(sP + s * kM2PBytes);
        const uint32_t v_base = smem_u32(sV + slot * kM2KvBytes);
        const int ksteps = tile_cols(j) >> 4;
        if (leader) {
          if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmem_base + 128 + s * 64, umma_desc_sw128(p_base + k * 32, 1024, 0),
                          umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, (t | k) != 0);
          } else {
            for (int k = 0; k < ksteps; ++k)
              umma_f16_ss(tmem_base + 128 + s * 64, umma_desc_sw128(p_base + k * 32, 1024, 0),
                          umma_desc_sw128(v_base + k * 2048, 1024, 8192), idesc_o, (t | k) != 0);
          }
          umma_commit(&o_done[s]);
          umma_commit(&v_empty[slot]);
        }
        __syncwarp();
        TR(13);
      };
      wait_bar(q_full, 0, 13);
      // One thread multiplexes both streams: whatever is ready is issued, PV first (it is on the streams'
      // critical path: the next P store waits for it; the next S is only needed a tile later).
      int s_next[2] = {0, 1}, pv_next[2] = {0, 1};
      int remaining = 2 * nkt;
      const long long t_start = clock64();
      uint32_t idle = 0;
      while (remaining > 0) {
        bool did = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int j = pv_next[s];
          if (j < nkt && j < s_next[s] && mbar_test(&p_full[s], (j >> 1) & 1) &&
              mbar_test(&v_full[j & (kM2Stages - 1)], (j / kM2Stages) & 1)) {
            TR(31);
            tc_fence_after();
            TR(12);
            issue_pv(j);
            pv_next[s] = j + 2;
            --remaining;
            did = true;
          }
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int j = s_next[s];
          if (j < nkt && (j < 2 || mbar_test(&s_free[s], ((j >> 1) - 1) & 1)) &&
              mbar_test(&k_full[j & (kM2Stages - 1)], (j / kM2Stages) & 1)) {
            tc_fence_after();
            TR(10);
            issue_s(j);
            s_next[s] = j + 2;
            --remaining;
            did = true;
          }
        }
        if (TRACE && !did) {
          int mask = 0;
          for (int s = 0; s < 2; ++s) {
            const int j = pv_next[s];
            if (j < nkt) {
              mask |= (mbar_test(&p_full[s], (j >> 1) & 1) ? 1 : 0) << (2 * s);
              mask |= (mbar_test(&v_full[j & (kM2Stages - 1)], (j / kM2Stages) & 1) ? 2 : 0) << (2 * s);
            }
          }
          TR(32 + mask);
        }
        if (!did) {
          // park briefly on the older pending P hand-off instead of burning the SMSP's issue slots
          const int sp = (pv_next[0] <= pv_next[1]) ? 0 : 1;
          const int jp = pv_next[sp];
          if (jp < nkt) mbar_try_wait_ns(&p_full[sp], (jp >> 1) & 1, 100);
          if ((++idle & 0x3FFu) == 0) {
            if (*reinterpret_cast<volatile int*>(&g_watchdog[0]) != 0) break;
            if (clock64() - t_start > kWatchdogCycles) {
              if (atomicCAS(&g_watchdog[0], 0, 18) == 0) {
                g_watchdog[1] = blockIdx.x;
                g_watchdog[2] = threadIdx.x;
                g_watchdog[3] = remaining;
              }
              break;
            }
          }
        }
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
      TR(1);
      // Single pass over S, one 32-column chunk at a time (holding all 64 scores plus the fp16 results would spill at
      // the 96 registers that 2 CTAs/SM allow, and spills are ruinous here: the smem carve-out leaves almost no
      // L1). The running offset m_ref is LAZY and is checked per chunk on the registers that feed the exponentials
      // anyway: it only moves when a chunk's row max exceeds it by more than 2^kMhsaTau (first chunk of the stream,
      // then rarely), in which case l, the TMEM-resident O row and — for the second chunk — the already
      // stored first half of P are rescaled. No separate max pass, no second trip to TMEM.
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
        TR(40);
        const float pm = need_mask ? mhsa_max_chunk<true>(sc0, kv0, p.n_tokens, kv_limit)
                                   : mhsa_max_chunk<false>(sc0, kv0, p.n_tokens, kv_limit);
        any0 = move_offset(pm, factor0);
        if (any0) l_run *= factor0;
        TR(41);
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
        l_tile = need_mask ? mhsa_exp_chunk<true, POLY>(sc0, c, m_use, kv0, p.n_tokens, kv_limit, ph)
                           : mhsa_exp_chunk<false, POLY>(sc0, c, m_use, kv0, p.n_tokens, kv_limit, ph);
        if (TRACE) {  // make the stamp wait for the exponentials (the packed halves are consumed much later)
          asm volatile("" ::"r"(*reinterpret_cast<uint32_t*>(&ph[15])) : "memory");
        }
        TR(42);
        __syncwarp();
        tmem_ld32(tS + 32, sc1);  // unconditional (columns beyond a short tail tile are stale but allocated)
        tmem_ld_wait();
        TR(43);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[s]);  // last read of S_j: the stream's next S MMA may start
      TR(3);
      // only the P store and the O rescale need the stream's previous PV retired (P buffer free, O quiescent)
      if (t > 0) {
        wait_bar(&o_done[s], (t - 1) & 1, 20);
        if (any0) rescale_o(factor0);
      }
      TR(4);
      if (row_active) {
        // fp16 P row -> 128B-swizzled K-major smem (8 16-byte slots = 64 keys)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(p_row + ((i ^ sw) * 16)) = *reinterpret_cast<uint4*>(&ph[4 * i]);
        if (second) {
          const float pm = need_mask ? mhsa_max_chunk<true>(sc1, kv0 + 32, p.n_tokens, kv_limit)
                                     : mhsa_max_chunk<false>(sc1, kv0 + 32, p.n_tokens, kv_limit);
          float factor1;
          if (move_offset(pm, factor1)) {  // rare: the first half of this tile was exponentiated against the old offset
            l_run *= factor1;
            l_tile *= factor1;
            if (t > 0) rescale_o(factor1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4* slot = reinterpret_cast<uint4*>(p_row + ((i ^ sw) * 16));
              uint4 q = *slot;
              __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 f2 = __half22float2(h[u]);
                h[u] = __floats2half2_rn(f2.x * factor1, f2.y * factor1);
              }
              *slot = q;
            }
          }
          const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
          l_tile += need_mask ? mhsa_exp_chunk<true, POLY>(sc1, c, m_use, kv0 + 32, p.n_tokens, kv_limit, ph)
                              : mhsa_exp_chunk<false, POLY>(sc1, c, m_use, kv0 + 32, p.n_tokens, kv_limit, ph);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(p_row + (((4 + i) ^ sw) * 16)) = *reinterpret_cast<uint4*>(&ph[4 * i]);
        }
        l_run += l_tile;
      }
      TR(5);
      fence_proxy_async_smem();  // P stores visible to the async (UMMA) proxy ...
      tc_fence_before();
      __syncwarp();              // ... before the warp's single elected arrival
      if (lane == 0) mbar_arrive(&p_full[s]);
      TR(6);
    }
    // ---- merge the two streams and write O / l; each warp emits 32 of the 64 output columns ----
    if (n_s > 0) wait_bar(&o_done[s], (n_s - 1) & 1, 25);
    TR(7);
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

  TR(8);
  tc_fence_before();
  __syncthreads();
  TR(99);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace lseg
