// lseg_b200 — shared device helpers for the sm_100a kernels.
//
// Thin inline-PTX wrappers (mbarrier, TMA, tcgen05/TMEM), UMMA descriptor builders and the
// host-side tensor-map factory. Everything here is written for sm_100a only; there is no
// fallback path.
#pragma once
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace lseg {

// ------------------------------------------------------------------------------------------
// Device-side watchdog: a barrier wait that exceeds kWatchdogCycles records (tag, block, role)
// in g_watchdog and makes every later wait in the grid fall through, so a protocol bug shows
// up as a reported error code instead of a hung GPU box.
// ------------------------------------------------------------------------------------------
constexpr long long kWatchdogCycles = 1ll << 31;  // ~1 s at 1.9 GHz
__device__ int g_watchdog[4];                     // [0]=tag (0 = ok), [1]=blockIdx.x, [2]=threadIdx.x, [3]=parity

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// suspend-time hint: the waiting thread sleeps in hardware until the phase completes or this many ns
// pass, instead of re-polling every ~100 cycles and stealing issue slots from the compute warps that
// share its SM sub-partition (ncu: 3.1 M TRYWAIT re-issues per MHSA launch without the hint).
constexpr uint32_t kTryWaitHintNs = 20000;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kTryWaitHintNs)
      : "memory");
  return ok != 0;
}
// Non-blocking phase test (for a thread that multiplexes several barriers), and a try_wait whose hardware
// suspend is bounded by `ns` instead of kTryWaitHintNs.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_ns(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, int tag) {
  // try_wait suspends the thread in hardware for a bounded time, so this loop is not a hot spin.
  // The watchdog bookkeeping (a global-memory flag read, ~1 us) runs only every 256 failed polls so
  // that it never sits on the critical path of a healthy pipeline.
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 0xFFu) != 0) continue;
    if (*reinterpret_cast<volatile int*>(&g_watchdog[0]) != 0) return;
    if (clock64() - t0 > kWatchdogCycles) {
      if (atomicCAS(&g_watchdog[0], 0, tag) == 0) {
        g_watchdog[1] = blockIdx.x;
        g_watchdog[2] = threadIdx.x;
        g_watchdog[3] = static_cast<int>(parity);
      }
      return;
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, tag);
}
// Fully inlined wait (hinted try_wait): for waits that sit between long-lived register state — the out-of-line
// mbar_wait_slow is a real call, and live registers around it are spilled per the ABI.
__device__ __forceinline__ void mbar_wait_inl(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 0xFFu) != 0) continue;
    if (*reinterpret_cast<volatile int*>(&g_watchdog[0]) != 0) return;
    if (clock64() - t0 > kWatchdogCycles) {
      if (atomicCAS(&g_watchdog[0], 0, tag) == 0) {
        g_watchdog[1] = blockIdx.x;
        g_watchdog[2] = threadIdx.x;
        g_watchdog[3] = static_cast<int>(parity);
      }
      return;
    }
  }
}
// Latency-critical variant: try_wait WITHOUT a suspend-time hint (the hardware re-checks after its short
// default window instead of parking the thread), for waits that sit on a kernel's critical path.
__device__ __forceinline__ bool mbar_try_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait_spin(bar, parity)) {
    if ((++polls & 0xFFFu) != 0) continue;
    if (*reinterpret_cast<volatile int*>(&g_watchdog[0]) != 0) return;
    if (clock64() - t0 > kWatchdogCycles) {
      if (atomicCAS(&g_watchdog[0], 0, tag) == 0) {
        g_watchdog[1] = blockIdx.x;
        g_watchdog[2] = threadIdx.x;
        g_watchdog[3] = static_cast<int>(parity);
      }
      return;
    }
  }
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// Every per-step kernel is launched with programmaticStreamSerializationAllowed: its CTAs may become resident and
// run their prologue (barrier init, TMEM alloc, tensor-map prefetch) while the previous kernel of the stream is
// still draining. griddep_wait() blocks until that previous grid has completed and its writes are visible; it
// must precede the first access to any global memory another kernel of the step reads or writes.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One lane of a converged warp (elect.sync): the MMA / TMA warps run converged with warp-uniform values (which the
// compiler then keeps in uniform registers) and only predicate the tcgen05 / bulk-copy instruction itself on this
// — a `if (lane == 0) { ... }` region makes every descriptor operand go through an ELECT + R2UR waterfall
// (~20 dependent instructions, ~130 clk per tcgen05.mma measured).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int warp_idx_sync() { return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0); }

// ---- proxies / fences --------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA stores: smem tile -> global through a tensor map (clips rows/columns outside the tensor).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// global[tile] += smem tile, performed by the TMA unit at L2 (element type from the tensor map: fp32)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their smem source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM --------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns; thread t of the warp receives lane (base_lane + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same 32 lanes x 32 columns shape as tmem_ld32
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 8- and 16-column variants (keep register pressure low on rarely taken paths)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA (tcgen05.mma) ------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle (layout_type 2 at bits [61,64), descriptor
// version 1 at bits [46,48)). Addresses / offsets are encoded >>4. Tiles must be 1024 B aligned.
//   K-major operand (rows of 64 halves = 128 B): SBO = 1024 B between 8-row groups, LBO unused.
//   MN-major operand (128 B along MN per K index): SBO = 1024 B between 8-k groups,
//   LBO = byte stride between 64-element MN chunks.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor for kind::f16 with fp16 A/B and fp32 accumulate.
//   bits [4,6) c_format=1 (F32); [7,10) a_format=0 (F16); [10,13) b_format=0 (F16);
//   bit 15 a_major, bit 16 b_major (0 = K-major, 1 = MN-major); [17,23) N>>3; [24,29) M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- CTA-pair (cta_group::2) variants ------------------------------------------------------
// Two CTAs of a cluster (ranks 2i, 2i+1 — one TPC) cooperate on one MMA: M = 256 rows split 128/128,
// each CTA stages its A half and half of the B tile; the even ("leader") CTA issues the MMA.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (same smem offset).
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once the pair's MMAs retire) on the mbarrier at this smem offset in every CTA of cta_mask.
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// Same, without release semantics: for hand-offs whose payload is TMEM (ordered by tcgen05.fence) — the default
// .release.cluster arrive drains the warp's earlier global/shared writes first (~1.6k clk after TMA-store epilogues).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// Arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

// ---- misc math ---------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x);
// Exact-erf GELU, 0.5 x (1 + erf(x / sqrt 2)) (timm Mlp / ProjectReadout use nn.GELU()), with erf from
// 2^x on the FMA/ALU pipes (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5] via the 1.5*2^23 rounding trick,
// minimax polynomial for 2^f (tools/fit_exp2.py: relative error 7.5e-5 / 2.7e-6 / 2.3e-7 for degree 3 / 4 / 5),
// exponent added by integer arithmetic on the bit pattern: ~9 issue slots instead of one MUFU instruction (8 clk
// of the MUFU pipe per warp, tools/probes/mufu_probe.cu). Kept as a measured experiment: neither the attention
// softmax (LSEG_MHSA_POLY=2|3|4: 64 -> 72-77 us) nor the GELU epilogue got faster — both are bound by issue
// slots / latency of their few warps, not by MUFU throughput. Requires -126 <= x <= 126 (the callers clamp).
template <int DEG>
__device__ __forceinline__ float exp2_poly(float x) {
  const float r = x + 12582912.f;  // 1.5 * 2^23: the low mantissa bits of r now hold round(x)
  const float f = x - (r - 12582912.f);
  float p;
  if (DEG == 3) {
    p = fmaf(f, 5.517166885e-02f, 2.426111221e-01f);
    p = fmaf(p, f, 6.932609855e-01f);
    p = fmaf(p, f, 9.999280736e-01f);
  } else if (DEG == 4) {
    p = fmaf(f, 9.570101897e-03f, 5.591786033e-02f);
    p = fmaf(p, f, 2.402474483e-01f);
    p = fmaf(p, f, 6.931218147e-01f);
    p = fmaf(p, f, 9.999992614e-01f);
  } else {
    p = fmaf(f, 1.327647190e-03f, 9.675541334e-03f);
    p = fmaf(p, f, 5.550713274e-02f);
    p = fmaf(p, f, 2.402211972e-01f);
    p = fmaf(p, f, 6.931469671e-01f);
    p = fmaf(p, f, 1.000000072e+00f);
  }
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// Exact-erf GELU x * Phi(x) with Phi(x) = 1 / (1 + 2^(x * P(x^2))): P is a degree-4 minimax fit (tools/fit_gelu.py:
// reweighted least squares against scipy erfc on [-7, 7], coefficients carry the -log2(e)); |gelu error| <= 3.7e-6 over
// all x in fp32 arithmetic, i.e. below the fp16 rounding of the stored activation wherever that activation is
// not already in fp16's subnormal range. 10 issue slots (2 MUFU) instead of the 20 of the Abramowitz-Stegun
// 7.1.26 form used before (the fc1 epilogue is issue-bound). P(u) < 0 for all u >= 0, so the sigmoid saturates
// monotonically outside the fitted range.
__device__ __forceinline__ float gelu_erf(float x) {
  const float u = x * x;
  float p = fmaf(u, -3.228989445e-06f, 8.823814435e-05f);
  p = fmaf(p, u, 3.602744768e-04f);
  p = fmaf(p, u, -1.052266877e-01f);
  p = fmaf(p, u, -2.302045391e+00f);
  // (exp2_poly<4> here instead of the MUFU ex2 was measured SLOWER: fc1 51.8 -> 58.9 us; the epilogue is bound by
  // issue slots of its 8 warps, not by the MUFU pipe)
  const float e = ex2_approx(x * p);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
// CLIP QuickGELU evaluated the way the fp16 reference does it: h = half(x); half(1.702*h);
// half(sigmoid(.)); half(h * .)  (x * torch.sigmoid(1.702 * x) on a HalfTensor).
__device__ __forceinline__ float quick_gelu(float x) {
  const float h = __half2float(__float2half_rn(x));
  const float t = __half2float(__float2half_rn(1.702f * h));
  const float s = __half2float(__float2half_rn(1.0f / (1.0f + __expf(-t))));
  return h * s;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
// Thread-local last-error string surfaced through lseg_last_error().
void set_error(const char* fmt, ...);
const char* get_error();

#define LSEG_CHECK_CUDA(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::lseg::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

// Kernel launch with the PDL attribute (see griddep_wait above); LSEG_NO_PDL=1 falls back to plain launches.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  static const bool no_pdl = getenv("LSEG_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Encode a tiled fp16 tensor map (rank 2..4), 128-byte swizzle, zero OOB fill.
// dims/box are innermost-first, strides_bytes[i] is the byte stride of dim i+1.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

int read_watchdog(int out[4], cudaStream_t stream);  // also clears it

}  // namespace lseg
