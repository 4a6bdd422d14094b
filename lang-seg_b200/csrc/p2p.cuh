// lseg_b200 — peer-memory plumbing for the one exchange of the path: gathering the logits of the batch shards on one
// rank (replaces the thread-per-GPU DataParallel gather of additional_utils/models.py:35-53).
//
// One process per GPU. A rank exposes a device buffer to its peers with CUDA IPC (NVLink 5 / NVSwitch peer access is
// enabled by cudaIpcOpenMemHandle); data moves either by the producing kernel storing straight into the peer mapping
// (lseg_forward_lowres with a peer address) or by a copy-engine transfer (lseg_p2p_copy) that overlaps the next step;
// completion is published with system-scope release / acquire flags written and polled by one-thread kernels on the
// CUDA streams involved — no host round trip, no NCCL call on the data path.
#pragma once
#include "common.cuh"

namespace lseg {

// flag[0] = value, after everything this stream did before (kernels and copies complete in stream order; the fence makes
// the flag store cumulative over them at system scope, so a peer that acquires the flag sees the data)
__global__ void p2p_signal_kernel(unsigned long long* flag, unsigned long long value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

// spin until every flags[i * stride] >= value (i < n). One thread; sleeps between polls. A wait longer than timeout_ns
// records tag 77 in the watchdog (lseg_read_watchdog) and returns, so a dead peer shows up as an error, not a hang.
__global__ void p2p_wait_kernel(const unsigned long long* flags, int n, int stride, unsigned long long value,
                                unsigned long long timeout_ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < n; ++i) {
    const unsigned long long* f = flags + static_cast<long long>(i) * stride;
    while (true) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= value) break;
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) {
        if (atomicCAS(&g_watchdog[0], 0, 77) == 0) {
          g_watchdog[1] = i;
          g_watchdog[2] = static_cast<int>(v);
          g_watchdog[3] = static_cast<int>(value);
        }
        return;
      }
      __nanosleep(500);
    }
  }
  __threadfence_system();
}

}  // namespace lseg
