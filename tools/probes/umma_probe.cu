// Micro-benchmark: issue-to-completion cost of back-to-back tcgen05.mma (kind::f16, M=128, K=16, SS operands)
// as a function of N, B-operand majorness and CTAs per SM. Standalone: nvcc -arch=sm_100a, run on the GPU box.
//   ./umma_probe  -> one line per configuration: cycles per UMMA (max over CTAs), implied TFLOP/s per SM-clock
#include <cstdio>
#include <cstdlib>
#include "../../lang-seg_b200/csrc/common.cuh"
using namespace lseg;

__global__ void __launch_bounds__(128, 2) probe(int n_mma, int N, int b_mn, int a_step, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a_base = smem_u32(smem);           // 16 KB A tile (128 x 64 K-major)
    const uint32_t b_base = smem_u32(smem + 16384);   // up to 32 KB B tile
    const uint32_t idesc = umma_idesc_f16(128, N, 0, b_mn);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int k = i & 3;
      const uint64_t ad = umma_desc_sw128(a_base + k * a_step, 1024, 0);
      const uint64_t bd = b_mn ? umma_desc_sw128(b_base + k * 2048, 1024, 8192) : umma_desc_sw128(b_base + k * 32, 1024, 0);
      umma_f16_ss(tm, ad, bd, idesc, 1);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 256);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 2 * 296 * sizeof(long long));
  long long h[2 * 296];
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 512;
  for (int ctas = 148; ctas <= 296; ctas += 148)
    for (int b_mn = 0; b_mn < 2; ++b_mn)
      for (int N : {16, 32, 64, 128, 256}) {
        if (b_mn && N > 64) continue;
        for (int rep = 0; rep < 2; ++rep) {
          probe<<<ctas, 128, smem>>>(n_mma, N, b_mn, 32, d);
          cudaDeviceSynchronize();
        }
        cudaError_t e = cudaGetLastError();
        cudaMemcpy(h, d, sizeof(long long) * 2 * ctas, cudaMemcpyDeviceToHost);
        long long mi = 0, mt = 0;
        for (int i = 0; i < ctas; ++i) {
          if (h[2 * i] > mi) mi = h[2 * i];
          if (h[2 * i + 1] > mt) mt = h[2 * i + 1];
        }
        const double per = double(mt) / n_mma;
        const double flop_clk_sm = 2.0 * 128 * N * 16 / per * (ctas / 148);
        printf("ctas/SM %d  B %s  N %3d : issue %.1f clk/UMMA, complete %.1f clk/UMMA -> %.0f flop/clk/SM (%.0f%% of 8192)  %s\n",
               ctas / 148, b_mn ? "MN-major" : "K-major ", N, double(mi) / n_mma, per, flop_clk_sm,
               100.0 * flop_clk_sm / 8192, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
