// Micro-benchmark: MUFU (ex2 / rcp) and FMA-pipe issue rates per SM sub-partition, as clk per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(float* out, long long* clk, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("{.reg .b32 t; shl.b32 t, %0, 3; add.s32 %0, t, %0;}" : "+r"(*reinterpret_cast<int*>(&a[i])));
      if (OP == 4) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&clk, 148 * sizeof(long long));
  const int iters = 2000;
  const char* names[5] = {"ex2", "rcp", "ffma", "shl+iadd", "tanh"};
  for (int op = 0; op < 5; ++op)
    for (int threads : {128, 256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) probe<0><<<148, threads>>>(out, clk, iters);
        if (op == 1) probe<1><<<148, threads>>>(out, clk, iters);
        if (op == 2) probe<2><<<148, threads>>>(out, clk, iters);
        if (op == 3) probe<3><<<148, threads>>>(out, clk, iters);
        if (op == 4) probe<4><<<148, threads>>>(out, clk, iters);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      long long m = 0;
      for (int i = 0; i < 148; ++i) m = h[i] > m ? h[i] : m;
      const double warps_per_smsp = threads / 32 / 4.0;
      const double per = double(m) / (iters * 8.0 * warps_per_smsp) / (op == 3 ? 2.0 : 1.0);
      printf("%-9s %4d thr/SM (%.0f warps/SMSP): %.2f clk per warp-instruction per SMSP\n", names[op], threads,
             warps_per_smsp, per);
    }
  return 0;
}
