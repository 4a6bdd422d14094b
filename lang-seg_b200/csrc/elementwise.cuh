// lseg_b200 — HBM-bound helper kernels (vectorised / warp-shuffle) for the LSeg forward path.
// Each kernel cites the reference line it restates; all are plain grid-stride or warp-per-row
// kernels sized in multiples of the SM count by their launchers.
#pragma once
#include "common.cuh"

namespace lseg {

// ------------------------------------------------------------------------------------------
// patchify: x fp32 NCHW [B,3,H,W] -> A fp16 [B*gh*gw, 768], column = c*256 + py*16 + px, which is
// the flattening of the patch-embed Conv2d(3,1024,k16,s16) weight (modules/models/lseg_vit.py:179),
// so the conv becomes one GEMM with the weight used as stored.
// ------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const float* __restrict__ x, __half* __restrict__ a, int B, int H, int W) {
  const int gh = H / 16, gw = W / 16;
  const long long total = static_cast<long long>(B) * gh * gw * 3 * 16 * 4;  // one thread = 4 pixels of a patch row
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px4 = static_cast<int>(i & 3);
    long long t = i >> 2;
    const int py = static_cast<int>(t & 15);
    t >>= 4;
    const int c = static_cast<int>(t % 3);
    t /= 3;
    const int gx = static_cast<int>(t % gw);
    t /= gw;
    const int gy = static_cast<int>(t % gh);
    const int b = static_cast<int>(t / gh);
    const float4 v = *reinterpret_cast<const float4*>(
        x + ((static_cast<long long>(b) * 3 + c) * H + gy * 16 + py) * W + gx * 16 + px4 * 4);
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    __half* dst = a + (static_cast<long long>(b) * gh * gw + gy * gw + gx) * 768 + c * 256 + py * 16 + px4 * 4;
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(dst) = o;
  }
}

// ------------------------------------------------------------------------------------------
// pos-embed resize, bilinear align_corners=False (modules/models/lseg_vit.py:149-163); weight-only,
// computed once per token grid. pos [1+g0*g0, D] -> out [1+gh*gw, D]; row 0 (cls) copied.
// ------------------------------------------------------------------------------------------
__global__ void pos_resize_kernel(const float* __restrict__ pos, float* __restrict__ out, int g0, int gh, int gw,
                                  int D) {
  const int row = blockIdx.x;  // 0 .. gh*gw
  if (row == 0) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) out[d] = pos[d];
    return;
  }
  const int oy = (row - 1) / gw, ox = (row - 1) % gw;
  const float sy = static_cast<float>(g0) / gh, sx = static_cast<float>(g0) / gw;
  float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.f);
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = y0 + ((y0 < g0 - 1) ? 1 : 0), x1 = x0 + ((x0 < g0 - 1) ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float* p00 = pos + (1 + y0 * g0 + x0) * static_cast<long long>(D);
  const float* p01 = pos + (1 + y0 * g0 + x1) * static_cast<long long>(D);
  const float* p10 = pos + (1 + y1 * g0 + x0) * static_cast<long long>(D);
  const float* p11 = pos + (1 + y1 * g0 + x1) * static_cast<long long>(D);
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    out[static_cast<long long>(row) * D + d] = hy * (hx * p00[d] + lx * p01[d]) + ly * (hx * p10[d] + lx * p11[d]);
}

// ------------------------------------------------------------------------------------------
// token assembly (modules/models/lseg_vit.py:188-193): x[b,0] = cls + pos[0];
// x[b,1+t] = patch[b*T+t] (bias already added by the GEMM) + pos[1+t].   fp32 [B, 1+T, D]
// ------------------------------------------------------------------------------------------
__global__ void assemble_tokens_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                       const float* __restrict__ pos, float* __restrict__ x, int B, int T, int D) {
  const int d4 = D / 4;
  const long long total = static_cast<long long>(B) * (T + 1) * d4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d4);
    const long long row = i / d4;
    const int t = static_cast<int>(row % (T + 1));
    const int b = static_cast<int>(row / (T + 1));
    const float4 pe = reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * D)[c];
    float4 v;
    if (t == 0)
      v = reinterpret_cast<const float4*>(cls)[c];
    else
      v = reinterpret_cast<const float4*>(patch + (static_cast<long long>(b) * T + (t - 1)) * D)[c];
    v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    reinterpret_cast<float4*>(x + row * D)[c] = v;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row, fp32 statistics (two-pass in registers), fp16 out.
// timm: eps 1e-6 on the fp32 residual stream; CLIP: eps 1e-5 on the fp16 stream computed in fp32
// (SURVEY.md Appendix A.1/A.2). C must be a multiple of 128 and <= 1024.
// ------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void layernorm_kernel(const TIn* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __half* __restrict__ y, long long M, int C,
                                 float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nv = C / 128;  // float4 groups per lane
  float v[32];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      if constexpr (sizeof(TIn) == 4) {
        const float4 q = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + row * C + col);
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
      } else {
        const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(x) + row * C + col);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = b.x; v[4 * i + 3] = b.y;
      }
      s += v[4 * i] + v[4 * i + 1] + v[4 * i + 2] + v[4 * i + 3];
    }
  }
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < nv) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = v[4 * i + j] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / C + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + col);
      const float4 bb = *reinterpret_cast<const float4*>(beta + col);
      __half2 h0 = __floats2half2_rn((v[4 * i] - mean) * rstd * g.x + bb.x, (v[4 * i + 1] - mean) * rstd * g.y + bb.y);
      __half2 h1 =
          __floats2half2_rn((v[4 * i + 2] - mean) * rstd * g.z + bb.z, (v[4 * i + 3] - mean) * rstd * g.w + bb.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&h0);
      o.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(y + row * C + col) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------
// readout split (modules/models/lseg_vit.py:79-90): tap fp32 [B, 1+T, D] ->
//   tok fp16 [B*T, D] (patch tokens) and cls fp16 [B, D].
// ------------------------------------------------------------------------------------------
__global__ void readout_split_kernel(const float* __restrict__ tap, __half* __restrict__ tok,
                                     __half* __restrict__ cls, int B, int T, int D) {
  const int d4 = D / 4;
  const long long total = static_cast<long long>(B) * (T + 1) * d4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d4);
    const long long row = i / d4;
    const int t = static_cast<int>(row % (T + 1));
    const int b = static_cast<int>(row / (T + 1));
    const float4 v = reinterpret_cast<const float4*>(tap + row * D)[c];
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    __half* dst = (t == 0) ? cls + static_cast<long long>(b) * D : tok + (static_cast<long long>(b) * T + (t - 1)) * D;
    reinterpret_cast<uint2*>(dst)[c] = o;
  }
}

// ------------------------------------------------------------------------------------------
// im2col for the one strided conv (act_postprocess4[4]: 3x3 s2 p1, modules/models/lseg_vit.py:516-522):
// NHWC fp16 [B,H,W,C] -> [B*Ho*Wo, 9*C], tap-major columns, zero halo.
// ------------------------------------------------------------------------------------------
__global__ void im2col_3x3_s2_kernel(const __half* __restrict__ x, __half* __restrict__ a, int B, int H, int W,
                                     int C) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int c8 = C / 8;
  const long long total = static_cast<long long>(B) * Ho * Wo * 9 * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c8);
    long long t = i / c8;
    const int tap = static_cast<int>(t % 9);
    t /= 9;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    const int iy = oy * 2 - 1 + tap / 3, ix = ox * 2 - 1 + tap % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * H + iy) * W + ix) * C)[c];
    reinterpret_cast<uint4*>(a + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * 9 * C + static_cast<long long>(tap) * C)[c] = v;
  }
}

// ------------------------------------------------------------------------------------------
// bilinear x2, align_corners=True, NHWC fp16 -> NHWC fp16 (fusion blocks, lseg_blocks.py:352-354).
// src = dst * (in-1)/(out-1), computed like ATen (float scale, float product).
// ------------------------------------------------------------------------------------------
__global__ void upsample2x_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W,
                                       int C) {
  const int Ho = 2 * H, Wo = 2 * W, c8 = C / 8;
  const float sh = (Ho > 1) ? static_cast<float>(H - 1) / (Ho - 1) : 0.f;
  const float sw = (Wo > 1) ? static_cast<float>(W - 1) / (Wo - 1) : 0.f;
  const long long total = static_cast<long long>(B) * Ho * Wo * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c8);
    long long t = i / c8;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    const float fy = sh * oy, fx = sw * ox;
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const __half* base = x + static_cast<long long>(b) * H * W * C;
    const uint4 q00 = reinterpret_cast<const uint4*>(base + (static_cast<long long>(y0) * W + x0) * C)[c];
    const uint4 q01 = reinterpret_cast<const uint4*>(base + (static_cast<long long>(y0) * W + x1) * C)[c];
    const uint4 q10 = reinterpret_cast<const uint4*>(base + (static_cast<long long>(y1) * W + x0) * C)[c];
    const uint4 q11 = reinterpret_cast<const uint4*>(base + (static_cast<long long>(y1) * W + x1) * C)[c];
    const __half2* a00 = reinterpret_cast<const __half2*>(&q00);
    const __half2* a01 = reinterpret_cast<const __half2*>(&q01);
    const __half2* a10 = reinterpret_cast<const __half2*>(&q10);
    const __half2* a11 = reinterpret_cast<const __half2*>(&q11);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f00 = __half22float2(a00[k]), f01 = __half22float2(a01[k]);
      const float2 f10 = __half22float2(a10[k]), f11 = __half22float2(a11[k]);
      oh[k] = __floats2half2_rn(hy * (hx * f00.x + lx * f01.x) + ly * (hx * f10.x + lx * f11.x),
                                hy * (hx * f00.y + lx * f01.y) + ly * (hx * f10.y + lx * f11.y));
    }
    reinterpret_cast<uint4*>(y + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * C)[c] = o;
  }
}

// ------------------------------------------------------------------------------------------
// pixel-feature normalisation (modules/models/lseg_net.py:191,194): row / ||row||_2 in fp32, cast to
// fp16, then multiply by logit_scale in fp16 (the reference's `logit_scale * image_features.half()`
// rounds the product to fp16 before the matmul). One warp per row of C=512.
// ------------------------------------------------------------------------------------------
__global__ void l2norm_scale_kernel(const float* __restrict__ x, __half* __restrict__ y, long long M, int C,
                                    float logit_scale) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nv = C / 128;
  float v[16];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < nv) {
      const float4 q = *reinterpret_cast<const float4*>(x + row * C + (i * 32 + lane) * 4);
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
      ss += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    }
  }
  const float nrm = sqrtf(warp_sum(ss));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < nv) {
      __half h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half u = __float2half_rn(v[4 * i + j] / nrm);
        h[j] = __float2half_rn(logit_scale * __half2float(u));
      }
      *reinterpret_cast<uint2*>(y + row * C + (i * 32 + lane) * 4) = *reinterpret_cast<uint2*>(h);
    }
  }
}

// text-feature normalisation in fp16 (modules/models/lseg_net.py:192): norm rounded to fp16, then
// an fp16 division. One warp per row.
__global__ void l2norm_f16_kernel(const __half* __restrict__ x, __half* __restrict__ y, int M, int C) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float f = __half2float(x[static_cast<long long>(row) * C + c]);
    ss += f * f;
  }
  const float nrm = __half2float(__float2half_rn(sqrtf(warp_sum(ss))));
  for (int c = lane; c < C; c += 32)
    y[static_cast<long long>(row) * C + c] = __float2half_rn(__half2float(x[static_cast<long long>(row) * C + c]) / nrm);
}

// ------------------------------------------------------------------------------------------
// output head (modules/models/lseg_net.py:196,203): fp16 logits [B,K,h,w] (values of the fp16
// matmul) -> .float() -> bilinear x2 align_corners=True -> fp32 NCHW [B,K,2h,2w].
// HBM-write-bound: each thread produces 4 consecutive outputs (one float4 store).
// ------------------------------------------------------------------------------------------
__global__ void upsample2x_nchw_kernel(const __half* __restrict__ x, float* __restrict__ y, long long planes, int H,
                                       int W) {
  const int Ho = 2 * H, Wo = 2 * W, w4 = Wo / 4;
  const float sh = (Ho > 1) ? static_cast<float>(H - 1) / (Ho - 1) : 0.f;
  const float sw = (Wo > 1) ? static_cast<float>(W - 1) / (Wo - 1) : 0.f;
  const long long total = planes * Ho * w4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xq = static_cast<int>(i % w4);
    long long t = i / w4;
    const int oy = static_cast<int>(t % Ho);
    const long long pl = t / Ho;
    const float fy = sh * oy;
    const int y0 = static_cast<int>(fy);
    const int y1 = min(y0 + 1, H - 1);
    const float ly = fy - y0, hy = 1.f - ly;
    const __half* r0 = x + (pl * H + y0) * W;
    const __half* r1 = x + (pl * H + y1) * W;
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ox = xq * 4 + k;
      const float fx = sw * ox;
      const int x0 = static_cast<int>(fx);
      const int x1 = min(x0 + 1, W - 1);
      const float lx = fx - x0, hx = 1.f - lx;
      o[k] = hy * (hx * __half2float(r0[x0]) + lx * __half2float(r0[x1])) +
             ly * (hx * __half2float(r1[x0]) + lx * __half2float(r1[x1]));
    }
    __stcs(reinterpret_cast<float4*>(y + (pl * Ho + oy) * Wo + xq * 4), make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ------------------------------------------------------------------------------------------
// CLIP text tower glue
// ------------------------------------------------------------------------------------------
// x = token_embedding(text).half() + positional_embedding.half()   (fp16 add)
__global__ void text_embed_kernel(const long long* __restrict__ tokens, const float* __restrict__ tok_emb,
                                  const float* __restrict__ pos_emb, __half* __restrict__ x, int K, int L, int Wd) {
  const long long total = static_cast<long long>(K) * L * Wd;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % Wd);
    const long long row = i / Wd;
    const int t = static_cast<int>(row % L);
    const long long id = tokens[row];
    x[i] = __hadd(__float2half_rn(tok_emb[id * Wd + d]), __float2half_rn(pos_emb[static_cast<long long>(t) * Wd + d]));
  }
}
// rows at the EOT position: text.argmax(-1) (first maximal id), gathered after ln_final.
__global__ void text_eot_gather_kernel(const long long* __restrict__ tokens, const __half* __restrict__ x,
                                       __half* __restrict__ out, int K, int L, int Wd) {
  const int k = blockIdx.x;
  if (k >= K) return;
  __shared__ int s_pos;
  if (threadIdx.x == 0) {
    long long best = tokens[static_cast<long long>(k) * L];
    int bp = 0;
    for (int t = 1; t < L; ++t) {
      const long long v = tokens[static_cast<long long>(k) * L + t];
      if (v > best) { best = v; bp = t; }
    }
    s_pos = bp;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < Wd; d += blockDim.x)
    out[static_cast<long long>(k) * Wd + d] = x[(static_cast<long long>(k) * L + s_pos) * Wd + d];
}

}  // namespace lseg
