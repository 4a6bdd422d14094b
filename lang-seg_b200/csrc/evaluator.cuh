// lseg_b200 — device side of the multi-scale / flip / sliding-window evaluator (mirror of
// LSeg_MultiEvalModule.forward, additional_utils/models.py:55-140, helpers :142-170).
//
// The reference builds every network input with a chain of torch ops per window — F.interpolate(align_corners=True) of
// the whole image, F.pad to the crop size with the normalised zero, a slice, a second F.pad, torch.flip — and replays
// the inverse chain on every network output (flip back, add, slice-accumulate into a canvas, count, divide, slice,
// F.interpolate back, add). Here that glue is three gather kernels:
//   eval_make_crops      image [3,h,w] -> all crop_size x crop_size network inputs of one scale (windows and their
//                        horizontal flips) in one launch: each output pixel samples the ORIGINAL image bilinearly at the
//                        scale's coordinates or takes the pad value — the resized / padded intermediates never exist;
//   eval_canvas          network outputs of the scale's windows -> overlap-averaged canvas [K,height,width]: per pixel the
//                        (plain + flipped-back) outputs of the covering windows are summed in the reference's window order
//                        and divided by their count;
//   eval_resize_add      scores[K,h,w] += bilinear(canvas -> (h,w), align_corners=True).
// Arithmetic follows the reference's order of operations (same adds in the same order, one division), so the batched
// result equals the sequential algorithm on the same network bit for bit; against torch's own interpolate kernels the
// bilinear weights agree to fp32 rounding (FMA contraction of the library build is not observable from here).
#pragma once
#include "common.cuh"

namespace lseg {

struct EvalWindow {
  int height, width;  // the scale's resized image
  int h0, w0;         // window origin in the (padded) resized image
  int flip;           // 1: this network input is the horizontally flipped crop
  int out_index;      // index of this input in the crop batch / of its output in the logits batch
};

// src index + weights of torch's bilinear align_corners=True (ATen area_pixel_compute_source_index / UpSample.cuh)
__device__ __forceinline__ void bilinear_ac_coord(int dst, int in, int out, int& i0, int& i1, float& l0, float& l1) {
  const float scale = (out > 1) ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float r = scale * static_cast<float>(dst);
  i0 = static_cast<int>(r);
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = r - static_cast<float>(i0);
  l0 = 1.f - l1;
}
__device__ __forceinline__ float bilinear_ac_blend(float v00, float v01, float v10, float v11, float h0l, float h1l, float w0l,
                                                   float w1l) {
  return __fadd_rn(__fmul_rn(h0l, __fadd_rn(__fmul_rn(w0l, v00), __fmul_rn(w1l, v01))),
                   __fmul_rn(h1l, __fadd_rn(__fmul_rn(w0l, v10), __fmul_rn(w1l, v11))));
}

// grid (ceil(crop/32), ceil(crop/8), n_inputs), block (32, 8): one thread = one pixel, all 3 channels
__global__ void eval_make_crops_kernel(const float* __restrict__ img, float* __restrict__ crops,
                                       const EvalWindow* __restrict__ wins, int h, int w, int crop, float pad0, float pad1,
                                       float pad2) {
  const EvalWindow win = wins[blockIdx.z];
  const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
  if (i >= crop || j >= crop) return;
  const int y = win.h0 + i;
  const int x = win.w0 + (win.flip ? crop - 1 - j : j);
  float* dst = crops + (static_cast<long long>(win.out_index) * 3) * crop * crop + static_cast<long long>(i) * crop + j;
  const long long cstride = static_cast<long long>(crop) * crop;
  if (y >= win.height || x >= win.width) {
    dst[0] = pad0;
    dst[cstride] = pad1;
    dst[2 * cstride] = pad2;
    return;
  }
  int y0, y1, x0, x1;
  float hy0, hy1, wx0, wx1;
  bilinear_ac_coord(y, h, win.height, y0, y1, hy0, hy1);
  bilinear_ac_coord(x, w, win.width, x0, x1, wx0, wx1);
  const long long plane = static_cast<long long>(h) * w;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* p = img + c * plane;
    dst[c * cstride] = bilinear_ac_blend(p[static_cast<long long>(y0) * w + x0], p[static_cast<long long>(y0) * w + x1],
                                         p[static_cast<long long>(y1) * w + x0], p[static_cast<long long>(y1) * w + x1], hy0,
                                         hy1, wx0, wx1);
  }
}

// canvas[k, y, x] (y < height, x < width) = sum over the scale's windows covering (y, x), in list order, of
//   (out[plain][k, y-h0, x-w0] + out[flipped][k, y-h0, crop-1-(x-w0)])   divided by the number of such windows.
// wins: the n_win PLAIN entries of the scale (out_index of the plain output; the flipped one is out_index + 1 when
// flip != 0 in the call). grid (ceil(width/32), ceil(height/8), K), block (32, 8)
__global__ void eval_canvas_kernel(const float* __restrict__ outs, float* __restrict__ canvas,
                                   const EvalWindow* __restrict__ wins, int n_win, int K, int crop, int height, int width,
                                   int flip, int whole) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, k = blockIdx.z;
  if (y >= height || x >= width) return;
  const long long cc = static_cast<long long>(crop) * crop;
  float acc = 0.f;
  int cnt = 0;
  for (int wi = 0; wi < n_win; ++wi) {
    const EvalWindow win = wins[wi];
    const int i = y - win.h0, j = x - win.w0;
    if (i < 0 || i >= crop || j < 0 || j >= crop) continue;
    const float* o = outs + (static_cast<long long>(win.out_index) * K + k) * cc + static_cast<long long>(i) * crop;
    float v = o[j];
    if (flip) v = __fadd_rn(v, o[static_cast<long long>(K) * cc + (crop - 1 - j)]);  // output += flip(flipped output)
    acc = cnt ? __fadd_rn(acc, v) : v;  // outputs[...] += output on a zero canvas
    ++cnt;
  }
  // sliding-window scales divide by count_norm; the single padded window of a small scale is used as is
  canvas[(static_cast<long long>(k) * height + y) * width + x] = whole ? acc : __fdiv_rn(acc, static_cast<float>(cnt));
}

// scores[k, i, j] += bilinear(canvas[k] (height x width) -> (h, w), align_corners=True).  grid (ceil(w/32), ceil(h/8), K)
__global__ void eval_resize_add_kernel(const float* __restrict__ canvas, float* __restrict__ scores, int height, int width,
                                       int h, int w) {
  const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y, k = blockIdx.z;
  if (i >= h || j >= w) return;
  int y0, y1, x0, x1;
  float hy0, hy1, wx0, wx1;
  bilinear_ac_coord(i, height, h, y0, y1, hy0, hy1);
  bilinear_ac_coord(j, width, w, x0, x1, wx0, wx1);
  const float* p = canvas + static_cast<long long>(k) * height * width;
  const float v = bilinear_ac_blend(p[static_cast<long long>(y0) * width + x0], p[static_cast<long long>(y0) * width + x1],
                                    p[static_cast<long long>(y1) * width + x0], p[static_cast<long long>(y1) * width + x1], hy0,
                                    hy1, wx0, wx1);
  float* s = scores + (static_cast<long long>(k) * h + i) * w + j;
  *s = __fadd_rn(*s, v);
}

// ------------------------------------------------------------------------------------------
// on-GPU preprocessing: ToTensor + Normalize(mean, std) + Resize (bilinear, torchvision's
// tensor resize = F.interpolate(align_corners=False), no antialias) of an 8-bit HWC image, as lseg_app.py:328-334 and
// modules/lseg_module.py:42-53 chain them, into the fp32 NCHW network input, optionally padded (bottom / right) with a
// constant (pad_image's -mean/std, additional_utils/models.py:145-156). grid (ceil(Wp/32), ceil(Hp/8)), block (32, 8)
// ------------------------------------------------------------------------------------------
__global__ void preprocess_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, int h, int w, int Ho, int Wo,
                                  int Hp, int Wp, float m0, float m1, float m2, float s0, float s1, float s2, float pad0,
                                  float pad1, float pad2) {
  const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
  if (i >= Hp || j >= Wp) return;
  const long long plane = static_cast<long long>(Hp) * Wp;
  float* dst = out + static_cast<long long>(i) * Wp + j;
  if (i >= Ho || j >= Wo) {
    dst[0] = pad0;
    dst[plane] = pad1;
    dst[2 * plane] = pad2;
    return;
  }
  // align_corners=False source coordinates (ATen: scale = in / out; src = max(scale * (dst + 0.5) - 0.5, 0))
  const float sy = static_cast<float>(h) / Ho, sx = static_cast<float>(w) / Wo;
  const float fy = fmaxf(sy * (i + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * (j + 0.5f) - 0.5f, 0.f);
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = y0 + ((y0 < h - 1) ? 1 : 0), x1 = x0 + ((x0 < w - 1) ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    auto px = [&](int yy, int xx) {  // ToTensor (u8 / 255) then Normalize ((v - mean) / std), per source pixel
      const float v = static_cast<float>(img[(static_cast<long long>(yy) * w + xx) * 3 + c]) / 255.f;
      return (v - mean[c]) / stdv[c];
    };
    dst[c * plane] = bilinear_ac_blend(px(y0, x0), px(y0, x1), px(y1, x0), px(y1, x1), hy, ly, hx, lx);
  }
}

}  // namespace lseg
