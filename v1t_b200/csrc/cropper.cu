// Image cropper, the step before the core: crop (nearest grid_sample on the shifted crop grid),
// bilinear down-sample to the core's input size and the optional behaviour planes — ONE gather kernel.
//   reference: src/v1t/models/image_cropper.py:104-112 (build_grid), :120-140 (forward:
//              F.grid_sample(mode="nearest", align_corners=True) -> transforms.Resize((36,64), antialias=False)
//              -> behaviours concatenated as constant planes when behavior_mode == 1).
//   ATen semantics restated (PyTorch 2.x, GridSampler.h / UpSample.h): nearest tap = nearbyint(((g + 1)/2)(W - 1))
//   (round-half-even), zeros outside the image; resize source index = max(scale (dst + 0.5) - 0.5, 0) with
//   scale = in/out, taps i0 = floor, i1 = min(i0 + 1, in - 1).
// A thread owns one output pixel of one sample and loops over the channels: its four bilinear taps are four crop
// pixels, each of which is one nearest-sampled input pixel, so the cropped intermediate never exists.
#include "common.cuh"

namespace v1t {
namespace {

__global__ void __launch_bounds__(256) crop_resize_kernel(v1t_crop_shape s, const float* __restrict__ images,
                                                          const float* __restrict__ grid,
                                                          const float* __restrict__ shifts,
                                                          const float* __restrict__ behaviors,
                                                          float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.out_h * s.out_w) return;
  const int oy = i / s.out_w, ox = i % s.out_w;
  int ys[2], xs[2];
  float wy[2], wx[2];
  if (s.out_h == s.crop_h && s.out_w == s.crop_w) {  // no resize
    ys[0] = ys[1] = oy;
    xs[0] = xs[1] = ox;
    wy[0] = wx[0] = 1.f;
    wy[1] = wx[1] = 0.f;
  } else {
    const float sy = (float)s.crop_h / (float)s.out_h, sx = (float)s.crop_w / (float)s.out_w;
    const float fy = fmaxf(sy * ((float)oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)ox + 0.5f) - 0.5f, 0.f);
    ys[0] = min((int)fy, s.crop_h - 1);
    xs[0] = min((int)fx, s.crop_w - 1);
    ys[1] = min(ys[0] + 1, s.crop_h - 1);
    xs[1] = min(xs[0] + 1, s.crop_w - 1);
    wy[1] = fy - (float)ys[0];
    wx[1] = fx - (float)xs[0];
    wy[0] = 1.f - wy[1];
    wx[0] = 1.f - wx[1];
  }
  const float shx = shifts ? shifts[2 * b] : 0.f, shy = shifts ? shifts[2 * b + 1] : 0.f;
  int64_t src[2][2];  // input pixel offset of each tap, -1 = outside (zeros padding)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float* g = grid + ((int64_t)ys[a] * s.crop_w + xs[c]) * 2;
      const float gx = g[0] + shx, gy = g[1] + shy;
      const float ix = (gx + 1.f) * 0.5f * (float)(s.in_w - 1), iy = (gy + 1.f) * 0.5f * (float)(s.in_h - 1);
      const float nx = nearbyintf(ix), ny = nearbyintf(iy);
      const bool ok = nx >= 0.f && nx <= (float)(s.in_w - 1) && ny >= 0.f && ny <= (float)(s.in_h - 1);
      src[a][c] = ok ? (int64_t)ny * s.in_w + (int64_t)nx : -1;
    }
  const int c_out = s.channels + s.behavior_planes;
  const int64_t plane_in = (int64_t)s.in_h * s.in_w, plane_out = (int64_t)s.out_h * s.out_w;
  for (int ch = 0; ch < s.channels; ++ch) {
    const float* img = images + ((int64_t)b * s.channels + ch) * plane_in;
    float v[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) v[a][c] = src[a][c] >= 0 ? __ldg(img + src[a][c]) : 0.f;
    // same association as ATen's upsample_bilinear2d: h0 (w0 v00 + w1 v01) + h1 (w0 v10 + w1 v11)
    out[((int64_t)b * c_out + ch) * plane_out + i] =
        wy[0] * (wx[0] * v[0][0] + wx[1] * v[0][1]) + wy[1] * (wx[0] * v[1][0] + wx[1] * v[1][1]);
  }
  for (int k = 0; k < s.behavior_planes; ++k)
    out[((int64_t)b * c_out + s.channels + k) * plane_out + i] = behaviors[(int64_t)b * s.behavior_planes + k];
}

}  // namespace
}  // namespace v1t

using namespace v1t;

extern "C" int v1t_crop_resize(const v1t_crop_shape* s, const float* images, const float* grid, const float* shifts,
                               const float* behaviors, float* out, void* stream) {
  V1T_CHECK_ARG(s, "crop_resize: null shape");
  V1T_CHECK_ARG(s->batch >= 0 && s->channels > 0 && s->in_h > 0 && s->in_w > 0 && s->crop_h > 0 && s->crop_w > 0 &&
                    s->out_h > 0 && s->out_w > 0 && s->behavior_planes >= 0,
                "crop_resize: bad shape");
  V1T_CHECK_ARG(s->batch <= 65535, "crop_resize: batch %d > 65535", s->batch);
  if (s->batch == 0) return V1T_OK;
  V1T_CHECK_ARG(images && grid && out, "crop_resize: null tensor");
  V1T_CHECK_ARG(s->behavior_planes == 0 || behaviors, "crop_resize: behaviour planes requested without behaviours");
  crop_resize_kernel<<<dim3(cdiv((int64_t)s->out_h * s->out_w, 256), s->batch), 256, 0, (cudaStream_t)stream>>>(
      *s, images, grid, shifts, behaviors, out);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
