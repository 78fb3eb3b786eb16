// anchors.cuh — anchor regeneration from a flat index, fp64, bit-exact with numpy
// (utils.py:230-353: generate_anchors_for_feature_map, gen_anchors, norm_boxes).
#pragma once
#include "common.cuh"

namespace od {

struct DevAnchorSpec {
  int32_t num_levels, num_ratios;
  double scales[OD_MAX_LEVELS];
  double sqrt_ratios[OD_MAX_RATIOS];  // IEEE sqrt done on the host == np.sqrt
  int32_t nx[OD_MAX_LEVELS];          // anchor columns per level: ceil(fmap_w / anchor_stride)
  int32_t step[OD_MAX_LEVELS];        // anchor_stride * fmap_stride (pixels between anchor centres)
  int64_t offset[OD_MAX_LEVELS + 1];  // first flat index of each level
  double norm_scale[4];               // (h-1, w-1, h-1, w-1)
};

inline int make_dev_anchor_spec(const od_anchor_spec* s, DevAnchorSpec* d) {
  if (!s) OD_FAIL(OD_ERR_NULL, "anchor spec is NULL");
  if (s->num_levels < 1 || s->num_levels > OD_MAX_LEVELS || s->num_ratios < 1 || s->num_ratios > OD_MAX_RATIOS ||
      s->anchor_stride < 1)
    OD_FAIL(OD_ERR_PARAM, "anchor spec out of range (levels %d, ratios %d, stride %d)", s->num_levels, s->num_ratios,
            s->anchor_stride);
  d->num_levels = s->num_levels;
  d->num_ratios = s->num_ratios;
  int64_t off = 0;
  for (int l = 0; l < s->num_levels; ++l) {
    const int64_t ny = (s->fmap_h[l] + s->anchor_stride - 1) / s->anchor_stride;
    const int64_t nx = (s->fmap_w[l] + s->anchor_stride - 1) / s->anchor_stride;
    d->scales[l] = s->scales[l];
    d->nx[l] = (int32_t)nx;
    d->step[l] = s->anchor_stride * s->fmap_stride[l];
    d->offset[l] = off;
    off += ny * nx * s->num_ratios;
  }
  for (int l = s->num_levels; l <= OD_MAX_LEVELS; ++l) d->offset[l] = off;
  for (int r = 0; r < s->num_ratios; ++r) d->sqrt_ratios[r] = __builtin_sqrt(s->ratios[r]);
  d->norm_scale[0] = d->norm_scale[2] = (double)(s->image_h - 1);
  d->norm_scale[1] = d->norm_scale[3] = (double)(s->image_w - 1);
  return OD_OK;
}

// Pixel-coordinate anchor (y1,x1,y2,x2) of flat index i in fp64.
__device__ __forceinline__ void anchor_pixel(const DevAnchorSpec& s, int64_t i, double out[4]) {
  int l = 0;
  while (l + 1 < s.num_levels && i >= s.offset[l + 1]) ++l;
  const int64_t rem = i - s.offset[l];
  const int32_t r = (int32_t)(rem % s.num_ratios);
  const int64_t cell = rem / s.num_ratios;
  const int64_t x = cell % s.nx[l], y = cell / s.nx[l];
  const double h = s.scales[l] / s.sqrt_ratios[r];
  const double w = s.scales[l] * s.sqrt_ratios[r];
  const double cy = (double)(y * s.step[l]);
  const double cx = (double)(x * s.step[l]);
  out[0] = cy - 0.5 * h;
  out[1] = cx - 0.5 * w;
  out[2] = cy + 0.5 * h;
  out[3] = cx + 0.5 * w;
}
// norm_boxes: (box - [0,0,1,1]) / (h-1,w-1,h-1,w-1), cast to fp32 (utils.py:193-196).
__device__ __forceinline__ float4 anchor_normalized(const DevAnchorSpec& s, int64_t i) {
  double p[4];
  anchor_pixel(s, i, p);
  return make_float4((float)((p[0] - 0.0) / s.norm_scale[0]), (float)((p[1] - 0.0) / s.norm_scale[1]),
                     (float)((p[2] - 1.0) / s.norm_scale[2]), (float)((p[3] - 1.0) / s.norm_scale[3]));
}

}  // namespace od
