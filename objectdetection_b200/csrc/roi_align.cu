// roi_align.cu — FPN level assignment + bilinear crop_and_resize gather (HBM-bound).
//
// Replaces MaskRCNN.roi_pooling (maskrcnn.py:74-187), tf.image.crop_and_resize
// (TF core/kernels/crop_and_resize_op.cc semantics) and FasterRCNN roi_pool (fastrcnn.py:22-70).
//
// Layout: feature maps NHWC fp32, so one bilinear tap is D*4 contiguous bytes (1 KiB at D=256)
// and the left/right taps of a bin are adjacent; every global access is a 16-byte vector and a
// warp reads/writes 512 contiguous bytes. The reference's per-level where/gather, concat and
// re-sort (maskrcnn.py:127-173) are not reproduced: each ROI writes straight to out[b*N+n].
//
//   1. *_meta_kernel : one thread per ROI -> RoiMeta (level base pointer + sampling grid origin/step)
//   2. crop_rows_kernel : one CTA per (ROI, output row); x-sample table in shared memory;
//      each thread keeps 4 bins x 4 taps of 16-byte loads in flight before blending.
#include "common.cuh"

namespace od {

struct __align__(16) RoiMeta {
  const float* base;  // image (level, batch) base; nullptr -> crop skipped (box_ind out of range)
  int32_t H, W;
  float in_y0, hs, in_x0, ws;  // in_y = in_y0 + y * hs   (TF: y1*(H-1) + y*height_scale)
};

struct LevelTable {
  const float* ptr[OD_MAX_LEVELS];
  int32_t H[OD_MAX_LEVELS];
  int32_t W[OD_MAX_LEVELS];
};

// maskrcnn.py:104-122
__device__ __forceinline__ int32_t roi_level_of(float4 r, int32_t image_h, int32_t image_w, int32_t min_level,
                                                int32_t max_level) {
  const float h = r.z - r.x;
  const float w = r.w - r.y;
  const float image_area = (float)(image_h * image_w);
  const float denom = 224.0f / sqrtf(image_area);
  const float v = sqrtf(h * w) / denom;
  const float lv = f_log(v) / f_log(2.0f);
  const float rr = rintf(lv);  // half to even
  int32_t level = (int32_t)(4u + (uint32_t)f_to_i32_x86(rr));
  level = max(level, min_level);
  level = min(level, max_level);
  return level;
}

__device__ __forceinline__ void fill_grid(RoiMeta& m, float4 box, int32_t ph, int32_t pw) {
  const float y1 = box.x, x1 = box.y, y2 = box.z, x2 = box.w;
  const float Hm1 = (float)(m.H - 1), Wm1 = (float)(m.W - 1);
  if (ph > 1) {
    m.hs = (y2 - y1) * Hm1 / (float)(ph - 1);
    m.in_y0 = y1 * Hm1;
  } else {
    m.hs = 0.0f;
    m.in_y0 = (float)(0.5 * (double)(y1 + y2) * (double)(m.H - 1));
  }
  if (pw > 1) {
    m.ws = (x2 - x1) * Wm1 / (float)(pw - 1);
    m.in_x0 = x1 * Wm1;
  } else {
    m.ws = 0.0f;
    m.in_x0 = (float)(0.5 * (double)(x1 + x2) * (double)(m.W - 1));
  }
}

__global__ void pyramid_meta_kernel(LevelTable lt, const float4* __restrict__ rois, int64_t total, int32_t N, int32_t D,
                                    int32_t image_h, int32_t image_w, int32_t min_level, int32_t num_levels,
                                    int32_t ph, int32_t pw, RoiMeta* __restrict__ meta, int32_t* __restrict__ level_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float4 r = rois[i];
  const int32_t level = roi_level_of(r, image_h, image_w, min_level, min_level + num_levels - 1);
  const int32_t l = level - min_level;
  RoiMeta m;
  m.H = lt.H[l];
  m.W = lt.W[l];
  m.base = lt.ptr[l] + (int64_t)(i / N) * m.H * m.W * D;
  fill_grid(m, r, ph, pw);
  meta[i] = m;
  if (level_out) level_out[i] = level;
}

// Generic tf.image.crop_and_resize meta. frcnn != 0: boxes are [n,5] (batch,x1,y1,x2,y2) pixels divided by
// (image_h, image_w) (fastrcnn.py:55-64).
__global__ void crop_meta_kernel(const float* __restrict__ image, int32_t B, int32_t H, int32_t W, int32_t D,
                                 const float* __restrict__ boxes, const int32_t* __restrict__ box_ind, int32_t n,
                                 int32_t ph, int32_t pw, int32_t frcnn, float image_h, float image_w,
                                 RoiMeta* __restrict__ meta) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 box;
  int32_t b;
  if (frcnn) {
    const float* p = boxes + 5 * (int64_t)i;
    b = (int32_t)p[0];
    box = make_float4(p[2] / image_h, p[1] / image_w, p[4] / image_h, p[3] / image_w);
  } else {
    box = reinterpret_cast<const float4*>(boxes)[i];
    b = box_ind[i];
  }
  RoiMeta m;
  m.H = H;
  m.W = W;
  m.base = (b >= 0 && b < B) ? image + (int64_t)b * H * W * D : nullptr;
  fill_grid(m, box, ph, pw);
  meta[i] = m;
}

struct XSample {
  int32_t left, right;
  float lerp;
  int32_t valid;
};
constexpr int kMaxPoolW = 64;
constexpr int kCropThreads = 128;
constexpr int kCropUnroll = 4;

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) {
  return make_float4(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t);
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(f_max(a.x, b.x), f_max(a.y, b.y), f_max(a.z, b.z), f_max(a.w, b.w));
}

__device__ __forceinline__ void build_xsamples(const RoiMeta& m, int32_t pw, XSample* xs) {
  for (int32_t x = threadIdx.x; x < pw; x += blockDim.x) {
    const float in_x = m.in_x0 + (float)x * m.ws;
    XSample s;
    s.valid = (in_x >= 0.0f) && (in_x <= (float)(m.W - 1));
    const float fl = floorf(in_x);
    s.left = s.valid ? (int32_t)fl : 0;
    s.right = s.valid ? (int32_t)ceilf(in_x) : 0;
    s.lerp = in_x - fl;
    xs[x] = s;
  }
}

// One CTA per (roi, output row y). out row is pw*D4 contiguous float4.
__global__ void __launch_bounds__(kCropThreads)
crop_rows_kernel(const RoiMeta* __restrict__ meta, int32_t ph, int32_t pw, int32_t D4, float extrap,
                 float4* __restrict__ out) {
  __shared__ XSample xs[kMaxPoolW];
  const int64_t item = blockIdx.x;
  const int64_t roi = item / ph;
  const int32_t y = (int32_t)(item - roi * ph);
  const RoiMeta m = meta[roi];
  if (m.base == nullptr) return;
  build_xsamples(m, pw, xs);
  __syncthreads();

  float4* orow = out + (roi * ph + y) * (int64_t)pw * D4;
  const int32_t total = pw * D4;
  const float in_y = m.in_y0 + (float)y * m.hs;
  const float4 ext4 = make_float4(extrap, extrap, extrap, extrap);
  if (!(in_y >= 0.0f) || !(in_y <= (float)(m.H - 1))) {
    for (int32_t e = threadIdx.x; e < total; e += kCropThreads) stg_cs_f4(orow + e, ext4);
    return;
  }
  const float fl = floorf(in_y);
  const int32_t top = (int32_t)fl, bot = (int32_t)ceilf(in_y);
  const float yl = in_y - fl;
  const float4* __restrict__ rtop = reinterpret_cast<const float4*>(m.base) + (int64_t)top * m.W * D4;
  const float4* __restrict__ rbot = reinterpret_cast<const float4*>(m.base) + (int64_t)bot * m.W * D4;

  for (int32_t e0 = threadIdx.x; e0 < total; e0 += kCropThreads * kCropUnroll) {
    float4 tl[kCropUnroll], tr[kCropUnroll], bl[kCropUnroll], br[kCropUnroll];
    float xl[kCropUnroll];
    int32_t ok[kCropUnroll];
#pragma unroll
    for (int u = 0; u < kCropUnroll; ++u) {
      const int32_t e = e0 + u * kCropThreads;
      ok[u] = 0;
      if (e < total) {
        const int32_t x = e / D4, c = e - x * D4;
        const XSample s = xs[x];
        ok[u] = s.valid ? 1 : 2;
        xl[u] = s.lerp;
        if (s.valid) {
          const int32_t lo = s.left * D4 + c, ro = s.right * D4 + c;
          tl[u] = ldg_f4(rtop + lo);
          tr[u] = ldg_f4(rtop + ro);
          bl[u] = ldg_f4(rbot + lo);
          br[u] = ldg_f4(rbot + ro);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kCropUnroll; ++u) {
      const int32_t e = e0 + u * kCropThreads;
      if (ok[u] == 1) {
        const float4 t = lerp4(tl[u], tr[u], xl[u]);
        const float4 b = lerp4(bl[u], br[u], xl[u]);
        stg_cs_f4(orow + e, lerp4(t, b, yl));
      } else if (ok[u] == 2) {
        stg_cs_f4(orow + e, ext4);
      }
    }
  }
}

// FasterRCNN roi_pool: crop 14x14 (ph=pw=14 grid in meta) fused with max_pool 2x2/2 -> 7x7.
// One CTA per (roi, pooled row). NOTE: a skipped crop (base == nullptr) leaves zeros in TF's
// crop output; the pooled output is then zero as well.
__global__ void __launch_bounds__(kCropThreads)
crop_pool2_rows_kernel(const RoiMeta* __restrict__ meta, int32_t ph, int32_t pw, int32_t D4,
                       float4* __restrict__ out) {
  __shared__ XSample xs[kMaxPoolW];
  const int32_t oh = ph / 2, ow = pw / 2;
  const int64_t item = blockIdx.x;
  const int64_t roi = item / oh;
  const int32_t oy = (int32_t)(item - roi * oh);
  const RoiMeta m = meta[roi];
  float4* orow = out + (roi * oh + oy) * (int64_t)ow * D4;
  const int32_t total = ow * D4;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m.base == nullptr) {
    for (int32_t e = threadIdx.x; e < total; e += kCropThreads) stg_cs_f4(orow + e, zero4);
    return;
  }
  build_xsamples(m, pw, xs);
  __syncthreads();
  int32_t top[2], bot[2], yok[2];
  float yl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float in_y = m.in_y0 + (float)(2 * oy + r) * m.hs;
    yok[r] = (in_y >= 0.0f) && (in_y <= (float)(m.H - 1));
    const float fl = floorf(in_y);
    top[r] = yok[r] ? (int32_t)fl : 0;
    bot[r] = yok[r] ? (int32_t)ceilf(in_y) : 0;
    yl[r] = in_y - fl;
  }
  const float4* __restrict__ img = reinterpret_cast<const float4*>(m.base);
  for (int32_t e = threadIdx.x; e < total; e += kCropThreads) {
    const int32_t ox = e / D4, c = e - ox * D4;
    float4 v[4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const XSample s = xs[2 * ox + q];
        float4 val = zero4;  // extrapolation_value = 0 (fastrcnn.py:68 default)
        if (yok[r] && s.valid) {
          const float4* rt = img + (int64_t)top[r] * m.W * D4;
          const float4* rb = img + (int64_t)bot[r] * m.W * D4;
          const float4 t = lerp4(ldg_f4(rt + s.left * D4 + c), ldg_f4(rt + s.right * D4 + c), s.lerp);
          const float4 b = lerp4(ldg_f4(rb + s.left * D4 + c), ldg_f4(rb + s.right * D4 + c), s.lerp);
          val = lerp4(t, b, yl[r]);
        }
        v[r * 2 + q] = val;
      }
    // same visiting order as the oracle: (0,0),(0,1),(1,0),(1,1)
    stg_cs_f4(orow + e, max4(max4(max4(v[0], v[1]), v[2]), v[3]));
  }
}

// ----------------------------------------------------------------------------- host side
static int launch_crop_rows(const RoiMeta* meta, int64_t n_rois, int32_t ph, int32_t pw, int32_t D, float extrap,
                            float* out, cudaStream_t st) {
  if (n_rois == 0) return OD_OK;
  const int64_t items = n_rois * ph;
  if (items > 0x7FFFFFFFll) OD_FAIL(OD_ERR_PARAM, "too many (roi,row) items: %lld", (long long)items);
  crop_rows_kernel<<<(unsigned)items, kCropThreads, 0, st>>>(meta, ph, pw, D / 4, extrap, reinterpret_cast<float4*>(out));
  OD_LAUNCH_CHECK("crop_rows_kernel");
  return OD_OK;
}

}  // namespace od

using namespace od;

extern "C" {

// The RoiMeta table lives at the tail of `pooled`? No: callers do not pass a workspace for this entry
// (the reference API has none), so the table is carved from a small per-call device allocation made with
// cudaMallocAsync on `stream` (stream-ordered, no synchronisation).
int od_pyramid_roi_align_forward(const DLTensor* const* fmaps, int32_t num_levels, int32_t min_level,
                                 const DLTensor* rois, int32_t image_h, int32_t image_w, int32_t pool_h,
                                 int32_t pool_w, DLTensor* pooled, DLTensor* roi_level, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!fmaps) OD_FAIL(OD_ERR_NULL, "fmaps is NULL");
  if (num_levels < 1 || num_levels > OD_MAX_LEVELS) OD_FAIL(OD_ERR_PARAM, "num_levels %d not in [1,%d]", num_levels, OD_MAX_LEVELS);
  if (pool_h < 1 || pool_w < 1 || pool_w > kMaxPoolW) OD_FAIL(OD_ERR_PARAM, "pool shape %dx%d unsupported (w <= %d)", pool_h, pool_w, kMaxPoolW);
  int dev = -1;
  OD_CHECK(check_tensor(rois, "rois", F32, 3, true, &dev));
  if (rois->shape[2] != 4) OD_FAIL(OD_ERR_SHAPE, "rois must be [B,N,4]");
  const int64_t B = rois->shape[0], N = rois->shape[1];
  LevelTable lt;
  int64_t D = -1;
  for (int l = 0; l < num_levels; ++l) {
    OD_CHECK(check_tensor(fmaps[l], "fmaps[l]", F32, 4, true, &dev));
    if (fmaps[l]->shape[0] != B) OD_FAIL(OD_ERR_SHAPE, "fmaps[%d] batch %lld != %lld", l, (long long)fmaps[l]->shape[0], (long long)B);
    if (D < 0) D = fmaps[l]->shape[3];
    if (fmaps[l]->shape[3] != D) OD_FAIL(OD_ERR_SHAPE, "fmaps[%d] depth mismatch", l);
    lt.ptr[l] = dptr<float>(fmaps[l]);
    lt.H[l] = (int32_t)fmaps[l]->shape[1];
    lt.W[l] = (int32_t)fmaps[l]->shape[2];
    if (reinterpret_cast<uintptr_t>(lt.ptr[l]) % 16) OD_FAIL(OD_ERR_LAYOUT, "fmaps[%d] not 16-byte aligned", l);
  }
  if (D % 4) OD_FAIL(OD_ERR_SHAPE, "depth %lld must be a multiple of 4", (long long)D);
  OD_CHECK(check_tensor(pooled, "pooled", F32, -1, true, &dev));
  const int64_t want = B * N * pool_h * pool_w * D;
  const bool shape5 = pooled->ndim == 5 && pooled->shape[0] == 1 && pooled->shape[1] == B * N && pooled->shape[2] == pool_h &&
                      pooled->shape[3] == pool_w && pooled->shape[4] == D;
  const bool shape4 = pooled->ndim == 4 && pooled->shape[0] == B * N && pooled->shape[1] == pool_h &&
                      pooled->shape[2] == pool_w && pooled->shape[3] == D;
  if (!(shape5 || shape4) || numel(pooled) != want) OD_FAIL(OD_ERR_SHAPE, "pooled must be [1,B*N,ph,pw,D] or [B*N,ph,pw,D]");
  if (reinterpret_cast<uintptr_t>(dptr<float>(pooled)) % 16) OD_FAIL(OD_ERR_LAYOUT, "pooled not 16-byte aligned");
  if (roi_level) {
    OD_CHECK(check_tensor(roi_level, "roi_level", I32, 2, true, &dev));
    if (roi_level->shape[0] != B || roi_level->shape[1] != N) OD_FAIL(OD_ERR_SHAPE, "roi_level must be [B,N]");
  }
  const int64_t total = B * N;
  if (total == 0) return OD_OK;
  RoiMeta* meta = nullptr;
  OD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&meta), sizeof(RoiMeta) * (size_t)total, st));
  pyramid_meta_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
      lt, dptr<float4>(rois), total, (int32_t)N, (int32_t)D, image_h, image_w, min_level, num_levels, pool_h, pool_w,
      meta, dptr<int32_t>(roi_level));
  OD_LAUNCH_CHECK("pyramid_meta_kernel");
  int rc = launch_crop_rows(meta, total, pool_h, pool_w, (int32_t)D, 0.0f, dptr<float>(pooled), st);
  cudaFreeAsync(meta, st);
  return rc;
}

int od_crop_and_resize(const DLTensor* image, const DLTensor* boxes, const DLTensor* box_ind, int32_t crop_h,
                       int32_t crop_w, float extrapolation_value, DLTensor* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = -1;
  OD_CHECK(check_tensor(image, "image", F32, 4, true, &dev));
  OD_CHECK(check_tensor(boxes, "boxes", F32, 2, true, &dev));
  OD_CHECK(check_tensor(box_ind, "box_ind", I32, 1, true, &dev));
  OD_CHECK(check_tensor(out, "out", F32, 4, true, &dev));
  if (crop_h < 1 || crop_w < 1 || crop_w > kMaxPoolW) OD_FAIL(OD_ERR_PARAM, "crop size %dx%d unsupported", crop_h, crop_w);
  const int64_t n = boxes->shape[0], D = image->shape[3];
  if (boxes->shape[1] != 4 || box_ind->shape[0] != n) OD_FAIL(OD_ERR_SHAPE, "boxes [n,4] / box_ind [n] mismatch");
  if (out->shape[0] != n || out->shape[1] != crop_h || out->shape[2] != crop_w || out->shape[3] != D)
    OD_FAIL(OD_ERR_SHAPE, "out must be [n,crop_h,crop_w,D]");
  if (D % 4) OD_FAIL(OD_ERR_SHAPE, "depth %lld must be a multiple of 4", (long long)D);
  if (reinterpret_cast<uintptr_t>(dptr<float>(image)) % 16 || reinterpret_cast<uintptr_t>(dptr<float>(out)) % 16 ||
      reinterpret_cast<uintptr_t>(dptr<float>(boxes)) % 16)
    OD_FAIL(OD_ERR_LAYOUT, "image/boxes/out must be 16-byte aligned");
  if (n == 0) return OD_OK;
  RoiMeta* meta = nullptr;
  OD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&meta), sizeof(RoiMeta) * (size_t)n, st));
  crop_meta_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
      dptr<float>(image), (int32_t)image->shape[0], (int32_t)image->shape[1], (int32_t)image->shape[2], (int32_t)D,
      dptr<float>(boxes), dptr<int32_t>(box_ind), (int32_t)n, crop_h, crop_w, 0, 1.f, 1.f, meta);
  OD_LAUNCH_CHECK("crop_meta_kernel");
  int rc = launch_crop_rows(meta, n, crop_h, crop_w, (int32_t)D, extrapolation_value, dptr<float>(out), st);
  cudaFreeAsync(meta, st);
  return rc;
}

int od_roi_pool_forward(const DLTensor* feature_map, const DLTensor* proposals, float image_h, float image_w,
                        DLTensor* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = -1;
  OD_CHECK(check_tensor(feature_map, "feature_map", F32, 4, true, &dev));
  OD_CHECK(check_tensor(proposals, "proposals", F32, 2, true, &dev));
  OD_CHECK(check_tensor(out, "out", F32, 4, true, &dev));
  const int64_t n = proposals->shape[0], D = feature_map->shape[3];
  if (proposals->shape[1] != 5) OD_FAIL(OD_ERR_SHAPE, "proposals must be [n,5]");
  if (out->shape[0] != n || out->shape[1] != 7 || out->shape[2] != 7 || out->shape[3] != D)
    OD_FAIL(OD_ERR_SHAPE, "out must be [n,7,7,D]");
  if (D % 4) OD_FAIL(OD_ERR_SHAPE, "depth %lld must be a multiple of 4", (long long)D);
  if (reinterpret_cast<uintptr_t>(dptr<float>(feature_map)) % 16 || reinterpret_cast<uintptr_t>(dptr<float>(out)) % 16)
    OD_FAIL(OD_ERR_LAYOUT, "feature_map/out must be 16-byte aligned");
  if (n == 0) return OD_OK;
  RoiMeta* meta = nullptr;
  OD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&meta), sizeof(RoiMeta) * (size_t)n, st));
  crop_meta_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
      dptr<float>(feature_map), (int32_t)feature_map->shape[0], (int32_t)feature_map->shape[1],
      (int32_t)feature_map->shape[2], (int32_t)D, dptr<float>(proposals), nullptr, (int32_t)n, 14, 14, 1, image_h,
      image_w, meta);
  OD_LAUNCH_CHECK("crop_meta_kernel");
  crop_pool2_rows_kernel<<<(unsigned)(n * 7), kCropThreads, 0, st>>>(meta, 14, 14, (int32_t)(D / 4), dptr<float4>(out));
  count_launches(1);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(meta, st);
  if (e != cudaSuccess) OD_FAIL(OD_ERR_CUDA, "launch crop_pool2_rows_kernel: %s", cudaGetErrorString(e));
  return OD_OK;
}

}  // extern "C"
