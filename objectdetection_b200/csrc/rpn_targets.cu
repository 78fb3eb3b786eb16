// rpn_targets.cu — RPN target builder: PreprareTrainData.build_rpn_targets (data_processor.py:173-294) with
// utils.intersection_over_union (utils.py:32-40), batched over images. SURVEY.md §8(f) rank 1: the step in front of
// the detection-head path in training (anchors x GT IoU, 261,888 x <=100 per image).
//
// The reference is numpy float64 throughout, so every kernel here computes in fp64 and the labels are bit-exact:
//   rpn_anchor_best_kernel : one thread per anchor, GT boxes in shared memory -> max IoU / first argmax per anchor
//   rpn_gt_best_kernel     : one CTA per (GT, image) over all anchors -> first argmax per GT ("best anchor" rule :232)
//   rpn_label_kernel       : -1 (< 0.3) / +1 (>= 0.7 or best anchor of a GT) / 0                       (:224-238)
//   rpn_subsample_kernel   : one CTA per image: balance to max_rpn_targets (:242-262) with the two
//                            np.random.choice draws replaced by explicit permutations (the list element at position
//                            q is dropped for the first `extra` entries q of the permutation with q < len(list)),
//                            then the fp64 box deltas of the surviving positives (:265-291).
#include "common.cuh"

namespace od {

constexpr int kRpnThreads = 256;
constexpr int kRpnSubThreads = 1024;

// utils.intersection_over_union: box = GT, boxes = anchors; union = (box_area + boxes_area) - intersection.
__device__ __forceinline__ double rpn_iou(const double* g, double g_area, double a0, double a1, double a2, double a3, double a_area) {
  const double y1 = fmax(g[0], a0), y2 = fmin(g[2], a2);
  const double x1 = fmax(g[1], a1), x2 = fmin(g[3], a3);
  const double inter = fmax(x2 - x1, 0.0) * fmax(y2 - y1, 0.0);
  // disjoint boxes: 0 / union is +-0 (the comparisons below treat both alike); skips the fp64 divide for ~99 % of pairs
  if (inter == 0.0) return 0.0;
  const double uni = g_area + a_area - inter;
  return inter / uni;
}

__global__ void __launch_bounds__(kRpnThreads)
rpn_anchor_best_kernel(const double* __restrict__ anchors, int A, const double* __restrict__ gt, const int32_t* __restrict__ gt_count,
                       int G, double* __restrict__ iou_max, int32_t* __restrict__ iou_arg) {
  extern __shared__ double s_gt[];   // [G][5]: y1,x1,y2,x2,area
  const int b = blockIdx.y;
  const int ng = min(gt_count[b], G);
  for (int j = threadIdx.x; j < ng; j += kRpnThreads) {
    const double* g = gt + ((int64_t)b * G + j) * 4;
    s_gt[j * 5 + 0] = g[0]; s_gt[j * 5 + 1] = g[1]; s_gt[j * 5 + 2] = g[2]; s_gt[j * 5 + 3] = g[3];
    s_gt[j * 5 + 4] = (g[2] - g[0]) * (g[3] - g[1]);
  }
  __syncthreads();
  const int a = blockIdx.x * kRpnThreads + threadIdx.x;
  if (a >= A) return;
  const double a0 = anchors[4 * (int64_t)a], a1 = anchors[4 * (int64_t)a + 1], a2 = anchors[4 * (int64_t)a + 2],
               a3 = anchors[4 * (int64_t)a + 3];
  const double a_area = (a2 - a0) * (a3 - a1);
  double best = 0.0;   // no GT: every anchor is background
  int arg = 0;
  for (int j = 0; j < ng; ++j) {
    const double v = rpn_iou(&s_gt[j * 5], s_gt[j * 5 + 4], a0, a1, a2, a3, a_area);
    if (j == 0 || v > best) {   // np.argmax: first maximum
      best = v;
      arg = j;
    }
  }
  iou_max[(int64_t)b * A + a] = best;
  iou_arg[(int64_t)b * A + a] = arg;
}

__global__ void __launch_bounds__(kRpnThreads)
rpn_gt_best_kernel(const double* __restrict__ anchors, int A, const double* __restrict__ gt, const int32_t* __restrict__ gt_count,
                   int G, int32_t* __restrict__ gt_best) {
  __shared__ double s_v[kRpnThreads];
  __shared__ int32_t s_i[kRpnThreads];
  const int j = blockIdx.x, b = blockIdx.y;
  if (j >= min(gt_count[b], G)) {
    if (threadIdx.x == 0) gt_best[(int64_t)b * G + j] = -1;
    return;
  }
  const double* gp = gt + ((int64_t)b * G + j) * 4;
  const double g[4] = {gp[0], gp[1], gp[2], gp[3]};
  const double g_area = (g[2] - g[0]) * (g[3] - g[1]);
  double best = -1.0;
  int arg = 0x7fffffff;
  for (int a = threadIdx.x; a < A; a += kRpnThreads) {
    const double a0 = anchors[4 * (int64_t)a], a1 = anchors[4 * (int64_t)a + 1], a2 = anchors[4 * (int64_t)a + 2],
                 a3 = anchors[4 * (int64_t)a + 3];
    const double v = rpn_iou(g, g_area, a0, a1, a2, a3, (a2 - a0) * (a3 - a1));
    if (v > best) {   // ascending a per thread: the first maximum wins
      best = v;
      arg = a;
    }
  }
  s_v[threadIdx.x] = best;
  s_i[threadIdx.x] = arg;
  __syncthreads();
  for (int o = kRpnThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double ov = s_v[threadIdx.x + o];
      const int oi = s_i[threadIdx.x + o];
      if (ov > s_v[threadIdx.x] || (ov == s_v[threadIdx.x] && oi < s_i[threadIdx.x])) {
        s_v[threadIdx.x] = ov;
        s_i[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) gt_best[(int64_t)b * G + j] = (s_i[0] == 0x7fffffff) ? 0 : s_i[0];
}

__global__ void rpn_label_kernel(const double* __restrict__ iou_max, int64_t total, int32_t* __restrict__ cls) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double m = iou_max[i];
  cls[i] = (m >= 0.7) ? 1 : ((m < 0.3) ? -1 : 0);
}
__global__ void rpn_label_best_kernel(const int32_t* __restrict__ gt_best, int G, int A, int32_t* __restrict__ cls) {
  const int j = threadIdx.x + blockIdx.x * blockDim.x, b = blockIdx.y;
  if (j >= G) return;
  const int a = gt_best[(int64_t)b * G + j];
  if (a >= 0) cls[(int64_t)b * A + a] = 1;
}

constexpr int kCompactItems = 8;
// Stable compaction of {i in [0,n) : pred(i)} by a 1024-thread CTA, 8 consecutive elements per thread and round;
// emit(i, rank) for every selected i; returns the count. scratch: 33 ints of shared memory.
template <typename Pred, typename Emit>
__device__ int cta_compact(int n, Pred pred, Emit emit, int* scratch) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int base = 0;
  for (int t0 = 0; t0 < n; t0 += kCompactItems * kRpnSubThreads) {
    const int i0 = t0 + kCompactItems * tid;
    bool p[kCompactItems];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < kCompactItems; ++u) {
      p[u] = (i0 + u < n) && pred(i0 + u);
      cnt += p[u] ? 1 : 0;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = scratch[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      scratch[lane] = w;   // inclusive warp totals
    }
    __syncthreads();
    int r = base + (warp ? scratch[warp - 1] : 0) + incl - cnt;
#pragma unroll
    for (int u = 0; u < kCompactItems; ++u)
      if (p[u]) emit(i0 + u, r++);
    base += scratch[31];
    __syncthreads();
  }
  return base;
}

__global__ void __launch_bounds__(kRpnSubThreads)
rpn_subsample_kernel(const double* __restrict__ anchors, int A, const double* __restrict__ gt, int G,
                     const int32_t* __restrict__ iou_arg, const int32_t* __restrict__ perm_pos,
                     const int32_t* __restrict__ perm_neg, int max_targets, double sd0, double sd1, double sd2, double sd3,
                     int32_t* __restrict__ cls, int32_t* __restrict__ list /*[B,A] scratch*/, double* __restrict__ target_bbox,
                     double* __restrict__ positive_anchors, int32_t* __restrict__ counts /*[B,4]*/) {
  __shared__ int scratch[33];
  const int b = blockIdx.x, tid = threadIdx.x;
  int32_t* c = cls + (int64_t)b * A;
  int32_t* lst = list + (int64_t)b * A;
  const int32_t* pp = perm_pos + (int64_t)b * A;
  const int32_t* pn = perm_neg + (int64_t)b * A;
  // positives: idx = where(cls == 1); drop `extra` of them (:242-247)
  const int n_pos0 = cta_compact(A, [&](int i) { return c[i] == 1; }, [&](int i, int r) { lst[r] = i; }, scratch);
  __syncthreads();
  const int extra_pos = n_pos0 - max_targets / 2;
  if (extra_pos > 0) {
    cta_compact(A, [&](int t) { const int q = pp[t]; return q >= 0 && q < n_pos0; },
                 [&](int t, int r) { if (r < extra_pos) c[lst[pp[t]]] = 0; }, scratch);
    __syncthreads();
  }
  const int n_pos = extra_pos > 0 ? n_pos0 - extra_pos : n_pos0;
  // negatives: idx = where(cls == -1); keep max_targets - n_pos of them (:249-253)
  const int n_neg0 = cta_compact(A, [&](int i) { return c[i] == -1; }, [&](int i, int r) { lst[r] = i; }, scratch);
  __syncthreads();
  const int extra_neg = n_neg0 - (max_targets - n_pos);
  if (extra_neg > 0) {
    cta_compact(A, [&](int t) { const int q = pn[t]; return q >= 0 && q < n_neg0; },
                 [&](int t, int r) { if (r < extra_neg) c[lst[pn[t]]] = 0; }, scratch);
    __syncthreads();
  }
  // regression targets of the surviving positives, ascending anchor index (:256-291)
  double* tb = target_bbox + (int64_t)b * max_targets * 4;
  double* pa = positive_anchors + (int64_t)b * max_targets * 4;
  for (int i = tid; i < max_targets * 4; i += kRpnSubThreads) {
    tb[i] = 0.0;
    pa[i] = 0.0;
  }
  __syncthreads();
  cta_compact(A, [&](int i) { return c[i] == 1; },
               [&](int i, int r) {
                 if (r >= max_targets) return;
                 const double* an = anchors + 4 * (int64_t)i;
                 const double* g = gt + ((int64_t)b * G + iou_arg[(int64_t)b * A + i]) * 4;
                 const double ah = an[2] - an[0], aw = an[3] - an[1];
                 const double acy = an[0] + 0.5 * ah, acx = an[1] + 0.5 * aw;
                 const double gh = g[2] - g[0], gw = g[3] - g[1];
                 const double gcy = g[0] + 0.5 * gh, gcx = g[1] + 0.5 * gw;
                 tb[r * 4 + 0] = ((gcy - acy) / ah) / sd0;
                 tb[r * 4 + 1] = ((gcx - acx) / aw) / sd1;
                 tb[r * 4 + 2] = log(gh / ah) / sd2;
                 tb[r * 4 + 3] = log(gw / aw) / sd3;
                 pa[r * 4 + 0] = an[0]; pa[r * 4 + 1] = an[1]; pa[r * 4 + 2] = an[2]; pa[r * 4 + 3] = an[3];
               },
               scratch);
  if (tid == 0) {
    int32_t* o = counts + (int64_t)b * 4;
    o[0] = n_pos0;
    o[1] = n_neg0;
    o[2] = n_pos;
    o[3] = extra_neg > 0 ? n_neg0 - extra_neg : n_neg0;
  }
}

}  // namespace od

using namespace od;

extern "C" {

size_t od_rpn_target_workspace_bytes(int64_t batch, int64_t num_anchors, int64_t num_gt) {
  Workspace w(nullptr, 0);
  w.take<double>((size_t)(batch * num_anchors));    // iou_max
  w.take<int32_t>((size_t)(batch * num_anchors));   // iou_arg
  w.take<int32_t>((size_t)(batch * num_anchors));   // index list scratch
  w.take<int32_t>((size_t)(batch * (num_gt > 0 ? num_gt : 1)));   // best anchor per GT
  return w.off + 256;
}

int od_rpn_target_forward(const DLTensor* anchors, const DLTensor* gt_boxes, const DLTensor* gt_count,
                          const DLTensor* perm_pos, const DLTensor* perm_neg, const od_rpn_target_params* params,
                          DLTensor* rpn_target_class, DLTensor* rpn_target_bbox, DLTensor* positive_anchors,
                          DLTensor* counts, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!params) OD_FAIL(OD_ERR_NULL, "params is NULL");
  int dev = -1;
  OD_CHECK(check_tensor(anchors, "anchors", F64, 2, true, &dev));
  OD_CHECK(check_tensor(gt_boxes, "gt_boxes", F64, 3, true, &dev));
  OD_CHECK(check_tensor(gt_count, "gt_count", I32, 1, true, &dev));
  OD_CHECK(check_tensor(perm_pos, "perm_pos", I32, 2, true, &dev));
  OD_CHECK(check_tensor(perm_neg, "perm_neg", I32, 2, true, &dev));
  OD_CHECK(check_tensor(rpn_target_class, "rpn_target_class", I32, 2, true, &dev));
  OD_CHECK(check_tensor(rpn_target_bbox, "rpn_target_bbox", F64, 3, true, &dev));
  OD_CHECK(check_tensor(positive_anchors, "positive_anchors", F64, 3, true, &dev));
  OD_CHECK(check_tensor(counts, "counts", I32, 2, true, &dev));
  const int64_t A = anchors->shape[0], B = gt_boxes->shape[0], G = gt_boxes->shape[1], T = params->max_rpn_targets;
  if (anchors->shape[1] != 4 || gt_boxes->shape[2] != 4) OD_FAIL(OD_ERR_SHAPE, "anchors [A,4] / gt_boxes [B,G,4] expected");
  if (gt_count->shape[0] != B) OD_FAIL(OD_ERR_SHAPE, "gt_count must be [B]");
  if (perm_pos->shape[0] != B || perm_pos->shape[1] != A || perm_neg->shape[0] != B || perm_neg->shape[1] != A)
    OD_FAIL(OD_ERR_SHAPE, "perm_pos / perm_neg must be [B,A]");
  if (rpn_target_class->shape[0] != B || rpn_target_class->shape[1] != A) OD_FAIL(OD_ERR_SHAPE, "rpn_target_class must be [B,A]");
  if (T < 0 || rpn_target_bbox->shape[0] != B || rpn_target_bbox->shape[1] != T || rpn_target_bbox->shape[2] != 4 ||
      positive_anchors->shape[0] != B || positive_anchors->shape[1] != T || positive_anchors->shape[2] != 4)
    OD_FAIL(OD_ERR_SHAPE, "rpn_target_bbox / positive_anchors must be [B,max_rpn_targets,4]");
  if (counts->shape[0] != B || counts->shape[1] != 4) OD_FAIL(OD_ERR_SHAPE, "counts must be [B,4]");
  if (A >= (1ll << 30) || G > 4096 || B > 65535) OD_FAIL(OD_ERR_PARAM, "supports A < 2^30, G <= 4096, B <= 65535");
  if (B == 0 || A == 0) return OD_OK;
  if (!ws) OD_FAIL(OD_ERR_WORKSPACE, "workspace is NULL");
  Workspace w(ws, ws_bytes);
  double* iou_max = w.take<double>((size_t)(B * A));
  int32_t* iou_arg = w.take<int32_t>((size_t)(B * A));
  int32_t* list = w.take<int32_t>((size_t)(B * A));
  int32_t* gt_best = w.take<int32_t>((size_t)(B * (G > 0 ? G : 1)));
  if (!w.ok()) OD_FAIL(OD_ERR_WORKSPACE, "workspace %zu < %zu bytes", ws_bytes, w.off);
  const double* an = dptr<double>(anchors);
  const double* gt = dptr<double>(gt_boxes);
  const int32_t* gc = dptr<int32_t>(gt_count);
  int32_t* cls = dptr<int32_t>(rpn_target_class);
  {
    const dim3 grid((unsigned)((A + kRpnThreads - 1) / kRpnThreads), (unsigned)B);
    const size_t smem = (size_t)(G > 0 ? G : 1) * 5 * sizeof(double);
    if (smem > 48 * 1024)
      OD_CUDA(cudaFuncSetAttribute(rpn_anchor_best_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpn_anchor_best_kernel<<<grid, kRpnThreads, smem, st>>>(an, (int)A, gt, gc, (int)G, iou_max, iou_arg);
    OD_LAUNCH_CHECK("rpn_anchor_best_kernel");
  }
  if (G > 0) {
    const dim3 grid((unsigned)G, (unsigned)B);
    rpn_gt_best_kernel<<<grid, kRpnThreads, 0, st>>>(an, (int)A, gt, gc, (int)G, gt_best);
    OD_LAUNCH_CHECK("rpn_gt_best_kernel");
  }
  {
    const int64_t total = B * A;
    rpn_label_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(iou_max, total, cls);
    OD_LAUNCH_CHECK("rpn_label_kernel");
  }
  if (G > 0) {
    const dim3 grid((unsigned)((G + 127) / 128), (unsigned)B);
    rpn_label_best_kernel<<<grid, 128, 0, st>>>(gt_best, (int)G, (int)A, cls);
    OD_LAUNCH_CHECK("rpn_label_best_kernel");
  }
  rpn_subsample_kernel<<<(unsigned)B, kRpnSubThreads, 0, st>>>(
      an, (int)A, gt, (int)G, iou_arg, dptr<int32_t>(perm_pos), dptr<int32_t>(perm_neg), (int)T, params->bbox_stddev[0],
      params->bbox_stddev[1], params->bbox_stddev[2], params->bbox_stddev[3], cls, list, dptr<double>(rpn_target_bbox),
      dptr<double>(positive_anchors), dptr<int32_t>(counts));
  OD_LAUNCH_CHECK("rpn_subsample_kernel");
  return OD_OK;
}

}  // extern "C"
