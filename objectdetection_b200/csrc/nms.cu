// nms.cu — greedy hard NMS with tf.image.non_max_suppression semantics as a bitmask-IoU kernel
// plus a single-CTA keep scan. Call sites replaced: proposals_tf.py:234, detection.py:177.
//
//   mask kernel : 64x64 IoU tiles over the upper triangle; thread t of a tile owns row i and builds the
//                 64-bit word "which later boxes j does i suppress" (fp32 IoU in TF's operation order,
//                 IEEE division, `> thr`). Diagonal tiles also emit their transpose with warp ballots
//                 (word j = which earlier boxes of the tile suppress j) for the scan.
//   scan kernel : one CTA per image walks the 64-box chunks in order. Inside a chunk the greedy
//                 recurrence kept_j = cand_j && !(sup_j & kept) is solved by warp-ballot fixed-point
//                 iteration (bit j is final after j+1 rounds; typically 2-4 rounds), then the rows of the
//                 kept boxes are OR-ed into the running `removed` bitmap in shared memory. Stops as soon
//                 as max_out boxes are kept.
#include "nms.cuh"
#include "topk.cuh"

namespace od {

constexpr int kScanThreads = 512;

__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ num_valid,
                const int32_t* __restrict__ group, int K, int W, float thr,
                unsigned long long* __restrict__ mask, uint32_t* __restrict__ diagT) {
  const int cb = blockIdx.x, rb = blockIdx.y, b = blockIdx.z;
  if (cb < rb) return;
  const int n = num_valid ? min(num_valid[b], K) : K;
  if (rb * 64 >= n || cb * 64 >= n) return;
  __shared__ CBox cbox[64];
  __shared__ int32_t cgrp[64];
  const int t = threadIdx.x;
  const float4* bx = boxes + (int64_t)b * K;
  {
    const int j = cb * 64 + t;
    const float4 v = (j < n) ? bx[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    cbox[t] = canon_box(v);
    cgrp[t] = (group && j < n) ? group[(int64_t)b * K + j] : 0;
  }
  __syncthreads();
  const int i = rb * 64 + t;
  unsigned long long bits = 0ull;
  if (i < n) {
    const CBox my = canon_box(bx[i]);
    const int32_t g = group ? group[(int64_t)b * K + i] : 0;
#pragma unroll 8
    for (int jj = 0; jj < 64; ++jj) {
      const int j = cb * 64 + jj;
      const bool hit = (j > i) && (j < n) && (tf_iou(my, cbox[jj]) > thr) && (g == cgrp[jj]);
      bits |= (unsigned long long)hit << jj;
    }
    mask[((int64_t)b * K + i) * W + cb] = bits;
  }
  if (cb == rb) {
    // transpose of the diagonal tile: word jj, bit t = "box t of this chunk suppresses box jj"
    uint32_t* dt = diagT + ((int64_t)b * W + cb) * 128;
    const int warp = t >> 5, lane = t & 31;
#pragma unroll 8
    for (int jj = 0; jj < 64; ++jj) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (bits >> jj) & 1ull);
      if (lane == 0) dt[jj * 2 + warp] = bal;
    }
  }
}

__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const uint32_t* __restrict__ diagT,
                const int32_t* __restrict__ num_valid, int K, int W, int max_out, int32_t* __restrict__ keep_pos,
                int32_t* __restrict__ num_kept, int32_t* __restrict__ keep_flag) {
  extern __shared__ unsigned long long removed[];  // [W]
  __shared__ unsigned long long kept_word;
  const int b = blockIdx.x;
  const int n = num_valid ? min(num_valid[b], K) : K;
  const int Wn = (n + 63) / 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int w = tid; w < W; w += kScanThreads) removed[w] = 0ull;
  if (keep_flag)
    for (int i = tid; i < K; i += kScanThreads) keep_flag[(int64_t)b * K + i] = 0;
  int kept_total = 0;
  const unsigned long long* mrow = mask + (int64_t)b * K * W;
  for (int c = 0; c < Wn && kept_total < max_out; ++c) {
    __syncthreads();
    if (warp == 0) {
      const unsigned long long word = removed[c];
      const uint32_t* dt = diagT + ((int64_t)b * W + c) * 128;
      const unsigned long long sup0 = ((unsigned long long)dt[lane * 2 + 1] << 32) | dt[lane * 2];
      const unsigned long long sup1 = ((unsigned long long)dt[(lane + 32) * 2 + 1] << 32) | dt[(lane + 32) * 2];
      const bool cand0 = (c * 64 + lane < n) && !((word >> lane) & 1ull);
      const bool cand1 = (c * 64 + lane + 32 < n) && !((word >> (lane + 32)) & 1ull);
      unsigned long long kept = (unsigned long long)__ballot_sync(0xffffffffu, cand0) |
                                ((unsigned long long)__ballot_sync(0xffffffffu, cand1) << 32);
      for (int it = 0; it < 64; ++it) {
        const bool k0 = cand0 && !(sup0 & kept);
        const bool k1 = cand1 && !(sup1 & kept);
        const unsigned long long nk = (unsigned long long)__ballot_sync(0xffffffffu, k0) |
                                      ((unsigned long long)__ballot_sync(0xffffffffu, k1) << 32);
        if (nk == kept) break;
        kept = nk;
      }
      // respect max_out: drop the highest set bits beyond the allowance
      int allow = max_out - kept_total;
      while (__popcll(kept) > allow) kept &= ~(1ull << (63 - __clzll((long long)kept)));
      if (lane == 0) kept_word = kept;
    }
    __syncthreads();
    const unsigned long long kept = kept_word;
    if (kept == 0ull) continue;
    if (tid < 64 && ((kept >> tid) & 1ull)) {
      const int pos = kept_total + __popcll(kept & ((1ull << tid) - 1ull));
      if (keep_pos) keep_pos[(int64_t)b * max_out + pos] = c * 64 + tid;
      if (keep_flag) keep_flag[(int64_t)b * K + c * 64 + tid] = 1;
    }
    kept_total += __popcll(kept);
    if (kept_total >= max_out) break;
    // OR the rows of the kept boxes into removed[c+1 .. Wn): 128 word lanes x 4 row groups
    const int rg = tid >> 7, wl = tid & 127;
    for (int w = c + 1 + wl; w < Wn; w += 128) {
      unsigned long long acc = 0ull;
      int rank = 0;
#pragma unroll 8
      for (int bit = 0; bit < 64; ++bit) {
        if ((kept >> bit) & 1ull) {
          if ((rank & 3) == rg) acc |= __ldg(&mrow[((int64_t)c * 64 + bit) * W + w]);
          ++rank;
        }
      }
      if (acc) atomicOr(&removed[w], acc);
    }
  }
  __syncthreads();
  if (keep_pos)
    for (int j = kept_total + tid; j < max_out; j += kScanThreads) keep_pos[(int64_t)b * max_out + j] = -1;
  if (num_kept && tid == 0) num_kept[b] = kept_total;
}

size_t nms_sorted_workspace_bytes(int64_t B, int64_t K) {
  const int64_t W = (K + 63) / 64;
  Workspace w(nullptr, 0);
  w.take<unsigned long long>((size_t)(B * K * W));
  w.take<uint32_t>((size_t)(B * W * 128));
  return w.off + 256;
}

int nms_sorted_launch(const float4* boxes, const int32_t* num_valid, const int32_t* group, int64_t B, int64_t K,
                      float thr, int64_t max_out, int32_t* keep_pos, int32_t* num_kept, int32_t* keep_flag,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B == 0) return OD_OK;
  if (B > 65535) OD_FAIL(OD_ERR_PARAM, "NMS batch %lld > 65535", (long long)B);
  if (K >= (1 << 22)) OD_FAIL(OD_ERR_PARAM, "NMS supports < 4M boxes per image");
  const int W = (int)((K + 63) / 64);
  Workspace w(ws, ws_bytes);
  unsigned long long* mask = w.take<unsigned long long>((size_t)(B * K * W));
  uint32_t* diagT = w.take<uint32_t>((size_t)(B * W * 128));
  if (!ws || !w.ok()) OD_FAIL(OD_ERR_WORKSPACE, "NMS workspace %zu < %zu bytes", ws_bytes, w.off);
  if (K > 0) {
    if (W > 65535) OD_FAIL(OD_ERR_PARAM, "NMS tile grid too large");
    const dim3 grid((unsigned)W, (unsigned)W, (unsigned)B);
    nms_mask_kernel<<<grid, 64, 0, st>>>(boxes, num_valid, group, (int)K, W, thr, mask, diagT);
    OD_LAUNCH_CHECK("nms_mask_kernel");
  }
  return nms_scan_launch(mask, diagT, num_valid, B, K, max_out, keep_pos, num_kept, keep_flag, st);
}

int nms_scan_launch(const unsigned long long* mask, const uint32_t* diagT, const int32_t* num_valid, int64_t B,
                    int64_t K, int64_t max_out, int32_t* keep_pos, int32_t* num_kept, int32_t* keep_flag,
                    cudaStream_t st) {
  const int W = (int)((K + 63) / 64);
  const size_t smem = (size_t)(W > 0 ? W : 1) * sizeof(unsigned long long);
  if (smem > 48 * 1024)
    OD_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_scan_kernel<<<(unsigned)B, kScanThreads, smem, st>>>(mask, diagT, num_valid, (int)K, W, (int)max_out, keep_pos,
                                                           num_kept, keep_flag);
  OD_LAUNCH_CHECK("nms_scan_kernel");
  return OD_OK;
}

// ---- unsorted front-end (tf.image.non_max_suppression on arbitrary score order)
__global__ void nms_build_keys_kernel(const float* __restrict__ scores, const int32_t* __restrict__ num_valid, int K,
                                      int64_t n_pow2, unsigned long long* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= n_pow2) return;
  const int n = num_valid ? min(num_valid[b], K) : K;
  keys[(int64_t)b * n_pow2 + i] = (i < n) ? composite_key(scores[(int64_t)b * K + i], (uint32_t)i) : 0ull;
}
__global__ void nms_gather_sorted_kernel(const float4* __restrict__ boxes, const unsigned long long* __restrict__ keys,
                                         const int32_t* __restrict__ num_valid, int K, int64_t n_pow2,
                                         float4* __restrict__ sorted) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const int n = num_valid ? min(num_valid[b], K) : K;
  if (i >= K) return;
  sorted[(int64_t)b * K + i] =
      (i < n) ? boxes[(int64_t)b * K + composite_index(keys[(int64_t)b * n_pow2 + i])] : make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void nms_map_back_kernel(const unsigned long long* __restrict__ keys, int64_t n_pow2, int max_out,
                                    int32_t* __restrict__ keep) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= max_out) return;
  const int32_t pos = keep[(int64_t)b * max_out + j];
  if (pos >= 0) keep[(int64_t)b * max_out + j] = (int32_t)composite_index(keys[(int64_t)b * n_pow2 + pos]);
}

}  // namespace od

using namespace od;

extern "C" {

size_t od_nms_workspace_bytes(int64_t batch, int64_t num_boxes) {
  Workspace w(nullptr, 0);
  w.take<unsigned long long>((size_t)(batch * next_pow2(num_boxes > 0 ? num_boxes : 1)));
  w.take<float4>((size_t)(batch * num_boxes));
  return w.off + 256 + nms_sorted_workspace_bytes(batch, num_boxes);
}

int od_nms(const DLTensor* boxes, const DLTensor* scores, const DLTensor* num_valid, float iou_threshold,
           int64_t max_out, DLTensor* keep_idx, DLTensor* num_kept, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = -1;
  OD_CHECK(check_tensor(boxes, "boxes", F32, 3, true, &dev));
  OD_CHECK(check_tensor(scores, "scores", F32, 2, true, &dev));
  OD_CHECK(check_tensor(keep_idx, "keep_idx", I32, 2, true, &dev));
  const int64_t B = boxes->shape[0], K = boxes->shape[1];
  if (boxes->shape[2] != 4 || scores->shape[0] != B || scores->shape[1] != K) OD_FAIL(OD_ERR_SHAPE, "boxes [B,K,4] / scores [B,K] mismatch");
  if (keep_idx->shape[0] != B || keep_idx->shape[1] != max_out) OD_FAIL(OD_ERR_SHAPE, "keep_idx must be [B,max_out]");
  if (num_valid) {
    OD_CHECK(check_tensor(num_valid, "num_valid", I32, 1, true, &dev));
    if (num_valid->shape[0] != B) OD_FAIL(OD_ERR_SHAPE, "num_valid must be [B]");
  }
  if (num_kept) {
    OD_CHECK(check_tensor(num_kept, "num_kept", I32, 1, true, &dev));
    if (num_kept->shape[0] != B) OD_FAIL(OD_ERR_SHAPE, "num_kept must be [B]");
  }
  if (reinterpret_cast<uintptr_t>(dptr<float>(boxes)) % 16) OD_FAIL(OD_ERR_LAYOUT, "boxes not 16-byte aligned");
  if (B == 0 || max_out == 0) return OD_OK;
  if (!ws) OD_FAIL(OD_ERR_WORKSPACE, "NMS workspace is NULL");
  const int64_t n_pow2 = next_pow2(K > 0 ? K : 1);
  Workspace w(ws, ws_bytes);
  unsigned long long* keys = w.take<unsigned long long>((size_t)(B * n_pow2));
  float4* sorted = w.take<float4>((size_t)(B * K));
  w.off = align_up(w.off, 256);
  if (!w.ok() || ws_bytes < w.off + nms_sorted_workspace_bytes(B, K) - 256)
    OD_FAIL(OD_ERR_WORKSPACE, "NMS workspace %zu bytes too small", ws_bytes);
  const int32_t* nv = dptr<int32_t>(num_valid);
  {
    const dim3 g((unsigned)((n_pow2 + 255) / 256), (unsigned)B);
    nms_build_keys_kernel<<<g, 256, 0, st>>>(dptr<float>(scores), nv, (int)K, n_pow2, keys);
    OD_LAUNCH_CHECK("nms_build_keys_kernel");
  }
  OD_CHECK(sort_u64_desc_launch(keys, B, n_pow2, st));
  if (K > 0) {
    const dim3 g((unsigned)((K + 255) / 256), (unsigned)B);
    nms_gather_sorted_kernel<<<g, 256, 0, st>>>(dptr<float4>(boxes), keys, nv, (int)K, n_pow2, sorted);
    OD_LAUNCH_CHECK("nms_gather_sorted_kernel");
  }
  OD_CHECK(nms_sorted_launch(sorted, nv, nullptr, B, K, iou_threshold, max_out, dptr<int32_t>(keep_idx),
                             dptr<int32_t>(num_kept), nullptr, w.base + w.off, ws_bytes - w.off, st));
  {
    const dim3 g((unsigned)((max_out + 255) / 256), (unsigned)B);
    nms_map_back_kernel<<<g, 256, 0, st>>>(keys, n_pow2, (int)max_out, dptr<int32_t>(keep_idx));
    OD_LAUNCH_CHECK("nms_map_back_kernel");
  }
  return OD_OK;
}

}  // extern "C"
