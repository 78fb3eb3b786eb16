// nms.cuh — internal interface of the bitmask NMS (nms.cu).
#pragma once
#include "common.cuh"

namespace od {

// Row stride (in 64-bit words) of the suppression bitmask for K boxes: ceil(K/64) rounded up to even, so that every
// mask row is 16-byte aligned (bulk-copy staging in the scan).
inline int64_t nms_mask_stride(int64_t K) { return (((K + 63) / 64) + 1) & ~(int64_t)1; }

// Workspace for nms_sorted_launch (mask + transposed diagonal blocks).
size_t nms_sorted_workspace_bytes(int64_t batch, int64_t K);

// Greedy hard NMS over boxes ALREADY in visiting order (score desc, index asc).
//   boxes      [B,K] float4 (y1,x1,y2,x2)
//   num_valid  [B] or nullptr: only the first num_valid[b] boxes take part
//   group      [B,K] or nullptr: boxes only suppress boxes of the same group (per-class NMS)
//   keep_pos   [B,max_out] or nullptr: kept positions in selection order, -1 padded
//   num_kept   [B] or nullptr
//   keep_flag  [B,K] or nullptr: 1 if position kept (0 otherwise, including invalid positions)
int nms_sorted_launch(const float4* boxes, const int32_t* num_valid, const int32_t* group, int64_t B, int64_t K,
                      float thr, int64_t max_out, int32_t* keep_pos, int32_t* num_kept, int32_t* keep_flag,
                      void* ws, size_t ws_bytes, cudaStream_t st);

// The keep scan alone, over a precomputed suppression bitmask:
//   mask  [B,K,mask_stride] (mask_stride >= W = ceil(K/64), even for the staged path): word (i,w) bit j = box i
//         suppresses box w*64+j (only w > i/64 is read; rows beyond num_valid are never read into a result)
//   diagT [B,W,64]: transposed diagonal tiles (word j of chunk c, bit t: box c*64+t suppresses box c*64+j, t < j)
int nms_scan_launch(const unsigned long long* mask, const unsigned long long* diagT, const int32_t* num_valid, int64_t B,
                    int64_t K, int64_t mask_stride, int64_t max_out, int32_t* keep_pos, int32_t* num_kept,
                    int32_t* keep_flag, cudaStream_t st);

}  // namespace od
