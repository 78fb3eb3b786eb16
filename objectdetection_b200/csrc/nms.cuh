// nms.cuh — internal interface of the bitmask NMS (nms.cu).
#pragma once
#include "common.cuh"

namespace od {

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

// The keep scan alone, over a precomputed suppression bitmask in the tile-major layout of nms.cu: with
// W = ceil(K/64), tile (rb, cb), cb > rb, is 64 consecutive words at ((b*W + rb)*W + cb)*64 (word r = which boxes of
// chunk cb does box rb*64+r suppress); the diagonal tile (rb, rb) holds the transpose (word j = which earlier boxes of
// the chunk suppress box rb*64+j). Tiles with cb < rb are never read. Rows/columns beyond num_valid must be zero or
// unwritten-but-masked: the scan only ORs rows of kept boxes.
int nms_scan_launch(const unsigned long long* mask, const int32_t* num_valid, int64_t B, int64_t K, int64_t max_out,
                    int32_t* keep_pos, int32_t* num_kept, int32_t* keep_flag, cudaStream_t st);

}  // namespace od
