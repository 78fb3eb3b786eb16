// nms.cu — greedy hard NMS with tf.image.non_max_suppression semantics as a bitmask-IoU kernel
// plus a single-CTA keep scan. Call sites replaced: proposals_tf.py:234, detection.py:177.
//
//   mask kernel : 64x64 IoU tiles over the upper triangle; a CTA owns 64 rows x 4 column tiles whose canonical boxes
//                 sit in shared memory; a thread owns row i and builds the 64-bit words "which later boxes j does
//                 i suppress" of two tiles. The
//                 pair test is TF's IoU in TF's operation order, but the IEEE division only runs when the
//                 intersection is positive (inter == 0 gives IoU 0 or NaN, never > thr for thr >= 0), which
//                 is the rare case. Diagonal tiles also emit their transpose with warp ballots (word j =
//                 which earlier boxes of the tile suppress j) for the scan.
//   scan kernel : one CTA per image walks the 64-box chunks in order. Inside a chunk the greedy recurrence
//                 kept_j = cand_j && !(sup_j & kept) is solved by warp-ballot fixed-point iteration (bit j is
//                 final after j+1 rounds; typically 2-4 rounds), then the mask rows of the kept boxes are
//                 OR-ed into the running `removed` bitmap in shared memory. The 64 mask rows of chunk c+1
//                 and its transposed diagonal tile are prefetched (cp.async into a double buffer / registers)
//                 while chunk c is resolved, so the per-chunk critical path never waits on L2. Stops as soon
//                 as max_out boxes are kept.
#include "nms.cuh"
#include "topk.cuh"

namespace od {

constexpr int kScanThreads = 512;

constexpr int kMaskColTiles = 4;   // column tiles (of 64 boxes) per CTA
constexpr int kMaskThreads = 128;  // 64 rows x 2 halves; half h owns column tiles h, h+2 of the CTA's span

template <bool FAST>  // FAST: thr >= 0, division skipped when the intersection is not positive
__global__ void __launch_bounds__(kMaskThreads)
nms_mask_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ num_valid,
                const int32_t* __restrict__ group, int K, int W, int Ws, float thr,
                unsigned long long* __restrict__ mask, uint32_t* __restrict__ diagT) {
  const int rb = blockIdx.y, b = blockIdx.z;
  const int cb0 = rb + blockIdx.x * kMaskColTiles;
  const int n = num_valid ? min(num_valid[b], K) : K;
  if (cb0 * 64 >= n) return;  // rb <= cb0: nothing valid in this span
  __shared__ float4 cbox[kMaskColTiles * 64];   // canonical corners; boxes >= n are zero-area (never hit)
  __shared__ float carea[kMaskColTiles * 64];
  __shared__ int32_t cgrp[kMaskColTiles * 64];
  const int t = threadIdx.x;
  const float4* bx = boxes + (int64_t)b * K;
  for (int q = t; q < kMaskColTiles * 64; q += kMaskThreads) {
    const int j = cb0 * 64 + q;
    const float4 v = (j < n) ? bx[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    const CBox c = canon_box(v);
    cbox[q] = make_float4(c.ymin, c.xmin, c.ymax, c.xmax);
    carea[q] = c.area;
    cgrp[q] = (group && j < n) ? group[(int64_t)b * K + j] : -1;
  }
  __syncthreads();
  const int r = t & 63, half = t >> 6;
  const int i = rb * 64 + r;
  const bool row_ok = i < n;
  const CBox my = canon_box(row_ok ? bx[i] : make_float4(0.f, 0.f, 0.f, 0.f));
  const int32_t g = (group && row_ok) ? group[(int64_t)b * K + i] : 0;
#pragma unroll
  for (int ct = 0; ct < kMaskColTiles / 2; ++ct) {
    const int tile = half + 2 * ct;
    const int cb = cb0 + tile;
    if (cb * 64 >= n) break;   // uniform per half (two warps)
    unsigned long long bits = 0ull;
    const float4* cbp = cbox + tile * 64;
    const float* cap = carea + tile * 64;
    const int32_t* cgp = cgrp + tile * 64;
    if (FAST) {
      // pass 1, branch-free: candidates = a superset of the pairs whose intersection has positive height and width.
      // Everything else has inter == 0 (or NaN) in TF's arithmetic and can never exceed a threshold >= 0.
      // (min(a,b) > max(c,d) needs a > d and b > c; a > c / b > d are the boxes' own validity, which pass 2 settles.)
      // The four differences run on the FMA pipe; a pair is a candidate iff all four are negative, i.e. the AND of
      // their sign bits is set, which one funnel shift appends to the word (bit jj ends up at position jj).
      uint32_t cand_lo = 0u, cand_hi = 0u;
#pragma unroll
      for (int jj = 31; jj >= 0; --jj) {
        const float4 c = cbp[jj];
        const uint32_t sgn = __float_as_uint(c.x - my.ymax) & __float_as_uint(my.ymin - c.z) &
                             __float_as_uint(c.y - my.xmax) & __float_as_uint(my.xmin - c.w);
        cand_lo = __funnelshift_l(sgn, cand_lo, 1);
      }
#pragma unroll
      for (int jj = 31; jj >= 0; --jj) {
        const float4 c = cbp[32 + jj];
        const uint32_t sgn = __float_as_uint(c.x - my.ymax) & __float_as_uint(my.ymin - c.z) &
                             __float_as_uint(c.y - my.xmax) & __float_as_uint(my.xmin - c.w);
        cand_hi = __funnelshift_l(sgn, cand_hi, 1);
      }
      unsigned long long m = ((unsigned long long)cand_hi << 32) | cand_lo;
      if (cb == rb) m &= (r == 63) ? 0ull : ~((2ull << r) - 1ull);   // only j > i
      // pass 2, rare: the exact TF IoU (operation order, IEEE division) of the candidates
      while (m) {
        const int jj = __ffsll((long long)m) - 1;
        m &= m - 1ull;
        const float4 c = cbp[jj];
        const float ih = f_min(my.ymax, c.z) - f_max(my.ymin, c.x);
        const float iw = f_min(my.xmax, c.w) - f_max(my.xmin, c.y);
        const float inter = f_max(ih, 0.0f) * f_max(iw, 0.0f);
        // inter > 0 implies both areas > 0 (inter <= area under monotone rounding), TF's area guard is moot
        const bool hit = (inter > 0.0f) && (inter / (my.area + cap[jj] - inter) > thr) && (!group || g == cgp[jj]);
        bits |= (unsigned long long)hit << jj;
      }
    } else {
#pragma unroll 8
      for (int jj = 0; jj < 64; ++jj) {
        const float4 c = cbp[jj];
        CBox o;
        o.ymin = c.x; o.xmin = c.y; o.ymax = c.z; o.xmax = c.w; o.area = cap[jj];
        const bool hit = (cb * 64 + jj < n) && (tf_iou(my, o) > thr) && (!group || g == cgp[jj]);
        bits |= (unsigned long long)hit << jj;
      }
    }
    if (cb == rb) bits &= (r == 63) ? 0ull : ~((2ull << r) - 1ull);   // only j > i
    if (!row_ok) bits = 0ull;
    if (row_ok) mask[((int64_t)b * K + i) * Ws + cb] = bits;
    if (cb == rb) {
      // transpose of the diagonal tile: word jj, bit r = "box r of this chunk suppresses box jj"
      uint32_t* dt = diagT + ((int64_t)b * W + cb) * 128;
      const int warp = (t >> 5) & 1, lane = t & 31;
#pragma unroll 8
      for (int jj = 0; jj < 64; ++jj) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (bits >> jj) & 1ull);
        if (lane == 0) dt[jj * 2 + warp] = bal;
      }
    }
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// NBUF >= 2: staged scan — Ws is even and NBUF buffers of (64 mask rows + the transposed diagonal tile) fit in shared
// memory; chunk c+NBUF-1 is prefetched with cp.async while chunk c is resolved. NBUF == 0: direct global loads.
template <int NBUF>
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const uint32_t* __restrict__ diagT,
                const int32_t* __restrict__ num_valid, int K, int W, int Ws, int max_out, int32_t* __restrict__ keep_pos,
                int32_t* __restrict__ num_kept, int32_t* __restrict__ keep_flag) {
  constexpr bool STAGED = NBUF >= 2;
  constexpr int DIST = STAGED ? NBUF - 1 : 1;
  extern __shared__ __align__(16) unsigned long long smem_u64[];
  unsigned long long* removed = smem_u64;       // [Ws]
  unsigned long long* stage = smem_u64 + Ws;    // [NBUF][64*Ws + 64] when STAGED
  const size_t buf_words = (size_t)64 * Ws + 64;
  __shared__ unsigned long long kept_word;
  const int b = blockIdx.x;
  const int n = num_valid ? min(num_valid[b], K) : K;
  const int Wn = (n + 63) / 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int w = tid; w < Ws; w += kScanThreads) removed[w] = 0ull;
  if (keep_flag)
    for (int i = tid; i < K; i += kScanThreads) keep_flag[(int64_t)b * K + i] = 0;
  int kept_total = 0;
  const unsigned long long* mrow = mask + (int64_t)b * K * Ws;
  const uint32_t* dbase = diagT + (int64_t)b * W * 128;

  // chunk c -> stage[c % NBUF]: its 64 mask rows, 16-byte units [(c+1)/2, ceil(Wn/2)), and its diagonal tile
  auto prefetch = [&](int c) {
    if (STAGED && c < Wn) {
      unsigned long long* dst = stage + (size_t)(c % (STAGED ? NBUF : 1)) * buf_words;
      const int u0 = (c + 1) >> 1, u1 = (Wn + 1) >> 1;
      const int r = tid >> 3;
      const int row = c * 64 + r;
      if (row < n)
        for (int u = u0 + (tid & 7); u < u1; u += 8)
          cp_async16(dst + (size_t)r * Ws + 2 * u, mrow + (size_t)row * Ws + 2 * u);
      if (tid < 32) cp_async16(dst + (size_t)64 * Ws + 2 * tid, dbase + (size_t)c * 128 + 4 * tid);
    }
    cp_async_commit();
  };
  unsigned long long sup0 = 0ull, sup1 = 0ull;   // !STAGED: warp 0 holds the diagonal tile of the current chunk
  auto load_diag = [&](int c, unsigned long long& s0, unsigned long long& s1) {
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(dbase + (size_t)c * 128) + lane);
    const uint2 d = __ldg(reinterpret_cast<const uint2*>(dbase + (size_t)c * 128) + lane + 32);
    s0 = ((unsigned long long)a.y << 32) | a.x;
    s1 = ((unsigned long long)d.y << 32) | d.x;
  };
  if (STAGED) {
#pragma unroll
    for (int c = 0; c < DIST; ++c) prefetch(c);
  } else if (warp == 0 && Wn > 0) {
    load_diag(0, sup0, sup1);
  }

  for (int c = 0; c < Wn; ++c) {
    if (STAGED) cp_async_wait<DIST - 1>();   // this thread's share of chunk c has landed
    __syncthreads();                         // everyone's share; removed[c] is final; buffer (c-1)%NBUF is free
    if (STAGED) prefetch(c + DIST);
    unsigned long long nsup0 = 0ull, nsup1 = 0ull;
    if (!STAGED && warp == 0 && c + 1 < Wn) load_diag(c + 1, nsup0, nsup1);
    const unsigned long long* buf = stage + (size_t)(c % (STAGED ? NBUF : 1)) * buf_words;
    if (warp == 0) {
      if (STAGED) {
        sup0 = buf[(size_t)64 * Ws + lane];
        sup1 = buf[(size_t)64 * Ws + lane + 32];
      }
      const unsigned long long word = removed[c];
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
      const int allow = max_out - kept_total;
      while (__popcll(kept) > allow) kept &= ~(1ull << (63 - __clzll((long long)kept)));
      if (lane == 0) kept_word = kept;
      if (!STAGED) {
        sup0 = nsup0;
        sup1 = nsup1;
      }
    }
    __syncthreads();
    const unsigned long long kept = kept_word;
    if (kept != 0ull) {
      if (tid < 64 && ((kept >> tid) & 1ull)) {
        const int pos = kept_total + __popcll(kept & ((1ull << tid) - 1ull));
        if (keep_pos) keep_pos[(int64_t)b * max_out + pos] = c * 64 + tid;
        if (keep_flag) keep_flag[(int64_t)b * K + c * 64 + tid] = 1;
      }
      kept_total += __popcll(kept);
      if (kept_total >= max_out) break;
      // OR the rows of the kept boxes into removed[c+1 .. Wn)
      if (STAGED) {
        // 8 row groups x 64 word lanes, rows from shared memory (unconditional loads, select by kept bit)
        const int rg = tid >> 6, wl = tid & 63;
        const unsigned int kbits = (unsigned int)(kept >> (rg * 8)) & 0xFFu;
        if (kbits)
          for (int w = c + 1 + wl; w < Wn; w += 64) {
            const unsigned long long* col = buf + (size_t)(rg * 8) * Ws + w;
            unsigned long long acc = 0ull;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const unsigned long long v = col[(size_t)r * Ws];
              acc |= ((kbits >> r) & 1u) ? v : 0ull;
            }
            if (acc) atomicOr(&removed[w], acc);
          }
      } else {
        // unconditional, fully pipelined global loads: 8 row groups x 64 word lanes
        const int rg = tid >> 6, wl = tid & 63;
        for (int w = c + 1 + wl; w < Wn; w += 64) {
          unsigned long long v[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int row = min(c * 64 + rg * 8 + r, K - 1);
            v[r] = __ldg(&mrow[(size_t)row * Ws + w]);
          }
          unsigned long long acc = 0ull;
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if ((kept >> (rg * 8 + r)) & 1ull) acc |= v[r];
          if (acc) atomicOr(&removed[w], acc);
        }
      }
    }
  }
  if (STAGED) cp_async_wait<0>();
  __syncthreads();
  if (keep_pos)
    for (int j = kept_total + tid; j < max_out; j += kScanThreads) keep_pos[(int64_t)b * max_out + j] = -1;
  if (num_kept && tid == 0) num_kept[b] = kept_total;
}

size_t nms_sorted_workspace_bytes(int64_t B, int64_t K) {
  const int64_t W = (K + 63) / 64, Ws = nms_mask_stride(K);
  Workspace w(nullptr, 0);
  w.take<unsigned long long>((size_t)(B * K * Ws));
  w.take<uint32_t>((size_t)(B * W * 128));
  return w.off + 256;
}

int nms_sorted_launch(const float4* boxes, const int32_t* num_valid, const int32_t* group, int64_t B, int64_t K,
                      float thr, int64_t max_out, int32_t* keep_pos, int32_t* num_kept, int32_t* keep_flag,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B == 0) return OD_OK;
  if (B > 65535) OD_FAIL(OD_ERR_PARAM, "NMS batch %lld > 65535", (long long)B);
  if (K >= (1 << 22)) OD_FAIL(OD_ERR_PARAM, "NMS supports < 4M boxes per image");
  const int W = (int)((K + 63) / 64), Ws = (int)nms_mask_stride(K);
  Workspace w(ws, ws_bytes);
  unsigned long long* mask = w.take<unsigned long long>((size_t)(B * K * Ws));
  uint32_t* diagT = w.take<uint32_t>((size_t)(B * W * 128));
  if (!ws || !w.ok()) OD_FAIL(OD_ERR_WORKSPACE, "NMS workspace %zu < %zu bytes", ws_bytes, w.off);
  if (K > 0) {
    if (W > 65535) OD_FAIL(OD_ERR_PARAM, "NMS tile grid too large");
    const dim3 grid((unsigned)((W + kMaskColTiles - 1) / kMaskColTiles), (unsigned)W, (unsigned)B);
    if (thr >= 0.0f)
      nms_mask_kernel<true><<<grid, kMaskThreads, 0, st>>>(boxes, num_valid, group, (int)K, W, Ws, thr, mask, diagT);
    else
      nms_mask_kernel<false><<<grid, kMaskThreads, 0, st>>>(boxes, num_valid, group, (int)K, W, Ws, thr, mask, diagT);
    OD_LAUNCH_CHECK("nms_mask_kernel");
  }
  return nms_scan_launch(mask, diagT, num_valid, B, K, Ws, max_out, keep_pos, num_kept, keep_flag, st);
}

int nms_scan_launch(const unsigned long long* mask, const uint32_t* diagT, const int32_t* num_valid, int64_t B,
                    int64_t K, int64_t mask_stride, int64_t max_out, int32_t* keep_pos, int32_t* num_kept,
                    int32_t* keep_flag, cudaStream_t st) {
  const int W = (int)((K + 63) / 64), Ws = (int)mask_stride;
  if (Ws < W) OD_FAIL(OD_ERR_PARAM, "NMS mask stride %d < %d words", Ws, W);
  const size_t plain = (size_t)(Ws > 0 ? Ws : 1) * sizeof(unsigned long long);
  const size_t per_buf = ((size_t)64 * Ws + 64) * sizeof(unsigned long long);
  const bool aligned = (Ws % 2 == 0) && (reinterpret_cast<uintptr_t>(mask) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(diagT) % 16 == 0);
  const size_t kSmemBudget = 200 * 1024;
  const int nbuf = !aligned ? 0 : (plain + 3 * per_buf <= kSmemBudget ? 3 : (plain + 2 * per_buf <= kSmemBudget ? 2 : 0));
  const size_t smem = plain + (size_t)nbuf * per_buf;
#define OD_SCAN_LAUNCH(NB)                                                                                          \
  do {                                                                                                              \
    if (smem > 48 * 1024)                                                                                           \
      OD_CUDA(cudaFuncSetAttribute(nms_scan_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    nms_scan_kernel<NB><<<(unsigned)B, kScanThreads, smem, st>>>(mask, diagT, num_valid, (int)K, W, Ws, (int)max_out, \
                                                                 keep_pos, num_kept, keep_flag);                    \
  } while (0)
  if (nbuf == 3) OD_SCAN_LAUNCH(3);
  else if (nbuf == 2) OD_SCAN_LAUNCH(2);
  else OD_SCAN_LAUNCH(0);
#undef OD_SCAN_LAUNCH
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
