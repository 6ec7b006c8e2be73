// nms.cu — multi-class rotated BEV NMS: the last step of FSF.simple_test.
//
// Reference: FrustumClusterHead._get_bboxes_single (projects/mmdet3d_plugin/models/dense_heads/frustum_cluster_head.py:
// 595-698): scores = sigmoid(cls_logits); bboxes = bbox_coder.decode(...); bboxes_for_nms = xywhr2xyxyr(bev);
// box3d_multiclass_nms(bboxes, bboxes_for_nms, scores, score_thr, max_num, cfg) with use_rotate_nms (nms_gpu of the
// un-vendored mmdet3d fork / iou3d): per class keep score > score_thr, greedy suppression in descending score order of
// boxes whose rotated BEV IoU with a kept box exceeds nms_thr, concatenate the classes, keep the max_num best scores.
// The IoU arithmetic of iou3d is not in the tree: the published definition (intersection polygon of the two rotated
// rectangles / union) is restated with Sutherland–Hodgman clipping; ties in score break towards the lower box index.
//
// B200 design: everything stays on the device and every step is a flat data-parallel pass —
//   candidates (class-major list of (class, box) with score > thr) → rank inside the class by counting → pairwise
//   suppression bit matrix (one thread per (row, 64-column word), boxes staged in shared memory) → one warp per class walks
//   its rows in order with the removed-set in registers → compaction → optional global top-k by counting rank.
// K is a few thousand queries: the counting ranks (O(T * T_c)) are microseconds and avoid a sort.
#include "common.cuh"

namespace fsfb {

struct Rect {  // rotated BEV rectangle: centre, half sizes, rotation, half diagonal
  float cx, cy, hx, hy, c, s, rad;
};

// box = (x, y, z, dx, dy, dz, yaw, ...).  Geometry of mmdet3d 0.x (the reference pins mmcv-full 1.3.9 / mmdet 2.14, README.md:
// 20-22): xywhr2xyxyr keeps dx along x and dy along y, and iou3d's rotate_around_center turns the axis-aligned corners by
// x' = dx cos + dy sin, y' = -dx sin + dy cos, i.e. CLOCKWISE by yaw — the same frame the point pooling uses (rot_angle = rz + pi/2
// with l along local x; in-tree anchor of that convention: fsd_bbox_head_fsd.py:307-309, rotation_3d_in_axis(local, roi_ry + pi/2)).
__device__ __forceinline__ Rect rect_from_box(const float* __restrict__ b) {
  Rect r;
  r.cx = b[0]; r.cy = b[1];
  r.hx = 0.5f * b[3]; r.hy = 0.5f * b[4];
  r.c = cosf(b[6]); r.s = -sinf(b[6]);
  r.rad = sqrtf(r.hx * r.hx + r.hy * r.hy);
  return r;
}

// Area of the intersection of two rotated rectangles (Sutherland–Hodgman: A's corners clipped by B's four edges), computed
// in B's frame around B's centre so that all magnitudes are box-sized.
__device__ float rect_intersection(const Rect& A, const Rect& B) {
  float px[8], py[8], qx[8], qy[8];
  // A's corners expressed in B's axis-aligned frame
  const float dx = A.cx - B.cx, dy = A.cy - B.cy;
  const float ox = dx * B.c + dy * B.s, oy = -dx * B.s + dy * B.c;          // A's centre in B's frame
  const float rc = A.c * B.c + A.s * B.s, rs = A.s * B.c - A.c * B.s;      // relative rotation
  const float sx[4] = {A.hx, -A.hx, -A.hx, A.hx}, sy[4] = {A.hy, A.hy, -A.hy, -A.hy};
  int n = 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    px[i] = ox + sx[i] * rc - sy[i] * rs;
    py[i] = oy + sx[i] * rs + sy[i] * rc;
  }
  // clip against x <= hx, x >= -hx, y <= hy, y >= -hy of B
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float lim = (e < 2) ? B.hx : B.hy;
    const float sign = (e & 1) ? -1.f : 1.f;
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const int j = (i + 1 == n) ? 0 : i + 1;
      const float ai = (e < 2) ? px[i] : py[i], aj = (e < 2) ? px[j] : py[j];
      const float di = lim - sign * ai, dj = lim - sign * aj;  // >= 0: inside
      if (di >= 0.f) {
        qx[m] = px[i]; qy[m] = py[i]; ++m;
      }
      if ((di >= 0.f) != (dj >= 0.f)) {
        const float t = di / (di - dj);
        qx[m] = px[i] + t * (px[j] - px[i]);
        qy[m] = py[i] + t * (py[j] - py[i]);
        ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) { px[i] = qx[i]; py[i] = qy[i]; }
    if (n == 0) return 0.f;
  }
  float area = 0.f;
  for (int i = 0; i < n; ++i) {
    const int j = (i + 1 == n) ? 0 : i + 1;
    area += px[i] * py[j] - px[j] * py[i];
  }
  return 0.5f * fabsf(area);
}

__device__ __forceinline__ float rect_iou(const Rect& A, const Rect& B) {
  // circumscribed circles apart: no intersection (most pairs of a scene; skips the polygon clipping)
  const float dx = A.cx - B.cx, dy = A.cy - B.cy, rr = A.rad + B.rad;
  if (dx * dx + dy * dy >= rr * rr) return 0.f;
  const float inter = rect_intersection(A, B);
  const float ua = 4.f * A.hx * A.hy + 4.f * B.hx * B.hy - inter;
  return inter / fmaxf(ua, 1e-8f);
}

// scores = sigmoid(logits) (optional); flags[c][i] = score > thr; counts[c]
__global__ void __launch_bounds__(256)
    k_nms_flags(const float* __restrict__ logits, int64_t k, int C, int64_t stride, int apply_sigmoid, float thr,
                float* __restrict__ scores, uint8_t* __restrict__ flags, int32_t* __restrict__ counts) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < k * C; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / C;
    const int c = (int)(t - i * C);
    float s = __ldg(logits + i * stride + c);
    if (apply_sigmoid) s = 1.f / (1.f + expf(-s));
    scores[i * C + c] = s;
    const bool f = s > thr;
    flags[(int64_t)c * k + i] = f ? 1 : 0;
    if (f) atomicAdd(counts + c, 1);
  }
}

// class offsets = exclusive scan of counts (C <= 64), max class size
__global__ void k_nms_offsets(const int32_t* __restrict__ counts, int C, int32_t* __restrict__ offsets, int32_t* __restrict__ maxc) {
  if (threadIdx.x == 0) {
    int run = 0, mx = 0;
    for (int c = 0; c < C; ++c) {
      offsets[c] = run;
      run += counts[c];
      mx = max(mx, counts[c]);
    }
    offsets[C] = run;
    maxc[0] = mx;
  }
}

// candidate t (flat = c * k + i, ascending) → slot offsets[c] + rank among its class by (score desc, box index asc).
// A block of 256 candidates walks the candidates of the classes it touches in shared-memory tiles (score, box index): the
// counting rank is O(T * T_c) compares but only O(T * T_c / 256) gathered loads.
__global__ void __launch_bounds__(256)
    k_nms_rank(const int32_t* __restrict__ flat, int64_t T, int64_t k, int C, const float* __restrict__ scores,
               const int32_t* __restrict__ offsets, int32_t* __restrict__ sorted_box, int32_t* __restrict__ sorted_cls,
               float* __restrict__ sorted_score) {
  __shared__ float s_score[256];
  __shared__ int s_box[256];
  __shared__ int s_lo, s_hi;
  for (int64_t t0 = (int64_t)blockIdx.x * 256; t0 < T; t0 += (int64_t)gridDim.x * 256) {   // block-uniform
    const int64_t t = t0 + threadIdx.x;
    const bool live = t < T;
    int c = 0, i = 0, beg = 0, end = 0;
    float s = 0.f;
    if (live) {
      const int64_t f = flat[t];
      c = (int)(f / k);
      i = (int)(f - (int64_t)c * k);
      s = scores[(int64_t)i * C + c];
      beg = offsets[c];
      end = offsets[c + 1];
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // candidates are class-major: the block's classes span [class of t0, class of its last candidate]
      const int64_t tl = min(t0 + 255, T - 1);
      s_lo = offsets[(int)(flat[t0] / k)];
      s_hi = offsets[(int)(flat[tl] / k) + 1];
    }
    __syncthreads();
    const int lo = s_lo, hi = s_hi;
    int rank = 0;
    for (int u0 = lo; u0 < hi; u0 += 256) {
      __syncthreads();
      const int u = u0 + (int)threadIdx.x;
      if (u < hi) {
        const int64_t fu = flat[u];
        const int cu = (int)(fu / k);
        const int ju = (int)(fu - (int64_t)cu * k);
        s_box[threadIdx.x] = ju;
        s_score[threadIdx.x] = scores[(int64_t)ju * C + cu];
      }
      __syncthreads();
      if (live) {
        const int a = max(beg, u0) - u0, b = min(end, min(hi, u0 + 256)) - u0;   // this thread's class inside the tile
        for (int x = a; x < b; ++x) {
          const float sj = s_score[x];
          rank += (sj > s) | ((sj == s) & (s_box[x] < i));
        }
      }
    }
    if (live) {
      sorted_box[beg + rank] = i;
      sorted_cls[beg + rank] = c;
      sorted_score[beg + rank] = s;
    }
  }
}

// mask[row][w] bit b: column (64 w + b) of the row's class comes later in the order and overlaps it by more than thr
constexpr int kNmsCols = 64;
__global__ void __launch_bounds__(kNmsCols)
    k_nms_mask(const float* __restrict__ boxes, int64_t box_stride, const int32_t* __restrict__ sorted_box,
               const int32_t* __restrict__ sorted_cls, const int32_t* __restrict__ offsets, int64_t T, int words, float thr,
               unsigned long long* __restrict__ mask) {
  // blockIdx.x: block of 64 rows (global sorted order); blockIdx.y: 64-column word inside the row's class
  __shared__ Rect s_col[kNmsCols];
  const int64_t row = (int64_t)blockIdx.x * kNmsCols + threadIdx.x;
  const int w = blockIdx.y;
  // rows of one block may straddle two classes: every thread looks up its own class, columns are staged per class below
  const int cls = row < T ? sorted_cls[row] : -1;
  const int beg = cls >= 0 ? offsets[cls] : 0, end = cls >= 0 ? offsets[cls + 1] : 0;
  Rect me;
  if (cls >= 0) me = rect_from_box(boxes + (int64_t)sorted_box[row] * box_stride);
  const int first_cls = sorted_cls[min((int64_t)blockIdx.x * kNmsCols, T - 1)];
  const int last_cls = sorted_cls[min((int64_t)blockIdx.x * kNmsCols + kNmsCols - 1, T - 1)];
  for (int c = first_cls; c <= last_cls; ++c) {
    const int cb = offsets[c], ce = offsets[c + 1];
    const int col0 = cb + w * kNmsCols;
    __syncthreads();
    if (col0 + (int)threadIdx.x < ce) s_col[threadIdx.x] = rect_from_box(boxes + (int64_t)sorted_box[col0 + threadIdx.x] * box_stride);
    __syncthreads();
    if (cls == c && col0 < ce) {
      unsigned long long bits = 0;
      const int ncol = (col0 + kNmsCols - 1 > row) ? min(kNmsCols, ce - col0) : 0;   // columns up to `row` never count
      for (int j = 0; j < ncol; ++j) {
        const int64_t col = col0 + j;
        if (col > row && rect_iou(me, s_col[j]) > thr) bits |= 1ull << j;
      }
      mask[row * words + w] = bits;
    }
  }
  (void)beg; (void)end;
}

// one CTA of 8 warps per class: greedy pass in score order, 64 candidates (one mask word) at a time.
//   1. the block's 64 diagonal words wait in shared memory (prefetched during the previous block: they are plain mask entries);
//      warp 0 runs the 64-step chain on registers + broadcast shared-memory reads (every lane the same work: no shuffle);
//   2. all 8 warps OR the surviving rows into the removed set (shared memory, words <= kNmsMaxWords: 65 536 candidates per
//      class): warp j takes survivors j, j + 8, ..., up to 8 independent loads in flight per lane.
// Two global-memory round trips per 64 candidates are on the serial path instead of one per candidate.
constexpr int kNmsMaxWords = 1024;
constexpr int kNmsReduceWarps = 8;

__global__ void __launch_bounds__(kNmsReduceWarps * 32)
    k_nms_reduce(const unsigned long long* __restrict__ mask, const int32_t* __restrict__ offsets, int words,
                 uint8_t* __restrict__ keep) {
  __shared__ unsigned long long removed[kNmsMaxWords];
  __shared__ unsigned long long s_diag[2][64];
  __shared__ unsigned long long s_kept;
  const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int beg = offsets[c], end = offsets[c + 1];
  for (int w = tid; w < words; w += kNmsReduceWarps * 32) removed[w] = 0ull;
  auto load_diag = [&](int b0, int slot) {   // threads 0..63 of the callers' range: row r of the block starting at b0
    const int r = tid & 63;
    s_diag[slot][r] = b0 + r < end ? mask[(int64_t)(b0 + r) * words + ((b0 - beg) >> 6)] : 0ull;
  };
  if (tid < 64 && beg < end) load_diag(beg, 0);
  __syncthreads();
  int slot = 0;
  for (int b0 = beg; b0 < end; b0 += 64, slot ^= 1) {
    const int wb = (b0 - beg) >> 6;
    const int nb = min(64, end - b0);
    if (warp == 0) {
      unsigned long long rem = removed[wb], kept = 0ull;
#pragma unroll 8
      for (int r = 0; r < 64; ++r) {
        const unsigned long long d = s_diag[slot][r];
        const bool alive = !((rem >> r) & 1ull) && r < nb;
        kept |= alive ? (1ull << r) : 0ull;
        rem |= alive ? d : 0ull;
      }
      if (lane == 0) s_kept = kept;
      if (lane < nb) keep[b0 + lane] = (uint8_t)((kept >> lane) & 1ull);
      if (lane + 32 < nb) keep[b0 + lane + 32] = (uint8_t)((kept >> (lane + 32)) & 1ull);
    } else if (warp >= 2 && warp < 4 && b0 + 64 < end) {
      load_diag(b0 + 64, slot ^ 1);   // 64 threads (warps 2, 3): the next block's diagonal words
    }
    __syncthreads();
    // words up to wb are final (columns before a row never count); OR the surviving rows into the later ones
    unsigned long long mine = 0ull, k = s_kept;
    for (int j = 0; k; ++j, k &= k - 1)   // survivors j = warp (mod 8), as a bit set of row numbers
      if ((j & (kNmsReduceWarps - 1)) == warp) mine |= k & (~k + 1ull);
    if (mine) {   // warp-uniform
      const unsigned long long* rows = mask + (int64_t)b0 * words;
      int r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {   // a warp owns at most 8 of the 64 rows
        r[u] = mine ? __ffsll((long long)mine) - 1 : -1;
        mine = mine ? (mine & (mine - 1)) : 0ull;
      }
      for (int w = (wb & ~31) + lane; w < words; w += 32) {
        if (w <= wb) continue;
        unsigned long long acc = 0ull;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc |= r[u] >= 0 ? rows[(int64_t)r[u] * words + w] : 0ull;
        if (acc) atomicOr(&removed[w], acc);
      }
    }
    __syncthreads();
  }
}

// kept rows → outputs (order of `kept_idx`: class-major, score descending, ties by box index).  When more than max_num survive,
// the output is the max_num best by (score desc, emitted position asc).  The kept list is one descending run per class, so a
// candidate's rank is its index inside its own class plus, for every other class, the number of entries that beat it — a binary
// search per class over the compacted scores (ties: earlier classes win, later classes lose), instead of P x P compares.
__global__ void __launch_bounds__(256)
    k_nms_kept_scores(const int32_t* __restrict__ kept_idx, int64_t P, const float* __restrict__ sorted_score,
                      float* __restrict__ kept_score) {
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < P; t += (int64_t)gridDim.x * 256) kept_score[t] = sorted_score[kept_idx[t]];
}

constexpr int kNmsMaxClasses = 64;
__global__ void __launch_bounds__(256)
    k_nms_emit(const int32_t* __restrict__ kept_idx, int64_t P, const int32_t* __restrict__ sorted_box,
               const int32_t* __restrict__ sorted_cls, const float* __restrict__ sorted_score, const float* __restrict__ kept_score,
               const float* __restrict__ boxes, int64_t box_stride, int box_dim, int64_t max_num, float* __restrict__ out_boxes,
               float* __restrict__ out_scores, long long* __restrict__ out_labels, int32_t* __restrict__ out_box_idx) {
  __shared__ int s_kb[kNmsMaxClasses + 1];   // class c occupies [s_kb[c], s_kb[c + 1]) of the kept list
  const bool select = P > max_num;
  if (select) {
    if (threadIdx.x <= kNmsMaxClasses) {
      const int c = threadIdx.x;
      int lo = 0, hi = (int)P;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sorted_cls[kept_idx[mid]] < c) lo = mid + 1;
        else hi = mid;
      }
      s_kb[c] = lo;
    }
    __syncthreads();
  }
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < P; t += (int64_t)gridDim.x * 256) {
    const int r = kept_idx[t];
    const float s = sorted_score[r];
    const int ct = sorted_cls[r];
    int64_t dst = t;
    if (select) {
      int rank = (int)t - s_kb[ct];
      for (int c = 0; c < kNmsMaxClasses && rank < max_num; ++c) {
        const int beg = s_kb[c], end = s_kb[c + 1];
        if (beg == end || c == ct) continue;
        int lo = beg, hi = end;   // first entry of the class that does not beat (s, position t)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const float v = kept_score[mid];
          const bool beats = c < ct ? (v >= s) : (v > s);
          if (beats) lo = mid + 1;
          else hi = mid;
        }
        rank += lo - beg;
      }
      if (rank >= max_num) continue;
      dst = rank;
    }
    const float* b = boxes + (int64_t)sorted_box[r] * box_stride;
    for (int d = 0; d < box_dim; ++d) out_boxes[dst * box_dim + d] = b[d];
    out_scores[dst] = s;
    out_labels[dst] = ct;
    if (out_box_idx) out_box_idx[dst] = sorted_box[r];
  }
}

}  // namespace fsfb

extern "C" {

/* Step 1: scores (sigmoid of logits when apply_sigmoid) and the class-major candidate flags; counts dev [C+2] i32 scratch
 * (counts[0..C), then class offsets are written by step 2 elsewhere). */
int fsfb_nms_flags(const float* logits, int64_t k, int num_classes, int64_t stride, int apply_sigmoid, float score_thr,
                   float* scores, uint8_t* flags, int32_t* counts, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(k >= 0 && num_classes >= 1 && num_classes <= 64 && stride >= num_classes && counts, "nms_flags: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FSFB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * num_classes, st));
  if (k == 0) return FSFB_OK;
  FSFB_CHECK_ARG(logits && scores && flags, "nms_flags: null pointer");
  const int grid = (int)std::min<int64_t>(ceil_div(k * num_classes, 256), (int64_t)kNumSMs * 8);
  FSFB_LAUNCH(k_nms_flags, grid, 256, 0, st, logits, k, num_classes, stride, apply_sigmoid, score_thr, scores, flags, counts);
  return FSFB_OK;
}

int fsfb_nms_workspace_bytes(int64_t candidates, int max_class, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && candidates >= 0 && max_class >= 0, "nms_workspace_bytes: bad argument");
  const int words = (int)ceil_div(std::max(max_class, 1), 64);
  Workspace ws(nullptr, 0);
  ws.take<int32_t>(std::max<int64_t>(candidates, 1));                     // sorted_box
  ws.take<int32_t>(std::max<int64_t>(candidates, 1));                     // sorted_cls
  ws.take<float>(std::max<int64_t>(candidates, 1));                       // sorted_score
  ws.take<unsigned long long>((size_t)std::max<int64_t>(candidates, 1) * words);  // mask
  ws.take<int32_t>(66);                                                   // offsets | maxc
  *bytes = ws.used;
  return FSFB_OK;
}

/* Step 2: rank the T candidates (`flat` = ascending indices of the set flags), build the suppression matrix and run the greedy
 * pass.  max_class = size of the largest class (host value, from the counts of step 1).  keep dev [T] u8 in sorted order;
 * the sorted arrays stay in the workspace for step 3. */
int fsfb_nms_suppress(const float* boxes, int64_t k, int64_t box_stride, const float* scores, int num_classes, const int32_t* flat,
                      int64_t candidates, const int32_t* counts, int max_class, float nms_thr, uint8_t* keep, void* workspace,
                      size_t workspace_bytes, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(k >= 0 && box_stride >= 7 && num_classes >= 1 && num_classes <= 64 && candidates >= 0 && max_class >= 0 &&
                     max_class <= 65536,
                 "nms_suppress: bad argument (at most 65536 candidates per class)");
  if (candidates == 0) return FSFB_OK;
  FSFB_CHECK_ARG(boxes && scores && flat && counts && keep, "nms_suppress: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int words = (int)ceil_div(std::max(max_class, 1), 64);
  Workspace ws(workspace, workspace_bytes);
  int32_t* sorted_box = ws.take<int32_t>(candidates);
  int32_t* sorted_cls = ws.take<int32_t>(candidates);
  float* sorted_score = ws.take<float>(candidates);
  unsigned long long* mask = ws.take<unsigned long long>((size_t)candidates * words);
  int32_t* offsets = ws.take<int32_t>(66);
  if (!ws.ok()) {
    set_error("nms_suppress: workspace too small (%zu given, %zu needed)", workspace_bytes, ws.used);
    return FSFB_ERR_CAPACITY;
  }
  FSFB_LAUNCH(k_nms_offsets, 1, 32, 0, st, counts, num_classes, offsets, offsets + 65);
  const int grid = (int)std::min<int64_t>(ceil_div(candidates, 256), (int64_t)kNumSMs * 8);
  FSFB_LAUNCH(k_nms_rank, grid, 256, 0, st, flat, candidates, k, num_classes, scores, offsets, sorted_box, sorted_cls, sorted_score);
  FSFB_CUDA(cudaMemsetAsync(mask, 0, (size_t)candidates * words * sizeof(unsigned long long), st));
  dim3 mgrid((unsigned)ceil_div(candidates, kNmsCols), (unsigned)words);
  FSFB_LAUNCH(k_nms_mask, mgrid, kNmsCols, 0, st, boxes, box_stride, sorted_box, sorted_cls, offsets, candidates, words, nms_thr, mask);
  FSFB_LAUNCH(k_nms_reduce, num_classes, kNmsReduceWarps * 32, 0, st, mask, offsets, words, keep);
  return FSFB_OK;
}

/* Step 3: emit the kept candidates (`kept_idx` = ascending indices of the set keep flags): boxes [min(P, max_num), box_dim],
 * scores, labels (class index), optional source box index; class-major in descending score, or — when more than max_num
 * survive — the max_num best scores in descending order (box3d_multiclass_nms). */
int fsfb_nms_emit(const float* boxes, int64_t box_stride, int box_dim, const int32_t* kept_idx, int64_t kept, int64_t candidates,
                  int max_class, int64_t max_num, float* out_boxes, float* out_scores, long long* out_labels, int32_t* out_box_idx,
                  void* workspace, size_t workspace_bytes, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(box_stride >= box_dim && box_dim >= 7 && kept >= 0 && candidates >= kept && max_num >= 1, "nms_emit: bad argument");
  if (kept == 0) return FSFB_OK;
  FSFB_CHECK_ARG(boxes && kept_idx && out_boxes && out_scores && out_labels, "nms_emit: null pointer");
  const int words = (int)ceil_div(std::max(max_class, 1), 64);
  Workspace ws(workspace, workspace_bytes);
  int32_t* sorted_box = ws.take<int32_t>(candidates);
  int32_t* sorted_cls = ws.take<int32_t>(candidates);
  float* sorted_score = ws.take<float>(candidates);
  unsigned long long* mask_mem = ws.take<unsigned long long>((size_t)candidates * words);
  if (!ws.ok()) {
    set_error("nms_emit: workspace does not match the one given to nms_suppress");
    return FSFB_ERR_CAPACITY;
  }
  // the suppression matrix is dead by now: its memory holds the compacted scores of the kept list
  float* kept_score = reinterpret_cast<float*>(mask_mem);
  const int grid = (int)std::min<int64_t>(ceil_div(kept, 256), (int64_t)kNumSMs * 8);
  if (kept > max_num)
    FSFB_LAUNCH(k_nms_kept_scores, grid, 256, 0, (cudaStream_t)stream, kept_idx, kept, sorted_score, kept_score);
  FSFB_LAUNCH(k_nms_emit, grid, 256, 0, (cudaStream_t)stream, kept_idx, kept, sorted_box, sorted_cls, sorted_score, kept_score, boxes,
              box_stride, box_dim, max_num, out_boxes, out_scores, out_labels, out_box_idx);
  return FSFB_OK;
}

}  // extern "C"
