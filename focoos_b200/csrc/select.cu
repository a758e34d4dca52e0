// Query selection and post-processing: row max, exact top-k (radix select + bitonic sort, ties broken by
// ascending index), row gather, box arithmetic and the fused DETR post-process (top-k over Q*C scores,
// label/query decode, threshold count, scale to the original image, round-half-even -> int32).
#include "common.cuh"

namespace fb200 {

template <typename T>
__global__ void row_select_kernel(const T* __restrict__ x, const uint8_t* __restrict__ valid, const float* __restrict__ fill,
                                  T* __restrict__ out, int64_t rows, int S, int C) {
  const int cv = C / 4;
  const int64_t total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int c = (i % cv) * 4;
    float v[4];
    if (valid[r % S]) load4(x + r * C + c, v);
    else load4(fill + c, v);
    store4(out + r * C + c, v);
  }
}

template <typename T>
__global__ void rowmax_kernel(const T* __restrict__ x, int64_t rows, int N, int pitch, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float m = -INFINITY;
  for (int j = lane; j < N; j += 32) m = fmaxf(m, to_f(x[row * pitch + j]));
  m = warp_max(m);
  if (lane == 0) out[row] = m;
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, int S, int C, int pitch, const int* __restrict__ idx, int K,
                                   int64_t total_rows, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // b*K + k
  if (row >= total_rows) return;
  const int b = row / K;
  const int s = idx[row];
  const T* sp = src + ((int64_t)b * S + s) * pitch;
  for (int c = lane * 4; c < C; c += 128) {
    float v[4];
    load4(sp + c, v);
    store4(out + row * C + c, v);
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float inv_sigmoidf_(float x) {
  x = fminf(fmaxf(x, 0.f), 1.f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.f - x, 1e-5f));
}

__global__ void box_op_kernel(int mode, const float* __restrict__ x, const float* __restrict__ ref, const int* __restrict__ idx,
                              float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 0) { out[i] = sigmoidf_(x[i]); return; }
  if (mode == 1) { out[i] = sigmoidf_(x[i] + inv_sigmoidf_(ref[i])); return; }
  if (mode == 2) { out[i] = x[i] + ref[(int64_t)idx[i >> 2] * 4 + (i & 3)]; return; }
  // mode 3: i indexes a box
  const float4 b = reinterpret_cast<const float4*>(x)[i];
  reinterpret_cast<float4*>(out)[i] = make_float4(b.x - 0.5f * b.z, b.y - 0.5f * b.w, b.x + 0.5f * b.z, b.y + 0.5f * b.w);
}

// ------------------------------------------------------------------------------------------------
// block-wide exact top-k.  1024 threads.  Result: sorted (key desc, idx asc) in smem `cand` (first K).
// ------------------------------------------------------------------------------------------------
constexpr int TOPK_THREADS = 1024;
constexpr int TOPK_MAXK = 1024;

__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

struct TopkSmem {
  unsigned long long cand[TOPK_MAXK];
  int hist[256];
  int warp_cnt[32];
  uint32_t prefix, mask;
  int remaining, n_greater, eq_base;
};

__device__ void topk_block(const float* __restrict__ x, int N, int K, TopkSmem& s) {
  const int tid = threadIdx.x;
  int Kp = 1;
  while (Kp < K) Kp <<= 1;
  if (tid == 0) { s.prefix = 0u; s.mask = 0u; s.remaining = K; s.n_greater = 0; s.eq_base = 0; }
  for (int i = tid; i < Kp; i += TOPK_THREADS) s.cand[i] = 0ull;
  __syncthreads();
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += TOPK_THREADS) s.hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s.prefix, mask = s.mask;
    for (int i = tid; i < N; i += TOPK_THREADS) {
      const uint32_t key = f2key(x[i]);
      if ((key & mask) == prefix) atomicAdd(&s.hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int rem = s.remaining, d = 255;
      for (; d > 0; --d) {
        if (s.hist[d] >= rem) break;
        rem -= s.hist[d];
      }
      s.remaining = rem;
      s.prefix = prefix | ((uint32_t)d << shift);
      s.mask = mask | (255u << shift);
    }
    __syncthreads();
  }
  const uint32_t thr = s.prefix;      // key of the K-th largest element
  const int need_eq = s.remaining;    // how many elements == thr are kept (lowest indices first)
  const int n_gt = K - need_eq;
  // pass A: everything strictly greater (order irrelevant, sorted afterwards)
  for (int i = tid; i < N; i += TOPK_THREADS) {
    const uint32_t key = f2key(x[i]);
    if (key > thr) {
      const int slot = atomicAdd(&s.n_greater, 1);
      s.cand[slot] = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    }
  }
  // pass B: equal elements in ascending index order
  const int lane = tid & 31, warp = tid >> 5;
  for (int base = 0; base < N; base += TOPK_THREADS) {
    const int i = base + tid;
    const bool eq = (i < N) && (f2key(x[i]) == thr);
    if (!__syncthreads_or(eq)) continue;
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s.warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s.eq_base;
    for (int w2 = 0; w2 < warp; ++w2) off += s.warp_cnt[w2];
    const int rank = off + __popc(bal & ((1u << lane) - 1u));
    if (eq && rank < need_eq)
      s.cand[n_gt + rank] = ((unsigned long long)thr << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w2 = 0; w2 < 32; ++w2) tot += s.warp_cnt[w2];
      s.eq_base += tot;
    }
    __syncthreads();
    if (s.eq_base >= need_eq) break;
  }
  __syncthreads();
  // bitonic sort, descending, Kp elements
  for (int size = 2; size <= Kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < Kp; i += TOPK_THREADS) {
        const int j = i ^ stride;
        if (j > i) {
          const unsigned long long a = s.cand[i], b = s.cand[j];
          const bool desc = ((i & size) == 0);
          if (desc ? (a < b) : (a > b)) { s.cand[i] = b; s.cand[j] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_kernel(const float* __restrict__ x, int N, int K, int* __restrict__ out_idx,
                                                            float* __restrict__ out_val) {
  __shared__ TopkSmem s;
  const int b = blockIdx.x;
  topk_block(x + (int64_t)b * N, N, K, s);
  for (int i = threadIdx.x; i < K; i += TOPK_THREADS) {
    const unsigned long long c = s.cand[i];
    out_idx[(int64_t)b * K + i] = (int)(0xffffffffu - (uint32_t)(c & 0xffffffffull));
    if (out_val) out_val[(int64_t)b * K + i] = key2f((uint32_t)(c >> 32));
  }
}

__global__ void __launch_bounds__(TOPK_THREADS) detr_postprocess_kernel(const float* __restrict__ scores, const float* __restrict__ boxes,
                                                                        const int* __restrict__ sizes, int Q, int C, int K, float thr,
                                                                        float* __restrict__ out_scores, int* __restrict__ out_labels,
                                                                        int* __restrict__ out_boxes, int* __restrict__ out_query,
                                                                        int* __restrict__ out_count) {
  __shared__ TopkSmem s;
  __shared__ int cnt;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) cnt = 0;
  topk_block(scores + (int64_t)b * Q * C, Q * C, K, s);
  const float Wimg = (float)sizes[b * 2 + 1], Himg = (float)sizes[b * 2 + 0];
  for (int i = threadIdx.x; i < K; i += TOPK_THREADS) {
    const unsigned long long c = s.cand[i];
    const int flat = (int)(0xffffffffu - (uint32_t)(c & 0xffffffffull));
    const float sc = key2f((uint32_t)(c >> 32));
    const int label = flat % C, q = flat / C;
    const float4 bx = reinterpret_cast<const float4*>(boxes)[(int64_t)b * Q + q];
    const int64_t o = (int64_t)b * K + i;
    out_scores[o] = sc;
    out_labels[o] = label;
    out_query[o] = q;
    out_boxes[o * 4 + 0] = (int)rintf(bx.x * Wimg);
    out_boxes[o * 4 + 1] = (int)rintf(bx.y * Himg);
    out_boxes[o * 4 + 2] = (int)rintf(bx.z * Wimg);
    out_boxes[o * 4 + 3] = (int)rintf(bx.w * Himg);
    if (sc > thr) atomicAdd(&cnt, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) out_count[b] = cnt;
}

// Evaluator post-process (fai_detr/processor.py:121-144 + detector_postprocess :19-57): per image top-k over the flattened [Q*C] scores (NO threshold),
// label / query decode, boxes scaled to the dataset entry's (height, width) as FLOATS, clipped to the image, empty boxes dropped; the survivors are
// written compacted in descending-score order.  One CTA per image.
__global__ void __launch_bounds__(TOPK_THREADS) detr_eval_postprocess_kernel(const float* __restrict__ scores, const float* __restrict__ boxes,
                                                                             const int* __restrict__ sizes, int Q, int C, int K,
                                                                             float* __restrict__ out_scores, int* __restrict__ out_labels,
                                                                             float* __restrict__ out_boxes, int* __restrict__ out_count) {
  __shared__ TopkSmem s;
  __shared__ int warp_tot[TOPK_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  topk_block(scores + (int64_t)b * Q * C, Q * C, K, s);
  const float Wimg = (float)sizes[b * 2 + 1], Himg = (float)sizes[b * 2 + 0];
  // K <= TOPK_MAXK = TOPK_THREADS: one candidate per thread
  float sc = 0.f; int label = 0; float4 bx = make_float4(0.f, 0.f, 0.f, 0.f); int keep = 0;
  if (tid < K) {
    const unsigned long long c = s.cand[tid];
    const int flat = (int)(0xffffffffu - (uint32_t)(c & 0xffffffffull));
    sc = key2f((uint32_t)(c >> 32));
    label = flat % C;
    bx = reinterpret_cast<const float4*>(boxes)[(int64_t)b * Q + flat / C];
    bx.x = fminf(fmaxf(bx.x * Wimg, 0.f), Wimg); bx.z = fminf(fmaxf(bx.z * Wimg, 0.f), Wimg);   // Boxes.scale then Boxes.clip
    bx.y = fminf(fmaxf(bx.y * Himg, 0.f), Himg); bx.w = fminf(fmaxf(bx.w * Himg, 0.f), Himg);
    keep = (bx.z - bx.x > 0.f) && (bx.w - bx.y > 0.f);                                          // Boxes.nonempty(threshold=0)
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_tot[warp] = __popc(m);
  __syncthreads();
  int base = 0, total = 0;
  for (int w = 0; w < TOPK_THREADS / 32; ++w) { if (w < warp) base += warp_tot[w]; total += warp_tot[w]; }
  if (keep) {
    const int64_t o = (int64_t)b * K + base + __popc(m & ((1u << lane) - 1u));
    out_scores[o] = sc;
    out_labels[o] = label;
    reinterpret_cast<float4*>(out_boxes)[o] = bx;
  }
  if (tid == 0) out_count[b] = total;
}

static inline unsigned grid_for(int64_t total, int threads) {
  int64_t g = cdiv(total, threads);
  const int64_t cap = 148LL * 32;
  return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace fb200
using namespace fb200;

extern "C" int fb200_row_select(const void* x, const uint8_t* valid, const float* fill, void* out, int dtype, int64_t rows,
                                int S, int C, void* stream) {
  FB_CHECK_ARG(x && valid && fill && out && C % 4 == 0 && S > 0, "row_select: bad arguments");
  FB_DISPATCH_DTYPE(dtype, T, (row_select_kernel<T><<<grid_for(rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, valid, fill, (T*)out, rows, S, C)));
  FB_CHECK_LAUNCH("row_select");
  return FB200_OK;
}

extern "C" int fb200_rowmax(const void* x, int dtype, int64_t rows, int N, int pitch, float* out, void* stream) {
  FB_CHECK_ARG(x && out && N > 0 && pitch >= N, "rowmax: bad arguments");
  FB_DISPATCH_DTYPE(dtype, T, (rowmax_kernel<T><<<(unsigned)cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>((const T*)x, rows, N, pitch, out)));
  FB_CHECK_LAUNCH("rowmax");
  return FB200_OK;
}

extern "C" int fb200_topk(const float* x, int B, int N, int K, int* out_idx, float* out_val, void* stream) {
  FB_CHECK_ARG(x && out_idx && B > 0, "topk: null pointer");
  FB_CHECK_ARG(K >= 1 && K <= TOPK_MAXK && K <= N, "topk: need 1 <= K <= min(N, %d) (K=%d N=%d)", TOPK_MAXK, K, N);
  topk_kernel<<<B, TOPK_THREADS, 0, (cudaStream_t)stream>>>(x, N, K, out_idx, out_val);
  FB_CHECK_LAUNCH("topk");
  return FB200_OK;
}

extern "C" int fb200_gather_rows(const void* src, int dtype, int B, int S, int C, int pitch, const int* idx, int K, void* out,
                                 void* stream) {
  FB_CHECK_ARG(src && idx && out && C % 4 == 0 && pitch % 4 == 0 && pitch >= C, "gather_rows: bad arguments");
  const int64_t rows = (int64_t)B * K;
  FB_DISPATCH_DTYPE(dtype, T, (gather_rows_kernel<T><<<(unsigned)cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>((const T*)src, S, C, pitch, idx, K, rows, (T*)out)));
  FB_CHECK_LAUNCH("gather_rows");
  return FB200_OK;
}

extern "C" int fb200_box_op(int mode, const float* x, const float* ref, const int* idx, float* out, int64_t n, void* stream) {
  FB_CHECK_ARG(x && out && n > 0 && mode >= 0 && mode <= 3, "box_op: bad arguments");
  FB_CHECK_ARG(mode != 1 || ref, "box_op: mode 1 needs ref");
  FB_CHECK_ARG(mode != 2 || (ref && idx && n % 4 == 0), "box_op: mode 2 needs anchors (ref), idx and n %% 4 == 0");
  FB_CHECK_ARG(mode != 3 || n % 4 == 0, "box_op: mode 3 needs n %% 4 == 0");
  const int64_t work = (mode == 3) ? n / 4 : n;
  box_op_kernel<<<(unsigned)cdiv(work, 256), 256, 0, (cudaStream_t)stream>>>(mode, x, ref, idx, out, work);
  FB_CHECK_LAUNCH("box_op");
  return FB200_OK;
}

extern "C" int fb200_detr_eval_postprocess(const float* scores, const float* boxes, const int* sizes, int B, int Q, int C, int K, float* out_scores,
                                           int* out_labels, float* out_boxes, int* out_count, void* stream) {
  FB_CHECK_ARG(scores && boxes && sizes && out_scores && out_labels && out_boxes && out_count, "detr_eval_postprocess: null pointer");
  FB_CHECK_ARG(B > 0 && K >= 1 && K <= TOPK_MAXK && (int64_t)K <= (int64_t)Q * C, "detr_eval_postprocess: bad K=%d", K);
  detr_eval_postprocess_kernel<<<B, TOPK_THREADS, 0, (cudaStream_t)stream>>>(scores, boxes, sizes, Q, C, K, out_scores, out_labels, out_boxes, out_count);
  FB_CHECK_LAUNCH("detr_eval_postprocess");
  return FB200_OK;
}

extern "C" int fb200_detr_postprocess(const float* scores, const float* boxes, const int* sizes, int B, int Q, int C, int K,
                                      float threshold, float* out_scores, int* out_labels, int* out_boxes, int* out_query,
                                      int* out_count, void* stream) {
  FB_CHECK_ARG(scores && boxes && sizes && out_scores && out_labels && out_boxes && out_query && out_count, "detr_postprocess: null pointer");
  FB_CHECK_ARG(B > 0 && K >= 1 && K <= TOPK_MAXK && (int64_t)K <= (int64_t)Q * C, "detr_postprocess: bad K=%d", K);
  detr_postprocess_kernel<<<B, TOPK_THREADS, 0, (cudaStream_t)stream>>>(scores, boxes, sizes, Q, C, K, threshold, out_scores, out_labels, out_boxes, out_query, out_count);
  FB_CHECK_LAUNCH("detr_postprocess");
  return FB200_OK;
}
