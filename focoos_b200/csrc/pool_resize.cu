// HBM-bound NHWC helpers: max/avg pooling, bilinear resize (align_corners=False), add.
// One thread handles 4 consecutive channels of one output pixel (8/16-byte vector accesses,
// consecutive threads -> consecutive channels -> coalesced rows).
#include "common.cuh"

namespace fb200 {

// 16-byte vectors: 4 floats or 8 halves per thread (these kernels are pure HBM streams; 8-byte fp16 accesses left half the bandwidth unused)
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&t.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&t.w));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]), c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
  }
};

template <typename T>
__global__ void maxpool3x3s2_v16_kernel(const T* __restrict__ x, int B, int H, int W, int C, int Ho, int Wo, T* __restrict__ out) {
  constexpr int N = Vec16<T>::N;
  const int cv = C / N;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (i % cv) * N;
    const int64_t pix = i / cv;
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * 2 - 1 + kh;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * 2 - 1 + kw;
        if (wi < 0 || wi >= W) continue;
        float v[8];
        Vec16<T>::load(x + (((int64_t)b * H + hi) * W + wi) * C + c, v);
#pragma unroll
        for (int j = 0; j < N; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    }
    Vec16<T>::store(out + pix * C + c, m);
  }
}

template <typename T>
__global__ void avgpool2x2_v16_kernel(const T* __restrict__ x, int B, int H, int W, int C, int Ho, int Wo, T* __restrict__ out) {
  constexpr int N = Vec16<T>::N;
  const int cv = C / N;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (i % cv) * N;
    const int64_t pix = i / cv;
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    const int h0 = ho * 2, w0 = wo * 2, h1 = min(h0 + 2, H), w1 = min(w0 + 2, W);
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (int hi = h0; hi < h1; ++hi)
      for (int wi = w0; wi < w1; ++wi) {
        float v[8];
        Vec16<T>::load(x + (((int64_t)b * H + hi) * W + wi) * C + c, v);
#pragma unroll
        for (int j = 0; j < N; ++j) s[j] += v[j];
      }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
#pragma unroll
    for (int j = 0; j < N; ++j) s[j] *= inv;
    Vec16<T>::store(out + pix * C + c, s);
  }
}

template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, int B, int H, int W, int C, int Ho, int Wo, T* __restrict__ out) {
  const int cv = C / 4;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (i % cv) * 4;
    const int64_t pix = i / cv;
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * 2 - 1 + kh;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * 2 - 1 + kw;
        if (wi < 0 || wi >= W) continue;
        float v[4];
        load4(x + (((int64_t)b * H + hi) * W + wi) * C + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    }
    store4(out + pix * C + c, m);
  }
}

template <typename T>
__global__ void avgpool2x2_kernel(const T* __restrict__ x, int B, int H, int W, int C, int Ho, int Wo, T* __restrict__ out) {
  const int cv = C / 4;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (i % cv) * 4;
    const int64_t pix = i / cv;
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    const int h0 = ho * 2, w0 = wo * 2, h1 = min(h0 + 2, H), w1 = min(w0 + 2, W);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int hi = h0; hi < h1; ++hi)
      for (int wi = w0; wi < w1; ++wi) {
        float v[4];
        load4(x + (((int64_t)b * H + hi) * W + wi) * C + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] += v[j];
      }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));  // ceil_mode, pad 0: divisor = clipped window
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] *= inv;
    store4(out + pix * C + c, s);
  }
}

// src index per ATen upsample_bilinear2d (align_corners=False): src = max((dst+0.5)*scale-0.5, 0), scale = in/out
template <typename T>
__global__ void resize_bilinear_kernel(const T* __restrict__ x, int B, int H, int W, int C, int x_pitch, T* __restrict__ out,
                                       int Ho, int Wo, int out_pitch, float sh, float sw) {
  const int cv = C / 4;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (i % cv) * 4;
    const int64_t pix = i / cv;
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    const float fh = fmaxf(((float)ho + 0.5f) * sh - 0.5f, 0.f), fw = fmaxf(((float)wo + 0.5f) * sw - 0.5f, 0.f);
    const int h0 = (int)fh, w0 = (int)fw;
    const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
    const float lh1 = fh - (float)h0, lh0 = 1.f - lh1, lw1 = fw - (float)w0, lw0 = 1.f - lw1;
    const T* base = x + (int64_t)b * H * W * x_pitch + c;
    float v00[4], v01[4], v10[4], v11[4], r[4];
    load4(base + ((int64_t)h0 * W + w0) * x_pitch, v00);
    load4(base + ((int64_t)h0 * W + w1) * x_pitch, v01);
    load4(base + ((int64_t)h1 * W + w0) * x_pitch, v10);
    load4(base + ((int64_t)h1 * W + w1) * x_pitch, v11);
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = lh0 * (lw0 * v00[j] + lw1 * v01[j]) + lh1 * (lw0 * v10[j] + lw1 * v11[j]);
    store4(out + pix * out_pitch + c, r);
  }
}

// Image resize of the pre-processing step (processor/base_processor.py:284-294: F.interpolate(images, size, mode="bilinear", align_corners=False) on the float image): a
// whole batch of decoded images - uint8 NHWC [B,H,W,3] or float NCHW [B,3,H,W] - straight to the model's float NCHW input, one launch, no padded intermediate.  Same
// source-index arithmetic and interpolation order as resize_bilinear_kernel above.
template <bool U8_NHWC>
__global__ void image_resize_kernel(const void* __restrict__ img, int B, int H, int W, float* __restrict__ out, int Ho, int Wo, float sh, float sw) {
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int wo = i % Wo, ho = (i / Wo) % Ho, b = i / ((int64_t)Wo * Ho);
    const float fh = fmaxf(((float)ho + 0.5f) * sh - 0.5f, 0.f), fw = fmaxf(((float)wo + 0.5f) * sw - 0.5f, 0.f);
    const int h0 = (int)fh, w0 = (int)fw;
    const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
    const float lh1 = fh - (float)h0, lh0 = 1.f - lh1, lw1 = fw - (float)w0, lw0 = 1.f - lw1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v00, v01, v10, v11;
      if (U8_NHWC) {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(img) + (int64_t)b * H * W * 3 + c;
        v00 = (float)p[((int64_t)h0 * W + w0) * 3]; v01 = (float)p[((int64_t)h0 * W + w1) * 3];
        v10 = (float)p[((int64_t)h1 * W + w0) * 3]; v11 = (float)p[((int64_t)h1 * W + w1) * 3];
      } else {
        const float* p = reinterpret_cast<const float*>(img) + ((int64_t)b * 3 + c) * H * W;
        v00 = p[(int64_t)h0 * W + w0]; v01 = p[(int64_t)h0 * W + w1];
        v10 = p[(int64_t)h1 * W + w0]; v11 = p[(int64_t)h1 * W + w1];
      }
      out[(((int64_t)b * 3 + c) * Ho + ho) * Wo + wo] = lh0 * (lw0 * v00 + lw1 * v01) + lh1 * (lw0 * v10 + lw1 * v11);
    }
  }
}

__global__ void resize_bilinear_h8_kernel(const __half* __restrict__ x, int B, int H, int W, int C, int x_pitch, __half* __restrict__ out,
                                          int Ho, int Wo, int out_pitch, float sh, float sw) {
  const int cv = C / 8;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (i % cv) * 8;
    const int64_t pix = i / cv;
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    const float fh = fmaxf(((float)ho + 0.5f) * sh - 0.5f, 0.f), fw = fmaxf(((float)wo + 0.5f) * sw - 0.5f, 0.f);
    const int h0 = (int)fh, w0 = (int)fw;
    const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
    const float lh1 = fh - (float)h0, lh0 = 1.f - lh1, lw1 = fw - (float)w0, lw0 = 1.f - lw1;
    const __half* base = x + (int64_t)b * H * W * x_pitch + c;
    float v00[8], v01[8], v10[8], v11[8], r[8];
    Vec16<__half>::load(base + ((int64_t)h0 * W + w0) * x_pitch, v00);
    Vec16<__half>::load(base + ((int64_t)h0 * W + w1) * x_pitch, v01);
    Vec16<__half>::load(base + ((int64_t)h1 * W + w0) * x_pitch, v10);
    Vec16<__half>::load(base + ((int64_t)h1 * W + w1) * x_pitch, v11);
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = lh0 * (lw0 * v00[j] + lw1 * v01[j]) + lh1 * (lw0 * v10[j] + lw1 * v11[j]);
    Vec16<__half>::store(out + pix * out_pitch + c, r);
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n4, int64_t bn4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float va[4], vb[4];
    load4(a + i * 4, va);
    load4(b + (i % bn4) * 4, vb);
#pragma unroll
    for (int j = 0; j < 4; ++j) va[j] += vb[j];
    store4(out + i * 4, va);
  }
}

// fp32 -> [hi | lo] fp16 pair (split-precision operands): hi = fp16(x), lo = fp16(x - hi)
__global__ void split_f32_pair_kernel(const float* __restrict__ x, int64_t rows, int C, int x_pitch, __half* __restrict__ out) {
  const int cv = C / 4;
  const int64_t total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int c = (i % cv) * 4;
    float v[4], hi[4], lo[4];
    load4(x + r * x_pitch + c, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half h = __float2half_rn(v[j]);
      hi[j] = __half2float(h);
      lo[j] = v[j] - hi[j];
    }
    store4(out + r * 2 * C + c, hi);
    store4(out + r * 2 * C + C + c, lo);
  }
}

// 8 values per thread: two 16-byte loads, two 16-byte stores (C % 8 == 0, 16-byte aligned rows)
__global__ void split_f32_pair8_kernel(const float* __restrict__ x, int64_t rows, int C, int x_pitch, __half* __restrict__ out) {
  const int cv = C / 8;
  const int64_t total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int c = (i % cv) * 8;
    const float4 a = *reinterpret_cast<const float4*>(x + r * x_pitch + c), b = *reinterpret_cast<const float4*>(x + r * x_pitch + c + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = __half2float(__float2half_rn(v[j]));
      lo[j] = v[j] - hi[j];
    }
    Vec16<__half>::store(out + r * 2 * C + c, hi);
    Vec16<__half>::store(out + r * 2 * C + C + c, lo);
  }
}

// ---- pair-format (fp32 value = fp16 hi plane + fp16 lo plane, see FB200_F16PAIR) variants of the three HBM-bound spatial operators between convs of the fp32-accurate
// mode: 8 channels per thread (16 B of hi + 16 B of lo), arithmetic in fp32 on hi + lo, the result re-split.  mode 0: max_pool2d(3,2,1); 1: AvgPool2d(2,2,ceil_mode);
// 2: bilinear resize (align_corners=False).
struct PairPtr { const __half* hi; int64_t lo_off; int pitch; };
__device__ __forceinline__ void pair_load8(const __half* hi, int64_t lo_off, float (&v)[8]) {
  float a[8], b[8];
  Vec16<__half>::load(hi, a);
  Vec16<__half>::load(hi + lo_off, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = a[j] + b[j];
}
__device__ __forceinline__ void pair_store8(__half* hi, int64_t lo_off, const float (&v)[8]) {
  float h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { h[j] = __half2float(__float2half_rn(v[j])); l[j] = v[j] - h[j]; }
  Vec16<__half>::store(hi, h);
  Vec16<__half>::store(hi + lo_off, l);
}

template <int MODE>
__global__ void pair_pool_kernel(const __half* __restrict__ x, int64_t x_lo, int x_pitch, int B, int H, int W, int C, __half* __restrict__ out, int64_t o_lo, int o_pitch,
                                 int Ho, int Wo, float sh, float sw) {
  const int cv = C / 8;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (i % cv) * 8;
    const int64_t pix = i / cv;
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    const __half* base = x + (int64_t)b * H * W * x_pitch + c;
    float r[8];
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = -INFINITY;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hi = ho * 2 - 1 + kh;
        if (hi < 0 || hi >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int wi = wo * 2 - 1 + kw;
          if (wi < 0 || wi >= W) continue;
          float v[8];
          pair_load8(base + ((int64_t)hi * W + wi) * x_pitch, x_lo, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = fmaxf(r[j], v[j]);
        }
      }
    } else if (MODE == 1) {
      const int h0 = ho * 2, w0 = wo * 2, h1 = min(h0 + 2, H), w1 = min(w0 + 2, W);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = 0.f;
      for (int hi = h0; hi < h1; ++hi)
        for (int wi = w0; wi < w1; ++wi) {
          float v[8];
          pair_load8(base + ((int64_t)hi * W + wi) * x_pitch, x_lo, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] += v[j];
        }
      const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] *= inv;
    } else {
      const float fh = fmaxf(((float)ho + 0.5f) * sh - 0.5f, 0.f), fw = fmaxf(((float)wo + 0.5f) * sw - 0.5f, 0.f);
      const int h0 = (int)fh, w0 = (int)fw;
      const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
      const float lh1 = fh - (float)h0, lh0 = 1.f - lh1, lw1 = fw - (float)w0, lw0 = 1.f - lw1;
      float v00[8], v01[8], v10[8], v11[8];
      pair_load8(base + ((int64_t)h0 * W + w0) * x_pitch, x_lo, v00);
      pair_load8(base + ((int64_t)h0 * W + w1) * x_pitch, x_lo, v01);
      pair_load8(base + ((int64_t)h1 * W + w0) * x_pitch, x_lo, v10);
      pair_load8(base + ((int64_t)h1 * W + w1) * x_pitch, x_lo, v11);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = lh0 * (lw0 * v00[j] + lw1 * v01[j]) + lh1 * (lw0 * v10[j] + lw1 * v11[j]);
    }
    pair_store8(out + pix * o_pitch + c, o_lo, r);
  }
}

static inline unsigned grid_for(int64_t total, int threads) {
  int64_t g = cdiv(total, threads);
  const int64_t cap = 148LL * 32;
  return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace fb200
using namespace fb200;

extern "C" int fb200_maxpool3x3s2(const void* x, int dtype, int B, int H, int W, int C, void* out, void* stream) {
  FB_CHECK_ARG(x && out && C % 4 == 0, "maxpool: null pointer or C %% 4 != 0");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (dtype == FB200_F16 && C % 8 == 0 && al16) {
    maxpool3x3s2_v16_kernel<__half><<<grid_for(total / 2, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)x, B, H, W, C, Ho, Wo, (__half*)out);
    FB_CHECK_LAUNCH("maxpool3x3s2");
    return FB200_OK;
  }
  FB_DISPATCH_DTYPE(dtype, T, (maxpool3x3s2_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, B, H, W, C, Ho, Wo, (T*)out)));
  FB_CHECK_LAUNCH("maxpool3x3s2");
  return FB200_OK;
}

extern "C" int fb200_avgpool2x2_ceil(const void* x, int dtype, int B, int H, int W, int C, void* out, void* stream) {
  FB_CHECK_ARG(x && out && C % 4 == 0, "avgpool: null pointer or C %% 4 != 0");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (dtype == FB200_F16 && C % 8 == 0 && al16) {
    avgpool2x2_v16_kernel<__half><<<grid_for(total / 2, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)x, B, H, W, C, Ho, Wo, (__half*)out);
    FB_CHECK_LAUNCH("avgpool2x2");
    return FB200_OK;
  }
  FB_DISPATCH_DTYPE(dtype, T, (avgpool2x2_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, B, H, W, C, Ho, Wo, (T*)out)));
  FB_CHECK_LAUNCH("avgpool2x2");
  return FB200_OK;
}

extern "C" int fb200_resize_bilinear(const void* x, int dtype, int B, int H, int W, int C, int x_pitch, void* out, int Ho,
                                     int Wo, int out_pitch, void* stream) {
  FB_CHECK_ARG(x && out && C % 4 == 0 && x_pitch % 4 == 0 && out_pitch % 4 == 0, "resize: null pointer or C/pitch %% 4 != 0");
  FB_CHECK_ARG(x_pitch >= C && out_pitch >= C && Ho > 0 && Wo > 0, "resize: bad shape");
  const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  if (dtype == FB200_F16 && C % 8 == 0 && x_pitch % 8 == 0 && out_pitch % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    resize_bilinear_h8_kernel<<<grid_for(total / 2, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)x, B, H, W, C, x_pitch, (__half*)out, Ho, Wo, out_pitch, sh, sw);
    FB_CHECK_LAUNCH("resize_bilinear");
    return FB200_OK;
  }
  FB_DISPATCH_DTYPE(dtype, T, (resize_bilinear_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, B, H, W, C, x_pitch, (T*)out, Ho, Wo, out_pitch, sh, sw)));
  FB_CHECK_LAUNCH("resize_bilinear");
  return FB200_OK;
}

extern "C" int fb200_image_resize(const void* images, int u8_nhwc, int B, int H, int W, float* out_nchw, int Ho, int Wo, void* stream) {
  FB_CHECK_ARG(images && out_nchw && B > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "image_resize: bad arguments");
  const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  const int64_t total = (int64_t)B * Ho * Wo;
  if (u8_nhwc) image_resize_kernel<true><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(images, B, H, W, out_nchw, Ho, Wo, sh, sw);
  else image_resize_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(images, B, H, W, out_nchw, Ho, Wo, sh, sw);
  FB_CHECK_LAUNCH("image_resize");
  return FB200_OK;
}

extern "C" int fb200_pair_pool(int mode, const void* x, int64_t x_lo_off, int x_pitch, int B, int H, int W, int C, void* out, int64_t out_lo_off, int out_pitch, int Ho, int Wo,
                               void* stream) {
  FB_CHECK_ARG(x && out && mode >= 0 && mode <= 2 && C % 8 == 0 && x_pitch % 8 == 0 && out_pitch % 8 == 0 && x_lo_off % 8 == 0 && out_lo_off % 8 == 0, "pair_pool: bad arguments");
  FB_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 && B > 0 && Ho > 0 && Wo > 0, "pair_pool: unaligned planes or empty output");
  FB_CHECK_ARG(mode != 0 || (Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1), "pair_pool: max_pool2d(3,2,1) output size");
  FB_CHECK_ARG(mode != 1 || (Ho == (H + 1) / 2 && Wo == (W + 1) / 2), "pair_pool: AvgPool2d(2,2,ceil) output size");
  const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
  const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  const __half* xp = static_cast<const __half*>(x);
  __half* op = static_cast<__half*>(out);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) pair_pool_kernel<0><<<grid_for(total, 256), 256, 0, st>>>(xp, x_lo_off, x_pitch, B, H, W, C, op, out_lo_off, out_pitch, Ho, Wo, sh, sw);
  else if (mode == 1) pair_pool_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(xp, x_lo_off, x_pitch, B, H, W, C, op, out_lo_off, out_pitch, Ho, Wo, sh, sw);
  else pair_pool_kernel<2><<<grid_for(total, 256), 256, 0, st>>>(xp, x_lo_off, x_pitch, B, H, W, C, op, out_lo_off, out_pitch, Ho, Wo, sh, sw);
  FB_CHECK_LAUNCH("pair_pool");
  return FB200_OK;
}

extern "C" int fb200_add(const void* a, const void* b, void* out, int dtype, int64_t rows, int64_t brows, int C, void* stream) {
  FB_CHECK_ARG(a && b && out && C % 4 == 0 && brows > 0 && rows % brows == 0, "add: bad arguments");
  const int64_t n4 = rows * C / 4, bn4 = brows * C / 4;
  FB_DISPATCH_DTYPE(dtype, T, (add_kernel<T><<<grid_for(n4, 256), 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (T*)out, n4, bn4)));
  FB_CHECK_LAUNCH("add");
  return FB200_OK;
}

extern "C" int fb200_split_f32_pair(const float* x, int64_t rows, int C, int x_pitch, void* out, void* stream) {
  FB_CHECK_ARG(x && out && C % 4 == 0 && x_pitch % 4 == 0 && x_pitch >= C, "split_f32_pair: bad arguments");
  if (C % 8 == 0 && x_pitch % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    split_f32_pair8_kernel<<<grid_for(rows * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(x, rows, C, x_pitch, (__half*)out);
  else
    split_f32_pair_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, rows, C, x_pitch, (__half*)out);
  FB_CHECK_LAUNCH("split_f32_pair");
  return FB200_OK;
}
