// LayerNorm(+residual) and small-sequence multi-head attention (L = 300/400, head_dim = 32).
#include "common.cuh"

namespace fb200 {

// one warp per row; C <= 1024, C % 4 == 0. Two-pass (mean, then centred variance) in registers.
template <typename T, int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        T* __restrict__ out, int64_t M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nv = C / 4;  // vectors per row
  float v[MAXV][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nv) {
      load4(x + row * C + vi * 4, v[i]);
      if (res) {
        float r[4];
        load4(res + row * C + vi * 4, r);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] += r[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) s += v[i][j];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nv) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = v[i][j] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nv) {
      float g[4], b[4], o[4];
      load4(gamma + vi * 4, g);
      load4(beta + vi * 4, b);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
      store4(out + row * C + vi * 4, o);
    }
  }
}

// grid: (B*heads, ceil(Lq/QT)); block 256 = 8 warps; each warp owns QT/8 queries.
// smem: Ks[Lk][33] + Vs[Lk][32] + per-warp probabilities P[8][Lk] + per-warp q[8][32]  (fp32)
constexpr int ATT_QT = 64;
template <typename T>
__global__ void __launch_bounds__(256) attention_kernel(const T* __restrict__ q, int q_pitch, const T* __restrict__ k,
                                                        int k_pitch, const T* __restrict__ v, int v_pitch, T* __restrict__ out,
                                                        int out_pitch, int Lq, int Lk, int heads, float scale) {
  extern __shared__ float sm[];
  float* Ks = sm;                   // [Lk][33]
  float* Vs = Ks + (size_t)Lk * 33; // [Lk][32]
  float* Ps = Vs + (size_t)Lk * 32; // [8][Lk]
  float* Qs = Ps + (size_t)8 * Lk;  // [8][32]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // cooperative load of K,V for (b,h): Lk rows x 32 channels; 8 threads x 4 elements per row
  for (int i = threadIdx.x; i < Lk * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 4;
    float kv[4], vv[4];
    load4(k + ((int64_t)b * Lk + r) * k_pitch + h * 32 + c, kv);
    load4(v + ((int64_t)b * Lk + r) * v_pitch + h * 32 + c, vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) { Ks[r * 33 + c + j] = kv[j]; Vs[r * 32 + c + j] = vv[j]; }
  }
  __syncthreads();
  const int q0 = blockIdx.y * ATT_QT;
  float* P = Ps + (size_t)warp * Lk;
  float* Q = Qs + warp * 32;
  for (int qi = q0 + warp; qi < min(q0 + ATT_QT, Lq); qi += 8) {
    Q[lane] = to_f(q[((int64_t)b * Lq + qi) * q_pitch + h * 32 + lane]) * scale;  // torch scales q before QK^T
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) s = fmaf(Q[d], Ks[j * 33 + d], s);
      P[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float e = expf(P[j] - mx);
      P[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o = 0.f;
    for (int j = 0; j < Lk; ++j) o = fmaf(P[j], Vs[j * 32 + lane], o);
    out[((int64_t)b * Lq + qi) * out_pitch + h * 32 + lane] = from_f<T>(o / sum);
    __syncwarp();
  }
}

}  // namespace fb200
using namespace fb200;

extern "C" int fb200_layernorm(const void* x, const void* res, const float* gamma, const float* beta, void* out, int dtype,
                               int64_t M, int C, float eps, void* stream) {
  FB_CHECK_ARG(x && gamma && beta && out, "layernorm: null pointer");
  FB_CHECK_ARG(C % 4 == 0 && C <= 1024 && M > 0, "layernorm: C must be a multiple of 4 and <= 1024 (got %d)", C);
  const unsigned grid = (unsigned)cdiv(M, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 256) { FB_DISPATCH_DTYPE(dtype, T, (layernorm_kernel<T, 2><<<grid, 256, 0, st>>>((const T*)x, (const T*)res, gamma, beta, (T*)out, M, C, eps))); }
  else { FB_DISPATCH_DTYPE(dtype, T, (layernorm_kernel<T, 8><<<grid, 256, 0, st>>>((const T*)x, (const T*)res, gamma, beta, (T*)out, M, C, eps))); }
  FB_CHECK_LAUNCH("layernorm");
  return FB200_OK;
}

extern "C" int fb200_attention(const void* q, int q_pitch, const void* k, int k_pitch, const void* v, int v_pitch, void* out,
                               int out_pitch, int dtype, int B, int Lq, int Lk, int heads, int head_dim, float scale,
                               void* stream) {
  FB_CHECK_ARG(q && k && v && out, "attention: null pointer");
  FB_CHECK_ARG(head_dim == 32, "attention: head_dim must be 32 (got %d)", head_dim);
  FB_CHECK_ARG(q_pitch % 4 == 0 && k_pitch % 4 == 0 && v_pitch % 4 == 0, "attention: pitches must be multiples of 4");
  const size_t smem = ((size_t)Lk * 33 + (size_t)Lk * 32 + (size_t)8 * Lk + 8 * 32) * sizeof(float);
  FB_CHECK_ARG(smem <= 227 * 1024, "attention: Lk=%d does not fit shared memory", Lk);
  dim3 grid(B * heads, (unsigned)cdiv(Lq, ATT_QT));
  cudaStream_t st = (cudaStream_t)stream;
  static bool configured = false;
  if (!configured) {  // raise the dynamic-smem ceiling once (not inside a stream capture)
    cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(attention_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    configured = true;
  }
  if (dtype == FB200_F32) {
    attention_kernel<float><<<grid, 256, smem, st>>>((const float*)q, q_pitch, (const float*)k, k_pitch, (const float*)v, v_pitch, (float*)out, out_pitch, Lq, Lk, heads, scale);
  } else if (dtype == FB200_F16) {
    attention_kernel<__half><<<grid, 256, smem, st>>>((const __half*)q, q_pitch, (const __half*)k, k_pitch, (const __half*)v, v_pitch, (__half*)out, out_pitch, Lq, Lk, heads, scale);
  } else { set_error("attention: bad dtype"); return FB200_ERR_INVALID; }
  FB_CHECK_LAUNCH("attention");
  return FB200_OK;
}
