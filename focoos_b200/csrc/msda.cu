// Multi-scale deformable attention core (RT-DETR decoder cross-attention), fully fused:
//   softmax over levels*points  +  sampling-location arithmetic  +  bilinear gather  +  weighted sum.
// One warp per (batch, query, head); lane = channel of the 32-wide head, so each bilinear tap is one
// coalesced 64/128-byte row segment of `value` (L2-resident: B*S*256 elements).  Gather/latency bound.
#include "common.cuh"

namespace fb200 {

constexpr int MSDA_MAX_LEVELS = 4;
struct MsdaShapes { int h[MSDA_MAX_LEVELS], w[MSDA_MAX_LEVELS], start[MSDA_MAX_LEVELS]; };

template <typename TV, typename TOA, typename TO>
__global__ void __launch_bounds__(256) msda_kernel(const TV* __restrict__ value, int v_pitch, const TOA* __restrict__ oa,
                                                   int oa_pitch, const float* __restrict__ ref, MsdaShapes sh, int L, int P,
                                                   int S, int Q, int heads, int64_t total, TO* __restrict__ out, int out_pitch) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (b*Q + q)*heads + h
  if (wid >= total) return;
  const int h = wid % heads;
  const int64_t bq = wid / heads;
  const int b = bq / Q;
  const int LP = L * P;
  const TOA* row = oa + bq * oa_pitch;
  // attention logits -> softmax (every lane holds logit of point `lane` if lane < LP)
  float logit = -INFINITY;
  if (lane < LP) logit = to_f(row[heads * LP * 2 + h * LP + lane]);
  const float mx = warp_max(logit);
  const float e = lane < LP ? expf(logit - mx) : 0.f;
  const float wgt = e / warp_sum(e);
  // sampling offsets for point `lane`
  float ox = 0.f, oy = 0.f;
  if (lane < LP) {
    ox = to_f(row[(h * LP + lane) * 2 + 0]);
    oy = to_f(row[(h * LP + lane) * 2 + 1]);
  }
  const float4 r = *reinterpret_cast<const float4*>(ref + bq * 4);
  // loc = ref_xy + off / P * ref_wh * 0.5   (modelling.py:871-874, same association order)
  const float locx = r.x + ox / (float)P * r.z * 0.5f;
  const float locy = r.y + oy / (float)P * r.w * 0.5f;
  const TV* vb = value + (int64_t)b * S * v_pitch + h * 32 + lane;
  float acc = 0.f;
#pragma unroll 4  // 16 independent gathers in flight per lane: the kernel is L2/HBM-latency bound
  for (int pt = 0; pt < LP; ++pt) {
    const int lvl = pt / P;
    const float lx = __shfl_sync(0xffffffffu, locx, pt), ly = __shfl_sync(0xffffffffu, locy, pt);
    const float aw = __shfl_sync(0xffffffffu, wgt, pt);
    const int H = sh.h[lvl], W = sh.w[lvl];
    // grid = 2*loc-1 ; grid_sample unnormalise (align_corners=False): ((g+1)*size-1)/2   (deformable.py:16,24-30)
    const float gx = 2.f * lx - 1.f, gy = 2.f * ly - 1.f;
    const float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float tx = ix - fx, ty = iy - fy;
    const float w_nw = (1.f - tx) * (1.f - ty), w_ne = tx * (1.f - ty), w_sw = (1.f - tx) * ty, w_se = tx * ty;
    const TV* vl = vb + (int64_t)sh.start[lvl] * v_pitch;
    float s = 0.f;
    const bool xin0 = x0 >= 0 && x0 < W, xin1 = x1 >= 0 && x1 < W, yin0 = y0 >= 0 && y0 < H, yin1 = y1 >= 0 && y1 < H;
    if (yin0 && xin0) s += to_f(vl[((int64_t)y0 * W + x0) * v_pitch]) * w_nw;
    if (yin0 && xin1) s += to_f(vl[((int64_t)y0 * W + x1) * v_pitch]) * w_ne;
    if (yin1 && xin0) s += to_f(vl[((int64_t)y1 * W + x0) * v_pitch]) * w_sw;
    if (yin1 && xin1) s += to_f(vl[((int64_t)y1 * W + x1) * v_pitch]) * w_se;
    acc = fmaf(s, aw, acc);
  }
  out[bq * out_pitch + h * 32 + lane] = from_f<TO>(acc);
}

}  // namespace fb200
using namespace fb200;

extern "C" int fb200_msda(const void* value, int v_dtype, int v_pitch, const void* oa, int oa_dtype, int oa_pitch,
                          const float* ref, const int* shapes_host, int L, int P, int B, int S, int Q, int heads, void* out,
                          int out_dtype, int out_pitch, void* stream) {
  FB_CHECK_ARG(value && oa && ref && shapes_host && out, "msda: null pointer");
  FB_CHECK_ARG(L >= 1 && L <= MSDA_MAX_LEVELS && L * P <= 32, "msda: levels*points must be <= 32 (L=%d P=%d)", L, P);
  MsdaShapes sh;
  int start = 0;
  for (int l = 0; l < L; ++l) {
    sh.h[l] = shapes_host[2 * l];
    sh.w[l] = shapes_host[2 * l + 1];
    sh.start[l] = start;
    start += sh.h[l] * sh.w[l];
  }
  FB_CHECK_ARG(start == S, "msda: sum of level sizes (%d) != S (%d)", start, S);
  FB_CHECK_ARG(oa_pitch >= heads * L * P * 3, "msda: oa_pitch too small");
  const int64_t total = (int64_t)B * Q * heads;
  const unsigned grid = (unsigned)cdiv(total, 8);
  cudaStream_t st = (cudaStream_t)stream;
#define MSDA_LAUNCH(TV, TOA, TO) \
  msda_kernel<TV, TOA, TO><<<grid, 256, 0, st>>>((const TV*)value, v_pitch, (const TOA*)oa, oa_pitch, ref, sh, L, P, S, Q, heads, total, (TO*)out, out_pitch)
  if (v_dtype == FB200_F32 && oa_dtype == FB200_F32 && out_dtype == FB200_F32) MSDA_LAUNCH(float, float, float);
  else if (v_dtype == FB200_F16 && oa_dtype == FB200_F32 && out_dtype == FB200_F16) MSDA_LAUNCH(__half, float, __half);
  else if (v_dtype == FB200_F16 && oa_dtype == FB200_F16 && out_dtype == FB200_F16) MSDA_LAUNCH(__half, __half, __half);
  else { set_error("msda: unsupported dtype combination %d/%d/%d", v_dtype, oa_dtype, out_dtype); return FB200_ERR_UNSUPPORTED; }
#undef MSDA_LAUNCH
  FB_CHECK_LAUNCH("msda");
  return FB200_OK;
}
