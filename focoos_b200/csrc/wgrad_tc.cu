// tcgen05 weight-gradient GEMM for sm_100a, fp32-accurate through split-precision fp16 products.
//
//   dW[co, kh, kw, ci] = sum over output pixels p of  dY[p, co] * X[pix(p, kh, kw), ci]          (stride-1 convs and linears)
//
// GEMM view per filter tap: D[M = Cout, N = Cin] = A^T B with the reduction over PIXELS.  In NHWC both operands are stored
// pixel-major with channels contiguous, i.e. they are "MN-major" for the tensor core: a TMA box of 64 pixels x 64 channels lands in
// shared memory as 8 swizzle atoms of (8 pixels x 128 B) - exactly the canonical SWIZZLE_128B MN-major layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp), so no transpose is ever materialised:
//   SBO = 1024 B (next 8-pixel group), LBO = 8192 B (next 64-channel block = the next TMA box), instruction descriptor with
//   a_major = b_major = MN.  The X box is the dY box shifted by the tap offset; out-of-bounds rows/columns (conv zero padding, ragged
//   pixel tiles) are zero-filled by the TMA unit for BOTH operands, so ragged tiles contribute exact zeros.
// fp32 accuracy: operands are the [hi | lo] fp16 pairs of the fp32 tensors (fb200_split_f32_pair); each 64-pixel K chunk issues
//   dY_hi*X_hi + dY_hi*X_lo + dY_lo*X_hi  into the same fp32 TMEM accumulator (error ~2^-21 relative, like the forward mode).
// Parallelism: work item = (128 x BLOCK_N weight tile, tap, pixel split); persistent CTAs, warp-specialised
//   (TMA producer / MMA issuer / TMEM allocator / 4 epilogue warps), 3-stage smem ring, 2 TMEM accumulator stages;
//   split partials are reduced in a fixed order by wgrad_reduce_kernel (reproducible, no atomics).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace fb200 {
namespace wg {

constexpr int BLOCK_M = 128;   // Cout tile
constexpr int BLOCK_K = 64;    // pixels per chunk
constexpr int STAGES = 3;
constexpr int BOX_BYTES = BLOCK_K * 64 * 2;  // one TMA box: 64 pixels x 64 channels fp16 = 8 KiB

struct WParams {
  int Cout, Cin, KH, KW, pad, stride;
  int BW, BH, tiles_w, tiles_h, B;     // pixel tile rectangle (BW*BH == 64) and counts
  int m_tiles, n_tiles, taps, splits;  // work decomposition
  int pt_total, pt_per_split;          // pixel tiles
  int total_items;
  float* part;                         // [splits][Cout][KH*KW][Cin]
  int single;                          // 1: plain fp16 operands, ONE product per chunk (the reference's fp16-autocast numerics class); 0: [hi | lo] pairs, three products
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();  // a descriptor / phase bug must not hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// MN-major SWIZZLE_128B shared-memory matrix descriptor: LBO = stride between 64-element MN blocks, SBO = stride between 8-row K groups
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(BOX_BYTES >> 4) << 16;  // leading byte offset: next 64-channel block (next TMA box)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: next group of 8 pixels
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = F32, A = B = F16, A and B MN-major (bits 15, 16), N at bits 17.., M = 128 at bits 24..
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BLOCK_N> __host__ __device__ constexpr int a_boxes() { return 2 * (BLOCK_M / 64); }   // hi + lo, 64 channels per box
template <int BLOCK_N> __host__ __device__ constexpr int b_boxes() { return 2 * (BLOCK_N / 64); }
template <int BLOCK_N> __host__ __device__ constexpr int stage_bytes() { return (a_boxes<BLOCK_N>() + b_boxes<BLOCK_N>()) * BOX_BYTES; }
template <int BLOCK_N> constexpr int smem_bytes() { return STAGES * stage_bytes<BLOCK_N>() + (2 * STAGES + 4) * 8 + 16 + 1024; }

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BOXES = a_boxes<BLOCK_N>(), B_BOXES = b_boxes<BLOCK_N>();
  constexpr int STAGE_BYTES = stage_bytes<BLOCK_N>();
  constexpr int A_HALF = (A_BOXES / 2) * BOX_BYTES, B_HALF = (B_BOXES / 2) * BOX_BYTES;  // bytes of the hi (or lo) part
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  // work item -> (split, tap, n tile, m tile); n fastest so CTAs running together share dY boxes in L2
  auto decode = [&](int item, int& mt, int& nt, int& tap, int& split) {
    nt = item % p.n_tiles; item /= p.n_tiles;
    mt = item % p.m_tiles; item /= p.m_tiles;
    tap = item % p.taps;
    split = item / p.taps;
  };

  if (warp == 0) {
    if (lane == 0) {  // ================================================================= TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int mt, nt, tap, split;
        decode(item, mt, nt, tap, split);
        const int m0 = mt * BLOCK_M, n0 = nt * BLOCK_N;
        const int kh = tap / p.KW, kw = tap - kh * p.KW;
        const int pt_begin = split * p.pt_per_split, pt_end = min(p.pt_total, pt_begin + p.pt_per_split);
        for (int pt = pt_begin; pt < pt_end; ++pt) {
          const int img = pt / tiles_per_img, rem = pt - img * tiles_per_img;
          const int h0 = (rem / p.tiles_w) * p.BH, w0 = (rem % p.tiles_w) * p.BW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(p.single ? STAGE_BYTES / 2 : STAGE_BYTES));
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BOXES * BOX_BYTES;
          const int halves = p.single ? 1 : 2;
          for (int half = 0; half < halves; ++half) {  // 0 = hi, 1 = lo (channel offset C in the pair tensor)
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_4d(&tmap_dy, &full_bar[stage], sa + half * A_HALF + j * BOX_BYTES, half * p.Cout + m0 + j * 64, w0, h0, img);
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_4d(&tmap_x, &full_bar[stage], sb + half * B_HALF + j * BOX_BYTES, half * p.Cin + n0 + j * 64, w0 * p.stride + kw - p.pad,
                          h0 * p.stride + kh - p.pad, img);  // stride 2: the X map traverses every second pixel (TMA element strides)
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ================================================================= MMA issuer
      constexpr uint32_t idesc = make_idesc_mn(BLOCK_N);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int mt, nt, tap, split;
        decode(item, mt, nt, tap, split);
        const int pt_begin = split * p.pt_per_split, pt_end = min(p.pt_total, pt_begin + p.pt_per_split);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        bool first = true;
        for (int pt = pt_begin; pt < pt_end; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BOXES * BOX_BYTES;
          const uint64_t a_hi = make_desc_mn(sa), a_lo = make_desc_mn(sa + A_HALF);
          const uint64_t b_hi = make_desc_mn(sb), b_lo = make_desc_mn(sb + B_HALF);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {  // 16 pixels = two 8-pixel groups = 2048 B: +128 in 16-byte units
            const uint64_t adv = (uint64_t)(k * 128);
            umma_f16(tmem_d, a_hi + adv, b_hi + adv, idesc, first ? 0u : 1u);
            first = false;
            if (!p.single) {
              umma_f16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_f16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {  // =================================================================== epilogue: TMEM -> split partial in global memory
    const int ew = warp & 3;         // TMEM lane quarter this warp may access
    const int row = ew * 32 + lane;  // accumulator row = output channel within the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int mt, nt, tap, split;
      decode(item, mt, nt, tap, split);
      const int co = mt * BLOCK_M + row, n0 = nt * BLOCK_N;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      float* dst = p.part + (((int64_t)split * p.Cout + co) * p.taps + tap) * p.Cin + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), r);
        if (co < p.Cout) {
          if (n0 + c0 + 32 <= p.Cin && (p.Cin & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < p.Cin) dst[c0 + j] = __uint_as_float(r[j]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(int64_t)k * n + i];
  out[i] = accumulate ? out[i] + s : s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// pair tensor [B,H,W,2C] fp16 -> 4-D map (channel, w, h, b) with a 64-channel x BW x BH box, SWIZZLE_128B, OOB = zero
static int encode_pair(CUtensorMap* m, const void* base, int C2, int W, int H, int B, int BW, int BH, const char* what, int stride = 1) {
  EncodeTiledFn fn = get_encode();
  if (!fn) { set_error("wgrad_tc: cuTensorMapEncodeTiled unavailable"); return FB200_ERR_CUDA; }
  const cuuint64_t gdim[4] = {(cuuint64_t)C2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)C2 * 2, (cuuint64_t)C2 * 2 * W, (cuuint64_t)C2 * 2 * W * H};
  // with a traversal stride s the box spans s*BW x s*BH input pixels and delivers ceil(s*BW / s) x ceil(s*BH / s) = BW x BH of them
  const cuuint32_t box[4] = {64, (cuuint32_t)(BW * stride), (cuuint32_t)(BH * stride), 1}, estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("wgrad_tc: cuTensorMapEncodeTiled(%s) failed with %d (C2=%d W=%d H=%d B=%d box %dx%d)", what, (int)r, C2, W, H, B, BW, BH); return FB200_ERR_CUDA; }
  return FB200_OK;
}

static int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

// pixel rectangle with EXACTLY 64 pixels (power-of-two width), maximising the covered fraction of the Ho x Wo map
static void choose_rect(int Ho, int Wo, int* BW, int* BH) {
  double best = -1.0;
  for (int bw = 1; bw <= 64; bw *= 2) {
    const int bh = 64 / bw;
    const double tiles = (double)((Wo + bw - 1) / bw) * (double)((Ho + bh - 1) / bh);
    const double eff = (double)Wo * Ho / (tiles * 64.0);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && bw > *BW)) { best = eff; *BW = bw; *BH = bh; }
  }
}

struct Plan { int BW, BH, tiles_w, tiles_h, m_tiles, n_tiles, block_n, splits, pt_total, pt_per_split; };
static Plan make_plan(int B, int Ho, int Wo, int Cin, int Cout, int taps) {
  Plan pl;
  pl.BW = 1; pl.BH = 64;
  choose_rect(Ho, Wo, &pl.BW, &pl.BH);
  pl.tiles_w = (Wo + pl.BW - 1) / pl.BW; pl.tiles_h = (Ho + pl.BH - 1) / pl.BH;
  pl.pt_total = B * pl.tiles_w * pl.tiles_h;
  pl.block_n = Cin > 64 ? 128 : 64;
  pl.m_tiles = (Cout + BLOCK_M - 1) / BLOCK_M;
  pl.n_tiles = (Cin + pl.block_n - 1) / pl.block_n;
  const int64_t base = (int64_t)pl.m_tiles * pl.n_tiles * taps;
  int64_t splits = (2 * (int64_t)num_sms() + base - 1) / base;           // ~2 work items per SM
  const int64_t max_splits = pl.pt_total / 8 > 0 ? pl.pt_total / 8 : 1;   // at least 8 pixel tiles (512 pixels) per item
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  pl.pt_per_split = (int)((pl.pt_total + splits - 1) / splits);
  pl.splits = (pl.pt_total + pl.pt_per_split - 1) / pl.pt_per_split;
  return pl;
}

}  // namespace wg
}  // namespace fb200

using namespace fb200;

/* 1 if fb200_conv_wgrad_tc supports the shape (else the caller uses fb200_conv_wgrad) */
extern "C" int fb200_conv_wgrad_tc_supported(int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW, int stride, int pad) {
  if ((stride != 1 && stride != 2) || KH != KW || (KH != 1 && KH != 3) || 2 * pad != KH - 1) return 0;
  if (Ho != (H + 2 * pad - KH) / stride + 1 || Wo != (W + 2 * pad - KW) / stride + 1) return 0;
  if (stride == 2 && KH != 3) return 0;
  if (Cin % 8 != 0 || Cout % 8 != 0) return 0;            // 16-byte global strides of the fp16 pair tensors
  if ((int64_t)B * Ho * Wo < 512) return 0;
  return 1;
}

extern "C" int64_t fb200_conv_wgrad_tc_workspace_bytes(int B, int Ho, int Wo, int Cin, int Cout, int KH, int KW) {
  const wg::Plan pl = wg::make_plan(B, Ho, Wo, Cin, Cout, KH * KW);
  return (int64_t)pl.splits * Cout * KH * KW * Cin * 4 + 16;
}

static int wgrad_tc_launch(const void* x_pair, int B, int H, int W, int Cin, const void* dy_pair, int Cout, int KH, int KW, int stride, int pad, float* dw,
                           int accumulate, void* workspace, void* stream, int single) {
  FB_CHECK_ARG(x_pair && dy_pair && dw && workspace, "conv_wgrad_tc: null pointer");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  FB_CHECK_ARG(fb200_conv_wgrad_tc_supported(B, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad), "conv_wgrad_tc: unsupported shape (B=%d H=%d W=%d Cin=%d Cout=%d k=%d s=%d)", B, H, W,
               Cin, Cout, KH, stride);
  FB_CHECK_ARG(((reinterpret_cast<uintptr_t>(x_pair) | reinterpret_cast<uintptr_t>(dy_pair) | reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0,
               "conv_wgrad_tc: pointers must be 16-byte aligned");
  using namespace wg;
  const int taps = KH * KW;
  const Plan pl = make_plan(B, Ho, Wo, Cin, Cout, taps);
  CUtensorMap tdy, tx;
  const int planes = single ? 1 : 2;  // channels per pixel of the operand tensors: C (plain fp16) or 2C ([hi | lo] pair)
  int rc = encode_pair(&tdy, dy_pair, planes * Cout, Wo, Ho, B, pl.BW, pl.BH, "dY");
  if (rc) return rc;
  rc = encode_pair(&tx, x_pair, planes * Cin, W, H, B, pl.BW, pl.BH, "X", stride);
  if (rc) return rc;
  WParams p;
  p.Cout = Cout; p.Cin = Cin; p.KH = KH; p.KW = KW; p.pad = pad; p.stride = stride;
  p.BW = pl.BW; p.BH = pl.BH; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.B = B;
  p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.taps = taps; p.splits = pl.splits;
  p.pt_total = pl.pt_total; p.pt_per_split = pl.pt_per_split;
  const int64_t items = (int64_t)pl.m_tiles * pl.n_tiles * taps * pl.splits;
  FB_CHECK_ARG(items <= 0x7fffffffLL, "conv_wgrad_tc: too many work items");
  p.total_items = (int)items;
  p.part = reinterpret_cast<float*>(workspace);
  p.single = single;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)(items < num_sms() ? items : num_sms());
  if (pl.block_n == 128) {
    static bool cfg = false;
    if (!cfg) { cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128>()); cfg = true; }
    static_assert(smem_bytes<128>() <= 227 * 1024, "shared memory budget exceeded");
    wgrad_tc_kernel<128><<<grid, 256, smem_bytes<128>(), st>>>(tdy, tx, p);
  } else {
    static bool cfg = false;
    if (!cfg) { cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<64>()); cfg = true; }
    wgrad_tc_kernel<64><<<grid, 256, smem_bytes<64>(), st>>>(tdy, tx, p);
  }
  FB_CHECK_LAUNCH("conv_wgrad_tc");
  const int64_t n = (int64_t)Cout * taps * Cin;
  wgrad_reduce_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(p.part, pl.splits, n, dw, accumulate);
  FB_CHECK_LAUNCH("conv_wgrad_tc(reduce)");
  return FB200_OK;
}

/* x_pair [B,H,W,2*Cin] fp16, dy_pair [B,Ho,Wo,2*Cout] fp16 (both from fb200_split_f32_pair, dense) -> dw [Cout][KH][KW][Cin] fp32 */
extern "C" int fb200_conv_wgrad_tc(const void* x_pair, int B, int H, int W, int Cin, const void* dy_pair, int Cout, int KH, int KW, int stride, int pad, float* dw,
                                   int accumulate, void* workspace, void* stream) {
  return wgrad_tc_launch(x_pair, B, H, W, Cin, dy_pair, Cout, KH, KW, stride, pad, dw, accumulate, workspace, stream, 0);
}

/* the same GEMM on PLAIN fp16 operands (x [B,H,W,Cin], dy [B,Ho,Wo,Cout], dense), one tensor-core product, fp32 accumulation and output: the arithmetic of
   the reference's training under torch.autocast(fp16) (trainer/trainer.py:735), used by the "amp" training precision */
extern "C" int fb200_conv_wgrad_tc_f16(const void* x, int B, int H, int W, int Cin, const void* dy, int Cout, int KH, int KW, int stride, int pad, float* dw,
                                       int accumulate, void* workspace, void* stream) {
  return wgrad_tc_launch(x, B, H, W, Cin, dy, Cout, KH, KW, stride, pad, dw, accumulate, workspace, stream, 1);
}
