// fp32 3x3 convolution (stride 1, pad 1) on the TF32 tensor cores — the VAE's convolutions (SURVEY.md §8 row f, "next":
// diffusers AutoencoderKL Encoder/Decoder ResnetBlock2D.conv1/conv2, Upsample2D.conv, conv_in-side 128..512 channels).
// The reference upcasts the SDXL VAE to fp32 (src/tryon_pipeline.py:913-915,1076-1093: force_upcast) and PyTorch runs
// its fp32 convolutions as TF32 products with fp32 accumulation by default; this kernel keeps exactly that arithmetic
// class (operands rounded to TF32 by the tensor core, fp32 accumulate, fp32 bias add, fp32 out) and replaces cuDNN's
// ~45 TFLOP/s on these shapes.
//
// Same machine as gemm2.cu, restated for 4-byte operands: NHWC fp32 activations == channels_last memory of the torch
// tensor, implicit GEMM through a 4-D TMA map (box = 32 channels x 128 pixels, shifted per tap, zero OOB fill = the
// padding), weights [9][Cout][Cin] fp32, a cluster of two CTAs per 256 x BN tile (each CTA stages its 128 pixels and
// half of the weight tile per 32-channel slab; 128-byte smem rows, SWIZZLE_128B), `tcgen05.mma.cta_group::2.kind::tf32`
// (K = 8 per instruction), fp32 accumulators double-buffered in TMEM, persistent tile loop.
//   warp 0: TMA producer   warp 1: MMA issuer (leader) + TMEM alloc   warps 2..9: epilogue (acc + bias (+ residual) -> 32x32 fp32
//   chunk staged in 128B-swizzled smem -> TMA store; rows/channels outside the tensor are clipped by the map)
#include "common.cuh"
#include "gemm_common.cuh"
#include "host.h"

namespace vton {

struct ConvF32Params {
  const float* bias;     // [Cout] or null
  const float* residual; // [B,H,W,Cout] fp32 NHWC or null: out = (acc + bias) + residual (the resnet's `x + h`)
  int B, H, W, Cout;
  int bw, bh, bb;        // pixel box of a 128-row tile (bw*bh*bb == 128)
  int tiles_x, tiles_y;
  int cin_slabs;         // Cin / 32
  int n_tiles;
};

constexpr int CT_BK = 32;                  // channels per slab (128 bytes of fp32)
constexpr int CT_A_BYTES = 128 * 128;      // 128 pixels x 128 B

template <int BN, int STAGES>
struct SmemT {
  static constexpr int B_BYTES = (BN / 2) * 128;
  static constexpr int STAGE_BYTES = CT_A_BYTES + B_BYTES;
  static constexpr int STORE_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int STORE_BYTES = 8 * 4096;            // one 32 x 32 fp32 staging buffer per epilogue warp
  static constexpr int BAR_OFFSET = STORE_OFFSET + STORE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
};

__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4)            // c_format = F32
         | (2u << 7)          // a_format = TF32
         | (2u << 10)         // b_format = TF32
         | ((N >> 3) << 17)   // n_dim
         | ((M >> 4) << 24);  // m_dim   (A and B K-major)
}
__device__ __forceinline__ void tc_mma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// F16IN: activations and weights arrive as fp16 (64 channels per 128-byte slab row, kind::f16 MMAs at twice the TF32 rate);
// accumulators, bias, residual and output stay fp32. The fp16 operand has the 10-bit mantissa TF32 would round an fp32 operand
// to, so the arithmetic class is the same; the producer is the VAE's GroupNorm(+SiLU) storing fp16 (norm_f32.cu, HALF_OUT).
template <int BN, int STAGES, bool F16IN>
__global__ void __launch_bounds__(320, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const ConvF32Params p, int m_pairs) {
  constexpr int BK = F16IN ? 64 : CT_BK;   // channels per slab (128 bytes)
  using L = SmemT<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int total_tiles = m_pairs * p.n_tiles;
  const int slabs = 9 * p.cin_slabs;
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 16);   // 8 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto tile_origin = [&](int m_tile, int* x0, int* y0, int* b0) {
    *x0 = (m_tile % p.tiles_x) * p.bw;
    *y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.bh;
    *b0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.bb;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters) {
        const int n_tile = t % p.n_tiles;
        const int m_tile = 2 * (t / p.n_tiles) + static_cast<int>(rank);
        const int n0 = n_tile * BN + static_cast<int>(rank) * (BN / 2);
        int x0, y0, b0;
        tile_origin(m_tile, &x0, &y0, &b0);
        for (int s = 0; s < slabs; ++s, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * L::STAGE_BYTES;
          const uint32_t b_dst = a_dst + CT_A_BYTES;
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
          const int tap = s / p.cin_slabs;
          const int c0 = (s - tap * p.cin_slabs) * BK;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma2_load_4d(a_dst, &tmA, full_bar(stage), c0, x0 + dx, y0 + dy, b0);
          tma2_load_2d(b_dst, &tmB, full_bar(stage), c0, tap * p.Cout + n0);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = F16IN ? make_idesc_f16(256, BN, 0) : make_idesc_tf32(256, BN);
      uint32_t it = 0;
      int tile_iter = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters, ++tile_iter) {
        const int acc = tile_iter & 1;
        const uint32_t acc_phase = (tile_iter >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int s = 0; s < slabs; ++s, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_src = smem_base + stage * L::STAGE_BYTES;
          const uint32_t b_src = a_src + CT_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 8 tf32 (or 16 halves) = 32 bytes per MMA step
            const uint64_t a_desc = make_smem_desc_sw128(a_src + k * 32, 0, 1024);
            const uint64_t b_desc = make_smem_desc_sw128(b_src + k * 32, 0, 1024);
            if (F16IN) tc_mma_f16_2cta(d, a_desc, b_desc, idesc, (s > 0 || k > 0) ? 1u : 0u);
            else tc_mma_tf32_2cta(d, a_desc, b_desc, idesc, (s > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_2cta(empty_bar(stage), 0x3);
        }
        tc_commit_2cta(tfull_bar(acc), 0x3);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;            // 0: even 32-column chunks, 1: odd chunks
    constexpr int NCHUNK = BN / 32;
    const uint32_t my_stage = smem_base + L::STORE_OFFSET + (half * 4 + quarter) * 4096;
    uint8_t* my_stage_gen = smem_raw + (my_stage - smem_u32(smem_raw));
    // this warp's 32 accumulator rows = a rectangular sub-box of the (bw, bh, bb) pixel box
    const int qx = (quarter * 32) % p.bw;
    const int qy = ((quarter * 32) / p.bw) % p.bh;
    const int qb = (quarter * 32) / (p.bw * p.bh);
    int tile_iter = 0;
    for (int t = cluster_id; t < total_tiles; t += n_clusters, ++tile_iter) {
      const int acc = tile_iter & 1;
      const uint32_t acc_phase = (tile_iter >> 1) & 1;
      const int n_tile = t % p.n_tiles;
      const int m_tile = 2 * (t / p.n_tiles) + static_cast<int>(rank);
      int x0, y0, b0;
      tile_origin(m_tile, &x0, &y0, &b0);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      // this thread's accumulator row = pixel (x0 + lx, y0 + ly, b0 + lb) of the tile's box (c fastest, then x, y, b)
      const float* res_row = nullptr;
      if (p.residual != nullptr) {
        const int r_local = quarter * 32 + lane;
        const int px = x0 + r_local % p.bw, py = y0 + (r_local / p.bw) % p.bh, pb = b0 + r_local / (p.bw * p.bh);
        if (px < p.W && py < p.H && pb < p.B)
          res_row = p.residual + ((static_cast<long long>(pb) * p.H + py) * p.W + px) * p.Cout;
      }
#pragma unroll 1
      for (int c = half; c < NCHUNK; c += 2) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        const int ncol = n_tile * BN + c * 32;
        if (p.bias != nullptr) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (ncol + g * 4 < p.Cout) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol + g * 4));
              v[g * 4 + 0] = __float_as_uint(__uint_as_float(v[g * 4 + 0]) + b4.x);
              v[g * 4 + 1] = __float_as_uint(__uint_as_float(v[g * 4 + 1]) + b4.y);
              v[g * 4 + 2] = __float_as_uint(__uint_as_float(v[g * 4 + 2]) + b4.z);
              v[g * 4 + 3] = __float_as_uint(__uint_as_float(v[g * 4 + 3]) + b4.w);
            }
          }
        }
        if (res_row != nullptr) {   // after the bias rounding, like torch's `x + conv(h)`
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (ncol + g * 4 < p.Cout) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(res_row + ncol + g * 4));
              v[g * 4 + 0] = __float_as_uint(__uint_as_float(v[g * 4 + 0]) + r4.x);
              v[g * 4 + 1] = __float_as_uint(__uint_as_float(v[g * 4 + 1]) + r4.y);
              v[g * 4 + 2] = __float_as_uint(__uint_as_float(v[g * 4 + 2]) + r4.z);
              v[g * 4 + 3] = __float_as_uint(__uint_as_float(v[g * 4 + 3]) + r4.w);
            }
          }
        }
        if (lane == 0) tma_store_wait_read<0>();   // the previous chunk's store has drained this warp's buffer
        __syncwarp();
        uint8_t* dst = my_stage_gen + lane * 128;  // row = 32 floats = 128 B, 16-byte chunk q at slot q ^ (row & 7)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(dst + ((q ^ (lane & 7)) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmOut, my_stage, ncol, x0 + qx, y0 + qy, b0 + qb);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

template <int BN, int STAGES, bool F16IN>
static int launch_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const ConvF32Params& p,
                       int m_pairs, cudaStream_t stream) {
  using L = SmemT<BN, STAGES>;
  auto kern = conv_tf32_kernel<BN, STAGES, F16IN>;
  static bool configured = false;
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = m_pairs * p.n_tiles;
  const int clusters = tiles < kSMs / 2 ? tiles : kSMs / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VTON_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, p, m_pairs));
  count_launch();
  return kOk;
}

// x: [B,H,W,Cin] fp32 NHWC (dense), w: [9][Cout][Cin] fp32 (tap-major), bias: [Cout] fp32 or null, residual: [B,H,W,Cout]
// fp32 NHWC or null (added after the bias), out: [B,H,W,Cout]
// in_fp16: x and w are fp16 (same layouts), Cin a multiple of 64; everything else stays fp32.
int conv3x3_f32_impl(const void* x, int B, int H, int W, int Cin, const void* w, int Cout, const void* bias,
                     const void* residual, void* out, int in_fp16, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && H > 0 && W > 0, "conv3x3_f32: empty input");
  VTON_CHECK_ARG(Cin % 32 == 0 && Cout % 32 == 0 && Cout >= 64, "conv3x3_f32: Cin=%d / Cout=%d must be multiples of 32 (Cout >= 64)", Cin, Cout);
  VTON_CHECK_ARG(!in_fp16 || Cin % 64 == 0, "conv3x3_f32: fp16 operands need Cin=%d to be a multiple of 64", Cin);
  const int esz = in_fp16 ? 2 : 4;          // operand element size
  const int bk = in_fp16 ? 64 : 32;         // channels per 128-byte slab row
  VTON_CHECK_ARG(x && w && out, "conv3x3_f32: null pointer");
  int bw = 1;
  while (bw < 128 && W % (bw * 2) == 0) bw *= 2;
  int bh = 1;
  while (bw * bh < 128 && bh * 2 <= H) bh *= 2;
  const int bb = 128 / (bw * bh);
  VTON_CHECK_ARG(bw * bh * bb == 128 && bw >= 8, "conv3x3_f32: W=%d H=%d cannot be tiled into 128-pixel boxes", W, H);
  const int bn = (Cout % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmB, tmOut;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(Cin) * esz, static_cast<uint64_t>(W) * Cin * esz,
                           static_cast<uint64_t>(H) * W * Cin * esz};
    uint32_t box[4] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bb)};
    if (int e = in_fp16 ? encode_tmap_f16(&tmA, x, 4, dims, strides, box) : encode_tmap_f32(&tmA, x, 4, dims, strides, box)) return e;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(9) * Cout};
    uint64_t strides[1] = {static_cast<uint64_t>(Cin) * esz};
    uint32_t box[2] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(bn / 2)};
    if (int e = in_fp16 ? encode_tmap_f16(&tmB, w, 2, dims, strides, box) : encode_tmap_f32(&tmB, w, 2, dims, strides, box)) return e;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(Cout) * 4, static_cast<uint64_t>(W) * Cout * 4,
                           static_cast<uint64_t>(H) * W * Cout * 4};
    const uint32_t sbw = bw < 32 ? bw : 32;
    const uint32_t sbh = static_cast<uint32_t>(bh) < 32 / sbw ? bh : 32 / sbw;
    const uint32_t sbb = 32 / (sbw * sbh);
    uint32_t box[4] = {32, sbw, sbh, sbb};
    if (int e = encode_tmap_f32(&tmOut, out, 4, dims, strides, box)) return e;
  }
  ConvF32Params p{};
  p.bias = static_cast<const float*>(bias);
  p.residual = static_cast<const float*>(residual);
  p.B = B;
  p.H = H;
  p.W = W;
  p.Cout = Cout;
  p.bw = bw;
  p.bh = bh;
  p.bb = bb;
  p.tiles_x = W / bw;
  p.tiles_y = cdiv(H, bh);
  p.cin_slabs = Cin / bk;
  p.n_tiles = cdiv(Cout, bn);
  const int m_tiles = p.tiles_x * p.tiles_y * cdiv(B, bb);
  const int m_pairs = cdiv(m_tiles, 2);
  if (in_fp16) {
    if (bn == 256) return launch_tf32<256, 5, true>(tmA, tmB, tmOut, p, m_pairs, stream);
    return launch_tf32<128, 6, true>(tmA, tmB, tmOut, p, m_pairs, stream);
  }
  if (bn == 256) return launch_tf32<256, 5, false>(tmA, tmB, tmOut, p, m_pairs, stream);
  return launch_tf32<128, 6, false>(tmA, tmB, tmOut, p, m_pairs, stream);
}

}  // namespace vton
