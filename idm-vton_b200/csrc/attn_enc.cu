// Encoder self-attention on tcgen05 for the CLIP towers around the denoising loop: the ViT-H image
// encoder (16 heads of 80, 257 tokens; src/tryon_pipeline.py:460-482 calls it twice per request) and the two text
// encoders (heads of 64, 77 tokens, causal mask; src/tryon_pipeline.py:511-743). The d = 64 kernels of the UNet path
// (attn.cu / attn6.cu) cannot take a head of 80, hence this one.
//
// One CTA = one (sample, head, 128-query tile), the machine of attn.cu: warp 0 TMA producer, warp 1 tcgen05.mma issuer,
// warps 2..5 softmax (thread = query row); S = Q K^T in tensor memory, P (fp16) through 128B-swizzled shared memory, P V in
// a second tensor-memory tile folded into register accumulators with the online-softmax rescale.
//
// Head dimension D = 16..96 in steps of 16: the head's columns are fetched as R = ceil(D/64) swizzled regions of 64
// columns straight out of the fused [tokens, 3*H*D] projection buffer (the second box starts at column h*D + 64 and runs
// into the next head's columns — those are never touched: Q K^T issues exactly D/16 k-steps, and the extra output columns
// of P V are never read back), so no padded copy of Q/K/V exists.
#include "common.cuh"
#include "host.h"

namespace vton {

struct EncAttnParams {
  __half* out;
  int ld_out;
  int B, H, N, D;
  int ksteps;  // D / 16
  int causal;  // key j visible to query i iff j <= i
  float scale_log2;
};

constexpr int EA_REGION = 128 * 128;  // 128 rows x 64 halves
constexpr int EA_STAGES = 2;

template <int R>
struct EncAttnSmem {
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + R * EA_REGION;
  static constexpr int OFF_V = OFF_K + EA_STAGES * R * EA_REGION;
  static constexpr int OFF_P = OFF_V + EA_STAGES * R * EA_REGION;
  static constexpr int OFF_BAR = OFF_P + 2 * EA_REGION;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

__device__ __forceinline__ float ea_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// R: 64-column regions per head; NCH: 32-column chunks of the output kept in registers (ceil(D/32))
template <int R, int NCH>
__global__ void __launch_bounds__(192, 1)
enc_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const EncAttnParams p) {
  using SM = EncAttnSmem<R>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + SM::OFF_BAR;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + EA_STAGES + s); };
  const uint32_t s_full = bar_base + 8u * (1 + 2 * EA_STAGES);
  const uint32_t p_full = s_full + 8;
  const uint32_t o_full = s_full + 16;
  const uint32_t tmem_slot = s_full + 24;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + SM::OFF_BAR + 8 * (4 + 2 * EA_STAGES));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;

  // causal: the keys past the last query of this tile are never visible to it
  const int n_keys = p.causal ? min(p.N, q_tile * 128 + 128) : p.N;
  const int total = (n_keys + 127) >> 7;
  constexpr uint32_t kTmemCols = 256;  // S: 128 columns, P V: R * 64

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < EA_STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_pv = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, R * EA_REGION);
#pragma unroll
      for (int r = 0; r < R; ++r)
        tma_load_3d(smem_base + SM::OFF_Q + r * EA_REGION, &tmQ, q_full, h * p.D + r * 64, q_tile * 128, b);
      for (int j = 0; j < total; ++j) {
        const int stage = j % EA_STAGES;
        const uint32_t phase = (j / EA_STAGES) & 1;
        mbar_wait(kv_empty(stage), phase ^ 1);
        mbar_expect_tx(kv_full(stage), 2 * R * EA_REGION);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          tma_load_3d(smem_base + SM::OFF_K + (stage * R + r) * EA_REGION, &tmK, kv_full(stage), h * p.D + r * 64, j * 128, b);
          tma_load_3d(smem_base + SM::OFF_V + (stage * R + r) * EA_REGION, &tmV, kv_full(stage), h * p.D + r * 64, j * 128, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0);  // S = Q K^T : B (keys x d) is K-major
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 1);   // O = P V   : B (d x keys) is MN-major, one 64-wide region
      mbar_wait(q_full, 0);
      for (int j = 0; j < total; ++j) {
        const int stage = j % EA_STAGES;
        const uint32_t phase = (j / EA_STAGES) & 1;
        mbar_wait(kv_full(stage), phase);
        tc_fence_after();
        const uint32_t qsrc = smem_base + SM::OFF_Q;
        const uint32_t ksrc = smem_base + SM::OFF_K + stage * R * EA_REGION;
        const uint32_t vsrc = smem_base + SM::OFF_V + stage * R * EA_REGION;
        for (int k = 0; k < p.ksteps; ++k) {
          const uint32_t off = (k >> 2) * EA_REGION + (k & 3) * 32;
          tc_mma_f16(tmem_s, make_smem_desc_sw128(qsrc + off, 0, 1024), make_smem_desc_sw128(ksrc + off, 0, 1024), idesc_s,
                     k > 0 ? 1u : 0u);
        }
        tc_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t psrc = smem_base + SM::OFF_P;
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t a_desc = make_smem_desc_sw128(psrc + (k >> 2) * EA_REGION + (k & 3) * 32, 0, 1024);
            const uint64_t b_desc = make_smem_desc_sw128(vsrc + r * EA_REGION + k * 2048, EA_REGION, 1024);
            tc_mma_f16(tmem_pv + r * 64, a_desc, b_desc, idesc_o, k > 0 ? 1u : 0u);
          }
        }
        tc_commit(kv_empty(stage));
        tc_commit(o_full);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int q_idx = q_tile * 128 + row;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o[NCH * 32];
#pragma unroll
    for (int i = 0; i < NCH * 32; ++i) o[i] = 0.f;
    const float sl2 = p.scale_log2;
    uint8_t* p_row = smem_gen + SM::OFF_P + row * 128;
    const int rx = row & 7;
    // number of keys visible to this query row (>= 1, so tile 0 always leaves a finite running maximum)
    const int row_keys = p.causal ? min(p.N, q_idx + 1) : p.N;

    for (int j = 0; j < total; ++j) {
      const int kv_valid = min(128, row_keys - j * 128);   // <= 0: no visible key in this tile (causal)
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_s + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (c * 32 + i < kv_valid) ? __uint_as_float(r[i]) : -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      // a tile with no visible key for this row (causal, later tiles): keep the running maximum (finite after tile 0)
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ea_exp2((m_run - m_new) * sl2);
      const float m_sc = m_new * sl2;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_s + lane_addr + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = c * 32 + 2 * i;
          const float p0 = (col < kv_valid) ? ea_exp2(__uint_as_float(r[2 * i]) * sl2 - m_sc) : 0.f;
          const float p1 = (col + 1 < kv_valid) ? ea_exp2(__uint_as_float(r[2 * i + 1]) * sl2 - m_sc) : 0.f;
          sum += p0 + p1;
          pk[i] = pack_h2(p0, p1);
        }
        // 32 key columns = 4 chunks of 16 B inside region (c >> 1), 16B-chunk index ((c & 1) * 4 + q)
        uint8_t* region = p_row + (c >> 1) * EA_REGION;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;
          uint4 v = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          *reinterpret_cast<uint4*>(region + ((chunk ^ rx) << 4)) = v;
        }
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1, 128);
      if (threadIdx.x == 64) mbar_arrive(p_full);
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_pv + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(r[i]);
      }
      tc_fence_before();
    }
    if (q_idx < p.N) {
      const float inv = 1.f / l_run;
      __half* dst = p.out + (static_cast<long long>(b) * p.N + q_idx) * p.ld_out + h * p.D;
#pragma unroll
      for (int g = 0; g < NCH * 4; ++g) {
        if (g * 8 < p.D) {
          uint4 ov;
          ov.x = pack_h2(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv);
          ov.y = pack_h2(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv);
          ov.z = pack_h2(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv);
          ov.w = pack_h2(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv);
          *reinterpret_cast<uint4*>(dst + g * 8) = ov;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

static int encode_head_tokens(CUtensorMap* tm, const void* base, long long ld, int cols, int n, int batch) {
  uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(n), static_cast<uint64_t>(batch)};
  uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(n) * ld * 2};
  uint32_t box[3] = {64, 128, 1};
  return encode_tmap_f16(tm, base, 3, dims, strides, box);
}

template <int R, int NCH>
static int launch_enc(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const EncAttnParams& p,
                      cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(enc_attn_kernel<R, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, EncAttnSmem<R>::TOTAL));
    configured = true;
  }
  dim3 grid((p.N + 127) / 128, p.H, p.B);
  enc_attn_kernel<R, NCH><<<grid, 192, EncAttnSmem<R>::TOTAL, stream>>>(tmQ, tmK, tmV, p);
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

// q / k / v: [B, N, >= H*D] views (row strides ldq / ldkv, e.g. the three column blocks of a fused QKV buffer);
// out: [B, N, H*D] (row stride ldo)
int enc_attn_impl(const void* q, long long ldq, const void* k, const void* v, long long ldkv, void* out, long long ldo,
                  int B, int H, int N, int D, float scale, int causal, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && H > 0 && N > 0, "encoder_attention: bad sizes B=%d H=%d N=%d", B, H, N);
  VTON_CHECK_ARG(D >= 16 && D <= 96 && D % 16 == 0, "encoder_attention: head dim %d unsupported (16..96, multiple of 16)", D);
  VTON_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0, "encoder_attention: row strides must be multiples of 8");
  VTON_CHECK_ARG(B <= 65535 && H <= 65535, "encoder_attention: grid too large");
  CUtensorMap tmQ, tmK, tmV;
  if (int e = encode_head_tokens(&tmQ, q, ldq, H * D, N, B)) return e;
  if (int e = encode_head_tokens(&tmK, k, ldkv, H * D, N, B)) return e;
  if (int e = encode_head_tokens(&tmV, v, ldkv, H * D, N, B)) return e;
  EncAttnParams p{};
  p.out = static_cast<__half*>(out);
  p.ld_out = static_cast<int>(ldo);
  p.B = B;
  p.H = H;
  p.N = N;
  p.D = D;
  p.ksteps = D / 16;
  p.causal = causal ? 1 : 0;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (D <= 64) return launch_enc<1, 2>(tmQ, tmK, tmV, p, stream);
  return launch_enc<2, 3>(tmQ, tmK, tmV, p, stream);
}

}  // namespace vton
