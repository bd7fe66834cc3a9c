// Flash-style attention on tcgen05 for head_dim 64 with a TWO-SEGMENT key/value stream (sm_100a).
//
// Replaces, on the try-on UNet's hot path:
//   * attn1 of src/attentionhacked_tryon.py:334-348 + ip_adapter/attention_processor.py:238-262 — self-attention whose
//     keys/values are [self tokens ; garment tokens]. Segment 0 = this sample's K/V, segment 1 = the cached garment K/V
//     of sample (b - kv1_off) % kv1_count; the torch.cat never happens. Query rows are the N self tokens only
//     (the reference computes and discards the Ng garment query rows).
//   * CFG-uncond samples (b < kv1_off) see ZERO garment features (src/tryon_pipeline.py:1796): K=V=0, so each of the N1
//     tokens adds exp(0 - m) to the softmax denominator and nothing to the numerator. Closed form, no KV traffic.
//   * attn2 (ip_adapter/attention_processor.py:1943-1995): called twice (text tokens, then IP tokens with
//     accumulate=1) — two independent softmaxes whose fp16 outputs are summed in fp16.
//
// One CTA = one (sample, head, 128-query tile). Warp 0: TMA producer, warp 1: tcgen05.mma issuer, warps 2..5: softmax
// (thread = query row). S = Q K^T lands in TMEM (fp32), softmax reads it with tcgen05.ld, writes P (fp16) into
// 128B-swizzled smem, P V goes through the tensor core into a second TMEM tile that is folded into register
// accumulators with the online-softmax rescale. Two CTAs are co-resident per SM so one CTA's softmax overlaps the other's MMAs.
#include "common.cuh"
#include "host.h"

namespace vton {

struct AttnParams {
  __half* out;
  int ld_out;
  int B, H, Nq, N0, N1;
  int kv1_off;    // segment-1 sample index = (b - kv1_off) % kv1_count; negative => zero K/V closed form
  int kv1_count;  // segment-1 sample index is taken modulo this count
  const int* kv1_base;  // optional device scalar added to the segment-1 sample index (hoisted per-step K/V)
  float scale_log2;
  int accumulate;
};

constexpr int AT_Q_BYTES = 128 * 128;      // 128 rows x 64 halves
constexpr int AT_KV_BYTES = 128 * 128;     // one K or V tile
constexpr int AT_P_BYTES = 2 * 128 * 128;  // 128 x 128 halves as two K-major regions
constexpr int AT_STAGES = 2;
constexpr int AT_OFF_Q = 0;
constexpr int AT_OFF_K = AT_OFF_Q + AT_Q_BYTES;
constexpr int AT_OFF_V = AT_OFF_K + AT_STAGES * AT_KV_BYTES;
constexpr int AT_OFF_P = AT_OFF_V + AT_STAGES * AT_KV_BYTES;
constexpr int AT_OFF_BAR = AT_OFF_P + AT_P_BYTES;
constexpr int AT_SMEM_TOTAL = AT_OFF_BAR + 256 + 1024;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(192, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
            const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
            const __grid_constant__ CUtensorMap tmV1, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + AT_OFF_BAR;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + AT_STAGES + s); };
  const uint32_t s_full = bar_base + 8u * (1 + 2 * AT_STAGES);
  const uint32_t p_full = s_full + 8;
  const uint32_t o_full = s_full + 16;
  const uint32_t tmem_slot = s_full + 24;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + AT_OFF_BAR + 8 * (4 + 2 * AT_STAGES));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;

  const int tiles0 = (p.N0 + 127) >> 7;
  int idx1 = -1;
  if (p.N1 > 0) {
    idx1 = b - p.kv1_off;
    if (idx1 >= 0) idx1 = idx1 % p.kv1_count + (p.kv1_base ? *p.kv1_base : 0);
  }
  const bool zero_kv = (p.N1 > 0) && (idx1 < 0);
  const int tiles1 = (p.N1 > 0 && idx1 >= 0) ? ((p.N1 + 127) >> 7) : 0;
  const int total = tiles0 + tiles1;
  constexpr uint32_t kTmemCols = 256;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK0);
    tma_prefetch_desc(&tmV0);
    if (tiles1) {
      tma_prefetch_desc(&tmK1);
      tma_prefetch_desc(&tmV1);
    }
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_STAGES; ++s) {
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
      mbar_expect_tx(q_full, AT_Q_BYTES);
      tma_load_3d(smem_base + AT_OFF_Q, &tmQ, q_full, h * 64, q_tile * 128, b);
      for (int j = 0; j < total; ++j) {
        const int stage = j % AT_STAGES;
        const uint32_t phase = (j / AT_STAGES) & 1;
        mbar_wait(kv_empty(stage), phase ^ 1);
        mbar_expect_tx(kv_full(stage), 2 * AT_KV_BYTES);
        const uint32_t kdst = smem_base + AT_OFF_K + stage * AT_KV_BYTES;
        const uint32_t vdst = smem_base + AT_OFF_V + stage * AT_KV_BYTES;
        if (j < tiles0) {
          tma_load_3d(kdst, &tmK0, kv_full(stage), h * 64, j * 128, b);
          tma_load_3d(vdst, &tmV0, kv_full(stage), h * 64, j * 128, b);
        } else {
          tma_load_3d(kdst, &tmK1, kv_full(stage), h * 64, (j - tiles0) * 128, idx1);
          tma_load_3d(vdst, &tmV1, kv_full(stage), h * 64, (j - tiles0) * 128, idx1);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0);  // S = Q K^T : B (keys x d) is K-major
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 1);   // O = P V   : B (d x keys) is MN-major
      mbar_wait(q_full, 0);
      for (int j = 0; j < total; ++j) {
        const int stage = j % AT_STAGES;
        const uint32_t phase = (j / AT_STAGES) & 1;
        mbar_wait(kv_full(stage), phase);
        tc_fence_after();
        const uint32_t qsrc = smem_base + AT_OFF_Q;
        const uint32_t ksrc = smem_base + AT_OFF_K + stage * AT_KV_BYTES;
        const uint32_t vsrc = smem_base + AT_OFF_V + stage * AT_KV_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc_mma_f16(tmem_s, make_smem_desc_sw128(qsrc + k * 32, 0, 1024), make_smem_desc_sw128(ksrc + k * 32, 0, 1024),
                     idesc_s, k > 0 ? 1u : 0u);
        }
        tc_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t psrc = smem_base + AT_OFF_P;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t a_desc = make_smem_desc_sw128(psrc + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(vsrc + k * 2048, 16384, 1024);
          tc_mma_f16(tmem_pv, a_desc, b_desc, idesc_o, k > 0 ? 1u : 0u);
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
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    const float sl2 = p.scale_log2;
    uint8_t* p_row = smem_gen + AT_OFF_P + row * 128;
    const int rx = row & 7;

    for (int j = 0; j < total; ++j) {
      const int kv_valid = (j < tiles0) ? min(128, p.N0 - j * 128) : min(128, p.N1 - (j - tiles0) * 128);
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
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2((m_run - m_new) * sl2);
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
          const float p0 = (col < kv_valid) ? fast_exp2(__uint_as_float(r[2 * i]) * sl2 - m_sc) : 0.f;
          const float p1 = (col + 1 < kv_valid) ? fast_exp2(__uint_as_float(r[2 * i + 1]) * sl2 - m_sc) : 0.f;
          sum += p0 + p1;
          pk[i] = pack_h2(p0, p1);
        }
        // 32 key columns = 4 chunks of 16 B inside region (c >> 1), 16B-chunk index ((c & 1) * 4 + q)
        uint8_t* region = p_row + (c >> 1) * 16384;
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
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_pv + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(r[i]);
      }
      tc_fence_before();
    }
    if (zero_kv) {
      // N1 all-zero key/value tokens: score 0 each (App. D.3)
      const float m_new = fmaxf(m_run, 0.f);
      const float alpha = fast_exp2((m_run - m_new) * sl2);
      l_run = l_run * alpha + static_cast<float>(p.N1) * fast_exp2(-m_new * sl2);
#pragma unroll
      for (int i = 0; i < 64; ++i) o[i] *= alpha;
    }
    if (q_idx < p.Nq) {
      const float inv = 1.f / l_run;
      __half* dst = p.out + (static_cast<long long>(b) * p.Nq + q_idx) * p.ld_out + h * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o[g * 8 + i] * inv;
        if (p.accumulate) {
          const uint4 old = *reinterpret_cast<const uint4*>(dst + g * 8);
          const uint32_t ow[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a = unpack_h2(ow[i]);
            v[2 * i] = a.x + round_h(v[2 * i]);
            v[2 * i + 1] = a.y + round_h(v[2 * i + 1]);
          }
        }
        uint4 ov;
        ov.x = pack_h2(v[0], v[1]);
        ov.y = pack_h2(v[2], v[3]);
        ov.z = pack_h2(v[4], v[5]);
        ov.w = pack_h2(v[6], v[7]);
        *reinterpret_cast<uint4*>(dst + g * 8) = ov;
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

static int g_attn_v2 = 1;   // 1: attn6.cu (pipelined, P in tensor memory) for Nq >= 256; 0: this one-tile kernel always
int attn6_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK0, const CUtensorMap& tmV0, const CUtensorMap& tmK1,
                 const CUtensorMap& tmV1, __half* out, int ld_out, int B, int H, int Nq, int N0, int N1, int kv1_off,
                 int kv1_count, const int* kv1_base, float scale_log2, int accumulate, int q_tiles, int poly,
                 cudaStream_t stream);
// attn6: exponentials (of every 4) evaluated by the FMA-pipe polynomial instead of the SFU. Measured on B200
// (profiles/r1_attention_poly_exp.jsonl): 0 -> 695, 1 -> 683, 2 -> 573 TFLOP/s at the 3072-token level: the extra FMA-pipe
// issue slots cost more than the SFU slots they free, so the default is 0.
static int g_attn_poly = 0;
void set_attn_poly(int n) { g_attn_poly = n; }
static int g_attn_qtiles = 0;  // attn6: query tiles per CTA (0 = by K/V length, 1, 2)
void set_attn_qtiles(int n) { g_attn_qtiles = n; }
void set_attn_v2(int on) { g_attn_v2 = on; }

static int encode_tokens(CUtensorMap* tm, const void* base, long long ld, int cols, int n, int batch) {
  uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(n), static_cast<uint64_t>(batch)};
  uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(n) * ld * 2};
  uint32_t box[3] = {64, 128, 1};
  return encode_tmap_f16(tm, base, 3, dims, strides, box);
}

// q: [B, Nq, >=H*64] (row stride ldq); k0/v0: [B, N0, .] (ldkv0); k1/v1: [B1, N1, .] (ldkv1); out: [B, Nq, .] (ldo)
int attn_impl(const void* q, long long ldq, const void* k0, const void* v0, long long ldkv0, const void* k1,
              const void* v1, long long ldkv1, void* out, long long ldo, int B, int H, int Nq, int N0, int N1, int B1,
              int kv1_off, int kv1_mod, const void* kv1_base, float scale, int accumulate, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && H > 0 && Nq > 0 && N0 > 0 && N1 >= 0, "attn: bad sizes B=%d H=%d Nq=%d N0=%d N1=%d", B, H, Nq, N0, N1);
  VTON_CHECK_ARG(ldq % 8 == 0 && ldkv0 % 8 == 0 && ldo % 8 == 0, "attn: row strides must be multiples of 8");
  VTON_CHECK_ARG(B <= 65535 && H <= 65535, "attn: grid too large");
  const bool has1 = N1 > 0 && B1 > 0 && k1 && v1;
  VTON_CHECK_ARG(N1 == 0 || has1 || kv1_off >= B, "attn: segment 1 declared (N1=%d) but no K/V given", N1);
  VTON_CHECK_ARG(!has1 || ldkv1 % 8 == 0, "attn: ldkv1 must be a multiple of 8");
  CUtensorMap tmQ, tmK0, tmV0, tmK1, tmV1;
  if (int e = encode_tokens(&tmQ, q, ldq, H * 64, Nq, B)) return e;
  if (int e = encode_tokens(&tmK0, k0, ldkv0, H * 64, N0, B)) return e;
  if (int e = encode_tokens(&tmV0, v0, ldkv0, H * 64, N0, B)) return e;
  tmK1 = tmK0;
  tmV1 = tmV0;
  if (has1) {
    if (int e = encode_tokens(&tmK1, k1, ldkv1, H * 64, N1, B1)) return e;
    if (int e = encode_tokens(&tmV1, v1, ldkv1, H * 64, N1, B1)) return e;
  }
  if (g_attn_v2 && Nq >= 256) {
    return attn6_launch(tmQ, tmK0, tmV0, tmK1, tmV1, static_cast<__half*>(out), static_cast<int>(ldo), B, H, Nq, N0, N1,
                          has1 ? kv1_off : (N1 > 0 ? B : 0), has1 ? (kv1_mod > 0 ? kv1_mod : B1) : 1,
                          has1 ? static_cast<const int*>(kv1_base) : nullptr, scale * 1.4426950408889634f, accumulate,
                          g_attn_qtiles, g_attn_poly, stream);
  }
  AttnParams p{};
  p.out = static_cast<__half*>(out);
  p.ld_out = static_cast<int>(ldo);
  p.B = B;
  p.H = H;
  p.Nq = Nq;
  p.N0 = N0;
  p.N1 = N1;
  p.kv1_off = has1 ? kv1_off : (N1 > 0 ? B : 0);
  p.kv1_count = has1 ? (kv1_mod > 0 ? kv1_mod : B1) : 1;
  p.kv1_base = has1 ? static_cast<const int*>(kv1_base) : nullptr;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.accumulate = accumulate;
  static bool configured = false;
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_TOTAL));
    configured = true;
  }
  dim3 grid((Nq + 127) / 128, H, B);
  attn_kernel<<<grid, 192, AT_SMEM_TOTAL, stream>>>(tmQ, tmK0, tmV0, tmK1, tmV1, p);
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vton
