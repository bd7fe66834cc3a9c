// Decoupled ping-pong flash attention on tcgen05 (sm_100a), head_dim 64, two-segment K/V stream — the throughput
// kernel behind b200vton_attention for Nq >= 256 (same semantics as attn.cu). Successor of attn5.cu; ncu on attn5
// showed the softmax warps parked 30% of their time on "S tile ready": S_g(j+1) could only be issued after P.V_g(j)
// because P overwrote the S columns, so softmax -> P.V -> S -> softmax was one serial chain per query tile. Here
//   * P gets its own TMEM columns, so S_g(j+1) = Q_g K(j+1)^T is issued as soon as the softmax warps have pulled
//     S_g(j) into registers (s_read barrier) and runs underneath the exponentials of tile j;
//   * P never crosses shared memory (tcgen05.st, then the TS-form MMA takes A = P from TMEM);
//   * fp32 softmax (ex2.approx.f32, fp32 row sums on the FMA pipe): measured faster than the packed-half variant,
//     the B200 SFU issues ex2.f16x2 as two ops anyway;
//   * 5-stage K/V ring (K(j+1) is consumed one softmax period before V(j));
//   * CTAs of samples with the garment segment (twice the K/V tiles) are scheduled first (blockIdx.z reversed), which
//     removes the long tail the short-first order had (17% of the launch at the config-2 shapes).
//
// One CTA owns 256 query rows of one (sample, head) as two 128-row tiles. Each K/V tile is loaded ONCE for both.
//   warp 0      TMA producer (Q0, Q1, then K/V tiles through the ring)
//   warp 1      tcgen05.mma issuer:  S0(0) S1(0) | [PV0(j-1) S0(j+1)] [PV1(j-1) S1(j+1)] | ...
//   warps 2..5  softmax group 0 (thread = query row of tile 0),  warps 6..9  softmax group 1 (tile 1)
// O accumulates IN TMEM across K/V tiles; the online-softmax rescale of O is lazy: the running exponent reference
// m_used only moves (and O / l are rescaled through tcgen05.ld/st) when a row maximum grows by more than 2^8, which
// keeps P <= 256 in fp16 and makes the correction rare after the first tile. O / l is exact in exact arithmetic for
// any reference m_used.
// TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)  (one CTA per SM).
#include "common.cuh"
#include "host.h"

namespace vton {

struct Attn6Params {
  __half* out;
  int ld_out;
  int B, H, Nq, N0, N1;
  int kv1_off, kv1_count;
  const int* kv1_base;
  float scale_log2;
  int accumulate;
};

constexpr int A6_TILE = 128 * 128;          // bytes of one 128 x 64 fp16 tile
// Shared memory of a CTA with G query tiles: Q tiles | K ring | V ring | barriers. K(j+1) is consumed one softmax
// period before V(j), so the rings are separate: a K slot is released as soon as its S MMAs retire.
template <int G> struct A6Cfg {
  static constexpr int KS = (G == 2) ? 4 : 3;
  static constexpr int VS = (G == 2) ? 3 : 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = G * A6_TILE;
  static constexpr int OFF_V = OFF_K + KS * A6_TILE;
  static constexpr int OFF_BAR = OFF_V + VS * A6_TILE;
  static constexpr int SMEM_TOTAL = OFF_BAR + 256 + 1024;   // G=2: 148.3 KB (one CTA/SM), G=1: 97.3 KB (two CTAs/SM)
  static constexpr int THREADS = 64 + 128 * G;
  static constexpr uint32_t TMEM_COLS = 256 * G;
  static constexpr uint32_t TM_O = 128 * G, TM_P = 128 * G + 64 * G;
};
constexpr float kLazyThreshold6 = 8.0f;     // log2 domain

__device__ __forceinline__ float ex2k_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA pipe (no SFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax polynomial for 2^f
// (max relative error 7.5e-5, an order below the fp16 rounding of P that follows), n added into the exponent field.
// POLY of every 4 exponentials can be evaluated here instead of the SFU (the trade FlashAttention-4 describes); on B200
// at d = 64 it measured slower (695 / 683 / 573 TFLOP/s for POLY = 0 / 1 / 2), so POLY = 0 is the default.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float xr = x + 12582912.f;   // 1.5 * 2^23: the integer nearest to x now sits in the low mantissa bits
  const float n = xr - 12582912.f;
  const float f = x - n;
  float p = fmaf(f, 0.05517132f, 0.24261054f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992812f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int G, int POLY>
__global__ void __launch_bounds__(A6Cfg<G>::THREADS, (G == 1) ? 2 : 1)
attn6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
             const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
             const __grid_constant__ CUtensorMap tmV1, const Attn6Params p) {
  using Cfg = A6Cfg<G>;
  constexpr int KS = Cfg::KS, VS = Cfg::VS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::OFF_BAR;
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (1 + KS + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (1 + 2 * KS + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (1 + 2 * KS + VS + s); };
  constexpr int kGroupBars = 1 + 2 * KS + 2 * VS;
  auto s_full = [&](int g) { return bar_base + 8u * (kGroupBars + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (kGroupBars + 2 + g); };
  auto o_full = [&](int g) { return bar_base + 8u * (kGroupBars + 4 + g); };
  auto s_read = [&](int g) { return bar_base + 8u * (kGroupBars + 6 + g); };
  const uint32_t tmem_slot = bar_base + 8u * (kGroupBars + 8);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::OFF_BAR + 8 * (kGroupBars + 8));
  static_assert(8 * (kGroupBars + 9) <= 256, "barrier block overflows");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_unit = blockIdx.x;                           // 128*G query rows
  const int h = blockIdx.y;
  const int b = p.B - 1 - static_cast<int>(blockIdx.z);   // two-segment samples (longer K/V stream) first

  const int tiles0 = (p.N0 + 127) >> 7;
  int idx1 = -1;
  if (p.N1 > 0) {
    idx1 = b - p.kv1_off;
    if (idx1 >= 0) idx1 = idx1 % p.kv1_count + (p.kv1_base ? *p.kv1_base : 0);
  }
  const bool zero_kv = (p.N1 > 0) && (idx1 < 0);
  const int tiles1 = (p.N1 > 0 && idx1 >= 0) ? ((p.N1 + 127) >> 7) : 0;
  const int total = tiles0 + tiles1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK0);
    tma_prefetch_desc(&tmV0);
    if (tiles1) {
      tma_prefetch_desc(&tmK1);
      tma_prefetch_desc(&tmV1);
    }
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
    }
    for (int s = 0; s < VS; ++s) {
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    for (int g = 0; g < G; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 4);   // one arrive per softmax warp of the group
      mbar_init(o_full(g), 1);
      mbar_init(s_read(g), 4);   // S_g pulled into registers by all four warps of the group
    }
    fence_barrier_init();
  }
  // Programmatic dependent launch: wait for the previous kernel BEFORE allocating tensor memory. A dependent CTA that
  // grabbed TMEM first and then waited could starve a primary CTA on the same SM that had signalled its dependents but
  // not yet allocated its own columns (library kernels signal at their very start): neither would ever proceed.
  pdl_wait();
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();   // this CTA holds everything it needs: the next kernel may start its own set-up

  if (warp == 0) {
    // ===================== TMA producer: Q, then K(0) | K(1) V(0) | K(2) V(1) | ... (consumption order) =====
    if (lane == 0) {
      mbar_expect_tx(q_full, G * A6_TILE);
#pragma unroll
      for (int g = 0; g < G; ++g)
        tma_load_3d(smem_base + Cfg::OFF_Q + g * A6_TILE, &tmQ, q_full, h * 64, (q_unit * G + g) * 128, b);
      auto load_tile = [&](uint32_t dst, uint32_t bar, const CUtensorMap* m0, const CUtensorMap* m1, int n) {
        mbar_expect_tx(bar, A6_TILE);
        if (n < tiles0) tma_load_3d(dst, m0, bar, h * 64, n * 128, b);
        else tma_load_3d(dst, m1, bar, h * 64, (n - tiles0) * 128, idx1);
      };
      for (int n = 0; n <= total; ++n) {
        if (n < total) {
          const int ks = n % KS;
          mbar_wait(k_empty(ks), ((n / KS) & 1) ^ 1);
          load_tile(smem_base + Cfg::OFF_K + ks * A6_TILE, k_full(ks), &tmK0, &tmK1, n);
        }
        if (n >= 1) {
          const int vs = (n - 1) % VS;
          mbar_wait(v_empty(vs), (((n - 1) / VS) & 1) ^ 1);
          load_tile(smem_base + Cfg::OFF_V + vs * A6_TILE, v_full(vs), &tmV0, &tmV1, n - 1);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 1);
      auto issue_s = [&](int g, int ks) {
        const uint32_t qsrc = smem_base + Cfg::OFF_Q + g * A6_TILE;
        const uint32_t ksrc = smem_base + Cfg::OFF_K + ks * A6_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tmem_base + g * 128, make_smem_desc_sw128(qsrc + k * 32, 0, 1024),
                     make_smem_desc_sw128(ksrc + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        tc_commit(s_full(g));
        if (g == G - 1) tc_commit(k_empty(ks));   // this K tile has been multiplied with every Q tile of the CTA
      };
      mbar_wait(q_full, 0);
      mbar_wait(k_full(0), 0);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < G; ++g) issue_s(g, 0);
      // iteration j: for each group, P.V of tile j-1 (its P just became ready) and then S of tile j+1 (its S(j) was
      // just read) — the two events are adjacent in the group's timeline, and the groups run half a period apart.
      for (int j = 0; j <= total; ++j) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (j >= 1) {
            const int vs = (j - 1) % VS;
            const uint32_t vsrc = smem_base + Cfg::OFF_V + vs * A6_TILE;
            if (g == 0) mbar_wait(v_full(vs), ((j - 1) / VS) & 1);
            mbar_wait(p_full(g), (j - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              // A = P from TMEM (16 keys = 8 packed 32-bit columns per K step); B = V tile, MN-major, 16 key rows/step
              tc_mma_f16_ts(tmem_base + Cfg::TM_O + g * 64, tmem_base + Cfg::TM_P + g * 64 + k * 8,
                            make_smem_desc_sw128(vsrc + k * 2048, A6_TILE, 1024), idesc_o,
                            (j > 1 || k > 0) ? 1u : 0u);
            }
            tc_commit(o_full(g));
            if (g == G - 1) tc_commit(v_empty(vs));
          }
          if (j + 1 < total) {
            const int ks = (j + 1) % KS;
            if (g == 0) mbar_wait(k_full(ks), ((j + 1) / KS) & 1);
            mbar_wait(s_read(g), j & 1);
            tc_fence_after();
            issue_s(g, ks);
          }
        }
      }
    }
  } else {
    // ===================== softmax (2 groups x 4 warps, thread = query row) =====================
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int q_idx = (q_unit * G + g) * 128 + row;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_addr;
    const uint32_t tO = tmem_base + Cfg::TM_O + g * 64 + lane_addr;
    const uint32_t tP = tmem_base + Cfg::TM_P + g * 64 + lane_addr;
    const float sl2 = p.scale_log2;
    float m_used = 0.f;
    float l_run = 0.f;

    for (int j = 0; j < total; ++j) {
      const int kv_valid = (j < tiles0) ? min(128, p.N0 - j * 128) : min(128, p.N1 - (j - tiles0) * 128);
      mbar_wait(s_full(g), j & 1);
      tc_fence_after();
      uint32_t s[4][32];
      tmem_ld_32x32(tS + 0, s[0]);
      tmem_ld_32x32(tS + 32, s[1]);
      tmem_ld_32x32(tS + 64, s[2]);
      tmem_ld_32x32(tS + 96, s[3]);
      tmem_ld_wait();
      // S_g(j) is in registers: the tensor core may overwrite the S columns with S_g(j+1) while we exponentiate
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_read(g));
      if (kv_valid < 128) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= kv_valid) s[c][i] = 0xff800000u;   // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(s[0][i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[1][i]));
        mx2 = fmaxf(mx2, __uint_as_float(s[2][i]));
        mx3 = fmaxf(mx3, __uint_as_float(s[3][i]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      float alpha = 1.f;
      bool rescale = false;
      if (j == 0) {
        m_used = mx;
      } else if ((mx - m_used) * sl2 > kLazyThreshold6) {
        alpha = ex2k_((m_used - mx) * sl2);
        m_used = mx;
        rescale = true;
      }
      const float msc = m_used * sl2;
      uint32_t pk[2][32];    // P row, packed pairs: 32-bit column c*16+i holds keys (c*32 + 2i, c*32 + 2i + 1)
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float e0 = ex2k_(__uint_as_float(s[c][2 * i]) * sl2 - msc);
          const float e1 = ex2k_(__uint_as_float(s[c][2 * i + 1]) * sl2 - msc);
          const float x2 = __uint_as_float(s[c][2 * i + 2]) * sl2 - msc;
          const float x3 = __uint_as_float(s[c][2 * i + 3]) * sl2 - msc;
          const float e2 = POLY >= 1 ? ex2_poly(x2) : ex2k_(x2);
          const float e3 = POLY >= 2 ? ex2_poly(x3) : ex2k_(x3);
          l0 += e0;
          l1 += e1;
          l2 += e2;
          l3 += e3;
          pk[c >> 1][(c & 1) * 16 + i] = pack_h2(e0, e1);
          pk[c >> 1][(c & 1) * 16 + i + 1] = pack_h2(e2, e3);
        }
      }
      l_run = l_run * alpha + ((l0 + l1) + (l2 + l3));
      if (j > 0) {
        // P.V_g(j-1) must have retired before O is rescaled or its P operand is overwritten (normally long done:
        // it was issued one softmax period ago)
        mbar_wait(o_full(g), (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(tO + c * 32, o);
          }
        }
      }
      tmem_st_32x32(tP, pk[0]);
      tmem_st_32x32(tP + 32, pk[1]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(g));
    }
    // ---- finalize
    mbar_wait(o_full(g), (total - 1) & 1);
    tc_fence_after();
    float o[64];
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tO, r0);
      tmem_ld_32x32(tO + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = __uint_as_float(r0[i]);
        o[32 + i] = __uint_as_float(r1[i]);
      }
    }
    if (zero_kv) {
      // N1 all-zero key/value tokens: score 0 each, value 0
      const float m_new = fmaxf(m_used, 0.f);
      const float beta = ex2k_((m_used - m_new) * sl2);
      l_run = l_run * beta + static_cast<float>(p.N1) * ex2k_(-m_new * sl2);
#pragma unroll
      for (int i = 0; i < 64; ++i) o[i] *= beta;
    }
    if (q_idx < p.Nq) {
      const float inv = 1.f / l_run;
      __half* dst = p.out + (static_cast<long long>(b) * p.Nq + q_idx) * p.ld_out + h * 64;
#pragma unroll
      for (int gq = 0; gq < 8; ++gq) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o[gq * 8 + i] * inv;
        if (p.accumulate) {
          const uint4 old = *reinterpret_cast<const uint4*>(dst + gq * 8);
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
        *reinterpret_cast<uint4*>(dst + gq * 8) = ov;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

int attn6_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK0, const CUtensorMap& tmV0, const CUtensorMap& tmK1,
                 const CUtensorMap& tmV1, __half* out, int ld_out, int B, int H, int Nq, int N0, int N1, int kv1_off,
                 int kv1_count, const int* kv1_base, float scale_log2, int accumulate, int q_tiles, int poly,
                 cudaStream_t stream) {
  Attn6Params p{};
  p.out = out;
  p.ld_out = ld_out;
  p.B = B;
  p.H = H;
  p.Nq = Nq;
  p.N0 = N0;
  p.N1 = N1;
  p.kv1_off = kv1_off;
  p.kv1_count = kv1_count;
  p.kv1_base = kv1_base;
  p.scale_log2 = scale_log2;
  p.accumulate = accumulate;
  if (q_tiles == 0) q_tiles = 1;
  if (poly < 0 || poly > 2) poly = 0;
  static bool configured = false;
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(attn6_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6Cfg<1>::SMEM_TOTAL));
    VTON_CUDA(cudaFuncSetAttribute(attn6_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6Cfg<1>::SMEM_TOTAL));
    VTON_CUDA(cudaFuncSetAttribute(attn6_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6Cfg<1>::SMEM_TOTAL));
    VTON_CUDA(cudaFuncSetAttribute(attn6_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6Cfg<2>::SMEM_TOTAL));
    VTON_CUDA(cudaFuncSetAttribute(attn6_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6Cfg<2>::SMEM_TOTAL));
    VTON_CUDA(cudaFuncSetAttribute(attn6_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6Cfg<2>::SMEM_TOTAL));
    configured = true;
  }
  auto go = [&](auto kern, auto cfg_tag) -> int {
    using Cfg = decltype(cfg_tag);
    constexpr int rows = (Cfg::THREADS - 64);   // 128 query rows per softmax group
    dim3 grid((Nq + rows - 1) / rows, H, B);
    VTON_CUDA(launch_kernel(kern, grid, dim3(Cfg::THREADS), Cfg::SMEM_TOTAL, stream, tmQ, tmK0, tmV0, tmK1, tmV1, p));
    return kOk;
  };
  // Two query tiles per CTA share every K/V tile (half the L2 -> smem traffic), but two independent one-tile CTAs per
  // SM hide each other's prologue, epilogue and softmax bubbles: measured faster at every config-2 shape (L1
  // self+garment 644 vs 632 TFLOP/s, L2 422 vs 368, garment batch-16 728 vs 644 / 513 vs 412), so it is the default.
  int rc;
  if (q_tiles == 1) {
    rc = poly == 0 ? go(attn6_kernel<1, 0>, A6Cfg<1>{}) : poly == 1 ? go(attn6_kernel<1, 1>, A6Cfg<1>{}) : go(attn6_kernel<1, 2>, A6Cfg<1>{});
  } else {
    rc = poly == 0 ? go(attn6_kernel<2, 0>, A6Cfg<2>{}) : poly == 1 ? go(attn6_kernel<2, 1>, A6Cfg<2>{}) : go(attn6_kernel<2, 2>, A6Cfg<2>{});
  }
  if (rc) return rc;
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vton
