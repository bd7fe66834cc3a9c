// P-in-TMEM variant of attn3.cu: the softmax writes P (packed fp16) back into the TMEM columns its S tile occupied
// (tcgen05.st) and the P.V MMA takes its A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc), so P never crosses
// shared memory: per K/V tile pair the smem traffic drops from ~260 KB (Q,K,V,P reads + P writes) to ~130 KB — shared
// memory bandwidth (128 B/clk/SM) was on a par with the SFU and TMEM-read limits of the d = 64 problem — and the
// generic->async proxy fence + swizzled stores leave the softmax critical path.
// Ping-pong flash attention on tcgen05 (sm_100a), head_dim 64, two-segment K/V stream, packed-half softmax — the
// throughput kernel behind b200vton_attention for Nq >= 256 (same semantics as attn.cu; structure of attn2.cu plus):
//   * P = exp2(x) is evaluated two elements per MUFU op (ex2.approx.ftz.f16x2 on x rounded to fp16) and lands directly
//     in the packed fp16 format the P.V MMA consumes — the B200 SFU (16 ex2/clk/SM) is the binding unit at d = 64;
//   * the softmax denominator is NOT summed by the CUDA cores: V is extended by a constant "ones" column (a second
//     MN-major swizzle atom reached through the UMMA descriptor's leading-byte-offset), so the tensor core produces
//     l = sum_j P_ij in fp32 as column 64 of the O accumulator, rescaled together with O;
//   * lazy-rescale threshold 2 (log2 domain): exponents stay <= 2 where fp16 arguments are still accurate.
//
// One CTA owns 256 query rows of one (sample, head) as two 128-row tiles. Each K/V tile is loaded ONCE for both.
//   warp 0      TMA producer (Q0, Q1, then K/V tiles through a 3-stage ring)
//   warp 1      tcgen05.mma issuer:  S0(j) S1(j) | PV0(j) S0(j+1) | PV1(j) S1(j+1) | ...
//   warps 2..5  softmax group 0 (thread = query row of tile 0),  warps 6..9  softmax group 1 (tile 1)
// While group 0 runs the softmax of S0(j+1) the tensor core executes PV1(j) and S1(j+1), and vice versa, so the MMA
// pipe and the exp/ALU pipes overlap inside one CTA. O accumulates IN TMEM across K/V tiles (tcgen05.mma accumulate);
// the online-softmax rescale of O is lazy: the running exponent reference m_used only moves (and O / l are rescaled
// through tcgen05.ld/st) when a row maximum grows by more than 2^8, which keeps P <= 256 in fp16 and makes the
// correction rare after the first tile. The result O / l is exact in exact arithmetic for any reference m_used.
// TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) (512 columns allocated, one CTA per SM).
#include "common.cuh"
#include "host.h"

namespace vton {

struct Attn5Params {
  __half* out;
  int ld_out;
  int B, H, Nq, N0, N1;
  int kv1_off, kv1_count;
  const int* kv1_base;
  float scale_log2;
  int accumulate;
};

constexpr int A5_TILE = 128 * 128;          // bytes of one 128 x 64 fp16 tile
constexpr int A5_STAGES = 4;
constexpr int A5_OFF_Q = 0;                 // Q0, Q1
constexpr int A5_OFF_KV = 2 * A5_TILE;      // stage s: K at +s*2*TILE, V at +s*2*TILE + TILE
constexpr int A5_OFF_ONES = A5_OFF_KV + A5_STAGES * 2 * A5_TILE;   // constant [128 keys x 64] tile: column 0 = 1, rest 0
constexpr int A5_OFF_BAR = A5_OFF_ONES + A5_TILE;
constexpr int A5_OCOLS = 80;                                // 64 value columns + the ones atom (16 columns, col 64 = l)
constexpr int A5_SMEM_TOTAL = A5_OFF_BAR + 256 + 1024;
constexpr float kLazyThreshold5 = 2.0f;     // log2 domain

__device__ __forceinline__ float ex2k_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
// two exponentials per SFU op: packed fp16 in, packed fp16 out
__device__ __forceinline__ uint32_t ex2_h2c(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld_32x16c(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16c(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__global__ void __launch_bounds__(320, 1)
attn5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
             const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
             const __grid_constant__ CUtensorMap tmV1, const Attn5Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + A5_OFF_BAR;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + A5_STAGES + s); };
  auto s_full = [&](int g) { return bar_base + 8u * (1 + 2 * A5_STAGES + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (3 + 2 * A5_STAGES + g); };
  auto o_full = [&](int g) { return bar_base + 8u * (5 + 2 * A5_STAGES + g); };
  const uint32_t tmem_slot = bar_base + 8u * (7 + 2 * A5_STAGES);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + A5_OFF_BAR + 8 * (7 + 2 * A5_STAGES));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x;
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
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK0);
    tma_prefetch_desc(&tmV0);
    if (tiles1) {
      tma_prefetch_desc(&tmK1);
      tma_prefetch_desc(&tmV1);
    }
    mbar_init(q_full, 1);
    for (int s = 0; s < A5_STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 4);   // one arrive per softmax warp of the group
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  if (warp >= 2 && warp < 6) {
    // constant B-operand tile: key row r, logical column 0 = 1.0, all else 0 (128B-swizzled like a V tile)
    const int r = (warp - 2) * 32 + lane;
    uint4* rowp = reinterpret_cast<uint4*>(smem_gen + A5_OFF_ONES + r * 128);
#pragma unroll
    for (int c = 0; c < 8; ++c) rowp[c] = make_uint4(0, 0, 0, 0);
    rowp[r & 7] = make_uint4(0x00003c00u, 0, 0, 0);   // chunk 0 sits at position (0 ^ (r & 7)); 0x3c00 = fp16 1.0
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * A5_TILE);
      tma_load_3d(smem_base + A5_OFF_Q, &tmQ, q_full, h * 64, q_pair * 256, b);
      tma_load_3d(smem_base + A5_OFF_Q + A5_TILE, &tmQ, q_full, h * 64, q_pair * 256 + 128, b);
      for (int j = 0; j < total; ++j) {
        const int stage = j % A5_STAGES;
        const uint32_t phase = (j / A5_STAGES) & 1;
        mbar_wait(kv_empty(stage), phase ^ 1);
        mbar_expect_tx(kv_full(stage), 2 * A5_TILE);
        const uint32_t kdst = smem_base + A5_OFF_KV + stage * 2 * A5_TILE;
        const uint32_t vdst = kdst + A5_TILE;
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
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, A5_OCOLS, 1);
      auto issue_s = [&](int g, int stage) {
        const uint32_t qsrc = smem_base + A5_OFF_Q + g * A5_TILE;
        const uint32_t ksrc = smem_base + A5_OFF_KV + stage * 2 * A5_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tmem_base + g * 128, make_smem_desc_sw128(qsrc + k * 32, 0, 1024),
                     make_smem_desc_sw128(ksrc + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        tc_commit(s_full(g));
      };
      mbar_wait(q_full, 0);
      mbar_wait(kv_full(0), 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < total; ++j) {
        const int stage = j % A5_STAGES;
        const int nstage = (j + 1) % A5_STAGES;
        const uint32_t vsrc = smem_base + A5_OFF_KV + stage * 2 * A5_TILE + A5_TILE;
        const uint32_t ones_src = smem_base + A5_OFF_ONES;
        for (int g = 0; g < 2; ++g) {
          mbar_wait(p_full(g), j & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // A = P from TMEM (the S_g columns, 16 keys = 8 packed 32-bit columns per K step);
            // B = [V | ones]: MN atom 0 = this V tile, MN atom 1 (via the leading byte offset) = the constant tile
            const uint64_t b_desc = make_smem_desc_sw128(vsrc + k * 2048, ones_src - vsrc, 1024);
            tc_mma_f16_ts(tmem_base + 256 + g * A5_OCOLS, tmem_base + g * 128 + k * 8, b_desc, idesc_o,
                          (j > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(o_full(g));
          if (g == 1) tc_commit(kv_empty(stage));   // K(j) and V(j) are no longer needed
          if (j + 1 < total) {
            if (g == 0) {
              mbar_wait(kv_full(nstage), ((j + 1) / A5_STAGES) & 1);
              tc_fence_after();
            }
            issue_s(g, nstage);
          }
        }
      }
    }
  } else {
    // ===================== softmax (2 groups x 4 warps, thread = query row) =====================
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int q_idx = q_pair * 256 + g * 128 + row;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_addr;
    const uint32_t tO = tmem_base + 256 + g * A5_OCOLS + lane_addr;
    const float sl2 = p.scale_log2;
    float m_used = 0.f;

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
      } else if ((mx - m_used) * sl2 > kLazyThreshold5) {
        alpha = ex2k_((m_used - mx) * sl2);
        m_used = mx;
        rescale = true;
      }
      const float msc = m_used * sl2;
      uint32_t pk[2][32];    // P row, packed pairs: 32-bit column c*16+i holds keys (c*32 + 2i, c*32 + 2i + 1)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = __uint_as_float(s[c][2 * i]) * sl2 - msc;
          const float x1 = __uint_as_float(s[c][2 * i + 1]) * sl2 - msc;
          pk[c >> 1][(c & 1) * 16 + i] = ex2_h2c(pack_h2(x0, x1));
        }
      }
      if (j > 0) {
        // S_g(j) was issued after PV_g(j-1) on the in-order tensor pipe, so its completion (s_full) already implies
        // that PV_g(j-1) has retired (O stable, old P dead); this wait returns at once and only keeps the barrier's
        // phase parity in lock-step for the final wait.
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
          {
            uint32_t o16[16];
            tmem_ld_32x16c(tO + 64, o16);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o16[i] = __float_as_uint(__uint_as_float(o16[i]) * alpha);
            tmem_st_32x16c(tO + 64, o16);
          }
          tmem_st_wait();
        }
      }
      // P overwrites the first 64 columns of this thread's S row (all 128 S columns are already in registers)
      tmem_st_32x32(tS, pk[0]);
      tmem_st_32x32(tS + 32, pk[1]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(g));
    }
    // ---- finalize
    mbar_wait(o_full(g), (total - 1) & 1);
    tc_fence_after();
    float o[64];
    float l_run;
    {
      uint32_t r0[32], r1[32], r2[16];
      tmem_ld_32x32(tO, r0);
      tmem_ld_32x32(tO + 32, r1);
      tmem_ld_32x16c(tO + 64, r2);
      tmem_ld_wait();
      l_run = __uint_as_float(r2[0]);   // column 64 = sum_j P_ij accumulated by the tensor core
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
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

int attn5_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK0, const CUtensorMap& tmV0, const CUtensorMap& tmK1,
                 const CUtensorMap& tmV1, __half* out, int ld_out, int B, int H, int Nq, int N0, int N1, int kv1_off,
                 int kv1_count, const int* kv1_base, float scale_log2, int accumulate, cudaStream_t stream) {
  Attn5Params p{};
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
  static bool configured = false;
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(attn5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A5_SMEM_TOTAL));
    configured = true;
  }
  dim3 grid((Nq + 255) / 256, H, B);
  attn5_kernel<<<grid, 320, A5_SMEM_TOTAL, stream>>>(tmQ, tmK0, tmV0, tmK1, tmV1, p);
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vton
