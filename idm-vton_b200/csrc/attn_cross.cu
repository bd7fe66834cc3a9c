// Decoupled cross-attention of the try-on / garment transformer blocks in ONE launch (sm_100a, head_dim 64):
//     out = fp16( fp16(softmax(Q Kt^T * scale) Vt) + fp16(ip_scale * fp16(softmax(Q Ki^T * scale) Vi)) )
// Kt/Vt = the 77 text tokens (attn2.to_k / to_v), Ki/Vi = the 16 IP-Adapter image tokens (to_k_ip / to_v_ip); the two
// softmaxes are independent and the fp16 outputs are summed in fp16, exactly the rounding points of the reference
// (ip_adapter/attention_processor.py IPAttnProcessor2_0.__call__: two F.scaled_dot_product_attention calls, then
// hidden_states + self.scale * ip_hidden_states; src/attentionhacked_tryon.py:368-380 calls it as attn2).
// With Ni = 0 it is the plain text cross-attention of the garment UNet (src/attentionhacked_garmnet.py:371-383).
//
// Both key sets fit ONE score tile: keys [0,80) = text (Nt <= 80, padding masked), keys [80,96) = image tokens
// (Ni <= 16). There is no K/V loop and no online softmax, the whole problem is launch/latency/HBM bound (Q in, O out:
// 31 MB at the C = 1280 level), so the kernel is a straight line per 128-query tile and relies on CTA co-residency
// (2 per SM by tensor memory) for overlap; the general kernel needed 2 launches and ~45 us for the same work.
//   warp 0      TMA: Q tile, K = [Kt ; Ki] (one barrier), V = [Vt ; Vi] (second barrier)
//   warp 1      tcgen05.mma: S = Q K^T (N = 96 or 80), then O_t = P[:, :80] Vt and O_i = P[:, 80:] Vi (A = P from TMEM)
//   warps 2..5  softmax (thread = query row): per-segment max / exp2 / sum in fp32, P (unnormalised, <= 1) packed to
//               fp16 into the TMEM columns S occupied; epilogue O_t / l_t and O_i / l_i, the fp16 roundings above,
//               staged through the (dead) Q tile in the 128B-swizzle layout and written with one TMA store per warp.
// TMEM (256 columns): S [0,96) (P aliases [0,48): packed chunk c only covers S columns already consumed),
//                     O_t [128,192), O_i [192,256).
#include "common.cuh"
#include "gemm_common.cuh"
#include "host.h"

namespace vton {

struct CrossParams {
  int Nq, Nt, Ni;
  float scale_log2;
  float ip_scale;
};

constexpr int XA_KT = 80;                       // padded text keys
constexpr int XA_KI = 16;                       // padded image keys
constexpr int XA_OFF_Q = 0;                     // 128 x 64 fp16, also the output staging tile
constexpr int XA_OFF_K = 128 * 128;             // 96 rows x 128 B
constexpr int XA_OFF_V = XA_OFF_K + 96 * 128;
constexpr int XA_OFF_BAR = XA_OFF_V + 96 * 128;
constexpr int XA_SMEM_TOTAL = XA_OFF_BAR + 64 + 1024;
constexpr uint32_t XA_TMEM_COLS = 256;
constexpr uint32_t XA_TM_OT = 128, XA_TM_OI = 192;

__device__ __forceinline__ float xa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void xa_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void xa_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void xa_tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(192, 2)
cross_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKt,
                  const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ CUtensorMap tmKi,
                  const __grid_constant__ CUtensorMap tmVi, const __grid_constant__ CUtensorMap tmO,
                  const CrossParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + XA_OFF_BAR;
  const uint32_t qk_full = bar_base, v_full = bar_base + 8, s_full = bar_base + 16, p_full = bar_base + 24,
                 o_full = bar_base + 32;
  const uint32_t tmem_slot = bar_base + 40;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + XA_OFF_BAR + 40);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const bool has_ip = p.Ni > 0;
  const int keys = has_ip ? (XA_KT + XA_KI) : XA_KT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKt);
    tma_prefetch_desc(&tmVt);
    tma_prefetch_desc(&tmO);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  // Programmatic dependent launch: wait for the previous kernel BEFORE allocating tensor memory. A dependent CTA that
  // grabbed TMEM first and then waited could starve a primary CTA on the same SM that had signalled its dependents but
  // not yet allocated its own columns (library kernels signal at their very start): neither would ever proceed.
  pdl_wait();
  if (warp == 1) tmem_alloc<XA_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();   // this CTA holds everything it needs: the next kernel may start its own set-up

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qk_full, 128 * 128 + keys * 128);
      tma_load_3d(smem_base + XA_OFF_Q, &tmQ, qk_full, h * 64, q_tile * 128, b);
      tma_load_3d(smem_base + XA_OFF_K, &tmKt, qk_full, h * 64, 0, b);
      if (has_ip) tma_load_3d(smem_base + XA_OFF_K + XA_KT * 128, &tmKi, qk_full, h * 64, 0, b);
      mbar_expect_tx(v_full, keys * 128);
      tma_load_3d(smem_base + XA_OFF_V, &tmVt, v_full, h * 64, 0, b);
      if (has_ip) tma_load_3d(smem_base + XA_OFF_V + XA_KT * 128, &tmVi, v_full, h * 64, 0, b);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = has_ip ? make_idesc_f16(128, XA_KT + XA_KI, 0) : make_idesc_f16(128, XA_KT, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 1);
      mbar_wait(qk_full, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_f16(tmem_base, make_smem_desc_sw128(smem_base + XA_OFF_Q + k * 32, 0, 1024),
                   make_smem_desc_sw128(smem_base + XA_OFF_K + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
      tc_commit(s_full);
      mbar_wait(v_full, 0);
      mbar_wait(p_full, 0);
      tc_fence_after();
      const uint32_t vsrc = smem_base + XA_OFF_V;
#pragma unroll
      for (int k = 0; k < XA_KT / 16; ++k)     // text keys: 16 per step = 8 packed P columns, 2048 B of V rows
        xa_mma_ts(tmem_base + XA_TM_OT, tmem_base + k * 8, make_smem_desc_sw128(vsrc + k * 2048, 128 * 128, 1024),
                  idesc_o, k > 0 ? 1u : 0u);
      if (has_ip)
        xa_mma_ts(tmem_base + XA_TM_OI, tmem_base + (XA_KT / 16) * 8,
                  make_smem_desc_sw128(vsrc + (XA_KT / 16) * 2048, 128 * 128, 1024), idesc_o, 0u);
      tc_commit(o_full);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr;
    const float sl2 = p.scale_log2;
    mbar_wait(s_full, 0);
    tc_fence_after();
    // pass 1: the two row maxima
    float mt = -INFINITY, mi = -INFINITY;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t s[32];
      tmem_ld_32x32(tS + c * 32, s);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = c * 32 + i;
        const float x = __uint_as_float(s[i]);
        if (col < XA_KT) {
          if (col < p.Nt) mt = fmaxf(mt, x);
        } else if (col - XA_KT < p.Ni) {
          mi = fmaxf(mi, x);
        }
      }
    }
    // pass 2: exponentials (unnormalised, <= 1), row sums, packed fp16 P over the consumed S columns
    const float mts = mt * sl2, mis = mi * sl2;
    float lt = 0.f, li = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t s[32];
      tmem_ld_32x32(tS + c * 32, s);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float e[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int col = c * 32 + 2 * i + u;
          const float x = __uint_as_float(s[2 * i + u]);
          float v = 0.f;
          if (col < XA_KT) {
            if (col < p.Nt) {
              v = xa_ex2(x * sl2 - mts);
              lt += v;
            }
          } else if (col - XA_KT < p.Ni) {
            v = xa_ex2(x * sl2 - mis);
            li += v;
          }
          e[u] = v;
        }
        pk[i] = pack_h2(e[0], e[1]);
      }
      xa_tmem_st16(tS + c * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full);

    // epilogue
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_t = 1.f / lt;
    const float inv_i = has_ip ? 1.f / li : 0.f;
    uint8_t* stage_row = smem_gen + XA_OFF_Q + row * 128;   // Q tile is dead once S has been produced
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t ot[32], oi[32];
      tmem_ld_32x32(tS + XA_TM_OT + c * 32, ot);
      if (has_ip) tmem_ld_32x32(tS + XA_TM_OI + c * 32, oi);
      tmem_ld_wait();
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float y0 = round_h(__uint_as_float(ot[2 * i]) * inv_t);
        float y1 = round_h(__uint_as_float(ot[2 * i + 1]) * inv_t);
        if (has_ip) {
          y0 += round_h(p.ip_scale * round_h(__uint_as_float(oi[2 * i]) * inv_i));
          y1 += round_h(p.ip_scale * round_h(__uint_as_float(oi[2 * i + 1]) * inv_i));
        }
        w[i] = pack_h2(y0, y1);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {   // 16-byte chunk (c*4 + g) of the row goes to slot chunk ^ (row & 7)
        const int chunk = c * 4 + g;
        *reinterpret_cast<uint4*>(stage_row + ((chunk ^ (row & 7)) << 4)) =
            make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      xa_tma_store_3d(&tmO, smem_base + XA_OFF_Q + quarter * 32 * 128, h * 64, q_tile * 128 + quarter * 32, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<XA_TMEM_COLS>(tmem_base);
  }
}

static int encode_rows(CUtensorMap* tm, const void* base, long long ld, int cols, int n, int batch, uint32_t box_rows) {
  uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(n), static_cast<uint64_t>(batch)};
  uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(n) * ld * 2};
  uint32_t box[3] = {64, box_rows, 1};
  return encode_tmap_f16(tm, base, 3, dims, strides, box);
}

// q/out: [B, Nq, >= H*64]; kt/vt: [B, Nt, .] (row stride ldkv_t); ki/vi: [B, Ni, .] (ldkv_i) or null with Ni = 0
int cross_attn_impl(const void* q, long long ldq, const void* kt, const void* vt, long long ldkv_t, int Nt,
                    const void* ki, const void* vi, long long ldkv_i, int Ni, void* out, long long ldo, int B, int H,
                    int Nq, float scale, float ip_scale, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && H > 0 && Nq > 0, "cross_attn: bad sizes B=%d H=%d Nq=%d", B, H, Nq);
  VTON_CHECK_ARG(Nt > 0 && Nt <= XA_KT && Ni >= 0 && Ni <= XA_KI, "cross_attn: needs 1 <= Nt <= 80 and Ni <= 16 (got %d, %d)", Nt, Ni);
  VTON_CHECK_ARG(q && kt && vt && out && (Ni == 0 || (ki && vi)), "cross_attn: null pointer");
  VTON_CHECK_ARG(ldq % 8 == 0 && ldkv_t % 8 == 0 && ldo % 8 == 0 && (Ni == 0 || ldkv_i % 8 == 0),
                 "cross_attn: row strides must be multiples of 8");
  VTON_CHECK_ARG(B <= 65535 && H <= 65535, "cross_attn: grid too large");
  CUtensorMap tmQ, tmKt, tmVt, tmKi, tmVi, tmO;
  if (int e = encode_rows(&tmQ, q, ldq, H * 64, Nq, B, 128)) return e;
  if (int e = encode_rows(&tmKt, kt, ldkv_t, H * 64, Nt, B, XA_KT)) return e;
  if (int e = encode_rows(&tmVt, vt, ldkv_t, H * 64, Nt, B, XA_KT)) return e;
  tmKi = tmKt;
  tmVi = tmVt;
  if (Ni > 0) {
    if (int e = encode_rows(&tmKi, ki, ldkv_i, H * 64, Ni, B, XA_KI)) return e;
    if (int e = encode_rows(&tmVi, vi, ldkv_i, H * 64, Ni, B, XA_KI)) return e;
  }
  if (int e = encode_rows(&tmO, out, ldo, H * 64, Nq, B, 32)) return e;
  CrossParams p{};
  p.Nq = Nq;
  p.Nt = Nt;
  p.Ni = Ni;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.ip_scale = ip_scale;
  static bool configured = false;
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(cross_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM_TOTAL));
    configured = true;
  }
  dim3 grid(cdiv(Nq, 128), H, B);
  VTON_CUDA(launch_kernel(cross_attn_kernel, grid, dim3(192), XA_SMEM_TOTAL, stream, tmQ, tmKt, tmVt, tmKi, tmVi, tmO, p));
  count_launch();
  return kOk;
}

}  // namespace vton
