// tcgen05 GEMM / implicit-GEMM 3x3 convolution for the IDM-VTON UNets (sm_100a).
//
//   linear : out[M,N] = epi( A[M,K] . W[N,K]^T )                (K8 in SURVEY.md 2.3: to_q/k/v/out, proj_in/out, FF)
//   conv3x3: out[b,y,x,:] = epi( sum_taps A[b,y+dy,x+dx,:] . W[tap][:, :]^T  (+ 1x1 shortcut in a 2nd accumulator) )
//            (K1/K2/K5: ResnetBlock2D conv1/conv2/conv_shortcut, conv_in, conv_out; NHWC, pad 1, stride 1)
//
// One CTA computes one 128 x BN output tile. Warp roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM
// allocator), warps 2..5 = epilogue (TMEM -> registers -> fp16 -> global). Operands are staged by TMA into
// 128B-swizzled shared memory, accumulators live in TMEM (fp32). The conv walks K as 9 taps x Cin/64 slabs with a 4-D
// tensor map over [B,H,W,C]; out-of-bounds box elements are zero-filled by TMA, which is the conv's zero padding.
// Epilogue rounding points replicate the reference's fp16-autocast path:
//   v = fp16(acc + bias); v = fp16(v + temb[b,n]); s = fp16(acc_sc + bias_sc); v = fp16(s + v); v = fp16(v + res)
#include "gemm_common.cuh"
#include "host.h"

namespace vton {

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
};

template <int BN, int STAGES, bool GEGLU>
__global__ void __launch_bounds__(192, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmS0, const __grid_constant__ CUtensorMap tmS1,
                 const __grid_constant__ CUtensorMap tmBs, const GemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int n_tile = tile % p.n_tiles;
  const int m_tile = tile / p.n_tiles;
  const int n0 = n_tile * BN;
  const int total_slabs = p.slabs_main + p.slabs_sc;
  constexpr uint32_t kTmemCols = (BN <= 64) ? 128 : (BN <= 128 ? 256 : 512);  // room for the shortcut accumulator

  // conv tile origin
  int b0 = 0, y0 = 0, x0 = 0;
  if (p.conv) {
    const int tx = m_tile % p.tiles_x;
    const int ty = (m_tile / p.tiles_x) % p.tiles_y;
    const int tb = m_tile / (p.tiles_x * p.tiles_y);
    x0 = tx * p.bw;
    y0 = ty * p.bh;
    b0 = tb * p.bb;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.slabs_sc) {
      tma_prefetch_desc(&tmS0);
      tma_prefetch_desc(&tmS1);
      tma_prefetch_desc(&tmBs);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  // Programmatic dependent launch: wait for the previous kernel BEFORE allocating tensor memory. A dependent CTA that
  // grabbed TMEM first and then waited could starve a primary CTA on the same SM that had signalled its dependents but
  // not yet allocated its own columns (library kernels signal at their very start): neither would ever proceed.
  pdl_wait();
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();   // this CTA holds everything it needs: the next kernel may start its own set-up

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int s = 0; s < total_slabs; ++s) {
        const int stage = s % STAGES;
        const uint32_t phase = (s / STAGES) & 1;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t a_dst = smem_base + stage * L::STAGE_BYTES;
        const uint32_t b_dst = a_dst + A_BYTES;
        mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
        if (s < p.slabs_main) {
          if (p.conv) {
            const int tap = s / p.cin_slabs;
            const int c0 = (s - tap * p.cin_slabs) * BK;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            tma_load_4d(a_dst, &tmA, full_bar(stage), c0, x0 * p.stride + dx, y0 * p.stride + dy, b0);
            tma_load_2d(b_dst, &tmB, full_bar(stage), c0, tap * p.cout + n0);
          } else {
            tma_load_2d(a_dst, &tmA, full_bar(stage), s * BK, m_tile * BM);
            tma_load_2d(b_dst, &tmB, full_bar(stage), s * BK, n0);
          }
        } else {
          const int ss = s - p.slabs_main;
          if (ss < p.sc_split)
            tma_load_4d(a_dst, &tmS0, full_bar(stage), ss * BK, x0, y0, b0);
          else
            tma_load_4d(a_dst, &tmS1, full_bar(stage), (ss - p.sc_split) * BK, x0, y0, b0);
          tma_load_2d(b_dst, &tmBs, full_bar(stage), ss * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, 0);
      for (int s = 0; s < total_slabs; ++s) {
        const int stage = s % STAGES;
        const uint32_t phase = (s / STAGES) & 1;
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_src = smem_base + stage * L::STAGE_BYTES;
        const uint32_t b_src = a_src + A_BYTES;
        const bool sc = s >= p.slabs_main;
        const uint32_t d = tmem_base + (sc ? BN : 0);
        const int s_local = sc ? s - p.slabs_main : s;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t a_desc = make_smem_desc_sw128(a_src + k * 32, 0, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(b_src + k * 32, 0, 1024);
          tc_mma_f16(d, a_desc, b_desc, idesc, (s_local > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
      }
      tc_commit(acc_bar);  // accumulators complete
    }
  } else {
    // ===================== epilogue (4 warps, thread = accumulator row) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    long long out_row;
    int sample;
    map_row(p, m_tile, quarter * 32 + lane, &out_row, &sample);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    epilogue_store<BN, GEGLU>(p, t_row, BN, n_tile, out_row, sample);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool GEGLU>
static int launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmS0,
                          const CUtensorMap& tmS1, const CUtensorMap& tmBs, const GemmParams& p, int grid,
                          cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gemm_conv_kernel<BN, STAGES, GEGLU>;
  static bool configured = false;
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  VTON_CUDA(launch_kernel(kern, dim3(grid), dim3(192), L::TOTAL, stream, tmA, tmB, tmS0, tmS1, tmBs, p));
  count_launch();
  return kOk;
}

int gemm2_dispatch(int bn, bool geglu, const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, int m_tiles,
                   cudaStream_t stream, int np = 1, const CUtensorMap* tmS0 = nullptr, const CUtensorMap* tmS1 = nullptr,
                   const CUtensorMap* tmBs = nullptr);

// force_bn encoding: 0 = automatic kernel / tile choice; 64..256 = 1-CTA kernel (gemm.cu) with that tile width;
// 1000 + {128,160,192,256} = 2-CTA persistent kernel (gemm2.cu) with that tile width.
static int g_auto_v2 = 1;
void set_auto_v2(int on) { g_auto_v2 = on; }
// "gemm_cluster4": 1 lets large linear layers with 256-wide tiles run in four-CTA clusters that multicast the A slabs.
// Measured SLOWER than two-CTA clusters at every config-2 shape (profiles/r1_gemm_cluster4.jsonl: 8192^3 1003 vs 1359
// TFLOP/s, FF2 763 vs 1298) — four-CTA multicast does not lower the L2 output load on this part and the coupled
// pairs lose slack — so it stays off; kept as a selectable variant (force_bn = 2256) with its parity test.
static int g_cluster4 = 0;
void set_cluster4(int on) { g_cluster4 = on; }

// Tile width of the 2-CTA kernel from a cost model instead of divisibility rules (round 2). Per 64-deep K slab a CTA pair
// spends max(MMA time, operand-delivery time): the 256 x BN x 64 MMA takes 2*BN clk (8192 dense fp16 FLOP/clk/SM), and each
// CTA pulls 16 KB of A + 64*BN bytes of W through the L2 -> SM path, whose chip-wide cap is ~6300 B/clk
// (/opt/skills/guides/B300_MICROARCH.md "LTS throughput cap"; measured here as 40-42 B/clk/SM with all SMs pulling). So
// narrow tiles are delivery-bound (BN 128: 577 clk per slab for 256 MMA clk) and the cost per output column falls with
// BN (4.5 / 3.9 / 3.5 / 3.0 clk per column and slab for BN 128 / 160 / 192 / 256). Against that stands quantisation:
// tiles = ceil(M/256) * ceil(N/BN) run in rounds of 74 clusters. Example the old rule got wrong: N = 640, M = 12288 —
// BN 160 gives 192 tiles = 3 rounds x 625 clk, BN 256 gives 144 tiles (the third column tile half empty) = 2 rounds x
// 769 clk: 18% less, which is where cuBLAS was ahead (profiles/r1_microbench_gemm2_final.jsonl: 1124 vs 786 TFLOP/s on
// [12288 x 640 x 2560]). Cout = 320 at M = 49152 still resolves to 160 (6 rounds either way).
static int pick_bn2(int N, bool geglu, int m_tiles, bool has_shortcut) {
  const int m_pairs = cdiv(m_tiles, 2);
  const int max_clusters = kSMs / 2;
  const int cands[4] = {256, 192, 160, 128};
  double best = 1e300;
  int best_bn = 0;
  for (int bn : cands) {
    if (geglu && ((bn != 256 && bn != 128) || N % bn != 0)) continue;
    if (has_shortcut && bn == 192) continue;   // two accumulators per tile: 2 * 192 columns do not fit a power-of-two alloc
    const long long tiles = static_cast<long long>(m_pairs) * cdiv(N, bn);
    const long long clusters = tiles < max_clusters ? tiles : max_clusters;
    const long long rounds = (tiles + clusters - 1) / clusters;
    double bw = 6300.0 / (2.0 * clusters);   // B/clk per SM when `clusters` pairs pull at once
    if (bw > 80.0) bw = 80.0;                // a lone SM does not get the whole crossbar
    const double mma = 2.0 * bn;
    const double l2 = (16384.0 + 64.0 * bn) / bw;
    const double cost = rounds * (mma > l2 ? mma : l2);
    if (cost < best * 0.999) {               // ties go to the wider tile (fewer bytes per FLOP)
      best = cost;
      best_bn = bn;
    }
  }
  return best_bn ? best_bn : 128;
}

// Returns the 2-CTA tile width to use, or 0 for the 1-CTA kernel.
static int choose_v2(int m_tiles, int N, bool geglu, bool has_shortcut, int force_bn) {
  int bn = 0;
  if (force_bn >= 1000) {
    bn = force_bn - 1000;
  } else if (force_bn == 0 && g_auto_v2 && m_tiles >= 4 && N >= 128 && !(geglu && N % 128 != 0)) {
    bn = pick_bn2(N, geglu, m_tiles, has_shortcut);
  }
  // with a fused shortcut the tile carries two accumulators (one TMEM stage): widths whose pair fits 512 columns
  if (has_shortcut && bn == 192) bn = 0;
  return bn;
}

static int pick_bn(int N, int force_bn) {
  if (force_bn) return force_bn;
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 160 == 0) return 160;
  if (N <= 64) return 64;
  return 128;
}

static int dispatch(int bn, bool geglu, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmS0,
                    const CUtensorMap& tmS1, const CUtensorMap& tmBs, GemmParams& p, int m_tiles,
                    cudaStream_t stream) {
  p.n_tiles = cdiv(p.N, bn);
  const int grid = m_tiles * p.n_tiles;
  if (geglu) {
    if (bn == 128) return launch_variant<128, 3, true>(tmA, tmB, tmS0, tmS1, tmBs, p, grid, stream);
    if (bn == 256) return launch_variant<256, 4, true>(tmA, tmB, tmS0, tmS1, tmBs, p, grid, stream);
    set_last_error("GEGLU epilogue supports BN 128/256 only (got %d)", bn);
    return kErrUnsupported;
  }
  switch (bn) {
    case 64: return launch_variant<64, 4, false>(tmA, tmB, tmS0, tmS1, tmBs, p, grid, stream);
    case 128: return launch_variant<128, 3, false>(tmA, tmB, tmS0, tmS1, tmBs, p, grid, stream);
    case 160: return launch_variant<160, 3, false>(tmA, tmB, tmS0, tmS1, tmBs, p, grid, stream);
    case 256: return launch_variant<256, 4, false>(tmA, tmB, tmS0, tmS1, tmBs, p, grid, stream);
  }
  set_last_error("unsupported BN %d", bn);
  return kErrUnsupported;
}

int gemm_f16_impl(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo, int M, int N,
                  int K, const void* bias, const void* residual, long long ldr, const void* rowvec, long long ld_rowvec,
                  int rows_per_sample, int flags, int force_bn, cudaStream_t stream) {
  const int geglu = flags & 1;
  VTON_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VTON_CHECK_ARG(K % 64 == 0, "gemm: K=%d must be a multiple of 64", K);
  VTON_CHECK_ARG(N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0, "gemm: N/lda/ldw/ldo must be multiples of 8");
  VTON_CHECK_ARG(!geglu || (N % 16 == 0 && !residual && !rowvec), "gemm: bad GEGLU configuration");
  int np = 1;
  if (force_bn >= 2000) {   // 2000 + bn: the 2-CTA kernel in four-CTA clusters (A multicast)
    np = 2;
    force_bn -= 1000;
  }
  const int bn2 = choose_v2(cdiv(M, BM), N, geglu != 0, false, force_bn);
  if (force_bn == 0 && g_cluster4 && bn2 == 256 && cdiv(N, 256) >= 2 && M >= 1024 && K >= 512) np = 2;
  VTON_CHECK_ARG(bn2 == 0 || bn2 == 128 || bn2 == 160 || bn2 == 192 || bn2 == 256, "gemm: bad 2-CTA tile width %d", bn2);
  int bn = bn2 ? bn2 : pick_bn(N, force_bn);
  if (geglu && bn != 128 && bn != 256) bn = 128;
  VTON_CHECK_ARG(!geglu || N % bn == 0, "gemm: GEGLU needs N %% BN == 0 (N=%d, BN=%d)", N, bn);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(np == 2 ? 64 : 128)};   // four-CTA clusters fetch A in 64-row halves
    if (int e = encode_tmap_f16(&tmA, A, 2, dims, strides, box)) return e;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn2 ? bn / 2 : bn)};   // 2-CTA: each CTA loads half the tile
    if (int e = encode_tmap_f16(&tmB, W, 2, dims, strides, box)) return e;
  }
  GemmParams p{};
  p.out = static_cast<__half*>(out);
  p.ld_out = static_cast<int>(ldo);
  p.M = M;
  p.N = N;
  p.bias = static_cast<const __half*>(bias);
  p.residual = static_cast<const __half*>(residual);
  p.ld_res = static_cast<int>(ldr);
  p.rowvec = static_cast<const __half*>(rowvec);
  p.ld_rowvec = static_cast<int>(ld_rowvec);
  p.rows_per_sample = rows_per_sample;
  p.slabs_main = K / 64;
  p.act_gelu = (flags & 4) ? 2 : ((flags & 2) ? 1 : 0);
  if (bn2) return gemm2_dispatch(bn, geglu != 0, tmA, tmB, p, cdiv(M, BM), stream, np);
  return dispatch(bn, geglu != 0, tmA, tmB, tmA, tmA, tmB, p, cdiv(M, BM), stream);
}

static void pick_box(int B, int H, int W, int* bw, int* bh, int* bb) {
  int w = 1;
  while (w < 128 && W % (w * 2) == 0) w *= 2;
  int h = 1;
  while (w * h < 128 && h * 2 <= H) h *= 2;
  *bw = w;
  *bh = h;
  *bb = 128 / (w * h);
  (void)B;
}

// pixel_step = 2 (stride-2 convolution): the box spans 2*bw x 2*bh input pixels and the TMA traversal stride picks every
// second one, so the tile that lands in shared memory is the same 128 rows x 64 channels as for stride 1 and the
// "im2col" of Downsample2D never exists in memory.
static int encode_nhwc(CUtensorMap* tm, const void* base, int B, int H, int W, int C, int ldc, int bw, int bh, int bb,
                       int pixel_step = 1) {
  uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(ldc) * 2, static_cast<uint64_t>(W) * ldc * 2,
                         static_cast<uint64_t>(H) * W * ldc * 2};
  uint32_t box[4] = {64, static_cast<uint32_t>(bw * pixel_step), static_cast<uint32_t>(bh * pixel_step),
                     static_cast<uint32_t>(bb)};
  uint32_t estr[4] = {1, static_cast<uint32_t>(pixel_step), static_cast<uint32_t>(pixel_step), 1};
  return encode_tmap_f16(tm, base, 4, dims, strides, box, pixel_step > 1 ? estr : nullptr);
}

// x: [B,Hin,Win,Cin] (channel stride ldx), w: [9][Cout][Cin] fp16 (tap-major), out: [B*H*W, ldo] with H = (Hin-1)/stride+1
int conv3x3_impl(const void* x, long long ldx, int B, int Hin, int Win, int Cin, const void* w, int Cout, const void* bias,
                 const void* temb, long long ld_temb, const void* sc0, int C0, const void* sc1, int C1, const void* w_sc,
                 const void* bias_sc, const void* residual, long long ldr, void* out, long long ldo, int force_bn,
                 int stride, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && Hin > 0 && Win > 0, "conv3x3: empty input");
  VTON_CHECK_ARG(stride == 1 || stride == 2, "conv3x3: stride %d unsupported", stride);
  VTON_CHECK_ARG(stride == 1 || (!w_sc && !residual), "conv3x3: stride 2 has no shortcut / residual form");
  const int H = (Hin - 1) / stride + 1, W = (Win - 1) / stride + 1;   // output size (padding 1)
  VTON_CHECK_ARG(Cin % 64 == 0, "conv3x3: Cin=%d must be a multiple of 64 (pad channels)", Cin);
  VTON_CHECK_ARG(Cout % 8 == 0 && ldo % 8 == 0 && ldx % 8 == 0, "conv3x3: Cout/ldo/ldx must be multiples of 8");
  VTON_CHECK_ARG(C0 % 64 == 0 && C1 % 64 == 0, "conv3x3: shortcut source channels must be multiples of 64");
  VTON_CHECK_ARG(!(w_sc && residual), "conv3x3: shortcut conv and identity residual are exclusive");
  int bw, bh, bb;
  pick_box(B, H, W, &bw, &bh, &bb);
  const int m_tiles_est = (W / bw) * cdiv(H, bh) * cdiv(B, bb);
  const int bn2 = choose_v2(m_tiles_est, Cout, false, w_sc != nullptr, force_bn);
  VTON_CHECK_ARG(bn2 == 0 || bn2 == 128 || bn2 == 160 || bn2 == 192 || bn2 == 256, "conv3x3: bad 2-CTA tile width %d", bn2);
  const int bn = bn2 ? bn2 : pick_bn(Cout, force_bn >= 1000 ? 0 : force_bn);
  CUtensorMap tmA, tmB, tmS0, tmS1, tmBs;
  VTON_CHECK_ARG(stride == 1 || (bw * 2 <= 256 && bh * 2 <= 256), "conv3x3: stride-2 box too large");
  if (int e = encode_nhwc(&tmA, x, B, Hin, Win, Cin, static_cast<int>(ldx), bw, bh, bb, stride)) return e;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(9) * Cout};
    uint64_t strides[1] = {static_cast<uint64_t>(Cin) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn2 ? bn / 2 : bn)};
    if (int e = encode_tmap_f16(&tmB, w, 2, dims, strides, box)) return e;
  }
  tmS0 = tmA;
  tmS1 = tmA;
  tmBs = tmB;
  GemmParams p{};
  if (w_sc) {
    VTON_CHECK_ARG(sc0 && C0 > 0, "conv3x3: shortcut needs a source");
    if (int e = encode_nhwc(&tmS0, sc0, B, H, W, C0, C0, bw, bh, bb)) return e;
    if (sc1 && C1 > 0) {
      if (int e = encode_nhwc(&tmS1, sc1, B, H, W, C1, C1, bw, bh, bb)) return e;
    }
    const int Csc = C0 + (sc1 ? C1 : 0);
    uint64_t dims[2] = {static_cast<uint64_t>(Csc), static_cast<uint64_t>(Cout)};
    uint64_t strides[1] = {static_cast<uint64_t>(Csc) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn2 ? bn / 2 : bn)};
    if (int e = encode_tmap_f16(&tmBs, w_sc, 2, dims, strides, box)) return e;
    p.slabs_sc = Csc / 64;
    p.sc_split = C0 / 64;
  }
  p.out = static_cast<__half*>(out);
  p.ld_out = static_cast<int>(ldo);
  p.M = B * H * W;
  p.N = Cout;
  p.bias = static_cast<const __half*>(bias);
  p.bias_sc = static_cast<const __half*>(bias_sc);
  p.residual = static_cast<const __half*>(residual);
  p.ld_res = static_cast<int>(ldr);
  p.rowvec = static_cast<const __half*>(temb);
  p.ld_rowvec = static_cast<int>(ld_temb);
  p.slabs_main = 9 * (Cin / 64);
  p.conv = 1;
  p.stride = stride;
  p.H = H;
  p.W = W;
  p.B = B;
  p.bw = bw;
  p.bh = bh;
  p.bb = bb;
  p.tiles_x = W / bw;
  p.tiles_y = cdiv(H, bh);
  p.cin_slabs = Cin / 64;
  p.cout = Cout;
  const int m_tiles = p.tiles_x * p.tiles_y * cdiv(B, bb);
  if (bn2) return gemm2_dispatch(bn, false, tmA, tmB, p, m_tiles, stream, 1, &tmS0, &tmS1, &tmBs);
  return dispatch(bn, false, tmA, tmB, tmS0, tmS1, tmBs, p, m_tiles, stream);
}

}  // namespace vton
