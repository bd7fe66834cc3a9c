// 2-CTA persistent tcgen05 GEMM / implicit-GEMM conv (sm_100a): the high-throughput path for the large linears and
// 3x3 convolutions of both UNets (same math, operands, epilogue and rounding points as gemm.cu).
//
// A cluster of two CTAs (one SM pair) owns a 256 x BN output tile: CTA r stages its own 128 rows of A and one half
// (BN/2 rows) of the weight tile per K slab, and the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) that reads
// both CTAs' shared memory, so every operand byte fetched from L2 feeds twice the math of the 1-CTA kernel.
// The kernel is persistent (one cluster per SM pair, static round-robin over tiles) and the fp32 accumulators are
// double-buffered in TMEM (2 x BN columns), so the epilogue of tile i (TMEM -> registers -> fp16 -> swizzled smem ->
// per-warp TMA stores) overlaps the main loop of tile i+1.
//   warp 0: TMA producer (both CTAs)      warp 1: MMA issuer (leader CTA) + TMEM alloc (both)
//   warps 2..9: epilogue (both CTAs; thread = accumulator row of the CTA's own 128-row half; the two warps that share a
//               TMEM lane quarter take the even / odd 32-column chunks of the tile)
// Barriers: full[s] lives in the leader (both CTAs' TMA bytes are credited to it), empty[s] / tmem_full[a] are
// multicast-committed to both CTAs, tmem_empty[a] lives in the leader and collects one arrive per epilogue warp (16).
#include "gemm_common.cuh"
#include "host.h"

namespace vton {

// One staging buffer per chunk a warp handles per tile: the epilogue never waits for a store inside a tile. (A one-slot
// ring that bought a sixth operand stage — the "deep pipeline" experiment queued at the end of round 1 — was measured in
// round 2 and removed: FF1 815 vs 1336 TFLOP/s, FF2 704 vs 1202, 8192^3 1287 vs 1365; profiles/r2_small_kernels.jsonl.
// The main loop is bound by the L2 -> SM delivery rate, not by bytes in flight, and the serialised staging stalls the
// epilogue.)
template <int BN, int STAGES>
struct Smem2 {
  static constexpr int B_BYTES = (BN / 2) * BK * 2;          // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue staging: per warp one 32-row x 32-column fp16 buffer (2 KB, 64B-swizzled) per chunk of the tile
  static constexpr int STORE_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int RING = (BN / 32 + 1) / 2;   // staging buffers per epilogue warp
  static constexpr int STORE_BYTES = 8 * RING * 2048;
  static constexpr int BAR_OFFSET = STORE_OFFSET + STORE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
};

// NP = MMA pairs per cluster. NP = 2 (linear layers only): a cluster of four CTAs owns two N-adjacent 256 x BN tiles
// of the same 256 rows; CTA (pair, r) fetches one 64-row half of the pair-independent A slab r and multicasts it to
// both pairs, so the L2 slices serve 24 KB instead of 32 KB per CTA and K slab — the chip-wide L2 output rate
// (~6300 B/clk, 42 B/clk/SM) is what bounds these GEMMs at ~65% tensor-pipe utilisation, not the tensor cores.
// A slot is recycled when BOTH pairs' MMAs have retired (empty barriers count NP multicast commits).
// SC = the resnet's 1x1 shortcut conv rides along (conv mode only): after the 9*Cin/64 main slabs the producer streams
// the Csc/64 slabs of the (two-source, never concatenated) block input and the MMAs accumulate them into a SECOND
// accumulator at columns [BN, 2BN), so conv2 and conv_shortcut keep their separately rounded fp16 outputs. With two
// accumulators per tile there is one TMEM stage instead of two; these tiles have >= 99 K slabs, so the exposed
// epilogue is a few percent.
template <int BN, int STAGES, bool GEGLU, int EPI, int NP = 1, bool SC = false>
__global__ void __launch_bounds__(320, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmOut, const GemmParams p, int m_pairs,
             const __grid_constant__ CUtensorMap tmS0, const __grid_constant__ CUtensorMap tmS1,
             const __grid_constant__ CUtensorMap tmBs) {
  using L = Smem2<BN, STAGES>;
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
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1;          // role inside the MMA pair (0 = leader)
  const uint32_t pr = crank >> 1;           // pair inside the cluster (always 0 when NP == 1)
  const int cluster_id = blockIdx.x / (2 * NP);
  const int n_clusters = gridDim.x / (2 * NP);
  const int n_super = (p.n_tiles + NP - 1) / NP;   // NP N-adjacent tiles per cluster step
  const int total_tiles = m_pairs * n_super;
  const int slabs = p.slabs_main + (SC ? p.slabs_sc : 0);
  constexpr uint32_t kTmemCols = 512;   // two accumulator stages of BN (<= 256) columns

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (SC) {
      tma_prefetch_desc(&tmS0);
      tma_prefetch_desc(&tmS1);
      tma_prefetch_desc(&tmBs);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), NP);   // one multicast commit per pair
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 16);   // 8 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  // Programmatic dependent launch: wait for the previous kernel BEFORE allocating tensor memory. A dependent CTA that
  // grabbed TMEM first and then waited could starve a primary CTA on the same SM that had signalled its dependents but
  // not yet allocated its own columns (library kernels signal at their very start): neither would ever proceed.
  pdl_wait();
  if (warp == 1) tmem_alloc_2cta(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();   // barriers of BOTH CTAs are initialised before any remote arrive / multicast commit / TMA
  tc_fence_after();
  pdl_launch_dependents();   // this cluster holds everything it needs: the next kernel may start its own set-up
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (each CTA loads its own halves) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters) {
        const int n_tile = (t % n_super) * NP + static_cast<int>(pr);
        const int m_tile = 2 * (t / n_super) + static_cast<int>(rank);
        const int n0 = n_tile * BN + static_cast<int>(rank) * (BN / 2);
        int b0 = 0, y0 = 0, x0 = 0;
        if (p.conv) {
          x0 = (m_tile % p.tiles_x) * p.bw;
          y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.bh;
          b0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.bb;
        }
        for (int s = 0; s < slabs; ++s, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * L::STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);   // both CTAs' bytes
          if (SC && s >= p.slabs_main) {
            const int ss = s - p.slabs_main;
            if (ss < p.sc_split)
              tma2_load_4d(a_dst, &tmS0, full_bar(stage), ss * BK, x0, y0, b0);
            else
              tma2_load_4d(a_dst, &tmS1, full_bar(stage), (ss - p.sc_split) * BK, x0, y0, b0);
            tma2_load_2d(b_dst, &tmBs, full_bar(stage), ss * BK, n0);
          } else if (p.conv) {
            const int tap = s / p.cin_slabs;
            const int c0 = (s - tap * p.cin_slabs) * BK;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            tma2_load_4d(a_dst, &tmA, full_bar(stage), c0, x0 * p.stride + dx, y0 * p.stride + dy, b0);
            tma2_load_2d(b_dst, &tmB, full_bar(stage), c0, tap * p.cout + n0);
          } else {
            if (NP == 2)   // my 64-row half of A slab `rank`, delivered to CTA (0, rank) and CTA (1, rank)
              tma2_load_2d_mc(a_dst + pr * (A_BYTES / 2), &tmA, full_bar(stage), s * BK, m_tile * BM + pr * (BM / 2),
                              static_cast<uint16_t>((1u << rank) | (1u << (2 + rank))));
            else
              tma2_load_2d(a_dst, &tmA, full_bar(stage), s * BK, m_tile * BM);
            tma2_load_2d(b_dst, &tmB, full_bar(stage), s * BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, BN, 0);
      uint32_t it = 0;
      int tile_iter = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters, ++tile_iter) {
        const int acc = SC ? 0 : (tile_iter & 1);
        const uint32_t acc_phase = SC ? (tile_iter & 1) : ((tile_iter >> 1) & 1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // both CTAs' epilogues drained this accumulator stage
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int s = 0; s < slabs; ++s, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_src = smem_base + stage * L::STAGE_BYTES;
          const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
          const bool sc_slab = SC && s >= p.slabs_main;
          const int s_local = sc_slab ? s - p.slabs_main : s;
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t a_desc = make_smem_desc_sw128(a_src + k * 32, 0, 1024);
            const uint64_t b_desc = make_smem_desc_sw128(b_src + k * 32, 0, 1024);
            tc_mma_f16_2cta(d + (sc_slab ? BN : 0), a_desc, b_desc, idesc, (s_local > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_2cta(empty_bar(stage), NP == 2 ? 0xF : 0x3);   // this pair is done with the slot (all CTAs hear it)
        }
        tc_commit_2cta(tfull_bar(acc), static_cast<uint16_t>(0x3u << (2 * pr)));   // wake this pair's epilogues
      }
    }
  } else {
    // ===================== epilogue (both CTAs) =====================
    // TMEM -> registers -> fused epilogue -> 64B-swizzled smem staging -> TMA store. Every warp owns a private ring of
    // staging buffers (one 32-row x 32-column buffer per chunk of the tile) and issues its own bulk stores, so the
    // epilogue needs no CTA-wide barrier and never waits for a store inside a tile: the buffers are only recycled at
    // the next tile, a full main loop later. Rows / columns outside the output are clipped by the tensor map.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;            // 0: even chunks, 1: odd chunks
    const int row = quarter * 32 + lane;
    constexpr int OUT_COLS = GEGLU ? BN / 2 : BN;
    constexpr int NCHUNK = OUT_COLS / 32;
    const uint32_t my_stage = smem_base + L::STORE_OFFSET + (half * 4 + quarter) * (L::RING * 2048);
    uint8_t* my_stage_gen = smem_raw + (my_stage - smem_u32(smem_raw));
    const int sw = (lane >> 1) & 3;
    // origin of this warp's 32 accumulator rows inside the conv tile box (rows are ordered x fastest, then y, then b)
    const int qx = p.conv ? (quarter * 32) % p.bw : 0;
    const int qy = p.conv ? ((quarter * 32) / p.bw) % p.bh : 0;
    const int qb = p.conv ? (quarter * 32) / (p.bw * p.bh) : 0;
    int tile_iter = 0;
    for (int t = cluster_id; t < total_tiles; t += n_clusters, ++tile_iter) {
      const int acc = SC ? 0 : (tile_iter & 1);
      const uint32_t acc_phase = SC ? (tile_iter & 1) : ((tile_iter >> 1) & 1);
      const int n_tile = (t % n_super) * NP + static_cast<int>(pr);
      const int m_tile = 2 * (t / n_super) + static_cast<int>(rank);
      long long out_row;
      int sample;
      map_row(p, m_tile, row, &out_row, &sample);
      int b0 = 0, y0 = 0, x0 = 0;
      if (p.conv) {
        x0 = (m_tile % p.tiles_x) * p.bw;
        y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.bh;
        b0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.bb;
      }
      const int out_n0 = n_tile * OUT_COLS;
      // The residual rows do not depend on the accumulator: fetch the first chunk's segment before waiting for the
      // main loop and every later one a chunk ahead, so their L2 latency never sits on the epilogue's critical path
      // (it was 45% of the epilogue's stall samples and doubled the time of the short-K out-projections).
      constexpr bool PRE_RES = ((EPI & EPI_RES) != 0) && ((EPI & EPI_RUNTIME) == 0) && !GEGLU;
      uint4 rcur[4], rnxt[4];
      auto res_fetch = [&](int c, uint4 (&dst)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ncol = out_n0 + c * 32 + g * 8;
          dst[g] = (c < NCHUNK && out_row >= 0 && ncol < p.N)
                       ? __ldg(reinterpret_cast<const uint4*>(p.residual + out_row * p.ld_res + ncol))
                       : make_uint4(0, 0, 0, 0);
        }
      };
      constexpr bool PRE_BIAS = ((EPI & EPI_BIAS) != 0) && ((EPI & EPI_RUNTIME) == 0) && !GEGLU;
      uint4 bcur[4], bnxt[4];
      auto bias_fetch = [&](int c, uint4 (&dst)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ncol = out_n0 + c * 32 + g * 8;
          dst[g] = (c < NCHUNK && ncol < p.N) ? __ldg(reinterpret_cast<const uint4*>(p.bias + ncol)) : make_uint4(0, 0, 0, 0);
        }
      };
      if (PRE_RES) res_fetch(half, rcur);
      if (PRE_BIAS) bias_fetch(half, bcur);
      if (lane == 0) tma_store_wait_read<0>();   // this warp's stores of the previous tile have drained the ring
      __syncwarp();
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c = half; c < NCHUNK; c += 2) {
        uint32_t pk[16];
        if (PRE_RES) res_fetch(c + 2, rnxt);
        if (PRE_BIAS) bias_fetch(c + 2, bnxt);
        epilogue_chunk<BN, GEGLU, EPI>(p, t_row, SC ? BN : 0, n_tile, out_row, sample, c, pk, PRE_RES ? rcur : nullptr,
                                       PRE_BIAS ? bcur : nullptr);
        if (PRE_RES) {
#pragma unroll
          for (int g = 0; g < 4; ++g) rcur[g] = rnxt[g];
        }
        if (PRE_BIAS) {
#pragma unroll
          for (int g = 0; g < 4; ++g) bcur[g] = bnxt[g];
        }
        const uint32_t slot = (c >> 1) * 2048;
        uint8_t* dst = my_stage_gen + slot + lane * 64;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(dst + ((q ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (p.conv)
            tma_store_4d(&tmOut, my_stage + slot, out_n0 + c * 32, x0 + qx, y0 + qy, b0 + qb);
          else
            tma_store_2d(&tmOut, my_stage + slot, out_n0 + c * 32, m_tile * BM + quarter * 32);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 2 * pr);   // pair leader's barrier, one arrive per warp
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still be reading our smem / signalling our barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct ScMaps {
  const CUtensorMap* s0;
  const CUtensorMap* s1;
  const CUtensorMap* bs;
};

template <int BN, int STAGES, bool GEGLU, int EPI, int NP = 1, bool SC = false>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const GemmParams& p,
                   int m_pairs, cudaStream_t stream, ScMaps sc = ScMaps{nullptr, nullptr, nullptr}) {
  using L = Smem2<BN, STAGES>;
  static_assert(!SC || 2 * BN <= 512, "two accumulators must fit the 512 TMEM columns");
  static_assert(L::TOTAL <= 232448, "shared memory of this variant exceeds the 227 KB a CTA may use");
  auto kern = gemm2_kernel<BN, STAGES, GEGLU, EPI, NP, SC>;
  static bool configured = false;
  static int max_clusters = kSMs / (2 * NP);
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    if (NP > 1) {
      // four-CTA clusters must sit inside one GPC: ask the driver how many fit on this part
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(kSMs / (2 * NP) * 2 * NP);
      q.blockDim = dim3(320);
      q.dynamicSmemBytes = L::TOTAL;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2 * NP;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      VTON_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &q));
      VTON_CHECK_ARG(n > 0, "gemm2: no %d-CTA cluster fits on this device", 2 * NP);
      if (n < max_clusters) max_clusters = n;
    }
    configured = true;
  }
  const int tiles = m_pairs * cdiv(p.n_tiles, NP);
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * NP * clusters);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * NP;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VTON_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, p, m_pairs, sc.s0 ? *sc.s0 : tmA, sc.s1 ? *sc.s1 : tmA,
                               sc.bs ? *sc.bs : tmB));
  count_launch();
  return kOk;
}

// bn in {128, 160, 192, 256}; weight-tile box = bn/2 rows

// np = MMA pairs per cluster: 2 expects tmA encoded with 64-row boxes (linear layers only, BN 256).
int gemm2_dispatch(int bn, bool geglu, const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, int m_tiles,
                   cudaStream_t stream, int np, const CUtensorMap* tmS0, const CUtensorMap* tmS1,
                   const CUtensorMap* tmBs) {
  p.n_tiles = cdiv(p.N, bn);
  const int m_pairs = cdiv(m_tiles, 2);
  // output tensor map: [rows, out columns] (linear) or [B,H,W,out columns] (conv), 32-column boxes, 64B swizzle
  CUtensorMap tmOut;
  const int out_cols = geglu ? p.N / 2 : p.N;
  if (p.conv) {
    uint64_t dims[4] = {static_cast<uint64_t>(out_cols), static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H),
                        static_cast<uint64_t>(p.B)};
    uint64_t strides[3] = {static_cast<uint64_t>(p.ld_out) * 2, static_cast<uint64_t>(p.W) * p.ld_out * 2,
                           static_cast<uint64_t>(p.H) * p.W * p.ld_out * 2};
    // a warp stores 32 consecutive tile rows = a rectangular sub-box of the (bw, bh, bb) pixel box
    const uint32_t sbw = p.bw < 32 ? p.bw : 32;
    const uint32_t sbh = static_cast<uint32_t>(p.bh) < 32 / sbw ? p.bh : 32 / sbw;
    const uint32_t sbb = 32 / (sbw * sbh);
    uint32_t box[4] = {32, sbw, sbh, sbb};
    if (int e = encode_tmap_f16(&tmOut, p.out, 4, dims, strides, box, nullptr, 64)) return e;
  } else {
    uint64_t dims[2] = {static_cast<uint64_t>(out_cols), static_cast<uint64_t>(p.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(p.ld_out) * 2};
    uint32_t box[2] = {32, 32};
    if (int e = encode_tmap_f16(&tmOut, p.out, 2, dims, strides, box, nullptr, 64)) return e;
  }
  // epilogue specialisation: the four combinations the UNets use get their own kernels, anything else is generic
  int epi = (p.bias ? EPI_BIAS : 0) | (p.rowvec ? EPI_ROWVEC : 0) | (p.residual ? EPI_RES : 0);
  if (p.act_gelu || !(epi == 0 || epi == EPI_BIAS || epi == (EPI_BIAS | EPI_ROWVEC) || epi == (EPI_BIAS | EPI_RES)))
    epi = EPI_RUNTIME;
  if (p.slabs_sc) {   // conv + fused 1x1 shortcut: runtime epilogue, one TMEM stage, weight-tile box of tmBs = bn/2 rows
    VTON_CHECK_ARG(p.conv && !geglu && np == 1 && tmS0 && tmS1 && tmBs, "gemm2: shortcut slabs need conv mode and their maps");
    const ScMaps sc{tmS0, tmS1, tmBs};
    switch (bn) {
      case 128: return launch2<128, 7, false, EPI_RUNTIME, 1, true>(tmA, tmB, tmOut, p, m_pairs, stream, sc);
      case 160: return launch2<160, 6, false, EPI_RUNTIME, 1, true>(tmA, tmB, tmOut, p, m_pairs, stream, sc);
      case 256: return launch2<256, 5, false, EPI_RUNTIME, 1, true>(tmA, tmB, tmOut, p, m_pairs, stream, sc);
    }
    set_last_error("gemm2: shortcut variant supports BN 128/160/256 (got %d)", bn);
    return kErrUnsupported;
  }
  if (np == 2) {
    VTON_CHECK_ARG(!p.conv && bn == 256, "gemm2: the four-CTA cluster variant covers linear layers with BN 256 only");
    if (geglu) return launch2<256, 5, true, 0, 2>(tmA, tmB, tmOut, p, m_pairs, stream);
    switch (epi) {
      case 0: return launch2<256, 5, false, 0, 2>(tmA, tmB, tmOut, p, m_pairs, stream);
      case EPI_BIAS: return launch2<256, 5, false, EPI_BIAS, 2>(tmA, tmB, tmOut, p, m_pairs, stream);
      case EPI_BIAS | EPI_RES: return launch2<256, 5, false, EPI_BIAS | EPI_RES, 2>(tmA, tmB, tmOut, p, m_pairs, stream);
      default: return launch2<256, 5, false, EPI_RUNTIME, 2>(tmA, tmB, tmOut, p, m_pairs, stream);
    }
  }
  if (geglu) {
    if (bn == 256) return launch2<256, 5, true, 0>(tmA, tmB, tmOut, p, m_pairs, stream);
    if (bn == 128) return launch2<128, 7, true, 0>(tmA, tmB, tmOut, p, m_pairs, stream);
    set_last_error("gemm2: GEGLU epilogue supports BN 128/256 only (got %d)", bn);
    return kErrUnsupported;
  }
#define VTON_G2(BN_, ST_)                                                                                  \
  switch (epi) {                                                                                           \
    case 0: return launch2<BN_, ST_, false, 0>(tmA, tmB, tmOut, p, m_pairs, stream);                        \
    case EPI_BIAS: return launch2<BN_, ST_, false, EPI_BIAS>(tmA, tmB, tmOut, p, m_pairs, stream);           \
    case EPI_BIAS | EPI_ROWVEC:                                                                              \
      return launch2<BN_, ST_, false, EPI_BIAS | EPI_ROWVEC>(tmA, tmB, tmOut, p, m_pairs, stream);           \
    case EPI_BIAS | EPI_RES:                                                                                 \
      return launch2<BN_, ST_, false, EPI_BIAS | EPI_RES>(tmA, tmB, tmOut, p, m_pairs, stream);              \
    default: return launch2<BN_, ST_, false, EPI_RUNTIME>(tmA, tmB, tmOut, p, m_pairs, stream);              \
  }
  switch (bn) {
    case 128: VTON_G2(128, 7)
    case 160: VTON_G2(160, 6)
    case 192: VTON_G2(192, 6)
    case 256: VTON_G2(256, 5)
  }
#undef VTON_G2
  set_last_error("gemm2: unsupported BN %d", bn);
  return kErrUnsupported;
}

}  // namespace vton
