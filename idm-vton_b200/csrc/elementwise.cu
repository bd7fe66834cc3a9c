// HBM-/latency-bound glue kernels of the denoise step (sm_100a): layout seams, samplers' data movement,
// timestep embeddings, skinny (M <= 16) linears, and the fused CFG + DDPM update.
// Reference call sites are cited per kernel; rounding points follow the fp16-autocast path.
#include "common.cuh"
#include "host.h"

namespace vton {

// ------------------------------------------------------------------------------------------------
// NCHW <-> NHWC seams. The reference's UNet API is NCHW (src/unet_hacked_tryon.py:1006); the engine is NHWC.
// scatter: dst[s, y, x, c_off + c] = src[s % Bs, c, y, x]   (CFG duplication `torch.cat([latents]*2)`,
//          src/tryon_pipeline.py:1769, and the 13-channel concat, :1777, become channel offsets)
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const __half* src, int Bs, int Cs, int HW, __half* dst, int Bd, int ldc, int c_off) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(Bd) * HW;
  if (i >= total) return;
  const int s = static_cast<int>(i / HW);
  const int px = static_cast<int>(i % HW);
  const int sb = s % Bs;
  for (int c = 0; c < Cs; ++c) dst[i * ldc + c_off + c] = src[(static_cast<long long>(sb) * Cs + c) * HW + px];
}

__global__ void nhwc_to_nchw_kernel(const __half* src, int B, int C, int HW, int ldc, __half* dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * C * HW;
  if (i >= total) return;
  const int px = static_cast<int>(i % HW);
  const int c = static_cast<int>((i / HW) % C);
  const int b = static_cast<int>(i / (static_cast<long long>(HW) * C));
  dst[i] = src[(static_cast<long long>(b) * HW + px) * ldc + c];
}

int nchw_to_nhwc_impl(const void* src, int Bs, int Cs, int H, int W, void* dst, int Bd, int ldc, int c_off,
                      cudaStream_t stream) {
  VTON_CHECK_ARG(Bs > 0 && Cs > 0 && H > 0 && W > 0 && Bd > 0 && c_off + Cs <= ldc, "nchw_to_nhwc: bad shape");
  const long long total = static_cast<long long>(Bd) * H * W;
  nchw_to_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const __half*>(src), Bs, Cs, H * W, static_cast<__half*>(dst), Bd, ldc, c_off);
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

int nhwc_to_nchw_impl(const void* src, int B, int C, int H, int W, int ldc, void* dst, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && C <= ldc, "nhwc_to_nchw: bad shape");
  const long long total = static_cast<long long>(B) * C * H * W;
  nhwc_to_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const __half*>(src), B, C, H * W, ldc, static_cast<__half*>(dst));
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Upsample2D data movement: nearest x2 (diffusers Upsample2D = F.interpolate(scale 2, nearest) then conv3x3;
// built at src/unet_block_hacked_tryon.py:2301,2443). NHWC, 16-byte vectors.
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* src, int B, int H, int W, int V, uint4* dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * 4 * H * W * V;
  if (i >= total) return;
  const int v = static_cast<int>(i % V);
  long long t = i / V;
  const int x = static_cast<int>(t % (2 * W));
  t /= 2 * W;
  const int y = static_cast<int>(t % (2 * H));
  const int b = static_cast<int>(t / (2 * H));
  dst[i] = src[((static_cast<long long>(b) * H + (y >> 1)) * W + (x >> 1)) * V + v];
}

int upsample2x_impl(const void* src, int B, int H, int W, int C, void* dst, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && H > 0 && W > 0 && C % 8 == 0, "upsample2x: bad shape");
  const long long total = static_cast<long long>(B) * 4 * H * W * (C / 8);
  upsample2x_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint4*>(src), B, H, W, C / 8, static_cast<uint4*>(dst));
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Downsample2D (conv3x3 stride 2 pad 1; src/unet_block_hacked_tryon.py:1113,1246): gather the strided patches into
// A[b*Ho*Wo, 9*C] (tap-major K) and run the plain GEMM against W[Cout, 9*C].
// ------------------------------------------------------------------------------------------------
__global__ void im2col_s2_kernel(const uint4* src, int B, int H, int W, int V, int Ho, int Wo, uint4* dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * Ho * Wo * 9 * V;
  if (i >= total) return;
  const int v = static_cast<int>(i % V);
  long long t = i / V;
  const int tap = static_cast<int>(t % 9);
  t /= 9;
  const int ox = static_cast<int>(t % Wo);
  t /= Wo;
  const int oy = static_cast<int>(t % Ho);
  const int b = static_cast<int>(t / Ho);
  const int iy = 2 * oy + tap / 3 - 1;
  const int ix = 2 * ox + tap % 3 - 1;
  uint4 val = make_uint4(0, 0, 0, 0);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = src[((static_cast<long long>(b) * H + iy) * W + ix) * V + v];
  dst[i] = val;
}

int im2col_s2_impl(const void* src, int B, int H, int W, int C, void* dst, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && H > 0 && W > 0 && C % 8 == 0, "im2col_s2: bad shape");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * 9 * (C / 8);
  im2col_s2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint4*>(src), B, H, W, C / 8, Ho, Wo, static_cast<uint4*>(dst));
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Sinusoidal timestep embedding (diffusers Timesteps, flip_sin_to_cos=True, freq_shift=0): out = [cos | sin],
// fp32 math, fp16 store (`t_emb.to(dtype=sample.dtype)`, src/unet_hacked_tryon.py:1134-1139).
// values: [n] floats on device; out: [n, dim]
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* values, int n, int dim, __half* out, int rows_repeat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half_dim = dim / 2;
  if (i >= n * rows_repeat * half_dim) return;
  const int k = i % half_dim;
  const int r = i / half_dim;
  const float t = values[r % n];
  const float freq = expf(-logf(10000.0f) * static_cast<float>(k) / static_cast<float>(half_dim));
  const float arg = t * freq;
  out[static_cast<long long>(r) * dim + k] = f2h(cosf(arg));
  out[static_cast<long long>(r) * dim + half_dim + k] = f2h(sinf(arg));
}

int timestep_embed_impl(const void* values, int n, int dim, int rows_repeat, void* out, cudaStream_t stream) {
  VTON_CHECK_ARG(n > 0 && dim > 0 && dim % 2 == 0 && rows_repeat > 0, "timestep_embed: bad shape");
  const int total = n * rows_repeat * (dim / 2);
  timestep_embed_kernel<<<(total + 127) / 128, 128, 0, stream>>>(static_cast<const float*>(values), n, dim,
                                                                  static_cast<__half*>(out), rows_repeat);
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Skinny linear for the embedding MLPs (M <= 16 rows): TimestepEmbedding linear_1/linear_2, add_embedding, and all
// per-resnet time_emb_proj batched into one call. One warp per output column; weights are read once.
//   x' = in_silu ? fp16(silu(x)) : x ; y = fp16(W x' + b) ; y = out_silu ? fp16(silu(y)) : y ; y = addend ? fp16(y + addend) : y
// ------------------------------------------------------------------------------------------------
constexpr int SK_MAXM = 16;

__global__ void __launch_bounds__(256)
skinny_linear_kernel(const __half* x, int ldx, int M, int K, const __half* W, long long ldw, int N, const __half* bias,
                     int in_silu, int out_silu, const __half* addend, int ld_add, __half* out, int ldo) {
  extern __shared__ __half xs[];  // [M][K] activated input
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    __half v = x[static_cast<long long>(m) * ldx + k];
    if (in_silu) v = f2h(silu_f(h2f(v)));
    xs[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  float acc[SK_MAXM];
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) acc[m] = 0.f;
  const __half* wrow = W + static_cast<long long>(n) * ldw;
  for (int k = lane * 8; k < K; k += 256) {
    const uint4 wu = *reinterpret_cast<const uint4*>(wrow + k);
    const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w};
    float wf[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_h2(ww[j]);
      wf[2 * j] = f.x;
      wf[2 * j + 1] = f.y;
    }
#pragma unroll
    for (int m = 0; m < SK_MAXM; ++m) {
      if (m < M) {
        const uint4 xu = *reinterpret_cast<const uint4*>(xs + m * K + k);
        const uint32_t xw[4] = {xu.x, xu.y, xu.z, xu.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_h2(xw[j]);
          acc[m] += f.x * wf[2 * j] + f.y * wf[2 * j + 1];
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  }
  if (lane == 0) {
    const float b = bias ? h2f(bias[n]) : 0.f;
    for (int m = 0; m < M; ++m) {
      float y = round_h(acc[m] + b);
      if (out_silu) y = round_h(silu_f(y));
      if (addend) y = round_h(y + h2f(addend[static_cast<long long>(m) * ld_add + n]));
      out[static_cast<long long>(m) * ldo + n] = f2h(y);
    }
  }
}

int skinny_linear_impl(const void* x, int ldx, int M, int K, const void* W, long long ldw, int N, const void* bias,
                       int in_silu, int out_silu, const void* addend, int ld_add, void* out, int ldo,
                       cudaStream_t stream) {
  VTON_CHECK_ARG(M > 0 && M <= SK_MAXM, "skinny_linear: M=%d out of range (1..%d)", M, SK_MAXM);
  VTON_CHECK_ARG(K % 8 == 0 && ldw % 8 == 0 && N > 0, "skinny_linear: K/ldw must be multiples of 8");
  const size_t smem = static_cast<size_t>(M) * K * sizeof(__half);
  VTON_CHECK_ARG(smem <= 96 * 1024, "skinny_linear: M*K too large");
  static bool configured = false;
  if (!configured) {
    VTON_CUDA(cudaFuncSetAttribute(skinny_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  skinny_linear_kernel<<<cdiv(N, 8), 256, smem, stream>>>(
      static_cast<const __half*>(x), ldx, M, K, static_cast<const __half*>(W), ldw, N, static_cast<const __half*>(bias),
      in_silu, out_silu, static_cast<const __half*>(addend), ld_add, static_cast<__half*>(out), ldo);
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Fused classifier-free guidance + DDPM ancestral step (src/tryon_pipeline.py:1814-1823; diffusers DDPMScheduler.step,
// epsilon prediction, fixed_small variance). Every op rounds to fp16 like the reference's fp16 tensor arithmetic:
//   g = u + fp16(gs * fp16(c - u));  x0 = fp16(fp16(x - fp16(sb * g)) * inv_sa);
//   prev = fp16(fp16(c0 * x0) + fp16(c1 * x));  out = prev + fp16(sigma * noise)   (noise == null at t == 0)
// eps: NHWC [2B, HW, ldc] (uncond rows first, then cond); latents/noise/out: NCHW [B, 4, HW]. Coefficients live on the
// device ([6] floats: gs, sb, inv_sa, c0, c1, sigma) so a captured CUDA graph can be replayed for every step.
// ------------------------------------------------------------------------------------------------
__global__ void cfg_ddpm_kernel(const __half* eps, int ldc, int B, int C, int HW, const __half* latents,
                                const __half* noise, const float* coef, int do_cfg, __half* out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * C * HW;
  if (i >= total) return;
  const int px = static_cast<int>(i % HW);
  const int c = static_cast<int>((i / HW) % C);
  const int b = static_cast<int>(i / (static_cast<long long>(HW) * C));
  const float gs = coef[0], sb = coef[1], inv_sa = coef[2], c0 = coef[3], c1 = coef[4], sigma = coef[5];
  float g;
  if (do_cfg) {
    const float u = h2f(eps[(static_cast<long long>(b) * HW + px) * ldc + c]);
    const float t = h2f(eps[(static_cast<long long>(b + B) * HW + px) * ldc + c]);
    g = round_h(u + round_h(gs * round_h(t - u)));
  } else {
    g = h2f(eps[(static_cast<long long>(b) * HW + px) * ldc + c]);
  }
  const float x = h2f(latents[i]);
  const float x0 = round_h(round_h(x - round_h(sb * g)) * inv_sa);
  float prev = round_h(round_h(c0 * x0) + round_h(c1 * x));
  if (noise) prev = round_h(prev + round_h(sigma * h2f(noise[i])));
  out[i] = f2h(prev);
}

int cfg_ddpm_impl(const void* eps, int ldc, int B, int C, int H, int W, const void* latents, const void* noise,
                  const void* coef, int do_cfg, void* out, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && C > 0 && C <= ldc && H > 0 && W > 0 && coef, "cfg_ddpm: bad arguments");
  const long long total = static_cast<long long>(B) * C * H * W;
  cfg_ddpm_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const __half*>(eps), ldc, B, C, H * W, static_cast<const __half*>(latents),
      static_cast<const __half*>(noise), static_cast<const float*>(coef), do_cfg, static_cast<__half*>(out));
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Pre / post-processing around the VAE (diffusers VaeImageProcessor as the pipeline uses it,
// src/tryon_pipeline.py:418-421, 1588-1602, 940-955, 1885). One launch each instead of ~10 small ATen kernels.
//   preprocess: image [B,3,H,W] fp32 -> init_image = 2x-1 (skipped when the batch already has negative values: diffusers
//               checks `image.min() < 0`; the minimum arrives as a device scalar, so there is no host sync);
//               mask [B,Cm,H,W] (Cm = 1 | 3: grayscale 0.299/0.587/0.114) -> binarised at 0.5;
//               masked_image = init_image * (mask < 0.5);  mask_latent = nearest resize to [B,1,H/s,W/s]
//               (F.interpolate default: source index = floor(dst * s)).
//   postprocess: decoder output [B,3,H,W] fp32 (NCHW, or NHWC memory with `nhwc` set) -> (x/2 + 0.5).clamp(0,1) as
//               fp32 NCHW ("pt") and/or uint8 NHWC (round(x*255): what "np" -> "pil" produces).
// ------------------------------------------------------------------------------------------------
__global__ void preprocess_kernel(const float* image, const float* mask_in, int Cm, const float* img_min, int B, int H,
                                  int W, int s, float* init_image, float* mask_bin, float* masked_image,
                                  __half* mask_latent) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long HW = static_cast<long long>(H) * W;
  if (i >= B * HW) return;
  const int b = static_cast<int>(i / HW);
  const long long px = i % HW;
  const bool normalize = !(*img_min < 0.f);
  float m;
  if (Cm == 3) {
    const float* mp = mask_in + static_cast<long long>(b) * 3 * HW + px;
    // separately rounded products and sums, like the three ATen kernels of the torch expression (no FMA contraction)
    m = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, mp[0]), __fmul_rn(0.587f, mp[HW])), __fmul_rn(0.114f, mp[2 * HW]));
  } else {
    m = mask_in[static_cast<long long>(b) * HW + px];
  }
  const float mb = m >= 0.5f ? 1.f : 0.f;
  mask_bin[i] = mb;
  const float keep = mb < 0.5f ? 1.f : 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const long long o = (static_cast<long long>(b) * 3 + c) * HW + px;
    float v = image[o];
    if (normalize) v = 2.0f * v - 1.0f;
    init_image[o] = v;
    masked_image[o] = v * keep;
  }
  const int y = static_cast<int>(px / W), x = static_cast<int>(px % W);
  if (y % s == 0 && x % s == 0 && y / s < H / s && x / s < W / s)
    mask_latent[(static_cast<long long>(b) * (H / s) + y / s) * (W / s) + x / s] = f2h(mb);
}

int preprocess_impl(const void* image, const void* mask, int Cm, const void* img_min, int B, int H, int W, int scale,
                    void* init_image, void* mask_bin, void* masked_image, void* mask_latent, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && H > 0 && W > 0 && (Cm == 1 || Cm == 3) && scale > 0 && H % scale == 0 && W % scale == 0,
                 "preprocess: bad shape B=%d H=%d W=%d Cm=%d scale=%d", B, H, W, Cm, scale);
  VTON_CHECK_ARG(image && mask && img_min && init_image && mask_bin && masked_image && mask_latent, "preprocess: null pointer");
  const long long total = static_cast<long long>(B) * H * W;
  preprocess_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const float*>(image), static_cast<const float*>(mask), Cm, static_cast<const float*>(img_min), B, H, W,
      scale, static_cast<float*>(init_image), static_cast<float*>(mask_bin), static_cast<float*>(masked_image),
      static_cast<__half*>(mask_latent));
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

__global__ void postprocess_kernel(const float* x, int nhwc, int B, int H, int W, float* out_pt, uint8_t* out_u8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long HW = static_cast<long long>(H) * W;
  if (i >= B * HW) return;
  const int b = static_cast<int>(i / HW);
  const long long px = i % HW;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = nhwc ? x[i * 3 + c] : x[(static_cast<long long>(b) * 3 + c) * HW + px];
    const float y = fminf(fmaxf(v / 2.f + 0.5f, 0.f), 1.f);
    if (out_pt) out_pt[(static_cast<long long>(b) * 3 + c) * HW + px] = y;
    if (out_u8) out_u8[i * 3 + c] = static_cast<uint8_t>(rintf(y * 255.f));
  }
}

int postprocess_impl(const void* x, int nhwc, int B, int H, int W, void* out_pt, void* out_u8, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && H > 0 && W > 0 && x && (out_pt || out_u8), "postprocess: bad arguments");
  const long long total = static_cast<long long>(B) * H * W;
  postprocess_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const float*>(x), nhwc, B, H, W, static_cast<float*>(out_pt), static_cast<uint8_t*>(out_u8));
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// CLIP embedding front ends (transformers CLIPVisionEmbeddings / CLIPTextEmbeddings, called from
// src/tryon_pipeline.py:468-470 and :592-596).
// patchify: the stride-P patch convolution of the ViT becomes a GEMM over rows [b, gy, gx] with K = (c, ky, kx) — the
//           order of the conv weight [Cout, 3, P, P] flattened — zero-padded to a multiple of 64 columns.
// token_embed: out[r] = fp16(token_embedding[ids[r]] + position_embedding[r % T]).
// ------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const __half* x, int C, int Hi, int Wi, int P, int gh, int gw, int K, int ldk, __half* out,
                                long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int col = static_cast<int>(i % ldk);
  const long long row = i / ldk;
  __half v = __float2half(0.f);
  if (col < K) {
    const int kx = col % P, ky = (col / P) % P, c = col / (P * P);
    const int gx = static_cast<int>(row % gw), gy = static_cast<int>((row / gw) % gh);
    const long long b = row / (static_cast<long long>(gw) * gh);
    v = x[((b * C + c) * Hi + gy * P + ky) * Wi + gx * P + kx];
  }
  out[i] = v;
}

int patchify_impl(const void* x, int B, int C, int Hi, int Wi, int P, void* out, int ldk, cudaStream_t stream) {
  VTON_CHECK_ARG(B > 0 && C > 0 && P > 0 && Hi >= P && Wi >= P, "patchify: bad shape");
  const int gh = Hi / P, gw = Wi / P, K = C * P * P;
  VTON_CHECK_ARG(ldk >= K, "patchify: row stride %d < K %d", ldk, K);
  const long long total = static_cast<long long>(B) * gh * gw * ldk;
  patchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const __half*>(x), C, Hi, Wi, P, gh, gw, K, ldk, static_cast<__half*>(out), total);
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

__global__ void token_embed_kernel(const long long* ids, int rows, int T, int V, int vocab, const uint4* tok, const uint4* pos,
                                   uint4* out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * V) return;
  const int r = static_cast<int>(i / V), c = static_cast<int>(i % V);
  long long id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4 a = tok[id * V + c];
  const uint4 b = pos[static_cast<long long>(r % T) * V + c];
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fa = unpack_h2(aw[j]), fb = unpack_h2(bw[j]);
    o[j] = pack_h2(fa.x + fb.x, fa.y + fb.y);
  }
  out[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

int token_embed_impl(const void* ids, int rows, int T, int C, int vocab, const void* tok, const void* pos, void* out,
                     cudaStream_t stream) {
  VTON_CHECK_ARG(rows > 0 && T > 0 && C > 0 && C % 8 == 0 && vocab > 0, "token_embed: bad shape rows=%d T=%d C=%d", rows, T, C);
  const int V = C / 8;
  const long long total = static_cast<long long>(rows) * V;
  token_embed_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const long long*>(ids), rows, T, V, vocab, static_cast<const uint4*>(tok), static_cast<const uint4*>(pos),
      static_cast<uint4*>(out));
  count_launch();
  VTON_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vton
