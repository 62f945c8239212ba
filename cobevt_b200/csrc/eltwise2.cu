// Round-2 HBM-bound kernels around the FAX path: uint8 camera-frame staging, the
// softmax / argmax post-process (row f2), nearest feature down-scaling (PyramidAxialEncoder scale < 1) and the
// MBConv pieces of the EfficientNet-B4 prefix (row a16): depthwise conv + BN + swish with the squeeze sums,
// the squeeze-excite gate, the gate application.  Reference lines are cited per entry point in
// include/cobevt_b200.h.
#include "common.h"
#include "sm100.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ float2 up2(uint32_t u, int bf16) {
  if (bf16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ uint32_t pk2(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void up8(const uint4& u, int bf16, float* f) {
  const float2 a = up2(u.x, bf16), b = up2(u.y, bf16), c = up2(u.z, bf16), d = up2(u.w, bf16);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pk8(const float* f, int bf16) {
  uint4 u;
  u.x = pk2(f[0], f[1], bf16); u.y = pk2(f[2], f[3], bf16); u.z = pk2(f[4], f[5], bf16); u.w = pk2(f[6], f[7], bf16);
  return u;
}
__device__ __forceinline__ float swish(float x) { return x / (1.0f + __expf(-x)); }

inline unsigned blocks_for(int64_t threads, int bs) { return (unsigned)((threads + bs - 1) / bs); }

// ------------------------------------------------------------------ uint8 frame staging
// (N,H,W,3) uint8 -> (N,Hp,Wp,4) 16-bit RGBx with a zero border: ((x / 255) - mean) / std per channel, optional
// BGR -> RGB swap.  One thread per output pixel: 3 byte loads (the 3-byte pixels of a warp are one 96-byte run),
// one 8-byte store.
__global__ void stage_rgb_u8_kernel(const uint8_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W, int Hp,
                                    int Wp, int pt, int pl, int bgr, float3 sc, float3 sh, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * Hp * Wp) return;
  const int xp = (int)(i % Wp);
  const int yp = (int)((i / Wp) % Hp);
  const int n = (int)(i / ((int64_t)Wp * Hp));
  const int ys = yp - pt, xs = xp - pl;
  uint2 u = make_uint2(0u, 0u);
  if (ys >= 0 && ys < H && xs >= 0 && xs < W) {
    const uint8_t* s = x + (((int64_t)n * H + ys) * W + xs) * 3;
    float c0 = (float)s[0], c1 = (float)s[1], c2 = (float)s[2];
    if (bgr) { const float t = c0; c0 = c2; c2 = t; }
    u.x = pk2(fmaf(c0, sc.x, sh.x), fmaf(c1, sc.y, sh.y), bf16);
    u.y = pk2(fmaf(c2, sc.z, sh.z), 0.f, bf16);
  }
  *reinterpret_cast<uint2*>(y + i * 4) = u;
}

// ------------------------------------------------------------------ softmax + argmax over the class axis (NCHW fp32)
__global__ void softmax_argmax_kernel(const float* __restrict__ seg, float* __restrict__ prob, int64_t* __restrict__ amap,
                                      int N, int C, int64_t HW) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * HW) return;
  const int64_t n = i / HW, p = i - n * HW;
  const float* s = seg + n * C * HW + p;
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, s[c * HW]);
  float sum = 0.f;
  for (int c = 0; c < C; ++c) sum += expf(s[c * HW] - mx);
  const float inv = 1.0f / sum;
  float best = -1.f;
  int arg = 0;
  float* o = prob + n * C * HW + p;
  for (int c = 0; c < C; ++c) {
    const float pr = expf(s[c * HW] - mx) * inv;
    o[c * HW] = pr;
    if (pr > best) { best = pr; arg = c; }     // first maximum wins, like torch.argmax
  }
  amap[i] = arg;
}

// ------------------------------------------------------------------ nearest resize (F.interpolate default mode)
__global__ void resize_nearest_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W, int C,
                                      int OH, int OW, float sy, float sx) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int cg = C / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * OH * OW * cg) return;
  const int c8 = (int)(idx % cg);
  int64_t t = idx / cg;
  const int ox = (int)(t % OW); t /= OW;
  const int oy = (int)(t % OH);
  const int n = (int)(t / OH);
  const int iy = min((int)floorf((float)oy * sy), H - 1), ix = min((int)floorf((float)ox * sx), W - 1);
  *reinterpret_cast<uint4*>(y + idx * 8) = *reinterpret_cast<const uint4*>(x + (((int64_t)n * H + iy) * W + ix) * C + c8 * 8);
}

// ------------------------------------------------------------------ depthwise KxK conv + bias (folded BN) + swish
// NHWC 16-bit, 8 channels (one 16-byte vector) per thread, padding (pt, pl) before / implied after ("static same"
// padding of efficientnet_pytorch is asymmetric).  Optionally accumulates the per-(image, channel) sums of the
// OUTPUT for the squeeze step: each thread owns ROWS consecutive output pixels of one channel group, block-level
// reduction through shared memory, then one atomicAdd per (block, channel).
template <int K>
__global__ void dwconv_kernel(const uint16_t* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                              uint16_t* __restrict__ y, float* __restrict__ sums, int N, int H, int W, int C, int OH, int OW,
                              int stride, int pt, int pl, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  // grid: (ceil(OH*OW / blockDim.y), C/8 groups folded into threadIdx.x, N)
  const int cg = C / 8;
  const int g = blockIdx.y * blockDim.x + threadIdx.x;          // channel group
  const int n = blockIdx.z;
  const int pix = blockIdx.x * blockDim.y + threadIdx.y;
  float acc[8];
  float part[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) part[j] = 0.f;
  if (g < cg && pix < OH * OW) {
    const int oy = pix / OW, ox = pix - oy * OW;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias[g * 8 + j];
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy * stride - pt + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < K; ++kx) {
        const int ix = ox * stride - pl + kx;
        if (ix < 0 || ix >= W) continue;
        float f[8];
        up8(*reinterpret_cast<const uint4*>(x + (((int64_t)n * H + iy) * W + ix) * C + g * 8), bf16, f);
        const float4 w0 = *reinterpret_cast<const float4*>(w + ((ky * K + kx) * C + g * 8));
        const float4 w1 = *reinterpret_cast<const float4*>(w + ((ky * K + kx) * C + g * 8 + 4));
        acc[0] = fmaf(f[0], w0.x, acc[0]); acc[1] = fmaf(f[1], w0.y, acc[1]);
        acc[2] = fmaf(f[2], w0.z, acc[2]); acc[3] = fmaf(f[3], w0.w, acc[3]);
        acc[4] = fmaf(f[4], w1.x, acc[4]); acc[5] = fmaf(f[5], w1.y, acc[5]);
        acc[6] = fmaf(f[6], w1.z, acc[6]); acc[7] = fmaf(f[7], w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = swish(acc[j]);
    const uint4 o = pk8(acc, bf16);
    *reinterpret_cast<uint4*>(y + (((int64_t)n * OH + oy) * OW + ox) * C + g * 8) = o;
    if (sums) up8(o, bf16, part);       // the squeeze averages the STORED (16-bit) activations
  }
  if (sums) {
    extern __shared__ float red[];       // [blockDim.y][blockDim.x * 8]
    float* mine = red + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = part[j];
    __syncthreads();
    if (threadIdx.y == 0 && g < cg) {
      for (int r = 1; r < blockDim.y; ++r) {
        const float* o = red + ((size_t)r * blockDim.x + threadIdx.x) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) part[j] += o[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(sums + (int64_t)n * C + g * 8 + j, part[j]);
    }
  }
}

// ------------------------------------------------------------------ squeeze-excite gate
// gate[n, c] = sigmoid(W2 swish(W1 mean[n] + b1) + b2); one block per image, Cse <= 64 hidden units in smem.
__global__ void se_gate_kernel(const float* __restrict__ sums, float inv_hw, const float* __restrict__ w1,
                               const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                               float* __restrict__ gate, int C, int Cse) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  extern __shared__ float sh[];          // mean[C] | hidden[Cse]
  float* mean = sh;
  float* hid = sh + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mean[c] = sums[(int64_t)n * C + c] * inv_hw;
  __syncthreads();
  for (int j = threadIdx.x; j < Cse; j += blockDim.x) {
    float a = b1[j];
    for (int c = 0; c < C; ++c) a = fmaf(w1[j * C + c], mean[c], a);
    hid[j] = swish(a);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = b2[c];
    for (int j = 0; j < Cse; ++j) a = fmaf(w2[c * Cse + j], hid[j], a);
    gate[(int64_t)n * C + c] = 1.0f / (1.0f + __expf(-a));
  }
}

// y[n, p, c] = x[n, p, c] * gate[n, c]   (in place allowed)
__global__ void scale_channels_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gate, uint16_t* __restrict__ y,
                                      int N, int64_t HW, int C, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int cg = C / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * HW * cg) return;
  const int c8 = (int)(idx % cg);
  const int64_t n = idx / (HW * cg);
  float f[8];
  up8(*reinterpret_cast<const uint4*>(x + idx * 8), bf16, f);
  const float4 g0 = *reinterpret_cast<const float4*>(gate + n * C + c8 * 8);
  const float4 g1 = *reinterpret_cast<const float4*>(gate + n * C + c8 * 8 + 4);
  f[0] *= g0.x; f[1] *= g0.y; f[2] *= g0.z; f[3] *= g0.w; f[4] *= g1.x; f[5] *= g1.y; f[6] *= g1.z; f[7] *= g1.w;
  *reinterpret_cast<uint4*>(y + idx * 8) = pk8(f, bf16);
}

// squeeze-excite gate + its application in one launch: every block recomputes the (tiny) gate of its image into shared
// memory -- C * Cse * 2 MACs against a launch saved per MBConv block -- then scales its slice of the map in place.
// grid (blocks per image, N); sums (N, C) as produced by dwconv_kernel.
__global__ void __launch_bounds__(256) se_scale_kernel(uint16_t* __restrict__ x, const float* __restrict__ sums, float inv_hw,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2, int64_t HW, int C,
                                                       int Cse, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  extern __shared__ float sh[];          // gate[C] | mean[C] | hidden[Cse]
  float* gate = sh;
  float* mean = sh + C;
  float* hid = sh + 2 * C;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mean[c] = sums[(int64_t)n * C + c] * inv_hw;
  __syncthreads();
  for (int j = threadIdx.x; j < Cse; j += blockDim.x) {
    float a = b1[j];
    for (int c = 0; c < C; ++c) a = fmaf(w1[j * C + c], mean[c], a);
    hid[j] = swish(a);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = b2[c];
    for (int j = 0; j < Cse; ++j) a = fmaf(w2[c * Cse + j], hid[j], a);
    gate[c] = 1.0f / (1.0f + __expf(-a));
  }
  __syncthreads();
  const int cg = C / 8;
  const int64_t nvec = HW * cg;
  uint16_t* xi = x + (int64_t)n * HW * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    float f[8];
    up8(*reinterpret_cast<const uint4*>(xi + i * 8), bf16, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= gate[c8 * 8 + j];
    *reinterpret_cast<uint4*>(xi + i * 8) = pk8(f, bf16);
  }
}

// ------------------------------------------------------------------ analytic BEV query projection (stage 0 of SinBEVT)
// q[b,n,p,:] = W' LN0(x[b,p,:] + e / (|e| + 1e-7)) + b',  e = a wx_p + b_ wy_p + c_{b,n}  (fax_modules.py:370-375,387-390
// and to_q :189-190, W' / b' = the projection with the LayerNorm affine folded in).  e is AFFINE in the cell's world
// coordinates, so W' e, |e|^2, mean(e) and mean(x e) are closed forms in (wx, wy) with per-channel / per-cell / per-view
// constants, and LayerNorm is (y - mu) rstd with mu, rstd scalars per token.  Hence
//   q = rstd (W'x)[b,p] + rstd inv (A wx + B wy + Cc_{b,n}) + (b' - rstd mu S)
// needs NO per-token reduction and NO GEMM over the (rows x views x cells) tokens: W'x is computed once per (row, cell) by
// the GEMM kernel, the per-cell statistics once per row map by row_stats9_kernel, and this kernel is a pure streaming
// write of q (the 503 MB LayerNorm output of bev_query_kernel and its re-read by the q-projection GEMM disappear).
// chan_tab: C x 20 fp32 = [a, b_, b_bev, w_cam(4), 0 | A, B, S, b' | Cb, M4(4), 0, 0, 0] with A = W'a, B = W'b_, S = W'1,
// Cb = W' b_bev, M4 = W' w_cam (five float4 per channel).  stats: rows x 12 fp32 = means over channels of
// [x, x^2, x a, x b_, x b_bev, x w_cam(4)].
template <int C>
__global__ void __launch_bounds__(256) row_stats9_kernel(const uint16_t* __restrict__ x, const float* __restrict__ tab,
                                                         float* __restrict__ stats, int64_t rows, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  constexpr int LPT = C / 8;                       // lanes per row (8 channels each)
  const int lane = threadIdx.x & 31, g = lane % LPT;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPT;
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  if (row < rows) {
    float f[8];
    up8(*reinterpret_cast<const uint4*>(x + row * C + g * 8), bf16, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 d0 = *reinterpret_cast<const float4*>(tab + (g * 8 + j) * 20);        // a, b_, b_bev, w_cam0
      const float4 d1 = *reinterpret_cast<const float4*>(tab + (g * 8 + j) * 20 + 4);    // w_cam1..3, 0
      const float v = f[j];
      acc[0] += v; acc[1] = fmaf(v, v, acc[1]);
      acc[2] = fmaf(v, d0.x, acc[2]); acc[3] = fmaf(v, d0.y, acc[3]); acc[4] = fmaf(v, d0.z, acc[4]);
      acc[5] = fmaf(v, d0.w, acc[5]); acc[6] = fmaf(v, d1.x, acc[6]); acc[7] = fmaf(v, d1.y, acc[7]); acc[8] = fmaf(v, d1.z, acc[8]);
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (row < rows && g == 0) {
    float* o = stats + row * 12;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = acc[k] * (1.0f / C);
  }
}

// One token = LPT lanes (8 channels each).  The per-(token, view) scalars (|e|, mean, variance) are computed by lane
// v of the token for view v -- all views at once instead of every lane repeating every view -- and broadcast with
// three shuffles; the channel math runs on packed fp32 pairs.  Measured 189 us for the 503 MB of the north_star stage 0
// (2.7 TB/s: pure-write streams on this part level off near 3 TB/s; a register-prefetch / 3-blocks-per-SM variant
// was slower, 199 us).
template <int C>
__global__ void __launch_bounds__(256) bev_q_kernel(const float* __restrict__ qx, int64_t x_rows, const float* __restrict__ stats,
                                                    const float* __restrict__ world, const float* __restrict__ tab,
                                                    const float* __restrict__ E_inv, uint16_t* __restrict__ out, int n, int P,
                                                    float qa, float qb, float qd, float ma, float mb, float eps, int bf16, int cpb) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  constexpr int LPT = C / 8, TPB = 256 / LPT;      // lanes per token, tokens per block iteration
  constexpr int MAXV = 8;
  __shared__ __align__(16) float sCc[MAXV][C];     // Cc_{b,v} = Cb - M4 t_v
  __shared__ __align__(16) float sV[MAXV][8];      // per view: qc, qe, qf, mc, t0..t3
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid % LPT, tok = tid / LPT;
  // ---- per-view constants of this agent row ------------------------------------------------------------
  for (int v = warp; v < MAXV; v += 8) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    if (v < n) {
      const float* Ei = E_inv + ((int64_t)b * n + v) * 16;
      t0 = Ei[3]; t1 = Ei[7]; t2 = Ei[11]; t3 = Ei[15];
    }
    float qc = 0.f, qe = 0.f, qf = 0.f, mc = 0.f;
    for (int ch = lane; ch < C; ch += 32) {
      const float4 d0 = *reinterpret_cast<const float4*>(tab + ch * 20), d1 = *reinterpret_cast<const float4*>(tab + ch * 20 + 4);
      const float4 c0 = *reinterpret_cast<const float4*>(tab + ch * 20 + 12), c1 = *reinterpret_cast<const float4*>(tab + ch * 20 + 16);
      const float c = d0.z - (d0.w * t0 + d1.x * t1 + d1.y * t2 + d1.z * t3);
      qc = fmaf(c, c, qc); qe = fmaf(d0.x, c, qe); qf = fmaf(d0.y, c, qf); mc += c;
      sCc[v][ch] = c0.x - (c0.y * t0 + c0.z * t1 + c0.w * t2 + c1.x * t3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      qc += __shfl_xor_sync(0xffffffffu, qc, o); qe += __shfl_xor_sync(0xffffffffu, qe, o);
      qf += __shfl_xor_sync(0xffffffffu, qf, o); mc += __shfl_xor_sync(0xffffffffu, mc, o);
    }
    if (lane == 0) {
      sV[v][0] = qc; sV[v][1] = 2.0f * qe; sV[v][2] = 2.0f * qf; sV[v][3] = mc * (1.0f / C);
      sV[v][4] = t0; sV[v][5] = t1; sV[v][6] = t2; sV[v][7] = t3;
    }
  }
  float A8[8], B8[8], S8[8], Q8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 q = *reinterpret_cast<const float4*>(tab + (g * 8 + j) * 20 + 8);
    A8[j] = q.x; B8[j] = q.y; S8[j] = q.z; Q8[j] = q.w;
  }
  __syncthreads();
  // this lane's view (for the scalar stage): lane g of a token handles view g (views >= LPT: see host check)
  const float4 va = *reinterpret_cast<const float4*>(sV[g < MAXV ? g : 0]), vb = *reinterpret_cast<const float4*>(sV[g < MAXV ? g : 0] + 4);
  const int64_t xrow0 = (x_rows == (int64_t)P) ? 0 : (int64_t)b * P;      // broadcast prior or per-row map
  const int base = blockIdx.x * cpb;
  constexpr float invC = 1.0f / C;
  const int lane_tok0 = lane - g;                                           // first lane of this token inside the warp
  for (int cc = tok; cc < cpb; cc += TPB) {
    const int cellr = base + cc;
    const bool live = cellr < P;
    const int cell = live ? cellr : P - 1;                                  // (shuffles below need every lane)
    const float wx = world[cell], wy = world[P + cell];
    const float* st = stats + (xrow0 + cell) * 12;
    const float4 s0 = *reinterpret_cast<const float4*>(st), s1 = *reinterpret_cast<const float4*>(st + 4);
    const float s8 = st[8];
    const float* qp = qx + (xrow0 + cell) * C + g * 8;
    const float4 q0 = *reinterpret_cast<const float4*>(qp), q1 = *reinterpret_cast<const float4*>(qp + 4);
    const float xq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float lin[8];                                                          // A wx + B wy (view independent)
#pragma unroll
    for (int j = 0; j < 8; ++j) lin[j] = fmaf(A8[j], wx, B8[j] * wy);
    // ---- scalars of (cell, view g), by lane g --------------------------------------------------------------
    float m_rstd, m_k1, m_k2;
    {
      const float ss = fmaxf(fmaf(wx, fmaf(qa, wx, fmaf(qd, wy, va.y)), fmaf(wy, fmaf(qb, wy, va.z), va.x)), 0.f);
      const float inv = __fdividef(1.0f, sqrtf(ss) + 1e-7f);
      const float mu = fmaf(inv, fmaf(ma, wx, fmaf(mb, wy, va.w)), s0.x);
      const float xe = fmaf(s0.z, wx, fmaf(s0.w, wy, s1.x)) - (s1.y * vb.x + s1.z * vb.y + s1.w * vb.z + s8 * vb.w);
      const float ey2 = fmaf(inv * inv, ss * invC, fmaf(2.0f * inv, xe, s0.y));
      m_rstd = rsqrtf(fmaxf(fmaf(-mu, mu, ey2), 0.f) + eps);
      m_k1 = m_rstd * inv;
      m_k2 = -m_rstd * mu;
    }
    for (int v = 0; v < n; ++v) {
      const float rstd = __shfl_sync(0xffffffffu, m_rstd, lane_tok0 + v);
      const float k1 = __shfl_sync(0xffffffffu, m_k1, lane_tok0 + v);
      const float k2 = __shfl_sync(0xffffffffu, m_k2, lane_tok0 + v);
      const float4 c0 = *reinterpret_cast<const float4*>(&sCc[v][g * 8]), c1 = *reinterpret_cast<const float4*>(&sCc[v][g * 8 + 4]);
      const float cc8[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float e0, e1, b0, b1;
        sm100::fadd2(e0, e1, lin[j], lin[j + 1], cc8[j], cc8[j + 1]);
        sm100::ffma2(b0, b1, S8[j], S8[j + 1], k2, k2, Q8[j], Q8[j + 1]);
        sm100::ffma2(b0, b1, e0, e1, k1, k1, b0, b1);
        sm100::ffma2(o[j], o[j + 1], xq[j], xq[j + 1], rstd, rstd, b0, b1);
      }
      if (live) *reinterpret_cast<uint4*>(out + (((int64_t)b * n + v) * P + cell) * C + g * 8) = pk8(o, bf16);
    }
  }
}

}  // namespace

using cobevt::fail;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int cobevt_stage_rgb_u8(const uint8_t* x, void* y, int32_t N, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                                   int32_t pad_top, int32_t pad_left, int32_t bgr, const float* mean3, const float* std3,
                                   int32_t dtype, void* stream) {
  if (!x || !y || N <= 0 || Hp < H + pad_top || Wp < W + pad_left) return fail(COBEVT_ERR_SHAPE, "stage_rgb_u8: bad args");
  float3 sc = make_float3(1.f / 255.f, 1.f / 255.f, 1.f / 255.f), sh = make_float3(0.f, 0.f, 0.f);
  if (mean3 && std3) {       // ((x / 255) - mean) / std  =  x * (1 / (255 std)) - mean / std
    sc = make_float3(1.f / (255.f * std3[0]), 1.f / (255.f * std3[1]), 1.f / (255.f * std3[2]));
    sh = make_float3(-mean3[0] / std3[0], -mean3[1] / std3[1], -mean3[2] / std3[2]);
  }
  const int64_t total = (int64_t)N * Hp * Wp;
  cobevt::launch_k(stage_rgb_u8_kernel, blocks_for(total, 256), 256, 0, ST(stream), x, (uint16_t*)y, N, H, W, Hp, Wp, pad_top, pad_left, bgr, sc, sh, dtype);
  return cobevt::check_launch("stage_rgb_u8");
}

extern "C" int cobevt_softmax_argmax(const float* seg, float* prob, int64_t* amap, int32_t N, int32_t C, int64_t HW, void* stream) {
  if (!seg || !prob || !amap || N <= 0 || C <= 0 || HW <= 0) return fail(COBEVT_ERR_SHAPE, "softmax_argmax: bad args");
  cobevt::launch_k(softmax_argmax_kernel, blocks_for((int64_t)N * HW, 256), 256, 0, ST(stream), seg, prob, amap, N, C, HW);
  return cobevt::check_launch("softmax_argmax");
}

extern "C" int cobevt_resize_nearest(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t OH, int32_t OW,
                                     float src_per_dst_h, float src_per_dst_w, void* stream) {
  if (!x || !y || C % 8 || OH <= 0 || OW <= 0) return fail(COBEVT_ERR_SHAPE, "resize_nearest: bad args");
  cobevt::launch_k(resize_nearest_kernel, blocks_for((int64_t)N * OH * OW * (C / 8), 256), 256, 0, ST(stream), 
      (const uint16_t*)x, (uint16_t*)y, N, H, W, C, OH, OW, src_per_dst_h, src_per_dst_w);
  return cobevt::check_launch("resize_nearest");
}

extern "C" int cobevt_dwconv_swish(const void* x, const float* w, const float* bias, void* y, float* sums, int32_t N, int32_t H,
                                   int32_t W, int32_t C, int32_t K, int32_t stride, int32_t pad_top, int32_t pad_left, int32_t OH,
                                   int32_t OW, int32_t dtype, void* stream) {
  if (!x || !w || !bias || !y || C % 8 || (K != 3 && K != 5) || stride < 1 || N <= 0 || N > 65535)
    return fail(COBEVT_ERR_SHAPE, "dwconv_swish: bad args (C %% 8, K in {3,5})");
  const int cg = C / 8;
  const int bx = cg >= 32 ? 32 : cg;                 // channel groups per block (adjacent lanes = adjacent 16-byte vectors)
  const int by = 256 / bx > 64 ? 64 : 256 / bx;      // output pixels per block
  dim3 block(bx, by), grid((OH * OW + by - 1) / by, (cg + bx - 1) / bx, N);
  const size_t sm = sums ? (size_t)bx * by * 8 * sizeof(float) : 0;
  if (sm > 48 * 1024) return fail(COBEVT_ERR_UNSUPPORTED, "dwconv_swish: reduction scratch too large");
  if (K == 3) cobevt::launch_k(dwconv_kernel<3>, grid, block, sm, ST(stream), (const uint16_t*)x, w, bias, (uint16_t*)y, sums, N, H, W, C, OH, OW, stride, pad_top, pad_left, dtype);
  else cobevt::launch_k(dwconv_kernel<5>, grid, block, sm, ST(stream), (const uint16_t*)x, w, bias, (uint16_t*)y, sums, N, H, W, C, OH, OW, stride, pad_top, pad_left, dtype);
  return cobevt::check_launch("dwconv_swish");
}

extern "C" int cobevt_se_gate(const float* sums, float inv_hw, const float* w1, const float* b1, const float* w2,
                              const float* b2, float* gate, int32_t N, int32_t C, int32_t Cse, void* stream) {
  if (!sums || !w1 || !b1 || !w2 || !b2 || !gate || N <= 0 || C <= 0 || Cse <= 0 || (C + Cse) * 4 > 48 * 1024)
    return fail(COBEVT_ERR_SHAPE, "se_gate: bad args");
  cobevt::launch_k(se_gate_kernel, N, 128, (size_t)(C + Cse) * sizeof(float), ST(stream), sums, inv_hw, w1, b1, w2, b2, gate, C, Cse);
  return cobevt::check_launch("se_gate");
}

extern "C" int cobevt_scale_channels(const void* x, const float* gate, void* y, int32_t N, int64_t HW, int32_t C, int32_t dtype,
                                     void* stream) {
  if (!x || !gate || !y || C % 8) return fail(COBEVT_ERR_SHAPE, "scale_channels: bad args");
  cobevt::launch_k(scale_channels_kernel, blocks_for((int64_t)N * HW * (C / 8), 256), 256, 0, ST(stream), (const uint16_t*)x, gate, (uint16_t*)y, N, HW, C, dtype);
  return cobevt::check_launch("scale_channels");
}

extern "C" int cobevt_row_stats9(const void* x, const float* chan_tab, float* stats, int64_t rows, int32_t C, int32_t dtype,
                                 void* stream) {
  if (!x || !chan_tab || !stats || rows <= 0 || (C != 32 && C != 64 && C != 128)) return fail(COBEVT_ERR_SHAPE, "row_stats9: bad args (C in {32,64,128})");
  const int64_t threads = rows * (C / 8);
  if (C == 128) cobevt::launch_k(row_stats9_kernel<128>, blocks_for(threads, 256), 256, 0, ST(stream), (const uint16_t*)x, chan_tab, stats, rows, dtype);
  else if (C == 64) cobevt::launch_k(row_stats9_kernel<64>, blocks_for(threads, 256), 256, 0, ST(stream), (const uint16_t*)x, chan_tab, stats, rows, dtype);
  else cobevt::launch_k(row_stats9_kernel<32>, blocks_for(threads, 256), 256, 0, ST(stream), (const uint16_t*)x, chan_tab, stats, rows, dtype);
  return cobevt::check_launch("row_stats9");
}

extern "C" int cobevt_bev_q_analytic(const float* qx, int64_t x_rows, const float* stats, const float* world, const float* chan_tab,
                                     const float* E_inv, void* out, int32_t BL, int32_t n, int32_t P, int32_t C, float qa, float qb,
                                     float qd, float ma, float mb, float eps, int32_t dtype, void* stream) {
  if (!qx || !stats || !world || !chan_tab || !E_inv || !out || BL <= 0 || n <= 0 || n > 8 || n > C / 8 || P <= 0 || BL > 65535 ||
      (C != 32 && C != 64 && C != 128) || (x_rows != P && x_rows != (int64_t)BL * P))
    return fail(COBEVT_ERR_SHAPE, "bev_q_analytic: bad args (C in {32,64,128}, n <= min(8, C/8), x_rows in {P, BL*P})");
  // cells per block: ~4 blocks per SM overall, at least one block iteration of tokens
  int64_t cpb = ((int64_t)P * BL + 148 * 4 - 1) / (148 * 4);
  const int tpb = 256 / (C / 8);
  cpb = (cpb + tpb - 1) / tpb * tpb;
  if (cpb < tpb) cpb = tpb;
  if (cpb > 4096) cpb = 4096;
  dim3 grid((unsigned)((P + cpb - 1) / cpb), (unsigned)BL);
  if (C == 128) cobevt::launch_k(bev_q_kernel<128>, grid, 256, 0, ST(stream), qx, x_rows, stats, world, chan_tab, E_inv, (uint16_t*)out, n, P, qa, qb, qd, ma, mb, eps, dtype, (int)cpb);
  else if (C == 64) cobevt::launch_k(bev_q_kernel<64>, grid, 256, 0, ST(stream), qx, x_rows, stats, world, chan_tab, E_inv, (uint16_t*)out, n, P, qa, qb, qd, ma, mb, eps, dtype, (int)cpb);
  else cobevt::launch_k(bev_q_kernel<32>, grid, 256, 0, ST(stream), qx, x_rows, stats, world, chan_tab, E_inv, (uint16_t*)out, n, P, qa, qb, qd, ma, mb, eps, dtype, (int)cpb);
  return cobevt::check_launch("bev_q_analytic");
}

extern "C" int cobevt_se_scale(void* x, const float* sums, float inv_hw, const float* w1, const float* b1, const float* w2,
                               const float* b2, int32_t N, int64_t HW, int32_t C, int32_t Cse, int32_t dtype, void* stream) {
  if (!x || !sums || !w1 || !b1 || !w2 || !b2 || N <= 0 || N > 65535 || C % 8 || Cse <= 0 || (2 * C + Cse) * 4 > 48 * 1024)
    return fail(COBEVT_ERR_SHAPE, "se_scale: bad args");
  int64_t bpi = (HW * (C / 8) + 256 * 4 - 1) / (256 * 4);          // ~4 vectors per thread
  if (bpi < 1) bpi = 1;
  if (bpi > 1024) bpi = 1024;
  dim3 grid((unsigned)bpi, (unsigned)N);
  cobevt::launch_k(se_scale_kernel, grid, 256, (size_t)(2 * C + Cse) * sizeof(float), ST(stream), (uint16_t*)x, sums, inv_hw, w1, b1, w2, b2, HW, C, Cse, dtype);
  return cobevt::check_launch("se_scale");
}
