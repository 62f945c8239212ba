// Round-2 HBM-bound kernels around the FAX path: uint8 camera-frame staging, the
// softmax / argmax post-process (row f2), nearest feature down-scaling (PyramidAxialEncoder scale < 1) and the
// MBConv pieces of the EfficientNet-B4 prefix (row a16): depthwise conv + BN + swish with the squeeze sums,
// the squeeze-excite gate, the gate application.  Reference lines are cited per entry point in
// include/cobevt_b200.h.
#include "common.h"
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
  stage_rgb_u8_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(x, (uint16_t*)y, N, H, W, Hp, Wp, pad_top, pad_left, bgr, sc, sh, dtype);
  return cobevt::check_launch("stage_rgb_u8");
}

extern "C" int cobevt_softmax_argmax(const float* seg, float* prob, int64_t* amap, int32_t N, int32_t C, int64_t HW, void* stream) {
  if (!seg || !prob || !amap || N <= 0 || C <= 0 || HW <= 0) return fail(COBEVT_ERR_SHAPE, "softmax_argmax: bad args");
  softmax_argmax_kernel<<<blocks_for((int64_t)N * HW, 256), 256, 0, ST(stream)>>>(seg, prob, amap, N, C, HW);
  return cobevt::check_launch("softmax_argmax");
}

extern "C" int cobevt_resize_nearest(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t OH, int32_t OW,
                                     float src_per_dst_h, float src_per_dst_w, void* stream) {
  if (!x || !y || C % 8 || OH <= 0 || OW <= 0) return fail(COBEVT_ERR_SHAPE, "resize_nearest: bad args");
  resize_nearest_kernel<<<blocks_for((int64_t)N * OH * OW * (C / 8), 256), 256, 0, ST(stream)>>>(
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
  if (K == 3) dwconv_kernel<3><<<grid, block, sm, ST(stream)>>>((const uint16_t*)x, w, bias, (uint16_t*)y, sums, N, H, W, C, OH, OW, stride, pad_top, pad_left, dtype);
  else dwconv_kernel<5><<<grid, block, sm, ST(stream)>>>((const uint16_t*)x, w, bias, (uint16_t*)y, sums, N, H, W, C, OH, OW, stride, pad_top, pad_left, dtype);
  return cobevt::check_launch("dwconv_swish");
}

extern "C" int cobevt_se_gate(const float* sums, float inv_hw, const float* w1, const float* b1, const float* w2,
                              const float* b2, float* gate, int32_t N, int32_t C, int32_t Cse, void* stream) {
  if (!sums || !w1 || !b1 || !w2 || !b2 || !gate || N <= 0 || C <= 0 || Cse <= 0 || (C + Cse) * 4 > 48 * 1024)
    return fail(COBEVT_ERR_SHAPE, "se_gate: bad args");
  se_gate_kernel<<<N, 128, (size_t)(C + Cse) * sizeof(float), ST(stream)>>>(sums, inv_hw, w1, b1, w2, b2, gate, C, Cse);
  return cobevt::check_launch("se_gate");
}

extern "C" int cobevt_scale_channels(const void* x, const float* gate, void* y, int32_t N, int64_t HW, int32_t C, int32_t dtype,
                                     void* stream) {
  if (!x || !gate || !y || C % 8) return fail(COBEVT_ERR_SHAPE, "scale_channels: bad args");
  scale_channels_kernel<<<blocks_for((int64_t)N * HW * (C / 8), 256), 256, 0, ST(stream)>>>((const uint16_t*)x, gate, (uint16_t*)y, N, HW, C, dtype);
  return cobevt::check_launch("scale_channels");
}
