// HBM-bound element-wise / gather kernels of the FAX path (coalesced 8/16-byte accesses,
// one warp per token row where a row reduction is needed).  Each entry point cites the
// reference lines it replaces; see include/cobevt_b200.h.
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
__device__ __forceinline__ uint16_t pk1(float a, int bf16) { return (uint16_t)(pk2(a, 0.f, bf16) & 0xFFFFu); }
__device__ __forceinline__ float up1(uint16_t u, int bf16) { return up2((uint32_t)u, bf16).x; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float2 warp_sum2(float a, float b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  return make_float2(a, b);
}

constexpr int MAXC_PER_LANE = 8;  // rows up to 256 channels per warp

// load C/32 consecutive channels of a 16-bit row for this lane (C % 64 == 0 -> pairs)
__device__ __forceinline__ void load_row(const uint16_t* row, int C, int lane, int bf16, float* v) {
  const int per = C / 32;  // 2, 4 or 8
  const uint16_t* p = row + lane * per;
  if (per == 4) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = up2(u.x, bf16), b = up2(u.y, bf16);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else if (per == 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = up2(u.x, bf16), b = up2(u.y, bf16), c = up2(u.z, bf16), d = up2(u.w, bf16);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  } else if (per == 2) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    float2 a = up2(u, bf16);
    v[0] = a.x; v[1] = a.y;
  } else {
    v[0] = up1(*p, bf16);
  }
}
__device__ __forceinline__ void store_row(uint16_t* row, int C, int lane, int bf16, const float* v) {
  const int per = C / 32;
  uint16_t* p = row + lane * per;
  if (per == 4) {
    uint2 u; u.x = pk2(v[0], v[1], bf16); u.y = pk2(v[2], v[3], bf16);
    *reinterpret_cast<uint2*>(p) = u;
  } else if (per == 8) {
    uint4 u; u.x = pk2(v[0], v[1], bf16); u.y = pk2(v[2], v[3], bf16); u.z = pk2(v[4], v[5], bf16); u.w = pk2(v[6], v[7], bf16);
    *reinterpret_cast<uint4*>(p) = u;
  } else if (per == 2) {
    *reinterpret_cast<uint32_t*>(p) = pk2(v[0], v[1], bf16);
  } else {
    *p = pk1(v[0], bf16);
  }
}

// normalise a row held across the warp: (v - mean) * rstd [* gamma + beta]
__device__ __forceinline__ void ln_row(float* v, int per, int C, int lane, float eps, const float* gamma, const float* beta) {
  float s = 0.f;
  for (int i = 0; i < per; ++i) s += v[i];
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int i = 0; i < per; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  for (int i = 0; i < per; ++i) {
    float o = (v[i] - mean) * rstd;
    if (gamma) o = fmaf(o, gamma[lane * per + i], beta[lane * per + i]);
    v[i] = o;
  }
}

// ------------------------------------------------------------------ pack_rgb8
__global__ void pack_rgb8_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, int64_t npix, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float r = x[i * 3], g = x[i * 3 + 1], b = x[i * 3 + 2];
  uint4 u;
  u.x = pk2(r, g, bf16); u.y = pk2(b, 0.f, bf16); u.z = 0; u.w = 0;
  *reinterpret_cast<uint4*>(y + i * 8) = u;
}

// ------------------------------------------------------------------ pack_rgb4_padded
// (N,H,W,3) fp32 -> (N, H+pt+pb, Wp, 4) 16-bit with a zero border (pt rows on top, pl columns on the
// left): the 7x7/s2 stem then needs no bounds checks and its taps start on 16-byte boundaries.
__global__ void pack_rgb4_pad_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W,
                                     int Hp, int Wp, int pt, int pl, int nchw, float3 mean, float3 istd, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  // two output pixels (one 16-byte store) per thread, Wp is even: twice the bytes in flight of the one-pixel version
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Wh = Wp >> 1;
  if (i >= (int64_t)N * Hp * Wh) return;
  const int xp = (int)(i % Wh) * 2;
  const int yp = (int)((i / Wh) % Hp);
  const int n = (int)(i / ((int64_t)Wh * Hp));
  const int ys = yp - pt;
  uint32_t o[4] = {0u, 0u, 0u, 0u};
  if (ys >= 0 && ys < H) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int xs = xp + k - pl;
      if (xs >= 0 && xs < W) {
        float r, g, b;
        if (nchw) {     // (N,3,H,W) planes (nuScenes batch['image'])
          const int64_t hw = (int64_t)H * W, off = (int64_t)n * 3 * hw + (int64_t)ys * W + xs;
          r = x[off]; g = x[off + hw]; b = x[off + 2 * hw];
        } else {
          const float* s = x + (((int64_t)n * H + ys) * W + xs) * 3;
          r = s[0]; g = s[1]; b = s[2];
        }
        o[2 * k] = pk2((r - mean.x) * istd.x, (g - mean.y) * istd.y, bf16);
        o[2 * k + 1] = pk2((b - mean.z) * istd.z, 0.f, bf16);
      }
    }
  }
  *reinterpret_cast<uint4*>(y + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------ maxpool 3x3 s2 p1
// Two horizontally adjacent outputs per thread (their 3x3 windows share a column: 15 loads instead of 18), every load
// issued up front with clamped coordinates (no control flow between them: 240 bytes in flight per thread), the
// maximum taken in the packed 16-bit domain (max is exact there), -inf for the padding.
template <bool BF16>
__device__ __forceinline__ uint32_t max2_16(uint32_t a, uint32_t b) {
  if constexpr (BF16) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  } else {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
}
template <bool BF16>
__global__ void __launch_bounds__(256) maxpool_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H,
                                                      int W, int C, int OH, int OW) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int cg = C / 8;
  const int OWp = (OW + 1) >> 1;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * OH * OWp * cg) return;
  const int c8 = (int)(idx % cg);
  int64_t t = idx / cg;
  const int oxp = (int)(t % OWp); t /= OWp;
  const int oy = (int)(t % OH);
  const int n = (int)(t / OH);
  const int ox0 = oxp * 2;
  constexpr uint32_t NINF = BF16 ? 0xFF80FF80u : 0xFC00FC00u;
  uint4 v[3][5];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    const bool vy = iy >= 0 && iy < H;
    const int iyc = min(max(iy, 0), H - 1);
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      const int ix = ox0 * 2 - 1 + kx;
      const bool ok = vy && ix >= 0 && ix < W;
      const int ixc = min(max(ix, 0), W - 1);
      const uint4 u = *reinterpret_cast<const uint4*>(x + (((int64_t)n * H + iyc) * W + ixc) * C + c8 * 8);
      v[ky][kx] = ok ? u : make_uint4(NINF, NINF, NINF, NINF);
    }
  }
  auto mx4 = [](uint4 a, uint4 b) {
    return make_uint4(max2_16<BF16>(a.x, b.x), max2_16<BF16>(a.y, b.y), max2_16<BF16>(a.z, b.z), max2_16<BF16>(a.w, b.w));
  };
  uint4 col[5];
#pragma unroll
  for (int kx = 0; kx < 5; ++kx) col[kx] = mx4(mx4(v[0][kx], v[1][kx]), v[2][kx]);
  const uint4 o0 = mx4(mx4(col[0], col[1]), col[2]);
  const uint4 o1 = mx4(mx4(col[2], col[3]), col[4]);
  uint16_t* yp = y + (((int64_t)n * OH + oy) * OW + ox0) * C + c8 * 8;
  *reinterpret_cast<uint4*>(yp) = o0;
  if (ox0 + 1 < OW) *reinterpret_cast<uint4*>(yp + C) = o1;
}

// ------------------------------------------------------------------ y = relu(x*s + b), two affine sets
// Four 16-byte vectors per thread, one block-stride apart: the four loads are in flight together (a streaming kernel
// with 16 bytes per thread in flight sits at ~4 TB/s).  SAME_CG: 256 % (C / 8) == 0, so all four vectors of a thread
// share one channel group and the scale / shift vectors are read once per thread; otherwise (56 / 112-channel
// EfficientNet features) they are read per vector.
template <bool SAME_CG>
__global__ void __launch_bounds__(256) bn_relu2_kernel(const uint16_t* __restrict__ x, const float* __restrict__ sa,
                                const float* __restrict__ ba, uint16_t* __restrict__ ya, const float* __restrict__ sb,
                                const float* __restrict__ bb, uint16_t* __restrict__ yb, int64_t nvec, int C, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  constexpr int VPT = 4;
  const int cg = C / 8;
  const int64_t base = (int64_t)blockIdx.x * (256 * VPT) + threadIdx.x;
  uint4 u[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int64_t idx = base + (int64_t)i * 256;
    u[i] = idx < nvec ? *reinterpret_cast<const uint4*>(x + idx * 8) : make_uint4(0, 0, 0, 0);
  }
  float s0[8], t0[8], s1[8], t1[8];
  auto load_affine = [&](int c0) {
    *reinterpret_cast<float4*>(s0) = *reinterpret_cast<const float4*>(sa + c0);
    *reinterpret_cast<float4*>(s0 + 4) = *reinterpret_cast<const float4*>(sa + c0 + 4);
    *reinterpret_cast<float4*>(t0) = *reinterpret_cast<const float4*>(ba + c0);
    *reinterpret_cast<float4*>(t0 + 4) = *reinterpret_cast<const float4*>(ba + c0 + 4);
    if (yb) {
      *reinterpret_cast<float4*>(s1) = *reinterpret_cast<const float4*>(sb + c0);
      *reinterpret_cast<float4*>(s1 + 4) = *reinterpret_cast<const float4*>(sb + c0 + 4);
      *reinterpret_cast<float4*>(t1) = *reinterpret_cast<const float4*>(bb + c0);
      *reinterpret_cast<float4*>(t1 + 4) = *reinterpret_cast<const float4*>(bb + c0 + 4);
    }
  };
  if (SAME_CG) load_affine((int)(base % cg) * 8);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int64_t idx = base + (int64_t)i * 256;
    if (idx >= nvec) continue;
    if (!SAME_CG) load_affine((int)(idx % cg) * 8);
    const uint32_t w4[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = up2(w4[j], bf16); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(v[j], s0[j], t0[j]), 0.f);
    uint4 w;
    w.x = pk2(o[0], o[1], bf16); w.y = pk2(o[2], o[3], bf16); w.z = pk2(o[4], o[5], bf16); w.w = pk2(o[6], o[7], bf16);
    *reinterpret_cast<uint4*>(ya + idx * 8) = w;
    if (yb) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(v[j], s1[j], t1[j]), 0.f);
      w.x = pk2(o[0], o[1], bf16); w.y = pk2(o[2], o[3], bf16); w.z = pk2(o[4], o[5], bf16); w.w = pk2(o[6], o[7], bf16);
      *reinterpret_cast<uint4*>(yb + idx * 8) = w;
    }
  }
}

// ------------------------------------------------------------------ ray-direction embedding
// out = normalize(W_img d - W_cam c)   (fax_modules.py:346-358).  blockIdx.y = view; every lane
// keeps the embedding weights of its C/32 channels in registers and the warp walks over cells.
constexpr int CELLS_PER_BLOCK = 256;
template <int PER>
__global__ void __launch_bounds__(256) ray_embed_kernel(const float* __restrict__ I_inv, const float* __restrict__ E_inv,
                                 const float* __restrict__ w_img, const float* __restrict__ w_cam,
                                 uint16_t* __restrict__ out, int NV, int h, int w, float image_h, float image_w,
                                 int C, int invert_I, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = blockIdx.y;
  float Ii[9];
#pragma unroll
  for (int r = 0; r < 9; ++r) Ii[r] = I_inv[v * 9 + r];
  if (invert_I) {   // batch['intrinsic'].inverse(), fax_modules.py:500-501 (adjugate / determinant)
    const float a = Ii[0], b_ = Ii[1], c_ = Ii[2], d_ = Ii[3], e_ = Ii[4], f = Ii[5], g = Ii[6], hh = Ii[7], k = Ii[8];
    const float A = e_ * k - f * hh, B = -(d_ * k - f * g), Cc = d_ * hh - e_ * g;
    const float idet = 1.0f / (a * A + b_ * B + c_ * Cc);
    Ii[0] = A * idet;  Ii[1] = -(b_ * k - c_ * hh) * idet; Ii[2] = (b_ * f - c_ * e_) * idet;
    Ii[3] = B * idet;  Ii[4] = (a * k - c_ * g) * idet;    Ii[5] = -(a * f - c_ * d_) * idet;
    Ii[6] = Cc * idet; Ii[7] = -(a * hh - b_ * g) * idet;  Ii[8] = (a * e_ - b_ * d_) * idet;
  }
  float Ei[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) Ei[r] = E_inv[v * 16 + r];
  constexpr int per = PER;
  float wi[PER][4], ce[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    {
      const int ch = lane * per + i;
      float cacc = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) { wi[i][r] = w_img[ch * 4 + r]; cacc += w_cam[ch * 4 + r] * Ei[r * 4 + 3]; }
      ce[i] = cacc;
    }
  }
  const int ncell = h * w;
  const int base = blockIdx.x * CELLS_PER_BLOCK;
  // two independent cells per iteration (ILP across the shuffle chains)
  for (int cc = warp; cc < CELLS_PER_BLOCK; cc += 16) {
    float e[2][PER];
    float ss[2] = {0.f, 0.f};
    int cell[2] = {base + cc, base + cc + 8};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int cl = cell[u] < ncell ? cell[u] : ncell - 1;
      const int ix = cl % w, iy = cl / w;
      const float px = (w > 1 ? (float)ix / (float)(w - 1) : 0.f) * image_w;
      const float py = (h > 1 ? (float)iy / (float)(h - 1) : 0.f) * image_h;
      const float cx = Ii[0] * px + Ii[1] * py + Ii[2];
      const float cy = Ii[3] * px + Ii[4] * py + Ii[5];
      const float cz = Ii[6] * px + Ii[7] * py + Ii[8];
      float d[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) d[r] = Ei[r * 4] * cx + Ei[r * 4 + 1] * cy + Ei[r * 4 + 2] * cz + Ei[r * 4 + 3];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        {
          const float a = wi[i][0] * d[0] + wi[i][1] * d[1] + wi[i][2] * d[2] + wi[i][3] * d[3] - ce[i];
          e[u][i] = a;
          ss[u] += a * a;
        }
      }
    }
    const float2 tot = warp_sum2(ss[0], ss[1]);
    const float inv[2] = {__fdividef(1.0f, sqrtf(tot.x) + 1e-7f), __fdividef(1.0f, sqrtf(tot.y) + 1e-7f)};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (cell[u] < ncell) {
#pragma unroll
        for (int i = 0; i < PER; ++i) e[u][i] *= inv[u];
        store_row(out + ((int64_t)v * ncell + cell[u]) * C, C, lane, bf16, e[u]);
      }
    }
  }
}

// ------------------------------------------------------------------ BEV positional query
// out[b, n, p, :] = LN0( x[b, p, :] + normalize(W_bev world_p + b_bev - W_cam c_{b,n}) )
// (fax_modules.py:370-375,387-390 followed by the LayerNorm of to_q, :189,216).
// blockIdx.y = b*n + view; lanes keep their channels' weights in registers.
template <int PER>
__global__ void __launch_bounds__(256) bev_query_kernel(const uint16_t* __restrict__ x, int64_t x_rows, const float* __restrict__ world,
                                 const float* __restrict__ w_bev, const float* __restrict__ b_bev,
                                 const float* __restrict__ w_cam, const float* __restrict__ E_inv,
                                 uint16_t* __restrict__ out, int BL, int n, int P, int C, float eps, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = blockIdx.y;
  const int b = v / n;
  const float* Ei = E_inv + (int64_t)v * 16;
  const float c0 = Ei[3], c1 = Ei[7], c2 = Ei[11], c3 = Ei[15];
  constexpr int per = PER;
  float wb[PER][2], ce[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    {
      const int ch = lane * per + i;
      wb[i][0] = w_bev[ch * 2];
      wb[i][1] = w_bev[ch * 2 + 1];
      ce[i] = b_bev[ch] - (w_cam[ch * 4] * c0 + w_cam[ch * 4 + 1] * c1 + w_cam[ch * 4 + 2] * c2 + w_cam[ch * 4 + 3] * c3);
    }
  }
  const int base = blockIdx.x * CELLS_PER_BLOCK;
  const float invC = 1.0f / (float)C;
  for (int cc = warp; cc < CELLS_PER_BLOCK; cc += 16) {
    float e[2][PER];
    float ss[2] = {0.f, 0.f};
    int cell[2] = {base + cc, base + cc + 8};
    float xv[2][PER];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int pp = cell[u] < P ? cell[u] : P - 1;
      const float wx = world[pp], wy = world[P + pp];
      const int64_t xr = (x_rows == (int64_t)P) ? (int64_t)pp : (int64_t)b * P + pp;   // broadcast prior or per-row map
      load_row(x + xr * C, C, lane, bf16, xv[u]);
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        {
          const float a = fmaf(wb[i][0], wx, fmaf(wb[i][1], wy, ce[i]));
          e[u][i] = a;
          ss[u] += a * a;
        }
      }
    }
    const float2 nrm = warp_sum2(ss[0], ss[1]);
    const float inv[2] = {__fdividef(1.0f, sqrtf(nrm.x) + 1e-7f), __fdividef(1.0f, sqrtf(nrm.y) + 1e-7f)};
    float s1[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < PER; ++i) { e[u][i] = fmaf(e[u][i], inv[u], xv[u][i]); s1[u] += e[u][i]; }
    const float2 sm = warp_sum2(s1[0], s1[1]);
    const float mean[2] = {sm.x * invC, sm.y * invC};
    float s2[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < PER; ++i) { const float dd = e[u][i] - mean[u]; s2[u] += dd * dd; }
    const float2 vr = warp_sum2(s2[0], s2[1]);
    const float rstd[2] = {rsqrtf(vr.x * invC + eps), rsqrtf(vr.y * invC + eps)};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (cell[u] < P) {
#pragma unroll
        for (int i = 0; i < PER; ++i) e[u][i] = (e[u][i] - mean[u]) * rstd[u];
        store_row(out + ((int64_t)v * P + cell[u]) * C, C, lane, bf16, e[u]);
      }
    }
  }
}

// ------------------------------------------------------------------ grouped-lane variants
// LPC lanes share one cell (CPL consecutive channels each, a 16- or 32-byte vector), so a warp handles
// 32 / LPC cells per instruction and the three row reductions are log2(LPC) shuffle steps instead of 5.
// These are the versions the entry points launch; the one-warp-per-cell kernels above remain for C
// values the grouped mapping does not cover.
template <int LPC>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPC / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// vector q of a lane covers channels q * (LPC * 8) + grp * 8 .. + 8: every warp-wide 16-byte access then
// touches whole 128-byte lines
template <int CPL, int LPC>
__device__ __forceinline__ int chan_of(int grp, int i) { return (i >> 3) * (LPC * 8) + grp * 8 + (i & 7); }
template <int CPL, int LPC>
__device__ __forceinline__ void load_vec(const uint16_t* row, int grp, int bf16, float (&v)[CPL]) {
#pragma unroll
  for (int q = 0; q < CPL / 8; ++q) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + q * (LPC * 8) + grp * 8);
    const float2 a = up2(u.x, bf16), b = up2(u.y, bf16), c = up2(u.z, bf16), d = up2(u.w, bf16);
    v[q * 8 + 0] = a.x; v[q * 8 + 1] = a.y; v[q * 8 + 2] = b.x; v[q * 8 + 3] = b.y;
    v[q * 8 + 4] = c.x; v[q * 8 + 5] = c.y; v[q * 8 + 6] = d.x; v[q * 8 + 7] = d.y;
  }
}
template <int CPL, int LPC>
__device__ __forceinline__ void store_vec(uint16_t* row, int grp, int bf16, const float (&v)[CPL]) {
#pragma unroll
  for (int q = 0; q < CPL / 8; ++q) {
    uint4 u;
    u.x = pk2(v[q * 8 + 0], v[q * 8 + 1], bf16); u.y = pk2(v[q * 8 + 2], v[q * 8 + 3], bf16);
    u.z = pk2(v[q * 8 + 4], v[q * 8 + 5], bf16); u.w = pk2(v[q * 8 + 6], v[q * 8 + 7], bf16);
    *reinterpret_cast<uint4*>(row + q * (LPC * 8) + grp * 8) = u;
  }
}

template <int CPL, int LPC>
__global__ void __launch_bounds__(256) ray_embed_grouped_kernel(const float* __restrict__ I_inv, const float* __restrict__ E_inv,
                                 const float* __restrict__ w_img, const float* __restrict__ w_cam,
                                 uint16_t* __restrict__ out, int NV, int h, int w, float image_h, float image_w,
                                 int invert_I, int bf16, int cpb) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  constexpr int C = CPL * LPC, CPW = 32 / LPC;          // channels, cells per warp and step
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane % LPC, sub = lane / LPC;         // channel group, cell slot inside the warp
  const int v = blockIdx.y;
  float Ii[9];
#pragma unroll
  for (int r = 0; r < 9; ++r) Ii[r] = I_inv[v * 9 + r];
  if (invert_I) {   // batch['intrinsic'].inverse(), fax_modules.py:500-501 (adjugate / determinant)
    const float a = Ii[0], b_ = Ii[1], c_ = Ii[2], d_ = Ii[3], e_ = Ii[4], f = Ii[5], g = Ii[6], hh = Ii[7], k = Ii[8];
    const float A = e_ * k - f * hh, B = -(d_ * k - f * g), Cc = d_ * hh - e_ * g;
    const float idet = 1.0f / (a * A + b_ * B + c_ * Cc);
    Ii[0] = A * idet;  Ii[1] = -(b_ * k - c_ * hh) * idet; Ii[2] = (b_ * f - c_ * e_) * idet;
    Ii[3] = B * idet;  Ii[4] = (a * k - c_ * g) * idet;    Ii[5] = -(a * f - c_ * d_) * idet;
    Ii[6] = Cc * idet; Ii[7] = -(a * hh - b_ * g) * idet;  Ii[8] = (a * e_ - b_ * d_) * idet;
  }
  // pixel -> ray direction is affine: d = M3 (px, py, 1) with M3 = E_inv[:, :3] I_inv (4 x 3) plus E_inv[:, 3]
  float M[4][3];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float e0 = E_inv[v * 16 + r * 4], e1 = E_inv[v * 16 + r * 4 + 1], e2 = E_inv[v * 16 + r * 4 + 2], e3 = E_inv[v * 16 + r * 4 + 3];
#pragma unroll
    for (int c = 0; c < 3; ++c) M[r][c] = e0 * Ii[c] + e1 * Ii[3 + c] + e2 * Ii[6 + c];
    M[r][2] += e3;
  }
  // per-channel: a = wi . d - ce  ->  a = A0 px + A1 py + A2   (3 FMAs per element instead of 16 + 4)
  float A0[CPL], A1[CPL], A2[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int ch = chan_of<CPL, LPC>(grp, i);
    const float w0 = w_img[ch * 4], w1 = w_img[ch * 4 + 1], w2 = w_img[ch * 4 + 2], w3 = w_img[ch * 4 + 3];
    float cacc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) cacc += w_cam[ch * 4 + r] * E_inv[v * 16 + r * 4 + 3];
    A0[i] = w0 * M[0][0] + w1 * M[1][0] + w2 * M[2][0] + w3 * M[3][0];
    A1[i] = w0 * M[0][1] + w1 * M[1][1] + w2 * M[2][1] + w3 * M[3][1];
    A2[i] = w0 * M[0][2] + w1 * M[1][2] + w2 * M[2][2] + w3 * M[3][2] - cacc;
  }
  const int ncell = h * w;
  const int base = blockIdx.x * cpb;
  const float sx = w > 1 ? image_w / (float)(w - 1) : 0.f, sy = h > 1 ? image_h / (float)(h - 1) : 0.f;
  for (int cc = warp * CPW + sub; cc < cpb; cc += 8 * CPW) {
    const int cell = base + cc;
    const int cl = cell < ncell ? cell : ncell - 1;
    const int iy = cl / w, ix = cl - iy * w;
    const float px = (float)ix * sx, py = (float)iy * sy;
    float e[CPL];
    float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; i += 2) {
      sm100::ffma2(e[i], e[i + 1], A1[i], A1[i + 1], py, py, A2[i], A2[i + 1]);
      sm100::ffma2(e[i], e[i + 1], A0[i], A0[i + 1], px, px, e[i], e[i + 1]);
      sm100::ffma2(ss0, ss1, e[i], e[i + 1], e[i], e[i + 1], ss0, ss1);
    }
    const float ss = group_sum<LPC>(ss0 + ss1);
    const float inv = __fdividef(1.0f, sqrtf(ss) + 1e-7f);
#pragma unroll
    for (int i = 0; i < CPL; i += 2) sm100::fmul2(e[i], e[i + 1], e[i], e[i + 1], inv, inv);
    if (cell < ncell) store_vec<CPL, LPC>(out + ((int64_t)v * ncell + cell) * C, grp, bf16, e);
  }
}

template <int CPL, int LPC>
__global__ void __launch_bounds__(256) bev_query_grouped_kernel(const uint16_t* __restrict__ x, int64_t x_rows, const float* __restrict__ world,
                                 const float* __restrict__ w_bev, const float* __restrict__ b_bev,
                                 const float* __restrict__ w_cam, const float* __restrict__ E_inv,
                                 uint16_t* __restrict__ out, int n, int P, float eps, int bf16, int cpb) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  constexpr int C = CPL * LPC, CPW = 32 / LPC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane % LPC, sub = lane / LPC;
  const int v = blockIdx.y;
  const int b = v / n;
  const float* Ei = E_inv + (int64_t)v * 16;
  const float c0 = Ei[3], c1 = Ei[7], c2 = Ei[11], c3 = Ei[15];
  float wb0[CPL], wb1[CPL], ce[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int ch = chan_of<CPL, LPC>(grp, i);
    wb0[i] = w_bev[ch * 2];
    wb1[i] = w_bev[ch * 2 + 1];
    ce[i] = b_bev[ch] - (w_cam[ch * 4] * c0 + w_cam[ch * 4 + 1] * c1 + w_cam[ch * 4 + 2] * c2 + w_cam[ch * 4 + 3] * c3);
  }
  // e(cell) = wb0 wx + wb1 wy + ce is affine in the cell's world coordinates, so its squared norm is a quadratic form
  // with six per-view constants: the L2 normalisation needs no per-cell reduction; the LayerNorm statistics of
  // y = x + e / (|e| + 1e-7) then come from ONE reduction round of (sum y, sum y^2) instead of three dependent ones.
  float qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f, qe = 0.f, qf = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    qa = fmaf(wb0[i], wb0[i], qa); qb = fmaf(wb1[i], wb1[i], qb); qc = fmaf(ce[i], ce[i], qc);
    qd = fmaf(wb0[i], wb1[i], qd); qe = fmaf(wb0[i], ce[i], qe); qf = fmaf(wb1[i], ce[i], qf);
  }
  qa = group_sum<LPC>(qa); qb = group_sum<LPC>(qb); qc = group_sum<LPC>(qc);
  qd = 2.0f * group_sum<LPC>(qd); qe = 2.0f * group_sum<LPC>(qe); qf = 2.0f * group_sum<LPC>(qf);
  const int base = blockIdx.x * cpb;
  constexpr float invC = 1.0f / (float)C;
  // the loads of the next cell are issued before the reductions of the current one (register prefetch)
  const int64_t xbase = (x_rows == (int64_t)P) ? 0 : (int64_t)b * P;   // broadcast prior or per-row map
  uint4 xn[CPL / 8];
  float wxn, wyn;
  auto fetch = [&](int cc) {
    const int cell = base + cc;
    const int pp = cell < P ? cell : P - 1;
    wxn = world[pp]; wyn = world[P + pp];
    const uint16_t* row = x + (xbase + pp) * C;
#pragma unroll
    for (int q = 0; q < CPL / 8; ++q) xn[q] = *reinterpret_cast<const uint4*>(row + q * (LPC * 8) + grp * 8);
  };
  fetch(warp * CPW + sub);
  for (int cc = warp * CPW + sub; cc < cpb; cc += 8 * CPW) {
    const int cell = base + cc;
    const float wx = wxn, wy = wyn;
    float xv[CPL], e[CPL];
#pragma unroll
    for (int q = 0; q < CPL / 8; ++q) {
      const float2 a = up2(xn[q].x, bf16), b2 = up2(xn[q].y, bf16), c = up2(xn[q].z, bf16), d = up2(xn[q].w, bf16);
      xv[q * 8 + 0] = a.x; xv[q * 8 + 1] = a.y; xv[q * 8 + 2] = b2.x; xv[q * 8 + 3] = b2.y;
      xv[q * 8 + 4] = c.x; xv[q * 8 + 5] = c.y; xv[q * 8 + 6] = d.x; xv[q * 8 + 7] = d.y;
    }
    if (cc + 8 * CPW < cpb) fetch(cc + 8 * CPW);
    const float ss = fmaxf(fmaf(wx, fmaf(qa, wx, fmaf(qd, wy, qe)), fmaf(wy, fmaf(qb, wy, qf), qc)), 0.f);
    const float inv = __fdividef(1.0f, sqrtf(ss) + 1e-7f);
    float s10 = 0.f, s11 = 0.f, s20 = 0.f, s21 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; i += 2) {
      sm100::ffma2(e[i], e[i + 1], wb1[i], wb1[i + 1], wy, wy, ce[i], ce[i + 1]);
      sm100::ffma2(e[i], e[i + 1], wb0[i], wb0[i + 1], wx, wx, e[i], e[i + 1]);
      sm100::ffma2(e[i], e[i + 1], e[i], e[i + 1], inv, inv, xv[i], xv[i + 1]);          // y = x + e / |e|
      sm100::fadd2(s10, s11, s10, s11, e[i], e[i + 1]);
      sm100::ffma2(s20, s21, e[i], e[i + 1], e[i], e[i + 1], s20, s21);
    }
    float s1 = s10 + s11, s2 = s20 + s21;
#pragma unroll
    for (int o = LPC / 2; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float mean = s1 * invC;
    const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, s2 * invC), 0.f) + eps);
    const float shift = -mean * rstd;
#pragma unroll
    for (int i = 0; i < CPL; i += 2) sm100::ffma2(e[i], e[i + 1], e[i], e[i + 1], rstd, rstd, shift, shift);
    if (cell < P) store_vec<CPL, LPC>(out + ((int64_t)v * P + cell) * C, grp, bf16, e);
  }
}

// ------------------------------------------------------------------ LayerNorm over rows
__global__ void layernorm_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, uint16_t* __restrict__ y, int64_t rows, int C,
                                 float eps, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float v[MAXC_PER_LANE];
  load_row(x + row * C, C, lane, bf16, v);
  ln_row(v, C / 32, C, lane, eps, gamma, beta);
  store_row(y + row * C, C, lane, bf16, v);
}

// ------------------------------------------------------------------ mean over views + LayerNorm
// x (B, L, P, C) -> y (B, P, C) = LN0(mean_l x)   (swap_fusion_modules.py:275-278)
__global__ void mean_views_ln_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int B, int L, int P,
                                     int C, float eps, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= (int64_t)B * P) return;
  const int p = (int)(row % P);
  const int b = (int)(row / P);
  const int per = C / 32;
  float acc[MAXC_PER_LANE];
  for (int i = 0; i < per; ++i) acc[i] = 0.f;
  for (int l = 0; l < L; ++l) {
    float v[MAXC_PER_LANE];
    load_row(x + (((int64_t)b * L + l) * P + p) * C, C, lane, bf16, v);
    for (int i = 0; i < per; ++i) acc[i] += v[i];
  }
  const float invL = 1.0f / (float)L;
  for (int i = 0; i < per; ++i) acc[i] *= invL;
  ln_row(acc, per, C, lane, eps, nullptr, nullptr);
  store_row(y + row * C, C, lane, bf16, acc);
}

// ------------------------------------------------------------------ regroup + STTF warp + ROI mask
// One warp per (b, l, h, w) output cell.  Restates corpbevt.py:28-64 (transpose, flip, affine
// warp about the map centre, flip, transpose) directly in (h, w) index space, fuse_utils.py:8-61
// (regroup: row prefix[b] + l, zero rows beyond record_len) and
// torch_transformation_utils.py:11-49 (nearest-warped ROI x cav mask).
__global__ void sttf_kernel(const uint16_t* __restrict__ x, const float* __restrict__ tm,
                            const int64_t* __restrict__ record_len, uint16_t* __restrict__ out,
                            uint8_t* __restrict__ mask, int B, int L, int H, int W, int C, float inv_disc,
                            int use_roi, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t cell = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (cell >= (int64_t)B * L * H * W) return;
  const int w = (int)(cell % W);
  const int h = (int)((cell / W) % H);
  const int l = (int)((cell / ((int64_t)W * H)) % L);
  const int b = (int)(cell / ((int64_t)W * H * L));
  int64_t start = 0;
  for (int i = 0; i < b; ++i) start += record_len[i];
  const int len = (int)record_len[b];
  const int per = C / 32;
  float acc[MAXC_PER_LANE];
  for (int i = 0; i < per; ++i) acc[i] = 0.f;
  uint8_t mk = 0;
  if (l < len) {
    // warp works on the transposed+flipped map t (rows i = w index, cols j = H-1-h): Ht = W, Wt = H
    const float* M = tm + ((int64_t)b * L + l) * 16;
    const float a = M[0], bb = M[1], tx = M[3] * inv_disc;
    const float c = M[4], d = M[5], ty = M[7] * inv_disc;
    const float cx = (float)H * 0.5f, cy = (float)W * 0.5f;    // centre = (Wt/2, Ht/2)
    const float t0 = cx - a * cx - bb * cy + tx;
    const float t1 = cy - c * cx - d * cy + ty;
    const float det = a * d - bb * c;
    const float ia = d / det, ib = -bb / det, ic = -c / det, id = a / det;
    const float xd = (float)(H - 1 - h), yd = (float)w;        // destination pixel in t-space
    const float sx = ia * (xd - t0) + ib * (yd - t1);
    const float sy = ic * (xd - t0) + id * (yd - t1);
    const uint16_t* src = x + (start + l) * (int64_t)H * W * C;
    const float x0 = floorf(sx), y0 = floorf(sy);
    const float fx = sx - x0, fy = sy - y0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int jx = (int)x0 + (t & 1), iy = (int)y0 + (t >> 1);   // t-space col / row
      const float wgt = ((t & 1) ? fx : 1.f - fx) * ((t >> 1) ? fy : 1.f - fy);
      if (jx < 0 || jx >= H || iy < 0 || iy >= W) continue;
      // t[i = iy][j = jx] = x[h = H-1-jx][w = iy]
      float v[MAXC_PER_LANE];
      load_row(src + ((int64_t)(H - 1 - jx) * W + iy) * C, C, lane, bf16, v);
      for (int i = 0; i < per; ++i) acc[i] = fmaf(v[i], wgt, acc[i]);
    }
    if (use_roi) {
      // the ROI mask is warped on the un-transposed (H, W) canvas (torch_transformation_utils.py:43):
      // centre (W/2, H/2), destination pixel (x = w, y = h)
      const float rcx = (float)W * 0.5f, rcy = (float)H * 0.5f;
      const float r0 = rcx - a * rcx - bb * rcy + tx;
      const float r1 = rcy - c * rcx - d * rcy + ty;
      const float mx = ia * ((float)w - r0) + ib * ((float)h - r1);
      const float my = ic * ((float)w - r0) + id * ((float)h - r1);
      const float rx = rintf(mx), ry = rintf(my);
      mk = (rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H) ? 1 : 0;
    } else {
      mk = 1;
    }
  }
  store_row(out + cell * C, C, lane, bf16, acc);
  if (lane == 0) mask[(((int64_t)b * H + h) * W + w) * L + l] = mk;
}


// ------------------------------------------------------------------ bilinear x2 upsample (align_corners=True)
// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True), nuscenes decoder.py:13
__global__ void upsample_bilinear2x_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W,
                                           int C, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int cg = C / 8, OH = 2 * H, OW = 2 * W;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * OH * OW * cg) return;
  const int c8 = (int)(idx % cg);
  int64_t t = idx / cg;
  const int ox = (int)(t % OW); t /= OW;
  const int oy = (int)(t % OH);
  const int n = (int)(t / OH);
  const float sy = OH > 1 ? (float)oy * (float)(H - 1) / (float)(OH - 1) : 0.f;
  const float sx = OW > 1 ? (float)ox * (float)(W - 1) / (float)(OW - 1) : 0.f;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float fy = sy - (float)y0, fx = sx - (float)x0;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
  const float wy[2] = {1.f - fy, fy}, wx[2] = {1.f - fx, fx};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (((int64_t)n * H + ys[a]) * W + xs[b]) * C + c8 * 8);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
      const float wgt = wy[a] * wx[b];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = up2(w4[j], bf16);
        acc[2 * j] = fmaf(f.x, wgt, acc[2 * j]);
        acc[2 * j + 1] = fmaf(f.y, wgt, acc[2 * j + 1]);
      }
    }
  uint4 o;
  o.x = pk2(acc[0], acc[1], bf16); o.y = pk2(acc[2], acc[3], bf16); o.z = pk2(acc[4], acc[5], bf16); o.w = pk2(acc[6], acc[7], bf16);
  *reinterpret_cast<uint4*>(y + idx * 8) = o;
}

// ------------------------------------------------------------------ y = relu(y + nearest_resize(up))
// DecoderBlock.forward, nuscenes decoder.py:28-36 (F.interpolate default = nearest)
__global__ void add_nearest_relu_kernel(uint16_t* __restrict__ y, const uint16_t* __restrict__ up, int N, int H, int W,
                                        int h, int w, int C, int bf16) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int cg = C / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * H * W * cg) return;
  const int c8 = (int)(idx % cg);
  int64_t t = idx / cg;
  const int ox = (int)(t % W); t /= W;
  const int oy = (int)(t % H);
  const int n = (int)(t / H);
  // torch nearest: src = floor(dst * (in / out))
  const int sy = min((int)floorf((float)oy * ((float)h / (float)H)), h - 1);
  const int sx = min((int)floorf((float)ox * ((float)w / (float)W)), w - 1);
  const uint4 a = *reinterpret_cast<const uint4*>(y + idx * 8);
  const uint4 b = *reinterpret_cast<const uint4*>(up + (((int64_t)n * h + sy) * w + sx) * C + c8 * 8);
  const uint32_t a4[4] = {a.x, a.y, a.z, a.w}, b4[4] = {b.x, b.y, b.z, b.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fa = up2(a4[j], bf16), fb = up2(b4[j], bf16);
    o[j] = pk2(fmaxf(fa.x + fb.x, 0.f), fmaxf(fa.y + fb.y, 0.f), bf16);
  }
  *reinterpret_cast<uint4*>(y + idx * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------ batched small-matrix inverse (3x3 / 4x4)
// batch['intrinsics'].inverse(), batch['extrinsics'].inverse()  (encoder_pyramid_axial.py:538-539)
template <int D>
__global__ void invert_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[D][2 * D];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) { a[r][c] = src[i * D * D + r * D + c]; a[r][D + c] = (r == c) ? 1.f : 0.f; }
#pragma unroll
  for (int col = 0; col < D; ++col) {          // Gauss-Jordan with partial pivoting
    int piv = col;
    float best = fabsf(a[col][col]);
#pragma unroll
    for (int r = 0; r < D; ++r) if (r > col && fabsf(a[r][col]) > best) { best = fabsf(a[r][col]); piv = r; }
#pragma unroll
    for (int r = 0; r < D; ++r)
      if (r == piv && piv != col) {
#pragma unroll
        for (int c = 0; c < 2 * D; ++c) { const float t = a[col][c]; a[col][c] = a[r][c]; a[r][c] = t; }
      }
    const float inv = 1.0f / a[col][col];
#pragma unroll
    for (int c = 0; c < 2 * D; ++c) a[col][c] *= inv;
#pragma unroll
    for (int r = 0; r < D; ++r)
      if (r != col) {
        const float f = a[r][col];
#pragma unroll
        for (int c = 0; c < 2 * D; ++c) a[r][c] = fmaf(-f, a[col][c], a[r][c]);
      }
  }
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) dst[i * D * D + r * D + c] = a[r][D + c];
}

// ------------------------------------------------------------------ layout converters (API boundary)
// dtype codes: 0 f16, 1 bf16, 2 f32
__device__ __forceinline__ float ld_any(const void* p, int64_t i, int dt) {
  if (dt == 2) return reinterpret_cast<const float*>(p)[i];
  return up1(reinterpret_cast<const uint16_t*>(p)[i], dt);
}
__device__ __forceinline__ void st_any(void* p, int64_t i, int dt, float v) {
  if (dt == 2) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<uint16_t*>(p)[i] = pk1(v, dt);
}
// src NCHW (C channels) -> dst NHWC with Cpad channels (zero filled), or the reverse
__global__ void permute_kernel(const void* __restrict__ src, void* __restrict__ dst, int N, int C, int H, int W,
                               int Cpad, int to_nhwc, int sdt, int ddt) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (to_nhwc) {
    const int64_t total = (int64_t)N * H * W * Cpad;
    if (idx >= total) return;
    const int c = (int)(idx % Cpad);
    const int64_t pix = idx / Cpad;
    const int64_t hw = (int64_t)H * W;
    const int n = (int)(pix / hw);
    const int64_t r = pix % hw;
    st_any(dst, idx, ddt, c < C ? ld_any(src, ((int64_t)n * C + c) * hw + r, sdt) : 0.f);
  } else {
    const int64_t total = (int64_t)N * C * H * W;
    if (idx >= total) return;
    const int64_t hw = (int64_t)H * W;
    const int64_t r = idx % hw;
    const int c = (int)((idx / hw) % C);
    const int n = (int)(idx / (hw * C));
    st_any(dst, idx, ddt, ld_any(src, ((int64_t)n * hw + r) * Cpad + c, sdt));
  }
}

inline unsigned blocks_for(int64_t threads, int bs) { return (unsigned)((threads + bs - 1) / bs); }
inline bool ok_c(int C) { return C == 32 || C == 64 || C == 128 || C == 256; }

}  // namespace

using cobevt::fail;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int cobevt_pack_rgb8(const float* x, void* y, int64_t npix, int32_t dtype, void* stream) {
  if (!x || !y || npix <= 0) return fail(COBEVT_ERR_SHAPE, "pack_rgb8: bad args");
  cobevt::launch_k(pack_rgb8_kernel, blocks_for(npix, 256), 256, 0, ST(stream), x, (uint16_t*)y, npix, dtype);
  return cobevt::check_launch("pack_rgb8");
}

extern "C" int cobevt_pack_rgb4_padded(const float* x, void* y, int32_t N, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                                        int32_t pad_top, int32_t pad_left, int32_t src_nchw, const float* mean3,
                                        const float* std3, int32_t dtype, void* stream) {
  if (!x || !y || N <= 0 || Hp < H + pad_top || Wp < W + pad_left || (Wp & 1)) return fail(COBEVT_ERR_SHAPE, "pack_rgb4_padded: bad args (Wp must be even)");
  const int64_t total = (int64_t)N * Hp * (Wp / 2);
  // mean3 / std3 are HOST pointers (3 floats each) or NULL: the Normalize of encoder_pyramid_axial.py:41-49
  float3 mean = make_float3(0.f, 0.f, 0.f), istd = make_float3(1.f, 1.f, 1.f);
  if (mean3 && std3) { mean = make_float3(mean3[0], mean3[1], mean3[2]); istd = make_float3(1.f / std3[0], 1.f / std3[1], 1.f / std3[2]); }
  cobevt::launch_k(pack_rgb4_pad_kernel, blocks_for(total, 256), 256, 0, ST(stream), x, (uint16_t*)y, N, H, W, Hp, Wp, pad_top, pad_left, src_nchw, mean, istd, dtype);
  return cobevt::check_launch("pack_rgb4_padded");
}

extern "C" int cobevt_maxpool3x3s2_nhwc(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C,
                                         int32_t dtype, void* stream) {
  if (!x || !y || C % 8) return fail(COBEVT_ERR_SHAPE, "maxpool: bad args");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)N * OH * ((OW + 1) / 2) * (C / 8);
  if (dtype == COBEVT_BF16) cobevt::launch_k(maxpool_kernel<true>, blocks_for(total, 256), 256, 0, ST(stream), (const uint16_t*)x, (uint16_t*)y, N, H, W, C, OH, OW);
  else cobevt::launch_k(maxpool_kernel<false>, blocks_for(total, 256), 256, 0, ST(stream), (const uint16_t*)x, (uint16_t*)y, N, H, W, C, OH, OW);
  return cobevt::check_launch("maxpool");
}

extern "C" int cobevt_bn_relu2(const void* x, const float* sa, const float* ba, void* ya, const float* sb,
                                const float* bb, void* yb, int64_t rows, int32_t C, int32_t dtype, void* stream) {
  if (!x || !sa || !ba || !ya || C % 8 || (yb && (!sb || !bb))) return fail(COBEVT_ERR_SHAPE, "bn_relu2: bad args");
  const int64_t nvec = rows * (C / 8);
  if (256 % (C / 8) == 0)
    cobevt::launch_k(bn_relu2_kernel<true>, blocks_for(nvec, 1024), 256, 0, ST(stream), (const uint16_t*)x, sa, ba, (uint16_t*)ya, sb, bb,
                                                                         (uint16_t*)yb, nvec, C, dtype);
  else
    cobevt::launch_k(bn_relu2_kernel<false>, blocks_for(nvec, 1024), 256, 0, ST(stream), (const uint16_t*)x, sa, ba, (uint16_t*)ya, sb, bb,
                                                                          (uint16_t*)yb, nvec, C, dtype);
  return cobevt::check_launch("bn_relu2");
}

// cells per block of the grouped kernels: the per-block set-up (embedding weights into registers) is ~100 loads
// per thread, so blocks are made as long as a grid of ~6 blocks per SM allows
static int cells_per_block(int64_t ncell, int views) {
  int64_t cpb = (ncell * views + 148 * 6 - 1) / (148 * 6);
  cpb = (cpb + 255) / 256 * 256;
  if (cpb < 256) cpb = 256;
  if (cpb > 8192) cpb = 8192;
  return (int)cpb;
}

extern "C" int cobevt_ray_embed(const float* I_inv, const float* E_inv, const float* w_img, const float* w_cam, void* out,
                                 int32_t NV, int32_t h, int32_t w, float image_h, float image_w, int32_t C,
                                 int32_t invert_intrinsic, int32_t dtype, void* stream) {
  if (!I_inv || !E_inv || !w_img || !w_cam || !out || !ok_c(C)) return fail(COBEVT_ERR_SHAPE, "ray_embed: bad args (C must be 64/128/256)");
  const int cpb = cells_per_block((int64_t)h * w, NV);
  dim3 gridg((h * w + cpb - 1) / cpb, NV);
  dim3 grid((h * w + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK, NV);
  if (C == 128) cobevt::launch_k(ray_embed_grouped_kernel<16, 8>, gridg, 256, 0, ST(stream), I_inv, E_inv, w_img, w_cam, (uint16_t*)out, NV, h, w, image_h, image_w, invert_intrinsic, dtype, cpb);
  else if (C == 64) cobevt::launch_k(ray_embed_grouped_kernel<8, 8>, gridg, 256, 0, ST(stream), I_inv, E_inv, w_img, w_cam, (uint16_t*)out, NV, h, w, image_h, image_w, invert_intrinsic, dtype, cpb);
  else if (C == 32) cobevt::launch_k(ray_embed_grouped_kernel<8, 4>, gridg, 256, 0, ST(stream), I_inv, E_inv, w_img, w_cam, (uint16_t*)out, NV, h, w, image_h, image_w, invert_intrinsic, dtype, cpb);
  else if (C == 256) cobevt::launch_k(ray_embed_grouped_kernel<16, 16>, gridg, 256, 0, ST(stream), I_inv, E_inv, w_img, w_cam, (uint16_t*)out, NV, h, w, image_h, image_w, invert_intrinsic, dtype, cpb);
  else if (C == 128) cobevt::launch_k(ray_embed_kernel<4>, grid, 256, 0, ST(stream), I_inv, E_inv, w_img, w_cam, (uint16_t*)out, NV, h, w, image_h, image_w, C, invert_intrinsic, dtype);
  else if (C == 32) cobevt::launch_k(ray_embed_kernel<1>, grid, 256, 0, ST(stream), I_inv, E_inv, w_img, w_cam, (uint16_t*)out, NV, h, w, image_h, image_w, C, invert_intrinsic, dtype);
  else if (C == 64) cobevt::launch_k(ray_embed_kernel<2>, grid, 256, 0, ST(stream), I_inv, E_inv, w_img, w_cam, (uint16_t*)out, NV, h, w, image_h, image_w, C, invert_intrinsic, dtype);
  else cobevt::launch_k(ray_embed_kernel<8>, grid, 256, 0, ST(stream), I_inv, E_inv, w_img, w_cam, (uint16_t*)out, NV, h, w, image_h, image_w, C, invert_intrinsic, dtype);
  return cobevt::check_launch("ray_embed");
}

extern "C" int cobevt_bev_query(const void* x, int64_t x_rows, const float* world, const float* w_bev, const float* b_bev,
                                 const float* w_cam, const float* E_inv, void* out, int32_t BL, int32_t n, int32_t P,
                                 int32_t C, float eps, int32_t dtype, void* stream) {
  if (!x || !world || !w_bev || !b_bev || !w_cam || !E_inv || !out || !ok_c(C) || x_rows <= 0)
    return fail(COBEVT_ERR_SHAPE, "bev_query: bad args");
  const int cpb = cells_per_block(P, BL * n);
  dim3 gridg((P + cpb - 1) / cpb, BL * n);
  dim3 grid((P + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK, BL * n);
  if (C == 128) cobevt::launch_k(bev_query_grouped_kernel<16, 8>, gridg, 256, 0, ST(stream), (const uint16_t*)x, x_rows, world, w_bev, b_bev, w_cam, E_inv, (uint16_t*)out, n, P, eps, dtype, cpb);
  else if (C == 64) cobevt::launch_k(bev_query_grouped_kernel<8, 8>, gridg, 256, 0, ST(stream), (const uint16_t*)x, x_rows, world, w_bev, b_bev, w_cam, E_inv, (uint16_t*)out, n, P, eps, dtype, cpb);
  else if (C == 32) cobevt::launch_k(bev_query_grouped_kernel<8, 4>, gridg, 256, 0, ST(stream), (const uint16_t*)x, x_rows, world, w_bev, b_bev, w_cam, E_inv, (uint16_t*)out, n, P, eps, dtype, cpb);
  else if (C == 256) cobevt::launch_k(bev_query_grouped_kernel<16, 16>, gridg, 256, 0, ST(stream), (const uint16_t*)x, x_rows, world, w_bev, b_bev, w_cam, E_inv, (uint16_t*)out, n, P, eps, dtype, cpb);
  else if (C == 128) cobevt::launch_k(bev_query_kernel<4>, grid, 256, 0, ST(stream), (const uint16_t*)x, x_rows, world, w_bev, b_bev, w_cam, E_inv, (uint16_t*)out, BL, n, P, C, eps, dtype);
  else if (C == 32) cobevt::launch_k(bev_query_kernel<1>, grid, 256, 0, ST(stream), (const uint16_t*)x, x_rows, world, w_bev, b_bev, w_cam, E_inv, (uint16_t*)out, BL, n, P, C, eps, dtype);
  else if (C == 64) cobevt::launch_k(bev_query_kernel<2>, grid, 256, 0, ST(stream), (const uint16_t*)x, x_rows, world, w_bev, b_bev, w_cam, E_inv, (uint16_t*)out, BL, n, P, C, eps, dtype);
  else cobevt::launch_k(bev_query_kernel<8>, grid, 256, 0, ST(stream), (const uint16_t*)x, x_rows, world, w_bev, b_bev, w_cam, E_inv, (uint16_t*)out, BL, n, P, C, eps, dtype);
  return cobevt::check_launch("bev_query");
}

extern "C" int cobevt_layernorm_rows(const void* x, const float* gamma, const float* beta, void* y, int64_t rows,
                                      int32_t C, float eps, int32_t dtype, void* stream) {
  if (!x || !y || !ok_c(C) || ((gamma == nullptr) != (beta == nullptr))) return fail(COBEVT_ERR_SHAPE, "layernorm: bad args");
  cobevt::launch_k(layernorm_kernel, blocks_for(rows * 32, 256), 256, 0, ST(stream), (const uint16_t*)x, gamma, beta, (uint16_t*)y, rows, C, eps, dtype);
  return cobevt::check_launch("layernorm");
}

extern "C" int cobevt_mean_views_ln(const void* x, void* y, int32_t B, int32_t L, int32_t P, int32_t C, float eps,
                                     int32_t dtype, void* stream) {
  if (!x || !y || !ok_c(C) || L <= 0) return fail(COBEVT_ERR_SHAPE, "mean_views_ln: bad args");
  cobevt::launch_k(mean_views_ln_kernel, blocks_for((int64_t)B * P * 32, 256), 256, 0, ST(stream), (const uint16_t*)x, (uint16_t*)y, B, L, P, C, eps, dtype);
  return cobevt::check_launch("mean_views_ln");
}

extern "C" int cobevt_sttf_regroup(const void* x, const float* tm, const int64_t* record_len, void* out, uint8_t* mask,
                                    int32_t B, int32_t L, int32_t H, int32_t W, int32_t C, float discrete_ratio,
                                    float downsample_rate, int32_t use_roi_mask, int32_t dtype, void* stream) {
  if (!x || !tm || !record_len || !out || !mask || !ok_c(C)) return fail(COBEVT_ERR_SHAPE, "sttf: bad args");
  const float inv_disc = 1.0f / (discrete_ratio * downsample_rate);
  const int64_t threads = (int64_t)B * L * H * W * 32;
  cobevt::launch_k(sttf_kernel, blocks_for(threads, 256), 256, 0, ST(stream), (const uint16_t*)x, tm, record_len, (uint16_t*)out, mask, B, L, H,
                                                                W, C, inv_disc, use_roi_mask, dtype);
  return cobevt::check_launch("sttf");
}

extern "C" int cobevt_upsample_bilinear2x(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t dtype,
                                           void* stream) {
  if (!x || !y || C % 8) return fail(COBEVT_ERR_SHAPE, "upsample_bilinear2x: bad args");
  const int64_t total = (int64_t)N * 4 * H * W * (C / 8);
  cobevt::launch_k(upsample_bilinear2x_kernel, blocks_for(total, 256), 256, 0, ST(stream), (const uint16_t*)x, (uint16_t*)y, N, H, W, C, dtype);
  return cobevt::check_launch("upsample_bilinear2x");
}

extern "C" int cobevt_add_nearest_relu(void* y, const void* up, int32_t N, int32_t H, int32_t W, int32_t h, int32_t w,
                                        int32_t C, int32_t dtype, void* stream) {
  if (!y || !up || C % 8) return fail(COBEVT_ERR_SHAPE, "add_nearest_relu: bad args");
  const int64_t total = (int64_t)N * H * W * (C / 8);
  cobevt::launch_k(add_nearest_relu_kernel, blocks_for(total, 256), 256, 0, ST(stream), (uint16_t*)y, (const uint16_t*)up, N, H, W, h, w, C, dtype);
  return cobevt::check_launch("add_nearest_relu");
}

extern "C" int cobevt_invert_matrices(const float* src, float* dst, int32_t n, int32_t dim, void* stream) {
  if (!src || !dst || n <= 0 || (dim != 3 && dim != 4)) return fail(COBEVT_ERR_SHAPE, "invert_matrices: dim must be 3 or 4");
  if (dim == 3) cobevt::launch_k(invert_kernel<3>, blocks_for(n, 128), 128, 0, ST(stream), src, dst, n);
  else cobevt::launch_k(invert_kernel<4>, blocks_for(n, 128), 128, 0, ST(stream), src, dst, n);
  return cobevt::check_launch("invert_matrices");
}

extern "C" int cobevt_permute(const void* src, void* dst, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                               int32_t to_nhwc, int32_t src_dtype, int32_t dst_dtype, void* stream) {
  if (!src || !dst || Cpad < C) return fail(COBEVT_ERR_SHAPE, "permute: bad args");
  const int64_t total = to_nhwc ? (int64_t)N * H * W * Cpad : (int64_t)N * C * H * W;
  cobevt::launch_k(permute_kernel, blocks_for(total, 256), 256, 0, ST(stream), src, dst, N, C, H, W, Cpad, to_nhwc, src_dtype, dst_dtype);
  return cobevt::check_launch("permute");
}
