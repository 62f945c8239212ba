// Implicit-GEMM convolution / token GEMM for sm_100a.
//
//   D[128 x BN] (TMEM, fp32) = A[128 x K] * B[BN x K]^T,   K = KH*KW*Cin
//
//   A : gathered on the fly from the NHWC activation (no im2col buffer): 128 producer
//       threads issue 16-byte cp.async (zero-fill outside the image) straight into a
//       K-major SWIZZLE_128B stage, 8 lanes per 128-byte channel run -> coalesced.
//   B : weights pre-packed on the host as ready-made SWIZZLE_128B tile images; one
//       cp.async.bulk (TMA engine) per stage, completing on the stage's mbarrier.
//   MMA : one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per 64-wide k-block;
//       tcgen05.commit releases the stage / signals the epilogue.
//   Epilogue : 4 warps, one TMEM lane (= output row) per thread: +bias, +residual,
//       ReLU/GELU, optional row LayerNorm-normalised second output, 16-bit / fp32 stores.
//
// Replaces the cuDNN/cuBLAS calls behind torchvision ResNet (resnet_ms.py:46-89), the
// FAX Bottleneck/downsample convs (fax_modules.py:472-489) and every nn.Linear of the
// token path (see include/cobevt_b200.h).
#include "sm100.cuh"
#include "common.h"
#include <cstring>
#include <vector>

using namespace sm100;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_PRODUCERS = 128;
constexpr int NUM_THREADS = 256;
constexpr int LAG = 2;  // cp.async groups kept in flight per producer thread
constexpr int kMaxDynSmem = 227 * 1024 - 2048;  // static smem (barriers) counts against the 227 KB

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 128) ? 3 : 4;
  static constexpr int kBStage = BN * BLOCK_K * 2;
  static constexpr int kStageBytes = A_STAGE_BYTES + kBStage;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kMinBlocks = (BN == 256) ? 1 : 2;
};

struct RowInfo {
  int32_t pixbase;  // n * Hin * Win
  int32_t iy0, ix0; // top-left logical input coordinate of the receptive field
  int32_t valid;
};
struct ChunkInfo {
  int16_t ky, kx;
  int16_t coff;
  int16_t valid;
};

struct KArgs {
  CobevtConvParams p;
  int32_t M, K, KB, n_tiles;
  uint32_t idesc;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

__device__ __forceinline__ uint32_t pack2(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u, int bf16) {
  if (bf16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, Cfg<BN>::kMinBlocks)
conv_gemm_kernel(const KArgs a) {
  using C = Cfg<BN>;
  constexpr int S = C::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages: A|B] [row table] [chunk table] [barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  RowInfo* rows = reinterpret_cast<RowInfo*>(smem + S * C::kStageBytes);
  ChunkInfo* chunks = reinterpret_cast<ChunkInfo*>(rows + BLOCK_M);
  __shared__ uint64_t full_bar[S], empty_bar[S], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const CobevtConvParams& p = a.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x % a.n_tiles;
  const int m_tile = blockIdx.x / a.n_tiles;
  const int KB = a.KB;
  const int Hlog = p.in_upsample ? p.Hin * 2 : p.Hin;
  const int Wlog = p.in_upsample ? p.Win * 2 : p.Win;

  // ---- one-time setup -------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], NUM_PRODUCERS + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(&tmem_base_s, C::kTmemCols);
    tmem_relinquish();
  }
  if (tid < BLOCK_M) {
    const int m = m_tile * BLOCK_M + tid;
    RowInfo ri;
    ri.valid = m < a.M;
    const int mm = ri.valid ? m : 0;
    const int ohw = p.OH * p.OW;
    const int n = mm / ohw;
    const int rem = mm - n * ohw;
    const int oy = rem / p.OW, ox = rem - oy * p.OW;
    ri.pixbase = n * p.Hin * p.Win;
    ri.iy0 = oy * p.stride_h - p.pad_h;
    ri.ix0 = ox * p.stride_w - p.pad_w;
    rows[tid] = ri;
  }
  for (int q = tid; q < KB * 8; q += NUM_THREADS) {
    const int k = q * 8;
    ChunkInfo ci;
    ci.valid = k < a.K;
    const int kk = ci.valid ? k : 0;
    const int tap = kk / p.Cin;
    ci.coff = (int16_t)(kk - tap * p.Cin);
    ci.ky = (int16_t)(tap / p.KW);
    ci.kx = (int16_t)(tap - ci.ky * p.KW);
    chunks[q] = ci;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // =========================== A producers ================================
    const int c8 = tid & 7;
    const int r0 = tid >> 3;
    int pixbase[8], iy0[8], ix0[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const RowInfo ri = rows[r0 + 16 * i];
      pixbase[i] = ri.pixbase;
      iy0[i] = ri.valid ? ri.iy0 : -(1 << 28);
      ix0[i] = ri.ix0;
    }
    const uint8_t* xbytes = reinterpret_cast<const uint8_t*>(p.x);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (kb / S) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      const ChunkInfo ci = chunks[kb * 8 + c8];
      uint8_t* sA = stage_base + s * C::kStageBytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + 16 * i;
        int iy = iy0[i] + ci.ky, ix = ix0[i] + ci.kx;
        const bool ok = ci.valid && iy >= 0 && iy < Hlog && ix >= 0 && ix < Wlog;
        if (p.in_upsample) { iy >>= 1; ix >>= 1; }
        const int64_t off = ((int64_t)(pixbase[i] + iy * p.Win + ix) * p.Cin + ci.coff) * 2;
        cp_async16(sA + sw128_off(r, c8), ok ? xbytes + off : xbytes, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (kb >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async_smem();
        mbar_arrive(&full_bar[(kb - LAG) % S]);
      }
    }
    // drain
    if (KB >= 2) {
      cp_async_wait<1>();
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[(KB - 2) % S]);
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    mbar_arrive(&full_bar[(KB - 1) % S]);
  } else {
    if (warp == 4 && lane == 0) {
      // =========================== B loader (TMA bulk) ======================
      const uint8_t* wb = reinterpret_cast<const uint8_t*>(p.w) + (int64_t)n_tile * KB * C::kBStage;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], C::kBStage);
        bulk_g2s(stage_base + s * C::kStageBytes + A_STAGE_BYTES, wb + (int64_t)kb * C::kBStage,
                 C::kBStage, &full_bar[s]);
      }
    } else if (warp == 5 && lane == 0) {
      // =========================== MMA issuer ===============================
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(stage_base + s * C::kStageBytes);
        const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BLOCK_K / 16; ++ks) {
          const uint64_t da = make_smem_desc(sa + ks * 32, 16, 1024, LAYOUT_SW128);
          const uint64_t db = make_smem_desc(sb + ks * 32, 16, 1024, LAYOUT_SW128);
          umma_f16(tmem_base, da, db, a.idesc, (kb | ks) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full_bar);
    }
    __syncwarp();
    // =========================== epilogue (warps 4..7) ======================
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after_sync();
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int64_t m = (int64_t)m_tile * BLOCK_M + row;
    const bool row_ok = m < a.M;
    const int bf16 = p.dtype;
    const int n_base = n_tile * BN;
    const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
    constexpr int G = (BN >= 32) ? 32 : 16;   // columns per TMEM load
    constexpr int NG = BN / G;

    float ln_shift = 0.f, ln_sum = 0.f, ln_sq = 0.f;
    float mean = 0.f, rstd = 0.f;
    const int passes = p.y_ln ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        uint32_t r[G];
        if constexpr (G == 32) tmem_ld_32x32b_x32(trow + g * G, r);
        else tmem_ld_32x32b_x16(trow + g * G, r);
        tmem_ld_wait();
        if (!row_ok) continue;
        const int n0 = n_base + g * G;
        float v[G];
#pragma unroll
        for (int j = 0; j < G; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < G; ++j) if (n0 + j < p.Cout) v[j] += __ldg(p.bias + n0 + j);
        }
        const int64_t mr = p.res_rows > 0 ? (m % p.res_rows) : m;
        if (p.res) {
          const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.res) + (mr * p.Cout + n0) * 2;
#pragma unroll
          for (int j = 0; j < G; j += 8) {
            if (n0 + j + 8 <= p.Cout) {
              const uint4 u = *reinterpret_cast<const uint4*>(rp + j * 2);
              const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = unpack2(w4[t], bf16);
                v[j + 2 * t] += f.x;
                v[j + 2 * t + 1] += f.y;
              }
            }
          }
        }
        if (p.res32) {
          const float* rp = p.res32 + mr * p.Cout + n0;
#pragma unroll
          for (int j = 0; j < G; j += 4) {
            if (n0 + j + 4 <= p.Cout) {
              const float4 f = *reinterpret_cast<const float4*>(rp + j);
              v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
            }
          }
        }
        if (p.act == COBEVT_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < G; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == COBEVT_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < G; ++j) v[j] = gelu_erf(v[j]);
        }
        if (pass == 0) {
          if (p.y_ln) {
            if (g == 0) ln_shift = v[0];
#pragma unroll
            for (int j = 0; j < G; ++j) {
              const float d = v[j] - ln_shift;
              ln_sum += d;
              ln_sq += d * d;
            }
          }
          if (p.y) {
            uint8_t* yp = reinterpret_cast<uint8_t*>(p.y) + (m * p.Cout + n0) * 2;
#pragma unroll
            for (int j = 0; j < G; j += 8) {
              if (n0 + j + 8 <= p.Cout) {
                uint4 u;
                u.x = pack2(v[j], v[j + 1], bf16);
                u.y = pack2(v[j + 2], v[j + 3], bf16);
                u.z = pack2(v[j + 4], v[j + 5], bf16);
                u.w = pack2(v[j + 6], v[j + 7], bf16);
                *reinterpret_cast<uint4*>(yp + j * 2) = u;
              } else {
                for (int t = 0; t < 8; ++t)
                  if (n0 + j + t < p.Cout) {
                    const uint32_t u = pack2(v[j + t], 0.f, bf16);
                    *reinterpret_cast<uint16_t*>(yp + (j + t) * 2) = (uint16_t)(u & 0xFFFFu);
                  }
              }
            }
          }
          if (p.y32) {
            float* yp = p.y32 + m * p.Cout + n0;
#pragma unroll
            for (int j = 0; j < G; j += 4) {
              if (n0 + j + 4 <= p.Cout) {
                *reinterpret_cast<float4*>(yp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
                for (int t = 0; t < 4; ++t) if (n0 + j + t < p.Cout) yp[j + t] = v[j + t];
              }
            }
          }
        } else {
          uint8_t* yp = reinterpret_cast<uint8_t*>(p.y_ln) + (m * p.Cout + n0) * 2;
#pragma unroll
          for (int j = 0; j < G; ++j) v[j] = (v[j] - mean) * rstd;
          if (p.ln_gamma) {
#pragma unroll
            for (int j = 0; j < G; ++j) v[j] = fmaf(v[j], __ldg(p.ln_gamma + n0 + j), __ldg(p.ln_beta + n0 + j));
          }
#pragma unroll
          for (int j = 0; j < G; j += 8) {
            uint4 u;
            u.x = pack2(v[j], v[j + 1], bf16);
            u.y = pack2(v[j + 2], v[j + 3], bf16);
            u.z = pack2(v[j + 4], v[j + 5], bf16);
            u.w = pack2(v[j + 6], v[j + 7], bf16);
            *reinterpret_cast<uint4*>(yp + j * 2) = u;
          }
        }
      }
      if (pass == 0 && p.y_ln) {
        const float inv = 1.0f / (float)BN;
        const float dm = ln_sum * inv;
        mean = ln_shift + dm;
        const float var = fmaxf(ln_sq * inv - dm * dm, 0.f);
        rstd = rsqrtf(var + p.ln_eps);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BN>
int launch(const KArgs& a, cudaStream_t st) {
  using C = Cfg<BN>;
  const size_t smem = 1024 + (size_t)C::kStages * C::kStageBytes + BLOCK_M * sizeof(RowInfo) +
                      (size_t)a.KB * 8 * sizeof(ChunkInfo);
  if (smem > kMaxDynSmem) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: smem %zu too large (K=%d)", smem, a.K);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (e != cudaSuccess) return cobevt::fail(COBEVT_ERR_CUDA, "conv: set smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int m_tiles = cobevt::ceil_div(a.M, BLOCK_M);
  conv_gemm_kernel<BN><<<m_tiles * a.n_tiles, NUM_THREADS, smem, st>>>(a);
  return cobevt::check_launch("conv_gemm_kernel");
}

}  // namespace

extern "C" int64_t cobevt_conv_packed_bytes(int32_t Cout, int32_t K, int32_t block_n) {
  const int64_t n_tiles = (Cout + block_n - 1) / block_n;
  const int64_t KB = (K + BLOCK_K - 1) / BLOCK_K;
  return n_tiles * KB * block_n * BLOCK_K * 2;
}

extern "C" int cobevt_conv_pack_weight(const uint16_t* w, int32_t Cout, int32_t K, int32_t block_n,
                                       uint16_t* out) {
  if (!w || !out || Cout <= 0 || K <= 0) return cobevt::fail(COBEVT_ERR_SHAPE, "pack: bad args");
  if (block_n != 16 && block_n != 32 && block_n != 64 && block_n != 128 && block_n != 256)
    return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "pack: block_n %d", block_n);
  const int n_tiles = (Cout + block_n - 1) / block_n;
  const int KB = (K + BLOCK_K - 1) / BLOCK_K;
  memset(out, 0, (size_t)cobevt_conv_packed_bytes(Cout, K, block_n));
  uint8_t* ob = reinterpret_cast<uint8_t*>(out);
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int kb = 0; kb < KB; ++kb) {
      uint8_t* img = ob + ((size_t)nt * KB + kb) * block_n * BLOCK_K * 2;
      for (int r = 0; r < block_n; ++r) {
        const int co = nt * block_n + r;
        if (co >= Cout) continue;
        for (int kk = 0; kk < BLOCK_K; ++kk) {
          const int k = kb * BLOCK_K + kk;
          if (k >= K) break;
          const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)kk >> 3) ^ ((uint32_t)r & 7u)) << 4) + (kk & 7) * 2;
          memcpy(img + off, &w[(size_t)co * K + k], 2);
        }
      }
    }
  return COBEVT_OK;
}

extern "C" int cobevt_conv2d_nhwc_fwd(const CobevtConvParams* p, void* stream) {
  if (!p || !p->x || !p->w) return cobevt::fail(COBEVT_ERR_SHAPE, "conv: null pointer");
  if (p->Cin % 8 != 0) return cobevt::fail(COBEVT_ERR_SHAPE, "conv: Cin %d not a multiple of 8", p->Cin);
  if (p->N <= 0 || p->OH <= 0 || p->OW <= 0 || p->Cout <= 0 || p->KH <= 0 || p->KW <= 0 || p->stride_h <= 0 || p->stride_w <= 0)
    return cobevt::fail(COBEVT_ERR_SHAPE, "conv: bad extents");
  if (p->dtype != COBEVT_F16 && p->dtype != COBEVT_BF16) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: dtype %d", p->dtype);
  if (!p->y && !p->y32 && !p->y_ln) return cobevt::fail(COBEVT_ERR_SHAPE, "conv: no output");
  if (p->y_ln && !(p->Cout == 128 && p->block_n == 128))
    return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: y_ln needs Cout == block_n == 128");
  if ((p->ln_gamma == nullptr) != (p->ln_beta == nullptr)) return cobevt::fail(COBEVT_ERR_SHAPE, "conv: ln_gamma/ln_beta must come together");
  if (p->Cout % 8 != 0 && (p->res || p->y_ln)) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: Cout %% 8 with res/ln");
  const int64_t M64 = (int64_t)p->N * p->OH * p->OW;
  if (M64 > (1ll << 30)) return cobevt::fail(COBEVT_ERR_SHAPE, "conv: M too large");
  KArgs a;
  a.p = *p;
  a.M = (int32_t)M64;
  a.K = p->KH * p->KW * p->Cin;
  a.KB = cobevt::ceil_div(a.K, BLOCK_K);
  a.n_tiles = cobevt::ceil_div(p->Cout, p->block_n);
  a.idesc = make_idesc_f16(128, p->block_n, p->dtype == COBEVT_BF16, false, false);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (p->block_n) {
    case 16: return launch<16>(a, st);
    case 32: return launch<32>(a, st);
    case 64: return launch<64>(a, st);
    case 128: return launch<128>(a, st);
    case 256: return launch<256>(a, st);
    default: return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: block_n %d", p->block_n);
  }
}
