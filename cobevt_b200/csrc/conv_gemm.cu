// Implicit-GEMM convolution / token GEMM for sm_100a -- persistent, warp-specialised.
//
//   D[128 x BN] (TMEM, fp32) = A[128 x K] * B[BN x K]^T,   K = KH*KW*Cin
//
// One CTA (20 warps) per SM walks over output tiles (128 rows x BN channels).  Warp roles:
//   warps 4-11     epilogue team 0: TMEM lane quarter (warp % 4) x column half; one output row per thread: all
//                  tcgen05.ld of the unit issued back to back, the accumulator is released as soon as it is in
//                  registers, then +bias (smem), +residual, ReLU / exact GELU, optional row LayerNorm second output
//                  (the two column halves exchange mean / M2 through smem); token GEMMs leave through a swizzled
//                  staging tile + one TMA tensor store per warp, convs store 16 bytes per thread
//   warps 0-3,     epilogue team 1 when the A operand comes from TMA (the teams take alternate tiles = alternate
//         12-15    TMEM accumulators, 4 per CTA, 2 for BN = 256); otherwise warps 0-3 are the cp.async A producers
//   warp  16       loader (one thread): weights pre-packed on the host as SWIZZLE_128B tile images by cp.async.bulk
//                  -- ALL k-blocks once when they fit (resident weights) -- and the A ring by TMA tensor loads
//   warp  17       MMA issuer: the whole warp walks the pipeline, the elected lane issues tcgen05.mma M=128 N=BN K=16
//                  and the tcgen05.commit that frees stages / publishes accumulators
//   warps 18-19    idle (they complete the control warpgroup for setmaxnreg: 104 registers for the 16 workers)
//
// Instances: conv_gemm_kernel<BN, RT>.  RT = true (BN = 128 token GEMMs with a plain 16-bit residual): the residual tile
// of a warp arrives in its staging buffer by TMA one tile ahead instead of by per-thread loads after the accumulator wait.
// All kernels are launched with programmatic dependent launch (common.h launch_k): prologue and resident weights are
// fetched before griddepcontrol.wait, activations only after it.  The warp index is broadcast from lane 0
// (sm100.cuh warp_index_uniform) so that role / tile state lives in uniform registers.
//
// A-operand modes (host-selected, see cobevt_conv2d_nhwc_fwd)
//   HALO    3x3 / stride 1 / pad 1 / Cin % 64 == 0 (the bulk of the ResNet encoder): the output tile is 16 rows x 8
//           pixels; its 18 x 10 x 64ch input halo is ONE TMA box per 64-channel slab and all nine taps read it in
//           place through shifted UMMA descriptors (start = slab + (ky*10+kx)*128 B, SBO = 10*128 B).  The
//           SWIZZLE_128B XOR is a function of absolute smem address bits (profiles/r01_umma_probe_halo.log), so
//           shifted starts need no re-layout.
//   LINEAR  1x1 / stride 1 (every nn.Linear): one 2-D TMA box (128 rows x 64 channels) per k-block.
//   TAP     other strides / kernel shapes with Cin in {8, 32, 64k}: one strided TMA box per filter tap.
//   PATCH   8-channel input, stride_w = 1 (the stem): the unswizzled input patch is the A operand in place
//           (overlapping 8 x 16 B core matrices, LBO = 16 B).
//   GATHER  fallback: per-thread cp.async im2col slices (nearest-x2-upsampled inputs, odd channel counts).
//
// Replaces the cuDNN/cuBLAS calls behind torchvision ResNet (resnet_ms.py:46-89), the FAX
// Bottleneck/downsample convs (fax_modules.py:472-489) and every nn.Linear of the token path
// (see include/cobevt_b200.h).
#include "sm100.cuh"
#include "common.h"
#include <cstring>
#include <cstdlib>

using namespace sm100;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_GATHER_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int HALO_H = 18, HALO_W = 10, TILE_H = 16, TILE_W = 8;
constexpr int A_HALO_BYTES = 24 * 1024;                // 180 rows x 128 B = 23040, padded to 1 KB multiple
constexpr int NUM_THREADS = 640;                       // 20 warps, see the role table in conv_gemm_kernel
constexpr int REGS_WORKERS = 104, REGS_CONTROL = 40;   // setmaxnreg split: 16 worker warps / loader + MMA warpgroup (pool 640 x 96)
constexpr int SA_GATHER = 6, SA_HALO = 4;       // A ring depth (all stages may be in flight)
constexpr int kMaxDynSmem = 227 * 1024 - 12288;        // static smem (barriers, bias, LN exchange) counts too
constexpr int MAX_STAGES = 8;
constexpr int STG_WARP_BYTES = 4096;                   // TMA-store staging per epilogue warp: 32 rows x <= 128 B

enum { MODE_GATHER = 0, MODE_HALO = 1, MODE_LINEAR = 2, MODE_TAP = 3, MODE_PATCH = 4 };
enum { A_SW128 = 0, A_SW64 = 1, A_NOSWZ = 2 };          // smem image of a TAP-mode A stage

struct KArgs {
  CobevtConvParams p;
  int32_t M, K, KB, n_tiles, m_tiles, total_tiles;
  int32_t mode, sa, sb, resident_b;
  int32_t tiles_y, tiles_x, nslab;      // HALO geometry
  int32_t cin_log2, kw_magic;           // GATHER tap decode without divisions (cin_log2 < 0: generic)
  int32_t use_tma;                      // A operand through TMA tensor loads (HALO / LINEAR / TAP) instead of per-thread cp.async
  int32_t tap_c, taps_per_kb, slabs_per_tap, a_layout;   // TAP: channels per tensor load, taps per 64-wide K block, ...
  int32_t patch_w, patch_h;             // PATCH: input patch (pixels) of one 16 x 8 output tile
  int32_t ntaps, tap_kw;                // TAP: tensor loads per tile and taps per filter row (1 when a row's taps are fused)
  int32_t dbg;                          // COBEVT_CONV_DEBUG bits (profiling experiments only): 1 no epilogue I/O, 2 no A loads, 4 no MMA
  uint32_t idesc;
  int32_t st_tma;                       // 16-bit outputs leave through per-warp smem tiles + TMA tensor stores
  int32_t res_tma;                      // the 16-bit residual arrives through the same staging tiles by TMA tensor loads, one tile ahead
  alignas(64) CUtensorMap tmY, tmYln;   // output maps: {Cout, M} box {U, 32}  or  {Cout, OW, OH, N} box {U, 8, 4, 1}
  alignas(64) CUtensorMap tmRes;        // residual map (same geometry as tmY) when res_tma
  alignas(64) CUtensorMap tmap;         // HALO: {C, W, H, N} box {64, 10, 18, 1}; LINEAR: {K, M} box {64, 128}; SWIZZLE_128B
};

// exact (erf) GELU of the reference (nn.GELU default).  erf by Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7, far below the 16-bit output rounding) = 2 MUFU + ~10 FMA instead of erff's ~30.
// Evaluated on pairs: packed fp32 (FFMA2 / FMUL2) halves the issue slots of the polynomial; 4 MUFU per pair.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  float z0, z1, t0, t1, p0, p1, q0, q1;
  fmul2(z0, z1, a0, a1, 0.70710678118654752f, 0.70710678118654752f);
  ffma2(t0, t1, z0, z1, 0.3275911f, 0.3275911f, 1.0f, 1.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(t0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(t1));
  ffma2(p0, p1, t0, t1, 1.061405429f, 1.061405429f, -1.453152027f, -1.453152027f);
  ffma2(p0, p1, p0, p1, t0, t1, 1.421413741f, 1.421413741f);
  ffma2(p0, p1, p0, p1, t0, t1, -0.284496736f, -0.284496736f);
  ffma2(p0, p1, p0, p1, t0, t1, 0.254829592f, 0.254829592f);
  fmul2(q0, q1, z0, z1, z0, z1);
  fmul2(q0, q1, q0, q1, -1.4426950408889634f, -1.4426950408889634f);      // exp(-z^2) = 2^(-z^2 log2 e)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q1) : "f"(q1));
  fmul2(p0, p1, p0, p1, t0, t1);
  fmul2(p0, p1, p0, p1, q0, q1);
  float e0 = 1.0f - p0, e1 = 1.0f - p1;                                    // erf(|x| / sqrt 2)
  fmul2(t0, t1, a0, a1, 0.5f, 0.5f);
  fmul2(x0, x1, x0, x1, 0.5f, 0.5f);
  ffma2(x0, x1, t0, t1, e0, e1, x0, x1);                                    // 0.5 x + 0.5 |x| erf
}
__device__ __forceinline__ uint32_t pack2(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// max(x, 0) on a packed pair
__device__ __forceinline__ uint32_t relu2(uint32_t u, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&u), __float2bfloat162_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __hmax2(*reinterpret_cast<__half2*>(&u), __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u, int bf16) {
  if (bf16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// tcgen05.mma with descriptors given as (lo, hi) 32-bit halves: only `lo` (start address field)
// changes between MMAs, so the issuing thread does one integer add per operand.
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                          uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// output row (TMEM lane) -> linear output pixel index, or -1
__device__ __forceinline__ int64_t out_row_index(const KArgs& a, int m_tile, int row) {
  const CobevtConvParams& p = a.p;
  if (a.mode == MODE_HALO || a.mode == MODE_TAP || a.mode == MODE_PATCH) {
    const int per_img = a.tiles_y * a.tiles_x;
    const int n = m_tile / per_img;
    const int t = m_tile - n * per_img;
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    const int oy = ty * TILE_H + (row >> 3), ox = tx * TILE_W + (row & 7);
    if (oy >= p.OH || ox >= p.OW) return -1;
    return ((int64_t)n * p.OH + oy) * p.OW + ox;
  }
  const int64_t m = (int64_t)m_tile * BLOCK_M + row;
  return m < a.M ? m : -1;
}

// one epilogue unit: U accumulator columns of this thread's row
template <int U>
__device__ __forceinline__ void tmem_ld_unit(uint32_t taddr, uint32_t (&r)[U]) {
  if constexpr (U == 16) {
    tmem_ld_32x32b_x16(taddr, r);
  } else {
#pragma unroll
    for (int g = 0; g < U / 32; ++g) {
      uint32_t t[32];
      tmem_ld_32x32b_x32(taddr + g * 32, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) r[g * 32 + j] = t[j];
    }
  }
}

// RT: residual by TMA (see the epilogue); a separate instance so that every other launch keeps its register allocation
template <int BN, bool RT>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_gemm_kernel(const __grid_constant__ KArgs a) {
  constexpr int kBStage = BN * BLOCK_K * 2;
  constexpr int NACC = BN >= 256 ? 2 : 4;         // accumulator buffers in TMEM (power of two)
  constexpr int kTmemCols = (NACC * BN) < 32 ? 32 : NACC * BN;
  constexpr int CW = BN >= 64 ? BN / 2 : BN;      // columns per epilogue warp
  constexpr int U = CW > 64 ? 64 : CW;            // columns per unit
  constexpr int NU = CW / U;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int SA = a.sa, SB = a.sb;
  const int a_stage_bytes = a.mode == MODE_HALO ? A_HALO_BYTES : A_GATHER_BYTES;
  uint8_t* sA = smem;
  uint8_t* sB = smem + SA * a_stage_bytes;
  uint8_t* sStage = sB + (a.resident_b ? a.KB * kBStage : SB * kBStage);   // 16 x STG_WARP_BYTES when a.st_tma
  __shared__ uint64_t a_full[MAX_STAGES], a_empty[MAX_STAGES], b_full[MAX_STAGES], b_empty[MAX_STAGES];
  __shared__ uint64_t acc_full[4], acc_empty[4], b_res_full;
  __shared__ uint64_t res_full[16];               // per epilogue warp: its residual tile has landed in the staging buffer
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sBias[512], sGamma[128], sBeta[128];
  __shared__ float2 sXch[2][2][2][128];           // [team][tile parity][column half][row] = (mean, M2)

  const CobevtConvParams& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = warp_index_uniform();                  // see sm100.cuh: keeps role / counter state in uniform registers
  const int KB = a.KB;

  // Programmatic dependent launch: the next kernel may be scheduled as soon as every CTA of this grid is resident;
  // this CTA's prologue (barriers, TMEM, constant bias / LN vectors, resident weights) runs before pdl_wait(), i.e.
  // possibly while the previous kernel of the stream is still draining.  Activations are touched only after it.
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&a_full[s], a.use_tma ? 1 : 128);
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(&b_res_full, 1);
    if (RT) for (int i = 0; i < 16; ++i) mbar_init(&res_full[i], 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  for (int i = tid; i < 512; i += NUM_THREADS) sBias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
  if (tid < 128) {
    sGamma[tid] = (p.ln_gamma && tid < p.Cout) ? p.ln_gamma[tid] : 1.f;
    sBeta[tid] = (p.ln_beta && tid < p.Cout) ? p.ln_beta[tid] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  // ---- roles (20 warps) ------------------------------------------------------------------------
  //   warps  4-11   epilogue team 0 (TMEM lane quarter = warp % 4, column half = (warp - 4) / 4)
  //   warps  0- 3   A operand by TMA : epilogue team 1, column half 0  |  otherwise: cp.async A producers
  //   warps 12-15   A operand by TMA : epilogue team 1, column half 1  |  otherwise: idle
  //   warp  16      loader: weights (bulk copies) and, in the TMA modes, the A ring (tensor loads)
  //   warp  17      MMA issuer            warps 18-19 idle (complete the control warpgroup for setmaxnreg)
  // The two epilogue teams take alternate tiles (= alternate TMEM accumulators), so two epilogues are in
  // flight: the epilogue of the token GEMMs is bound by per-warp latency, not by any pipe.
  const int n_teams = a.use_tma ? 2 : 1;
  int team = -1, e = 0;
  if (warp >= 4 && warp < 12) { team = 0; e = warp - 4; }
  else if (a.use_tma && warp < 4) { team = 1; e = warp; }
  else if (a.use_tma && warp >= 12 && warp < 16) { team = 1; e = warp - 8; }

  if (warp < 4 && !a.use_tma) {
    // ================================ A producers ========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_WORKERS));
    pdl_wait();
    const uint8_t* xbytes = reinterpret_cast<const uint8_t*>(p.x);
    uint32_t it = 0;   // stage uses issued so far (across tiles)
    int ps = 0;        // stage of use `it`            (incremental: no runtime divisions in the loops)
    uint32_t pph = 1;  // parity to wait for on a_empty[ps]
    auto next_stage = [&]() { if (++ps == SA) { ps = 0; pph ^= 1u; } };
    if (a.mode == MODE_LINEAR) {
      // 1x1 / stride 1: A rows are contiguous token rows, no coordinate arithmetic at all
      const int c8 = tid & 7, r0 = tid >> 3;
      const int64_t row_bytes = (int64_t)p.Cin * 2;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles;
        const int m0 = m_tile * BLOCK_M;
        const uint8_t* base = xbytes + (int64_t)m0 * row_bytes + c8 * 16;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = ps;
          mbar_wait(&a_empty[s], pph);
          next_stage();
          const bool kvalid = kb * BLOCK_K + c8 * 8 < a.K;
          uint8_t* dst = sA + s * A_GATHER_BYTES;
          const uint8_t* src = base + kb * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r0 + 16 * i;
            const bool ok = kvalid && (m0 + r) < a.M;
            if (!COBEVT_DBG(a, 2)) cp_async16(dst + sw128_off(r, c8), ok ? src + r * row_bytes : xbytes, ok ? 16u : 0u);
          }
          cp_async_mbar_arrive_noinc(&a_full[s]);   // async arrival when this stage's bytes have landed
        }
      }
    } else if (a.mode == MODE_GATHER) {
      const int Hlog = p.in_upsample ? p.Hin * 2 : p.Hin;
      const int Wlog = p.in_upsample ? p.Win * 2 : p.Win;
      const int c8 = tid & 7, r0 = tid >> 3;
      const int ohw = p.OH * p.OW;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles;
        int pixbase[8], iy0[8], ix0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m_tile * BLOCK_M + r0 + 16 * i;
          const bool ok = m < a.M;
          const int mm = ok ? m : 0;
          const int n = mm / ohw;
          const int rem = mm - n * ohw;
          const int oy = rem / p.OW, ox = rem - oy * p.OW;
          pixbase[i] = n * p.Hin * p.Win;
          iy0[i] = ok ? oy * p.stride_h - p.pad_h : -(1 << 28);
          ix0[i] = ox * p.stride_w - p.pad_w;
        }
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = ps;
          mbar_wait(&a_empty[s], pph);
          next_stage();
          // chunk -> (tap, channel offset)
          const int k = kb * BLOCK_K + c8 * 8;
          const bool kvalid = k < a.K;
          const int kk = kvalid ? k : 0;
          int tap, coff, ky;
          if (a.cin_log2 >= 0) {      // power-of-two Cin, small KW: shifts + multiply-high
            tap = kk >> a.cin_log2;
            coff = kk & (p.Cin - 1);
            ky = (tap * a.kw_magic) >> 16;
          } else {
            tap = kk / p.Cin;
            coff = kk - tap * p.Cin;
            ky = tap / p.KW;
          }
          const int kx = tap - ky * p.KW;
          uint8_t* dst = sA + s * A_GATHER_BYTES;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r0 + 16 * i;
            int iy = iy0[i] + ky, ix = ix0[i] + kx;
            const bool ok = kvalid && iy >= 0 && iy < Hlog && ix >= 0 && ix < Wlog;
            if (p.in_upsample) { iy >>= 1; ix >>= 1; }
            const int64_t off = ((int64_t)(pixbase[i] + iy * p.Win + ix) * p.Cin + coff) * 2;
            if (!COBEVT_DBG(a, 2)) cp_async16(dst + sw128_off(r, c8), ok ? xbytes + off : xbytes, ok ? 16u : 0u);
          }
          cp_async_mbar_arrive_noinc(&a_full[s]);   // async arrival when this stage's bytes have landed
        }
      }
    } else {
      // HALO: 180 halo pixels x 8 chunks = 1440 pieces per 64-channel slab, 12 per thread
      const int per_img = a.tiles_y * a.tiles_x;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles;
        const int n = m_tile / per_img;
        const int t = m_tile - n * per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        int32_t poff[12];   // element offset of the halo pixel (channel 0) or -1
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int idx = tid + 128 * i;
          const int pix = idx >> 3;
          const int hy = pix / HALO_W, hx = pix - hy * HALO_W;
          const int iy = ty * TILE_H - 1 + hy, ix = tx * TILE_W - 1 + hx;
          const bool ok = pix < HALO_H * HALO_W && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
          poff[i] = ok ? ((n * p.Hin + iy) * p.Win + ix) * p.Cin : -1;
        }
        for (int slab = 0; slab < a.nslab; ++slab, ++it) {
          const int s = ps;
          mbar_wait(&a_empty[s], pph);
          next_stage();
          uint8_t* dst = sA + s * A_HALO_BYTES;
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const int idx = tid + 128 * i;
            const int pix = idx >> 3, chunk = idx & 7;
            if (pix < HALO_H * HALO_W) {
              const bool ok = poff[i] >= 0;
              const int64_t off = ((int64_t)(ok ? poff[i] : 0) + slab * 64 + chunk * 8) * 2;
              if (!COBEVT_DBG(a, 2)) cp_async16(dst + pix * 128 + ((chunk ^ (pix & 7)) << 4), xbytes + (ok ? off : 0), ok ? 16u : 0u);
            }
          }
          cp_async_mbar_arrive_noinc(&a_full[s]);   // async arrival when this stage's bytes have landed
        }
      }
    }
  } else if (team >= 0) {
    // ================================ epilogue (8 warps per team) ==========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_WORKERS));
    pdl_wait();                       // residual reads and output stores follow the previous kernel
    const int q4 = e & 3;             // TMEM lane quarter (== warp % 4)
    const int hf = e >> 2;            // column half
    const bool active = (BN >= 64) || (hf == 0);
    const int row = q4 * 32 + lane;
    const int bf16 = p.dtype;
    const bool do_ln = p.y_ln != nullptr;      // host guarantees BN == Cout == 128 -> CW == U == 64
    uint32_t acc_it = 0;
    // Residual by TMA (token GEMMs): the warp's 32 x 64 residual tile is requested into its staging buffer BEFORE the
    // accumulator wait -- right after the previous tile's stores have left the buffer, i.e. about one tile ahead of its
    // use -- so the epilogue reads it from shared memory instead of stalling on 8 global loads per thread after the wait
    // (that stall was 25 % of this kernel's samples on the residual + LayerNorm GEMMs, profiles/r02_ncu_full.txt).
    const bool res_tma = RT && active && !COBEVT_DBG(a, 9);
    uint8_t* const stg_gen = sStage + (team * 8 + e) * STG_WARP_BYTES;
    uint64_t* const res_bar = &res_full[team * 8 + e];
    uint32_t res_par = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++acc_it) {
      const int m_tile = tile / a.n_tiles;
      const int n_tile = tile - m_tile * a.n_tiles;
      const int buf = acc_it & (NACC - 1);
      if ((int)(acc_it % (uint32_t)n_teams) != team) continue;      // the other team's tile
      if (RT && res_tma && lane == 0) {
        bulk_wait_read0();                                          // the previous tile's stores have read the buffer
        mbar_arrive_expect_tx(res_bar, 32 * U * 2);
        tma_load_2d(stg_gen, &a.tmRes, res_bar, n_tile * BN + hf * CW, m_tile * BLOCK_M + q4 * 32);
      }
      mbar_wait(&acc_full[buf], (acc_it / NACC) & 1);
      tc_fence_after_sync();
      if (!active) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      if (COBEVT_DBG(a, 8)) {   // profiling experiment: no TMEM read / math at all
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      const int64_t m = out_row_index(a, m_tile, row);
      const bool row_ok = m >= 0;
      const int64_t mr = p.res_rows > 0 ? (int64_t)((uint32_t)(row_ok ? m : 0) % (uint32_t)p.res_rows) : m;
      const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * BN + hf * CW;
      // TMA-store path: the warp's 32 x U tile is written to its staging buffer in the output map's swizzle
      // (inner box = U * 2 bytes: SWIZZLE_128B / 64B / 32B) and leaves as one tensor store (full lines, clipped
      // at the tensor bounds, no store instructions or address arithmetic in this warp).
      const uint32_t stg = smem_u32(sStage) + (uint32_t)(team * 8 + e) * STG_WARP_BYTES;
      const bool st_tma = a.st_tma && !COBEVT_DBG(a, 1);
      auto tma_out = [&](const CUtensorMap* tm, int n0c, bool relu, const float (&vv)[U]) {
        constexpr int CH16 = U / 8;                     // 16-byte pieces per row
        constexpr int RPL = 128 / (U * 2);              // rows per 128-byte line
        if (lane == 0) bulk_wait_read0();               // the previous store has finished reading the buffer
        __syncwarp();
#pragma unroll
        for (int ck = 0; ck < CH16; ++ck) {
          uint32_t w0 = pack2(vv[ck * 8 + 0], vv[ck * 8 + 1], bf16), w1 = pack2(vv[ck * 8 + 2], vv[ck * 8 + 3], bf16);
          uint32_t w2 = pack2(vv[ck * 8 + 4], vv[ck * 8 + 5], bf16), w3 = pack2(vv[ck * 8 + 6], vv[ck * 8 + 7], bf16);
          if (relu) { w0 = relu2(w0, bf16); w1 = relu2(w1, bf16); w2 = relu2(w2, bf16); w3 = relu2(w3, bf16); }
          sts128(stg + (uint32_t)(lane * (U * 2)) + ((((uint32_t)ck ^ ((uint32_t)(lane / RPL) & (CH16 - 1)))) << 4), w0, w1, w2, w3);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (a.mode == MODE_LINEAR || a.mode == MODE_GATHER) {
            tma_store_2d(tm, stg, n0c, m_tile * BLOCK_M + q4 * 32);
          } else {
            const int per_img = a.tiles_y * a.tiles_x;
            const int n = m_tile / per_img;
            const int t = m_tile - n * per_img;
            const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
            tma_store_4d(tm, stg, n0c, tx * TILE_W, ty * TILE_H + q4 * 4, n);
          }
          bulk_commit();
        }
      };
#pragma unroll 1
      for (int u = 0; u < NU; ++u) {
        const int col = hf * CW + u * U;           // column inside the tile
        const int n0 = n_tile * BN + col;          // global output channel
        uint32_t r[U];
        tmem_ld_unit<U>(trow + u * U, r);
        tmem_ld_wait();
        if (u == NU - 1) {                         // accumulator is in registers: hand the buffer back
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        const bool live = row_ok && n0 < p.Cout && !COBEVT_DBG(a, 1);
        float v[U];
        {
          const float4* b4 = reinterpret_cast<const float4*>(sBias + (n0 & 511));   // broadcast LDS.128
#pragma unroll
          for (int j = 0; j < U; j += 4) {
            const float4 bb = b4[j >> 2];
            v[j] = __uint_as_float(r[j]) + bb.x;
            v[j + 1] = __uint_as_float(r[j + 1]) + bb.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + bb.z;
            v[j + 3] = __uint_as_float(r[j + 3]) + bb.w;
          }
        }
        if (RT && res_tma) {
          mbar_wait(res_bar, res_par);
          res_par ^= 1u;
          if constexpr (U == 64) {                   // the tile sits in the output map's SWIZZLE_128B image: this lane's row
#pragma unroll
            for (int ck = 0; ck < 8; ++ck) {
              const uint4 rr = lds128(stg + (uint32_t)(lane * 128) + ((((uint32_t)ck ^ ((uint32_t)lane & 7u))) << 4));
              const uint32_t w4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const float2 f = unpack2(w4[k2], bf16);
                v[ck * 8 + 2 * k2] += f.x;
                v[ck * 8 + 2 * k2 + 1] += f.y;
              }
            }
          }
        } else if (live && p.res) {
          const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.res) + (mr * p.Cout + n0) * 2;
#pragma unroll
          for (int j0 = 0; j0 < U; j0 += 32) {
            uint4 rr[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
              rr[t] = (j0 + t * 8 < U && n0 + j0 + t * 8 + 8 <= p.Cout)
                          ? *reinterpret_cast<const uint4*>(rp + (j0 + t * 8) * 2) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (j0 + t * 8 < U) {
                const uint32_t w4[4] = {rr[t].x, rr[t].y, rr[t].z, rr[t].w};
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                  const float2 f = unpack2(w4[k2], bf16);
                  v[j0 + t * 8 + 2 * k2] += f.x;
                  v[j0 + t * 8 + 2 * k2 + 1] += f.y;
                }
              }
            }
          }
        }
        if (live && p.res32) {
          const float* rp = p.res32 + mr * p.Cout + n0;
#pragma unroll
          for (int j = 0; j < U; j += 4) {
            if (n0 + j + 4 <= p.Cout) {
              const float4 f = *reinterpret_cast<const float4*>(rp + j);
              v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
            }
          }
        }
        const bool relu_packed = p.act == COBEVT_ACT_RELU && !p.y32 && !do_ln;   // max(.,0) on packed pairs
        if (p.act == COBEVT_ACT_RELU && !relu_packed) {
#pragma unroll
          for (int j = 0; j < U; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == COBEVT_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < U; j += 2) gelu_erf2(v[j], v[j + 1]);
        } else if (p.act == COBEVT_ACT_SWISH) {        // x * sigmoid(x): MBConv expand / stem (efficientnet_pytorch Swish)
#pragma unroll
          for (int j = 0; j < U; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
        }
        if (st_tma && p.y) {
          tma_out(&a.tmY, n0, relu_packed, v);
        } else
        if (live && p.y) {
          uint8_t* yp = reinterpret_cast<uint8_t*>(p.y) + (m * p.Cout + n0) * 2;
#pragma unroll
          for (int j = 0; j < U; j += 8) {
            if (n0 + j + 8 <= p.Cout) {
              uint4 w;
              w.x = pack2(v[j], v[j + 1], bf16);
              w.y = pack2(v[j + 2], v[j + 3], bf16);
              w.z = pack2(v[j + 4], v[j + 5], bf16);
              w.w = pack2(v[j + 6], v[j + 7], bf16);
              if (relu_packed) { w.x = relu2(w.x, bf16); w.y = relu2(w.y, bf16); w.z = relu2(w.z, bf16); w.w = relu2(w.w, bf16); }
              *reinterpret_cast<uint4*>(yp + j * 2) = w;
            } else {
              for (int t = 0; t < 8; ++t)
                if (n0 + j + t < p.Cout) {
                  const uint32_t w = pack2(relu_packed ? fmaxf(v[j + t], 0.f) : v[j + t], 0.f, bf16);
                  *reinterpret_cast<uint16_t*>(yp + (j + t) * 2) = (uint16_t)(w & 0xFFFFu);
                }
            }
          }
        }
        if (live && p.y32) {
          float* yp = p.y32 + m * p.Cout + n0;
#pragma unroll
          for (int j = 0; j < U; j += 4) {
            if (n0 + j + 4 <= p.Cout) {
              *reinterpret_cast<float4*>(yp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              for (int t = 0; t < 4; ++t) if (n0 + j + t < p.Cout) yp[j + t] = v[j + t];
            }
          }
        }
        if constexpr (BN == 32 || BN == 64 || BN == 128) {
          if (do_ln) {
            // exact two-pass statistics on this warp's U columns (in registers); for BN >= 64 the two
            // column halves combine (mean, M2) through shared memory (Chan et al.)
            float s1 = 0.f;
#pragma unroll
            for (int j = 0; j < U; ++j) s1 += v[j];
            const float mh = s1 * (1.0f / U);
            float m2 = 0.f;
#pragma unroll
            for (int j = 0; j < U; ++j) { const float d = v[j] - mh; m2 += d * d; }
            float mean = mh, var = m2 * (1.0f / U);
            if constexpr (BN >= 64) {
              const int par = (acc_it / (uint32_t)n_teams) & 1;   // exchange slots are double-buffered per team
              sXch[team][par][hf][row] = make_float2(mh, m2);
              asm volatile("bar.sync %0, 64;" ::"r"(1 + team * 4 + q4) : "memory");
              const float2 o = sXch[team][par][hf ^ 1][row];
              mean = 0.5f * (mh + o.x);
              const float dlt = mh - o.x;
              var = (m2 + o.y + dlt * dlt * (float)(U / 2)) * (1.0f / (2 * U));
            }
            const float rstd = rsqrtf(var + p.ln_eps);
            if (st_tma) {
#pragma unroll
              for (int j = 0; j < U; ++j) v[j] = fmaf((v[j] - mean) * rstd, sGamma[(col + j) & 127], sBeta[(col + j) & 127]);
              tma_out(&a.tmYln, n0, false, v);
            } else if (live) {
              uint8_t* yp = reinterpret_cast<uint8_t*>(p.y_ln) + (m * p.Cout + n0) * 2;
#pragma unroll
              for (int j = 0; j < U; ++j) v[j] = fmaf((v[j] - mean) * rstd, sGamma[(col + j) & 127], sBeta[(col + j) & 127]);
#pragma unroll
              for (int j = 0; j < U; j += 8) {
                uint4 w;
                w.x = pack2(v[j], v[j + 1], bf16);
                w.y = pack2(v[j + 2], v[j + 3], bf16);
                w.z = pack2(v[j + 4], v[j + 5], bf16);
                w.w = pack2(v[j + 6], v[j + 7], bf16);
                *reinterpret_cast<uint4*>(yp + j * 2) = w;
              }
            }
          }
        }
      }
    }
    if (a.st_tma && lane == 0) bulk_wait0();      // every tensor store of this warp has completed
  } else if (warp == 16) {
    // ================================ loader ==============================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONTROL));
    if (lane == 0) {
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
      if (a.resident_b) {
        mbar_arrive_expect_tx(&b_res_full, (uint32_t)(KB * kBStage));
        for (int kb = 0; kb < KB; ++kb) bulk_g2s(sB + kb * kBStage, wbase + (int64_t)kb * kBStage, kBStage, &b_res_full);
      }
      pdl_wait();                         // weights are constants; the A operand is the previous kernel's output
      int bs = 0, ps = 0;                 // B / A ring positions and the parities to wait for
      uint32_t bph = 1, pph = 1;
      auto load_b = [&](const uint8_t* wb, int kb) {
        mbar_wait(&b_empty[bs], bph);
        mbar_arrive_expect_tx(&b_full[bs], kBStage);
        bulk_g2s(sB + bs * kBStage, wb + (int64_t)kb * kBStage, kBStage, &b_full[bs]);
        if (++bs == SB) { bs = 0; bph ^= 1u; }
      };
      if (a.use_tma) {
        // One thread feeds both rings: the A stage is a 3-D halo box (conv padding = TMA out-of-bounds zero
        // fill) or a 128-row token slab, written by the TMA engine directly in the SWIZZLE_128B image the
        // UMMA descriptors read; the weight stages follow in the MMA's consumption order.
        const int per_img = a.tiles_y * a.tiles_x;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
          const int m_tile = tile / a.n_tiles;
          const int n_tile = tile - m_tile * a.n_tiles;
          const uint8_t* wb = wbase + (int64_t)n_tile * KB * kBStage;
          if (a.mode == MODE_HALO) {
            const int n = m_tile / per_img;
            const int t = m_tile - n * per_img;
            const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
            for (int slab = 0; slab < a.nslab; ++slab) {
              mbar_wait(&a_empty[ps], pph);
              mbar_arrive_expect_tx(&a_full[ps], HALO_H * HALO_W * 128);
              tma_load_4d(sA + ps * A_HALO_BYTES, &a.tmap, &a_full[ps], slab * 64, tx * TILE_W - 1, ty * TILE_H - 1, n);
              if (++ps == SA) { ps = 0; pph ^= 1u; }
              if (!a.resident_b)
                for (int tap = 0; tap < 9; ++tap) load_b(wb, tap * a.nslab + slab);
            }
          } else if (a.mode == MODE_PATCH) {
            // 8-channel input (one pixel = 16 bytes = one core-matrix row): the raw input patch of the tile is
            // loaded once, unswizzled; the MMA descriptors walk it in place (see the issuer)
            const int n = m_tile / per_img;
            const int t = m_tile - n * per_img;
            const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
            mbar_wait(&a_empty[ps], pph);
            mbar_arrive_expect_tx(&a_full[ps], (uint32_t)(a.patch_w * a.patch_h * 16));
            // (pixels and channels are one merged dimension, so a patch row is ONE 176-byte request, not 11)
            tma_load_3d(sA + ps * A_GATHER_BYTES, &a.tmap, &a_full[ps], (tx * TILE_W * p.stride_w - p.pad_w) * 8,
                        ty * TILE_H * p.stride_h - p.pad_h, n);
            if (++ps == SA) { ps = 0; pph ^= 1u; }
          } else if (a.mode == MODE_TAP) {
            // general strided conv: one strided 16 x 8 pixel box per filter tap (and 64-channel slab); the
            // taps of a K block sit side by side in the stage (whole 128-byte rows, 64-byte rows or 16-byte
            // core-matrix columns, see a_layout).  Padding = out-of-bounds zero fill.
            const int n = m_tile / per_img;
            const int t = m_tile - n * per_img;
            const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
            const int x0 = tx * TILE_W * p.stride_w - p.pad_w, y0 = ty * TILE_H * p.stride_h - p.pad_h;
            const int ntaps = a.ntaps;
            const uint32_t tap_bytes = (uint32_t)BLOCK_M * a.tap_c * 2;
            int tap = 0, slab = 0, ky = 0, kx = 0;
            for (int kb = 0; kb < KB; ++kb) {
              mbar_wait(&a_empty[ps], pph);
              uint8_t* dst = sA + ps * A_GATHER_BYTES;
              const int nload = a.slabs_per_tap > 1 ? 1 : (ntaps - tap < a.taps_per_kb ? ntaps - tap : a.taps_per_kb);
              mbar_arrive_expect_tx(&a_full[ps], nload * tap_bytes);
              for (int j = 0; j < nload; ++j) {
                tma_load_4d(dst + j * tap_bytes, &a.tmap, &a_full[ps], slab * 64, x0 + kx, y0 + ky, n);
                if (++slab == a.slabs_per_tap) { slab = 0; ++tap; if (++kx == a.tap_kw) { kx = 0; ++ky; } }
              }
              if (++ps == SA) { ps = 0; pph ^= 1u; }
              if (!a.resident_b) load_b(wb, kb);
            }
          } else {
            for (int kb = 0; kb < KB; ++kb) {
              mbar_wait(&a_empty[ps], pph);
              mbar_arrive_expect_tx(&a_full[ps], A_GATHER_BYTES);
              tma_load_2d(sA + ps * A_GATHER_BYTES, &a.tmap, &a_full[ps], kb * BLOCK_K, m_tile * BLOCK_M);
              if (++ps == SA) { ps = 0; pph ^= 1u; }
              if (!a.resident_b) load_b(wb, kb);
            }
          }
        }
      } else if (!a.resident_b) {
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
          const int n_tile = tile % a.n_tiles;
          const uint8_t* wb = wbase + (int64_t)n_tile * KB * kBStage;
          int slab = 0, tap = 0;
          for (int j = 0; j < KB; ++j) {
            // consumption order: GATHER kb = j; HALO (slab-major, taps inner) kb = tap*nslab + slab
            int kb = j;
            if (a.mode == MODE_HALO) { kb = tap * a.nslab + slab; if (++tap == 9) { tap = 0; ++slab; } }
            load_b(wb, kb);
          }
        }
      }
    }
  } else if (warp == 17) {
    // ================================ MMA issuer ==========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONTROL));
    // The whole warp walks the pipeline (waits, stage counters and descriptor arithmetic stay on the
    // uniform datapath); one elected lane issues the tcgen05 instructions.
    {
      uint32_t acc_it = 0;
      int sa_ = 0, sb_ = 0;            // current A / B stage and the parity to wait for
      uint32_t aph = 0, bph = 0;
      // descriptor halves: `hi` is constant, `lo` = start address >> 4 | LBO field
      const uint64_t tA_g = make_smem_desc(0, 16, 1024, LAYOUT_SW128);
      const uint64_t tA_h = make_smem_desc(0, 16, HALO_W * 128, LAYOUT_SW128);
      const uint32_t lo0 = (uint32_t)tA_g;                 // LBO bits only (same for all three)
      const uint32_t hiA_g = (uint32_t)(tA_g >> 32), hiA_h = (uint32_t)(tA_h >> 32), hiB = hiA_g;
      const uint64_t tA_64 = make_smem_desc(0, 16, 512, LAYOUT_SW64);
      const uint64_t tA_n = make_smem_desc(0, 2048, 128, LAYOUT_NONE);
      const uint32_t hiA_64 = (uint32_t)(tA_64 >> 32), hiA_n = (uint32_t)(tA_n >> 32), lo_n = (uint32_t)tA_n;
      const uint32_t idesc = a.idesc;
      const bool do_mma = !COBEVT_DBG(a, 4);
      if (a.resident_b) mbar_wait(&b_res_full, 0);
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++acc_it) {
        const int buf = acc_it & (NACC - 1);
        mbar_wait(&acc_empty[buf], ((acc_it / NACC) & 1) ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * BN;
        uint32_t accum = 0;
        if (a.mode == MODE_PATCH) {
          // A row m = output pixel (y, x) of the 16 x 8 tile; its K slice for filter row ky is the KW*8 contiguous
          // elements starting at input pixel (y*sh + ky, x).  In the unswizzled patch, pixels are 16 bytes apart, so
          // the 8 x 16 B core matrix of K chunk kx is 8 consecutive pixels starting at x + kx: overlapping core
          // matrices, LBO = 16 B between K chunks, SBO = sh patch rows between 8-row groups.  No im2col copy exists.
          mbar_wait(&a_full[sa_], aph);
          tc_fence_after_sync();
          const uint32_t pitch = (uint32_t)a.patch_w * 16u;
          const uint64_t tA_p = make_smem_desc(0, 16, (uint32_t)p.stride_h * pitch, LAYOUT_NONE);
          const uint32_t lo_p = (uint32_t)tA_p, hi_p = (uint32_t)(tA_p >> 32);
          const uint32_t a_addr = smem_u32(sA + sa_ * A_GATHER_BYTES);
          const uint32_t b_base = lo0 | (smem_u32(sB) >> 4);
          if (elect_one()) {
            if (do_mma) {
              if (p.KH == 7 && p.KW == 4) {           // the stem: fully unrolled, descriptor deltas are immediates
                const uint32_t a0 = lo_p | (a_addr >> 4), pstep = pitch >> 4;
#pragma unroll
                for (int j = 0; j < 14; ++j)
                  umma_lohi(d_tmem, a0 + (j >> 1) * pstep + (j & 1) * 2, hi_p,
                            b_base + (j >> 2) * (kBStage >> 4) + (j & 3) * 2, hiB, idesc, j ? 1u : 0u);
              } else {
                int j = 0;
                for (int ky = 0; ky < p.KH; ++ky)
                  for (int kp = 0; kp < p.KW / 2; ++kp, ++j)
                    umma_lohi(d_tmem, lo_p | ((a_addr + ky * pitch + kp * 32) >> 4), hi_p,
                              b_base + (j >> 2) * (kBStage >> 4) + (j & 3) * 2, hiB, idesc, j ? 1u : 0u);
              }
            }
            umma_commit(&a_empty[sa_]);
          }
          if (++sa_ == SA) { sa_ = 0; aph ^= 1u; }
        } else if (a.mode != MODE_HALO) {
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&a_full[sa_], aph);
            if (!a.use_tma) fence_proxy_async_smem();     // cp.async (generic proxy) bytes -> visible to the UMMA (async proxy)
            uint32_t sb_addr;
            if (a.resident_b) {
              sb_addr = smem_u32(sB + kb * kBStage);
            } else {
              mbar_wait(&b_full[sb_], bph);
              sb_addr = smem_u32(sB + sb_ * kBStage);
            }
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(sA + sa_ * A_GATHER_BYTES) >> 4;
            const uint32_t a_lo = lo0 | a_addr;
            const uint32_t b_lo = lo0 | (sb_addr >> 4);
            if (elect_one()) {
              if (do_mma) {
                if (a.a_layout == A_SW128) {
#pragma unroll
                  for (int ks = 0; ks < BLOCK_K / 16; ++ks) umma_lohi(d_tmem, a_lo + ks * 2, hiA_g, b_lo + ks * 2, hiB, idesc, (accum | ks) ? 1u : 0u);
                } else {
                  // TAP stages of narrow inputs: 32 channels -> two SWIZZLE_64B tiles (one per tap, 8 KB);
                  // 8 channels -> eight 2 KB columns of 8 x 16 B core matrices (no swizzle, LBO = next tap)
                  const int rem = a.K - kb * BLOCK_K;
                  const int nks = rem >= BLOCK_K ? BLOCK_K / 16 : (rem + 15) / 16;
                  for (int ks = 0; ks < nks; ++ks) {
                    const uint32_t alo = a.a_layout == A_SW64 ? (lo0 | (a_addr + (ks >> 1) * 512 + (ks & 1) * 2))
                                                               : (lo_n | (a_addr + ks * 256));
                    umma_lohi(d_tmem, alo, a.a_layout == A_SW64 ? hiA_64 : hiA_n, b_lo + ks * 2, hiB, idesc, (accum | ks) ? 1u : 0u);
                  }
                }
              }
              umma_commit(&a_empty[sa_]);
              if (!a.resident_b) umma_commit(&b_empty[sb_]);
            }
            accum = 1;
            if (++sa_ == SA) { sa_ = 0; aph ^= 1u; }
            if (!a.resident_b) { if (++sb_ == SB) { sb_ = 0; bph ^= 1u; } }
          }
        } else {
          const uint32_t tap_stride = (uint32_t)(a.nslab * kBStage) >> 4;     // resident B: next tap, same slab
          for (int slab = 0; slab < a.nslab; ++slab) {
            mbar_wait(&a_full[sa_], aph);
            if (!a.use_tma) fence_proxy_async_smem();
            tc_fence_after_sync();
            const uint32_t a_lo = lo0 | (smem_u32(sA + sa_ * A_HALO_BYTES) >> 4);
            uint32_t b_res_lo = lo0 | (smem_u32(sB + slab * kBStage) >> 4);
            constexpr int kTapRows[9] = {0, 1, 2, HALO_W, HALO_W + 1, HALO_W + 2, 2 * HALO_W, 2 * HALO_W + 1, 2 * HALO_W + 2};
            if (a.resident_b) {
              if (elect_one()) {
                if (do_mma) {
#pragma unroll
                  for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                    for (int ks = 0; ks < BLOCK_K / 16; ++ks)
                      umma_lohi(d_tmem, a_lo + kTapRows[tap] * 8 + ks * 2, hiA_h, b_res_lo + tap * tap_stride + ks * 2, hiB, idesc,
                                (accum | tap | ks) ? 1u : 0u);
                  }
                }
                umma_commit(&a_empty[sa_]);
              }
            } else {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                mbar_wait(&b_full[sb_], bph);
                tc_fence_after_sync();
                const uint32_t b_lo = lo0 | (smem_u32(sB + sb_ * kBStage) >> 4);
                if (elect_one()) {
                  if (do_mma) {
#pragma unroll
                    for (int ks = 0; ks < BLOCK_K / 16; ++ks)
                      umma_lohi(d_tmem, a_lo + kTapRows[tap] * 8 + ks * 2, hiA_h, b_lo + ks * 2, hiB, idesc, (accum | tap | ks) ? 1u : 0u);
                  }
                  umma_commit(&b_empty[sb_]);
                  if (tap == 8) umma_commit(&a_empty[sa_]);
                }
                if (++sb_ == SB) { sb_ = 0; bph ^= 1u; }
              }
            }
            accum = 1;
            if (++sa_ == SA) { sa_ = 0; aph ^= 1u; }
          }
        }
        if (elect_one()) umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONTROL));   // idle half of the control warpgroup
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_WORKERS));   // warps 12-15 without a second team
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}


// tensor map of the A operand for the TMA-fed modes; returns false if the layout cannot be described
bool make_a_tmap(KArgs& a) {
  cobevt::EncodeTiledFn enc = cobevt::get_encode_tiled();
  if (!enc) return false;
  const CobevtConvParams& p = a.p;
  const CUtensorMapDataType dt = p.dtype == COBEVT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r;
  if (a.mode == MODE_HALO) {
    const cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)p.N};
    const cuuint64_t strides[3] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Win * p.Cin * 2, (cuuint64_t)p.Hin * p.Win * p.Cin * 2};
    const cuuint32_t box[4] = {64, HALO_W, HALO_H, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(&a.tmap, dt, 4, const_cast<void*>(p.x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (a.mode == MODE_PATCH) {
    // {W*8, H, N}: the 8 channels and the pixels of an image row are one contiguous dimension
    const cuuint64_t dims[3] = {(cuuint64_t)p.Win * 8, (cuuint64_t)p.Hin, (cuuint64_t)p.N};
    const cuuint64_t strides[2] = {(cuuint64_t)p.Win * 16, (cuuint64_t)p.Hin * p.Win * 16};
    const cuuint32_t box[3] = {(cuuint32_t)a.patch_w * 8, (cuuint32_t)a.patch_h, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    r = enc(&a.tmap, dt, 3, const_cast<void*>(p.x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (a.mode == MODE_TAP) {
    // (fused filter rows: the KW taps of one row are KW*Cin contiguous elements starting at the pixel, i.e. a
    //  tensor of overlapping windows -- pixel stride Cin elements, window length tap_c -- is loaded instead)
    const bool fused_row = a.tap_kw == 1 && p.KW > 1;
    const cuuint64_t dims[4] = {(cuuint64_t)(fused_row ? a.tap_c : p.Cin), (cuuint64_t)(fused_row ? p.Win - p.KW + 1 : p.Win), (cuuint64_t)p.Hin, (cuuint64_t)p.N};
    const cuuint64_t strides[3] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Win * p.Cin * 2, (cuuint64_t)p.Hin * p.Win * p.Cin * 2};
    const cuuint32_t box[4] = {(cuuint32_t)a.tap_c, (cuuint32_t)((TILE_W - 1) * p.stride_w + 1), (cuuint32_t)((TILE_H - 1) * p.stride_h + 1), 1};
    const cuuint32_t es[4] = {1, (cuuint32_t)p.stride_w, (cuuint32_t)p.stride_h, 1};
    const CUtensorMapSwizzle sw = a.a_layout == A_SW128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : a.a_layout == A_SW64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    r = enc(&a.tmap, dt, 4, const_cast<void*>(p.x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)a.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.Cin * 2};
    const cuuint32_t box[2] = {BLOCK_K, BLOCK_M};
    const cuuint32_t es[2] = {1, 1};
    r = enc(&a.tmap, dt, 2, const_cast<void*>(p.x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS;
}


// output tensor maps for the TMA-store epilogue; U = columns per epilogue unit of this BN
bool make_out_tmaps(KArgs& a, int U, bool single_unit) {
  cobevt::EncodeTiledFn enc = cobevt::get_encode_tiled();
  if (!enc) return false;
  const CobevtConvParams& p = a.p;
  const CUtensorMapDataType dt = p.dtype == COBEVT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = U == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (U == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  auto one = [&](CUtensorMap* tm, void* base) -> bool {
    if (!base) return true;
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    CUresult r;
    if (a.mode == MODE_LINEAR || a.mode == MODE_GATHER) {
      const cuuint64_t dims[2] = {(cuuint64_t)p.Cout, (cuuint64_t)a.M};
      const cuuint64_t strides[1] = {(cuuint64_t)p.Cout * 2};
      const cuuint32_t box[2] = {(cuuint32_t)U, 32};
      const cuuint32_t es[2] = {1, 1};
      r = enc(tm, dt, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)p.OW, (cuuint64_t)p.OH, (cuuint64_t)p.N};
      const cuuint64_t strides[3] = {(cuuint64_t)p.Cout * 2, (cuuint64_t)p.OW * p.Cout * 2, (cuuint64_t)p.OH * p.OW * p.Cout * 2};
      const cuuint32_t box[4] = {(cuuint32_t)U, TILE_W, 4, 1};
      const cuuint32_t es[4] = {1, 1, 1, 1};
      r = enc(tm, dt, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS;
  };
  if (!(one(&a.tmY, p.y) && one(&a.tmYln, p.y_ln))) return false;
  // residual through the same tiles: plain [M][Cout] 16-bit residual of a single-unit (BN = 128) token GEMM
  static int no_rt = -1;
  if (no_rt < 0) { const char* e = getenv("COBEVT_CONV_NORESTMA"); no_rt = e ? atoi(e) : 0; }
  a.res_tma = 0;
  if (!no_rt && single_unit && p.res && p.res_rows == 0 && U == 64 && (a.mode == MODE_LINEAR || a.mode == MODE_GATHER) &&
      one(&a.tmRes, const_cast<void*>(p.res)))
    a.res_tma = 1;
  return true;
}

template <int BN>
int launch(KArgs& a, cudaStream_t st) {
  constexpr int kBStage = BN * BLOCK_K * 2;
  constexpr int CWh = BN >= 64 ? BN / 2 : BN;            // mirrors the epilogue's CW / U
  constexpr int Uh = CWh > 64 ? 64 : CWh;
  const int a_bytes = a.mode == MODE_HALO ? A_HALO_BYTES : A_GATHER_BYTES;
  static int no_tst = -1;
  if (no_tst < 0) { const char* e = getenv("COBEVT_CONV_NOTMASTORE"); no_tst = e ? atoi(e) : 0; }
  // 16-bit outputs through TMA tensor stores when the staging tiles (16 warps x 4 KB) fit next to the rings
  // (token GEMMs only: they are store-bound; for the 3x3 convs the staging area costs ring depth and the narrow
  //  per-pixel boxes are no better than direct stores -- measured, profiles/r01i_launch_detail.log)
  const bool want_st = !no_tst && a.use_tma && a.mode == MODE_LINEAR && (a.p.y || a.p.y_ln) && a.p.Cout % 8 == 0 && Uh == 64;
  const int sa_full = a.mode == MODE_HALO ? SA_HALO : SA_GATHER;
  const int cand_st[3] = {1, 1, 0};
  const int cand_sa[3] = {sa_full, a.mode == MODE_HALO ? 3 : 4, sa_full};
  int64_t smem = 0;
  bool fit = false;
  for (int c = want_st ? 0 : 2; c < 3 && !fit; ++c) {
    a.st_tma = cand_st[c];
    a.sa = cand_sa[c];
    const int budget = kMaxDynSmem - 1024 - (a.st_tma ? 16 * STG_WARP_BYTES : 0);
    const int64_t b_all = (int64_t)a.KB * kBStage;
    a.resident_b = (a.n_tiles == 1 && b_all <= budget - a.sa * a_bytes && b_all < (1 << 20)) ? 1 : 0;
    if (a.resident_b) {
      a.sb = 1;
      smem = 1024 + (int64_t)a.sa * a_bytes + b_all;
      fit = true;
    } else {
      int sb = (budget - a.sa * a_bytes) / kBStage;
      if (sb > MAX_STAGES) sb = MAX_STAGES;
      if (sb < 3 && a.st_tma) continue;                      // the weight ring matters more than the store path
      if (sb < 2) { a.sa = 2; sb = (budget - a.sa * a_bytes) / kBStage; }
      if (sb < 2) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: no smem for the B ring (BN=%d)", BN);
      a.sb = sb;
      smem = 1024 + (int64_t)a.sa * a_bytes + (int64_t)sb * kBStage;
      fit = true;
    }
  }
  a.res_tma = 0;
  if (a.st_tma && !make_out_tmaps(a, Uh, BN == 128)) {      // no tensor maps: direct stores (the staging area stays unused)
    a.st_tma = 0;
    a.res_tma = 0;
  }
  if (a.st_tma) smem += 16 * STG_WARP_BYTES;
  if (a.mode == MODE_PATCH && !a.resident_b) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: PATCH mode needs the resident weight image");
  static bool attr_set[cobevt::kMaxDevices][2] = {};         // per device and instance (and per BN: one static per instantiation)
  const int dev = cobevt::current_device();
  void (*kern)(const KArgs) = conv_gemm_kernel<BN, false>;
  if constexpr (BN == 128) {
    if (a.res_tma) kern = conv_gemm_kernel<BN, true>;
  }
  if (!attr_set[dev][a.res_tma]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (e != cudaSuccess) return cobevt::fail(COBEVT_ERR_CUDA, "conv: set smem attr: %s", cudaGetErrorString(e));
    attr_set[dev][a.res_tma] = true;
  }
  const int num_sms = cobevt::num_sms();
  const int grid = a.total_tiles < num_sms ? a.total_tiles : num_sms;
  cobevt::launch_k(kern, grid, NUM_THREADS, (size_t)smem, st, a);
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
  if (p->y_ln && !(p->Cout == p->block_n && (p->Cout == 32 || p->Cout == 64 || p->Cout == 128)))
    return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: y_ln needs Cout == block_n in {32, 64, 128}");
  if ((p->ln_gamma == nullptr) != (p->ln_beta == nullptr)) return cobevt::fail(COBEVT_ERR_SHAPE, "conv: ln_gamma/ln_beta must come together");
  if (p->Cout > 512) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: Cout %d > 512", p->Cout);
  if (p->Cout % 8 != 0 && (p->res || p->y_ln)) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "conv: Cout %% 8 with res/ln");
  const int64_t M64 = (int64_t)p->N * p->OH * p->OW;
  if (M64 > (1ll << 30)) return cobevt::fail(COBEVT_ERR_SHAPE, "conv: M too large");
  if ((int64_t)p->N * p->Hin * p->Win * p->Cin >= (1ll << 31)) return cobevt::fail(COBEVT_ERR_SHAPE, "conv: input too large for 32-bit offsets");
  KArgs a;
  memset(&a, 0, sizeof(a));
  a.p = *p;
  a.M = (int32_t)M64;
  a.K = p->KH * p->KW * p->Cin;
  a.KB = cobevt::ceil_div(a.K, BLOCK_K);
  a.n_tiles = cobevt::ceil_div(p->Cout, p->block_n);
  a.idesc = make_idesc_f16(128, p->block_n, p->dtype == COBEVT_BF16, false, false);
  const bool halo = p->KH == 3 && p->KW == 3 && p->stride_h == 1 && p->stride_w == 1 && p->pad_h == 1 && p->pad_w == 1 &&
                    p->Cin % 64 == 0 && !p->in_upsample && p->OH == p->Hin && p->OW == p->Win;
  if (halo) {
    a.mode = MODE_HALO;
    a.tiles_y = cobevt::ceil_div(p->OH, TILE_H);
    a.tiles_x = cobevt::ceil_div(p->OW, TILE_W);
    a.nslab = p->Cin / 64;
    a.m_tiles = p->N * a.tiles_y * a.tiles_x;
  } else {
    const bool linear = p->KH == 1 && p->KW == 1 && p->stride_h == 1 && p->stride_w == 1 && p->pad_h == 0 &&
                        p->pad_w == 0 && !p->in_upsample;
    a.mode = linear ? MODE_LINEAR : MODE_GATHER;
    a.m_tiles = cobevt::ceil_div(a.M, BLOCK_M);
    a.cin_log2 = -1;
    if ((p->Cin & (p->Cin - 1)) == 0 && p->KW <= 16 && p->KH * p->KW < 4096) {
      int l = 0;
      while ((1 << l) < p->Cin) ++l;
      a.cin_log2 = l;
      a.kw_magic = (65536 + p->KW - 1) / p->KW;
    }
  }
  {
    static int no_tma = -1, no_tap = -1;
    if (no_tma < 0) { const char* e = getenv("COBEVT_CONV_NOTMA"); no_tma = e ? atoi(e) : 0; }
    if (no_tap < 0) { const char* e = getenv("COBEVT_CONV_NOTAP"); no_tap = e ? atoi(e) : 0; }   // 1 no TAP, 2 no PATCH, 4 no fused rows, 8 only Cin>=64 TAP
    a.use_tma = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(p->x) & 15) == 0;
    if (!no_tma && aligned && (a.mode == MODE_HALO || a.mode == MODE_LINEAR)) {
      a.use_tma = make_a_tmap(a) ? 1 : 0;
    } else if (!no_tma && aligned && a.mode == MODE_GATHER && !p->in_upsample &&
               (p->Cin == 8 || p->Cin == 32 || p->Cin % 64 == 0) && p->stride_h <= 8 && p->stride_w <= 8 && !(no_tap & 1) &&
               !((no_tap & 8) && p->Cin < 64)) {
      // strided / narrow convolutions: per-tap strided tensor loads over 16 x 8 output-pixel tiles
      KArgs t = a;
      t.mode = MODE_TAP;
      t.tiles_y = cobevt::ceil_div(p->OH, TILE_H);
      t.tiles_x = cobevt::ceil_div(p->OW, TILE_W);
      t.m_tiles = p->N * t.tiles_y * t.tiles_x;
      t.tap_c = p->Cin < 64 ? p->Cin : 64;
      t.taps_per_kb = p->Cin < 64 ? 64 / p->Cin : 1;
      t.slabs_per_tap = p->Cin < 64 ? 1 : p->Cin / 64;
      t.a_layout = p->Cin == 8 ? A_NOSWZ : (p->Cin == 32 ? A_SW64 : A_SW128);
      t.ntaps = p->KH * p->KW;
      t.tap_kw = p->KW;
      if (!(no_tap & 4) && p->stride_w == 1 && p->pad_w == 0 && p->KW > 1 && p->Cin < 64 && (p->Cin * p->KW == 32 || p->Cin * p->KW == 64)) {
        t.tap_c = p->Cin * p->KW;           // one load per filter row
        t.taps_per_kb = 64 / t.tap_c;
        t.a_layout = t.tap_c == 32 ? A_SW64 : A_SW128;
        t.ntaps = p->KH;
        t.tap_kw = 1;
      }
      // 8-channel inputs with unit stride along W: the raw input patch of a tile is the A operand (PATCH mode),
      // provided the whole packed weight can stay resident
      const int pw = TILE_W + p->KW - 1, ph = (TILE_H - 1) * p->stride_h + p->KH;
      if (!(no_tap & 2) && p->Cin == 8 && p->stride_w == 1 && p->KW % 2 == 0 && pw * 8 <= 256 && ph <= 256 && pw * ph * 16 <= A_GATHER_BYTES &&
          a.n_tiles == 1 && (int64_t)a.KB * p->block_n * BLOCK_K * 2 <= 96 * 1024) {
        t.mode = MODE_PATCH;
        t.patch_w = pw;
        t.patch_h = ph;
        t.a_layout = A_SW128;
      }
      if (make_a_tmap(t)) { a = t; a.use_tma = 1; }
    }
  }
  a.total_tiles = a.m_tiles * a.n_tiles;
#ifdef COBEVT_DEBUG_KNOBS
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("COBEVT_CONV_DEBUG"); dbg = e ? atoi(e) : 0; } a.dbg = dbg; }
#endif
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
