// Windowed / dilated-grid multi-head attention (head_dim = 32) for sm_100a.
//
// Three kernels share the operand layouts (Q / K: K-major SWIZZLE_64B tiles, P: 16-bit K-major A operand, V: MN-major
// SWIZZLE_64B, i.e. the same row-per-key image as K -- no transpose) and the token addressing: window
// ('(x w1)(y w2)') and grid ('(w1 x)(w2 y)') partitions are address arithmetic / TMA boxes on the NHWC token maps,
// so none of the reference's rearrange copies exist.  Zero-padded key positions use a constant token
// (Linear(LayerNorm(0)) of the reference, survey quirk Q2).
//
//   window_attn_v3_kernel   windows of <= 704 keys (all FAX cross attentions of the north_star config and the
//                           FuseBEVT window / grid attention): K/V resident, FOUR softmax warpgroups, O accumulated
//                           in TMEM with lazy rescale, K/V by TMA; longer multi-view windows run as key segments
//                           + cobevt_attn_merge.
//   window_attn_v2_kernel   705 .. 1024 keys: K/V resident, two softmax warpgroups, register-level two-chunk pipeline.
//   window_attn_kernel      anything longer: K/V streamed in 128-key chunks, one 128-thread CTA per query tile.
//
// Per chunk:  S = Q K^T (tcgen05.mma) -> TMEM -> registers (tcgen05.ld); online softmax in fp32 with exp2 on the
// MUFU (scale * log2 e folded; by default two of the eight key groups of a chunk use the FMA-pipe polynomial instead);
// P -> shared memory; O (+)= P V (tcgen05.mma).
//
// window_attn_v3_kernel<BF16, POLY, QLIN, BIAS> instances: unbiased launches run <., 0x44, false, false>; the biased ones
// (FuseBEVT, global self-attention) run <., 0, false, true>, the only instance with the bias code: it reads the
// key-group-major bias copy (bias_layout 1) and starts its key-chunk walk at chunk (window % chunks) so that the CTAs of
// a head do not hammer the same L2 lines.  QLIN (one carried query offset) is kept behind COBEVT_ATTN_QLIN=1.
//
// Reference: CrossWinAttention.forward fax_modules.py:198-248, partitions
// fax_modules.py:399-433; Attention.forward fax_modules.py:132-176;
// swap_fusion_modules.Attention.forward :87-128.
#include <cstdlib>
#include <type_traits>
#include "sm100.cuh"
#include "common.h"

using namespace sm100;

namespace {

constexpr int TILE_Q = 128;
constexpr int CHUNK = 128;   // keys per chunk
constexpr int DH = 32;       // head dim
constexpr int SQ_BYTES = TILE_Q * DH * 2;        // 8 KB
constexpr int SK_BYTES = CHUNK * DH * 2;         // 8 KB
constexpr int SP_BYTES = TILE_Q * CHUNK * 2;     // 32 KB
constexpr int SMEM_BYTES = 1024 + SQ_BYTES + 2 * (2 * SK_BYTES) + SP_BYTES;
constexpr int TMEM_COLS = 256;  // S: [0,128)  O: [128,160)

struct AArgs {
  CobevtAttnParams p;
  int32_t n_win, q_tiles, Lq, Lk, n_chunks, q_passes, C, q_dbuf;
  int32_t dbg;   // COBEVT_ATTN_DEBUG (timing experiments only; results are wrong): 1 no MUFU, 2 no P stores, 4 no proxy fence, 8 second warpgroup idle
  int32_t q_nx, q_ny, k_nx, k_ny;  // windows along h / w
  uint32_t idesc_qk, idesc_pv;
  int32_t kv_vt, kv_segs;             // views of the K/V tensors (batch stride in views); key segments per launch
  int32_t bias_ld;                    // keys per bias row = keys of the WHOLE window (all segments); == Lk without segments
  int32_t kv_tma;                     // resident K / V windows are fetched by two 4-D TMA tensor loads
  alignas(64) CUtensorMap tmK, tmV;   // {C, Wv, Hv, B*views}, box {32, w2 (span), w1 (span), views}, SWIZZLE_64B
};

struct Tok {
  int n, py, px;
  bool in_range;  // inside the valid (stored) extent
};

// token t (0 .. views*w1*w2) of window `win` -> plane / pixel
__device__ __forceinline__ Tok token_of(const CobevtTokenMap& m, int nx, int ny, int win, int t) {
  Tok r;
  const int ww = m.w1 * m.w2;
  r.n = t / ww;
  const int tt = t - r.n * ww;
  const int i = tt / m.w2, j = tt - i * m.w2;
  const int X = win / ny, Y = win - X * ny;
  if (m.grid) {
    r.py = i * nx + X;
    r.px = j * ny + Y;
  } else {
    r.py = X * m.w1 + i;
    r.px = Y * m.w2 + j;
  }
  r.in_range = r.py < m.Hv && r.px < m.Wv;
  return r;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes for a pair of logits (FA-4 style software exponential): the MUFU does 16 ex2 per
// clock per SM, which bounds this head size (128 tensor FLOPs per exponential), so a compile-time share of every
// chunk's exponentials is evaluated here instead.  x = i + f (i = floor, by an add with round-to-minus-infinity
// against 1.5 * 2^23, which also leaves i in the low mantissa bits); 2^f by the degree-3 minimax polynomial
// (max relative error 8.8e-5, a fifth of the fp16 rounding P receives anyway); 2^i by adding i to the exponent field.
// Inputs are clamped to >= -126 (masked keys arrive as -inf): the result is then a normal number <= 2^-126 * 2,
// which the 16-bit P rounds to zero exactly like the MUFU path.
__device__ __forceinline__ void poly_exp2_pair(float& x0, float& x1) {
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  float t0, t1, r0, r1, f0, f1, p0, p1;
  asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\t"
      "add.rm.ftz.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(t0), "=f"(t1) : "f"(x0), "f"(x1), "f"(12582912.0f));
  fadd2(r0, r1, t0, t1, -12582912.0f, -12582912.0f);                 // floor(x) as a float (exact)
  ffma2(f0, f1, r0, r1, -1.0f, -1.0f, x0, x1);                        // f = x - floor(x) in [0, 1)
  ffma2(p0, p1, f0, f1, 0.077119089663028717f, 0.077119089663028717f, 0.227564394474029541f, 0.227564394474029541f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.695146143436431885f, 0.695146143436431885f);
  ffma2(p0, p1, p0, p1, f0, f1, 1.0f, 1.0f);
  x0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(128, 2) window_attn_kernel(const __grid_constant__ AArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + SQ_BYTES;              // buffer b: K at sKV + b*2*SK_BYTES, V right after
  uint8_t* sP = sKV + 4 * SK_BYTES;
  __shared__ uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_base_s;

  const CobevtAttnParams& p = a.p;
  const int tid = threadIdx.x, warp = tid >> 5;
  int item = blockIdx.x;
  const int q_tile = item % a.q_tiles; item /= a.q_tiles;
  const int head = item % p.heads;     item /= p.heads;
  const int win = item % a.n_win;
  const int b = item / a.n_win;
  const int bf16 = p.dtype;

  if (tid == 0) {
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  uint32_t ph_s = 0, ph_o = 0;

  const uint8_t* qb = reinterpret_cast<const uint8_t*>(p.q);
  const uint8_t* kb_ = reinterpret_cast<const uint8_t*>(p.k);
  const uint8_t* vb_ = reinterpret_cast<const uint8_t*>(p.v);
  const float sl2 = p.scale * 1.4426950408889634f;

  // my query row
  const int qrow = q_tile * TILE_Q + tid;      // index inside the window's row list
  const int rows_per_pass = p.q_mean ? p.qmap.w1 * p.qmap.w2 : a.Lq;
  const bool row_ok = qrow < rows_per_pass;

  float acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.f;

  // issue the cp.async loads of KV chunk `c` into buffer `buf`
  auto load_kv = [&](int c, int buf) {
    const int key = c * CHUNK + tid;
    uint8_t* dK = sKV + buf * 2 * SK_BYTES;
    uint8_t* dV = dK + SK_BYTES;
    const uint8_t* srcK = kb_;
    const uint8_t* srcV = vb_;
    uint32_t nb = 0;
    if (key < a.Lk) {
      const Tok t = token_of(p.kvmap, a.k_nx, a.k_ny, win, key);
      if (t.in_range) {
        const int64_t off = ((((int64_t)b * p.kvmap.views + t.n) * p.kvmap.Hv + t.py) * p.kvmap.Wv + t.px) * p.kv_stride + head * DH;
        srcK = kb_ + off * 2;
        srcV = vb_ + off * 2;
        nb = 16;
      } else if (p.k_pad) {
        srcK = reinterpret_cast<const uint8_t*>(p.k_pad) + head * DH * 2;
        srcV = reinterpret_cast<const uint8_t*>(p.v_pad) + head * DH * 2;
        nb = 16;
      }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      cp_async16(dK + sw64_off(tid, ch), srcK + (nb ? ch * 16 : 0), nb);
      cp_async16(dV + sw64_off(tid, ch), srcV + (nb ? ch * 16 : 0), nb);
    }
    cp_async_commit();
  };

  for (int pass = 0; pass < a.q_passes; ++pass) {
    // ---- Q tile ------------------------------------------------------------
    {
      const uint8_t* src = qb;
      uint32_t nb = 0;
      if (row_ok) {
        const int t_idx = p.q_mean ? pass * rows_per_pass + qrow : qrow;
        const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, t_idx);
        const int64_t off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.q_stride + head * DH;
        src = qb + off * 2;
        nb = 16;
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) cp_async16(sQ + sw64_off(tid, ch), src + (nb ? ch * 16 : 0), nb);
      cp_async_commit();
    }
    load_kv(0, 0);

    float m_run = -INFINITY, l_run = 0.f;
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;

    for (int c = 0; c < a.n_chunks; ++c) {
      const int buf = c & 1;
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncthreads();
      // ---- S = Q K^T ----------------------------------------------------------
      if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(sQ);
        const uint32_t sb = smem_u32(sKV + buf * 2 * SK_BYTES);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          const uint64_t da = make_smem_desc(sa + ks * 32, 16, 512, LAYOUT_SW64);
          const uint64_t db = make_smem_desc(sb + ks * 32, 16, 512, LAYOUT_SW64);
          umma_f16(tmem_base, da, db, a.idesc_qk, ks ? 1u : 0u);
        }
        umma_commit(&bar_s);
      }
      // prefetch next chunk into the other buffer (its previous user, chunk c-1, has retired)
      if (c + 1 < a.n_chunks) load_kv(c + 1, buf ^ 1);

      mbar_wait(&bar_s, ph_s);
      ph_s ^= 1;
      tc_fence_after_sync();

      // ---- softmax on this row (two sweeps over the TMEM row: max, then exp) ------
      const int key0 = c * CHUNK;
      auto logits = [&](int g, float (&x)[32]) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + g * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]) * sl2;
        const int kg = key0 + g * 32;
        if (p.bias && row_ok) {
          const float* bp = p.bias + ((int64_t)head * a.Lq + qrow) * a.Lk + kg;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (kg + j < a.Lk) x[j] = fmaf(__ldg(bp + j), 1.4426950408889634f, x[j]);
        }
        if (p.key_mask) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (kg + j < a.Lk) {
              const Tok t = token_of(p.kvmap, a.k_nx, a.k_ny, win, kg + j);
              const uint8_t mk = p.key_mask[(((int64_t)b * p.kvmap.Hv + t.py) * p.kvmap.Wv + t.px) * p.kvmap.views + t.n];
              if (mk == 0) x[j] = -INFINITY;
            }
          }
        }
        if (kg + 32 > a.Lk) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (kg + j >= a.Lk) x[j] = -INFINITY;
        }
      };
      float mx = m_run;
#pragma unroll 1
      for (int g = 0; g < CHUNK / 32; ++g) {
        float x[32];
        logits(g, x);
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, x[j]);
      }
      const float m_safe = (mx == -INFINITY) ? 0.f : mx;
      const float alpha = fast_exp2(m_run - m_safe);
      float sum = 0.f;
#pragma unroll 1
      for (int g = 0; g < CHUNK / 32; ++g) {
        float x[32];
        logits(g, x);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x[j] = fast_exp2(x[j] - m_safe);
          sum += x[j];
        }
        // P -> smem (A operand, two 64-key SW128 halves); this group = 4 x 16-byte chunks
        uint8_t* dst = sP + (g >> 1) * (TILE_Q * 128);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 u;
          u.x = pack_h2(x[cc * 8 + 0], x[cc * 8 + 1], bf16);
          u.y = pack_h2(x[cc * 8 + 2], x[cc * 8 + 3], bf16);
          u.z = pack_h2(x[cc * 8 + 4], x[cc * 8 + 5], bf16);
          u.w = pack_h2(x[cc * 8 + 6], x[cc * 8 + 7], bf16);
          *reinterpret_cast<uint4*>(dst + sw128_off(tid, (g & 1) * 4 + cc)) = u;
        }
      }
      l_run = l_run * alpha + sum;
      m_run = mx;
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      // ---- O_c = P V -------------------------------------------------------------
      if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(sP);
        const uint32_t sv = smem_u32(sKV + buf * 2 * SK_BYTES + SK_BYTES);
#pragma unroll
        for (int ks = 0; ks < CHUNK / 16; ++ks) {
          const uint64_t da = make_smem_desc(sa + (ks >> 2) * (TILE_Q * 128) + (ks & 3) * 32, 16, 1024, LAYOUT_SW128);
          const uint64_t db = make_smem_desc(sv + ks * 16 * 64, 512, 512, LAYOUT_SW64);
          umma_f16(tmem_base + CHUNK, da, db, a.idesc_pv, ks ? 1u : 0u);
        }
        umma_commit(&bar_o);
      }
      mbar_wait(&bar_o, ph_o);
      ph_o ^= 1;
      tc_fence_after_sync();
      {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + CHUNK, r);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] = fmaf(o[d], alpha, __uint_as_float(r[d]));
      }
      tc_fence_before_sync();
    }
    const float inv = (l_run > 0.f) ? 1.0f / (l_run * (float)a.q_passes) : 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] = fmaf(o[d], inv, acc[d]);
    __syncthreads();  // all reads of sQ / KV buffers by the tensor core are complete (bar_o)
  }

  // ---- write the row --------------------------------------------------------------
  if (row_ok) {
    const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, qrow);
    int64_t off;
    if (p.q_mean) off = (((int64_t)b * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.out_stride + head * DH;
    else off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.out_stride + head * DH;
    uint8_t* op = reinterpret_cast<uint8_t*>(p.out) + off * 2;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint4 u;
      u.x = pack_h2(acc[ch * 8 + 0], acc[ch * 8 + 1], bf16);
      u.y = pack_h2(acc[ch * 8 + 2], acc[ch * 8 + 3], bf16);
      u.z = pack_h2(acc[ch * 8 + 4], acc[ch * 8 + 5], bf16);
      u.w = pack_h2(acc[ch * 8 + 6], acc[ch * 8 + 7], bf16);
      *reinterpret_cast<uint4*>(op + ch * 16) = u;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}


// =====================================================================================
// Resident-KV kernel: one CTA per (batch row, window, head).  The window's K/V (<= 1024 keys) are
// loaded ONCE into shared memory; two independent 128-thread warpgroups walk over the 128-row
// query tiles / query views; each warpgroup has its own MMA-issuing thread (warps 8 / 9) and talks to
// it only through mbarriers -- the softmax threads never wait for each other.  The tensor-core work is
// software-pipelined over 64-key chunks so no MMA latency sits on the exp critical path:
//     iteration c:  issue S(c+1)=Q K(c+1)^T  |  softmax of S(c) (held in 64 registers)  |
//                   issue O(c)=P(c) V(c)     |  fold the finished O(c-1) into the running output
// S and O are double-buffered in TMEM (2x64 + 2x32 columns per warpgroup), P in shared memory.
// Same operand layouts as the streaming kernel above (V used as an MN-major B operand).
// =====================================================================================
constexpr int CH = 64;                                 // keys per chunk
constexpr int KV_CAP_KEYS = 1024;
constexpr int V2_THREADS = 384;   // 2 softmax warpgroups + 1 warpgroup holding the 2 MMA-issuer warps (8, 9)
constexpr int V2_REGS_SOFTMAX = 224, V2_REGS_ISSUER = 40;   // setmaxnreg split of the 64K register file
constexpr int SP2_BYTES = TILE_Q * CH * 2;             // 16 KB per P buffer
// fixed part: barriers/alignment + Q buffers (1 or 2 per warpgroup) + 2x2 P buffers;  + keys * (128 B K/V + 4 B mask)
__host__ __device__ constexpr int v2_fixed_smem(int q_dbuf) { return 1024 + 2 * (1 + q_dbuf) * SQ_BYTES + 2 * 2 * SP2_BYTES; }

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Volatile forms for the softmax inner loop of the four-warpgroup kernel: volatile asm statements keep their source
// order, so the packing of key group cc-1 stays BETWEEN the exponentials of groups cc and cc+1.  Left to itself ptxas
// sometimes hoists all 64 MUFU.EX2 of a chunk above every F2FP / STS (seen after an unrelated template change: +10 % on
// the stage-0 attention), which puts the whole pack + store tail behind the last exponential.
__device__ __forceinline__ float fast_exp2_v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2_v(float lo, float hi) {
  uint32_t d;
  if (BF16) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// The whole CTA is one continuous stream of 64-key chunks per warpgroup: chunk g = (pass P, chunk c)
// where a "pass" is one (query tile, query view) pair.  The issuer runs two chunks ahead across pass
// boundaries (the next pass' Q tile is prefetched into a second buffer while the current one is
// being processed), so the softmax threads never see a pipeline refill.
template <bool BF16>
__global__ void __launch_bounds__(V2_THREADS, 1) window_attn_v2_kernel(const __grid_constant__ AArgs a) {
  pdl_trigger();
  pdl_wait();
  // roles: warps 0-3 = softmax group 0, warps 4-7 = softmax group 1, warp 8 / 9 lane 0 = their MMA issuers
  // (warps 10, 11 only exist so that the issuers form a full warpgroup for setmaxnreg)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_ch = (a.Lk + CH - 1) / CH;
  const int keys_pad = n_ch * CH;
  const int q_dbuf = a.q_dbuf;                     // 1: two Q buffers per warpgroup (prefetch distance 1)
  uint8_t* sQ_all = smem;
  uint8_t* sP_all = sQ_all + 2 * (1 + q_dbuf) * SQ_BYTES;
  uint8_t* sK = sP_all + 4 * SP2_BYTES;
  uint8_t* sV = sK + keys_pad * 64;
  float* kmask = reinterpret_cast<float*>(sV + keys_pad * 64);
  __shared__ __align__(8) uint64_t bars[2][10];    // per warpgroup: S0 S1 | O0 O1 | P0 P1 | Q0 Q1 | F0 F1
  enum : uint32_t { BAR_S = 0, BAR_O = 16, BAR_P = 32, BAR_Q = 48, BAR_F = 64 };   // byte offsets inside bars[wg]
  __shared__ uint32_t tmem_base_s;

  const CobevtAttnParams& p = a.p;
  const int tid = threadIdx.x;
  const int warp = warp_index_uniform();                  // see sm100.cuh: keeps role / counter state in uniform registers
  const bool is_mma = warp >= 8;
  const int wg = is_mma ? warp - 8 : warp >> 2, wtid = tid & 127;
  int item = blockIdx.x;
  const int head = item % p.heads; item /= p.heads;
  const int win = item % a.n_win;
  const int b = item / a.n_win;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        mbar_init(&bars[i][j], 1); mbar_init(&bars[i][2 + j], 1); mbar_init(&bars[i][4 + j], 128); mbar_init(&bars[i][6 + j], 128);
        mbar_init(&bars[i][8 + j], 128);
      }
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  const uint8_t* qb = reinterpret_cast<const uint8_t*>(p.q);
  const uint8_t* kb_ = reinterpret_cast<const uint8_t*>(p.k);
  const uint8_t* vb_ = reinterpret_cast<const uint8_t*>(p.v);

  // ---- resident K / V: 4 lanes per token (64 contiguous bytes), all threads -----------------
  for (int idx = tid; idx < keys_pad * 4; idx += V2_THREADS) {
    const int tok = idx >> 2, piece = idx & 3;
    const uint8_t* srcK = kb_;
    const uint8_t* srcV = vb_;
    uint32_t nb = 0;
    float km = 0.f;
    if (tok < a.Lk) {
      const Tok t = token_of(p.kvmap, a.k_nx, a.k_ny, win, tok);
      if (t.in_range) {
        const int64_t off = ((((int64_t)b * p.kvmap.views + t.n) * p.kvmap.Hv + t.py) * p.kvmap.Wv + t.px) * p.kv_stride + head * DH;
        srcK = kb_ + off * 2 + piece * 16;
        srcV = vb_ + off * 2 + piece * 16;
        nb = 16;
        if (p.key_mask && piece == 0 &&
            p.key_mask[(((int64_t)b * p.kvmap.Hv + t.py) * p.kvmap.Wv + t.px) * p.kvmap.views + t.n] == 0)
          km = -INFINITY;
      } else if (p.k_pad) {
        srcK = reinterpret_cast<const uint8_t*>(p.k_pad) + head * DH * 2 + piece * 16;
        srcV = reinterpret_cast<const uint8_t*>(p.v_pad) + head * DH * 2 + piece * 16;
        nb = 16;
      }
    } else {
      km = -INFINITY;
    }
    cp_async16(sK + sw64_off(tok, piece), srcK, nb);
    cp_async16(sV + sw64_off(tok, piece), srcV, nb);
    if (piece == 0) kmask[tok] = km;
  }
  cp_async_commit();
  cp_async_wait<0>();
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s + wg * 256;   // S0 [0,64) S1 [64,128) O0 [128,160) O1 [160,192)
  uint8_t* sQ = sQ_all + wg * (1 + q_dbuf) * SQ_BYTES;
  uint8_t* sP = sP_all + wg * 2 * SP2_BYTES;
  const int rows_per_pass = p.q_mean ? p.qmap.w1 * p.qmap.w2 : a.Lq;
  const int n_my_tiles = (COBEVT_DBG(a, 8) && wg == 1) ? 0 : (a.q_tiles - wg + 1) >> 1;      // tiles wg, wg+2, ...
  const int NP = n_my_tiles * a.q_passes;                // passes of this warpgroup
  const int total = NP * n_ch;                           // chunks of this warpgroup

  if (is_mma) {
    // =============================== MMA issuer of warpgroup `wg` ===========================
    // S(k) = Q K(k)^T may overwrite TMEM buffer k&1 as soon as S(k-2) sits in the softmax threads'
    // registers (barrier F(k-2)); O(g) = P(g) V(g) follows the P(g) barrier.  The softmax threads hold
    // two chunks of logits in registers, so the issuer runs three chunks ahead of the exps.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(V2_REGS_ISSUER));
    if (warp < 10 && total > 0) {     // whole warp walks the protocol, one elected lane issues
      const uint64_t dQK = make_smem_desc(0, 16, 512, LAYOUT_SW64);
      const uint64_t dP = make_smem_desc(0, 16, 1024, LAYOUT_SW128);
      const uint64_t dV = make_smem_desc(0, 512, 512, LAYOUT_SW64);
      const uint32_t qk_lo = (uint32_t)dQK, qk_hi = (uint32_t)(dQK >> 32);
      const uint32_t p_lo = (uint32_t)dP, p_hi = (uint32_t)(dP >> 32);
      const uint32_t v_lo = (uint32_t)dV, v_hi = (uint32_t)(dV >> 32);
      const uint32_t q_addr = smem_u32(sQ) >> 4, k_addr = smem_u32(sK) >> 4, v_addr = smem_u32(sV) >> 4;
      const uint32_t p_addr = smem_u32(sP) >> 4;
      auto mma = [&](uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
      };
      int gq = 0, Pq = 0, cq = 0;          // next S chunk to issue: flat index, pass, chunk in pass
      auto next_qk = [&]() {                // S(gq) = Q(Pq) K(cq)^T  -> S buffer gq&1
        if (gq >= 2) {                      // S(gq-2) has left that buffer
          mbar_wait(&bars[wg][8 + (gq & 1)], (uint32_t)((gq - 2) >> 1) & 1u);
          tc_fence_after_sync();
        }
        if (cq == 0) {
          mbar_wait(&bars[wg][6 + (Pq & q_dbuf)], (uint32_t)(Pq >> q_dbuf) & 1u);
          tc_fence_after_sync();
        }
        const uint32_t qa = q_addr + (Pq & q_dbuf) * (SQ_BYTES / 16);
        const uint32_t kb = k_addr + cq * (CH * 64 / 16);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < DH / 16; ++ks)
            mma(tmem_base + (gq & 1) * CH, qk_lo | (qa + ks * 2), qk_hi, qk_lo | (kb + ks * 2), qk_hi, a.idesc_qk, ks ? 1u : 0u);
          umma_commit(&bars[wg][gq & 1]);
        }
        ++gq;
        if (++cq == n_ch) { cq = 0; ++Pq; }
      };
      next_qk();
      if (total > 1) next_qk();
      if (total > 2) next_qk();
      int c = 0;
      for (int g = 0; g < total; ++g) {
        mbar_wait(&bars[wg][4 + (g & 1)], (uint32_t)(g >> 1) & 1u);     // P(g) written, O(g-2) in registers
        tc_fence_after_sync();
        {                                                           // O(g) = P(g) V(c)  -> O buffer g&1
          const uint32_t pa = p_addr + (g & 1) * (SP2_BYTES / 16), va = v_addr + c * (CH * 64 / 16);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < CH / 16; ++ks)
              mma(tmem_base + 2 * CH + (g & 1) * DH, p_lo | (pa + ks * 2), p_hi, v_lo | (va + ks * 64), v_hi, a.idesc_pv, ks ? 1u : 0u);
            umma_commit(&bars[wg][2 + (g & 1)]);
          }
        }
        if (gq < total) next_qk();
        if (++c == n_ch) c = 0;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(V2_REGS_SOFTMAX));
    if (total > 0) {
    // =============================== softmax warpgroup `wg` ==================================
    // Register-level software pipeline, one 64-key chunk per step, two chunks of logits in registers:
    //   x  = S(g)   : exps + P(g) stores (64 MUFU, the step's critical resource)
    //   xn = S(g+1) : landed under the first exps; its scale / mask / bias / row max run on the ALU
    //                 pipes in the shadow of the exps of S(g)
    // and when x is dead it receives S(g+2) straight away.  O(g-1) is read from TMEM under the last
    // exps.  The only serial pieces of a step are the proxy fence + arrive and the 16-FMA fold of O.
    // The loop is unrolled by two so the arrays swap roles without register moves; the per-tile output
    // accumulator (mean over query views) lives in 32 spare TMEM columns.
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_acc = t_row + 192;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float L2E = 1.4426950408889634f;
    const bool any_mask = p.key_mask != nullptr;
    const bool bias_vec = (a.Lk & 3) == 0;
    const uint32_t p_swz = (smem_u32(sP) + (uint32_t)wtid * 128u) | (((uint32_t)wtid & 7u) << 4);
    const uint32_t bar_wg = smem_u32(&bars[wg][0]);
    const uint32_t km_s = smem_u32(kmask);
    const float inv_passes = 1.0f / (float)a.q_passes;

    // ---- Q tiles: this thread copies 4 x 16 B per pass; the row addresses are computed once per
    //      query tile (in 16-byte units), a pass (= query view) only adds the plane stride ----------
    uint32_t qoff16[4];
    const uint32_t plane16 = p.q_mean ? (uint32_t)(((int64_t)p.qmap.Hv * p.qmap.Wv * p.q_stride) >> 3) : 0u;
    const uint32_t q_dst = smem_u32(sQ);
    // S(k) is read from TMEM two steps before its exps, so the Q tile of a pass must be resident two
    // chunks before the pass starts: request it q_dist passes ahead
    const int q_dist = (q_dbuf && n_ch == 1) ? 2 : 1;
    const int q_issue_c = q_dbuf ? 0 : n_ch - 2;       // (single buffer: host guarantees n_ch >= 3)
    int nq = 0, nq_pass = 0, nq_tile = wg;             // next pass whose Q has not been requested yet
    auto issue_next_q = [&]() {
      if (nq_pass == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = wtid + 128 * i;
          const int r = idx >> 2, piece = idx & 3;
          const int rr = nq_tile * TILE_Q + r;
          qoff16[i] = 0xFFFFFFFFu;
          if (rr < rows_per_pass) {
            const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, rr);
            const int64_t off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.q_stride + head * DH;
            qoff16[i] = (uint32_t)(off >> 3) + piece;
          }
        }
      }
      const uint32_t dq = q_dst + (uint32_t)(nq & q_dbuf) * SQ_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = wtid + 128 * i;
        const bool ok = qoff16[i] != 0xFFFFFFFFu;
        const uint8_t* src = ok ? qb + ((size_t)(qoff16[i] + (uint32_t)nq_pass * plane16) << 4) : qb;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dq + sw64_off(idx >> 2, idx & 3)), "l"(src), "r"(ok ? 16u : 0u) : "memory");
      }
      cp_async_commit();
      ++nq;
      if (++nq_pass == a.q_passes) { nq_pass = 0; nq_tile += 2; }
    };
    int nq_done = 0;                                    // passes whose Q has been published to the issuer
    auto complete_q = [&]() {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      for (; nq_done < nq; ++nq_done) mbar_arrive_a(bar_wg + BAR_Q + (uint32_t)(nq_done & q_dbuf) * 8u);
    };
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;

    int g = 0, c = 0, P = 0, pass = 0, q_tile = wg;      // flat chunk, chunk in pass, flat pass, view, query tile
    float l_run = 0.f, l_fin = 0.f, alpha_prev = 0.f;
    float m_cur = -INFINITY, alpha_cur = 0.f;             // running max incl. chunk g / its rescale factor
    bool plain_cur = true;

    // the pass that just ended ((pass, q_tile) already advanced): out/acc (+)= o / (l * passes)
    auto finish_pass = [&]() {
      const float inv = (l_fin > 0.f) ? inv_passes / l_fin : 0.f;
      const bool last_of_tile = pass == 0;
      const bool first_of_tile = (last_of_tile ? a.q_passes - 1 : pass - 1) == 0;
      if (!first_of_tile) {
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_acc, ra);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DH; d += 2)
          ffma2(o[d], o[d + 1], o[d], o[d + 1], inv, inv, __uint_as_float(ra[d]), __uint_as_float(ra[d + 1]));
      } else {
#pragma unroll
        for (int d = 0; d < DH; d += 2) fmul2(o[d], o[d + 1], o[d], o[d + 1], inv, inv);
      }
      if (last_of_tile) {
        const int qrow = (q_tile - 2) * TILE_Q + wtid;
        if (qrow < rows_per_pass) {
          const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, qrow);
          int64_t off;
          if (p.q_mean) off = (((int64_t)b * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.out_stride + head * DH;
          else off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.out_stride + head * DH;
          uint8_t* op = reinterpret_cast<uint8_t*>(p.out) + off * 2;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 u;
            u.x = pack2<BF16>(o[ch * 8 + 0], o[ch * 8 + 1]);
            u.y = pack2<BF16>(o[ch * 8 + 2], o[ch * 8 + 3]);
            u.z = pack2<BF16>(o[ch * 8 + 4], o[ch * 8 + 5]);
            u.w = pack2<BF16>(o[ch * 8 + 6], o[ch * 8 + 7]);
            *reinterpret_cast<uint4*>(op + ch * 16) = u;
          }
        }
      } else {
        uint32_t ra[32];
#pragma unroll
        for (int d = 0; d < DH; ++d) ra[d] = __float_as_uint(o[d]);
        tmem_st_32x32b_x32(t_acc, ra);
        tmem_st_wait();
      }
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = 0.f;
    };

    // scale / mask / bias of a chunk of raw logits (chunk `cc` of a pass of query tile `qt`) and its
    // running row max on top of `m_base`; returns the new running max.  `plain`: x stays raw (the scale
    // is folded into the exp), otherwise x is rewritten as log2-domain logits.
    auto prep_max = [&](float (&x)[CH], int cc, int qt, float m_base, bool& plain) -> float {
      const int key0 = cc * CH;
      const bool use_km = any_mask || (key0 + CH > a.Lk);
      plain = !use_km && !(p.bias != nullptr);
      float mx;
      if (plain) {
        float mr[4] = {x[0], x[1], x[2], x[3]};          // 4 independent chains (ILP)
#pragma unroll
        for (int j = 4; j < CH; ++j) mr[j & 3] = fmaxf(mr[j & 3], x[j]);
        mx = fmaxf(m_base, fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3])) * sl2);
      } else {
        if (use_km) {
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) {
            const float4 k4 = lds128f(km_s + (uint32_t)(key0 + 4 * j) * 4u);
            ffma2(x[4 * j + 0], x[4 * j + 1], x[4 * j + 0], x[4 * j + 1], sl2, sl2, k4.x, k4.y);
            ffma2(x[4 * j + 2], x[4 * j + 3], x[4 * j + 2], x[4 * j + 3], sl2, sl2, k4.z, k4.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; j += 2) fmul2(x[j], x[j + 1], x[j], x[j + 1], sl2, sl2);
        }
        const int qrow = qt * TILE_Q + wtid;
        if (p.bias && qrow < rows_per_pass) {
          const float* bp = p.bias + ((int64_t)head * a.Lq + qrow) * a.Lk + key0;
          if (bias_vec && key0 + CH <= a.Lk) {
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
              ffma2(x[4 * j + 0], x[4 * j + 1], b4.x, b4.y, L2E, L2E, x[4 * j + 0], x[4 * j + 1]);
              ffma2(x[4 * j + 2], x[4 * j + 3], b4.z, b4.w, L2E, L2E, x[4 * j + 2], x[4 * j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (key0 + j < a.Lk) x[j] = fmaf(__ldg(bp + j), L2E, x[j]);
          }
        }
        float mr[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
        for (int j = 4; j < CH; ++j) mr[j & 3] = fmaxf(mr[j & 3], x[j]);
        mx = fmaxf(m_base, fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3])));
      }
      return mx;
    };

    // one chunk.  x = logits of chunk g (prepared, max known); xn = raw logits of chunk g+1 (in flight)
    auto step = [&](auto par_c, float (&x)[CH], float (&xn)[CH]) {
      constexpr int PAR = decltype(par_c)::value;       // == g & 1: buffer indices are compile-time
      if (c == 0) l_run = 0.f;
      const bool q_now = (c == q_issue_c) && (nq < NP) && (nq <= P + q_dist);
      if (q_now) issue_next_q();            // (the tensor core has finished reading the buffer it overwrites:
                                            //  this thread has waited for every S of the pass that used it)
      const float m_safe = (m_cur == -INFINITY) ? 0.f : m_cur;
      const float nm = -m_safe;
      const float es = plain_cur ? sl2 : 1.0f;            // exp2(x * es - m)
      float sm[4] = {0.f, 0.f, 0.f, 0.f};
      // exps of 8 logits + their 16 bytes of P(g)
      auto exps8 = [&](int cc) {
        const int j = cc * 8;
        ffma2(x[j], x[j + 1], x[j], x[j + 1], es, es, nm, nm);
        ffma2(x[j + 2], x[j + 3], x[j + 2], x[j + 3], es, es, nm, nm);
        ffma2(x[j + 4], x[j + 5], x[j + 4], x[j + 5], es, es, nm, nm);
        ffma2(x[j + 6], x[j + 7], x[j + 6], x[j + 7], es, es, nm, nm);
        if (!COBEVT_DBG(a, 1)) {
#pragma unroll
          for (int t = 0; t < 8; ++t) x[j + t] = fast_exp2(x[j + t]);
        }
        fadd2(sm[0], sm[1], sm[0], sm[1], x[j], x[j + 1]);
        fadd2(sm[2], sm[3], sm[2], sm[3], x[j + 2], x[j + 3]);
        fadd2(sm[0], sm[1], sm[0], sm[1], x[j + 4], x[j + 5]);
        fadd2(sm[2], sm[3], sm[2], sm[3], x[j + 6], x[j + 7]);
        if (!COBEVT_DBG(a, 2))
        sts128((p_swz ^ ((uint32_t)cc << 4)) + PAR * SP2_BYTES,
               pack2<BF16>(x[j + 0], x[j + 1]), pack2<BF16>(x[j + 2], x[j + 3]),
               pack2<BF16>(x[j + 4], x[j + 5]), pack2<BF16>(x[j + 6], x[j + 7]));
      };
      exps8(0);
      exps8(1);
      if (q_now) complete_q();
      tmem_ld_wait();                        // xn = S(g+1) has landed
      tc_fence_before_sync();
      if (g + 1 < total) mbar_arrive_a(bar_wg + BAR_F + (PAR ^ 1) * 8);   // its TMEM buffer is free
      // ---- exps 16..47, with the preparation of the next chunk in their shadow --------------------
      exps8(2);
      exps8(3);
      float m_next = m_cur, alpha_next = 0.f;
      bool plain_next = true;
      if (g + 1 < total) {
        const bool new_pass = (c + 1 == n_ch);
        const float m_base = new_pass ? -INFINITY : m_cur;
        m_next = prep_max(xn, new_pass ? 0 : c + 1, (new_pass && pass + 1 == a.q_passes) ? q_tile + 2 : q_tile, m_base, plain_next);
        alpha_next = fast_exp2(m_base - ((m_next == -INFINITY) ? 0.f : m_next));
      }
      exps8(4);
      exps8(5);
      // ---- O(g-1): its load lands under the last exps ------------------------------------------
      uint32_t ro[32];
      {
        constexpr int ob = PAR ^ 1;         // == (g - 1) & 1
        if (g > 0) mbar_wait_a(bar_wg + BAR_O + ob * 8, (uint32_t)((g - 1) >> 1) & 1u);
        tc_fence_after_sync();
        tmem_ld_32x32b_x32(t_row + 2 * CH + ob * DH, ro);
      }
      exps8(6);
      exps8(7);
      l_run = l_run * alpha_cur + ((sm[0] + sm[1]) + (sm[2] + sm[3]));
      if (!COBEVT_DBG(a, 4)) fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive_a(bar_wg + BAR_P + PAR * 8);
      tmem_ld_wait();                       // ro has landed
      // ---- fold the previous chunk's O (possibly the previous pass' last) -------------------
      if (g > 0) {
#pragma unroll
        for (int d = 0; d < DH; d += 2)
          ffma2(o[d], o[d + 1], o[d], o[d + 1], alpha_prev, alpha_prev, __uint_as_float(ro[d]), __uint_as_float(ro[d + 1]));
        if (c == 0) finish_pass();          // that was the last chunk of the previous pass
      }
      // ---- x is dead: it receives S(g+2) ------------------------------------------------------
      {
        if (g + 2 < total) mbar_wait_a(bar_wg + BAR_S + PAR * 8, (uint32_t)((g + 2) >> 1) & 1u);
        tc_fence_after_sync();
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(t_row + PAR * CH, r0);
        tmem_ld_32x32b_x32(t_row + PAR * CH + 32, r1);
#pragma unroll
        for (int j = 0; j < 32; ++j) { x[j] = __uint_as_float(r0[j]); x[32 + j] = __uint_as_float(r1[j]); }
      }
      alpha_prev = alpha_cur;
      m_cur = m_next; alpha_cur = alpha_next; plain_cur = plain_next;
      ++g;
      if (++c == n_ch) {
        c = 0; l_fin = l_run; ++P;
        if (++pass == a.q_passes) { pass = 0; q_tile += 2; }
      }
    };

    // ---- prologue: Q of the first pass(es), S(0) prepared, S(1) in flight -------------------------
    issue_next_q();
    if (q_dist == 2 && NP > 1) issue_next_q();
    complete_q();
    float xa[CH], xb[CH];
    {
      mbar_wait_a(bar_wg + BAR_S, 0);
      tc_fence_after_sync();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(t_row, r0);
      tmem_ld_32x32b_x32(t_row + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) { xa[j] = __uint_as_float(r0[j]); xa[32 + j] = __uint_as_float(r1[j]); }
      tc_fence_before_sync();
      mbar_arrive_a(bar_wg + BAR_F);
      m_cur = prep_max(xa, 0, q_tile, -INFINITY, plain_cur);
      alpha_cur = 0.f;
      if (total > 1) mbar_wait_a(bar_wg + BAR_S + 8, 0);
      tc_fence_after_sync();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(t_row + CH, s0);
      tmem_ld_32x32b_x32(t_row + CH + 32, s1);
#pragma unroll
      for (int j = 0; j < 32; ++j) { xb[j] = __uint_as_float(s0[j]); xb[32 + j] = __uint_as_float(s1[j]); }
    }
    while (g < total) {
      step(std::integral_constant<int, 0>{}, xa, xb);
      if (g < total) step(std::integral_constant<int, 1>{}, xb, xa);
    }
    tmem_ld_wait();                          // drain the last (stale) logit prefetch
    {   // the last chunk's O
      const int ob = (g - 1) & 1;
      mbar_wait_a(bar_wg + BAR_O + (uint32_t)ob * 8u, (uint32_t)((g - 1) >> 1) & 1u);
      tc_fence_after_sync();
      uint32_t ro[32];
      tmem_ld_32x32b_x32(t_row + 2 * CH + ob * DH, ro);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < DH; d += 2)
        ffma2(o[d], o[d + 1], o[d], o[d + 1], alpha_prev, alpha_prev, __uint_as_float(ro[d]), __uint_as_float(ro[d + 1]));
      finish_pass();
    }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

// =====================================================================================
// Four-warpgroup resident-KV kernel.  The softmax of this head size (32) is bound by per-warp
// instruction latency, not by any pipe: one query row per thread needs ~100 registers for a 32-key
// chunk, so four softmax warpgroups (16 warps, 112 registers each) fit where the two-warpgroup
// kernel above (64-key chunks, 224 registers) has only 8 warps to hide its dependency chains.
//   warps 0-15   softmax warpgroups 0..3        warps 16-19   lane 0 = MMA issuer of warpgroup (warp-16)
// Work split inside the CTA (one CTA = one (batch row, window, head), K/V resident as above):
//   * >= 3 query tiles: warpgroup w takes tiles w, w+4, ... with all their passes (query views);
//   * 1 or 2 query tiles: the G = 4 / tiles warpgroups that share a tile split its passes (mean over
//     query views: partial means add up) or, with a single pass, its key chunks (independent online
//     softmax streams, log-sum-exp merge); the partials meet once, at the end, through shared memory.
// TMEM per warpgroup (128 columns): S [0,64) O [64,96) ACC [96,128); S and the P tile in shared memory are
// single-buffered -- the other three warpgroups cover the tensor-core round trips.  O accumulates in TMEM over
// the chunks of a pass (lazy rescale, see below): no tensor-core round trip sits on the per-chunk path.
// =====================================================================================
constexpr int CH3 = 64;                                  // keys per chunk
constexpr int V3_WG = 4;
constexpr int V3_THREADS = (V3_WG + 1) * 128;            // 640
constexpr int V3_REGS_SOFTMAX = 112, V3_REGS_ISSUER = 32;   // the CTA pool is 640 x 96 = 61440 = 16*32*112 + 4*32*32
constexpr int SP3_BYTES = TILE_Q * CH3 * 2;              // 16 KB: one P buffer per warpgroup
constexpr int V3_POLY_DEFAULT = 1;     // 2 of 8 key groups.  Null result while the chunk loop still spilled (profiles/r02_attn_ab.log);
                                       // with the MUFU queue as the top stall: attention family 3183 -> 3077 us (profiles/r02g_attn_bias_ab.log)
__host__ __device__ constexpr int v3_fixed_smem() { return 1024 + V3_WG * (2 * SQ_BYTES + SP3_BYTES); }

// POLY: bit cc set -> the 8 exponentials of 8-key group cc (of the 8 groups of a 64-key chunk) use poly_exp2_pair
// QLIN: the four query rows a thread fetches per pass (rows r, r+32, r+64, r+96 of the tile) are a constant address step
//       apart (one query plane per pass and 32 % window width == 0, decided on the host), so ONE offset is carried
//       through the chunk loop instead of four -- the loop state of this kernel spills.
// BIAS: the additive-bias code (and the per-CTA chunk rotation that goes with it) exists only in the instances the
//       biased attentions use; the unbiased launches keep their shorter chunk loop.
template <bool BF16, int POLY, bool QLIN, bool BIAS>
__global__ void __launch_bounds__(V3_THREADS, 1) window_attn_v3_kernel(const __grid_constant__ AArgs a) {
  pdl_trigger();   // barrier init and the TMEM allocation below may overlap the previous kernel's tail
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_ch = (a.Lk + CH3 - 1) / CH3;
  const int keys_pad = n_ch * CH3;
  uint8_t* sQ_all = smem;                                   // per warpgroup: 2 x 8 KB
  uint8_t* sP_all = sQ_all + V3_WG * 2 * SQ_BYTES;          // per warpgroup: 16 KB (also the merge scratch)
  uint8_t* sK = sP_all + V3_WG * SP3_BYTES;
  uint8_t* sV = sK + keys_pad * 64;
  float* kmask = reinterpret_cast<float*>(sV + keys_pad * 64);
  __shared__ __align__(8) uint64_t bars[V3_WG][8];          // per warpgroup: S F | O - | P - | Q0 Q1
  enum : uint32_t { BAR_S = 0, BAR_F = 8, BAR_O = 16, BAR_P = 32, BAR_Q = 48 };
  __shared__ __align__(8) uint64_t bar_k, bar_v;            // K / V windows landed (TMA path): S = Q K^T starts as soon as K
                                                            // is in, the V window keeps streaming under the first softmax.
  // This kernel sits at its register limit (112 / 32 after setmaxnreg): two further changes were measured SLOWER and
  // removed -- K as several TMA boxes with their own barriers (-4 %, profiles/r02_attn_kbox_ab.log) and fetching the dense
  // bias of a chunk before the S wait (-60 %: 490 more spill bytes land in every path, profiles/r02_attn_bias_prefetch_ab.log)
  __shared__ float2 sML[V3_WG][TILE_Q];                     // (running max, sum) partials of the key-split merge
  __shared__ uint32_t tmem_base_s;

  const CobevtAttnParams& p = a.p;
  const int tid = threadIdx.x;
  const int warp = warp_index_uniform();                  // see sm100.cuh: keeps role / counter state in uniform registers
  const bool is_mma = warp >= 4 * V3_WG;
  const int wg = is_mma ? warp - 4 * V3_WG : warp >> 2, wtid = tid & 127;
  int item = blockIdx.x;
  const int seg = item % a.kv_segs; item /= a.kv_segs;          // key segment (1 unless key-segment mode)
  const int head = item % p.heads; item /= p.heads;
  const int win = item % a.n_win;
  const int b = item / a.n_win;
  const int view0 = p.kv_view0 + seg * p.kvmap.views;           // first K/V view of this CTA

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < V3_WG; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        mbar_init(&bars[i][j], j == 0 ? 1 : 128); mbar_init(&bars[i][2 + j], 1); mbar_init(&bars[i][4 + j], 128); mbar_init(&bars[i][6 + j], 128);
      }
    mbar_init(&bar_k, 1);
    mbar_init(&bar_v, 1);
    mbar_fence_init();
    pdl_wait();
    if (a.kv_tma) {
      // the whole key window of every view is one strided box: 2 instructions instead of 16 B per thread
      const int X = win / a.k_ny, Y = win - X * a.k_ny;
      const int c1 = p.kvmap.grid ? Y : Y * p.kvmap.w2, c2 = p.kvmap.grid ? X : X * p.kvmap.w1;
      mbar_arrive_expect_tx(&bar_k, (uint32_t)a.Lk * 64u);
      tma_load_4d(sK, &a.tmK, &bar_k, head * DH, c1, c2, b * a.kv_vt + view0);
      mbar_arrive_expect_tx(&bar_v, (uint32_t)a.Lk * 64u);
      tma_load_4d(sV, &a.tmV, &bar_v, head * DH, c1, c2, b * a.kv_vt + view0);
    }
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  pdl_wait();                                               // Q, K, V, masks: outputs of earlier kernels
  const uint8_t* qb = reinterpret_cast<const uint8_t*>(p.q);
  const uint8_t* kb_ = reinterpret_cast<const uint8_t*>(p.k);
  const uint8_t* vb_ = reinterpret_cast<const uint8_t*>(p.v);

  if (a.kv_tma) {
    // key mask vector + zero rows behind the window (their P is 0, but 0 x garbage could be NaN)
    for (int tok = tid; tok < keys_pad; tok += V3_THREADS) {
      float km = 0.f;
      if (tok >= a.Lk) {
        km = -INFINITY;
#pragma unroll
        for (int piece = 0; piece < 4; ++piece) {
          *reinterpret_cast<uint4*>(sK + sw64_off(tok, piece)) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sV + sw64_off(tok, piece)) = make_uint4(0, 0, 0, 0);
        }
      } else if (p.key_mask) {
        const Tok t = token_of(p.kvmap, a.k_nx, a.k_ny, win, tok);
        if (p.key_mask[(((int64_t)b * p.kvmap.Hv + t.py) * p.kvmap.Wv + t.px) * p.kvmap.views + t.n] == 0) km = -INFINITY;
      }
      kmask[tok] = km;
    }
  } else
  // ---- resident K / V by cp.async: 4 lanes per token (64 contiguous bytes), all threads --------
  for (int idx = tid; idx < keys_pad * 4; idx += V3_THREADS) {
    const int tok = idx >> 2, piece = idx & 3;
    const uint8_t* srcK = kb_;
    const uint8_t* srcV = vb_;
    uint32_t nb = 0;
    float km = 0.f;
    if (tok < a.Lk) {
      const Tok t = token_of(p.kvmap, a.k_nx, a.k_ny, win, tok);
      if (t.in_range) {
        const int64_t off = ((((int64_t)b * a.kv_vt + view0 + t.n) * p.kvmap.Hv + t.py) * p.kvmap.Wv + t.px) * p.kv_stride + head * DH;
        srcK = kb_ + off * 2 + piece * 16;
        srcV = vb_ + off * 2 + piece * 16;
        nb = 16;
        if (p.key_mask && piece == 0 &&
            p.key_mask[(((int64_t)b * p.kvmap.Hv + t.py) * p.kvmap.Wv + t.px) * p.kvmap.views + t.n] == 0)
          km = -INFINITY;
      } else if (p.k_pad) {
        srcK = reinterpret_cast<const uint8_t*>(p.k_pad) + head * DH * 2 + piece * 16;
        srcV = reinterpret_cast<const uint8_t*>(p.v_pad) + head * DH * 2 + piece * 16;
        nb = 16;
      }
    } else {
      km = -INFINITY;
    }
    cp_async16(sK + sw64_off(tok, piece), srcK, nb);
    cp_async16(sV + sw64_off(tok, piece), srcV, nb);
    if (piece == 0) kmask[tok] = km;
  }
  cp_async_commit();
  cp_async_wait<0>();
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s + wg * 128;
  uint8_t* sQ = sQ_all + wg * 2 * SQ_BYTES;
  uint8_t* sP = sP_all + wg * SP3_BYTES;
  const int rows_per_pass = p.q_mean ? p.qmap.w1 * p.qmap.w2 : a.Lq;

  // ---- work split ---------------------------------------------------------------------------
  const int T = a.q_tiles >= 3 ? V3_WG : a.q_tiles;       // warpgroups are spread over T tiles at a time
  const int G = V3_WG / T;                                // warpgroups sharing one tile (1, 2 or 4)
  const int t0 = wg % T, sub = wg / T;
  const bool split_pass = G > 1 && a.q_passes > 1;        // the group splits the query views
  const bool split_keys = G > 1 && a.q_passes == 1;       // the group splits the key chunks
  const int n_my_tiles = t0 < a.q_tiles ? (a.q_tiles - t0 + T - 1) / T : 0;    // == 1 whenever G > 1
  const int passes_per_tile = split_pass ? (a.q_passes > sub ? (a.q_passes - sub + G - 1) / G : 0) : a.q_passes;
  const int n_uch = split_keys ? (n_ch > sub ? (n_ch - sub + G - 1) / G : 0) : n_ch;   // chunks per pass
  const int c_first = split_keys ? sub : 0, c_step = split_keys ? G : 1;
  // Biased attentions: every CTA of a head reads the same dense bias, and CTAs launched together run in lockstep, so
  // dozens of SMs ask the L2 for the same lines at the same moment (measured: the FuseBEVT attentions got SLOWER with
  // the coalesced bias layout, which concentrates a warp's request on 4 lines).  Each CTA therefore walks its key
  // chunks from a different starting chunk (the online softmax does not care about the order).
  // (the start chunk depends on the window only -- never on the batch row -- so a sample's bits do not depend on its
  //  slot in the batch: the N-rank shard check compares bit for bit)
  const int c_rot = (BIAS && n_uch > 1) ? win % n_uch : 0;
  auto chunk_of = [&](int kk) {                               // step kk of a pass -> key chunk
    int r = kk + c_rot;
    if (r >= n_uch) r -= n_uch;
    return c_first + r * c_step;
  };
  const int NP = n_my_tiles * passes_per_tile;
  const int total = (COBEVT_DBG(a, 8) && wg > 0) ? 0 : NP * n_uch;

  if (is_mma) {
    // =============================== MMA issuer of warpgroup `wg` ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(V3_REGS_ISSUER));
    // the whole warp walks the protocol (waits, counters, descriptor arithmetic stay warp-uniform), one
    // elected lane issues the tcgen05 instructions
    if (total > 0) {
      const uint64_t dA = make_smem_desc(0, 16, 512, LAYOUT_SW64);      // Q, K: K-major 64-byte rows
      const uint64_t dP = make_smem_desc(0, 16, 1024, LAYOUT_SW128);    // P: K-major 128-byte rows
      const uint64_t dV = make_smem_desc(0, 512, 512, LAYOUT_SW64);     // V: MN-major
      const uint32_t a_lo = (uint32_t)dA, a_hi = (uint32_t)(dA >> 32);
      const uint32_t p_lo = (uint32_t)dP, p_hi = (uint32_t)(dP >> 32);
      const uint32_t v_lo = (uint32_t)dV, v_hi = (uint32_t)(dV >> 32);
      const uint32_t q_addr = smem_u32(sQ) >> 4, k_addr = smem_u32(sK) >> 4, v_addr = smem_u32(sV) >> 4;
      const uint32_t p_addr = smem_u32(sP) >> 4;
      auto mma = [&](uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
      };
      if (a.kv_tma) mbar_wait(&bar_k, 0);
      int gq = 0, Pq = 0, kq = 0;          // next S chunk to issue: flat index, pass, chunk in pass
      auto next_qk = [&]() {                // S(gq) = Q(Pq) K(chunk)^T
        if (kq == 0) {
          mbar_wait(&bars[wg][6 + (Pq & 1)], (uint32_t)(Pq >> 1) & 1u);
          tc_fence_after_sync();
        }
        const uint32_t qa = q_addr + (Pq & 1) * (SQ_BYTES / 16);
        const uint32_t kb = k_addr + chunk_of(kq) * (CH3 * 64 / 16);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < DH / 16; ++ks)
            mma(tmem_base, a_lo | (qa + ks * 2), a_hi, a_lo | (kb + ks * 2), a_hi, a.idesc_qk, ks ? 1u : 0u);
          umma_commit(&bars[wg][0]);
        }
        ++gq;
        if (++kq == n_uch) { kq = 0; ++Pq; }
      };
      next_qk();
      int k = 0;
      for (int g = 0; g < total; ++g) {
        if (gq < total) {                   // S(g) has left TMEM: S(g+1) may overwrite it
          mbar_wait(&bars[wg][1], (uint32_t)g & 1u);
          tc_fence_after_sync();
          next_qk();
        }
        mbar_wait(&bars[wg][4], (uint32_t)g & 1u);                  // P(g) written
        if (g == 0 && a.kv_tma) mbar_wait(&bar_v, 0);               // the V window has landed
        tc_fence_after_sync();
        {                                                           // O (+)= P(g) V(chunk)
          const uint32_t va = v_addr + chunk_of(k) * (CH3 * 64 / 16);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < CH3 / 16; ++ks)
              mma(tmem_base + CH3, p_lo | (p_addr + ks * 2), p_hi, v_lo | (va + ks * 64), v_hi, a.idesc_pv, (k | ks) ? 1u : 0u);
            umma_commit(&bars[wg][2]);
          }
        }
        if (++k == n_uch) k = 0;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(V3_REGS_SOFTMAX));
    // =============================== softmax warpgroup `wg` ==================================
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_o = t_row + CH3, t_acc = t_row + CH3 + DH;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float L2E = 1.4426950408889634f;
    const bool any_mask = p.key_mask != nullptr;
    const bool bias_vec = ((a.Lk | a.bias_ld) & 3) == 0;
    const int bias_k0 = view0 * (p.kvmap.w1 * p.kvmap.w2);     // first key of this CTA's views within the whole window
    const uint32_t p_swz = (smem_u32(sP) + (uint32_t)wtid * 128u) | (((uint32_t)wtid & 7u) << 4);          // SW128 row
    const uint32_t bar_wg = smem_u32(&bars[wg][0]);
    const uint32_t km_s = smem_u32(kmask);
    const float inv_passes = 1.0f / (float)a.q_passes;
    float o[DH];                                           // the warpgroup's last (or only) tile result
    float m_fin = -INFINITY, l_fin = 0.f;                  // state of the last finished pass
    bool have_acc = false;                                 // ACC (TMEM) holds earlier passes of the current tile

    if (total > 0) {
    // ---- Q tiles (see the two-warpgroup kernel) -------------------------------------------------
    uint32_t qoff16[QLIN ? 1 : 4];
    const uint32_t plane16 = p.q_mean ? (uint32_t)(((int64_t)p.qmap.Hv * p.qmap.Wv * p.q_stride) >> 3) : 0u;
    const uint32_t q_dst = smem_u32(sQ);
    // request the Q tile of pass Pn = pass pjn of tile `tile` (the caller derives both from the loop's own counters)
    auto issue_q = [&](int Pn, int pjn, int tile) {
      if (pjn == 0) {
        if constexpr (QLIN) {
          const int rr = tile * TILE_Q + (wtid >> 2);
          const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, rr < rows_per_pass ? rr : 0);
          const int64_t off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.q_stride + head * DH;
          qoff16[0] = (uint32_t)(off >> 3) + (uint32_t)(wtid & 3);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int idx = wtid + 128 * i;
            const int r = idx >> 2, piece = idx & 3;
            const int rr = tile * TILE_Q + r;
            qoff16[i] = 0xFFFFFFFFu;
            if (rr < rows_per_pass) {
              const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, rr);
              const int64_t off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.q_stride + head * DH;
              qoff16[i] = (uint32_t)(off >> 3) + piece;
            }
          }
        }
      }
      const uint32_t view = split_pass ? (uint32_t)(sub + pjn * G) : (uint32_t)pjn;
      const uint32_t dq = q_dst + (uint32_t)(Pn & 1) * SQ_BYTES;
      // QLIN: 32 tile rows further = 32 / w2 window rows further (x nx pixel rows for the dilated grid partition)
      const uint32_t step16 = QLIN ? (uint32_t)(((32 / p.qmap.w2) * (p.qmap.grid ? a.q_nx : 1) * p.qmap.Wv * p.q_stride) >> 3) : 0u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = wtid + 128 * i;
        bool ok;
        uint32_t o16;
        if constexpr (QLIN) {
          ok = tile * TILE_Q + (idx >> 2) < rows_per_pass;
          o16 = qoff16[0] + (uint32_t)i * step16;
        } else {
          ok = qoff16[i] != 0xFFFFFFFFu;
          o16 = qoff16[i];
        }
        const uint8_t* src = ok ? qb + ((size_t)(o16 + view * plane16) << 4) : qb;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dq + sw64_off(idx >> 2, idx & 3)), "l"(src), "r"(ok ? 16u : 0u) : "memory");
      }
      cp_async_commit();
    };
    auto complete_q = [&](int Pn) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive_a(bar_wg + BAR_Q + (uint32_t)(Pn & 1) * 8u);
    };
    auto write_rows = [&](int q_tile, float (&o)[DH]) {    // o -> output rows of query tile q_tile
      const int qrow = q_tile * TILE_Q + wtid;
      if (qrow < rows_per_pass) {
        const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, qrow);
        int64_t off;
        if (p.q_mean) off = (((int64_t)b * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.out_stride + head * DH;
        else off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.out_stride + head * DH;
        uint8_t* op = reinterpret_cast<uint8_t*>(p.out) + off * 2;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 u;
          u.x = pack2<BF16>(o[ch * 8 + 0], o[ch * 8 + 1]);
          u.y = pack2<BF16>(o[ch * 8 + 2], o[ch * 8 + 3]);
          u.z = pack2<BF16>(o[ch * 8 + 4], o[ch * 8 + 5]);
          u.w = pack2<BF16>(o[ch * 8 + 6], o[ch * 8 + 7]);
          *reinterpret_cast<uint4*>(op + ch * 16) = u;
        }
      }
    };

    int g = 0, k = 0, P = 0, pj = 0, q_tile = t0;        // flat chunk, chunk in pass, flat pass, pass in tile, tile
    float m_run = 0.f, l_run = 0.f;

    // the pass that just ended ((pj, q_tile) already advanced).  Key split: keep (m, l, o) raw for the
    // merge.  Otherwise o <- o / (l * passes) + earlier passes of this tile; the tile's last pass leaves
    // the sum in `o` (written out, or handed to the pass-split merge), the others park it in TMEM.
    auto finish_pass = [&](bool last_unit, float (&o)[DH]) {
      if (p.part) {     // key-segment mode: un-normalised row + (max, sum) for cobevt_attn_merge; one pass per tile
        const int qrow = (last_unit ? q_tile : q_tile - T) * TILE_Q + wtid;
        if (qrow < rows_per_pass) {
          float4* dst = reinterpret_cast<float4*>(p.part + seg * p.part_segment_stride +
                                                  ((((int64_t)b * a.n_win + win) * p.heads + head) * a.Lq + qrow) * 36);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) dst[ch] = make_float4(o[ch * 4 + 0], o[ch * 4 + 1], o[ch * 4 + 2], o[ch * 4 + 3]);
          dst[8] = make_float4(m_fin, l_fin, 0.f, 0.f);
        }
        return;
      }
      if (split_keys) return;
      const float inv = (l_fin > 0.f) ? inv_passes / l_fin : 0.f;
      const bool last_of_tile = last_unit || pj == 0;
      if (have_acc) {
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_acc, ra);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DH; d += 2)
          ffma2(o[d], o[d + 1], o[d], o[d + 1], inv, inv, __uint_as_float(ra[d]), __uint_as_float(ra[d + 1]));
      } else {
#pragma unroll
        for (int d = 0; d < DH; d += 2) fmul2(o[d], o[d + 1], o[d], o[d + 1], inv, inv);
      }
      if (last_of_tile) {
        have_acc = false;
        if (G == 1) {
          write_rows(last_unit ? q_tile : q_tile - T, o);
        }
      } else {
        uint32_t ra[32];
#pragma unroll
        for (int d = 0; d < DH; ++d) ra[d] = __float_as_uint(o[d]);
        tmem_st_32x32b_x32(t_acc, ra);
        tmem_st_wait();
        have_acc = true;
      }
    };
    // the finished pass' O: TMEM -> o
    auto load_o = [&](int glast, float (&o)[DH]) {
      mbar_wait_a(bar_wg + BAR_O, (uint32_t)glast & 1u);
      tc_fence_after_sync();
      uint32_t ro[32];
      tmem_ld_32x32b_x32(t_o, ro);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = __uint_as_float(ro[d]);
      tc_fence_before_sync();
    };

    issue_q(0, 0, t0);
    complete_q(0);
    for (; g < total; ++g) {
      const bool q_now = (k == 0) && (P + 1 < NP);
      if (q_now) {                                  // pass P + 1: the next view of this tile, or view 0 of the next tile
        const bool wrap = pj + 1 == passes_per_tile;
        issue_q(P + 1, wrap ? 0 : pj + 1, wrap ? q_tile + T : q_tile);
      }
      const int key0 = chunk_of(k) * CH3;
      mbar_wait_fast(bar_wg + BAR_S, (uint32_t)g & 1u);
      tc_fence_after_sync();
      float x[CH3];
      {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(t_row, r0);
        tmem_ld_32x32b_x32(t_row + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { x[j] = __uint_as_float(r0[j]); x[32 + j] = __uint_as_float(r1[j]); }
      }
      tc_fence_before_sync();
      mbar_arrive_a(bar_wg + BAR_F);                // S(g) sits in registers: its TMEM columns are free
      const bool use_km = any_mask || (key0 + CH3 > a.Lk);
      const bool plain = !use_km && !BIAS;
      float cmax;                                   // chunk max, log2 domain
      if (plain) {
        float mr[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
        for (int j = 4; j < CH3; ++j) mr[j & 3] = fmaxf(mr[j & 3], x[j]);
        cmax = fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3])) * sl2;
      } else {
        if (use_km) {
#pragma unroll
          for (int j = 0; j < CH3 / 4; ++j) {
            const float4 k4 = lds128f(km_s + (uint32_t)(key0 + 4 * j) * 4u);
            ffma2(x[4 * j + 0], x[4 * j + 1], x[4 * j + 0], x[4 * j + 1], sl2, sl2, k4.x, k4.y);
            ffma2(x[4 * j + 2], x[4 * j + 3], x[4 * j + 2], x[4 * j + 3], sl2, sl2, k4.z, k4.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH3; j += 2) fmul2(x[j], x[j + 1], x[j], x[j + 1], sl2, sl2);
        }
        const int qrow = q_tile * TILE_Q + wtid;
        if (BIAS && p.bias_layout == 1 && qrow < rows_per_pass) {
          // key-group-major bias (heads, Lk/4, Lq, 4): lane = query row, so one load instruction of the warp covers 512
          // contiguous bytes (4 L1 wavefronts) instead of 32 rows 4 x Lk bytes apart (32 wavefronts) -- the row-major
          // form made the biased attentions L1-wavefront bound (2.2 us per 64-key chunk step)
          const float4* bp4 = reinterpret_cast<const float4*>(p.bias) +
                              ((int64_t)head * (a.bias_ld >> 2) + ((bias_k0 + key0) >> 2)) * a.Lq + qrow;
#pragma unroll
          for (int j = 0; j < CH3 / 4; ++j) {
            if (key0 + 4 * j < a.Lk) {
              const float4 b4 = __ldg(bp4 + (int64_t)j * a.Lq);
              ffma2(x[4 * j + 0], x[4 * j + 1], b4.x, b4.y, L2E, L2E, x[4 * j + 0], x[4 * j + 1]);
              ffma2(x[4 * j + 2], x[4 * j + 3], b4.z, b4.w, L2E, L2E, x[4 * j + 2], x[4 * j + 3]);
            }
          }
        } else if (BIAS && qrow < rows_per_pass) {
          const float* bp = p.bias + ((int64_t)head * a.Lq + qrow) * a.bias_ld + bias_k0 + key0;
          if (bias_vec && key0 + CH3 <= a.Lk) {
#pragma unroll
            for (int j = 0; j < CH3 / 4; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
              ffma2(x[4 * j + 0], x[4 * j + 1], b4.x, b4.y, L2E, L2E, x[4 * j + 0], x[4 * j + 1]);
              ffma2(x[4 * j + 2], x[4 * j + 3], b4.z, b4.w, L2E, L2E, x[4 * j + 2], x[4 * j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH3; ++j)
              if (key0 + j < a.Lk) x[j] = fmaf(__ldg(bp + j), L2E, x[j]);
          }
        }
        float mr[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
        for (int j = 4; j < CH3; ++j) mr[j & 3] = fmaxf(mr[j & 3], x[j]);
        cmax = fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3]));
      }
      // ---- reference max.  O accumulates in TMEM across the chunks of a pass, so the exps stay relative
      //      to a *stale* reference m_run and O is only rescaled when the row max has grown by more than
      //      2^8 (P <= 256 is harmless in 16 bits; sums and O are fp32).  The result is the exact softmax.
      if (k == 0) {
        m_run = (cmax == -INFINITY) ? 0.f : cmax;
        l_run = 0.f;
      } else {
        const bool grow = cmax > m_run + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          const float fac = grow ? fast_exp2(m_run - cmax) : 1.0f;
          if (grow) m_run = cmax;
          l_run *= fac;
          // every P V issued so far (chunks <= g-1) must have landed before O is touched
          mbar_wait_a(bar_wg + BAR_O, (uint32_t)(g - 1) & 1u);
          tc_fence_after_sync();
          uint32_t ro[32];
          tmem_ld_32x32b_x32(t_o, ro);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < DH; ++d) ro[d] = __float_as_uint(__uint_as_float(ro[d]) * fac);
          tmem_st_32x32b_x32(t_o, ro);
          tmem_st_wait();
          tc_fence_before_sync();
        }
      }
      const float nm = -m_run;
      const float es = plain ? sl2 : 1.0f;
      float sm[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[CH3 / 2];                         // P in 16 bits, packed one key group behind the exponentials
      auto pack_group = [&](int cc) {
        const int j = cc * 8;
        pk[cc * 4 + 0] = pack2_v<BF16>(x[j + 0], x[j + 1]);
        pk[cc * 4 + 1] = pack2_v<BF16>(x[j + 2], x[j + 3]);
        pk[cc * 4 + 2] = pack2_v<BF16>(x[j + 4], x[j + 5]);
        pk[cc * 4 + 3] = pack2_v<BF16>(x[j + 6], x[j + 7]);
      };
#pragma unroll
      for (int cc = 0; cc < CH3 / 8; ++cc) {
        const int j = cc * 8;
        ffma2(x[j], x[j + 1], x[j], x[j + 1], es, es, nm, nm);
        ffma2(x[j + 2], x[j + 3], x[j + 2], x[j + 3], es, es, nm, nm);
        ffma2(x[j + 4], x[j + 5], x[j + 4], x[j + 5], es, es, nm, nm);
        ffma2(x[j + 6], x[j + 7], x[j + 6], x[j + 7], es, es, nm, nm);
        if ((POLY >> cc) & 1) {
#pragma unroll
          for (int t = 0; t < 8; t += 2) poly_exp2_pair(x[j + t], x[j + t + 1]);
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t) x[j + t] = fast_exp2_v(x[j + t]);
        }
        if (cc >= 1) pack_group(cc - 1);
        fadd2(sm[0], sm[1], sm[0], sm[1], x[j], x[j + 1]);
        fadd2(sm[2], sm[3], sm[2], sm[3], x[j + 2], x[j + 3]);
        fadd2(sm[0], sm[1], sm[0], sm[1], x[j + 4], x[j + 5]);
        fadd2(sm[2], sm[3], sm[2], sm[3], x[j + 6], x[j + 7]);
      }
      pack_group(CH3 / 8 - 1);
      // the P buffer (and O) are free once P V of the previous chunk has landed
      if (g >= 1) mbar_wait_fast(bar_wg + BAR_O, (uint32_t)(g - 1) & 1u);
#pragma unroll
      for (int cc = 0; cc < CH3 / 8; ++cc)
        sts128(p_swz ^ ((uint32_t)cc << 4), pk[cc * 4 + 0], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
      l_run += (sm[0] + sm[1]) + (sm[2] + sm[3]);
      if (q_now) complete_q(P + 1);
      // ---- a new pass overwrites O: the finished one has to leave TMEM first ----------------------
      if (k == 0 && g > 0) {
        float ot[DH];
        load_o(g - 1, ot);
        finish_pass(false, ot);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive_a(bar_wg + BAR_P);
      if (++k == n_uch) {
        k = 0; l_fin = l_run; m_fin = m_run; ++P;
        if (++pj == passes_per_tile) { pj = 0; q_tile += T; }
      }
    }
    load_o(g - 1, o);
    if (pj == 0) q_tile -= T;                  // back to the tile of the last pass
    finish_pass(true, o);
    } else {
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = 0.f;  // nothing to do for this warpgroup: an empty partial
    }

    // ---- partials of a shared tile meet here ---------------------------------------------------
    if (G > 1) {
      if (sub > 0) {                            // park (m, l, o) in this warpgroup's P buffers (all its MMAs retired)
        const uint32_t base = smem_u32(sP) + (uint32_t)wtid * 128u;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (((uint32_t)ch ^ ((uint32_t)wtid & 7u)) << 4)),
                       "f"(o[ch * 4 + 0]), "f"(o[ch * 4 + 1]), "f"(o[ch * 4 + 2]), "f"(o[ch * 4 + 3]) : "memory");
        sML[wg][wtid] = make_float2(m_fin, l_fin);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + t0), "r"(G * 128) : "memory");
      if (sub == 0) {
        float m_all = m_fin, l_all = l_fin;
        if (split_keys) {
          for (int s = 1; s < G; ++s) m_all = fmaxf(m_all, sML[t0 + s * T][wtid].x);
          const float ms = (m_all == -INFINITY) ? 0.f : m_all;
          const float w0 = fast_exp2(m_fin - ms);
          l_all = l_fin * w0;
#pragma unroll
          for (int d = 0; d < DH; ++d) o[d] *= w0;
        }
        for (int s = 1; s < G; ++s) {
          const int ow = t0 + s * T;
          const float2 ml = sML[ow][wtid];
          const float ms = (m_all == -INFINITY) ? 0.f : m_all;
          const float ws = split_keys ? fast_exp2(ml.x - ms) : 1.0f;
          if (split_keys) l_all = fmaf(ml.y, ws, l_all);
          const uint32_t base = smem_u32(sP_all + ow * SP3_BYTES) + (uint32_t)wtid * 128u;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const float4 f = lds128f(base + (((uint32_t)ch ^ ((uint32_t)wtid & 7u)) << 4));
            o[ch * 4 + 0] = fmaf(f.x, ws, o[ch * 4 + 0]);
            o[ch * 4 + 1] = fmaf(f.y, ws, o[ch * 4 + 1]);
            o[ch * 4 + 2] = fmaf(f.z, ws, o[ch * 4 + 2]);
            o[ch * 4 + 3] = fmaf(f.w, ws, o[ch * 4 + 3]);
          }
        }
        if (split_keys) {
          const float inv = l_all > 0.f ? 1.0f / l_all : 0.f;
#pragma unroll
          for (int d = 0; d < DH; ++d) o[d] *= inv;
        }
        // rows of tile t0 (token addressing as in write_rows)
        const int qrow = t0 * TILE_Q + wtid;
        if (qrow < rows_per_pass) {
          const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, qrow);
          int64_t off;
          if (p.q_mean) off = (((int64_t)b * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.out_stride + head * DH;
          else off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.out_stride + head * DH;
          uint8_t* op = reinterpret_cast<uint8_t*>(p.out) + off * 2;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 u;
            u.x = pack2<BF16>(o[ch * 8 + 0], o[ch * 8 + 1]);
            u.y = pack2<BF16>(o[ch * 8 + 2], o[ch * 8 + 3]);
            u.z = pack2<BF16>(o[ch * 8 + 4], o[ch * 8 + 5]);
            u.w = pack2<BF16>(o[ch * 8 + 6], o[ch * 8 + 7]);
            *reinterpret_cast<uint4*>(op + ch * 16) = u;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

// ---- merge of key-segment partials (see cobevt_attn_merge in the header) -----------------------------
__global__ void attn_merge_kernel(const float* __restrict__ part, int S, int64_t seg_stride, uint16_t* __restrict__ out,
                                  CobevtTokenMap qmap, int nx, int ny, int B, int n_win, int heads, int Lq,
                                  int rows_per_pass, int passes, int q_mean, int out_stride, int bf16) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * n_win * heads * rows_per_pass;
  if (idx >= total) return;
  const int r = (int)(idx % rows_per_pass);
  int64_t rest = idx / rows_per_pass;
  const int head = (int)(rest % heads); rest /= heads;
  const int win = (int)(rest % n_win);
  const int b = (int)(rest / n_win);
  float acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.f;
  for (int pass = 0; pass < passes; ++pass) {
    const int t = pass * rows_per_pass + r;
    const float* base = part + ((((int64_t)b * n_win + win) * heads + head) * Lq + t) * 36;
    float M = -INFINITY;
    for (int s = 0; s < S; ++s) M = fmaxf(M, base[s * seg_stride + 32]);
    const float Ms = (M == -INFINITY) ? 0.f : M;
    float L = 0.f, o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
    for (int s = 0; s < S; ++s) {
      const float4* ps = reinterpret_cast<const float4*>(base + s * seg_stride);
      const float4 ml = ps[8];
      const float w = fast_exp2(ml.x - Ms);
      L = fmaf(ml.y, w, L);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float4 f = ps[ch];
        o[ch * 4 + 0] = fmaf(f.x, w, o[ch * 4 + 0]); o[ch * 4 + 1] = fmaf(f.y, w, o[ch * 4 + 1]);
        o[ch * 4 + 2] = fmaf(f.z, w, o[ch * 4 + 2]); o[ch * 4 + 3] = fmaf(f.w, w, o[ch * 4 + 3]);
      }
    }
    const float inv = L > 0.f ? 1.0f / (L * (float)passes) : 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] = fmaf(o[d], inv, acc[d]);
  }
  const Tok t = token_of(qmap, nx, ny, win, r);
  int64_t off;
  if (q_mean) off = (((int64_t)b * qmap.Hv + t.py) * qmap.Wv + t.px) * out_stride + head * DH;
  else off = ((((int64_t)b * qmap.views + t.n) * qmap.Hv + t.py) * qmap.Wv + t.px) * out_stride + head * DH;
  uint4* op = reinterpret_cast<uint4*>(out + off);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 u;
    u.x = pack_h2(acc[ch * 8 + 0], acc[ch * 8 + 1], bf16); u.y = pack_h2(acc[ch * 8 + 2], acc[ch * 8 + 3], bf16);
    u.z = pack_h2(acc[ch * 8 + 4], acc[ch * 8 + 5], bf16); u.w = pack_h2(acc[ch * 8 + 6], acc[ch * 8 + 7], bf16);
    op[ch] = u;
  }
}

}  // namespace


extern "C" int cobevt_window_attn_fwd(const CobevtAttnParams* p, void* stream) {
  if (!p || !p->q || !p->k || !p->v || (!p->out && !p->part)) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: null pointer");
  const CobevtTokenMap& qm = p->qmap;
  const CobevtTokenMap& km = p->kvmap;
  if (qm.w1 <= 0 || qm.w2 <= 0 || km.w1 <= 0 || km.w2 <= 0 || qm.H % qm.w1 || qm.W % qm.w2 || km.H % km.w1 || km.W % km.w2)
    return cobevt::fail(COBEVT_ERR_SHAPE, "attn: map extents not divisible by the window");
  if (qm.Hv != qm.H || qm.Wv != qm.W) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: padded query maps are not supported");
  if (km.Hv > km.H || km.Wv > km.W) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: valid extent exceeds padded extent");
  if ((p->k_pad == nullptr) != (p->v_pad == nullptr)) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: k_pad/v_pad must come together");
  AArgs a;
  a.p = *p;
  a.kv_vt = p->kv_views_total > 0 ? p->kv_views_total : km.views;
  a.kv_segs = p->kv_segments > 1 ? p->kv_segments : 1;
  if (a.kv_segs > 1 && !p->part) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: kv_segments without part");
  if (p->kv_view0 < 0 || p->kv_view0 + a.kv_segs * km.views > a.kv_vt) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: kv_view0 + kvmap.views exceeds kv_views_total");
  if (p->part) {      // key-segment mode: every query token is a plain row; bias rows span the keys of ALL K/V views
    a.p.q_mean = 0;
  }
  a.bias_ld = a.kv_vt * km.w1 * km.w2;
  a.q_dbuf = 0;
  a.kv_tma = 0;
  a.dbg = 0;
#ifdef COBEVT_DEBUG_KNOBS
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("COBEVT_ATTN_DEBUG"); dbg = e ? atoi(e) : 0; } a.dbg = dbg; }
#endif
  a.q_nx = qm.H / qm.w1; a.q_ny = qm.W / qm.w2;
  a.k_nx = km.H / km.w1; a.k_ny = km.W / km.w2;
  a.n_win = a.q_nx * a.q_ny;
  if (a.n_win != a.k_nx * a.k_ny)
    return cobevt::fail(COBEVT_ERR_SHAPE, "attn: %d query windows vs %d key windows", a.n_win, a.k_nx * a.k_ny);
  a.Lq = qm.views * qm.w1 * qm.w2;
  a.Lk = km.views * km.w1 * km.w2;
  a.n_chunks = cobevt::ceil_div(a.Lk, CHUNK);
  a.q_passes = a.p.q_mean ? qm.views : 1;
  const int rows = a.p.q_mean ? qm.w1 * qm.w2 : a.Lq;
  a.q_tiles = cobevt::ceil_div(rows, TILE_Q);
  a.C = p->heads * DH;
  if (a.p.q_stride == 0) a.p.q_stride = a.C;
  if (a.p.kv_stride == 0) a.p.kv_stride = a.C;
  if (a.p.out_stride == 0) a.p.out_stride = a.C;
  if (a.p.q_stride < a.C || a.p.kv_stride < a.C || a.p.out_stride < a.C || (a.p.q_stride | a.p.kv_stride | a.p.out_stride) % 8)
    return cobevt::fail(COBEVT_ERR_SHAPE, "attn: token strides must be >= heads*32 and multiples of 8");
  if (p->bias && p->q_mean) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: bias with q_mean");
  if (p->bias_layout != 0 && p->bias_layout != 1) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: bias_layout must be 0 or 1");
  const bool bias_kq = p->bias && p->bias_layout == 1;
  if (bias_kq && (((a.Lk | a.bias_ld | (km.w1 * km.w2)) & 3) != 0))
    return cobevt::fail(COBEVT_ERR_SHAPE, "attn: key-group-major bias needs key counts that are multiples of 4");
  if (p->part && !p->out) {}   // (out is ignored in key-segment mode)
  if (p->dtype != COBEVT_F16 && p->dtype != COBEVT_BF16) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: dtype");
  const bool bf = p->dtype == COBEVT_BF16;
  a.idesc_qk = make_idesc_f16(128, CHUNK, bf, false, false);
  a.idesc_pv = make_idesc_f16(128, DH, bf, false, true);
  const int64_t items = (int64_t)p->B * a.n_win * p->heads * a.q_tiles;
  if (items <= 0 || items > 0x7fffffff) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: bad grid");
  static int force_kernel = -1;     // COBEVT_ATTN_KERNEL=2 forces the two-warpgroup kernel (A/B timing)
  if (force_kernel < 0) { const char* e = getenv("COBEVT_ATTN_KERNEL"); force_kernel = e ? atoi(e) : 0; }
  {
    const int keys_pad = cobevt::ceil_div(a.Lk, CH3) * CH3;
    const int smem = v3_fixed_smem() + keys_pad * (128 + 4);
    if (p->part && !(smem <= 227 * 1024 && a.q_tiles >= 3))
      return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: key-segment mode needs a resident segment (<= 704 keys) and >= 3 query tiles");
    if (smem <= 227 * 1024 && (force_kernel != 2 || p->part)) {
      static int attr_v3_dev[cobevt::kMaxDevices][6][2] = {};        // per device, per kernel instance
      int* attr_v3 = nullptr;
      // share of the exponentials evaluated on the FMA pipes: COBEVT_ATTN_POLY = 0 none (default), 1 = 2/8, 2 = 3/8, 3 = 4/8
      static int poly_sel = -1;
      if (poly_sel < 0) { const char* e = getenv("COBEVT_ATTN_POLY"); poly_sel = e ? atoi(e) : V3_POLY_DEFAULT; if (poly_sel < 0 || poly_sel > 3) poly_sel = V3_POLY_DEFAULT; }
      typedef void (*KernT)(const AArgs);
      static const KernT kerns[2][6] = {
          {window_attn_v3_kernel<false, 0x00, false, false>, window_attn_v3_kernel<false, 0x44, false, false>,
           window_attn_v3_kernel<false, 0x92, false, false>, window_attn_v3_kernel<false, 0xAA, false, false>,
           window_attn_v3_kernel<false, 0x00, true, false>, window_attn_v3_kernel<false, 0x00, false, true>},
          {window_attn_v3_kernel<true, 0x00, false, false>, window_attn_v3_kernel<true, 0x44, false, false>,
           window_attn_v3_kernel<true, 0x92, false, false>, window_attn_v3_kernel<true, 0xAA, false, false>,
           window_attn_v3_kernel<true, 0x00, true, false>, window_attn_v3_kernel<true, 0x00, false, true>}};
      // COBEVT_ATTN_QLIN=1 selects the single-offset instance.  Off by default: it was 0.7 % faster when introduced, but
      // since the bias code moved into its own instance ptxas hoists all 64 MUFU.EX2 of a chunk above the pack / store
      // tail in THIS instance only (+10 % on the stage-0 attention, profiles/r02g_attn_bias_ab.log)
      static int qlin_on = -1;
      if (qlin_on < 0) { const char* e = getenv("COBEVT_ATTN_QLIN"); qlin_on = (e && e[0] == '1') ? 1 : 0; }
      const bool qlin = qlin_on && poly_sel == 0 && !p->bias && (p->q_mean || p->qmap.views == 1) && p->qmap.w2 > 0 && 32 % p->qmap.w2 == 0;
      const int ksel = p->bias ? 5 : (qlin ? 4 : poly_sel);
      KernT kern = kerns[bf][ksel];
      attr_v3 = attr_v3_dev[cobevt::current_device()][ksel];
      if (smem > attr_v3[bf]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cobevt::fail(COBEVT_ERR_CUDA, "attn: set smem attr: %s", cudaGetErrorString(e));
        attr_v3[bf] = smem;
      }
      a.q_dbuf = 1;
      a.idesc_qk = make_idesc_f16(128, CH3, bf, false, false);
      {
        static int no_tma = -1;
        if (no_tma < 0) { const char* e = getenv("COBEVT_ATTN_NOTMA"); no_tma = e ? atoi(e) : 0; }
        a.kv_tma = 0;
        const int sx = km.grid ? a.k_nx : 1, sy = km.grid ? a.k_ny : 1;      // traversal strides along H / W
        const int span_h = (km.w1 - 1) * sx + 1, span_w = (km.w2 - 1) * sy + 1;
        cobevt::EncodeTiledFn enc = cobevt::get_encode_tiled();
        if (!no_tma && enc && km.Hv == km.H && km.Wv == km.W && sx <= 8 && sy <= 8 && span_h <= 256 && span_w <= 256 &&
            km.views <= 256 && ((reinterpret_cast<uintptr_t>(p->k) | reinterpret_cast<uintptr_t>(p->v)) & 15) == 0) {
          const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
          const cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)km.Wv, (cuuint64_t)km.Hv, (cuuint64_t)p->B * a.kv_vt};
          const cuuint64_t strides[3] = {(cuuint64_t)a.p.kv_stride * 2, (cuuint64_t)km.Wv * a.p.kv_stride * 2,
                                         (cuuint64_t)km.Hv * km.Wv * a.p.kv_stride * 2};
          const cuuint32_t box[4] = {DH, (cuuint32_t)span_w, (cuuint32_t)span_h, (cuuint32_t)km.views};
          const cuuint32_t es[4] = {1, (cuuint32_t)sy, (cuuint32_t)sx, 1};
          const CUresult r1 = enc(&a.tmK, dt, 4, const_cast<void*>(p->k), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          const CUresult r2 = enc(&a.tmV, dt, 4, const_cast<void*>(p->v), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          a.kv_tma = (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS) ? 1 : 0;
        }
      }
      const int64_t ctas = (int64_t)p->B * a.n_win * p->heads * a.kv_segs;
      cobevt::launch_k(kern, (unsigned)ctas, V3_THREADS, smem, reinterpret_cast<cudaStream_t>(stream), a);
      return cobevt::check_launch("window_attn_v3_kernel");
    }
  }
  if (p->kv_view0 != 0 || a.kv_vt != km.views || a.kv_segs > 1)
    return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: a K/V view subset needs the resident four-warpgroup kernel (<= 704 keys)");
  if (bias_kq) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: key-group-major bias needs the resident four-warpgroup kernel (<= 704 keys per call)");
  if (a.Lk <= KV_CAP_KEYS) {
    const int keys_pad = cobevt::ceil_div(a.Lk, CH) * CH;
    a.q_dbuf = (v2_fixed_smem(1) + keys_pad * (128 + 4) <= 227 * 1024) ? 1 : 0;
    const int smem = v2_fixed_smem(a.q_dbuf) + keys_pad * (128 + 4);
    static int attr_v2_dev[cobevt::kMaxDevices][2] = {{0, 0}};
    int* attr_v2 = attr_v2_dev[cobevt::current_device()];
    auto kern = bf ? window_attn_v2_kernel<true> : window_attn_v2_kernel<false>;
    if (smem > attr_v2[bf]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return cobevt::fail(COBEVT_ERR_CUDA, "attn: set smem attr: %s", cudaGetErrorString(e));
      attr_v2[bf] = smem;
    }
    a.idesc_qk = make_idesc_f16(128, CH, bf, false, false);
    const int64_t ctas = (int64_t)p->B * a.n_win * p->heads;
    cobevt::launch_k(kern, (unsigned)ctas, V2_THREADS, smem, reinterpret_cast<cudaStream_t>(stream), a);
    return cobevt::check_launch("window_attn_v2_kernel");
  }
  static bool attr_set_dev[cobevt::kMaxDevices] = {false};
  bool& attr_set = attr_set_dev[cobevt::current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return cobevt::fail(COBEVT_ERR_CUDA, "attn: set smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cobevt::launch_k(window_attn_kernel, (unsigned)items, 128, SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), a);
  return cobevt::check_launch("window_attn_kernel");
}

extern "C" int cobevt_attn_merge(const float* part, int32_t segments, int64_t segment_stride, void* out,
                                 const CobevtTokenMap* qmap, int32_t B, int32_t heads, int32_t q_mean,
                                 int32_t out_stride, int32_t dtype, void* stream) {
  if (!part || !out || !qmap || segments <= 0 || B <= 0 || heads <= 0) return cobevt::fail(COBEVT_ERR_SHAPE, "attn_merge: bad args");
  if (qmap->w1 <= 0 || qmap->w2 <= 0 || qmap->H % qmap->w1 || qmap->W % qmap->w2 || qmap->Hv != qmap->H || qmap->Wv != qmap->W)
    return cobevt::fail(COBEVT_ERR_SHAPE, "attn_merge: bad query map");
  if (dtype != COBEVT_F16 && dtype != COBEVT_BF16) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn_merge: dtype");
  const int nx = qmap->H / qmap->w1, ny = qmap->W / qmap->w2, n_win = nx * ny;
  const int Lq = qmap->views * qmap->w1 * qmap->w2;
  const int passes = q_mean ? qmap->views : 1;
  const int rpp = q_mean ? qmap->w1 * qmap->w2 : Lq;
  if (out_stride == 0) out_stride = heads * DH;
  const int64_t total = (int64_t)B * n_win * heads * rpp;
  cobevt::launch_k(attn_merge_kernel, (unsigned)((total + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream), 
      part, segments, segment_stride, reinterpret_cast<uint16_t*>(out), *qmap, nx, ny, B, n_win, heads, Lq, rpp, passes, q_mean,
      out_stride, dtype == COBEVT_BF16 ? 1 : 0);
  return cobevt::check_launch("attn_merge_kernel");
}

extern "C" int64_t cobevt_attn_workspace_bytes(const CobevtAttnParams* p) {
  // key-segment mode scratch (`part`): kv_segments x B x windows x heads x Lq x 36 floats; 0 for a single resident pass
  if (!p || p->kv_segments <= 1) return 0;
  const CobevtTokenMap& qm = p->qmap;
  if (qm.w1 <= 0 || qm.w2 <= 0 || qm.H % qm.w1 || qm.W % qm.w2) return -1;
  const int64_t n_win = (int64_t)(qm.H / qm.w1) * (qm.W / qm.w2);
  const int64_t Lq = (int64_t)qm.views * qm.w1 * qm.w2;
  return (int64_t)p->kv_segments * p->B * n_win * p->heads * Lq * 36 * (int64_t)sizeof(float);
}

extern "C" int32_t cobevt_attn_resident_keys(void) {
  return ((227 * 1024 - v3_fixed_smem()) / (128 + 4)) / CH3 * CH3;
}
