// Windowed / dilated-grid multi-head attention (head_dim = 32) for sm_100a.
//
// One CTA (128 threads, one query row per thread) handles one
// (batch row, window, head, 128-row query tile):
//   S  = Q K^T          tcgen05.mma  M=128 N=128 K=32   (Q, K: K-major SWIZZLE_64B tiles)
//   P  = softmax(...)   TMEM -> registers (tcgen05.ld), online max/sum in fp32, exp2 on MUFU,
//                       P written as 16-bit K-major SWIZZLE_128B A-operand
//   O += P V            tcgen05.mma  M=128 N=32  K=128  (V: MN-major SWIZZLE_64B, i.e. the
//                       same row-per-key image as K -- no transpose)
// Keys/values stream through a double-buffered 128-key chunk; window ('(x w1)(y w2)') and
// grid ('(w1 x)(w2 y)') partitions are address arithmetic on the NHWC token maps, so none of
// the reference's rearrange copies exist.  Zero-padded key positions use a constant token
// (Linear(LayerNorm(0)) of the reference, survey quirk Q2).
//
// Reference: CrossWinAttention.forward fax_modules.py:198-248, partitions
// fax_modules.py:399-433; Attention.forward fax_modules.py:132-176;
// swap_fusion_modules.Attention.forward :87-128.
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
  int32_t n_win, q_tiles, Lq, Lk, n_chunks, q_passes, C;
  int32_t q_nx, q_ny, k_nx, k_ny;  // windows along h / w
  uint32_t idesc_qk, idesc_pv;
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

__device__ __forceinline__ uint32_t pack_h2(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(128, 2) window_attn_kernel(const AArgs a) {
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
      if (tid == 0) {
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
      if (tid == 0) {
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
constexpr int V2_THREADS = 320;   // 2 softmax warpgroups + 2 MMA-issuer warps
constexpr int SP2_BYTES = TILE_Q * CH * 2;             // 16 KB per P buffer
constexpr int V2_FIXED_SMEM = 1024 + 2 * SQ_BYTES + 2 * 2 * SP2_BYTES;   // + keys * (128 B K/V + 4 B mask)

__device__ __forceinline__ void wg_barrier(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}

__global__ void __launch_bounds__(V2_THREADS, 1) window_attn_v2_kernel(const AArgs a) {
  // roles: warps 0-3 = softmax group 0, warps 4-7 = softmax group 1, warp 8 / 9 lane 0 = their MMA issuers
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_ch = (a.Lk + CH - 1) / CH;
  const int keys_pad = n_ch * CH;
  uint8_t* sQ_all = smem;
  uint8_t* sP_all = sQ_all + 2 * SQ_BYTES;
  uint8_t* sK = sP_all + 4 * SP2_BYTES;
  uint8_t* sV = sK + keys_pad * 64;
  float* kmask = reinterpret_cast<float*>(sV + keys_pad * 64);
  __shared__ uint64_t bar_s[2][2], bar_o[2][2], bar_p[2][2], bar_q[2];
  __shared__ uint32_t tmem_base_s;

  const CobevtAttnParams& p = a.p;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool is_mma = warp >= 8;
  const int wg = is_mma ? warp - 8 : warp >> 2, wtid = tid & 127;
  int item = blockIdx.x;
  const int head = item % p.heads; item /= p.heads;
  const int win = item % a.n_win;
  const int b = item / a.n_win;
  const int bf16 = p.dtype;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { mbar_init(&bar_s[i][j], 1); mbar_init(&bar_o[i][j], 1); mbar_init(&bar_p[i][j], 128); }
    mbar_init(&bar_q[0], 128); mbar_init(&bar_q[1], 128);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  const uint8_t* qb = reinterpret_cast<const uint8_t*>(p.q);
  const uint8_t* kb_ = reinterpret_cast<const uint8_t*>(p.k);
  const uint8_t* vb_ = reinterpret_cast<const uint8_t*>(p.v);

  // ---- resident K / V: 4 lanes per token (64 contiguous bytes), all 256 threads ----------
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
  uint8_t* sQ = sQ_all + wg * SQ_BYTES;
  uint8_t* sP = sP_all + wg * 2 * SP2_BYTES;
  const int rows_per_pass = p.q_mean ? p.qmap.w1 * p.qmap.w2 : a.Lq;

  if (is_mma) {
    // =============================== MMA issuer of warpgroup `wg` ===========================
    if ((tid & 31) == 0) {
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
      auto issue_qk = [&](int c) {       // S(c) = Q K(c)^T  -> S buffer c&1
        const uint32_t kb = k_addr + c * (CH * 64 / 16);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks)
          mma(tmem_base + (c & 1) * CH, qk_lo | (q_addr + ks * 2), qk_hi, qk_lo | (kb + ks * 2), qk_hi, a.idesc_qk, ks ? 1u : 0u);
        umma_commit(&bar_s[wg][c & 1]);
      };
      auto issue_pv = [&](int c) {       // O(c) = P(c) V(c)  -> O buffer c&1
        const uint32_t pa = p_addr + (c & 1) * (SP2_BYTES / 16), va = v_addr + c * (CH * 64 / 16);
#pragma unroll
        for (int ks = 0; ks < CH / 16; ++ks)
          mma(tmem_base + 2 * CH + (c & 1) * DH, p_lo | (pa + ks * 2), p_hi, v_lo | (va + ks * 64), v_hi, a.idesc_pv, ks ? 1u : 0u);
        umma_commit(&bar_o[wg][c & 1]);
      };
      uint32_t ph_q = 0, ph_p[2] = {0, 0};
      for (int q_tile = wg; q_tile < a.q_tiles; q_tile += 2) {
        for (int pass = 0; pass < a.q_passes; ++pass) {
          mbar_wait(&bar_q[wg], ph_q);
          ph_q ^= 1;
          tc_fence_after_sync();
          issue_qk(0);
          if (n_ch > 1) issue_qk(1);
          for (int c = 0; c < n_ch; ++c) {
            mbar_wait(&bar_p[wg][c & 1], ph_p[c & 1]);     // P(c) written, S(c) consumed, O(c-2) folded
            ph_p[c & 1] ^= 1;
            tc_fence_after_sync();
            issue_pv(c);
            if (c + 2 < n_ch) issue_qk(c + 2);
          }
        }
      }
    }
  } else {
  // =============================== softmax warpgroup `wg` ==================================
  const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t ph_s[2] = {0, 0}, ph_o[2] = {0, 0};
  const float sl2 = p.scale * 1.4426950408889634f;
  const bool any_mask = p.key_mask != nullptr;
  const bool bias_vec = (a.Lk & 3) == 0;
  bool first_q = true;

  for (int q_tile = wg; q_tile < a.q_tiles; q_tile += 2) {
    const int qrow = q_tile * TILE_Q + wtid;
    const bool row_ok = qrow < rows_per_pass;
    float acc[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] = 0.f;

    for (int pass = 0; pass < a.q_passes; ++pass) {
      // ---- Q tile: 4 lanes per row.  Every MMA that read the previous Q / P / S / O has retired:
      //      this thread waited on the last s_ready and o_ready of the previous pass. -------------
      (void)first_q;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = wtid + 128 * i;
        const int r = idx >> 2, piece = idx & 3;
        const int rr = q_tile * TILE_Q + r;
        const uint8_t* src = qb;
        uint32_t nb = 0;
        if (rr < rows_per_pass) {
          const int t_idx = p.q_mean ? pass * rows_per_pass + rr : rr;
          const Tok t = token_of(p.qmap, a.q_nx, a.q_ny, win, t_idx);
          const int64_t off = ((((int64_t)b * p.qmap.views + t.n) * p.qmap.Hv + t.py) * p.qmap.Wv + t.px) * p.q_stride + head * DH;
          src = qb + off * 2 + piece * 16;
          nb = 16;
        }
        cp_async16(sQ + sw64_off(r, piece), src, nb);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&bar_q[wg]);

      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
      float o[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = 0.f;

      for (int c = 0; c < n_ch; ++c) {
        const int sb = c & 1;
        const int key0 = c * CH;
        const bool use_km = any_mask || (key0 + CH > a.Lk);
        mbar_wait(&bar_s[wg][sb], ph_s[sb]);
        ph_s[sb] ^= 1;
        tc_fence_after_sync();
        // ---- logits of this row: 64 registers ------------------------------------------
        float x[CH];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(t_row + sb * CH, r0);
          tmem_ld_32x32b_x32(t_row + sb * CH + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { x[j] = __uint_as_float(r0[j]); x[32 + j] = __uint_as_float(r1[j]); }
        }
        float mx;
        const bool plain = !use_km && !(p.bias != nullptr);
        if (plain) {
          float mr[4] = {x[0], x[1], x[2], x[3]};          // 4 independent chains (ILP)
#pragma unroll
          for (int j = 4; j < CH; ++j) mr[j & 3] = fmaxf(mr[j & 3], x[j]);
          mx = fmaxf(m_run, fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3])) * sl2);
        } else {
          if (use_km) {
            const float4* km4 = reinterpret_cast<const float4*>(kmask + key0);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 k4 = km4[j];
              x[4 * j + 0] = fmaf(x[4 * j + 0], sl2, k4.x);
              x[4 * j + 1] = fmaf(x[4 * j + 1], sl2, k4.y);
              x[4 * j + 2] = fmaf(x[4 * j + 2], sl2, k4.z);
              x[4 * j + 3] = fmaf(x[4 * j + 3], sl2, k4.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) x[j] *= sl2;
          }
          if (p.bias && row_ok) {
            const float* bp = p.bias + ((int64_t)head * a.Lq + qrow) * a.Lk + key0;
            if (bias_vec && key0 + CH <= a.Lk) {
#pragma unroll
              for (int j = 0; j < CH / 4; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
                x[4 * j + 0] = fmaf(b4.x, 1.4426950408889634f, x[4 * j + 0]);
                x[4 * j + 1] = fmaf(b4.y, 1.4426950408889634f, x[4 * j + 1]);
                x[4 * j + 2] = fmaf(b4.z, 1.4426950408889634f, x[4 * j + 2]);
                x[4 * j + 3] = fmaf(b4.w, 1.4426950408889634f, x[4 * j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (key0 + j < a.Lk) x[j] = fmaf(__ldg(bp + j), 1.4426950408889634f, x[j]);
            }
          }
          float mr[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
          for (int j = 4; j < CH; ++j) mr[j & 3] = fmaxf(mr[j & 3], x[j]);
          mx = fmaxf(m_run, fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3])));
        }
        const float m_safe = (mx == -INFINITY) ? 0.f : mx;
        const float alpha = fast_exp2(m_run - m_safe);
        float sm[4] = {0.f, 0.f, 0.f, 0.f};
        if (plain) {
#pragma unroll
          for (int j = 0; j < CH; ++j) { x[j] = fast_exp2(fmaf(x[j], sl2, -m_safe)); sm[j & 3] += x[j]; }
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j) { x[j] = fast_exp2(x[j] - m_safe); sm[j & 3] += x[j]; }
        }
        const float sum = (sm[0] + sm[1]) + (sm[2] + sm[3]);
        l_run = l_run * alpha + sum;
        m_run = mx;
        // P(c) -> smem buffer sb (its previous reader, PV(c-2), retired: this thread folded O(c-2))
        {
          uint8_t* dst = sP + sb * SP2_BYTES;
#pragma unroll
          for (int cc = 0; cc < CH / 8; ++cc) {
            uint4 u;
            u.x = pack_h2(x[cc * 8 + 0], x[cc * 8 + 1], bf16);
            u.y = pack_h2(x[cc * 8 + 2], x[cc * 8 + 3], bf16);
            u.z = pack_h2(x[cc * 8 + 4], x[cc * 8 + 5], bf16);
            u.w = pack_h2(x[cc * 8 + 6], x[cc * 8 + 7], bf16);
            *reinterpret_cast<uint4*>(dst + sw128_off(wtid, cc)) = u;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&bar_p[wg][sb]);
        // ---- fold the previous chunk's O into the running output -------------------------
        if (c > 0) {
          const int ob = (c - 1) & 1;
          mbar_wait(&bar_o[wg][ob], ph_o[ob]);
          ph_o[ob] ^= 1;
          tc_fence_after_sync();
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + 2 * CH + ob * DH, r);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < DH; ++d) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(r[d]));
          tc_fence_before_sync();
        }
        alpha_prev = alpha;
      }
      {   // last chunk
        const int ob = (n_ch - 1) & 1;
        mbar_wait(&bar_o[wg][ob], ph_o[ob]);
        ph_o[ob] ^= 1;
        tc_fence_after_sync();
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + 2 * CH + ob * DH, r);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(r[d]));
        tc_fence_before_sync();
      }
      const float inv = (l_run > 0.f) ? 1.0f / (l_run * (float)a.q_passes) : 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc[d] = fmaf(o[d], inv, acc[d]);
    }

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
  }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

}  // namespace

extern "C" int cobevt_window_attn_fwd(const CobevtAttnParams* p, void* stream) {
  if (!p || !p->q || !p->k || !p->v || !p->out) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: null pointer");
  const CobevtTokenMap& qm = p->qmap;
  const CobevtTokenMap& km = p->kvmap;
  if (qm.w1 <= 0 || qm.w2 <= 0 || km.w1 <= 0 || km.w2 <= 0 || qm.H % qm.w1 || qm.W % qm.w2 || km.H % km.w1 || km.W % km.w2)
    return cobevt::fail(COBEVT_ERR_SHAPE, "attn: map extents not divisible by the window");
  if (qm.Hv != qm.H || qm.Wv != qm.W) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: padded query maps are not supported");
  if (km.Hv > km.H || km.Wv > km.W) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: valid extent exceeds padded extent");
  if ((p->k_pad == nullptr) != (p->v_pad == nullptr)) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: k_pad/v_pad must come together");
  AArgs a;
  a.p = *p;
  a.q_nx = qm.H / qm.w1; a.q_ny = qm.W / qm.w2;
  a.k_nx = km.H / km.w1; a.k_ny = km.W / km.w2;
  a.n_win = a.q_nx * a.q_ny;
  if (a.n_win != a.k_nx * a.k_ny)
    return cobevt::fail(COBEVT_ERR_SHAPE, "attn: %d query windows vs %d key windows", a.n_win, a.k_nx * a.k_ny);
  a.Lq = qm.views * qm.w1 * qm.w2;
  a.Lk = km.views * km.w1 * km.w2;
  a.n_chunks = cobevt::ceil_div(a.Lk, CHUNK);
  a.q_passes = p->q_mean ? qm.views : 1;
  const int rows = p->q_mean ? qm.w1 * qm.w2 : a.Lq;
  a.q_tiles = cobevt::ceil_div(rows, TILE_Q);
  a.C = p->heads * DH;
  if (a.p.q_stride == 0) a.p.q_stride = a.C;
  if (a.p.kv_stride == 0) a.p.kv_stride = a.C;
  if (a.p.out_stride == 0) a.p.out_stride = a.C;
  if (a.p.q_stride < a.C || a.p.kv_stride < a.C || a.p.out_stride < a.C || (a.p.q_stride | a.p.kv_stride | a.p.out_stride) % 8)
    return cobevt::fail(COBEVT_ERR_SHAPE, "attn: token strides must be >= heads*32 and multiples of 8");
  if (p->bias && p->q_mean) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: bias with q_mean");
  if (p->dtype != COBEVT_F16 && p->dtype != COBEVT_BF16) return cobevt::fail(COBEVT_ERR_UNSUPPORTED, "attn: dtype");
  const bool bf = p->dtype == COBEVT_BF16;
  a.idesc_qk = make_idesc_f16(128, CHUNK, bf, false, false);
  a.idesc_pv = make_idesc_f16(128, DH, bf, false, true);
  const int64_t items = (int64_t)p->B * a.n_win * p->heads * a.q_tiles;
  if (items <= 0 || items > 0x7fffffff) return cobevt::fail(COBEVT_ERR_SHAPE, "attn: bad grid");
  if (a.Lk <= KV_CAP_KEYS) {
    const int keys_pad = cobevt::ceil_div(a.Lk, CH) * CH;
    const int smem = V2_FIXED_SMEM + keys_pad * (128 + 4);
    static int attr_v2 = 0;
    if (smem > attr_v2) {
      cudaError_t e = cudaFuncSetAttribute(window_attn_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return cobevt::fail(COBEVT_ERR_CUDA, "attn: set smem attr: %s", cudaGetErrorString(e));
      attr_v2 = smem;
    }
    a.idesc_qk = make_idesc_f16(128, CH, bf, false, false);
    const int64_t ctas = (int64_t)p->B * a.n_win * p->heads;
    window_attn_v2_kernel<<<(unsigned)ctas, V2_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    return cobevt::check_launch("window_attn_v2_kernel");
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return cobevt::fail(COBEVT_ERR_CUDA, "attn: set smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  window_attn_kernel<<<(unsigned)items, 128, SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return cobevt::check_launch("window_attn_kernel");
}
