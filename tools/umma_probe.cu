// Standalone probe: validates the UMMA shared-memory descriptor / instruction
// descriptor conventions used by the library kernels against a host reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
// Run on a B200; prints one line per case: name, max abs error, PASS/FAIL.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <functional>
#include "../cobevt_b200/csrc/sm100.cuh"

using namespace sm100;

struct ProbeArgs {
  uint32_t a_bytes, b_bytes;          // smem image sizes (multiples of 16)
  uint32_t a_layout, a_lbo, a_sbo;    // descriptor fields for A
  uint32_t b_layout, b_lbo, b_sbo;    // descriptor fields for B
  uint32_t idesc;
  uint32_t n;                         // N (columns of D), multiple of 16, <= 256
  uint32_t nk;                        // number of K=16 steps
  uint32_t a_koff[32];                // byte offset of A start per k-step
  uint32_t b_koff[32];
  uint32_t use_bulk;                  // 1: load images with cp.async.bulk
  uint32_t a_base_off[32];            // descriptor base_offset field per k-step (bits 49..51)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img,
             float* __restrict__ d_out, ProbeArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_load;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((args.a_bytes + 1023u) & ~1023u);
  const uint32_t tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_load, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (args.use_bulk) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_load, args.a_bytes + args.b_bytes);
      bulk_g2s(sa, a_img, args.a_bytes, &bar_load);
      bulk_g2s(sb, b_img, args.b_bytes, &bar_load);
    }
    mbar_wait(&bar_load, 0);
  } else {
    for (uint32_t i = tid * 16; i < args.a_bytes; i += 128 * 16)
      *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(a_img + i);
    for (uint32_t i = tid * 16; i < args.b_bytes; i += 128 * 16)
      *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(b_img + i);
    fence_proxy_async_smem();
    __syncthreads();
  }

  if (tid == 0) {
    tc_fence_after_sync();
    for (uint32_t ks = 0; ks < args.nk; ++ks) {
      uint64_t da = make_smem_desc(smem_u32(sa) + args.a_koff[ks], args.a_lbo, args.a_sbo, args.a_layout) | ((uint64_t)(args.a_base_off[ks] & 7u) << 49);
      uint64_t db = make_smem_desc(smem_u32(sb) + args.b_koff[ks], args.b_lbo, args.b_sbo, args.b_layout);
      umma_f16(tmem_base, da, db, args.idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after_sync();

  // each thread = one row of D
  const uint32_t row = tid;
  for (uint32_t c0 = 0; c0 < args.n; c0 += 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(tmem_base + ((warp * 32u) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d_out[row * args.n + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// ---------------------------------------------------------------------------
// micro-benchmarks that size the attention design: TMEM read rate, ex2 rate
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) ldtm_bench(float* out, int iters, long long* cyc) {
  __shared__ uint32_t tmem_base_s;
  const uint32_t warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tb = tmem_base_s + ((warp * 32u) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tb + c * 32 + (it & 3) * 128, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc += __uint_as_float(r[j]);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * 128 + threadIdx.x] = acc;
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

__global__ void __launch_bounds__(256, 1) ex2_bench(float* out, int iters, long long* cyc, int mode) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[j])); x[j] = y - 1.0f; }
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { uint32_t xi = __float_as_uint(x[j]), y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(xi)); x[j] = __uint_as_float((y & 0x3fff3fffu) ^ 0x80008000u); }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

static void run_micro() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  long long h;
  const int it = 2000;
  ldtm_bench<<<1, 128>>>(out, it, cyc); cudaDeviceSynchronize();
  ldtm_bench<<<1, 128>>>(out, it, cyc); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDTM 32x32b.x32: %lld cyc for %d x 4 loads/warp, 4 warps -> %.1f B/cyc/SM (%.1f cyc per 128x32 fp32 block)\n",
         h, it, (double)it * 4 * 128 * 32 * 4 / h, (double)h / (it * 4));
  for (int mode = 0; mode < 2; ++mode) {
    ex2_bench<<<1, 256>>>(out, it, cyc, mode); cudaDeviceSynchronize();
    ex2_bench<<<1, 256>>>(out, it, cyc, mode); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ex2 %s: %lld cyc for %d x 8 per thread, 256 threads -> %.2f instr/cyc/SM\n",
           mode ? "f16x2" : "f32", h, it, (double)it * 8 * 256 / h);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("micro error %s\n", cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

// ---------------------------------------------------------------------------
static uint16_t f2h(float f) { __half h = __float2half(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
static uint16_t f2bf(float f) { __nv_bfloat16 h = __float2bfloat16(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float bf2f(uint16_t u) { __nv_bfloat16 h; memcpy(&h, &u, 2); return __bfloat162float(h); }

static uint32_t sw128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }
static uint32_t sw64(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

using Place = std::function<uint32_t(uint32_t row, uint32_t k)>;  // byte offset of element

static int run_case(const char* name, uint32_t M, uint32_t N, uint32_t K, bool bf16,
                    Place placeA, uint32_t a_bytes, Place placeB, uint32_t b_bytes,
                    ProbeArgs args) {
  std::vector<float> A(M * K), B(N * K);
  srand(1234);
  for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.0f;
  for (auto& v : B) v = (float)((rand() % 13) - 6) / 4.0f;
  std::vector<uint8_t> ai(a_bytes, 0), bi(b_bytes, 0);
  for (uint32_t r = 0; r < M; ++r)
    for (uint32_t k = 0; k < K; ++k) {
      uint16_t u = bf16 ? f2bf(A[r * K + k]) : f2h(A[r * K + k]);
      memcpy(&ai[placeA(r, k)], &u, 2);
    }
  for (uint32_t n = 0; n < N; ++n)
    for (uint32_t k = 0; k < K; ++k) {
      uint16_t u = bf16 ? f2bf(B[n * K + k]) : f2h(B[n * K + k]);
      memcpy(&bi[placeB(n, k)], &u, 2);
    }
  std::vector<float> ref(M * N, 0.f);
  for (uint32_t r = 0; r < M; ++r)
    for (uint32_t n = 0; n < N; ++n) {
      float s = 0;
      for (uint32_t k = 0; k < K; ++k) {
        float a = bf16 ? bf2f(f2bf(A[r * K + k])) : h2f(f2h(A[r * K + k]));
        float b = bf16 ? bf2f(f2bf(B[n * K + k])) : h2f(f2h(B[n * K + k]));
        s += a * b;
      }
      ref[r * N + n] = s;
    }
  uint8_t *da, *db; float* dd;
  cudaMalloc(&da, a_bytes); cudaMalloc(&db, b_bytes); cudaMalloc(&dd, M * N * 4);
  cudaMemcpy(da, ai.data(), a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bi.data(), b_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xFF, M * N * 4);
  args.a_bytes = a_bytes; args.b_bytes = b_bytes; args.n = N; args.nk = K / 16;
  uint32_t smem = ((a_bytes + 1023u) & ~1023u) + b_bytes + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe_kernel<<<1, 128, smem>>>(da, db, dd, args);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-44s CUDA ERROR %s\n", name, cudaGetErrorString(e));
    exit(3);  // context is dead after a trap
  }
  std::vector<float> out(M * N);
  cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (uint32_t i = 0; i < M * N; ++i) {
    double d = fabs((double)out[i] - ref[i]);
    if (!(d == d)) d = 1e30;
    if (d > maxerr) maxerr = d;
  }
  printf("%-44s maxerr %.4g  %s\n", name, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
  cudaFree(da); cudaFree(db); cudaFree(dd);
  return maxerr < 1e-3;
}


// A = im2col view of an unswizzled 8-channel input patch read IN PLACE: core matrices overlap (LBO = 16 B),
// 8-row groups are `sh` patch rows apart.  B no-swizzle as in case 4.
static int run_patch_case(const char* name, uint32_t pw, uint32_t ph, uint32_t sh, uint32_t KW, uint32_t ky0) {
  const uint32_t M = 128, N = 64, K = KW * 8;
  std::vector<float> P(ph * pw * 8), B(N * K);
  srand(77);
  for (auto& v : P) v = (float)((rand() % 17) - 8) / 8.0f;
  for (auto& v : B) v = (float)((rand() % 13) - 6) / 4.0f;
  const uint32_t a_bytes = ph * pw * 16, b_bytes = N * K * 2;
  std::vector<uint8_t> ai(a_bytes, 0), bi(b_bytes, 0);
  for (uint32_t i = 0; i < ph * pw * 8; ++i) { uint16_t u = f2h(P[i]); memcpy(&ai[i * 2], &u, 2); }
  for (uint32_t n = 0; n < N; ++n)
    for (uint32_t k = 0; k < K; ++k) { uint16_t u = f2h(B[n * K + k]); memcpy(&bi[(k / 8) * (N * 16) + n * 16 + (k % 8) * 2], &u, 2); }
  std::vector<float> ref(M * N, 0.f);
  for (uint32_t m = 0; m < M; ++m)
    for (uint32_t n = 0; n < N; ++n) {
      float s = 0;
      const uint32_t y = m / 8, x = m % 8;
      for (uint32_t k = 0; k < K; ++k) s += h2f(f2h(P[((y * sh + ky0) * pw + x + k / 8) * 8 + k % 8])) * h2f(f2h(B[n * K + k]));
      ref[m * N + n] = s;
    }
  ProbeArgs a{};
  a.a_layout = LAYOUT_NONE; a.a_lbo = 16; a.a_sbo = sh * pw * 16;
  a.b_layout = LAYOUT_NONE; a.b_lbo = N * 16; a.b_sbo = 128;
  a.idesc = make_idesc_f16(M, N, false, false, false);
  for (uint32_t ks = 0; ks < K / 16; ++ks) { a.a_koff[ks] = ky0 * pw * 16 + ks * 32; a.b_koff[ks] = ks * 2 * N * 16; }
  uint8_t *da, *db; float* dd;
  cudaMalloc(&da, a_bytes); cudaMalloc(&db, b_bytes); cudaMalloc(&dd, M * N * 4);
  cudaMemcpy(da, ai.data(), a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bi.data(), b_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xFF, M * N * 4);
  a.a_bytes = a_bytes; a.b_bytes = b_bytes; a.n = N; a.nk = K / 16;
  uint32_t smem = ((a_bytes + 1023u) & ~1023u) + b_bytes + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe_kernel<<<1, 128, smem>>>(da, db, dd, a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s CUDA ERROR %s\n", name, cudaGetErrorString(e)); exit(3); }
  std::vector<float> out(M * N);
  cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad_rows = 0;
  for (uint32_t m = 0; m < M; ++m) {
    double em = 0;
    for (uint32_t n = 0; n < N; ++n) { double d = fabs((double)out[m * N + n] - ref[m * N + n]); if (!(d == d)) d = 1e30; if (d > em) em = d; }
    if (em > 1e-3) ++bad_rows;
    if (em > maxerr) maxerr = em;
  }
  printf("%-44s maxerr %.4g  bad rows %d  %s\n", name, maxerr, bad_rows, maxerr < 1e-3 ? "PASS" : "FAIL");
  cudaFree(da); cudaFree(db); cudaFree(dd);
  return maxerr < 1e-3;
}

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  int id = 0;
  auto want = [&](int i) { return only < 0 || only == i; };
  if (only == 100) {
    run_patch_case("patch pw=11 sh=2 ky=0", 11, 37, 2, 4, 0);
    run_patch_case("patch pw=11 sh=2 ky=3", 11, 37, 2, 4, 3);
    run_patch_case("patch pw=8 sh=1 ky=0 (cores 128B-aligned)", 8, 16, 1, 2, 0);
    run_patch_case("patch pw=16 sh=1 ky=1", 16, 18, 1, 4, 1);
    return 0;
  }
  if (only < 0) run_micro();

  // ---- case 0/1: SW128 K-major A and B, K = 128 (two 64-wide k-blocks) -------------
  for (int bulk = 0; bulk < 2; ++bulk, ++id) {
    if (!want(id)) continue;
    const uint32_t M = 128, N = 64, K = 128;
    ProbeArgs a{}; a.use_bulk = bulk;
    a.a_layout = LAYOUT_SW128; a.a_lbo = 16; a.a_sbo = 1024;
    a.b_layout = LAYOUT_SW128; a.b_lbo = 16; a.b_sbo = 1024;
    a.idesc = make_idesc_f16(M, N, false, false, false);
    for (uint32_t ks = 0; ks < K / 16; ++ks) {
      a.a_koff[ks] = (ks / 4) * (M * 128) + (ks % 4) * 32;
      a.b_koff[ks] = (ks / 4) * (N * 128) + (ks % 4) * 32;
    }
    auto pA = [=](uint32_t r, uint32_t k) { return (k / 64) * (M * 128) + sw128(r, (k % 64) / 8) + (k % 8) * 2; };
    auto pB = [=](uint32_t r, uint32_t k) { return (k / 64) * (N * 128) + sw128(r, (k % 64) / 8) + (k % 8) * 2; };
    run_case(bulk ? "sw128 K-major f16 (bulk copy)" : "sw128 K-major f16 (st.shared)", M, N, K, false, pA, M * K * 2, pB, N * K * 2, a);
  }
  // ---- case 2: same, bf16, N = 256 ---------------------------------------------------
  if (want(id)) {
    const uint32_t M = 128, N = 256, K = 64;
    ProbeArgs a{};
    a.a_layout = LAYOUT_SW128; a.a_lbo = 16; a.a_sbo = 1024;
    a.b_layout = LAYOUT_SW128; a.b_lbo = 16; a.b_sbo = 1024;
    a.idesc = make_idesc_f16(M, N, true, false, false);
    for (uint32_t ks = 0; ks < K / 16; ++ks) { a.a_koff[ks] = ks * 32; a.b_koff[ks] = ks * 32; }
    auto pA = [=](uint32_t r, uint32_t k) { return sw128(r, k / 8) + (k % 8) * 2; };
    run_case("sw128 K-major bf16 N=256", M, N, K, true, pA, M * K * 2, pA, N * K * 2, a);
  }
  ++id;
  // ---- case 3: SW64 K-major, K = 32 (QK^T shape) --------------------------------------
  if (want(id)) {
    const uint32_t M = 128, N = 128, K = 32;
    ProbeArgs a{};
    a.a_layout = LAYOUT_SW64; a.a_lbo = 16; a.a_sbo = 512;
    a.b_layout = LAYOUT_SW64; a.b_lbo = 16; a.b_sbo = 512;
    a.idesc = make_idesc_f16(M, N, false, false, false);
    for (uint32_t ks = 0; ks < K / 16; ++ks) { a.a_koff[ks] = ks * 32; a.b_koff[ks] = ks * 32; }
    auto pA = [=](uint32_t r, uint32_t k) { return sw64(r, k / 8) + (k % 8) * 2; };
    run_case("sw64 K-major f16 K=32", M, N, K, false, pA, M * K * 2, pA, N * K * 2, a);
  }
  ++id;
  // ---- case 4: no-swizzle K-major ------------------------------------------------------
  if (want(id)) {
    const uint32_t M = 128, N = 64, K = 32;
    ProbeArgs a{};
    a.a_layout = LAYOUT_NONE; a.a_lbo = M * 16; a.a_sbo = 128;
    a.b_layout = LAYOUT_NONE; a.b_lbo = N * 16; a.b_sbo = 128;
    a.idesc = make_idesc_f16(M, N, false, false, false);
    for (uint32_t ks = 0; ks < K / 16; ++ks) { a.a_koff[ks] = ks * 2 * M * 16; a.b_koff[ks] = ks * 2 * N * 16; }
    auto pA = [=](uint32_t r, uint32_t k) { return (k / 8) * (M * 16) + r * 16 + (k % 8) * 2; };
    auto pB = [=](uint32_t r, uint32_t k) { return (k / 8) * (N * 16) + r * 16 + (k % 8) * 2; };
    run_case("no-swizzle K-major f16 (LBO=k-stride)", M, N, K, false, pA, M * K * 2, pB, N * K * 2, a);
  }
  ++id;
  // ---- case 5: P.V : A sw128 K-major (K = 128 keys), B = V MN-major SW64 ----------------
  // V image: key j is a 64-byte row (32 dims), 16B chunks swizzled by SW64.
  for (int hyp = 0; hyp < 3; ++hyp, ++id) {
    if (!want(id)) continue;
    const uint32_t M = 128, N = 32, K = 128;
    ProbeArgs a{};
    a.a_layout = LAYOUT_SW128; a.a_lbo = 16; a.a_sbo = 1024;
    a.b_layout = LAYOUT_SW64;
    if (hyp == 0) { a.b_lbo = 512; a.b_sbo = 512; }
    if (hyp == 1) { a.b_lbo = 16;  a.b_sbo = 512; }
    if (hyp == 2) { a.b_lbo = 512; a.b_sbo = 64; }
    a.idesc = make_idesc_f16(M, N, false, false, true);
    for (uint32_t ks = 0; ks < K / 16; ++ks) {
      a.a_koff[ks] = (ks / 4) * (M * 128) + (ks % 4) * 32;
      a.b_koff[ks] = ks * 16 * 64;
    }
    auto pA = [=](uint32_t r, uint32_t k) { return (k / 64) * (M * 128) + sw128(r, (k % 64) / 8) + (k % 8) * 2; };
    // B element (n = dim, k = key): key row k, dim n
    auto pB = [=](uint32_t n, uint32_t k) { return sw64(k, n / 8) + (n % 8) * 2; };
    char nm[96]; snprintf(nm, 96, "PV: B=V MN-major SW64 hyp%d (lbo %u sbo %u)", hyp, a.b_lbo, a.b_sbo);
    run_case(nm, M, N, K, false, pA, M * K * 2, pB, K * N * 2, a);
  }
  // ---- case 8: P.V with V no-swizzle MN-major -------------------------------------------
  // canonical ((8,n),(8,k)) in elements: 8 dims contiguous (16B), k rows... image: key j row of 64B unswizzled
  for (int hyp = 0; hyp < 2; ++hyp, ++id) {
    if (!want(id)) continue;
    const uint32_t M = 128, N = 32, K = 64;
    ProbeArgs a{};
    a.a_layout = LAYOUT_SW128; a.a_lbo = 16; a.a_sbo = 1024;
    a.b_layout = LAYOUT_NONE;
    // INTERLEAVE MN-major: ((1,n),(8,k)):((X,SBO),(1,LBO)) in uint128 units:
    //   8 k-rows contiguous 16B apart; n-groups (8 elems) stride SBO; k-groups stride LBO.
    // image: [n8 group][key][8 dims] : offset = (n/8)*(K*16) + k*16 + (n%8)*2
    if (hyp == 0) { a.b_sbo = K * 16; a.b_lbo = 128; }
    if (hyp == 1) { a.b_lbo = K * 16; a.b_sbo = 128; }
    a.idesc = make_idesc_f16(M, N, false, false, true);
    for (uint32_t ks = 0; ks < K / 16; ++ks) { a.a_koff[ks] = ks * 32; a.b_koff[ks] = ks * 16 * 16; }
    auto pA = [=](uint32_t r, uint32_t k) { return sw128(r, k / 8) + (k % 8) * 2; };
    auto pB = [=](uint32_t n, uint32_t k) { return (n / 8) * (K * 16) + k * 16 + (n % 8) * 2; };
    char nm[96]; snprintf(nm, 96, "PV: B=V MN-major no-swizzle hyp%d", hyp);
    run_case(nm, M, N, K, false, pA, M * K * 2, pB, K * N * 2, a);
  }
  // ---- case 10: M=128, N=128, K=64*3 SW128 multi-block accumulate -----------------------
  if (want(id)) {
    const uint32_t M = 128, N = 128, K = 192;
    ProbeArgs a{}; a.use_bulk = 1;
    a.a_layout = LAYOUT_SW128; a.a_lbo = 16; a.a_sbo = 1024;
    a.b_layout = LAYOUT_SW128; a.b_lbo = 16; a.b_sbo = 1024;
    a.idesc = make_idesc_f16(M, N, false, false, false);
    for (uint32_t ks = 0; ks < K / 16; ++ks) {
      a.a_koff[ks] = (ks / 4) * (M * 128) + (ks % 4) * 32;
      a.b_koff[ks] = (ks / 4) * (N * 128) + (ks % 4) * 32;
    }
    auto pA = [=](uint32_t r, uint32_t k) { return (k / 64) * (M * 128) + sw128(r, (k % 64) / 8) + (k % 8) * 2; };
    run_case("sw128 K-major f16 N=128 K=192", M, N, K, false, pA, M * K * 2, pA, N * K * 2, a);
  }
  ++id;

  // ---- strip cases: A rows live in a contiguous strip, swizzled by ABSOLUTE smem address;
  //      the tile for "tap j" starts j rows (j*128 B) into the strip.  hyp 0: base_offset = j,
  //      hyp 1: base_offset = 0.
  for (int j = 1; j <= 3; j += 2)
    for (int hyp = 0; hyp < 2; ++hyp, ++id) {
      if (!want(id)) continue;
      const uint32_t M = 128, N = 64, K = 64;
      ProbeArgs a{};
      a.a_layout = LAYOUT_SW128; a.a_lbo = 16; a.a_sbo = 1024;
      a.b_layout = LAYOUT_SW128; a.b_lbo = 16; a.b_sbo = 1024;
      a.idesc = make_idesc_f16(M, N, false, false, false);
      for (uint32_t ks = 0; ks < K / 16; ++ks) {
        a.a_koff[ks] = j * 128 + ks * 32; a.b_koff[ks] = ks * 32;
        a.a_base_off[ks] = hyp == 0 ? (uint32_t)j : 0u;
      }
      auto pA = [=](uint32_t r, uint32_t k) { uint32_t row = r + j; return row * 128u + ((((k / 8) ^ (row & 7u))) << 4) + (k % 8) * 2; };
      auto pB = [=](uint32_t r, uint32_t k) { return sw128(r, k / 8) + (k % 8) * 2; };
      char nm[96]; snprintf(nm, 96, "strip: start +%d rows, base_offset=%s", j, hyp == 0 ? "j" : "0");
      run_case(nm, M, N, K, false, pA, (M + 8) * 128, pB, N * K * 2, a);
    }
  // ---- 2-D halo tile: 18 x 16-pixel pitch, output tile 16 rows x 8 px; tap (ky,kx) starts at
  //      (ky*16 + kx) rows; SBO = 16*128 = 2048.
  for (int tap = 0; tap < 3; ++tap)
    for (int hyp = 0; hyp < 2; ++hyp, ++id) {
      if (!want(id)) continue;
      const int ky = tap, kx = tap;   // (0,0) (1,1) (2,2)
      const uint32_t M = 128, N = 64, K = 64;
      ProbeArgs a{};
      a.a_layout = LAYOUT_SW128; a.a_lbo = 16; a.a_sbo = 2048;
      a.b_layout = LAYOUT_SW128; a.b_lbo = 16; a.b_sbo = 1024;
      a.idesc = make_idesc_f16(M, N, false, false, false);
      const uint32_t start_row = ky * 16 + kx;
      for (uint32_t ks = 0; ks < K / 16; ++ks) {
        a.a_koff[ks] = start_row * 128 + ks * 32; a.b_koff[ks] = ks * 32;
        a.a_base_off[ks] = hyp == 0 ? (start_row & 7u) : 0u;
      }
      // logical A row r = (gy = r / 8, gx = r % 8) -> halo row (ky + gy) * 16 + kx + gx
      auto pA = [=](uint32_t r, uint32_t k) { uint32_t row = (ky + r / 8) * 16 + kx + (r % 8); return row * 128u + ((((k / 8) ^ (row & 7u))) << 4) + (k % 8) * 2; };
      auto pB = [=](uint32_t r, uint32_t k) { return sw128(r, k / 8) + (k % 8) * 2; };
      char nm[96]; snprintf(nm, 96, "halo 18x16: tap (%d,%d), SBO 2048, base_offset=%s", ky, kx, hyp == 0 ? "row&7" : "0");
      run_case(nm, M, N, K, false, pA, 18 * 16 * 128, pB, N * K * 2, a);
    }
  // ---- halo with pitch 10 pixels (SBO = 1280, groups have different swizzle phases)
  for (int hyp = 0; hyp < 2; ++hyp, ++id) {
    if (!want(id)) continue;
    const int ky = 1, kx = 1;
    const uint32_t M = 128, N = 64, K = 64;
    ProbeArgs a{};
    a.a_layout = LAYOUT_SW128; a.a_lbo = 16; a.a_sbo = 1280;
    a.b_layout = LAYOUT_SW128; a.b_lbo = 16; a.b_sbo = 1024;
    a.idesc = make_idesc_f16(M, N, false, false, false);
    const uint32_t start_row = ky * 10 + kx;
    for (uint32_t ks = 0; ks < K / 16; ++ks) {
      a.a_koff[ks] = start_row * 128 + ks * 32; a.b_koff[ks] = ks * 32;
      a.a_base_off[ks] = hyp == 0 ? (start_row & 7u) : 0u;
    }
    auto pA = [=](uint32_t r, uint32_t k) { uint32_t row = (ky + r / 8) * 10 + kx + (r % 8); return row * 128u + ((((k / 8) ^ (row & 7u))) << 4) + (k % 8) * 2; };
    auto pB = [=](uint32_t r, uint32_t k) { return sw128(r, k / 8) + (k % 8) * 2; };
    char nm[96]; snprintf(nm, 96, "halo pitch 10 (SBO 1280), base_offset=%s", hyp == 0 ? "row&7" : "0");
    run_case(nm, M, N, K, false, pA, 18 * 10 * 128 + 1024, pB, N * K * 2, a);
  }
  return 0;
}
