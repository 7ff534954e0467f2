// Plain fp32 CUDA-core GEMM (reference-accurate fallback + small/odd shapes).
//   C[M,N] (+)= A[M,K] . op(B) (+ bias[N]),  op(B) = B[K,N] (transB=0) or B[N,K]^T (transB=1)
// 128x128x8 CTA tile, 256 threads, 8x8 register tile, register double-buffered global loads.
// The tcgen05 kernels in gemm_tc.cu are the fast path for the large contractions; this kernel is
// the exact-fp32 yardstick they are tested against and serves shapes they do not cover.
#include "common.cuh"

namespace regcn {

constexpr int BM = 128, BN = 128, BK = 8, TBG = 256;

template <bool TRANSB>
__global__ void __launch_bounds__(TBG) gemm_f32_kernel(
    const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
    int M, int N, int K, const float* __restrict__ bias, int accumulate, int k_per_split, float* __restrict__ ws) {
  pdl_grid_sync();
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);

  // global -> register staging: A tile 128 rows x 8 k = 256 float4 (one per thread)
  const int a_row = tid >> 1, a_k4 = (tid & 1) * 4;
  // B tile: transB: 128 n-rows x 8 k (same mapping); else 8 k-rows x 128 n = 256 float4
  const int b_row = TRANSB ? (tid >> 1) : (tid >> 5);
  const int b_c4 = TRANSB ? (tid & 1) * 4 : (tid & 31) * 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra, rb;
  auto load_tiles = [&](int k0) {
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    const int gm = m0 + a_row, gk = k0 + a_k4;
    if (gm < M && gk < kend) ra = ldg4(A + (size_t)gm * lda + gk);  // K % 4 == 0 keeps float4 in range
    if (TRANSB) {
      const int gn = n0 + b_row, gk2 = k0 + b_c4;
      if (gn < N && gk2 < kend) rb = ldg4(B + (size_t)gn * ldb + gk2);
    } else {
      const int gk2 = k0 + b_row, gn = n0 + b_c4;
      if (gk2 < kend && gn < N) {
        if (gn + 3 < N) rb = ldg4(B + (size_t)gk2 * ldb + gn);
        else {
          const float* p = B + (size_t)gk2 * ldb + gn;
          rb.x = __ldg(p);
          if (gn + 1 < N) rb.y = __ldg(p + 1);
          if (gn + 2 < N) rb.z = __ldg(p + 2);
        }
      }
    }
  };
  auto store_tiles = [&](int buf) {
    As[buf][a_k4 + 0][a_row] = ra.x; As[buf][a_k4 + 1][a_row] = ra.y;
    As[buf][a_k4 + 2][a_row] = ra.z; As[buf][a_k4 + 3][a_row] = ra.w;
    if (TRANSB) {
      Bs[buf][b_c4 + 0][b_row] = rb.x; Bs[buf][b_c4 + 1][b_row] = rb.y;
      Bs[buf][b_c4 + 2][b_row] = rb.z; Bs[buf][b_c4 + 3][b_row] = rb.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][b_row][b_c4]) = rb;
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool has_next = k0 + BK < kend;
    if (has_next) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (has_next) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  const bool split = gridDim.z > 1;
  float* Cout = split ? ws + (size_t)blockIdx.z * (size_t)M * (size_t)N : C;
  const int ldo = split ? N : ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = acc[i][j];
      if (!split) {
        if (bias) v += __ldg(bias + gn);
        if (accumulate) v += Cout[(size_t)gm * ldo + gn];
      }
      Cout[(size_t)gm * ldo + gn] = v;
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, float* __restrict__ C, int ldc,
                                     int M, int N, const float* __restrict__ bias, int accumulate) {
  pdl_grid_sync();
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)M * N;
  if (idx >= total) return;
  const int m = (int)(idx / N), n = (int)(idx - (size_t)m * N);
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += ws[(size_t)s * total + idx];
  if (bias) v += __ldg(bias + n);
  if (accumulate) v += C[(size_t)m * ldc + n];
  C[(size_t)m * ldc + n] = v;
}

size_t gemm_f32_workspace_bytes(int M, int N, int split_k) {
  return split_k > 1 ? (size_t)split_k * (size_t)M * (size_t)N * sizeof(float) : 0;
}

int gemm_f32(const float* A, int lda, const float* B, int ldb, int transB, float* C, int ldc, int M, int N, int K,
             const float* bias, int accumulate, int split_k, float* ws, size_t ws_bytes, cudaStream_t st) {
  if (!A || !B || !C) { set_last_error("gemm_f32: null pointer"); return REGCN_ERR_NULL; }
  if (M < 0 || N <= 0 || K <= 0 || (K & 3) || (lda & 3) || (ldb & 3) || lda < K || (transB ? ldb < K : ldb < N) || ldc < N) {
    set_last_error("gemm_f32: bad dims M=%d N=%d K=%d lda=%d ldb=%d ldc=%d transB=%d (need K,lda,ldb %% 4 == 0)", M, N, K, lda, ldb, ldc, transB);
    return REGCN_ERR_DIM;
  }
  if (((uintptr_t)A | (uintptr_t)B) & 15) { set_last_error("gemm_f32: A/B must be 16-byte aligned"); return REGCN_ERR_DIM; }
  if (M == 0) return REGCN_OK;
  if (split_k < 1) split_k = 1;
  int k_per = ((K + split_k - 1) / split_k + BK - 1) / BK * BK;
  split_k = (K + k_per - 1) / k_per;
  if (split_k > 1 && (!ws || ws_bytes < gemm_f32_workspace_bytes(M, N, split_k))) {
    set_last_error("gemm_f32: split-K workspace too small"); return REGCN_ERR_WORKSPACE;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, split_k);
  if (transB) launch_k(gemm_f32_kernel<true>, grid, TBG, 0, st, A, lda, B, ldb, C, ldc, M, N, K, bias, accumulate, k_per, ws);
  else launch_k(gemm_f32_kernel<false>, grid, TBG, 0, st, A, lda, B, ldb, C, ldc, M, N, K, bias, accumulate, k_per, ws);
  if (split_k > 1) {
    const size_t total = (size_t)M * N;
    launch_k(splitk_reduce_kernel, (unsigned)((total + 255) / 256), 256, 0, st, ws, split_k, C, ldc, M, N, bias, accumulate);
  }
  return check_launch("gemm_f32");
}

}  // namespace regcn
