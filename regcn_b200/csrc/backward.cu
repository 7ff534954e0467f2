// Training-side kernels: the backward of the evolution step, the train-mode ConvTransE/R tower
// (batch-statistics BatchNorm, dropout), the cross-entropy gradient and the clipped Adam update.
//   src/rrgcn.py:197-223 (get_loss), src/main.py:235-246 (backward, clip_grad_norm_(1.0), Adam step)
//   rgcn/layers.py:222-279 (UnionRGCNLayer), src/decoder.py:29-52,78-100 (ConvTransR/E, mode "train")
// Every reduction is a fixed-order tree (warp-segmented over a CSR row, slab partials + one finishing pass):
// no float atomics, so a training step is bit-reproducible.
#include "common.cuh"
#include "internal.h"
#include <cub/cub.cuh>

namespace regcn {

#define ROW_PROLOGUE(M_)                                                               \
  const int lane = threadIdx.x & 31;                                                   \
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);         \
  if (row >= (M_)) return;                                                             \
  const int nvec = d >> 2;

static inline unsigned rgrid(int M) { return (unsigned)(((size_t)M * 32 + 255) / 256); }
static inline unsigned egrid(size_t n, int per_thread = 1) {
  size_t t = (n + per_thread - 1) / per_thread;
  size_t b = (t + 255) / 256;
  return (unsigned)(b < 1 ? 1 : b);
}
static inline int chk_d(const char* who, int d) {
  if (d <= 0 || (d & 3) || d > 256) { set_last_error("%s: d=%d unsupported (need d%%4==0, d<=256)", who, d); return REGCN_ERR_UNSUPPORTED; }
  return REGCN_OK;
}

// counter-based dropout mask: element i of call `seed` is kept iff u(i) >= p
__device__ __forceinline__ uint32_t mix32(uint32_t seed, unsigned long long idx) {
  uint32_t x = (uint32_t)idx * 0x9E3779B1u ^ ((uint32_t)(idx >> 32) * 0x85EBCA77u) ^ seed;
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  x += seed * 0x27d4eb2fU;
  x ^= x >> 15; x *= 0x2c1b3c6dU; x ^= x >> 12; x *= 0x297a2d39U; x ^= x >> 15;
  return x;
}
__device__ __forceinline__ float drop_scale(uint32_t seed, unsigned long long idx, float p, float inv_keep) {
  const float u = (float)(mix32(seed, idx) >> 8) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}
// gradient mask recovered from a forward output z: mode 0 none, 1 z > 0 (relu then dropout), 2 z != 0 (dropout)
__device__ __forceinline__ float masked(float g, const float* Z, size_t i, int mode, float scale) {
  if (mode == 0) return g;
  const float z = Z[i];
  if (mode == 1) return z > 0.f ? g * scale : 0.f;
  return z != 0.f ? g * scale : 0.f;
}

// =====================================================================================================
// CSR gather-sum: out[row] (+)= row_w[row] * sum_{j in row} col_w[col_j] * (X[col_j] (+ X[col_j + col2_off]))
// One warp per row, 32 column ids at a time broadcast by shuffle, fixed summation order (CSR order).
//   * aggregate backward w.r.t. h (rgcn/layers.py:257-279): the snapshot graph holds every edge with its inverse, so
//     the in-neighbour multiset of u equals its out-neighbour multiset: dh[u] = sum_{w in N_in(u)} norm[w] dAgg[w]
//     -- the forward CSR with the weight moved to the gathered row;
//   * aggregate backward w.r.t. the relation table: rows = relation types, columns = edge destinations;
//   * relation mean-pool backward (src/rrgcn.py:161-166): rows = entities, columns = relations, col2_off = R;
//   * decoder gathers E[s], rel[r] backward (src/decoder.py:81-82): rows = table rows, columns = query ids.
// =====================================================================================================
// Long rows (hub entities of a Zipf-shaped snapshot, popular relations) are finished by the whole CTA: every warp first
// sums the leading kGatherHead columns of its own row; the tails of the CTA's 8 rows are then split 32 columns at a
// time over the 8 warps and folded through shared memory in warp order -- still a fixed summation order.
constexpr int kGatherHead = 64;
template <int RV>
__device__ __forceinline__ void gather_span(WarpRow<RV>& acc, const float* __restrict__ X, int ldx,
                                            const float* __restrict__ col_w, const int* __restrict__ col, int j0, int e,
                                            int stride, int nvec, int lane, int col2_off,
                                            const float* __restrict__ rho = nullptr, float gamma = 0.f,
                                            const int* __restrict__ partner = nullptr, int row = 0) {
  for (; j0 < e; j0 += stride) {
    const int n = min(32, e - j0);
    int c = 0;
    float w = 0.f;
    if (lane < n) {
      c = __ldg(col + j0 + lane);
      w = col_w ? __ldg(col_w + c) : 1.f;
      if (rho) {   // radius-difference edge weight exp(-gamma |rho_a - rho_b|), hyperbolic_layers.py:232-234
        const int other = partner ? __ldg(partner + j0 + lane) : row;
        w *= expf(-gamma * fabsf(__ldg(rho + c) - __ldg(rho + other)));
      }
    }
    // four gathered rows in flight per step: the loads of columns k..k+3 are issued before the first FMA (a row is
    // an 800-byte L2/HBM access, so one load per iteration leaves the warp waiting a full memory latency per edge);
    // the FMAs still run in CSR order, so the sum is bit-identical to the one-at-a-time loop
    for (int k = 0; k < n; k += 4) {
      float4 v[4][RV];
      float wk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = min(k + u, n - 1);
        const int ck = __shfl_sync(0xffffffffu, c, kk);
        wk[u] = __shfl_sync(0xffffffffu, w, kk);
        const float* xr = X + (size_t)ck * ldx;
#pragma unroll
        for (int i = 0; i < RV; ++i) {
          const int cc = lane + i * kWarp;
          v[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cc < nvec) {
            v[u][i] = *reinterpret_cast<const float4*>(xr + 4 * cc);
            if (col2_off) v[u][i] = f4_add(v[u][i], *reinterpret_cast<const float4*>(xr + (size_t)col2_off * ldx + 4 * cc));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k + u < n) {
#pragma unroll
          for (int i = 0; i < RV; ++i) acc.v[i] = f4_fma(wk[u], v[u][i], acc.v[i]);
        }
      }
    }
  }
}

template <int RV>
__global__ void __launch_bounds__(256) csr_gather_sum_kernel(const float* __restrict__ X, int ldx,
                                                             const float* __restrict__ col_w,
                                                             const float* __restrict__ row_w,
                                                             const int* __restrict__ rowptr, const int* __restrict__ col,
                                                             int nrows, int d, int col2_off, float* __restrict__ out,
                                                             int ldo, int accumulate, const float* __restrict__ rho,
                                                             float gamma, const int* __restrict__ partner) {
  pdl_grid_sync();
  __shared__ float4 part[8][32 * RV];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row0 = blockIdx.x * 8;
  const int row = row0 + wid;
  const int nvec = d >> 2;
  WarpRow<RV> acc;
  acc.zero();
  if (row < nrows) {
    const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    gather_span(acc, X, ldx, col_w, col, b, min(e, b + kGatherHead), 32, nvec, lane, col2_off, rho, gamma, partner, row);
  }
  for (int rr = 0; rr < 8; ++rr) {                        // block-uniform loop: the tails of this CTA's rows
    if (row0 + rr >= nrows) break;
    const int b2 = __ldg(rowptr + row0 + rr) + kGatherHead, e2 = __ldg(rowptr + row0 + rr + 1);
    if (e2 <= b2) continue;
    WarpRow<RV> p;
    p.zero();
    gather_span(p, X, ldx, col_w, col, b2 + 32 * wid, e2, 32 * 8, nvec, lane, col2_off, rho, gamma, partner, row0 + rr);
#pragma unroll
    for (int i = 0; i < RV; ++i) part[wid][lane + i * kWarp] = p.v[i];
    __syncthreads();
    if (wid == rr) {
#pragma unroll
      for (int w = 0; w < 8; ++w)
#pragma unroll
        for (int i = 0; i < RV; ++i) acc.v[i] = f4_add(acc.v[i], part[w][lane + i * kWarp]);
    }
    __syncthreads();
  }
  if (row >= nrows) return;
  if (row_w) acc.scale(__ldg(row_w + row));
  float* o = out + (size_t)row * ldo;
  if (accumulate) {
    WarpRow<RV> prev;
    prev.load_plain(o, nvec, lane);
    acc.zip(prev, [](float a, float p) { return a + p; });
  }
  acc.store(o, nvec, lane);
}

int csr_gather_sum(const float* X, int ldx, const float* col_w, const float* row_w, const int* rowptr, const int* col,
                   int nrows, int d, int col2_off, float* out, int ldo, int accumulate, cudaStream_t st, const float* rho,
                   float gamma, const int* partner) {
  if (!X || !rowptr || !col || !out) { set_last_error("csr_gather_sum: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk_d("csr_gather_sum", d)) return e;
  if ((ldx & 3) || (ldo & 3) || ldx < d || ldo < d) { set_last_error("csr_gather_sum: bad pitch"); return REGCN_ERR_DIM; }
  if (nrows <= 0) return REGCN_OK;
  if (d <= 128) launch_k(csr_gather_sum_kernel<1>, rgrid(nrows), 256, 0, st, X, ldx, col_w, row_w, rowptr, col, nrows, d, col2_off, out, ldo, accumulate, rho, gamma, partner);
  else launch_k(csr_gather_sum_kernel<2>, rgrid(nrows), 256, 0, st, X, ldx, col_w, row_w, rowptr, col, nrows, d, col2_off, out, ldo, accumulate, rho, gamma, partner);
  return check_launch("csr_gather_sum");
}

// =====================================================================================================
// group_by_key: stable counting order of n int32 keys in [0, nkeys): rowptr (nkeys+1), perm (n) = original positions
// grouped by key, vals_out[i] = vals[perm[i]] (optional).  The transposed indices of the backward gathers.
// =====================================================================================================
__global__ void iota_kernel(int* __restrict__ out, int n) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}
__global__ void group_rowptr_kernel(const int* __restrict__ sorted, int n, int nkeys, int* __restrict__ rowptr) {
  pdl_grid_sync();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nkeys) return;
  int lo = 0, hi = n;                                   // first position with key >= k
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sorted[mid] < k) lo = mid + 1; else hi = mid;
  }
  rowptr[k] = lo;
}
__global__ void gather_i32_kernel(const int* __restrict__ vals, const int* __restrict__ perm, int n, int* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = vals[perm[i]];
}
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
size_t group_by_key_workspace_bytes(int n) {
  size_t cb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cb, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, n > 0 ? n : 1);
  return align256(cb) + 2 * align256((size_t)(n > 0 ? n : 1) * sizeof(int)) + 256;
}
int group_by_key(const int* keys, int n, int nkeys, const int* vals, int* rowptr, int* perm, int* vals_out, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  if (!rowptr || (n > 0 && (!keys || !perm)) || (vals && !vals_out)) { set_last_error("group_by_key: null pointer"); return REGCN_ERR_NULL; }
  if (n < 0 || nkeys <= 0) { set_last_error("group_by_key: bad sizes"); return REGCN_ERR_DIM; }
  if (n == 0) {
    cudaMemsetAsync(rowptr, 0, (size_t)(nkeys + 1) * sizeof(int), st);
    return check_launch("group_by_key");
  }
  if (!ws || ws_bytes < group_by_key_workspace_bytes(n)) { set_last_error("group_by_key: workspace too small"); return REGCN_ERR_WORKSPACE; }
  size_t cb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cb, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, n);
  char* base = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  void* cubtmp = base;
  int* sorted = (int*)(base + align256(cb));
  int* iota = (int*)(base + align256(cb) + align256((size_t)n * sizeof(int)));
  launch_k(iota_kernel, egrid(n), 256, 0, st, iota, n);
  int end_bit = 1;
  while (end_bit < 31 && (1 << end_bit) < nkeys) ++end_bit;
  cub::DeviceRadixSort::SortPairs(cubtmp, cb, keys, sorted, (const int*)iota, perm, n, 0, end_bit, st);
  launch_k(group_rowptr_kernel, egrid(nkeys + 1), 256, 0, st, (const int*)sorted, n, nkeys, rowptr);
  if (vals) launch_k(gather_i32_kernel, egrid(n), 256, 0, st, vals, (const int*)perm, n, vals_out);
  return check_launch("group_by_key");
}

// row id of every CSR position (the key array for transposing a CSR with group_by_key) and 1/row-length
__global__ void expand_rowptr_kernel(const int* __restrict__ rowptr, int nrows, int nnz, int* __restrict__ rowid,
                                     float* __restrict__ inv_len) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz && rowid) {
    int lo = 0, hi = nrows;                             // last row with rowptr[row] <= i
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (rowptr[mid + 1] <= i) lo = mid + 1; else hi = mid;
    }
    rowid[i] = lo;
  }
  if (i < nrows && inv_len) {
    const int len = rowptr[i + 1] - rowptr[i];
    inv_len[i] = len > 0 ? 1.0f / (float)len : 0.f;
  }
}
int expand_rowptr(const int* rowptr, int nrows, int nnz, int* rowid, float* inv_len, cudaStream_t st) {
  if (!rowptr || (!rowid && !inv_len)) { set_last_error("expand_rowptr: null pointer"); return REGCN_ERR_NULL; }
  const int n = nnz > nrows ? nnz : nrows;
  if (n <= 0) return REGCN_OK;
  launch_k(expand_rowptr_kernel, egrid(n), 256, 0, st, rowptr, nrows, nnz, rowid, inv_len);
  return check_launch("expand_rowptr");
}

// =====================================================================================================
// Row-kernel backwards
// =====================================================================================================
// y = x / max(|x|, 1e-12)  ->  dx = (dy - y <y,dy>) / max(|x|, 1e-12).  On entry x holds the input row, on exit dx.
template <int RV>
__device__ __forceinline__ void row_normalize_bwd(WarpRow<RV>& x, const WarpRow<RV>& dy) {
  const float n = fmaxf(sqrtf(x.sumsq()), 1e-12f);
  const float inv = 1.0f / n;
  x.scale(inv);
  const float s = x.dot(dy);
  x.zip(dy, [=](float y, float g) { return (g - y * s) * inv; });
}

template <int RV>
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ dx, int M, int d) {
  pdl_grid_sync();
  ROW_PROLOGUE(M)
  WarpRow<RV> a, g;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  g.load_plain(dy + (size_t)row * d, nvec, lane);
  row_normalize_bwd(a, g);
  a.store(dx + (size_t)row * d, nvec, lane);
}
int normalize_bwd(const float* x, const float* dy, float* dx, int M, int d, cudaStream_t st) {
  if (!x || !dy || !dx) { set_last_error("normalize_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk_d("normalize_bwd", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(normalize_bwd_kernel<1>, rgrid(M), 256, 0, st, x, dy, dx, M, d);
  else launch_k(normalize_bwd_kernel<2>, rgrid(M), 256, 0, st, x, dy, dx, M, d);
  return check_launch("normalize_bwd");
}

// GRU gates backward (nn.GRUCell, src/rrgcn.py:168-174): recomputes r, z, n from the saved pre-activations.
//   r = s(ir+hr), z = s(iz+hz), n = tanh(in + r*hn), h' = (h - n) z + n, out = [normalize](h')
template <int RV>
__global__ void __launch_bounds__(256) gru_gate_bwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                           const float* __restrict__ hprev, const float* __restrict__ dout,
                                                           int M, int d, int normalize, float* __restrict__ dgi,
                                                           float* __restrict__ dgh, float* __restrict__ dhprev) {
  pdl_grid_sync();
  ROW_PROLOGUE(M)
  WarpRow<RV> r, z, n, hn, h, g;
  {
    WarpRow<RV> t;
    const float* gir = gi + (size_t)row * 3 * d;
    const float* ghr = gh + (size_t)row * 3 * d;
    r.load_plain(gir, nvec, lane); t.load_plain(ghr, nvec, lane);
    r.zip(t, [](float a, float b) { return sigmoidf_(b + a); });
    z.load_plain(gir + d, nvec, lane); t.load_plain(ghr + d, nvec, lane);
    z.zip(t, [](float a, float b) { return sigmoidf_(b + a); });
    n.load_plain(gir + 2 * d, nvec, lane); hn.load_plain(ghr + 2 * d, nvec, lane);
    t = hn;
    t.zip(r, [](float a, float rr) { return a * rr; });
    n.zip(t, [](float a, float b) { return tanhf(a + b); });
  }
  h.load_plain(hprev + (size_t)row * d, nvec, lane);
  g.load_plain(dout + (size_t)row * d, nvec, lane);
  if (normalize) {
    WarpRow<RV> hp = h;                                  // h' = (h - n) z + n
    hp.zip(n, [](float a, float b) { return a - b; });
    hp.zip(z, [](float a, float b) { return a * b; });
    hp.zip(n, [](float a, float b) { return a + b; });
    row_normalize_bwd(hp, g);
    g = hp;
  }
  // g = dL/dh'
  WarpRow<RV> dz = h;                                    // dz_pre = g (h - n) z (1 - z)
  dz.zip(n, [](float a, float b) { return a - b; });
  dz.zip(g, [](float a, float b) { return a * b; });
  dz.zip(z, [](float a, float zz) { return a * zz * (1.0f - zz); });
  WarpRow<RV> dn = g;                                    // dn_pre = g (1 - z) (1 - n^2)
  dn.zip(z, [](float a, float zz) { return a * (1.0f - zz); });
  dn.zip(n, [](float a, float nn) { return a * (1.0f - nn * nn); });
  WarpRow<RV> dr = dn;                                   // dr_pre = dn_pre hn r (1 - r)
  dr.zip(hn, [](float a, float b) { return a * b; });
  dr.zip(r, [](float a, float rr) { return a * rr * (1.0f - rr); });
  WarpRow<RV> dhn = dn;                                  // d gh_n = dn_pre r
  dhn.zip(r, [](float a, float rr) { return a * rr; });
  g.zip(z, [](float a, float zz) { return a * zz; });   // dhprev = g z
  float* o = dgi + (size_t)row * 3 * d;
  dr.store(o, nvec, lane); dz.store(o + d, nvec, lane); dn.store(o + 2 * d, nvec, lane);
  o = dgh + (size_t)row * 3 * d;
  dr.store(o, nvec, lane); dz.store(o + d, nvec, lane); dhn.store(o + 2 * d, nvec, lane);
  g.store(dhprev + (size_t)row * d, nvec, lane);
}
int gru_gate_bwd(const float* gi, const float* gh, const float* hprev, const float* dout, int M, int d, int normalize,
                 float* dgi, float* dgh, float* dhprev, cudaStream_t st) {
  if (!gi || !gh || !hprev || !dout || !dgi || !dgh || !dhprev) { set_last_error("gru_gate_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk_d("gru_gate_bwd", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(gru_gate_bwd_kernel<1>, rgrid(M), 256, 0, st, gi, gh, hprev, dout, M, d, normalize, dgi, dgh, dhprev);
  else launch_k(gru_gate_bwd_kernel<2>, rgrid(M), 256, 0, st, gi, gh, hprev, dout, M, d, normalize, dgi, dgh, dhprev);
  return check_launch("gru_gate_bwd");
}

// UnionRGCNLayer apply step backward (rgcn/layers.py:241-253): out = dropout(rrelu(P + where(indeg>0, L_loop, L_evolve))).
// The masks are recovered from `out` itself: with dropout, out == 0 marks a dropped element.
//   dP = dout * f,  dL = [indeg>0 ? dP : 0 | indeg>0 ? 0 : dP]   (N x 2d)
template <int RV>
__global__ void __launch_bounds__(256) union_combine_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout,
                                                                const int* __restrict__ indeg, int N, int d, float p,
                                                                float* __restrict__ dP, float* __restrict__ dL) {
  pdl_grid_sync();
  ROW_PROLOGUE(N)
  WarpRow<RV> o, g;
  g.load_plain(dout + (size_t)row * d, nvec, lane);
  if (out) {
    o.load_plain(out + (size_t)row * d, nvec, lane);
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    const float zero_f = p > 0.f ? 0.f : kRReluSlope;
    g.zip(o, [=](float gg, float oo) { return gg * (oo > 0.f ? inv_keep : (oo < 0.f ? kRReluSlope * inv_keep : zero_f)); });
  }
  g.store(dP + (size_t)row * d, nvec, lane);
  if (dL) {
    WarpRow<RV> zr;
    zr.zero();
    const bool act = __ldg(indeg + row) > 0;
    float* l = dL + (size_t)row * 2 * d;
    (act ? g : zr).store(l, nvec, lane);
    (act ? zr : g).store(l + d, nvec, lane);
  }
}
int union_combine_bwd(const float* out, const float* dout, const int* indeg, int N, int d, float p, float* dP, float* dL,
                      cudaStream_t st) {
  if (!dout || !dP || (dL && !indeg)) { set_last_error("union_combine_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk_d("union_combine_bwd", d)) return e;
  if (N <= 0) return REGCN_OK;
  if (d <= 128) launch_k(union_combine_bwd_kernel<1>, rgrid(N), 256, 0, st, out, dout, indeg, N, d, p, dP, dL);
  else launch_k(union_combine_bwd_kernel<2>, rgrid(N), 256, 0, st, out, dout, indeg, N, d, p, dP, dL);
  return check_launch("union_combine_bwd");
}

// Time gate backward (src/rrgcn.py:176-178): out = g c' + (1-g) h, g = s(G+b), c' = [normalize](cur)
//   dG = dout (c' - h) g (1-g);  dcur = [normalize_bwd](dout g);  dh = dout (1-g)   (the direct path only)
template <int RV>
__global__ void __launch_bounds__(256) time_gate_bwd_kernel(const float* __restrict__ G, const float* __restrict__ bias,
                                                            const float* __restrict__ cur, const float* __restrict__ h,
                                                            const float* __restrict__ dout, int N, int d, int normalize_cur,
                                                            float* __restrict__ dG, float* __restrict__ dcur,
                                                            float* __restrict__ dh) {
  pdl_grid_sync();
  ROW_PROLOGUE(N)
  WarpRow<RV> g, b, c, hh, go;
  g.load_plain(G + (size_t)row * d, nvec, lane);
  b.load(bias, nvec, lane);
  c.load_plain(cur + (size_t)row * d, nvec, lane);
  hh.load_plain(h + (size_t)row * d, nvec, lane);
  go.load_plain(dout + (size_t)row * d, nvec, lane);
  g.zip(b, [](float a, float bb) { return sigmoidf_(a + bb); });
  WarpRow<RV> cn = c;
  if (normalize_cur) row_l2normalize(cn);
  cn.zip(hh, [](float a, float bb) { return a - bb; });            // c' - h
  cn.zip(go, [](float a, float gg) { return a * gg; });
  cn.zip(g, [](float a, float s) { return a * s * (1.0f - s); });  // dG
  cn.store(dG + (size_t)row * d, nvec, lane);
  WarpRow<RV> dc = go;
  dc.zip(g, [](float a, float s) { return a * s; });               // dL/dc'
  if (normalize_cur) { row_normalize_bwd(c, dc); dc = c; }
  dc.store(dcur + (size_t)row * d, nvec, lane);
  go.zip(g, [](float a, float s) { return a * (1.0f - s); });
  go.store(dh + (size_t)row * d, nvec, lane);
}
int time_gate_bwd(const float* G, const float* bias, const float* cur, const float* h, const float* dout, int N, int d,
                  int normalize_cur, float* dG, float* dcur, float* dh, cudaStream_t st) {
  if (!G || !bias || !cur || !h || !dout || !dG || !dcur || !dh) { set_last_error("time_gate_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk_d("time_gate_bwd", d)) return e;
  if (N <= 0) return REGCN_OK;
  if (d <= 128) launch_k(time_gate_bwd_kernel<1>, rgrid(N), 256, 0, st, G, bias, cur, h, dout, N, d, normalize_cur, dG, dcur, dh);
  else launch_k(time_gate_bwd_kernel<2>, rgrid(N), 256, 0, st, G, bias, cur, h, dout, N, d, normalize_cur, dG, dcur, dh);
  return check_launch("time_gate_bwd");
}

// =====================================================================================================
// Elementwise: tanh backward, dropout
// =====================================================================================================
__global__ void __launch_bounds__(256) tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                       float* __restrict__ dx, size_t n4) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = reinterpret_cast<const float4*>(y)[i];
  const float4 g = reinterpret_cast<const float4*>(dy)[i];
  reinterpret_cast<float4*>(dx)[i] = make_float4(g.x * (1.f - a.x * a.x), g.y * (1.f - a.y * a.y),
                                                 g.z * (1.f - a.z * a.z), g.w * (1.f - a.w * a.w));
}
int tanh_bwd(const float* y, const float* dy, float* dx, size_t n, cudaStream_t st) {
  if (!y || !dy || !dx) { set_last_error("tanh_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (n & 3) { set_last_error("tanh_bwd: n must be a multiple of 4"); return REGCN_ERR_DIM; }
  if (n == 0) return REGCN_OK;
  launch_k(tanh_bwd_kernel, egrid(n / 4), 256, 0, st, y, dy, dx, n / 4);
  return check_launch("tanh_bwd");
}

__global__ void __launch_bounds__(256) dropout_kernel(float* __restrict__ x, size_t n, float p, float inv_keep, uint32_t seed) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) x[i] *= drop_scale(seed, i, p, inv_keep);
}
int dropout_inplace(float* x, size_t n, float p, uint32_t seed, cudaStream_t st) {
  if (!x) { set_last_error("dropout: null pointer"); return REGCN_ERR_NULL; }
  if (!(p >= 0.f && p < 1.f)) { set_last_error("dropout: p must be in [0,1)"); return REGCN_ERR_DIM; }
  if (n == 0 || p == 0.f) return REGCN_OK;
  launch_k(dropout_kernel, egrid(n), 256, 0, st, x, n, p, 1.0f / (1.0f - p), seed);
  return check_launch("dropout");
}

// =====================================================================================================
// Column reductions over a (rows x cols) row-major matrix in two fixed-order stages:
//   stage A: partial[slab][col] over rows of the slab (32 columns x 8 row lanes per CTA, smem tree);
//   stage B: per channel (group of L consecutive columns) sum over slabs and the L columns, in double.
// MODE 0: (x, x^2) -> BatchNorm batch statistics;  MODE 1: (dy, dy*xhat) -> BatchNorm backward sums, dy masked by
// the forward output;  MODE 2: (x) -> bias gradients / plain column sums.
// =====================================================================================================
constexpr int kSlabRows = 64;
template <int MODE>
__global__ void __launch_bounds__(256) col_reduce_kernel(const float* __restrict__ A, const float* __restrict__ Z,
                                                         const float* __restrict__ Y, int ld, int rows, int cols,
                                                         int mask_mode, float mask_scale, const float* __restrict__ mean,
                                                         const float* __restrict__ invstd, int L, int C,
                                                         float* __restrict__ p0, float* __restrict__ p1) {
  pdl_grid_sync();
  __shared__ float s0[8][33], s1[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * kSlabRows;
  const int r1 = min(rows, r0 + kSlabRows);
  float a = 0.f, b = 0.f;
  if (col < cols) {
    float mu = 0.f, is = 1.f;
    if (MODE == 1) { const int ch = (col / L) % C; mu = mean[ch]; is = invstd[ch]; }
    for (int r = r0 + ty; r < r1; r += 8) {
      const size_t i = (size_t)r * ld + col;
      if (MODE == 0) { const float v = A[i]; a += v; b += v * v; }
      else if (MODE == 1) { const float g = masked(A[i], Z, i, mask_mode, mask_scale); a += g; b += g * ((Y[i] - mu) * is); }
      else { a += A[i]; }
    }
  }
  s0[ty][tx] = a; s1[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && col < cols) {
    float x = 0.f, y = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { x += s0[k][tx]; y += s1[k][tx]; }
    p0[(size_t)blockIdx.y * cols + col] = x;
    if (MODE != 2) p1[(size_t)blockIdx.y * cols + col] = y;
  }
}

// stage B.  One CTA (128 threads) per channel.
//  mode 0: mean, invstd (biased variance, eps) + running-stat update (momentum, unbiased variance)   nn.BatchNorm1d
//  mode 1: o0 = sum0, o1 = sum1 (floats): dbeta, dgamma and the two sums of the BatchNorm backward
//  mode 2: o0 (+)= sum0
__global__ void __launch_bounds__(128) col_finalize_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                                           int nslab, int cols, int L, int mode, double count, float eps,
                                                           float momentum, float* __restrict__ o0, float* __restrict__ o1,
                                                           float* __restrict__ run_mean, float* __restrict__ run_var,
                                                           int accumulate) {
  pdl_grid_sync();
  __shared__ double r0[128], r1[128];
  const int ch = blockIdx.x;
  double a = 0.0, b = 0.0;
  const int per = nslab * L;
  for (int i = threadIdx.x; i < per; i += 128) {
    const int s = i / L, l = i - s * L;
    const size_t idx = (size_t)s * cols + (size_t)ch * L + l;
    a += (double)p0[idx];
    if (mode != 2) b += (double)p1[idx];
  }
  r0[threadIdx.x] = a; r1[threadIdx.x] = b;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a = r0[0]; b = r1[0];
    if (mode == 0) {
      const double mu = a / count;
      double var = b / count - mu * mu;
      if (var < 0.0) var = 0.0;
      o0[ch] = (float)mu;
      o1[ch] = (float)(1.0 / sqrt(var + (double)eps));
      if (run_mean) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        run_mean[ch] = (float)((1.0 - momentum) * (double)run_mean[ch] + momentum * mu);
        run_var[ch] = (float)((1.0 - momentum) * (double)run_var[ch] + momentum * unb);
      }
    } else if (mode == 1) {
      o0[ch] = (float)a; o1[ch] = (float)b;
    } else {
      o0[ch] = (accumulate ? o0[ch] : 0.f) + (float)a;
    }
  }
}

size_t col_reduce_workspace_bytes(int rows, int cols) {
  const size_t nslab = (size_t)(rows + kSlabRows - 1) / kSlabRows;
  return 2 * (nslab < 1 ? 1 : nslab) * (size_t)cols * sizeof(float);
}
static int col_reduce_launch(int mode, const float* A, const float* Z, const float* Y, int ld, int rows, int cols,
                             int mask_mode, float mask_scale, const float* mean, const float* invstd, int L, int C,
                             float eps, float momentum, float* o0, float* o1, float* run_mean, float* run_var,
                             int accumulate, float* ws, size_t ws_bytes, const char* who, cudaStream_t st) {
  if (!A || !o0 || !ws) { set_last_error("%s: null pointer", who); return REGCN_ERR_NULL; }
  if (rows <= 0 || cols <= 0 || L <= 0 || cols % L) { set_last_error("%s: bad dims rows=%d cols=%d L=%d", who, rows, cols, L); return REGCN_ERR_DIM; }
  if (ws_bytes < col_reduce_workspace_bytes(rows, cols)) { set_last_error("%s: workspace too small", who); return REGCN_ERR_WORKSPACE; }
  const int nslab = (rows + kSlabRows - 1) / kSlabRows;
  float* p0 = ws;
  float* p1 = ws + (size_t)nslab * cols;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)nslab);
  if (mode == 0) launch_k(col_reduce_kernel<0>, grid, 256, 0, st, A, Z, Y, ld, rows, cols, mask_mode, mask_scale, mean, invstd, L, C, p0, p1);
  else if (mode == 1) launch_k(col_reduce_kernel<1>, grid, 256, 0, st, A, Z, Y, ld, rows, cols, mask_mode, mask_scale, mean, invstd, L, C, p0, p1);
  else launch_k(col_reduce_kernel<2>, grid, 256, 0, st, A, Z, Y, ld, rows, cols, mask_mode, mask_scale, mean, invstd, L, C, p0, p1);
  launch_k(col_finalize_kernel, (unsigned)(cols / L), 128, 0, st, (const float*)p0, (const float*)p1, nslab, cols, L, mode,
           (double)rows * (double)L, eps, momentum, o0, o1, run_mean, run_var, accumulate);
  return check_launch(who);
}

// BatchNorm1d batch statistics of X viewed as (B, C, L): mean (C), invstd (C); running stats updated in place.
int bn_stats(const float* X, int B, int C, int L, float eps, float momentum, float* mean, float* invstd, float* run_mean,
             float* run_var, float* ws, size_t ws_bytes, cudaStream_t st) {
  if (!invstd) { set_last_error("bn_stats: null pointer"); return REGCN_ERR_NULL; }
  return col_reduce_launch(0, X, nullptr, nullptr, C * L, B, C * L, 0, 1.f, nullptr, nullptr, L, C, eps, momentum, mean,
                           invstd, run_mean, run_var, 0, ws, ws_bytes, "bn_stats", st);
}
// BatchNorm backward sums: sum_dy (C) = dbeta, sum_dy_xhat (C) = dgamma; dy = masked(dZ, Z).
int bn_bwd_stats(const float* dZ, const float* Z, const float* Y, int B, int C, int L, int mask_mode, float mask_scale,
                 const float* mean, const float* invstd, float* sum_dy, float* sum_dy_xhat, float* ws, size_t ws_bytes,
                 cudaStream_t st) {
  if (!Y || !mean || !invstd || !sum_dy_xhat || (mask_mode && !Z)) { set_last_error("bn_bwd_stats: null pointer"); return REGCN_ERR_NULL; }
  return col_reduce_launch(1, dZ, Z, Y, C * L, B, C * L, mask_mode, mask_scale, mean, invstd, L, C, 0.f, 0.f, sum_dy,
                           sum_dy_xhat, nullptr, nullptr, 0, ws, ws_bytes, "bn_bwd_stats", st);
}
int col_sum(const float* X, int ld, int rows, int cols, int L, float* out, int accumulate, float* ws, size_t ws_bytes,
            cudaStream_t st) {
  return col_reduce_launch(2, X, nullptr, nullptr, ld, rows, cols, 0, 1.f, nullptr, nullptr, L, cols / (L > 0 ? L : 1), 0.f, 0.f,
                           out, nullptr, nullptr, nullptr, accumulate, ws, ws_bytes, "col_sum", st);
}

// out = dropout(relu?(gamma (x - mean) invstd + beta)) over X viewed as (B, C, L); mean == NULL: no normalisation.
__global__ void __launch_bounds__(256) bn_act_drop_kernel(const float* __restrict__ X, size_t total, int C, int L,
                                                          const float* __restrict__ mean, const float* __restrict__ invstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          int relu, float p, float inv_keep, uint32_t seed,
                                                          float* __restrict__ out) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = X[i];
  if (mean) {
    const int ch = (int)((i / (size_t)L) % (size_t)C);
    v = (v - mean[ch]) * invstd[ch] * gamma[ch] + beta[ch];
  }
  if (relu) v = fmaxf(v, 0.f);
  if (p > 0.f) v *= drop_scale(seed, i, p, inv_keep);
  out[i] = v;
}
__global__ void __launch_bounds__(256) bn_act_drop_vec4_kernel(const float* __restrict__ X, size_t total4, int C, int L,
                                                               const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               int relu, float p, float inv_keep, uint32_t seed,
                                                               float* __restrict__ out) {
  pdl_grid_sync();
  const size_t i4 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const size_t i = i4 * 4;
  const float4 x = *reinterpret_cast<const float4*>(X + i);
  float v[4] = {x.x, x.y, x.z, x.w};
  if (mean) {
    const int ch = (int)((i / (size_t)L) % (size_t)C);
    const float mu = mean[ch], is = invstd[ch], ga = gamma[ch], be = beta[ch];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (v[k] - mu) * is * ga + be;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (relu) v[k] = fmaxf(v[k], 0.f);
    if (p > 0.f) v[k] *= drop_scale(seed, i + k, p, inv_keep);      // same per-element counter as the scalar kernel
  }
  *reinterpret_cast<float4*>(out + i) = make_float4(v[0], v[1], v[2], v[3]);
}
int bn_act_drop(const float* X, int B, int C, int L, const float* mean, const float* invstd, const float* gamma,
                const float* beta, int relu, float p, uint32_t seed, float* out, cudaStream_t st) {
  if (!X || !out || (mean && (!invstd || !gamma || !beta))) { set_last_error("bn_act_drop: null pointer"); return REGCN_ERR_NULL; }
  if (!(p >= 0.f && p < 1.f)) { set_last_error("bn_act_drop: p must be in [0,1)"); return REGCN_ERR_DIM; }
  const size_t total = (size_t)B * C * L;
  if (total == 0) return REGCN_OK;
  if (L % 4 == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    launch_k(bn_act_drop_vec4_kernel, egrid(total / 4), 256, 0, st, X, total / 4, C, L, mean, invstd, gamma, beta, relu, p,
             p > 0.f ? 1.0f / (1.0f - p) : 1.0f, seed, out);
    return check_launch("bn_act_drop");
  }
  launch_k(bn_act_drop_kernel, egrid(total), 256, 0, st, X, total, C, L, mean, invstd, gamma, beta, relu, p,
           p > 0.f ? 1.0f / (1.0f - p) : 1.0f, seed, out);
  return check_launch("bn_act_drop");
}

// dX = gamma invstd (dy - sum_dy/n - xhat sum_dy_xhat/n), dy = masked(dZ, Z); then the mask of the layer before
// (out_mode / out_src: the forward INPUT-side dropout, recovered from its output) is applied to dX.
// mean == NULL: no BatchNorm in between, dX = masked(dZ, Z) (* out mask).
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dZ, const float* __restrict__ Z,
                                                           const float* __restrict__ Y, size_t total, int C, int L,
                                                           int mask_mode, float mask_scale, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ sum_dy,
                                                           const float* __restrict__ sum_dy_xhat, float inv_n,
                                                           const float* __restrict__ out_src, int out_mode, float out_scale,
                                                           float* __restrict__ dX) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  float g = masked(dZ[i], Z, i, mask_mode, mask_scale);
  if (mean) {
    const int ch = (int)((i / (size_t)L) % (size_t)C);
    const float is = invstd[ch];
    const float xh = (Y[i] - mean[ch]) * is;
    g = gamma[ch] * is * (g - sum_dy[ch] * inv_n - xh * sum_dy_xhat[ch] * inv_n);
  }
  g = masked(g, out_src, i, out_mode, out_scale);
  dX[i] = g;
}
// Same arithmetic, four consecutive elements of one channel per thread (L % 4 == 0): 16-byte accesses and one 64-bit
// division per four elements -- the scalar kernel ran at 2.6 TB/s on the (B, 50, d) conv features.
__device__ __forceinline__ float mask1(float g, float z, int mode, float scale) {
  if (mode == 0) return g;
  if (mode == 1) return z > 0.f ? g * scale : 0.f;
  return z != 0.f ? g * scale : 0.f;
}
__global__ void __launch_bounds__(256) bn_bwd_apply_vec4_kernel(const float* __restrict__ dZ, const float* __restrict__ Z,
                                                                const float* __restrict__ Y, size_t total4, int C, int L,
                                                                int mask_mode, float mask_scale, const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ sum_dy,
                                                                const float* __restrict__ sum_dy_xhat, float inv_n,
                                                                const float* __restrict__ out_src, int out_mode, float out_scale,
                                                                float* __restrict__ dX) {
  pdl_grid_sync();
  const size_t i4 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const size_t i = i4 * 4;
  const float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 dz = *reinterpret_cast<const float4*>(dZ + i);
  const float4 zz = mask_mode ? *reinterpret_cast<const float4*>(Z + i) : z0;
  float g[4] = {mask1(dz.x, zz.x, mask_mode, mask_scale), mask1(dz.y, zz.y, mask_mode, mask_scale),
                mask1(dz.z, zz.z, mask_mode, mask_scale), mask1(dz.w, zz.w, mask_mode, mask_scale)};
  if (mean) {
    const int ch = (int)((i / (size_t)L) % (size_t)C);
    const float is = invstd[ch], mu = mean[ch], ga = gamma[ch], s0 = sum_dy[ch], s1 = sum_dy_xhat[ch];
    const float4 yy = *reinterpret_cast<const float4*>(Y + i);
    const float y[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (y[k] - mu) * is;
      g[k] = ga * is * (g[k] - s0 * inv_n - xh * s1 * inv_n);
    }
  }
  const float4 oo = out_mode ? *reinterpret_cast<const float4*>(out_src + i) : z0;
  *reinterpret_cast<float4*>(dX + i) = make_float4(mask1(g[0], oo.x, out_mode, out_scale), mask1(g[1], oo.y, out_mode, out_scale),
                                                   mask1(g[2], oo.z, out_mode, out_scale), mask1(g[3], oo.w, out_mode, out_scale));
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
int bn_bwd_apply(const float* dZ, const float* Z, const float* Y, int B, int C, int L, int mask_mode, float mask_scale,
                 const float* mean, const float* invstd, const float* gamma, const float* sum_dy,
                 const float* sum_dy_xhat, const float* out_src, int out_mode, float out_scale, float* dX, cudaStream_t st) {
  if (!dZ || !dX || (mask_mode && !Z) || (out_mode && !out_src) || (mean && (!Y || !invstd || !gamma || !sum_dy || !sum_dy_xhat))) {
    set_last_error("bn_bwd_apply: null pointer"); return REGCN_ERR_NULL;
  }
  const size_t total = (size_t)B * C * L;
  if (total == 0) return REGCN_OK;
  if (L % 4 == 0 && aligned16(dZ) && aligned16(dX) && (!mask_mode || aligned16(Z)) && (!mean || aligned16(Y)) &&
      (!out_mode || aligned16(out_src))) {
    launch_k(bn_bwd_apply_vec4_kernel, egrid(total / 4), 256, 0, st, dZ, Z, Y, total / 4, C, L, mask_mode, mask_scale, mean,
             invstd, gamma, sum_dy, sum_dy_xhat, 1.0f / ((float)B * (float)L), out_src, out_mode, out_scale, dX);
    return check_launch("bn_bwd_apply");
  }
  launch_k(bn_bwd_apply_kernel, egrid(total), 256, 0, st, dZ, Z, Y, total, C, L, mask_mode, mask_scale, mean, invstd, gamma,
           sum_dy, sum_dy_xhat, 1.0f / ((float)B * (float)L), out_src, out_mode, out_scale, dX);
  return check_launch("bn_bwd_apply");
}

// =====================================================================================================
// ConvTransE / ConvTransR tower, train mode (src/decoder.py:29-52, 78-100)
// =====================================================================================================
// X0[b,0,:] = first[triples[b,col0]], X0[b,1,:] = second[triples[b,col1]]                  (:80-83)
template <int RV>
__global__ void __launch_bounds__(256) dec_gather_stack_kernel(const float* __restrict__ first, const float* __restrict__ second,
                                                               const int64_t* __restrict__ triples, int col0, int col1,
                                                               int B, int d, float* __restrict__ X0) {
  pdl_grid_sync();
  ROW_PROLOGUE(2 * B)
  const int b = row >> 1, ch = row & 1;
  const int64_t id = triples[(size_t)b * 3 + (ch ? col1 : col0)];
  WarpRow<RV> r;
  r.load_plain((ch ? second : first) + (size_t)id * d, nvec, lane);
  r.store(X0 + (size_t)row * d, nvec, lane);
}
int dec_gather_stack(const float* first, const float* second, const int64_t* triples, int col0, int col1, int B, int d,
                     float* X0, cudaStream_t st) {
  if (!first || !second || !triples || !X0) { set_last_error("dec_gather_stack: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk_d("dec_gather_stack", d)) return e;
  if (B <= 0) return REGCN_OK;
  if (d <= 128) launch_k(dec_gather_stack_kernel<1>, rgrid(2 * B), 256, 0, st, first, second, triples, col0, col1, B, d, X0);
  else launch_k(dec_gather_stack_kernel<2>, rgrid(2 * B), 256, 0, st, first, second, triples, col0, col1, B, d, X0);
  return check_launch("dec_gather_stack");
}

constexpr int kConvK = 3;      // Conv1d(2, C, 3, padding=1): the only kernel size the reference instantiates (:19,:67)
constexpr int kConvMaxC = 64;
constexpr int kConvMaxD = 256;

// x1 = dropout(bn0(X0)) (stored: the conv weight gradient and the dropout mask need it); y = conv1(x1) + bias   (:84-87)
__global__ void __launch_bounds__(256) dec_conv_fwd_kernel(const float* __restrict__ X0, int B, int d, int C,
                                                           const float* __restrict__ mean0, const float* __restrict__ invstd0,
                                                           const float* __restrict__ g0, const float* __restrict__ b0,
                                                           float p, float inv_keep, uint32_t seed,
                                                           const float* __restrict__ W, const float* __restrict__ bias,
                                                           float* __restrict__ X1, float* __restrict__ Y) {
  pdl_grid_sync();
  __shared__ float xs[2][kConvMaxD + 2];
  __shared__ float ws[kConvMaxC * 2 * kConvK + kConvMaxC];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < C * 2 * kConvK + C; i += blockDim.x) ws[i] = i < C * 2 * kConvK ? W[i] : bias[i - C * 2 * kConvK];
  for (int i = threadIdx.x; i < 2 * (d + 2); i += blockDim.x) {
    const int ch = i / (d + 2), j = i - ch * (d + 2) - 1;
    float v = 0.f;
    if (j >= 0 && j < d) {
      const size_t idx = ((size_t)b * 2 + ch) * d + j;
      v = (X0[idx] - mean0[ch]) * invstd0[ch] * g0[ch] + b0[ch];
      if (p > 0.f) v *= drop_scale(seed, idx, p, inv_keep);
      X1[idx] = v;
    }
    xs[ch][j + 1] = v;
  }
  __syncthreads();
  const float* wb = ws + C * 2 * kConvK;
  for (int o = threadIdx.x; o < C * d; o += blockDim.x) {
    const int c = o / d, j = o - c * d;
    const float* w = ws + c * 2 * kConvK;
    float acc = wb[c];
#pragma unroll
    for (int k = 0; k < kConvK; ++k) acc += w[k] * xs[0][j + k] + w[kConvK + k] * xs[1][j + k];
    Y[((size_t)b * C + c) * d + j] = acc;
  }
}
int dec_conv_fwd(const float* X0, int B, int d, int C, int ksz, const float* mean0, const float* invstd0, const float* g0,
                 const float* b0, float p, uint32_t seed, const float* W, const float* bias, float* X1, float* Y,
                 cudaStream_t st) {
  if (!X0 || !mean0 || !invstd0 || !g0 || !b0 || !W || !bias || !X1 || !Y) { set_last_error("dec_conv_fwd: null pointer"); return REGCN_ERR_NULL; }
  if (ksz != kConvK || C <= 0 || C > kConvMaxC || d <= 0 || d > kConvMaxD) { set_last_error("dec_conv_fwd: unsupported C=%d d=%d ksz=%d", C, d, ksz); return REGCN_ERR_UNSUPPORTED; }
  if (B <= 0) return REGCN_OK;
  launch_k(dec_conv_fwd_kernel, (unsigned)B, 256, 0, st, X0, B, d, C, mean0, invstd0, g0, b0, p, p > 0.f ? 1.0f / (1.0f - p) : 1.0f, seed, W, bias, X1, Y);
  return check_launch("dec_conv_fwd");
}

// dX1[b,ci,j] = sum_{c,k} dY[b,c,j+1-k] W[c,ci,k], then the input-dropout mask (x1 != 0) / (1-p)
__global__ void __launch_bounds__(256) dec_conv_bwd_input_kernel(const float* __restrict__ dY, const float* __restrict__ X1,
                                                                 int B, int d, int C, const float* __restrict__ W,
                                                                 float p, float inv_keep, float* __restrict__ dX1) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  float* dys = sm;                                 // C x (d+2), zero padded
  float* ws = sm + (size_t)C * (d + 2);            // C x 2 x 3
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < C * 2 * kConvK; i += blockDim.x) ws[i] = W[i];
  for (int i = threadIdx.x; i < C * (d + 2); i += blockDim.x) {
    const int c = i / (d + 2), j = i - c * (d + 2) - 1;
    dys[i] = (j >= 0 && j < d) ? dY[((size_t)b * C + c) * d + j] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 2 * d; o += blockDim.x) {
    const int ci = o / d, j = o - ci * d;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* w = ws + (c * 2 + ci) * kConvK;
      const float* g = dys + c * (d + 2) + j;      // g[m] = dY[c, j - 1 + m]; y[j'] uses x[j' + k - 1] -> j' = j + 1 - k -> m = 2 - k
#pragma unroll
      for (int k = 0; k < kConvK; ++k) acc += g[2 - k] * w[k];
    }
    const size_t idx = ((size_t)b * 2 + ci) * d + j;
    if (p > 0.f) acc = X1[idx] != 0.f ? acc * inv_keep : 0.f;
    dX1[idx] = acc;
  }
}
int dec_conv_bwd_input(const float* dY, const float* X1, int B, int d, int C, int ksz, const float* W, float p, float* dX1,
                       cudaStream_t st) {
  if (!dY || !X1 || !W || !dX1) { set_last_error("dec_conv_bwd_input: null pointer"); return REGCN_ERR_NULL; }
  if (ksz != kConvK || C <= 0 || C > kConvMaxC || d <= 0 || d > kConvMaxD) { set_last_error("dec_conv_bwd_input: unsupported shape"); return REGCN_ERR_UNSUPPORTED; }
  if (B <= 0) return REGCN_OK;
  const size_t smem = ((size_t)C * (d + 2) + (size_t)C * 2 * kConvK) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(dec_conv_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
  launch_k(dec_conv_bwd_input_kernel, (unsigned)B, 256, smem, st, dY, X1, B, d, C, W, p, p > 0.f ? 1.0f / (1.0f - p) : 1.0f, dX1);
  return check_launch("dec_conv_bwd_input");
}

// dW[c,ci,k] = sum_{b,j} dY[b,c,j] x1[b,ci,j+k-1];  dbias[c] = sum_{b,j} dY[b,c,j].
// CTA = (slab of kConvSlab queries) x (group of kConvCG channels); thread j owns position j and keeps the 7 partial sums
// of each channel of the group in registers; block tree at the end -> partial[slab][C*7].
constexpr int kConvCG = 10;
constexpr int kConvSlab = 16;
__global__ void __launch_bounds__(256) dec_conv_bwd_weight_kernel(const float* __restrict__ dY, const float* __restrict__ X1,
                                                                  int B, int d, int C, float* __restrict__ partial) {
  pdl_grid_sync();
  __shared__ float red[8][kConvCG * 7];
  const int c0 = blockIdx.y * kConvCG;
  const int b0 = blockIdx.x * kConvSlab, b1 = min(B, b0 + kConvSlab);
  const int j = threadIdx.x;
  float acc[kConvCG][7];
#pragma unroll
  for (int c = 0; c < kConvCG; ++c)
#pragma unroll
    for (int q = 0; q < 7; ++q) acc[c][q] = 0.f;
  if (j < d) {
    // no staging, no barriers in the loop: the three taps of both input rows come straight from L1 (neighbouring
    // threads share them), the dY loads are coalesced over j, and consecutive queries overlap freely
#pragma unroll 2
    for (int b = b0; b < b1; ++b) {
      const float* xr = X1 + (size_t)b * 2 * d;
      float x[6];
      x[0] = j > 0 ? xr[j - 1] : 0.f;         x[1] = xr[j];         x[2] = j + 1 < d ? xr[j + 1] : 0.f;
      x[3] = j > 0 ? xr[d + j - 1] : 0.f;     x[4] = xr[d + j];     x[5] = j + 1 < d ? xr[d + j + 1] : 0.f;
      const float* gy = dY + ((size_t)b * C + c0) * d + j;
#pragma unroll
      for (int c = 0; c < kConvCG; ++c) {
        if (c0 + c < C) {
          const float g = gy[(size_t)c * d];
#pragma unroll
          for (int q = 0; q < 6; ++q) acc[c][q] = fmaf(g, x[q], acc[c][q]);
          acc[c][6] += g;
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < kConvCG; ++c)
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const float s = warp_sum(acc[c][q]);
      if (lane == 0) red[wid][c * 7 + q] = s;
    }
  __syncthreads();
  if (threadIdx.x < kConvCG * 7) {
    const int c = threadIdx.x / 7, q = threadIdx.x - c * 7;
    if (c0 + c < C) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
      // layout of one partial row: dW (C x 2 x 3) then dbias (C)
      const size_t col = q < 6 ? (size_t)(c0 + c) * 6 + q : (size_t)C * 6 + (c0 + c);
      partial[(size_t)blockIdx.x * (C * 7) + col] = s;
    }
  }
}
size_t dec_conv_bwd_weight_workspace_bytes(int B, int C) {
  const size_t nslab = (size_t)(B + kConvSlab - 1) / kConvSlab;
  return (nslab < 1 ? 1 : nslab) * (size_t)C * 7 * sizeof(float) + col_reduce_workspace_bytes((int)nslab, C * 7);
}
int dec_conv_bwd_weight(const float* dY, const float* X1, int B, int d, int C, int ksz, float* dW, float* ws,
                        size_t ws_bytes, cudaStream_t st) {
  if (!dY || !X1 || !dW || !ws) { set_last_error("dec_conv_bwd_weight: null pointer"); return REGCN_ERR_NULL; }
  if (ksz != kConvK || C <= 0 || C > kConvMaxC || d <= 0 || d > kConvMaxD) { set_last_error("dec_conv_bwd_weight: unsupported shape"); return REGCN_ERR_UNSUPPORTED; }
  if (ws_bytes < dec_conv_bwd_weight_workspace_bytes(B, C)) { set_last_error("dec_conv_bwd_weight: workspace too small"); return REGCN_ERR_WORKSPACE; }
  if (B <= 0) return REGCN_OK;
  const int nslab = (B + kConvSlab - 1) / kConvSlab;
  float* partial = ws;
  float* ws2 = ws + (size_t)nslab * C * 7;
  dim3 grid((unsigned)nslab, (unsigned)((C + kConvCG - 1) / kConvCG));
  launch_k(dec_conv_bwd_weight_kernel, grid, 256, 0, st, dY, X1, B, d, C, partial);
  // dW (C*6 floats) followed by dbias (C floats) in one output buffer of C*7 floats
  return col_sum(partial, C * 7, nslab, C * 7, 1, dW, 0, ws2, ws_bytes - (size_t)nslab * C * 7 * sizeof(float), st);
}

// =====================================================================================================
// Cross entropy over materialised logits (nn.CrossEntropyLoss, src/rrgcn.py:87-88,218-223)
// =====================================================================================================
__device__ __forceinline__ float block_max(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
  __syncthreads();
  return r;
}
// ce[b] = lse[b] - S[b, target]; one CTA per row
__global__ void __launch_bounds__(256) ce_lse_rows_kernel(const float* __restrict__ S, int64_t ld, int B, int N,
                                                          const int64_t* __restrict__ triples, int target_col,
                                                          float* __restrict__ ce, float* __restrict__ lse) {
  pdl_grid_sync();
  __shared__ float sh[8];
  const int b = blockIdx.x;
  const float* s = S + (size_t)b * ld;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) m = fmaxf(m, s[j]);
  m = block_max(m, sh);
  float a = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) a += expf(s[j] - m);
  a = block_sum(a, sh);
  if (threadIdx.x == 0) {
    const float l = m + logf(a);
    lse[b] = l;
    ce[b] = l - s[triples[(size_t)b * 3 + target_col]];
  }
}
// in place: S[b,j] = (exp(S[b,j] - lse[b]) - [j == target]) * (*gscale) / B; padding columns [N, ld) are zeroed so the
// matrix can be the K-major operand of the two gradient GEMMs
__global__ void __launch_bounds__(256) softmax_grad_rows_kernel(float* __restrict__ S, int64_t ld, int B, int N,
                                                                const int64_t* __restrict__ triples, int target_col,
                                                                const float* __restrict__ lse,
                                                                const float* __restrict__ gscale, float inv_B) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  float* s = S + (size_t)b * ld;
  const float l = lse[b];
  const float sc = (gscale ? *gscale : 1.0f) * inv_B;
  const int t = (int)triples[(size_t)b * 3 + target_col];
  for (int j = threadIdx.x; j < (int)ld; j += blockDim.x) {
    float v = 0.f;
    if (j < N) v = (expf(s[j] - l) - (j == t ? 1.f : 0.f)) * sc;
    s[j] = v;
  }
}
int ce_lse_rows(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, float* ce, float* lse,
                float* loss, cudaStream_t st) {
  if (!S || !triples || !ce || !lse) { set_last_error("ce_lse_rows: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0 || N <= 0) return REGCN_OK;
  launch_k(ce_lse_rows_kernel, (unsigned)B, 256, 0, st, S, ld, B, N, triples, target_col, ce, lse);
  if (loss) return mean_f32(ce, B, loss, st);
  return check_launch("ce_lse_rows");
}
int softmax_grad_rows(float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, const float* lse,
                      const float* gscale, cudaStream_t st) {
  if (!S || !triples || !lse) { set_last_error("softmax_grad_rows: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0 || N <= 0) return REGCN_OK;
  launch_k(softmax_grad_rows_kernel, (unsigned)B, 256, 0, st, S, ld, B, N, triples, target_col, lse, gscale, 1.0f / (float)B);
  return check_launch("softmax_grad_rows");
}

// =====================================================================================================
// Transpose (+ TF32 split): out[c][r] = X[r][c]; columns [rows, ldo) of the output are zero filled so that the result
// is a K-major GEMM operand with K = ldo.  Outputs: plain (optional) and hi/lo (optional).
// =====================================================================================================
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ X, int rows, int cols, int ldx,
                                                              float* __restrict__ out, float* __restrict__ hi,
                                                              float* __restrict__ lo, int ldo) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? X[(size_t)r * ldx + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;                  // output row c, output column r
    if (c < cols && r < ldo) {
      const float v = tile[tx][i];
      const size_t o = (size_t)c * ldo + r;
      if (out) out[o] = v;
      if (hi) { float h, l; split_tf32_1(v, h, l); hi[o] = h; lo[o] = l; }
    }
  }
}
int transpose_split(const float* X, int rows, int cols, int ldx, float* out, float* hi, float* lo, int ldo, cudaStream_t st) {
  if (!X || (!out && !hi) || (hi && !lo)) { set_last_error("transpose_split: null pointer"); return REGCN_ERR_NULL; }
  if (ldo < rows || ldx < cols) { set_last_error("transpose_split: bad pitch"); return REGCN_ERR_DIM; }
  if (rows <= 0 || cols <= 0) return REGCN_OK;
  dim3 grid((unsigned)((ldo + 31) / 32), (unsigned)((cols + 31) / 32));
  launch_k(transpose_split_kernel, grid, 256, 0, st, X, rows, cols, ldx, out, hi, lo, ldo);
  return check_launch("transpose_split");
}

// =====================================================================================================
// Optimizer: clip_grad_norm_(max_norm) + torch.optim.Adam(lr, weight_decay) over ONE flat parameter buffer
// (src/main.py:194, 243-246).  Stage 1: fixed-order sum of squares (double); stage 2: the update.
// =====================================================================================================
constexpr int kNormBlocks = 296;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, size_t n, double* __restrict__ partial) {
  pdl_grid_sync();
  __shared__ double sh[256];
  double a = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)g[i];
    a += v * v;
  }
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void sumsq_finalize_kernel(const double* __restrict__ partial, int np, float* __restrict__ total_norm) {
  pdl_grid_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < np; ++i) a += partial[i];
    *total_norm = (float)sqrt(a);
  }
}
// total_norm (device float): the clip coefficient is min(1, max_norm / (total_norm + 1e-6)) like clip_grad_norm_
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                                                   float wd, float bc1, float bc2_sqrt, float max_norm,
                                                   const float* __restrict__ total_norm) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float coef = 1.0f;
  if (total_norm && max_norm > 0.f) coef = fminf(1.0f, max_norm / (*total_norm + 1e-6f));
  const float pi = p[i];
  float gi = g[i] * coef;
  gi = fmaf(wd, pi, gi);
  const float mi = b1 * m[i] + (1.0f - b1) * gi;          // exp_avg.lerp_(grad, 1-b1)
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}
size_t adam_workspace_bytes(void) { return (size_t)kNormBlocks * sizeof(double) + 16; }
int grad_norm(const float* g, size_t n, float* total_norm, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!g || !total_norm || !ws) { set_last_error("grad_norm: null pointer"); return REGCN_ERR_NULL; }
  if (ws_bytes < adam_workspace_bytes() || ((uintptr_t)ws & 7)) { set_last_error("grad_norm: workspace too small / unaligned"); return REGCN_ERR_WORKSPACE; }
  launch_k(sumsq_partial_kernel, kNormBlocks, 256, 0, st, g, n, (double*)ws);
  launch_k(sumsq_finalize_kernel, 1, 32, 0, st, (const double*)ws, (int)kNormBlocks, total_norm);
  return check_launch("grad_norm");
}
int adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps, float wd,
              int step, float max_norm, const float* total_norm, cudaStream_t st) {
  if (!p || !g || !m || !v) { set_last_error("adam_step: null pointer"); return REGCN_ERR_NULL; }
  if (step < 1) { set_last_error("adam_step: step counts from 1"); return REGCN_ERR_DIM; }
  if (n == 0) return REGCN_OK;
  const float bc1 = (float)(1.0 - pow((double)b1, (double)step));
  const float bc2s = (float)sqrt(1.0 - pow((double)b2, (double)step));
  launch_k(adam_kernel, egrid(n), 256, 0, st, p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2s, max_norm, total_norm);
  return check_launch("adam_step");
}

// =====================================================================================================
// Static-graph constraint: RGCNBlockLayer backward (rgcn/layers.py:147-179 through
// RGCNLayer.forward :48-91, static configuration of src/rrgcn.py:101-106,146-152) and the angle loss (:225-247).
// =====================================================================================================
// dh[u] = sum_{in-edges (w->u, r')} norm[w] * dAgg[w]_b . W[inv(r')]_b^T : every edge (u->v, r) of the snapshot graph has
// its inverse (v->u, r +- R) as an in-edge of u, so the forward CSR-by-destination serves again.
__global__ void __launch_bounds__(256) block_aggregate_bwd_h_kernel(
    const float* __restrict__ dAgg, const float* __restrict__ W, const int* __restrict__ rowptr,
    const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted, const float* __restrict__ norm, int N,
    int d_in, int d_out, int nb, int R, float* __restrict__ dh) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= N) return;
  const int si = d_in / nb, so = d_out / nb;
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  for (int j0 = 0; j0 < d_in; j0 += kWarp) {
    const int j = j0 + lane;
    if (j >= d_in) continue;
    const int b = j / si, i = j - b * si;
    float acc = 0.f;
    for (int e = beg; e < end; ++e) {
      const int w = __ldg(src_sorted + e), t = __ldg(etype_sorted + e);
      const int ti = t < R ? t + R : t - R;
      const float* gp = dAgg + (size_t)w * d_out + b * so;
      const float* wp = W + (size_t)ti * ((size_t)nb * si * so) + ((size_t)b * si + i) * so;
      float m = 0.f;
      for (int o = 0; o < so; ++o) m = fmaf(gp[o], __ldg(wp + o), m);
      acc = fmaf(__ldg(norm + w), m, acc);
    }
    dh[(size_t)row * d_in + j] = acc;
  }
}
// dW[r]_b[i][o] = sum_{e: type r} norm[dst] h[src]_{b,i} dAgg[dst]_{b,o}; CTA = (relation, split of its edge list),
// partial (nsplit, R2*wsz) summed afterwards by col_sum (fixed order).
__global__ void __launch_bounds__(256) block_aggregate_bwd_w_kernel(
    const float* __restrict__ h, const float* __restrict__ dAgg, const int* __restrict__ type_rowptr,
    const int* __restrict__ type_src, const int* __restrict__ type_dst, const float* __restrict__ norm, int d_in,
    int d_out, int nb, int nsplit, int R2, float* __restrict__ partial) {
  pdl_grid_sync();
  const int r = blockIdx.x, sp = blockIdx.y;
  const int si = d_in / nb, so = d_out / nb;
  const int wsz = nb * si * so;
  const int tb = __ldg(type_rowptr + r), te = __ldg(type_rowptr + r + 1);
  const int per = (te - tb + nsplit - 1) / nsplit;
  const int e0 = tb + sp * per, e1 = min(te, e0 + per);
  for (int idx = threadIdx.x; idx < wsz; idx += blockDim.x) {
    const int b = idx / (si * so), rem = idx - b * si * so;
    const int i = rem / so, o = rem - i * so;
    float acc = 0.f;
    for (int e = e0; e < e1; ++e) {
      const int sr = __ldg(type_src + e), ds = __ldg(type_dst + e);
      acc = fmaf(__ldg(norm + ds) * h[(size_t)sr * d_in + b * si + i], dAgg[(size_t)ds * d_out + b * so + o], acc);
    }
    partial[((size_t)sp * R2 + r) * wsz + idx] = acc;
  }
}
constexpr int kBlockWSplit = 16;
size_t block_aggregate_bwd_w_workspace_bytes(int R2, int d_in, int d_out, int nb) {
  const size_t wsz = (size_t)nb * (d_in / nb) * (d_out / nb);
  return (size_t)kBlockWSplit * R2 * wsz * sizeof(float) + col_reduce_workspace_bytes(kBlockWSplit, (int)(R2 * wsz));
}
int block_aggregate_bwd(const float* h, const float* dAgg, const float* W, const int* rowptr, const int* src_sorted,
                        const int* etype_sorted, const float* norm, const int* type_rowptr, const int* type_src,
                        const int* type_dst, int N, int R2, int d_in, int d_out, int nb, float* dh, float* dW, float* ws,
                        size_t ws_bytes, cudaStream_t st) {
  if (!h || !dAgg || !W || !rowptr || !src_sorted || !etype_sorted || !norm) { set_last_error("block_aggregate_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (nb <= 0 || d_in % nb || d_out % nb || (R2 & 1)) { set_last_error("block_aggregate_bwd: bad num_bases / relation count"); return REGCN_ERR_UNSUPPORTED; }
  if (N <= 0) return REGCN_OK;
  if (dh) launch_k(block_aggregate_bwd_h_kernel, rgrid(N), 256, 0, st, dAgg, W, rowptr, src_sorted, etype_sorted, norm, N, d_in, d_out, nb, R2 / 2, dh);
  if (dW) {
    if (!type_rowptr || !type_src || !type_dst || !ws) { set_last_error("block_aggregate_bwd: null pointer"); return REGCN_ERR_NULL; }
    if (ws_bytes < block_aggregate_bwd_w_workspace_bytes(R2, d_in, d_out, nb)) { set_last_error("block_aggregate_bwd: workspace too small"); return REGCN_ERR_WORKSPACE; }
    const int wsz = nb * (d_in / nb) * (d_out / nb);
    float* partial = ws;
    float* ws2 = ws + (size_t)kBlockWSplit * R2 * wsz;
    dim3 grid((unsigned)R2, (unsigned)kBlockWSplit);
    launch_k(block_aggregate_bwd_w_kernel, grid, 256, 0, st, h, dAgg, type_rowptr, type_src, type_dst, norm, d_in, d_out, nb, (int)kBlockWSplit, R2, partial);
    return col_sum(partial, R2 * wsz, kBlockWSplit, R2 * wsz, 1, dW, 0, ws2, ws_bytes - (size_t)kBlockWSplit * R2 * wsz * sizeof(float), st);
  }
  return check_launch("block_aggregate_bwd");
}

// Angle loss term of one history step (src/rrgcn.py:225-247): sim = <s, e/|e|> (layer_norm) or <s,e>/(|s||e|);
// term[row] = weight * max(cos_step - sim, 0).  The caller sums the (L*N) terms with col_sum.
template <int RV>
__global__ void __launch_bounds__(256) static_angle_fwd_kernel(const float* __restrict__ S, const float* __restrict__ E,
                                                               int N, int d, float cos_step, float weight,
                                                               int layer_norm, float* __restrict__ term) {
  pdl_grid_sync();
  ROW_PROLOGUE(N)
  WarpRow<RV> s, e;
  s.load_plain(S + (size_t)row * d, nvec, lane);
  e.load_plain(E + (size_t)row * d, nvec, lane);
  float sim;
  if (layer_norm) {
    row_l2normalize(e);
    sim = s.dot(e);
  } else {
    const float c = sqrtf(s.sumsq()) * sqrtf(e.sumsq());
    sim = s.dot(e) / c;
  }
  const float v = cos_step - sim;
  if (lane == 0) term[row] = v > 0.f ? weight * v : 0.f;
}
// dS (+)= -w m ehat (through normalize(s) when not layer_norm), dE = normalize_bwd(e, -w m shat); m = [cos - sim > 0]
template <int RV>
__global__ void __launch_bounds__(256) static_angle_bwd_kernel(const float* __restrict__ S, const float* __restrict__ E,
                                                               int N, int d, float cos_step, float weight,
                                                               int layer_norm, const float* __restrict__ gscale,
                                                               float* __restrict__ dS, int accumulate_dS,
                                                               float* __restrict__ dE) {
  pdl_grid_sync();
  ROW_PROLOGUE(N)
  WarpRow<RV> s, e;
  s.load_plain(S + (size_t)row * d, nvec, lane);
  e.load_plain(E + (size_t)row * d, nvec, lane);
  WarpRow<RV> eh = e, sh = s;
  row_l2normalize(eh);
  float sim;
  if (layer_norm) sim = s.dot(eh);
  else { sh.scale(1.0f / sqrtf(s.sumsq())); eh = e; eh.scale(1.0f / sqrtf(e.sumsq())); sim = sh.dot(eh); }
  const float g = (cos_step - sim > 0.f) ? -weight * (gscale ? *gscale : 1.0f) : 0.f;
  WarpRow<RV> gs = eh, ge = sh;                // d sim / d s(hat) = ehat, d sim / d ehat = s(hat)
  gs.scale(g);
  ge.scale(g);
  if (!layer_norm) {                            // s enters through s/|s| (no clamp in the reference: plain division)
    const float n = sqrtf(s.sumsq());
    const float dotv = sh.dot(gs);
    gs.zip(sh, [=](float a, float y) { return (a - y * dotv) / n; });
    const float ne = sqrtf(e.sumsq());
    const float dote = eh.dot(ge);
    ge.zip(eh, [=](float a, float y) { return (a - y * dote) / ne; });
  } else {
    row_normalize_bwd(e, ge);
    ge = e;
  }
  float* ds = dS + (size_t)row * d;
  if (accumulate_dS) {
    WarpRow<RV> prev;
    prev.load_plain(ds, nvec, lane);
    gs.zip(prev, [](float a, float p) { return a + p; });
  }
  gs.store(ds, nvec, lane);
  ge.store(dE + (size_t)row * d, nvec, lane);
}
int static_angle_fwd(const float* S, const float* E, int N, int d, float cos_step, float weight, int layer_norm,
                     float* term, cudaStream_t st) {
  if (!S || !E || !term) { set_last_error("static_angle_fwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk_d("static_angle_fwd", d)) return e;
  if (N <= 0) return REGCN_OK;
  if (d <= 128) launch_k(static_angle_fwd_kernel<1>, rgrid(N), 256, 0, st, S, E, N, d, cos_step, weight, layer_norm, term);
  else launch_k(static_angle_fwd_kernel<2>, rgrid(N), 256, 0, st, S, E, N, d, cos_step, weight, layer_norm, term);
  return check_launch("static_angle_fwd");
}
int static_angle_bwd(const float* S, const float* E, int N, int d, float cos_step, float weight, int layer_norm,
                     const float* gscale, float* dS, int accumulate_dS, float* dE, cudaStream_t st) {
  if (!S || !E || !dS || !dE) { set_last_error("static_angle_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk_d("static_angle_bwd", d)) return e;
  if (N <= 0) return REGCN_OK;
  if (d <= 128) launch_k(static_angle_bwd_kernel<1>, rgrid(N), 256, 0, st, S, E, N, d, cos_step, weight, layer_norm, gscale, dS, accumulate_dS, dE);
  else launch_k(static_angle_bwd_kernel<2>, rgrid(N), 256, 0, st, S, E, N, d, cos_step, weight, layer_norm, gscale, dS, accumulate_dS, dE);
  return check_launch("static_angle_bwd");
}

}  // namespace regcn

// ---------------------------------------------------------------------------------------------------------------
// C ABI (declared in include/regcn_b200.h, "training" section)
// ---------------------------------------------------------------------------------------------------------------
using namespace regcn;
#define ST(s) ((cudaStream_t)(s))
extern "C" {
int regcn_csr_gather_sum(const float* X, int ldx, const float* col_w, const float* row_w, const int32_t* rowptr,
                         const int32_t* col, int nrows, int d, int col2_off, float* out, int ldo, int accumulate,
                         const float* rho, float gamma, const int32_t* partner, void* stream) {
  return csr_gather_sum(X, ldx, col_w, row_w, rowptr, col, nrows, d, col2_off, out, ldo, accumulate, ST(stream), rho, gamma,
                        partner);
}
size_t regcn_group_by_key_workspace_bytes(int n) { return group_by_key_workspace_bytes(n); }
int regcn_group_by_key(const int32_t* keys, int n, int nkeys, const int32_t* vals, int32_t* rowptr, int32_t* perm,
                       int32_t* vals_out, void* workspace, size_t workspace_bytes, void* stream) {
  return group_by_key(keys, n, nkeys, vals, rowptr, perm, vals_out, workspace, workspace_bytes, ST(stream));
}
int regcn_expand_rowptr(const int32_t* rowptr, int nrows, int nnz, int32_t* rowid, float* inv_len, void* stream) {
  return expand_rowptr(rowptr, nrows, nnz, rowid, inv_len, ST(stream));
}
int regcn_normalize_bwd(const float* x, const float* dy, float* dx, int M, int d, void* stream) {
  return normalize_bwd(x, dy, dx, M, d, ST(stream));
}
int regcn_gru_gate_bwd(const float* gi, const float* gh, const float* hprev, const float* dout, int M, int d,
                       int normalize, float* dgi, float* dgh, float* dhprev, void* stream) {
  return gru_gate_bwd(gi, gh, hprev, dout, M, d, normalize, dgi, dgh, dhprev, ST(stream));
}
int regcn_union_combine_bwd(const float* out, const float* dout, const int32_t* indeg, int N, int d, float p, float* dP,
                            float* dL, void* stream) {
  return union_combine_bwd(out, dout, indeg, N, d, p, dP, dL, ST(stream));
}
int regcn_time_gate_bwd(const float* G, const float* bias, const float* cur, const float* h, const float* dout, int N,
                        int d, int normalize_cur, float* dG, float* dcur, float* dh, void* stream) {
  return time_gate_bwd(G, bias, cur, h, dout, N, d, normalize_cur, dG, dcur, dh, ST(stream));
}
int regcn_tanh_bwd(const float* y, const float* dy, float* dx, size_t n, void* stream) {
  return tanh_bwd(y, dy, dx, n, ST(stream));
}
int regcn_dropout(float* x, size_t n, float p, uint32_t seed, void* stream) {
  return dropout_inplace(x, n, p, seed, ST(stream));
}
size_t regcn_col_reduce_workspace_bytes(int rows, int cols) { return col_reduce_workspace_bytes(rows, cols); }
int regcn_bn_stats(const float* X, int B, int C, int L, float eps, float momentum, float* mean, float* invstd,
                   float* running_mean, float* running_var, float* workspace, size_t workspace_bytes, void* stream) {
  return bn_stats(X, B, C, L, eps, momentum, mean, invstd, running_mean, running_var, workspace, workspace_bytes, ST(stream));
}
int regcn_bn_bwd_stats(const float* dZ, const float* Z, const float* Y, int B, int C, int L, int mask_mode,
                       float mask_scale, const float* mean, const float* invstd, float* sum_dy, float* sum_dy_xhat,
                       float* workspace, size_t workspace_bytes, void* stream) {
  return bn_bwd_stats(dZ, Z, Y, B, C, L, mask_mode, mask_scale, mean, invstd, sum_dy, sum_dy_xhat, workspace,
                      workspace_bytes, ST(stream));
}
int regcn_col_sum(const float* X, int ld, int rows, int cols, float* out, int accumulate, float* workspace,
                  size_t workspace_bytes, void* stream) {
  return col_sum(X, ld, rows, cols, 1, out, accumulate, workspace, workspace_bytes, ST(stream));
}
int regcn_bn_act_drop(const float* X, int B, int C, int L, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, int relu, float p, uint32_t seed, float* out, void* stream) {
  return bn_act_drop(X, B, C, L, mean, invstd, gamma, beta, relu, p, seed, out, ST(stream));
}
int regcn_bn_bwd_apply(const float* dZ, const float* Z, const float* Y, int B, int C, int L, int mask_mode,
                       float mask_scale, const float* mean, const float* invstd, const float* gamma,
                       const float* sum_dy, const float* sum_dy_xhat, const float* out_src, int out_mode,
                       float out_scale, float* dX, void* stream) {
  return bn_bwd_apply(dZ, Z, Y, B, C, L, mask_mode, mask_scale, mean, invstd, gamma, sum_dy, sum_dy_xhat, out_src,
                      out_mode, out_scale, dX, ST(stream));
}
int regcn_dec_gather_stack(const float* first, const float* second, const int64_t* triples, int col0, int col1, int B,
                           int d, float* X0, void* stream) {
  return dec_gather_stack(first, second, triples, col0, col1, B, d, X0, ST(stream));
}
int regcn_dec_conv_fwd(const float* X0, int B, int d, int C, int ksz, const float* mean0, const float* invstd0,
                       const float* gamma0, const float* beta0, float p, uint32_t seed, const float* W,
                       const float* bias, float* X1, float* Y, void* stream) {
  return dec_conv_fwd(X0, B, d, C, ksz, mean0, invstd0, gamma0, beta0, p, seed, W, bias, X1, Y, ST(stream));
}
int regcn_dec_conv_bwd_input(const float* dY, const float* X1, int B, int d, int C, int ksz, const float* W, float p,
                             float* dX1, void* stream) {
  return dec_conv_bwd_input(dY, X1, B, d, C, ksz, W, p, dX1, ST(stream));
}
size_t regcn_dec_conv_bwd_weight_workspace_bytes(int B, int C) { return dec_conv_bwd_weight_workspace_bytes(B, C); }
int regcn_dec_conv_bwd_weight(const float* dY, const float* X1, int B, int d, int C, int ksz, float* dW_db,
                              float* workspace, size_t workspace_bytes, void* stream) {
  return dec_conv_bwd_weight(dY, X1, B, d, C, ksz, dW_db, workspace, workspace_bytes, ST(stream));
}
int regcn_ce_lse_rows(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, float* ce,
                      float* lse, float* loss, void* stream) {
  return ce_lse_rows(S, ld, B, N, triples, target_col, ce, lse, loss, ST(stream));
}
int regcn_softmax_grad_rows(float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, const float* lse,
                            const float* gscale, void* stream) {
  return softmax_grad_rows(S, ld, B, N, triples, target_col, lse, gscale, ST(stream));
}
int regcn_transpose_split(const float* X, int rows, int cols, int ldx, float* out, float* out_hi, float* out_lo, int ldo,
                          void* stream) {
  return transpose_split(X, rows, cols, ldx, out, out_hi, out_lo, ldo, ST(stream));
}
size_t regcn_adam_workspace_bytes(void) { return adam_workspace_bytes(); }
int regcn_grad_norm(const float* g, size_t n, float* total_norm, void* workspace, size_t workspace_bytes, void* stream) {
  return grad_norm(g, n, total_norm, workspace, workspace_bytes, ST(stream));
}
int regcn_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, float max_norm, const float* total_norm, void* stream) {
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, max_norm, total_norm, ST(stream));
}
size_t regcn_block_aggregate_bwd_workspace_bytes(int R2, int d_in, int d_out, int num_bases) {
  return block_aggregate_bwd_w_workspace_bytes(R2, d_in, d_out, num_bases);
}
int regcn_block_aggregate_bwd(const float* h, const float* dAgg, const float* W, const int32_t* rowptr,
                              const int32_t* src_sorted, const int32_t* etype_sorted, const float* norm,
                              const int32_t* type_rowptr, const int32_t* type_src, const int32_t* type_dst, int N, int R2,
                              int d_in, int d_out, int num_bases, float* dh, float* dW, float* workspace,
                              size_t workspace_bytes, void* stream) {
  return block_aggregate_bwd(h, dAgg, W, rowptr, src_sorted, etype_sorted, norm, type_rowptr, type_src, type_dst, N, R2,
                             d_in, d_out, num_bases, dh, dW, workspace, workspace_bytes, ST(stream));
}
int regcn_static_angle_fwd(const float* S, const float* E, int N, int d, float cos_step, float weight, int layer_norm,
                           float* term, void* stream) {
  return static_angle_fwd(S, E, N, d, cos_step, weight, layer_norm, term, ST(stream));
}
int regcn_static_angle_bwd(const float* S, const float* E, int N, int d, float cos_step, float weight, int layer_norm,
                           const float* gscale, float* dS, int accumulate_dS, float* dE, void* stream) {
  return static_angle_bwd(S, E, N, d, cos_step, weight, layer_norm, gscale, dS, accumulate_dS, dE, ST(stream));
}
}  // extern "C"
