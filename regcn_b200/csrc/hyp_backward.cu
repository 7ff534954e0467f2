// Backward kernels of the hyperbolic evolution step (second half): HyperbolicRecurrentRGCN.get_loss
// in train mode for the hyperbolic_uvrgcn encoder + hyperbolic_convtranse decoder (hyperbolic_model.py:722-890,941-1088,
// hyperbolic_layers.py:222-323, hyperbolic_ops.py:38-233,395-435, hyperbolic_decoder.py:360-413).
//
// Every Poincare row map on the path is RADIAL: y = s(n) x with n = max(|x|, eps) -- exp_0 (with its projection),
// log_0, project_to_ball, F.normalize, the tangent normalisation exp_0(normalize(log_0 x)) -- so one backward kernel
// serves them all:  dx = s dy + (s'(n)/n) <x,dy> x   (s' = 0 where the norm clamp is active).
#include "common.cuh"
#include "internal.h"

namespace regcn {

#define ROWP(M_)                                                                       \
  const int lane = threadIdx.x & 31;                                                   \
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);         \
  if (row >= (M_)) return;                                                             \
  const int nvec = d >> 2;
static inline unsigned rg(int M) { return (unsigned)(((size_t)M * 32 + 255) / 256); }
static inline int chk(const char* who, int d) {
  if (d <= 0 || (d & 3) || d > 256) { set_last_error("%s: d=%d unsupported (need d%%4==0, d<=256)", who, d); return REGCN_ERR_UNSUPPORTED; }
  return REGCN_OK;
}

// mode 0 log_0, 1 exp_0 (+ projection), 2 project_to_ball, 3 exp_0(F.normalize(log_0 x))   -> (s, s')
__device__ __forceinline__ void radial_coeff(int mode, float nraw, const Curv& k, float& s, float& sp) {
  const float n = fmaxf(nraw, kEps);
  const bool free_n = nraw > kEps;                 // below eps the norm clamp holds n constant
  const float sc = k.sqrt_c;
  if (mode == 0) {
    const float u = sc * n;
    const bool clamped = u >= 1.0f - kEps;
    const float a = atanhf(fminf(u, 1.0f - kEps));
    s = a / (sc * n);
    sp = (clamped ? 0.f : 1.0f / (n * (1.0f - k.c * n * n))) - a / (sc * n * n);
  } else if (mode == 1) {
    const float t = tanhf(sc * n);
    const float m = t / sc;                        // |exp_0(v)| before the projection
    if (m > k.proj_max) { s = k.proj_max / n; sp = -k.proj_max / (n * n); }
    else { s = t / (sc * n); sp = (1.0f - t * t) / n - t / (sc * n * n); }
  } else if (mode == 2) {
    if (n > k.proj_max) { s = k.proj_max / n; sp = -k.proj_max / (n * n); }
    else { s = 1.0f; sp = 0.f; }
  } else {
    const float K = fminf(tanhf(sc) / sc, k.proj_max);   // |exp_0(unit vector)|, projected
    s = K / n; sp = -K / (n * n);
  }
  if (!free_n) sp = 0.f;
}

template <int RV>
__global__ void __launch_bounds__(256) radial_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dx, int M, int d, int mode, Curv cv) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a, g;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  g.load_plain(dy + (size_t)row * d, nvec, lane);
  const float nraw = sqrtf(a.sumsq());
  float s, sp;
  radial_coeff(mode, nraw, cv, s, sp);
  const float n = fmaxf(nraw, kEps);
  const float coef = sp / n * a.dot(g);
  g.zip(a, [=](float gg, float xx) { return s * gg + coef * xx; });
  g.store(dx + (size_t)row * d, nvec, lane);
}
int radial_bwd(const float* x, const float* dy, float* dx, int M, int d, int mode, double c, cudaStream_t st) {
  if (!x || !dy || !dx) { set_last_error("radial_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("radial_bwd", d)) return e;
  if (mode < 0 || mode > 3 || !(c > 0)) { set_last_error("radial_bwd: bad mode / curvature"); return REGCN_ERR_DIM; }
  if (M <= 0) return REGCN_OK;
  Curv cv = make_curv(c);
  if (d <= 128) launch_k(radial_bwd_kernel<1>, rg(M), 256, 0, st, x, dy, dx, M, d, mode, cv);
  else launch_k(radial_bwd_kernel<2>, rg(M), 256, 0, st, x, dy, dx, M, d, mode, cv);
  return check_launch("radial_bwd");
}

// radius (hyperbolic_ops.py:206): rho = max(|x|, eps);   apply_radius (:222-233): y = x / max(|x|,eps) * clamp(r, eps, rmax)
template <int RV>
__global__ void __launch_bounds__(256) row_radius_kernel(const float* __restrict__ x, int M, int d, float* __restrict__ rho) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  const float n = fmaxf(sqrtf(a.sumsq()), kEps);
  if (lane == 0) rho[row] = n;
}
// dx (+)= drho * x / |x|
template <int RV>
__global__ void __launch_bounds__(256) row_radius_bwd_kernel(const float* __restrict__ x, const float* __restrict__ drho,
                                                             int M, int d, float* __restrict__ dx) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  const float nraw = sqrtf(a.sumsq());
  const float f = nraw > kEps ? __ldg(drho + row) / nraw : 0.f;
  a.scale(f);
  a.store(dx + (size_t)row * d, nvec, lane);
}
template <int RV>
__global__ void __launch_bounds__(256) apply_radius_kernel(const float* __restrict__ x, const float* __restrict__ r, int M,
                                                           int d, Curv cv, float* __restrict__ y) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  row_apply_radius(a, __ldg(r + row), cv);
  a.store(y + (size_t)row * d, nvec, lane);
}
template <int RV>
__global__ void __launch_bounds__(256) apply_radius_bwd_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                               const float* __restrict__ dy, int M, int d, Curv cv,
                                                               float* __restrict__ dx, float* __restrict__ dr) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a, g;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  g.load_plain(dy + (size_t)row * d, nvec, lane);
  const float nraw = sqrtf(a.sumsq());
  const float n = fmaxf(nraw, kEps);
  const float rv = __ldg(r + row);
  const float rr = clampf_(rv, kEps, cv.radius_max);
  a.scale(1.0f / n);                                     // unit direction
  const float dotv = a.dot(g);
  if (lane == 0) dr[row] = (rv > kEps && rv < cv.radius_max) ? dotv : 0.f;
  const bool free_n = nraw > kEps;
  const float f = rr / n;
  g.zip(a, [=](float gg, float u) { return f * (gg - (free_n ? u * dotv : 0.f)); });
  g.store(dx + (size_t)row * d, nvec, lane);
}
int row_radius(const float* x, int M, int d, float* rho, cudaStream_t st) {
  if (!x || !rho) { set_last_error("row_radius: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("row_radius", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(row_radius_kernel<1>, rg(M), 256, 0, st, x, M, d, rho);
  else launch_k(row_radius_kernel<2>, rg(M), 256, 0, st, x, M, d, rho);
  return check_launch("row_radius");
}
int row_radius_bwd(const float* x, const float* drho, int M, int d, float* dx, cudaStream_t st) {
  if (!x || !drho || !dx) { set_last_error("row_radius_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("row_radius_bwd", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(row_radius_bwd_kernel<1>, rg(M), 256, 0, st, x, drho, M, d, dx);
  else launch_k(row_radius_bwd_kernel<2>, rg(M), 256, 0, st, x, drho, M, d, dx);
  return check_launch("row_radius_bwd");
}
int apply_radius_fwd(const float* x, const float* r, int M, int d, double c, float* y, cudaStream_t st) {
  if (!x || !r || !y) { set_last_error("apply_radius: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("apply_radius", d)) return e;
  if (M <= 0) return REGCN_OK;
  Curv cv = make_curv(c);
  if (d <= 128) launch_k(apply_radius_kernel<1>, rg(M), 256, 0, st, x, r, M, d, cv, y);
  else launch_k(apply_radius_kernel<2>, rg(M), 256, 0, st, x, r, M, d, cv, y);
  return check_launch("apply_radius");
}
int apply_radius_bwd(const float* x, const float* r, const float* dy, int M, int d, double c, float* dx, float* dr,
                     cudaStream_t st) {
  if (!x || !r || !dy || !dx || !dr) { set_last_error("apply_radius_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("apply_radius_bwd", d)) return e;
  if (M <= 0) return REGCN_OK;
  Curv cv = make_curv(c);
  if (d <= 128) launch_k(apply_radius_bwd_kernel<1>, rg(M), 256, 0, st, x, r, dy, M, d, cv, dx, dr);
  else launch_k(apply_radius_bwd_kernel<2>, rg(M), 256, 0, st, x, r, dy, M, d, cv, dx, dr);
  return check_launch("apply_radius_bwd");
}

// Elementwise pieces.  op 0: y = clamp(x, -lim, lim)  (torch.clamp: gradient 1 inside the closed interval)
//                      op 1: y = 0.9 tanh(x) + 0.1 x  (HyperbolicConvTransE entity activation, hyperbolic_decoder.py:378)
//                      op 2: y = -x;  op 3: y = relu(x)  (RotH's residual tangent MLP, :1028-1030)
__global__ void __launch_bounds__(256) eltwise_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n,
                                                          int op, float lim) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  y[i] = op == 0 ? clampf_(v, -lim, lim) : op == 1 ? 0.9f * tanhf(v) + 0.1f * v : op == 2 ? -v : fmaxf(v, 0.f);
}
__global__ void __launch_bounds__(256) eltwise_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dx, size_t n, int op, float lim) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  float f;
  if (op == 0) f = (v >= -lim && v <= lim) ? 1.f : 0.f;
  else if (op == 1) { const float t = tanhf(v); f = 0.9f * (1.0f - t * t) + 0.1f; }
  else if (op == 2) f = -1.f;
  else f = v > 0.f ? 1.f : 0.f;
  dx[i] = dy[i] * f;
}
int eltwise_fwd(const float* x, float* y, size_t n, int op, float lim, cudaStream_t st) {
  if (!x || !y) { set_last_error("eltwise_fwd: null pointer"); return REGCN_ERR_NULL; }
  if (n == 0) return REGCN_OK;
  launch_k(eltwise_fwd_kernel, (unsigned)((n + 255) / 256), 256, 0, st, x, y, n, op, lim);
  return check_launch("eltwise_fwd");
}
int eltwise_bwd(const float* x, const float* dy, float* dx, size_t n, int op, float lim, cudaStream_t st) {
  if (!x || !dy || !dx) { set_last_error("eltwise_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (n == 0) return REGCN_OK;
  launch_k(eltwise_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, st, x, dy, dx, n, op, lim);
  return check_launch("eltwise_bwd");
}

// TemporalRadiusEvolution scalars (hyperbolic_ops.py:406-425) and _static_radius (hyperbolic_model.py:715-720):
//   rs = min(clamp(raw, rmin, rmax), cap);  new_r = beta rs + (1-beta) dyn + clamp(delta, -eps_r, eps_r)
// dyn / delta NULL: new_r = rs (the plain static radius).
__global__ void radius_combine_kernel(const float* __restrict__ raw, const float* __restrict__ dyn,
                                      const float* __restrict__ delta, int M, float rmin, float rmax, float cap, float beta,
                                      float eps_r, float* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float rs = fminf(clampf_(raw[i], rmin, rmax), cap);
  out[i] = dyn ? beta * rs + (1.0f - beta) * dyn[i] + clampf_(delta[i], -eps_r, eps_r) : rs;
}
__global__ void radius_combine_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ delta,
                                          const float* __restrict__ g, int M, float rmin, float rmax, float cap, float beta,
                                          float eps_r, int has_dyn, float* __restrict__ draw, float* __restrict__ ddyn,
                                          float* __restrict__ ddelta) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float r = raw[i], gi = g[i];
  const float cl = clampf_(r, rmin, rmax);
  const float pass = (r >= rmin && r <= rmax && cl <= cap) ? 1.f : 0.f;      // torch.clamp / torch.min gradients
  draw[i] = gi * pass * (has_dyn ? beta : 1.0f);
  if (has_dyn) {
    ddyn[i] = gi * (1.0f - beta);
    const float dl = delta[i];
    ddelta[i] = (dl >= -eps_r && dl <= eps_r) ? gi : 0.f;
  }
}
int radius_combine(const float* raw, const float* dyn, const float* delta, int M, float rmin, float rmax, double c,
                   float beta, float eps_r, float* out, cudaStream_t st) {
  if (!raw || !out || ((dyn != nullptr) != (delta != nullptr))) { set_last_error("radius_combine: null pointer"); return REGCN_ERR_NULL; }
  if (M <= 0) return REGCN_OK;
  const float cap = (float)(1.0 / sqrt(c) - 1e-6);
  launch_k(radius_combine_kernel, (unsigned)((M + 255) / 256), 256, 0, st, raw, dyn, delta, M, rmin, rmax, cap, beta, eps_r, out);
  return check_launch("radius_combine");
}
int radius_combine_bwd(const float* raw, const float* delta, const float* g, int M, float rmin, float rmax, double c,
                       float beta, float eps_r, float* draw, float* ddyn, float* ddelta, cudaStream_t st) {
  if (!raw || !g || !draw || ((ddyn != nullptr) != (ddelta != nullptr)) || (ddyn && !delta)) { set_last_error("radius_combine_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (M <= 0) return REGCN_OK;
  const float cap = (float)(1.0 / sqrt(c) - 1e-6);
  launch_k(radius_combine_bwd_kernel, (unsigned)((M + 255) / 256), 256, 0, st, raw, delta, g, M, rmin, rmax, cap, beta, eps_r,
           ddyn ? 1 : 0, draw, ddyn, ddelta);
  return check_launch("radius_combine_bwd");
}

// radius_mlp = Linear(d, 1) (hyperbolic_ops.py:390-392,407): delta[n] = <t[n], w> + b;  backward: dt = ddelta (x) w and the
// scaled rows ddelta[n] t[n] whose column sums are dw (col_sum); db = sum ddelta.
template <int RV>
__global__ void __launch_bounds__(256) row_dot_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                                      const float* __restrict__ b, int M, int d, float* __restrict__ out) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a, ww;
  a.load_plain(t + (size_t)row * d, nvec, lane);
  ww.load(w, nvec, lane);
  const float v = a.dot(ww) + (b ? __ldg(b) : 0.f);
  if (lane == 0) out[row] = v;
}
template <int RV>
__global__ void __launch_bounds__(256) row_dot_bwd_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                                          const float* __restrict__ dout, int M, int d,
                                                          float* __restrict__ dt, float* __restrict__ scaled) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a, ww;
  a.load_plain(t + (size_t)row * d, nvec, lane);
  ww.load(w, nvec, lane);
  const float g = __ldg(dout + row);
  ww.scale(g);
  ww.store(dt + (size_t)row * d, nvec, lane);
  a.scale(g);
  a.store(scaled + (size_t)row * d, nvec, lane);
}
int row_dot(const float* t, const float* w, const float* b, int M, int d, float* out, cudaStream_t st) {
  if (!t || !w || !out) { set_last_error("row_dot: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("row_dot", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(row_dot_kernel<1>, rg(M), 256, 0, st, t, w, b, M, d, out);
  else launch_k(row_dot_kernel<2>, rg(M), 256, 0, st, t, w, b, M, d, out);
  return check_launch("row_dot");
}
int row_dot_bwd(const float* t, const float* w, const float* dout, int M, int d, float* dt, float* scaled, cudaStream_t st) {
  if (!t || !w || !dout || !dt || !scaled) { set_last_error("row_dot_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("row_dot_bwd", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(row_dot_bwd_kernel<1>, rg(M), 256, 0, st, t, w, dout, M, d, dt, scaled);
  else launch_k(row_dot_bwd_kernel<2>, rg(M), 256, 0, st, t, w, dout, M, d, dt, scaled);
  return check_launch("row_dot_bwd");
}

// Radius-difference edge weights (hyperbolic_layers.py:232-234): w_e = exp(-gamma |rho_src - rho_dst|) multiplies the message
// of edge e.  Gradient w.r.t. the radii: with m_e = h_tan[src] + rel[type] and v = dst,
//   s_e = -gamma sign(rho_src - rho_dst) norm[v] w_e <dAgg[v], m_e>;   drho[src] += s_e,  drho[dst] -= s_e.
// One warp per destination row: the 200-wide dot per in-edge, its own -sum written directly, s_e stored per CSR position
// for the by-source pass (edge_scalar_gather).
template <int RV>
__global__ void __launch_bounds__(256) edge_radius_grad_kernel(
    const float* __restrict__ ht, const float* __restrict__ rel, const float* __restrict__ dagg,
    const int* __restrict__ rowptr, const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted,
    const float* __restrict__ norm, const float* __restrict__ rho, float gamma, int N, int d,
    float* __restrict__ s_edge, float* __restrict__ drho_dst) {
  pdl_grid_sync();
  ROWP(N)
  const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  WarpRow<RV> g;
  g.load_plain(dagg + (size_t)row * d, nvec, lane);
  const float nv = __ldg(norm + row), rv = __ldg(rho + row);
  float acc = 0.f;
  for (int p = b; p < e; ++p) {
    const int u = __ldg(src_sorted + p), t = __ldg(etype_sorted + p);
    WarpRow<RV> m, r;
    m.load_plain(ht + (size_t)u * d, nvec, lane);
    r.load(rel + (size_t)t * d, nvec, lane);
    m.zip(r, [](float a, float bb) { return a + bb; });
    const float dotv = g.dot(m);
    const float diff = __ldg(rho + u) - rv;
    const float w = expf(-gamma * fabsf(diff));
    const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    const float s = -gamma * sgn * nv * w * dotv;
    if (lane == 0) s_edge[p] = s;
    acc += s;
  }
  if (lane == 0) drho_dst[row] = -acc;
}
// out[row] (+)= sum_{j in row} vals[perm[j]]  (per-source sums of the per-edge scalars; fixed order)
__global__ void edge_scalar_gather_kernel(const float* __restrict__ vals, const int* __restrict__ rowptr,
                                          const int* __restrict__ perm, int nrows, float* __restrict__ out, int accumulate) {
  pdl_grid_sync();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  float a = 0.f;
  for (int j = rowptr[row]; j < rowptr[row + 1]; ++j) a += vals[perm[j]];
  out[row] = (accumulate ? out[row] : 0.f) + a;
}
int edge_radius_grad(const float* ht, const float* rel, const float* dagg, const int* rowptr, const int* src_sorted,
                     const int* etype_sorted, const float* norm, const float* rho, float gamma, int N, int d,
                     float* s_edge, float* drho_dst, cudaStream_t st) {
  if (!ht || !rel || !dagg || !rowptr || !src_sorted || !etype_sorted || !norm || !rho || !s_edge || !drho_dst) { set_last_error("edge_radius_grad: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("edge_radius_grad", d)) return e;
  if (N <= 0) return REGCN_OK;
  if (d <= 128) launch_k(edge_radius_grad_kernel<1>, rg(N), 256, 0, st, ht, rel, dagg, rowptr, src_sorted, etype_sorted, norm, rho, gamma, N, d, s_edge, drho_dst);
  else launch_k(edge_radius_grad_kernel<2>, rg(N), 256, 0, st, ht, rel, dagg, rowptr, src_sorted, etype_sorted, norm, rho, gamma, N, d, s_edge, drho_dst);
  return check_launch("edge_radius_grad");
}
int edge_scalar_gather(const float* vals, const int* rowptr, const int* perm, int nrows, float* out, int accumulate,
                       cudaStream_t st) {
  if (!vals || !rowptr || !perm || !out) { set_last_error("edge_scalar_gather: null pointer"); return REGCN_ERR_NULL; }
  if (nrows <= 0) return REGCN_OK;
  launch_k(edge_scalar_gather_kernel, (unsigned)((nrows + 255) / 256), 256, 0, st, vals, rowptr, perm, nrows, out, accumulate);
  return check_launch("edge_scalar_gather");
}

// loss_radius (hyperbolic_model.py:1066-1073): lambda * mean_{i in ids} (rs[i] - target[i])^2 over the entities of the batch;
// term[j] per id (summed by col_sum) and the gradient w.r.t. radius_static (through _static_radius' clamps).
__global__ void radius_mse_kernel(const float* __restrict__ raw, const float* __restrict__ target,
                                  const int64_t* __restrict__ ids, int n, float rmin, float rmax, float cap, float scale,
                                  float* __restrict__ term) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t i = ids[j];
  const float rs = fminf(clampf_(raw[i], rmin, rmax), cap);
  const float df = rs - target[i];
  term[j] = scale * df * df;
}
__global__ void radius_mse_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ target,
                                      const int64_t* __restrict__ ids, int n, float rmin, float rmax, float cap, float scale,
                                      const float* __restrict__ gscale, float* __restrict__ draw) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t i = ids[j];                              // ids are unique: no write conflicts
  const float r = raw[i];
  const float cl = clampf_(r, rmin, rmax);
  const float rs = fminf(cl, cap);
  const float pass = (r >= rmin && r <= rmax && cl <= cap) ? 1.f : 0.f;
  draw[i] = 2.0f * scale * (rs - target[i]) * pass * (gscale ? *gscale : 1.0f);
}
int radius_mse(const float* raw, const float* target, const int64_t* ids, int n, float rmin, float rmax, double c,
               float lambda, float* term, cudaStream_t st) {
  if (!raw || !target || !ids || !term) { set_last_error("radius_mse: null pointer"); return REGCN_ERR_NULL; }
  if (n <= 0) return REGCN_OK;
  launch_k(radius_mse_kernel, (unsigned)((n + 255) / 256), 256, 0, st, raw, target, ids, n, rmin, rmax,
           (float)(1.0 / sqrt(c) - 1e-6), lambda / (float)n, term);
  return check_launch("radius_mse");
}
int radius_mse_bwd(const float* raw, const float* target, const int64_t* ids, int n, float rmin, float rmax, double c,
                   float lambda, const float* gscale, float* draw, cudaStream_t st) {
  if (!raw || !target || !ids || !draw) { set_last_error("radius_mse_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (n <= 0) return REGCN_OK;
  launch_k(radius_mse_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, st, raw, target, ids, n, rmin, rmax,
           (float)(1.0 / sqrt(c) - 1e-6), lambda / (float)n, gscale, draw);
  return check_launch("radius_mse_bwd");
}

}  // namespace regcn

using namespace regcn;
#define ST(s) ((cudaStream_t)(s))
extern "C" {
int regcn_radial_bwd(const float* x, const float* dy, float* dx, int M, int d, int mode, double c, void* stream) {
  return radial_bwd(x, dy, dx, M, d, mode, c, ST(stream));
}
int regcn_row_radius(const float* x, int M, int d, float* rho, void* stream) { return row_radius(x, M, d, rho, ST(stream)); }
int regcn_row_radius_bwd(const float* x, const float* drho, int M, int d, float* dx, void* stream) {
  return row_radius_bwd(x, drho, M, d, dx, ST(stream));
}
int regcn_apply_radius(const float* x, const float* r, int M, int d, double c, float* y, void* stream) {
  return apply_radius_fwd(x, r, M, d, c, y, ST(stream));
}
int regcn_apply_radius_bwd(const float* x, const float* r, const float* dy, int M, int d, double c, float* dx, float* dr,
                           void* stream) {
  return apply_radius_bwd(x, r, dy, M, d, c, dx, dr, ST(stream));
}
int regcn_eltwise_fwd(const float* x, float* y, size_t n, int op, float lim, void* stream) {
  return eltwise_fwd(x, y, n, op, lim, ST(stream));
}
int regcn_eltwise_bwd(const float* x, const float* dy, float* dx, size_t n, int op, float lim, void* stream) {
  return eltwise_bwd(x, dy, dx, n, op, lim, ST(stream));
}
int regcn_radius_combine(const float* raw, const float* dyn, const float* delta, int M, float rmin, float rmax, double c,
                         float beta, float eps_r, float* out, void* stream) {
  return radius_combine(raw, dyn, delta, M, rmin, rmax, c, beta, eps_r, out, ST(stream));
}
int regcn_radius_combine_bwd(const float* raw, const float* delta, const float* g, int M, float rmin, float rmax, double c,
                             float beta, float eps_r, float* draw, float* ddyn, float* ddelta, void* stream) {
  return radius_combine_bwd(raw, delta, g, M, rmin, rmax, c, beta, eps_r, draw, ddyn, ddelta, ST(stream));
}
int regcn_row_dot(const float* t, const float* w, const float* b, int M, int d, float* out, void* stream) {
  return row_dot(t, w, b, M, d, out, ST(stream));
}
int regcn_row_dot_bwd(const float* t, const float* w, const float* dout, int M, int d, float* dt, float* scaled,
                      void* stream) {
  return row_dot_bwd(t, w, dout, M, d, dt, scaled, ST(stream));
}
int regcn_edge_radius_grad(const float* ht, const float* rel, const float* dagg, const int32_t* rowptr,
                           const int32_t* src_sorted, const int32_t* etype_sorted, const float* norm, const float* rho,
                           float gamma, int N, int d, float* s_edge, float* drho_dst, void* stream) {
  return edge_radius_grad(ht, rel, dagg, rowptr, src_sorted, etype_sorted, norm, rho, gamma, N, d, s_edge, drho_dst,
                          ST(stream));
}
int regcn_edge_scalar_gather(const float* vals, const int32_t* rowptr, const int32_t* perm, int nrows, float* out,
                             int accumulate, void* stream) {
  return edge_scalar_gather(vals, rowptr, perm, nrows, out, accumulate, ST(stream));
}
int regcn_radius_mse(const float* raw, const float* target, const int64_t* ids, int n, float rmin, float rmax, double c,
                     float lambda, float* term, void* stream) {
  return radius_mse(raw, target, ids, n, rmin, rmax, c, lambda, term, ST(stream));
}
int regcn_radius_mse_bwd(const float* raw, const float* target, const int64_t* ids, int n, float rmin, float rmax, double c,
                         float lambda, const float* gscale, float* draw, void* stream) {
  return radius_mse_bwd(raw, target, ids, n, rmin, rmax, c, lambda, gscale, draw, ST(stream));
}
}

// =====================================================================================================================
// Distance decoders in training (HyperbolicMuRP / MuRPRel .loss, hyperbolic_decoder.py:647-928; the scoring core
// _chunked_hyperbolic_ce_loss :182-307 in the <q,e>, |q|^2, |e|^2 form of a18).
// =====================================================================================================================
namespace regcn {

// mobius_add without its final projection (a separate radial node): z = (a x + b y) / den,
//   a = 1 + 2c<x,y> + c|y|^2, b = 1 - c|x|^2, den = 1 + 2c<x,y> + c^2|x|^2|y|^2 + eps      (hyperbolic_ops.py:135-142)
template <int RV>
__global__ void __launch_bounds__(256) mobius_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int M,
                                                         int d, float c, float* __restrict__ z) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a, b;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  b.load_plain(y + (size_t)row * d, nvec, lane);
  const float X = a.sumsq(), Y = b.sumsq(), xy = a.dot(b);
  const float ca = 1.0f + 2.0f * c * xy + c * Y, cb = 1.0f - c * X;
  const float den = 1.0f + 2.0f * c * xy + c * c * X * Y + kEps;
  a.zip(b, [=](float xx, float yy) { return (ca * xx + cb * yy) / den; });
  a.store(z + (size_t)row * d, nvec, lane);
}
template <int RV>
__global__ void __launch_bounds__(256) mobius_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ dz, int M, int d, float c,
                                                         float* __restrict__ dx, float* __restrict__ dy) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a, b, g;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  b.load_plain(y + (size_t)row * d, nvec, lane);
  g.load_plain(dz + (size_t)row * d, nvec, lane);
  const float X = a.sumsq(), Y = b.sumsq(), xy = a.dot(b);
  const float ca = 1.0f + 2.0f * c * xy + c * Y, cb = 1.0f - c * X;
  const float den = 1.0f + 2.0f * c * xy + c * c * X * Y + kEps;
  const float gx = g.dot(a), gy = g.dot(b);
  const float gz = (ca * gx + cb * gy) / den;                  // <g, z>
  const float inv = 1.0f / den;
  // dx = (a/den) g + (2c gx/den) y - (2c gy/den) x - (gz/den)(2c y + 2c^2 Y x)
  const float kx_g = ca * inv, kx_y = 2.0f * c * gx * inv - gz * inv * 2.0f * c,
              kx_x = -2.0f * c * gy * inv - gz * inv * 2.0f * c * c * Y;
  // dy = (b/den) g + (gx/den)(2c x + 2c y) - (gz/den)(2c x + 2c^2 X y)
  const float ky_g = cb * inv, ky_x = 2.0f * c * gx * inv - gz * inv * 2.0f * c,
              ky_y = 2.0f * c * gx * inv - gz * inv * 2.0f * c * c * X;
  WarpRow<RV> ox, oy;
#pragma unroll
  for (int i = 0; i < RV; ++i) {
    ox.v[i] = make_float4(kx_g * g.v[i].x + kx_y * b.v[i].x + kx_x * a.v[i].x, kx_g * g.v[i].y + kx_y * b.v[i].y + kx_x * a.v[i].y,
                          kx_g * g.v[i].z + kx_y * b.v[i].z + kx_x * a.v[i].z, kx_g * g.v[i].w + kx_y * b.v[i].w + kx_x * a.v[i].w);
    oy.v[i] = make_float4(ky_g * g.v[i].x + ky_x * a.v[i].x + ky_y * b.v[i].x, ky_g * g.v[i].y + ky_x * a.v[i].y + ky_y * b.v[i].y,
                          ky_g * g.v[i].z + ky_x * a.v[i].z + ky_y * b.v[i].z, ky_g * g.v[i].w + ky_x * a.v[i].w + ky_y * b.v[i].w);
  }
  ox.store(dx + (size_t)row * d, nvec, lane);
  oy.store(dy + (size_t)row * d, nvec, lane);
}
int mobius_fwd(const float* x, const float* y, int M, int d, double c, float* z, cudaStream_t st) {
  if (!x || !y || !z) { set_last_error("mobius_fwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("mobius_fwd", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(mobius_fwd_kernel<1>, rg(M), 256, 0, st, x, y, M, d, (float)c, z);
  else launch_k(mobius_fwd_kernel<2>, rg(M), 256, 0, st, x, y, M, d, (float)c, z);
  return check_launch("mobius_fwd");
}
int mobius_bwd(const float* x, const float* y, const float* dz, int M, int d, double c, float* dx, float* dy, cudaStream_t st) {
  if (!x || !y || !dz || !dx || !dy) { set_last_error("mobius_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("mobius_bwd", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(mobius_bwd_kernel<1>, rg(M), 256, 0, st, x, y, dz, M, d, (float)c, dx, dy);
  else launch_k(mobius_bwd_kernel<2>, rg(M), 256, 0, st, x, y, dz, M, d, (float)c, dx, dy);
  return check_launch("mobius_bwd");
}

// z = x * y elementwise (MuRP's diagonal relation map, :748-752); the backward is the same kernel twice.
__global__ void __launch_bounds__(256) eltwise_mul_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          float* __restrict__ z, size_t n) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) z[i] = x[i] * y[i];
}
int eltwise_mul(const float* x, const float* y, float* z, size_t n, cudaStream_t st) {
  if (!x || !y || !z) { set_last_error("eltwise_mul: null pointer"); return REGCN_ERR_NULL; }
  if (n == 0) return REGCN_OK;
  launch_k(eltwise_mul_kernel, (unsigned)((n + 255) / 256), 256, 0, st, x, y, z, n);
  return check_launch("eltwise_mul");
}

// out[r] += alpha * s[r] * x[r]   (the |q|^2 / |e|^2 terms of the distance-score gradient)
template <int RV>
__global__ void __launch_bounds__(256) row_axpy_kernel(const float* __restrict__ x, const float* __restrict__ s, float alpha,
                                                       int M, int d, float* __restrict__ out) {
  pdl_grid_sync();
  ROWP(M)
  WarpRow<RV> a, o;
  a.load_plain(x + (size_t)row * d, nvec, lane);
  o.load_plain(out + (size_t)row * d, nvec, lane);
  const float f = alpha * __ldg(s + row);
  o.zip(a, [=](float oo, float xx) { return fmaf(f, xx, oo); });
  o.store(out + (size_t)row * d, nvec, lane);
}
int row_axpy(const float* x, const float* s, float alpha, int M, int d, float* out, cudaStream_t st) {
  if (!x || !s || !out) { set_last_error("row_axpy: null pointer"); return REGCN_ERR_NULL; }
  if (int e = chk("row_axpy", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(row_axpy_kernel<1>, rg(M), 256, 0, st, x, s, alpha, M, d, out);
  else launch_k(row_axpy_kernel<2>, rg(M), 256, 0, st, x, s, alpha, M, d, out);
  return check_launch("row_axpy");
}

// Gradient of the proxy-distance score S = scale (margin - n^2) + bias, n = min(|(-q)(+)e|, proj_max), through its three
// scalars per (query, candidate): dot D, x2 = |q|^2, y2 = |e|^2 (a18's form; hyperbolic_decoder.py:164-172).
// In place: D[b,n] <- dS * dS/dD; H[b,n] <- dS * dS/dy2; per row: gx[b] = sum_n dS * dS/dx2, gs[b] = sum_n dS (margin - n^2),
// gm[b] = sum_n dS * scale.  One CTA per query row.
__global__ void __launch_bounds__(256) hyp_dist_grad_kernel(float* __restrict__ D, const float* __restrict__ dS,
                                                            float* __restrict__ H, int64_t ld, int B, int N,
                                                            const float* __restrict__ x2, const float* __restrict__ y2,
                                                            float c, float pm, const float* __restrict__ scale_margin,
                                                            float* __restrict__ gx, float* __restrict__ gs,
                                                            float* __restrict__ gm) {
  pdl_grid_sync();
  __shared__ float sh[3][8];
  const int b = blockIdx.x;
  const float scale = scale_margin[0], margin = scale_margin[1];
  const float X = x2[b];
  float ax = 0.f, as = 0.f, am = 0.f;
  for (int j = threadIdx.x; j < (int)ld; j += blockDim.x) {
    const size_t i = (size_t)b * ld + j;
    float g1 = 0.f, h = 0.f;
    if (j < N) {
      const float g = dS[i];
      const float u = -D[i], Y = y2[j];
      const float A = 1.0f + 2.0f * c * u + c * Y, Bc = 1.0f - c * X;
      const float num = fmaxf(A * A * X + 2.0f * A * Bc * u + Bc * Bc * Y, 0.f);
      const float den = 1.0f + 2.0f * c * u + c * c * X * Y + kEps;
      const float n2 = num / (den * den);
      const bool clamped = sqrtf(num) / fabsf(den) > pm;
      const float n2c = clamped ? pm * pm : n2;
      as += g * (margin - n2c);
      am += g * scale;
      if (!clamped) {
        const float inv2 = 1.0f / (den * den), k3 = 2.0f * num * inv2 / den;
        const float dnum_u = 4.0f * c * A * X + 4.0f * c * Bc * u + 2.0f * A * Bc;
        const float dnum_X = A * A - 2.0f * c * A * u - 2.0f * c * Bc * Y;
        const float dnum_Y = 2.0f * c * A * X + 2.0f * c * Bc * u + Bc * Bc;
        const float dn2_u = dnum_u * inv2 - k3 * (2.0f * c);
        const float dn2_X = dnum_X * inv2 - k3 * (c * c * Y);
        const float dn2_Y = dnum_Y * inv2 - k3 * (c * c * X);
        g1 = g * scale * dn2_u;                     // dS/dD = scale dn2/du   (u = -D, S = -scale n^2)
        ax += g * (-scale) * dn2_X;
        h = g * (-scale) * dn2_Y;
      }
    }
    D[i] = g1;
    H[i] = h;
  }
  float v[3] = {ax, as, am};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float s = warp_sum(v[q]);
    if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sh[threadIdx.x][w];
    (threadIdx.x == 0 ? gx : threadIdx.x == 1 ? gs : gm)[b] = s;
  }
}
int hyp_dist_grad(float* D, const float* dS, float* H, int64_t ld, int B, int N, const float* x2, const float* y2, double c,
                  const float* scale_margin, float* gx, float* gs, float* gm, cudaStream_t st) {
  if (!D || !dS || !H || !x2 || !y2 || !scale_margin || !gx || !gs || !gm) { set_last_error("hyp_dist_grad: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0 || N <= 0) return REGCN_OK;
  const Curv cv = make_curv(c);
  launch_k(hyp_dist_grad_kernel, (unsigned)B, 256, 0, st, D, dS, H, ld, B, N, x2, y2, cv.c, cv.proj_max, scale_margin, gx, gs, gm);
  return check_launch("hyp_dist_grad");
}

}  // namespace regcn

extern "C" {
int regcn_mobius_fwd(const float* x, const float* y, int M, int d, double c, float* z, void* stream) {
  return regcn::mobius_fwd(x, y, M, d, c, z, (cudaStream_t)stream);
}
int regcn_mobius_bwd(const float* x, const float* y, const float* dz, int M, int d, double c, float* dx, float* dy,
                     void* stream) {
  return regcn::mobius_bwd(x, y, dz, M, d, c, dx, dy, (cudaStream_t)stream);
}
int regcn_eltwise_mul(const float* x, const float* y, float* z, size_t n, void* stream) {
  return regcn::eltwise_mul(x, y, z, n, (cudaStream_t)stream);
}
int regcn_row_axpy(const float* x, const float* s, float alpha, int M, int d, float* out, void* stream) {
  return regcn::row_axpy(x, s, alpha, M, d, out, (cudaStream_t)stream);
}
int regcn_hyp_dist_grad(float* D, const float* dS, float* H, int64_t ld, int B, int N, const float* x2, const float* y2,
                        double c, const float* scale_margin, float* gx, float* gs, float* gm, void* stream) {
  return regcn::hyp_dist_grad(D, dS, H, ld, B, N, x2, y2, c, scale_margin, gx, gs, gm, (cudaStream_t)stream);
}
}

// =====================================================================================================================
// RotH / RotHRel / AttH in training: Givens rotation / reflection of the tangent pairs (hyperbolic_decoder.py:1033-1051,
// 1380-1401), forward and backward w.r.t. the vector and the angles.
// =====================================================================================================================
namespace regcn {
// mode 0 rotation: (c x1 - s x2, s x1 + c x2);  mode 1 reflection: (c x1 + s x2, s x1 - c x2).
// ang: (B, d/2) or, with ang_bcast, one (d/2) vector shared by all rows.
__global__ void __launch_bounds__(256) givens_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ang,
                                                         int ang_bcast, size_t npairs, int half, int mode,
                                                         float* __restrict__ y) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npairs) return;
  const float a = ang[ang_bcast ? i % half : i];
  float s, c;
  sincosf(a, &s, &c);
  const float2 v = reinterpret_cast<const float2*>(x)[i];
  reinterpret_cast<float2*>(y)[i] = mode == 0 ? make_float2(c * v.x - s * v.y, s * v.x + c * v.y)
                                              : make_float2(c * v.x + s * v.y, s * v.x - c * v.y);
}
__global__ void __launch_bounds__(256) givens_bwd_kernel(const float* __restrict__ x, const float* __restrict__ ang,
                                                         const float* __restrict__ dy, int ang_bcast, size_t npairs,
                                                         int half, int mode, float* __restrict__ dx,
                                                         float* __restrict__ dang) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npairs) return;
  const float a = ang[ang_bcast ? i % half : i];
  float s, c;
  sincosf(a, &s, &c);
  const float2 v = reinterpret_cast<const float2*>(x)[i];
  const float2 g = reinterpret_cast<const float2*>(dy)[i];
  if (mode == 0) {
    reinterpret_cast<float2*>(dx)[i] = make_float2(c * g.x + s * g.y, -s * g.x + c * g.y);
    dang[i] = g.x * (-s * v.x - c * v.y) + g.y * (c * v.x - s * v.y);
  } else {
    reinterpret_cast<float2*>(dx)[i] = make_float2(c * g.x + s * g.y, s * g.x - c * g.y);
    dang[i] = g.x * (-s * v.x + c * v.y) + g.y * (c * v.x + s * v.y);
  }
}
int givens_fwd(const float* x, const float* ang, int ang_bcast, int B, int d, int mode, float* y, cudaStream_t st) {
  if (!x || !ang || !y) { set_last_error("givens_fwd: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 1) || mode < 0 || mode > 1) { set_last_error("givens_fwd: bad d / mode"); return REGCN_ERR_DIM; }
  const size_t np = (size_t)B * (d / 2);
  if (np == 0) return REGCN_OK;
  launch_k(givens_fwd_kernel, (unsigned)((np + 255) / 256), 256, 0, st, x, ang, ang_bcast, np, d / 2, mode, y);
  return check_launch("givens_fwd");
}
int givens_bwd(const float* x, const float* ang, const float* dy, int ang_bcast, int B, int d, int mode, float* dx,
               float* dang, cudaStream_t st) {
  if (!x || !ang || !dy || !dx || !dang) { set_last_error("givens_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 1) || mode < 0 || mode > 1) { set_last_error("givens_bwd: bad d / mode"); return REGCN_ERR_DIM; }
  const size_t np = (size_t)B * (d / 2);
  if (np == 0) return REGCN_OK;
  launch_k(givens_bwd_kernel, (unsigned)((np + 255) / 256), 256, 0, st, x, ang, dy, ang_bcast, np, d / 2, mode, dx, dang);
  return check_launch("givens_bwd");
}
}  // namespace regcn
extern "C" {
int regcn_givens_fwd(const float* x, const float* ang, int ang_bcast, int B, int d, int mode, float* y, void* stream) {
  return regcn::givens_fwd(x, ang, ang_bcast, B, d, mode, y, (cudaStream_t)stream);
}
int regcn_givens_bwd(const float* x, const float* ang, const float* dy, int ang_bcast, int B, int d, int mode, float* dx,
                     float* dang, void* stream) {
  return regcn::givens_bwd(x, ang, dy, ang_bcast, B, d, mode, dx, dang, (cudaStream_t)stream);
}
}

// =====================================================================================================================
// AttH / AttHRel attention mix (hyperbolic_decoder.py:1434-1445, 1617-1625): a = sigmoid(<w, u>) over 2d features,
// mixed = a rot + (1 - a) ref.  One warp per query row.
// =====================================================================================================================
namespace regcn {
__global__ void __launch_bounds__(256) attn_mix_fwd_kernel(const float* __restrict__ w, int w_bcast, const float* __restrict__ u,
                                                           const float* __restrict__ rot, const float* __restrict__ ref,
                                                           int B, int d, float* __restrict__ a_out, float* __restrict__ mixed) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= B) return;
  const float* wr = w_bcast ? w : w + (size_t)row * 2 * d;
  const float* ur = u + (size_t)row * 2 * d;
  float acc = 0.f;
  for (int j = lane; j < 2 * d; j += 32) acc = fmaf(wr[j], ur[j], acc);
  const float a = sigmoidf_(warp_sum(acc));
  if (lane == 0) a_out[row] = a;
  for (int j = lane; j < d; j += 32) {
    const size_t i = (size_t)row * d + j;
    mixed[i] = a * rot[i] + (1.0f - a) * ref[i];
  }
}
__global__ void __launch_bounds__(256) attn_mix_bwd_kernel(const float* __restrict__ w, int w_bcast, const float* __restrict__ u,
                                                           const float* __restrict__ rot, const float* __restrict__ ref,
                                                           const float* __restrict__ a_in, const float* __restrict__ g, int B,
                                                           int d, float* __restrict__ dw, float* __restrict__ du,
                                                           float* __restrict__ drot, float* __restrict__ dref) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= B) return;
  const float a = a_in[row];
  float acc = 0.f;
  for (int j = lane; j < d; j += 32) {
    const size_t i = (size_t)row * d + j;
    const float gg = g[i];
    acc = fmaf(gg, rot[i] - ref[i], acc);
    drot[i] = a * gg;
    dref[i] = (1.0f - a) * gg;
  }
  const float dlogit = warp_sum(acc) * a * (1.0f - a);
  const float* wr = w_bcast ? w : w + (size_t)row * 2 * d;
  const float* ur = u + (size_t)row * 2 * d;
  for (int j = lane; j < 2 * d; j += 32) {
    const size_t i = (size_t)row * 2 * d + j;
    dw[i] = dlogit * ur[j];
    du[i] = dlogit * wr[j];
  }
}
int attn_mix_fwd(const float* w, int w_bcast, const float* u, const float* rot, const float* ref, int B, int d, float* a,
                 float* mixed, cudaStream_t st) {
  if (!w || !u || !rot || !ref || !a || !mixed) { set_last_error("attn_mix_fwd: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0 || d <= 0) return REGCN_OK;
  launch_k(attn_mix_fwd_kernel, rg(B), 256, 0, st, w, w_bcast, u, rot, ref, B, d, a, mixed);
  return check_launch("attn_mix_fwd");
}
int attn_mix_bwd(const float* w, int w_bcast, const float* u, const float* rot, const float* ref, const float* a,
                 const float* g, int B, int d, float* dw, float* du, float* drot, float* dref, cudaStream_t st) {
  if (!w || !u || !rot || !ref || !a || !g || !dw || !du || !drot || !dref) { set_last_error("attn_mix_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0 || d <= 0) return REGCN_OK;
  launch_k(attn_mix_bwd_kernel, rg(B), 256, 0, st, w, w_bcast, u, rot, ref, a, g, B, d, dw, du, drot, dref);
  return check_launch("attn_mix_bwd");
}
}  // namespace regcn
extern "C" {
int regcn_attn_mix_fwd(const float* w, int w_bcast, const float* u, const float* rot, const float* ref, int B, int d,
                       float* a, float* mixed, void* stream) {
  return regcn::attn_mix_fwd(w, w_bcast, u, rot, ref, B, d, a, mixed, (cudaStream_t)stream);
}
int regcn_attn_mix_bwd(const float* w, int w_bcast, const float* u, const float* rot, const float* ref, const float* a,
                       const float* g, int B, int d, float* dw, float* du, float* drot, float* dref, void* stream) {
  return regcn::attn_mix_bwd(w, w_bcast, u, rot, ref, a, g, B, d, dw, du, drot, dref, (cudaStream_t)stream);
}
}

// =====================================================================================================================
// Relation-specific curvature in training (--plus-relation-specific-curvature, hyperbolic_decoder.py:66-86,145-163): the
// true-distance score S = scale (margin - 2/(sqrt(c_q+eps)+eps) atanh(min(sqrt(c_q+eps) n, 1-1e-6))) with a per-query
// curvature, differentiated w.r.t. the dot, the two squared norms and c_q.
// =====================================================================================================================
namespace regcn {
__global__ void __launch_bounds__(256) hyp_truedist_grad_kernel(float* __restrict__ D, const float* __restrict__ dS,
                                                                float* __restrict__ H, int64_t ld, int B, int N,
                                                                const float* __restrict__ x2, const float* __restrict__ y2,
                                                                const float* __restrict__ row_c,
                                                                const float* __restrict__ scale_margin,
                                                                float* __restrict__ gx, float* __restrict__ gs,
                                                                float* __restrict__ gm, float* __restrict__ gc) {
  pdl_grid_sync();
  __shared__ float sh[4][8];
  const int b = blockIdx.x;
  const float scale = scale_margin[0], margin = scale_margin[1];
  const float X = x2[b], c = row_c[b];
  const float sc = sqrtf(c + kEps), inv = sc + kEps;
  const float dsc = 0.5f / sc;                          // d sqrt(c+eps)/dc = d inv/dc
  const float nmax = 1.0f / inv - kEps;
  float ax = 0.f, as = 0.f, am = 0.f, ac = 0.f;
  for (int j = threadIdx.x; j < (int)ld; j += blockDim.x) {
    const size_t i = (size_t)b * ld + j;
    float g1 = 0.f, h = 0.f;
    if (j < N) {
      const float g = dS[i];
      const float Dv = D[i], Y = y2[j];
      const float a = 1.0f - 2.0f * c * Dv + c * Y, bb = 1.0f - c * X;
      const float num = fmaxf(a * a * X - 2.0f * a * bb * Dv + bb * bb * Y, 0.f);
      const float den = 1.0f - 2.0f * c * Dv + c * c * X * Y + kEps;
      const float nraw = sqrtf(num) / fabsf(den);
      const bool lo = nraw < kEps, hi = nraw > nmax;
      const float n = lo ? kEps : (hi ? nmax : nraw);
      const float argr = sc * n;
      const bool aclamp = argr >= 0.999999f;
      const float arg = fminf(argr, 0.999999f);
      const float at = atanhf(arg);
      const float dist = 2.0f / inv * at;
      as += g * (margin - dist);
      am += g * scale;
      const float ddist_dn = aclamp ? 0.f : (2.0f / inv) * sc / (1.0f - arg * arg);
      // d dist / d c with n held fixed, plus the moving upper bound of n
      float ddist_dc = -2.0f / (inv * inv) * dsc * at + (aclamp ? 0.f : (2.0f / inv) * dsc * n / (1.0f - arg * arg));
      if (hi) ddist_dc += ddist_dn * (-dsc / (inv * inv));
      if (!lo && !hi && num > 0.f) {
        const float q = Y - 2.0f * Dv;
        const float dnum_D = -4.0f * c * a * X - 2.0f * a * bb + 4.0f * bb * c * Dv;
        const float dnum_X = a * a + 2.0f * a * c * Dv - 2.0f * bb * c * Y;
        const float dnum_Y = 2.0f * a * c * X - 2.0f * bb * c * Dv + bb * bb;
        const float dnum_c = 2.0f * a * X * q - 2.0f * bb * Dv * q + 2.0f * a * Dv * X - 2.0f * bb * X * Y;
        const float i2n = 0.5f / num, iden = 1.0f / den;
        const float dn_D = n * (dnum_D * i2n - (-2.0f * c) * iden);
        const float dn_X = n * (dnum_X * i2n - (c * c * Y) * iden);
        const float dn_Y = n * (dnum_Y * i2n - (c * c * X) * iden);
        const float dn_c = n * (dnum_c * i2n - (-2.0f * Dv + 2.0f * c * X * Y) * iden);
        g1 = g * (-scale) * ddist_dn * dn_D;
        ax += g * (-scale) * ddist_dn * dn_X;
        h = g * (-scale) * ddist_dn * dn_Y;
        ddist_dc += ddist_dn * dn_c;
      }
      ac += g * (-scale) * ddist_dc;
    }
    D[i] = g1;
    H[i] = h;
  }
  float v[4] = {ax, as, am, ac};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float s = warp_sum(v[q]);
    if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sh[threadIdx.x][w];
    (threadIdx.x == 0 ? gx : threadIdx.x == 1 ? gs : threadIdx.x == 2 ? gm : gc)[b] = s;
  }
}
// c_q = max(1e-5, min(softplus(raw[r mod R]), upper)): per-query d raw = dc_q * sigmoid(raw) where neither clamp is active
__global__ void rel_curvature_bwd_kernel(const float* __restrict__ raw, const int64_t* __restrict__ triples, int B, int R,
                                         float upper, const float* __restrict__ dcq, float* __restrict__ draw_q,
                                         int* __restrict__ base_rel) {
  pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int r = (int)(triples[3 * (size_t)b + 1] % R);
  const float x = raw[r];
  const float sp = x > 20.f ? x : log1pf(expf(x));
  const float sg = x > 20.f ? 1.f : 1.0f / (1.0f + expf(-x));
  draw_q[b] = (sp <= upper && sp >= 1e-5f) ? dcq[b] * sg : 0.f;
  base_rel[b] = r;
}
int hyp_truedist_grad(float* D, const float* dS, float* H, int64_t ld, int B, int N, const float* x2, const float* y2,
                      const float* row_c, const float* scale_margin, float* gx, float* gs, float* gm, float* gc,
                      cudaStream_t st) {
  if (!D || !dS || !H || !x2 || !y2 || !row_c || !scale_margin || !gx || !gs || !gm || !gc) { set_last_error("hyp_truedist_grad: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0 || N <= 0) return REGCN_OK;
  launch_k(hyp_truedist_grad_kernel, (unsigned)B, 256, 0, st, D, dS, H, ld, B, N, x2, y2, row_c, scale_margin, gx, gs, gm, gc);
  return check_launch("hyp_truedist_grad");
}
int rel_curvature_bwd(const float* raw, const int64_t* triples, int B, int R, double c, double cmax, const float* dcq,
                      float* draw_q, int* base_rel, cudaStream_t st) {
  if (!raw || !triples || !dcq || !draw_q || !base_rel) { set_last_error("rel_curvature_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0) return REGCN_OK;
  float upper = 0.999f * (float)c;
  if (cmax > 0 && (float)cmax < upper) upper = (float)cmax;
  launch_k(rel_curvature_bwd_kernel, (unsigned)((B + 255) / 256), 256, 0, st, raw, triples, B, R, upper, dcq, draw_q, base_rel);
  return check_launch("rel_curvature_bwd");
}
}  // namespace regcn
extern "C" {
int regcn_hyp_truedist_grad(float* D, const float* dS, float* H, int64_t ld, int B, int N, const float* x2, const float* y2,
                            const float* row_c, const float* scale_margin, float* gx, float* gs, float* gm, float* gc,
                            void* stream) {
  return regcn::hyp_truedist_grad(D, dS, H, ld, B, N, x2, y2, row_c, scale_margin, gx, gs, gm, gc, (cudaStream_t)stream);
}
int regcn_rel_curvature_bwd(const float* raw, const int64_t* triples, int B, int R, double c, double cmax, const float* dcq,
                            float* draw_q, int32_t* base_rel, void* stream) {
  return regcn::rel_curvature_bwd(raw, triples, B, R, c, cmax, dcq, draw_q, base_rel, (cudaStream_t)stream);
}
}
