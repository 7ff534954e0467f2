// Row-wise fused epilogue kernels: one warp owns one row of d <= 256 floats in registers, so every
// per-row reduction (L2 norm, exp_0/log_0, projection, radius) is a 5-step shuffle, never a re-read.
//   K3  GRU gates            src/rrgcn.py:168-174 (nn.GRUCell), hyperbolic_model.py:815-824
//   K5  self-loop combine    rgcn/layers.py:226-255, hyperbolic_layers.py:273-323, :649-694
//   K8  hyperbolic row maps  hyperbolic_ops.py:38-233, 395-435; hyperbolic_model.py:715-720,779-782
//   K9  time gate            src/rrgcn.py:176-178, hyperbolic_model.py:829-867
#include "common.cuh"

namespace regcn {

#define ROW_KERNEL_PROLOGUE(M_)                                                        \
  const int lane = threadIdx.x & 31;                                                   \
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);         \
  if (row >= (M_)) return;                                                             \
  const int nvec = d >> 2;

static inline unsigned row_grid(int M) { return (unsigned)(((size_t)M * 32 + 255) / 256); }
static inline int check_d(const char* who, int d) {
  if (d <= 0 || (d & 3) || d > 256) { set_last_error("%s: d=%d unsupported (need d%%4==0, d<=256)", who, d); return REGCN_ERR_UNSUPPORTED; }
  return REGCN_OK;
}

// ---- F.normalize rows / elementwise entity activation for the decoders ---------------------
// mode 0: out = x / max(|x|,1e-12)                    (F.normalize, src/rrgcn.py:154,170,176,190)
// mode 1: out = tanh(x)                               (ConvTransE entity table, src/decoder.py:79)
// mode 2: t = log_0(x); out = 0.9*tanh(t) + 0.1*t     (HyperbolicConvTransE, hyperbolic_decoder.py:377-379)
// mode 3: out = log_0(x);  mode 4: out = exp_0(x);  mode 5: out = project_to_ball(x)
// mode 6: out = exp_0(normalize(log_0(x)))            (predict-time layer_norm, hyperbolic_model.py:926-929)
// mode 7: identity; with out == NULL only the row |x|^2 is produced
// mode 8: out = exp_0(rrelu(x))                       (HyperbolicRGCNLayer tail, hyperbolic_layers.py:149-159)
// mode 9: out = tanh(x / max(|x|,1e-12))              (predict-time F.normalize + ConvTransE activation in one pass,
//                                                      src/rrgcn.py:190 + src/decoder.py:79)
// sumsq (optional): |out|^2 per row (for the norm/dot form of the hyperbolic scores)
template <int RV>
__global__ void __launch_bounds__(256) row_map_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int d,
                                                      int mode, Curv cv, float* __restrict__ sumsq,
                                                      float* __restrict__ out_hi, float* __restrict__ out_lo) {
  pdl_grid_sync();
  ROW_KERNEL_PROLOGUE(M)
  WarpRow<RV> r;
  r.load_plain(x + (size_t)row * d, nvec, lane);
  switch (mode) {
    case 0: row_l2normalize(r); break;
    case 1: r.map([](float a) { return tanhf(a); }); break;
    case 2: row_log0(r, cv); r.map([](float a) { return 0.9f * tanhf(a) + 0.1f * a; }); break;
    case 3: row_log0(r, cv); break;
    case 4: row_exp0(r, cv); break;
    case 5: row_project(r, cv); break;
    case 6: row_log0(r, cv); row_l2normalize(r); row_exp0(r, cv); break;
    case 8: r.map([](float a) { return rreluf_(a); }); row_exp0(r, cv); break;
    case 9: row_l2normalize(r); r.map([](float a) { return tanhf(a); }); break;
    default: break;  // mode 7: identity (row |x|^2 only)
  }
  if (out) r.store(out + (size_t)row * d, nvec, lane);
  if (out_hi) r.store_split(out_hi + (size_t)row * d, out_lo + (size_t)row * d, nvec, lane);
  if (sumsq) {
    float s = r.sumsq();
    if (lane == 0) sumsq[row] = s;
  }
}

int row_map(const float* x, float* out, int M, int d, int mode, double c, float* sumsq, float* out_hi, float* out_lo,
            cudaStream_t st) {
  if (!x || (!out && !sumsq && !out_hi) || (out_hi && !out_lo)) { set_last_error("row_map: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("row_map", d)) return e;
  if (mode < 0 || mode > 9) { set_last_error("row_map: bad mode %d", mode); return REGCN_ERR_DIM; }
  if (M <= 0) return REGCN_OK;
  Curv cv = make_curv(c > 0 ? c : 1.0);
  if (d <= 128) launch_k(row_map_kernel<1>, row_grid(M), 256, 0, st, x, out, M, d, mode, cv, sumsq, out_hi, out_lo);
  else launch_k(row_map_kernel<2>, row_grid(M), 256, 0, st, x, out, M, d, mode, cv, sumsq, out_hi, out_lo);
  return check_launch("row_map");
}

// ---- K3: GRU gates ---------------------------------------------------------------------------
// gi = x.W_ih^T + b_ih, gh = h.W_hh^T + b_hh (both M x 3d, gate order r,z,n like nn.GRUCell)
// r = s(gi_r+gh_r); z = s(gi_z+gh_z); n = tanh(gi_n + r*gh_n); h' = (h - n)*z + n; optional F.normalize
template <int RV>
__global__ void __launch_bounds__(256) gru_gate_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                       const float* __restrict__ hprev, float* __restrict__ out,
                                                       int M, int d, int normalize, float* __restrict__ out_hi,
                                                       float* __restrict__ out_lo) {
  pdl_grid_sync();
  ROW_KERNEL_PROLOGUE(M)
  WarpRow<RV> ir, iz, in_, hr, hz, hn, h;
  const float* gir = gi + (size_t)row * 3 * d;
  const float* ghr = gh + (size_t)row * 3 * d;
  ir.load_plain(gir, nvec, lane); iz.load_plain(gir + d, nvec, lane); in_.load_plain(gir + 2 * d, nvec, lane);
  hr.load_plain(ghr, nvec, lane); hz.load_plain(ghr + d, nvec, lane); hn.load_plain(ghr + 2 * d, nvec, lane);
  h.load_plain(hprev + (size_t)row * d, nvec, lane);
  ir.zip(hr, [](float a, float b) { return sigmoidf_(b + a); });   // reset gate
  iz.zip(hz, [](float a, float b) { return sigmoidf_(b + a); });   // update ("input") gate
  hn.zip(ir, [](float a, float r) { return a * r; });
  in_.zip(hn, [](float a, float b) { return tanhf(a + b); });      // new gate
  h.zip(in_, [](float hh, float n) { return hh - n; });
  h.zip(iz, [](float a, float z) { return a * z; });
  h.zip(in_, [](float a, float n) { return a + n; });
  if (normalize) row_l2normalize(h);
  h.store(out + (size_t)row * d, nvec, lane);
  if (out_hi) h.store_split(out_hi + (size_t)row * d, out_lo + (size_t)row * d, nvec, lane);
}

int gru_gate(const float* gi, const float* gh, const float* hprev, float* out, int M, int d, int normalize,
             float* out_hi, float* out_lo, cudaStream_t st) {
  if (!gi || !gh || !hprev || !out) { set_last_error("gru_gate: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("gru_gate", d)) return e;
  if (M <= 0) return REGCN_OK;
  if (d <= 128) launch_k(gru_gate_kernel<1>, row_grid(M), 256, 0, st, gi, gh, hprev, out, M, d, normalize, out_hi, out_lo);
  else launch_k(gru_gate_kernel<2>, row_grid(M), 256, 0, st, gi, gh, hprev, out, M, d, normalize, out_hi, out_lo);
  return check_launch("gru_gate");
}

// ---- K5: self-loop combine ---------------------------------------------------------------------
// P  (N x d)      neighbour term: agg.W_n (uvrgcn) or the aggregated tangent (lgcn, block layers)
// L  (N x 2d)     h.[W_loop | W_evolve] (null when self_loop is off); row picks cols [0,d) if indeg>0 else [d,2d)
// S  (N x d)      prev.W_skip (pre-bias) for the skip gate, null when off; skip_bias (d); prev (N x d)
// Euclidean  (hyper=0): out = act(P + loop)  /  act(s*(P+loop) + (1-s)*prev)          rgcn/layers.py:241-251
// Hyperbolic (hyper=1): t = clamp(P,+-10) (+loop, skip) ; t = clamp(t,+-10); t = act(t); out = exp_0(t)
//                       optional ht_next = log_0(out), radius_next = max(|out|,eps)     hyperbolic_layers.py:296-323
// act: 0 none, 1 rrelu(eval slope 11/48)
template <int RV>
__global__ void __launch_bounds__(256) union_combine_kernel(
    const float* __restrict__ P, const float* __restrict__ L, const int* __restrict__ indeg,
    const float* __restrict__ S, const float* __restrict__ skip_bias, const float* __restrict__ prev,
    int N, int d, int act, int hyper, Curv cv, float* __restrict__ out, float* __restrict__ ht_next,
    float* __restrict__ radius_next, int ldL, float* __restrict__ out_hi, float* __restrict__ out_lo,
    float* __restrict__ ht_hi, float* __restrict__ ht_lo, const int* __restrict__ active_pos) {
  pdl_grid_sync();
  ROW_KERNEL_PROLOGUE(N)
  WarpRow<RV> t, u;
  if (active_pos) {
    // sparse-snapshot form: active rows carry the complete pre-activation [agg | h].[W_n ; W_loop] in the compact
    // matrix P; inactive rows take h.W_evolve from L
    const int ap = __ldg(active_pos + row);
    if (ap >= 0) t.load_plain(P + (size_t)ap * d, nvec, lane);
    else t.load_plain(L + (size_t)row * ldL, nvec, lane);
  } else {
    t.load_plain(P + (size_t)row * d, nvec, lane);
  }
  if (hyper) t.map([](float a) { return clampf_(a, -10.f, 10.f); });
  if (L && !active_pos) {
    const int sel = __ldg(indeg + row) > 0 ? 0 : d;
    u.load_plain(L + (size_t)row * ldL + sel, nvec, lane);
    t.zip(u, [](float a, float b) { return a + b; });
  }
  if (S) {
    WarpRow<RV> g, b, p;
    g.load_plain(S + (size_t)row * d, nvec, lane);
    b.load(skip_bias, nvec, lane);
    p.load_plain(prev + (size_t)row * d, nvec, lane);
    g.zip(b, [](float a, float bb) { return sigmoidf_(a + bb); });
    t.zip(g, [](float a, float s) { return s * a; });
    p.zip(g, [](float pp, float s) { return (1.0f - s) * pp; });
    t.zip(p, [](float a, float bb) { return a + bb; });
  }
  if (hyper) t.map([](float a) { return clampf_(a, -10.f, 10.f); });
  if (act == 1) t.map([](float a) { return rreluf_(a); });
  if (hyper) row_exp0(t, cv);
  t.store(out + (size_t)row * d, nvec, lane);
  if (out_hi) t.store_split(out_hi + (size_t)row * d, out_lo + (size_t)row * d, nvec, lane);
  if (hyper && (ht_next || radius_next || ht_hi)) {
    float n = fmaxf(sqrtf(t.sumsq()), kEps);
    if (radius_next && lane == 0) radius_next[row] = n;
    if (ht_next || ht_hi) {
      row_log0(t, cv);
      if (ht_next) t.store(ht_next + (size_t)row * d, nvec, lane);
      if (ht_hi) t.store_split(ht_hi + (size_t)row * d, ht_lo + (size_t)row * d, nvec, lane);
    }
  }
}

int union_combine(const float* P, const float* L, const int* indeg, const float* S, const float* skip_bias,
                  const float* prev, int N, int d, int act, int hyper, double c, float* out, float* ht_next,
                  float* radius_next, int ldL, float* out_hi, float* out_lo, float* ht_hi, float* ht_lo,
                  const int* active_pos, cudaStream_t st) {
  if (!P || !out || (L && !indeg && !active_pos) || (S && (!skip_bias || !prev)) || (active_pos && (hyper || !L))) { set_last_error("union_combine: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("union_combine", d)) return e;
  if (N <= 0) return REGCN_OK;
  Curv cv = make_curv(hyper ? c : 1.0);
  if (d <= 128) launch_k(union_combine_kernel<1>, row_grid(N), 256, 0, st, P, L, indeg, S, skip_bias, prev, N, d, act, hyper, cv, out, ht_next, radius_next, ldL > 0 ? ldL : 2 * d, out_hi, out_lo, ht_hi, ht_lo, active_pos);
  else launch_k(union_combine_kernel<2>, row_grid(N), 256, 0, st, P, L, indeg, S, skip_bias, prev, N, d, act, hyper, cv, out, ht_next, radius_next, ldL > 0 ? ldL : 2 * d, out_hi, out_lo, ht_hi, ht_lo, active_pos);
  return check_launch("union_combine");
}

// ---- K9 (Euclidean): h' = s(G+b) * [normalize](cur) + (1 - s(G+b)) * h        src/rrgcn.py:176-178
template <int RV>
__global__ void __launch_bounds__(256) time_gate_kernel(const float* __restrict__ G, const float* __restrict__ bias,
                                                        const float* __restrict__ cur, const float* __restrict__ h,
                                                        float* __restrict__ out, int N, int d, int normalize_cur,
                                                        int ldg, float* __restrict__ out_hi, float* __restrict__ out_lo,
                                                        const int* __restrict__ row_idx, int act) {
  pdl_grid_sync();
  ROW_KERNEL_PROLOGUE(N)
  WarpRow<RV> g, b, c, hh;
  // row_idx: `cur` is a compact matrix (one row per listed entity), everything else is indexed by the entity id
  const int erow = row_idx ? __ldg(row_idx + row) : row;
  g.load_plain(G + (size_t)erow * ldg, nvec, lane);
  b.load(bias, nvec, lane);
  c.load_plain(cur + (size_t)row * d, nvec, lane);
  hh.load_plain(h + (size_t)erow * d, nvec, lane);
  if (act == 1) c.map([](float a) { return rreluf_(a); });
  if (normalize_cur) row_l2normalize(c);
  g.zip(b, [](float a, float bb) { return sigmoidf_(a + bb); });
  c.zip(g, [](float a, float s) { return s * a; });
  hh.zip(g, [](float a, float s) { return (1.0f - s) * a; });
  c.zip(hh, [](float a, float bb) { return a + bb; });
  c.store(out + (size_t)erow * d, nvec, lane);
  if (out_hi) c.store_split(out_hi + (size_t)erow * d, out_lo + (size_t)erow * d, nvec, lane);
}

int time_gate(const float* G, const float* bias, const float* cur, const float* h, float* out, int N, int d,
              int normalize_cur, int ldg, float* out_hi, float* out_lo, cudaStream_t st, const int* row_idx, int act) {
  if (!G || !bias || !cur || !h || !out) { set_last_error("time_gate: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("time_gate", d)) return e;
  if (N <= 0) return REGCN_OK;
  if (d <= 128) launch_k(time_gate_kernel<1>, row_grid(N), 256, 0, st, G, bias, cur, h, out, N, d, normalize_cur, ldg > 0 ? ldg : d, out_hi, out_lo, row_idx, act);
  else launch_k(time_gate_kernel<2>, row_grid(N), 256, 0, st, G, bias, cur, h, out, N, d, normalize_cur, ldg > 0 ? ldg : d, out_hi, out_lo, row_idx, act);
  return check_launch("time_gate");
}

// ---- shared-trajectory evolution (engine.cu: regcn_regcn_evolve_shared) ------------------------------------------
// G history windows evolved as one recurrence keep their entity state COMPACT: rows [0, N0) are the trajectory every
// entity follows until it first receives an edge in its window (the inactive-row update is row-local, and all windows
// start from the same table, so that trajectory is shared by all G windows); a row (g, v) of the union numbering gets
// its own compact row the first time it is active.  One warp per active row of the step: assigns the compact row
// (first activity: copies the shared row), publishes the row's state to the union-numbered table the edge kernels
// gather from, and records the compact positions of the step's active rows.
__global__ void __launch_bounds__(256) shared_rows_update_kernel(const int* __restrict__ active_rows, int n_active, int N0,
                                                                 int d, int* __restrict__ cpos, int* __restrict__ count,
                                                                 float* __restrict__ h_c, float* __restrict__ x_full,
                                                                 int* __restrict__ act_c, int* __restrict__ apos_c) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (i >= n_active) return;
  const int a = __ldg(active_rows + i);
  int p = 0, fresh = 0;
  if (lane == 0) {
    p = cpos[a];
    if (p < 0) {
      p = N0 + atomicAdd(count, 1);
      cpos[a] = p;
      fresh = 1;
    }
    act_c[i] = p;
    apos_c[p] = i;
  }
  p = __shfl_sync(0xffffffffu, p, 0);
  fresh = __shfl_sync(0xffffffffu, fresh, 0);
  const float4* src = reinterpret_cast<const float4*>(h_c + (size_t)(fresh ? a % N0 : p) * d);
  float4* own = reinterpret_cast<float4*>(h_c + (size_t)p * d);
  float4* pub = reinterpret_cast<float4*>(x_full + (size_t)a * d);
  for (int k = lane; k < (d >> 2); k += 32) {
    const float4 v = src[k];
    if (fresh) own[k] = v;
    pub[k] = v;
  }
}

// out[r] = h_c[cpos[r] >= 0 ? cpos[r] : r mod N0]: the union-numbered table of the last step
__global__ void __launch_bounds__(256) shared_rows_expand_kernel(const float* __restrict__ h_c, const int* __restrict__ cpos,
                                                                 int N, int N0, int d, float* __restrict__ out) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int r = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (r >= N) return;
  const int p = __ldg(cpos + r);
  const float4* src = reinterpret_cast<const float4*>(h_c + (size_t)(p >= 0 ? p : r % N0) * d);
  float4* dst = reinterpret_cast<float4*>(out + (size_t)r * d);
  for (int k = lane; k < (d >> 2); k += 32) dst[k] = src[k];
}

int shared_rows_update(const int* active_rows, int n_active, int N0, int d, int* cpos, int* count, float* h_c, float* x_full,
                       int* act_c, int* apos_c, cudaStream_t st) {
  if (!active_rows || !cpos || !count || !h_c || !x_full || !act_c || !apos_c) { set_last_error("shared_rows_update: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("shared_rows_update", d)) return e;
  if (n_active <= 0) return REGCN_OK;
  launch_k(shared_rows_update_kernel, row_grid(n_active), 256, 0, st, active_rows, n_active, N0, d, cpos, count, h_c, x_full, act_c, apos_c);
  return check_launch("shared_rows_update");
}
int shared_rows_expand(const float* h_c, const int* cpos, int N, int N0, int d, float* out, cudaStream_t st) {
  if (!h_c || !cpos || !out) { set_last_error("shared_rows_expand: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("shared_rows_expand", d)) return e;
  if (N <= 0) return REGCN_OK;
  launch_k(shared_rows_expand_kernel, row_grid(N), 256, 0, st, h_c, cpos, N, N0, d, out);
  return check_launch("shared_rows_expand");
}

// ---- K8: hyperbolic init / tangent prep / gate + radius evolution ------------------------------
struct RadiusCfg {
  float rmin, rmax, rs_cap;   // _static_radius: min(clamp(rs, rmin, rmax), 1/sqrt(c) - 1e-6)   hyperbolic_model.py:715-720
  float beta, eps_r;          // TemporalRadiusEvolution anchor_beta, epsilon                    hyperbolic_ops.py:376-388
};
__device__ __forceinline__ float static_radius(float raw, const RadiusCfg& rc) {
  return fminf(clampf_(raw, rc.rmin, rc.rmax), rc.rs_cap);
}

// h0 = apply_radius(exp_0([normalize](emb)) | project(emb), static_radius)     hyperbolic_model.py:773-782
template <int RV>
__global__ void __launch_bounds__(256) hyp_init_kernel(const float* __restrict__ emb, const float* __restrict__ radius_static,
                                                       int N, int d, int normalize, int on_manifold, Curv cv, RadiusCfg rc,
                                                       float* __restrict__ out) {
  pdl_grid_sync();
  ROW_KERNEL_PROLOGUE(N)
  WarpRow<RV> r;
  r.load_plain(emb + (size_t)row * d, nvec, lane);
  if (on_manifold) row_project(r, cv);
  else {
    if (normalize) row_l2normalize(r);
    row_exp0(r, cv);
  }
  if (radius_static) row_apply_radius(r, static_radius(__ldg(radius_static + row), rc), cv);
  r.store(out + (size_t)row * d, nvec, lane);
}

// ht = log_0(h); pt = clamp(ht, +-10); radius = max(|h|, eps)          hyperbolic_model.py:802,842-846; layers :270
template <int RV>
__global__ void __launch_bounds__(256) hyp_tangent_kernel(const float* __restrict__ h, int N, int d, Curv cv,
                                                          float* __restrict__ ht, float* __restrict__ pt,
                                                          float* __restrict__ radius, float* __restrict__ ht_hi,
                                                          float* __restrict__ ht_lo, float* __restrict__ pt_hi,
                                                          float* __restrict__ pt_lo) {
  pdl_grid_sync();
  ROW_KERNEL_PROLOGUE(N)
  WarpRow<RV> r;
  r.load_plain(h + (size_t)row * d, nvec, lane);
  if (radius) {
    float n = fmaxf(sqrtf(r.sumsq()), kEps);
    if (lane == 0) radius[row] = n;
  }
  row_log0(r, cv);
  if (ht) r.store(ht + (size_t)row * d, nvec, lane);
  if (ht_hi) r.store_split(ht_hi + (size_t)row * d, ht_lo + (size_t)row * d, nvec, lane);
  if (pt || pt_hi) {
    r.map([](float a) { return clampf_(a, -10.f, 10.f); });
    if (pt) r.store(pt + (size_t)row * d, nvec, lane);
    if (pt_hi) r.store_split(pt_hi + (size_t)row * d, pt_lo + (size_t)row * d, nvec, lane);
  }
}

// cur = project(h2); [layer_norm: cur = exp_0(normalize(log_0(cur)))]
// ct = clamp(log_0(cur)); g = s(G+b); nt = g*ct + (1-g)*pt; h = project(exp_0(nt))
// residual radius evolution (or plain apply_radius)                      hyperbolic_model.py:829-867, ops:406-425
template <int RV>
__global__ void __launch_bounds__(256) hyp_time_gate_kernel(
    const float* __restrict__ h2, const float* __restrict__ pt, const float* __restrict__ G,
    const float* __restrict__ bias, const float* __restrict__ radius_static, const float* __restrict__ rw,
    float rb, int N, int d, int layer_norm, int residual, Curv cv, RadiusCfg rc, float* __restrict__ out) {
  pdl_grid_sync();
  ROW_KERNEL_PROLOGUE(N)
  WarpRow<RV> cur, p, g, b;
  cur.load_plain(h2 + (size_t)row * d, nvec, lane);
  row_project(cur, cv);
  if (layer_norm) { row_log0(cur, cv); row_l2normalize(cur); row_exp0(cur, cv); }
  row_log0(cur, cv);
  cur.map([](float a) { return clampf_(a, -10.f, 10.f); });
  p.load_plain(pt + (size_t)row * d, nvec, lane);
  g.load_plain(G + (size_t)row * d, nvec, lane);
  b.load(bias, nvec, lane);
  g.zip(b, [](float a, float bb) { return sigmoidf_(a + bb); });
  cur.zip(g, [](float a, float s) { return s * a; });
  p.zip(g, [](float a, float s) { return (1.0f - s) * a; });
  cur.zip(p, [](float a, float bb) { return a + bb; });
  row_exp0(cur, cv);
  row_project(cur, cv);
  const float rs = static_radius(__ldg(radius_static + row), rc);
  if (residual) {
    WarpRow<RV> t = cur, w;
    const float dyn = fmaxf(sqrtf(cur.sumsq()), kEps);
    row_log0(t, cv);
    w.load(rw, nvec, lane);
    float delta = t.dot(w) + rb;
    delta = clampf_(delta, -rc.eps_r, rc.eps_r);
    const float base = rc.beta * rs + (1.0f - rc.beta) * dyn;
    row_apply_radius(cur, base + delta, cv);
  } else {
    row_apply_radius(cur, rs, cv);
  }
  cur.store(out + (size_t)row * d, nvec, lane);
}

static RadiusCfg make_rc(double c, float rmin, float rmax, float beta, float eps_r) {
  RadiusCfg rc;
  rc.rmin = rmin; rc.rmax = rmax; rc.rs_cap = (float)(1.0 / sqrt(c) - 1e-6); rc.beta = beta; rc.eps_r = eps_r;
  return rc;
}

int hyp_init(const float* emb, const float* radius_static, int N, int d, int normalize, int on_manifold, double c,
             float rmin, float rmax, float* out, cudaStream_t st) {
  if (!emb || !out) { set_last_error("hyp_init: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("hyp_init", d)) return e;
  if (!(c > 0)) { set_last_error("hyp_init: curvature must be > 0"); return REGCN_ERR_DIM; }
  Curv cv = make_curv(c);
  RadiusCfg rc = make_rc(c, rmin, rmax, 1.f, 0.f);
  if (d <= 128) launch_k(hyp_init_kernel<1>, row_grid(N), 256, 0, st, emb, radius_static, N, d, normalize, on_manifold, cv, rc, out);
  else launch_k(hyp_init_kernel<2>, row_grid(N), 256, 0, st, emb, radius_static, N, d, normalize, on_manifold, cv, rc, out);
  return check_launch("hyp_init");
}

int hyp_tangent(const float* h, int N, int d, double c, float* ht, float* pt, float* radius, float* ht_hi,
                float* ht_lo, float* pt_hi, float* pt_lo, cudaStream_t st) {
  if (!h) { set_last_error("hyp_tangent: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("hyp_tangent", d)) return e;
  if (!(c > 0)) { set_last_error("hyp_tangent: curvature must be > 0"); return REGCN_ERR_DIM; }
  Curv cv = make_curv(c);
  if (d <= 128) launch_k(hyp_tangent_kernel<1>, row_grid(N), 256, 0, st, h, N, d, cv, ht, pt, radius, ht_hi, ht_lo, pt_hi, pt_lo);
  else launch_k(hyp_tangent_kernel<2>, row_grid(N), 256, 0, st, h, N, d, cv, ht, pt, radius, ht_hi, ht_lo, pt_hi, pt_lo);
  return check_launch("hyp_tangent");
}

int hyp_time_gate(const float* h2, const float* pt, const float* G, const float* bias, const float* radius_static,
                  const float* rw, float rb, int N, int d, int layer_norm, int residual, double c, float rmin,
                  float rmax, float beta, float eps_r, float* out, cudaStream_t st) {
  if (!h2 || !pt || !G || !bias || !radius_static || !out || (residual && !rw)) { set_last_error("hyp_time_gate: null pointer"); return REGCN_ERR_NULL; }
  if (int e = check_d("hyp_time_gate", d)) return e;
  if (!(c > 0)) { set_last_error("hyp_time_gate: curvature must be > 0"); return REGCN_ERR_DIM; }
  Curv cv = make_curv(c);
  RadiusCfg rc = make_rc(c, rmin, rmax, beta, eps_r);
  if (d <= 128) launch_k(hyp_time_gate_kernel<1>, row_grid(N), 256, 0, st, h2, pt, G, bias, radius_static, rw, rb, N, d, layer_norm, residual, cv, rc, out);
  else launch_k(hyp_time_gate_kernel<2>, row_grid(N), 256, 0, st, h2, pt, G, bias, radius_static, rw, rb, N, d, layer_norm, residual, cv, rc, out);
  return check_launch("hyp_time_gate");
}

}  // namespace regcn
