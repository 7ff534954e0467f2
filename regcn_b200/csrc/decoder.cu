// Decoder-side kernels.
//   K10 ConvTransE / ConvTransR query tower   src/decoder.py:29-52,78-95; hyperbolic_decoder.py:376-406,478-507
//   K12 RotH / MuRP query builder             hyperbolic_decoder.py:1064-1086 (RotH), :744-765 (MuRP), :1243-1251 (RotHRel)
//   K13 hyperbolic score epilogue             hyperbolic_decoder.py:89-179 (proxy-distance branch) in norm/dot form
// The dense contractions between these stages (FC 50d->d, the d->d projections, Q.E^T) run on the GEMM kernels.
#include "common.cuh"

namespace regcn {

// ---- K10a: feature map  F[b, c*d + i] = relu(bn1(conv1d(bn0([x0;x1]))))  (eval-mode BatchNorm) --------
// x0 = ent[idx0[b]], x1 = second[idx1[b]] (relation row for ConvTransE, object row for ConvTransR).
// bn params are pre-folded on the host into scale/shift: y = x*scale + shift.
__global__ void __launch_bounds__(256) convtranse_features_kernel(
    const float* __restrict__ ent, const float* __restrict__ second, const int64_t* __restrict__ triples,
    int col0, int col1, int B, int d, int C, int ksz,
    const float* __restrict__ bn0_scale, const float* __restrict__ bn0_shift,   // (2)
    const float* __restrict__ conv_w, const float* __restrict__ conv_b,         // (C,2,ksz), (C)
    const float* __restrict__ bn1_scale, const float* __restrict__ bn1_shift,   // (C)
    float* __restrict__ F, float* __restrict__ F_hi, float* __restrict__ F_lo) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  const int pad = ksz / 2;
  const int ld = d + 2 * pad;
  float* y0 = sm;            // padded, bn0 applied
  float* y1 = sm + ld;
  float* wsm = sm + 2 * ld;  // C*2*ksz weights, then C conv bias, C bn1 scale, C bn1 shift
  const int b = blockIdx.x;
  const int64_t i0 = triples[3 * (size_t)b + col0];
  const int64_t i1 = triples[3 * (size_t)b + col1];
  const float s0 = bn0_scale[0], t0 = bn0_shift[0], s1 = bn0_scale[1], t1 = bn0_shift[1];
  for (int i = threadIdx.x; i < ld; i += blockDim.x) {
    const int j = i - pad;
    const bool in = j >= 0 && j < d;
    y0[i] = in ? fmaf(__ldg(ent + (size_t)i0 * d + j), s0, t0) : 0.f;
    y1[i] = in ? fmaf(__ldg(second + (size_t)i1 * d + j), s1, t1) : 0.f;
  }
  const int nw = C * 2 * ksz;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) wsm[i] = conv_w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    wsm[nw + i] = conv_b[i];
    wsm[nw + C + i] = bn1_scale[i];
    wsm[nw + 2 * C + i] = bn1_shift[i];
  }
  __syncthreads();
  const size_t fb = (size_t)b * C * d;
  if ((d & 3) == 0 && ksz <= 5) {
    // thread <-> (4 consecutive positions, a slice of the channels): the padded inputs of the 4 positions stay in
    // registers, per-channel weights are shared-memory broadcasts, and every store is a 16-byte vector, coalesced along i
    const int ngrp = d >> 2;                                   // position groups per channel row
    const int nslice = max(1, (int)blockDim.x / ngrp);         // channel slices handled concurrently
    const int pg = threadIdx.x % ngrp, cs = threadIdx.x / ngrp;
    if (cs < nslice) {
      float a0[8], a1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool ok = k < ksz + 3;
        a0[k] = ok ? y0[4 * pg + k] : 0.f;
        a1[k] = ok ? y1[4 * pg + k] : 0.f;
      }
      for (int c = cs; c < C; c += nslice) {
        const float* w = wsm + c * 2 * ksz;
        const float cb = wsm[nw + c], bs = wsm[nw + C + c], bt = wsm[nw + 2 * C + c];
        float o4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float acc = cb;
#pragma unroll
          for (int k = 0; k < 5; ++k) if (k < ksz) acc = fmaf(w[k], a0[u + k], acc);
#pragma unroll
          for (int k = 0; k < 5; ++k) if (k < ksz) acc = fmaf(w[ksz + k], a1[u + k], acc);
          o4[u] = fmaxf(fmaf(acc, bs, bt), 0.f);
        }
        const size_t o = fb + (size_t)c * d + 4 * pg;
        if (F) st4(F + o, make_float4(o4[0], o4[1], o4[2], o4[3]));
        if (F_hi) {
          float4 h, l;
          split_tf32_1(o4[0], h.x, l.x); split_tf32_1(o4[1], h.y, l.y);
          split_tf32_1(o4[2], h.z, l.z); split_tf32_1(o4[3], h.w, l.w);
          st4(F_hi + o, h);
          st4(F_lo + o, l);
        }
      }
    }
    return;
  }
  // generic shape: thread <-> position i (fixed), loop over channels
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    float a0[7], a1[7];
    for (int k = 0; k < ksz && k < 7; ++k) { a0[k] = y0[i + k]; a1[k] = y1[i + k]; }
    for (int c = 0; c < C; ++c) {
      const float* w = wsm + c * 2 * ksz;
      float acc = wsm[nw + c];
      for (int k = 0; k < ksz; ++k) acc = fmaf(w[k], a0[k], acc);
      for (int k = 0; k < ksz; ++k) acc = fmaf(w[ksz + k], a1[k], acc);
      acc = fmaf(acc, wsm[nw + C + c], wsm[nw + 2 * C + c]);
      acc = fmaxf(acc, 0.f);
      const size_t o = fb + (size_t)c * d + i;
      if (F) F[o] = acc;
      if (F_hi) { float h, l; split_tf32_1(acc, h, l); F_hi[o] = h; F_lo[o] = l; }
    }
  }
}

int convtranse_features(const float* ent, const float* second, const int64_t* triples, int col0, int col1, int B,
                        int d, int C, int ksz, const float* bn0_scale, const float* bn0_shift, const float* conv_w,
                        const float* conv_b, const float* bn1_scale, const float* bn1_shift, float* F, float* F_hi,
                        float* F_lo, cudaStream_t st) {
  if (!ent || !second || !triples || !bn0_scale || !bn0_shift || !conv_w || !conv_b || !bn1_scale || !bn1_shift ||
      (!F && !F_hi) || (F_hi && !F_lo)) {
    set_last_error("convtranse_features: null pointer"); return REGCN_ERR_NULL;
  }
  if (B <= 0) return REGCN_OK;
  if (d <= 0 || C <= 0 || ksz <= 0 || !(ksz & 1) || ksz > 7) { set_last_error("convtranse_features: bad dims d=%d C=%d k=%d", d, C, ksz); return REGCN_ERR_DIM; }
  const size_t smem = ((size_t)2 * (d + 2 * (ksz / 2)) + (size_t)C * 2 * ksz + 3 * (size_t)C) * sizeof(float);
  if (smem > 48 * 1024) { set_last_error("convtranse_features: shared memory %zu too large", smem); return REGCN_ERR_UNSUPPORTED; }
  launch_k(convtranse_features_kernel, B, 256, smem, st, ent, second, triples, col0, col1, B, d, C, ksz, bn0_scale, bn0_shift,
                                                   conv_w, conv_b, bn1_scale, bn1_shift, F, F_hi, F_lo);
  return check_launch("convtranse_features");
}

// ---- K10b: x = relu(x*scale + shift) per feature (bn2 folded; scale==null -> plain relu, the B==1 case) ----
__global__ void affine_relu_kernel(float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                   size_t total, int d, int relu) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i % d);
  float v = x[i];
  if (scale) v = fmaf(v, __ldg(scale + j), __ldg(shift + j));
  x[i] = relu ? fmaxf(v, 0.f) : v;
}

int affine_relu(float* x, const float* scale, const float* shift, int M, int d, int relu, cudaStream_t st) {
  if (!x || (scale && !shift)) { set_last_error("affine_relu: null pointer"); return REGCN_ERR_NULL; }
  const size_t total = (size_t)M * d;
  if (!total) return REGCN_OK;
  launch_k(affine_relu_kernel, (unsigned)((total + 255) / 256), 256, 0, st, x, scale, shift, total, d, relu);
  return check_launch("affine_relu");
}

// ---- K12a: gather + tangent map of the query subjects:  out[b] = log_0(project?(E[idx[b]])) ----
template <int RV>
__global__ void __launch_bounds__(256) gather_log0_kernel(const float* __restrict__ E, const int64_t* __restrict__ triples,
                                                          int col, int B, int d, int project, Curv cv, float* __restrict__ out) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= B) return;
  const int nvec = d >> 2;
  WarpRow<RV> r;
  r.load(E + (size_t)triples[3 * (size_t)row + col] * d, nvec, lane);
  if (project) row_project(r, cv);
  row_log0(r, cv);
  r.store(out + (size_t)row * d, nvec, lane);
}

int gather_log0(const float* E, const int64_t* triples, int col, int B, int d, int project, double c, float* out, cudaStream_t st) {
  if (!E || !triples || !out) { set_last_error("gather_log0: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || d > 256) { set_last_error("gather_log0: d=%d unsupported", d); return REGCN_ERR_UNSUPPORTED; }
  if (B <= 0) return REGCN_OK;
  Curv cv = make_curv(c);
  const unsigned grid = (unsigned)(((size_t)B * 32 + 255) / 256);
  if (d <= 128) launch_k(gather_log0_kernel<1>, grid, 256, 0, st, E, triples, col, B, d, project, cv, out);
  else launch_k(gather_log0_kernel<2>, grid, 256, 0, st, E, triples, col, B, d, project, cv, out);
  return check_launch("gather_log0");
}

// ---- K12b: query = project(exp_0(rot(s_tan))) (+)_c project(exp_0(v_r)) ----------------------------------
// kind 0 RotH   : rot = Givens(s_tan[b], ang[r_b])   ang (2R, d/2), trans (2R, d)       hyperbolic_decoder.py:1074-1085
// kind 1 MuRP   : rot = diag[r_b] * s_tan[b]         ang = diag (2R, d)                  :751-764
// kind 2 RotHRel: rot = Givens(s_tan[b], global_rot (d/2)); query = (-exp_0(rot)) (+)_c E[o_b]   :1247-1251
// Pair (x_{2i}, x_{2i+1}) lives inside one float4, so the rotation needs no shuffles.
__device__ __forceinline__ float4 givens4(float4 x, float a0, float a1) {
  float s0, c0, s1, c1;
  sincosf(a0, &s0, &c0);
  sincosf(a1, &s1, &c1);
  return make_float4(c0 * x.x - s0 * x.y, s0 * x.x + c0 * x.y, c1 * x.z - s1 * x.w, s1 * x.z + c1 * x.w);
}

template <int RV>
__device__ __forceinline__ void row_mobius_add(WarpRow<RV>& x, const WarpRow<RV>& y, const Curv& cv) {
  // mobius_add (hyperbolic_ops.py:135-143)
  const float x_sq = x.sumsq(), y_sq = y.sumsq(), xy = x.dot(y);
  const float c = cv.c;
  const float a = 1.0f + 2.0f * c * xy + c * y_sq;
  const float b = 1.0f - c * x_sq;
  const float den = 1.0f + 2.0f * c * xy + c * c * x_sq * y_sq + kEps;
  x.zip(y, [=](float xx, float yy) { return (a * xx + b * yy) / den; });
  row_project(x, cv);
}

template <int RV>
__global__ void __launch_bounds__(256) hyp_query_kernel(
    const float* __restrict__ s_tan, const float* __restrict__ ang, const float* __restrict__ trans,
    const float* __restrict__ E, const int64_t* __restrict__ triples, int B, int d, int kind, Curv cv,
    float* __restrict__ Q, float* __restrict__ q_sumsq) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= B) return;
  const int nvec = d >> 2;
  const int64_t r = triples[3 * (size_t)row + 1];
  WarpRow<RV> x, y;
  x.load_plain(s_tan + (size_t)row * d, nvec, lane);
#pragma unroll
  for (int i = 0; i < RV; ++i) {
    const int cidx = lane + i * kWarp;
    if (cidx < nvec) {
      if (kind == 1) {
        float4 dg = ldg4(ang + (size_t)r * d + 4 * cidx);
        x.v[i] = make_float4(dg.x * x.v[i].x, dg.y * x.v[i].y, dg.z * x.v[i].z, dg.w * x.v[i].w);
      } else {
        const float* ap = kind == 0 ? ang + (size_t)r * (d / 2) + 2 * cidx : ang + 2 * cidx;
        x.v[i] = givens4(x.v[i], __ldg(ap), __ldg(ap + 1));
      }
    }
  }
  row_exp0(x, cv);
  if (kind == 2) {
    x.map([](float a) { return -a; });
    y.load(E + (size_t)triples[3 * (size_t)row + 2] * d, nvec, lane);
  } else {
    row_project(x, cv);
    y.load(trans + (size_t)r * d, nvec, lane);
    row_exp0(y, cv);
    row_project(y, cv);
  }
  row_mobius_add(x, y, cv);
  x.store(Q + (size_t)row * d, nvec, lane);
  if (q_sumsq) {
    const float s = x.sumsq();
    if (lane == 0) q_sumsq[row] = s;
  }
}

int hyp_query(const float* s_tan, const float* ang, const float* trans, const float* E, const int64_t* triples, int B,
              int d, int kind, double c, float* Q, float* q_sumsq, cudaStream_t st) {
  if (!s_tan || !ang || !triples || !Q || (kind != 2 && !trans) || (kind == 2 && !E)) { set_last_error("hyp_query: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || d > 256 || kind < 0 || kind > 2) { set_last_error("hyp_query: d=%d kind=%d unsupported", d, kind); return REGCN_ERR_UNSUPPORTED; }
  if (B <= 0) return REGCN_OK;
  Curv cv = make_curv(c);
  const unsigned grid = (unsigned)(((size_t)B * 32 + 255) / 256);
  if (d <= 128) launch_k(hyp_query_kernel<1>, grid, 256, 0, st, s_tan, ang, trans, E, triples, B, d, kind, cv, Q, q_sumsq);
  else launch_k(hyp_query_kernel<2>, grid, 256, 0, st, s_tan, ang, trans, E, triples, B, d, kind, cv, Q, q_sumsq);
  return check_launch("hyp_query");
}

// ---- K13: score epilogue on a dense dot-product matrix -------------------------------------------------
// In:  S[b,n] = <q_b, e_n>.  Out: scale*(margin - |project((-q_b) (+)_c e_n)|^2) + bias[n] + qbias[b]
// with x = -q: x_sq = |q|^2, y_sq = |e|^2, xy = -<q,e>; A = 1+2c*xy+c*y_sq; Bc = 1-c*x_sq;
// |num|^2 = A^2 x_sq + 2 A Bc xy + Bc^2 y_sq; den = 1+2c*xy+c^2 x_sq y_sq + eps; n = sqrt(|num|^2)/den,
// clamp_norm -> min(n, proj_max) (for n >= eps).                       hyperbolic_decoder.py:164-172, ops:135-143
__global__ void hyp_score_epilogue_kernel(float* __restrict__ S, size_t ld, int B, int N, const float* __restrict__ q_sumsq,
                                          const float* __restrict__ e_sumsq, const float* __restrict__ bias,
                                          const float* __restrict__ qbias, Curv cv, const float* __restrict__ scale_margin,
                                          const float* __restrict__ row_c) {
  pdl_grid_sync();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= N) return;
  const float scale = scale_margin[0], margin = scale_margin[1];
  const float dot = S[(size_t)b * ld + n];
  float v = row_c ? hyp_dist_score_from_dot(dot, __ldg(q_sumsq + b), __ldg(e_sumsq + n), __ldg(row_c + b), scale, margin)
                  : hyp_score_from_dot(dot, __ldg(q_sumsq + b), __ldg(e_sumsq + n), cv.c, cv.proj_max, scale, margin);
  if (bias) v = __fadd_rn(v, __ldg(bias + n));
  if (qbias) v = __fadd_rn(v, __ldg(qbias + b));
  S[(size_t)b * ld + n] = v;
}

int hyp_score_epilogue(float* S, int ld, int B, int N, const float* q_sumsq, const float* e_sumsq, const float* bias,
                       const float* qbias, double c, const float* scale_margin, const float* row_c, cudaStream_t st) {
  if (!S || !q_sumsq || !e_sumsq || !scale_margin) { set_last_error("hyp_score_epilogue: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0 || N <= 0) return REGCN_OK;
  if (B > 65535) { set_last_error("hyp_score_epilogue: B=%d > 65535 (chunk the queries)", B); return REGCN_ERR_DIM; }
  Curv cv = make_curv(c);
  dim3 grid((N + 255) / 256, B);
  launch_k(hyp_score_epilogue_kernel, grid, 256, 0, st, S, (size_t)ld, B, N, q_sumsq, e_sumsq, bias, qbias, cv, scale_margin, row_c);
  return check_launch("hyp_score_epilogue");
}


// Per-query curvature of the relation-specific-curvature decoders (hyperbolic_decoder.py:66-86,1020-1026):
// c_q = max(1e-5, min(softplus(raw[r mod R]), 0.999 * c, cmax)).
__global__ void rel_curvature_kernel(const float* __restrict__ raw, const int64_t* __restrict__ triples, int B, int R,
                                     float upper, float* __restrict__ out) {
  pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int r = (int)(triples[3 * (size_t)b + 1] % R);
  const float x = __ldg(raw + r);
  const float sp = x > 20.f ? x : log1pf(expf(x));       // F.softplus (beta 1, threshold 20)
  out[b] = fmaxf(fminf(sp, upper), 1e-5f);
}

int rel_curvature(const float* raw, const int64_t* triples, int B, int R, double c, double cmax, float* out,
                  cudaStream_t st) {
  if (!raw || !triples || !out) { set_last_error("rel_curvature: null pointer"); return REGCN_ERR_NULL; }
  if (R <= 0) { set_last_error("rel_curvature: R=%d", R); return REGCN_ERR_DIM; }
  if (B <= 0) return REGCN_OK;
  // upper = min(0.999 * c, cmax) formed in fp32 like the reference's tensors (new_tensor(float(...)))
  float upper = 0.999f * (float)c;
  if (cmax > 0 && (float)cmax < upper) upper = (float)cmax;
  launch_k(rel_curvature_kernel, (B + 255) / 256, 256, 0, st, raw, triples, B, R, upper, out);
  return check_launch("rel_curvature");
}


// ---- AttH / AttHRel query builder (hyperbolic_decoder.py:1283-1512, 1515-1700) ------------------------------------
// mode 0 (HyperbolicAttH.forward :1403-1480): per-relation tables rot, ref (2R, d/2), attn (2R, 2d), rel (2R, d), trans (2R, d)
//   a = s(<attn[r,:d], x> + <attn[r,d:], rel[r]>);  m = a Rot(x) + (1-a) Ref(x);  q = project(exp_0(m)) (+) project(exp_0(trans[r]))
// mode 1 (HyperbolicAttHRel.forward :1593-1640): global rot, ref (d/2), attn (2d); o = E[o_b]
//   a = s(<attn[:d], x> + <attn[d:], log_0(o)>);  q = (-exp_0(m)) (+) o
__device__ __forceinline__ float4 givens_ref4(float4 x, float a0, float a1) {
  float s0, c0, s1, c1;
  sincosf(a0, &s0, &c0);
  sincosf(a1, &s1, &c1);
  return make_float4(c0 * x.x + s0 * x.y, s0 * x.x - c0 * x.y, c1 * x.z + s1 * x.w, s1 * x.z - c1 * x.w);
}
template <int RV>
__global__ void __launch_bounds__(256) atth_query_kernel(
    const float* __restrict__ s_tan, const float* __restrict__ rot, const float* __restrict__ ref,
    const float* __restrict__ attn, const float* __restrict__ rel, const float* __restrict__ trans,
    const float* __restrict__ E, const int64_t* __restrict__ triples, int B, int d, int mode, Curv cv,
    float* __restrict__ Q, float* __restrict__ q_sumsq) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= B) return;
  const int nvec = d >> 2;
  const int64_t r = triples[3 * (size_t)row + 1];
  WarpRow<RV> x, y, w, u;
  x.load_plain(s_tan + (size_t)row * d, nvec, lane);
  const float* aw = mode == 0 ? attn + (size_t)r * 2 * d : attn;
  w.load(aw, nvec, lane);
  float logit = x.dot(w);
  if (mode == 0) {
    u.load(rel + (size_t)r * d, nvec, lane);
  } else {
    y.load(E + (size_t)triples[3 * (size_t)row + 2] * d, nvec, lane);
    u = y;
    row_log0(u, cv);
  }
  w.load(aw + d, nvec, lane);
  logit += u.dot(w);
  const float a = sigmoidf_(logit);
  const float* rp = mode == 0 ? rot + (size_t)r * (d / 2) : rot;
  const float* fp = mode == 0 ? ref + (size_t)r * (d / 2) : ref;
#pragma unroll
  for (int i = 0; i < RV; ++i) {
    const int cidx = lane + i * kWarp;
    if (cidx < nvec) {
      const float4 ro = givens4(x.v[i], __ldg(rp + 2 * cidx), __ldg(rp + 2 * cidx + 1));
      const float4 re = givens_ref4(x.v[i], __ldg(fp + 2 * cidx), __ldg(fp + 2 * cidx + 1));
      const float b1 = 1.0f - a;
      x.v[i] = make_float4(a * ro.x + b1 * re.x, a * ro.y + b1 * re.y, a * ro.z + b1 * re.z, a * ro.w + b1 * re.w);
    }
  }
  row_exp0(x, cv);
  if (mode == 0) {
    row_project(x, cv);
    y.load(trans + (size_t)r * d, nvec, lane);
    row_exp0(y, cv);
    row_project(y, cv);
  } else {
    x.map([](float v) { return -v; });
  }
  row_mobius_add(x, y, cv);
  x.store(Q + (size_t)row * d, nvec, lane);
  if (q_sumsq) {
    const float s = x.sumsq();
    if (lane == 0) q_sumsq[row] = s;
  }
}
int atth_query(const float* s_tan, const float* rot, const float* ref, const float* attn, const float* rel,
               const float* trans, const float* E, const int64_t* triples, int B, int d, int mode, double c, float* Q,
               float* q_sumsq, cudaStream_t st) {
  if (!s_tan || !rot || !ref || !attn || !triples || !Q || (mode == 0 && (!rel || !trans)) || (mode == 1 && !E)) { set_last_error("atth_query: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || d > 256 || mode < 0 || mode > 1) { set_last_error("atth_query: d=%d mode=%d unsupported", d, mode); return REGCN_ERR_UNSUPPORTED; }
  if (B <= 0) return REGCN_OK;
  Curv cv = make_curv(c);
  const unsigned grid = (unsigned)(((size_t)B * 32 + 255) / 256);
  if (d <= 128) launch_k(atth_query_kernel<1>, grid, 256, 0, st, s_tan, rot, ref, attn, rel, trans, E, triples, B, d, mode, cv, Q, q_sumsq);
  else launch_k(atth_query_kernel<2>, grid, 256, 0, st, s_tan, rot, ref, attn, rel, trans, E, triples, B, d, mode, cv, Q, q_sumsq);
  return check_launch("atth_query");
}
}  // namespace regcn
