// Shared device helpers for the regcn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define REGCN_OK 0
#define REGCN_ERR_NULL (-1)
#define REGCN_ERR_DIM (-2)
#define REGCN_ERR_WORKSPACE (-3)
#define REGCN_ERR_UNSUPPORTED (-4)

namespace regcn {

void set_last_error(const char* fmt, ...);
int check_launch(const char* what);

// Opt-in per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).
// slot 0 = tensor-core GEMM (work = algorithmic flops), slot 1 = union aggregate (work = algorithmic bytes).
enum { PROF_GEMM_TC = 0, PROF_AGGREGATE = 1, PROF_NUM_SLOTS = 2 };
bool prof_on();
void prof_begin(int slot, cudaStream_t st);
void prof_end(int slot, double work, cudaStream_t st);

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// The hot path is a chain of ~100 short kernels per evaluated timestamp; a plain stream serialises them with a
// ~2 us launch gap each.  Every kernel here starts with pdl_grid_sync(): it first lets the NEXT kernel of the stream
// be scheduled (its CTAs become resident as ours retire and park at their own wait), then waits until the PREVIOUS
// kernel has completed and flushed its memory.  Kernels are launched through launch_k() with the
// programmatic-stream-serialization attribute; correctness never depends on the overlap (the wait is a full grid
// dependency), only the launch latency is hidden.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
// per-thread opt-out for the next launches: a persistent kernel launched early parks one CTA per SM at its wait and
// would take the SMs the two-stream schedule leaves free for the side stream
void pdl_suppress(bool on);
bool pdl_suppressed();
void count_kernel_launch();     // feeds regcn_kernel_launches() (bench.py reports it as gpu_launches)
template <typename... P, typename... A>
inline void launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && !pdl_suppressed()) ? 1 : 0;
  count_kernel_launch();
  cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

constexpr int kWarp = 32;
constexpr float kEps = 1e-6f;                   // HyperbolicOps.EPS (hyperbolic_ops.py:28)
constexpr float kRReluSlope = (1.0f / 8.0f + 1.0f / 3.0f) * 0.5f;  // F.rrelu eval slope, 11/48

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float rreluf_(float x) { return x >= 0.f ? x : x * kRReluSlope; }
__device__ __forceinline__ float clampf_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 acc) {
  return make_float4(fmaf(s, a.x, acc.x), fmaf(s, a.y, acc.y), fmaf(s, a.z, acc.z), fmaf(s, a.w, acc.w));
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// TF32 operand split: hi = rna_tf32(a), lo = rna_tf32(a - hi).  rna (round to nearest, ties away from zero, to the 10-bit
// mantissa) is done on the bit pattern -- add half an ulp to the magnitude, clear the low 13 bits -- which is what
// cvt.rna.tf32.f32 returns for every finite input (and keeps Inf / NaN); the PTX instruction costs 4 SASS instructions
// for its special cases, 12 per element for the split, this costs 5.
__device__ __forceinline__ float rna_tf32(float a) { return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32_1(float a, float& hi, float& lo) {
  hi = rna_tf32(a);
  lo = rna_tf32(a - hi);
}

// ---------------------------------------------------------------------------
// A row of d floats (d % 4 == 0, d <= 4*32*RV) held by one warp: lane l owns the
// float4 chunks l, l+32, ...  RV = 2 covers d <= 256 (d = 200 -> 50 chunks).
// ---------------------------------------------------------------------------
template <int RV>
struct WarpRow {
  float4 v[RV];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < RV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void load(const float* row, int nvec, int lane) {
#pragma unroll
    for (int i = 0; i < RV; ++i) {
      int c = lane + i * kWarp;
      v[i] = c < nvec ? ldg4(row + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void load_plain(const float* row, int nvec, int lane) {
#pragma unroll
    for (int i = 0; i < RV; ++i) {
      int c = lane + i * kWarp;
      v[i] = c < nvec ? *reinterpret_cast<const float4*>(row + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void store(float* row, int nvec, int lane) const {
#pragma unroll
    for (int i = 0; i < RV; ++i) {
      int c = lane + i * kWarp;
      if (c < nvec) st4(row + 4 * c, v[i]);
    }
  }
  // TF32 operand split for the tensor-core GEMMs: hi = rna_tf32(x), lo = rna_tf32(x - hi)
  __device__ __forceinline__ void store_split(float* hi, float* lo, int nvec, int lane) const {
#pragma unroll
    for (int i = 0; i < RV; ++i) {
      int c = lane + i * kWarp;
      if (c < nvec) {
        float4 h, l;
        split_tf32_1(v[i].x, h.x, l.x); split_tf32_1(v[i].y, h.y, l.y);
        split_tf32_1(v[i].z, h.z, l.z); split_tf32_1(v[i].w, h.w, l.w);
        st4(hi + 4 * c, h);
        st4(lo + 4 * c, l);
      }
    }
  }
  __device__ __forceinline__ float sumsq() const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < RV; ++i) s += f4_dot(v[i], v[i]);
    return warp_sum(s);
  }
  __device__ __forceinline__ float dot(const WarpRow& o) const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < RV; ++i) s += f4_dot(v[i], o.v[i]);
    return warp_sum(s);
  }
  __device__ __forceinline__ void scale(float s) {
#pragma unroll
    for (int i = 0; i < RV; ++i) v[i] = f4_scale(v[i], s);
  }
  template <typename F>
  __device__ __forceinline__ void map(F f) {
#pragma unroll
    for (int i = 0; i < RV; ++i) {
      v[i].x = f(v[i].x); v[i].y = f(v[i].y); v[i].z = f(v[i].z); v[i].w = f(v[i].w);
    }
  }
  template <typename F>
  __device__ __forceinline__ void zip(const WarpRow& o, F f) {
#pragma unroll
    for (int i = 0; i < RV; ++i) {
      v[i].x = f(v[i].x, o.v[i].x); v[i].y = f(v[i].y, o.v[i].y);
      v[i].z = f(v[i].z, o.v[i].z); v[i].w = f(v[i].w, o.v[i].w);
    }
  }
};

// ---------------------------------------------------------------------------
// Poincare-ball row maps, restating hyperbolic_ops.py:38-233 exactly
// (same clamp order, same epsilons).  All take sqrt_c and c as fp32.
// ---------------------------------------------------------------------------
struct Curv {
  float c, sqrt_c;
  float proj_max;    // (1/sqrt(c) - eps) - eps : clamp bound inside clamp_norm via project_to_ball (:51-53,:72-74)
  float radius_max;  // 1/sqrt(c) - eps         : apply_radius upper clamp (:229-230)
};
// Bounds are formed in double like the reference's python-float arithmetic, then rounded to fp32.
inline Curv make_curv(double c) {
  Curv k;
  k.c = (float)c;
  k.sqrt_c = (float)sqrt(c);
  k.proj_max = (float)(1.0 / sqrt(c) - 1e-6 - 1e-6);
  k.radius_max = (float)(1.0 / sqrt(c) - 1e-6);
  return k;
}

// clamp_norm(x, max_norm): n = max(|x|, eps); x * (min(n, max_norm - eps) / n)   (:51-53)
template <int RV>
__device__ __forceinline__ void row_project(WarpRow<RV>& x, const Curv& k) {
  float n = fmaxf(sqrtf(x.sumsq()), kEps);
  float cn = fminf(n, k.proj_max);
  x.scale(cn / n);
}
// exp_map_zero (:91-95): project( tanh(sqrt_c*n) * (v/n) / sqrt_c )
template <int RV>
__device__ __forceinline__ void row_exp0(WarpRow<RV>& v, const Curv& k) {
  float n = fmaxf(sqrtf(v.sumsq()), kEps);
  float t = tanhf(k.sqrt_c * n);
  const float sc = k.sqrt_c;
  v.map([=](float a) { return t * (a / n) / sc; });
  row_project(v, k);
}
// log_map_zero (:112-116): atanh(min(sqrt_c*n, 1-eps)) * x / (sqrt_c*n)
template <int RV>
__device__ __forceinline__ void row_log0(WarpRow<RV>& x, const Curv& k) {
  float n = fmaxf(sqrtf(x.sumsq()), kEps);
  float sn = fminf(k.sqrt_c * n, 1.0f - kEps);
  float a = atanhf(sn);
  float den = k.sqrt_c * n;
  x.map([=](float e) { return a * e / den; });
}
// apply_radius (:222-233): x / max(|x|,eps) * clamp(r, eps, 1/sqrt(c) - eps)
template <int RV>
__device__ __forceinline__ void row_apply_radius(WarpRow<RV>& x, float r, const Curv& k) {
  float n = fmaxf(sqrtf(x.sumsq()), kEps);
  float rr = clampf_(r, kEps, k.radius_max);
  x.map([=](float e) { return e / n * rr; });
}
// F.normalize: x / max(|x|, 1e-12)
template <int RV>
__device__ __forceinline__ void row_l2normalize(WarpRow<RV>& x) {
  float n = fmaxf(sqrtf(x.sumsq()), 1e-12f);
  x.map([=](float e) { return e / n; });
}

// K13 score from a dot product, written with explicit IEEE operations (no FMA contraction) so that the dense
// epilogue kernel, the fused count epilogue and the pair-score pass produce bit-identical values.
__device__ __forceinline__ float hyp_score_from_dot(float dot, float x_sq, float y_sq, float c, float proj_max,
                                                    float scale, float margin) {
  const float xy = -dot;
  const float two_c_xy = __fmul_rn(__fmul_rn(2.0f, c), xy);
  const float a = __fadd_rn(__fadd_rn(1.0f, two_c_xy), __fmul_rn(c, y_sq));
  const float b = __fsub_rn(1.0f, __fmul_rn(c, x_sq));
  const float t1 = __fmul_rn(__fmul_rn(a, a), x_sq);
  const float t2 = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, a), b), xy);
  const float t3 = __fmul_rn(__fmul_rn(b, b), y_sq);
  const float num_sq = fmaxf(__fadd_rn(__fadd_rn(t1, t2), t3), 0.f);
  const float den = __fadd_rn(__fadd_rn(__fadd_rn(1.0f, two_c_xy), __fmul_rn(__fmul_rn(__fmul_rn(c, c), x_sq), y_sq)), kEps);
  float n = __fdiv_rn(__fsqrt_rn(num_sq), fabsf(den));
  n = fminf(n, proj_max);
  return __fmul_rn(scale, __fsub_rn(margin, __fmul_rn(n, n)));
}

// K13, true-distance branch (hyperbolic_decoder.py:145-163: use_hyperbolic_distance with a per-query curvature c_q):
//   score = scale * (margin - 2/(sqrt(c_q)+eps) * atanh(min(sqrt(c_q) * n, 1-eps))),   n = |(-q) (+)_{c_q} e|
// from the same three scalars <q,e>, |q|^2, |e|^2.  sqrt(c_q) is sqrt(c_q + eps) like the reference (:148).
__device__ __forceinline__ float hyp_dist_score_from_dot(float dot, float x_sq, float y_sq, float cq, float scale,
                                                         float margin) {
  const float sqrt_c = __fsqrt_rn(__fadd_rn(cq, kEps));
  const float two_c_xy = __fmul_rn(__fmul_rn(2.0f, cq), dot);
  const float a = __fadd_rn(__fsub_rn(1.0f, two_c_xy), __fmul_rn(cq, y_sq));
  const float b = __fsub_rn(1.0f, __fmul_rn(cq, x_sq));
  // |a*(-q) + b*e|^2 = a^2 x_sq - 2ab<q,e> + b^2 y_sq
  const float t1 = __fmul_rn(__fmul_rn(a, a), x_sq);
  const float t2 = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, a), b), dot);
  const float t3 = __fmul_rn(__fmul_rn(b, b), y_sq);
  const float num_sq = fmaxf(__fadd_rn(__fsub_rn(t1, t2), t3), 0.f);
  const float den = __fadd_rn(__fadd_rn(__fsub_rn(1.0f, two_c_xy), __fmul_rn(__fmul_rn(__fmul_rn(cq, cq), x_sq), y_sq)), kEps);
  float n = fmaxf(__fdiv_rn(__fsqrt_rn(num_sq), fabsf(den)), kEps);
  const float inv = __fadd_rn(sqrt_c, kEps);
  n = fminf(n, __fsub_rn(__fdiv_rn(1.0f, inv), kEps));
  const float arg = fminf(__fmul_rn(sqrt_c, n), 0.999999f);
  const float dist = __fmul_rn(__fdiv_rn(2.0f, inv), atanhf(arg));
  return __fmul_rn(scale, __fsub_rn(margin, dist));
}

// Rank contribution of candidate j (score s) against target t (score st): stable-sort position rule (rank.cu).
__device__ __forceinline__ int rank_beats(float s, int j, float st, int t) {
  return (s > st || (s == st && j < t)) ? 1 : 0;
}

}  // namespace regcn
