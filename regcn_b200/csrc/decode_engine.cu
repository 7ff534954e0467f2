// One-call decode + rank of an evaluated timestamp for the ConvTransE / ConvTransR pair (src/rrgcn.py:190-193,
// src/decoder.py:29-52,78-100, rgcn/utils.py:136-166): what the evaluation loop issues per timestamp after the
// evolution -- [F.normalize] -> tanh(E) -> entity query tower -> fused score/count against all entities + filter
// correction -> relation query tower -> (B,2R) scores -> raw/filtered relation ranks -> four rank vectors packed for
// one device->host copy.  The same kernels, in the same order and with the same arithmetic as the per-op Python path
// (regcn_b200/evaluate.py), so the ranks are bit-identical; what disappears is ~30 Python->C round trips and as many
// allocator calls per timestamp (the loop is host-bound at ICEWS sizes).
#include "common.cuh"
#include "internal.h"
#include <stdlib.h>

namespace regcn {

__global__ void extract_target_kernel(const int64_t* __restrict__ triples, int B, int col, int* __restrict__ out) {
  pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = (int)triples[(size_t)b * 3 + col];
}
__global__ void pack_ranks_kernel(const int* __restrict__ raw, const int* __restrict__ filt, const int* __restrict__ raw_r,
                                  const int* __restrict__ filt_r, int B, int* __restrict__ out) {
  pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  out[b] = raw[b] + 1;
  out[B + b] = filt[b] + 1;
  out[2 * B + b] = raw_r[b] + 1;
  out[3 * B + b] = filt_r[b] + 1;
}

static inline size_t al256(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
// REGCN_FUSED_TOWER=0: feature map through memory (regcn_convtranse_features + FC GEMM), as regcn_b200/decoder.py reads it
static bool fused_tower() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("REGCN_FUSED_TOWER"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

struct DecodePlan {
  size_t emb_n, e_all, e_hi, e_lo, f_hi, f_lo, q, q_hi, q_lo, gemm_ws, a_hi, a_lo, b_hi, b_lo, ps, target, raw, filt,
      r_hi, r_lo, score_rel, tscore_r, raw_r, filt_r, total, gemm_ws_bytes;
  int split_k;
};
static DecodePlan plan_decode(int N, int R2, int d, int B, int C, int P) {
  DecodePlan p;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += al256(bytes); return at; };
  const size_t nd = (size_t)N * d * 4, bd = (size_t)B * d * 4, bf = (size_t)B * C * d * 4, pd = (size_t)P * d * 4;
  p.emb_n = take(nd); p.e_all = take(nd); p.e_hi = take(nd); p.e_lo = take(nd);
  p.f_hi = take(bf); p.f_lo = p.f_hi;     // (one fp32 feature matrix; f_lo kept as an alias)
  p.q = take(bd); p.q_hi = take(bd); p.q_lo = take(bd);
  // the FC GEMM's split-K choice of ConvTransE._tower (regcn_b200/decoder.py)
  int sk = (148 * 2) / (((B + 127) / 128) * ((d + 127) / 128));
  sk = sk < 1 ? 1 : (sk > 16 ? 16 : sk);
  const int kmax = (C * d) / 512 > 1 ? (C * d) / 512 : 1;
  p.split_k = sk < kmax ? sk : kmax;
  {
    const size_t a = gemm_tf32_workspace_bytes(B, d, p.split_k), b = convtrans_fc_workspace_bytes(B, d);
    p.gemm_ws_bytes = a > b ? a : b;
  }
  p.gemm_ws = take(p.gemm_ws_bytes);
  p.a_hi = take(pd); p.a_lo = take(pd); p.b_hi = take(pd); p.b_lo = take(pd);
  p.ps = take((size_t)P * 4);
  p.target = take((size_t)B * 4); p.raw = take((size_t)B * 4); p.filt = take((size_t)B * 4);
  p.r_hi = take((size_t)R2 * d * 4); p.r_lo = take((size_t)R2 * d * 4);
  p.score_rel = take((size_t)B * R2 * 4);
  p.tscore_r = take((size_t)B * 4); p.raw_r = take((size_t)B * 4); p.filt_r = take((size_t)B * 4);
  p.total = o + 256;
  return p;
}

size_t convtrans_decode_rank_workspace_bytes(int N, int R2, int d, int B, int C, int P) {
  return plan_decode(N, R2, d, B, C, P).total;
}

// tower parameter block of one decoder (device pointers): bn0 scale, shift (2) | conv weight (C,2,k), bias (C) |
// bn1 scale, shift (C) | fc weight hi, lo (d, C*d) | fc bias (d) | bn2 scale, shift (d) | fc weight hi, lo in the
// reduction order of the fused tower (regcn_convtrans_fc_pack_weight; only read when that path runs)
enum { TW_BN0_S = 0, TW_BN0_B, TW_CONV_W, TW_CONV_B, TW_BN1_S, TW_BN1_B, TW_FC_HI, TW_FC_LO, TW_FC_B, TW_BN2_S, TW_BN2_B, TW_FCZ_HI, TW_FCZ_LO, TW_NUM };

static int run_tower(const float* first, const float* second, const int64_t* triples, int col0, int col1, int B, int d,
                     int C, int ksz, const void* const* tw, const DecodePlan& pl, char* ws, int bn2, cudaStream_t st) {
  float* feat = (float*)(ws + pl.f_hi);      // the feature map as ONE fp32 matrix: the FC GEMM splits it to TF32 on chip
  float* q = (float*)(ws + pl.q);
  int e;
  if (fused_tower() && ksz == 3 && !(d & 3) && C <= 64) {
    // the feature map is computed inside the FC GEMM's operand ring (gemm_tc.cu, conv-producer mode): never written
    e = convtrans_fc(first, second, triples, col0, col1, B, d, C, ksz, (const float*)tw[TW_BN0_S], (const float*)tw[TW_BN0_B],
                     (const float*)tw[TW_CONV_W], (const float*)tw[TW_CONV_B], (const float*)tw[TW_BN1_S],
                     (const float*)tw[TW_BN1_B], (const float*)tw[TW_FCZ_HI], (const float*)tw[TW_FCZ_LO], 16 * C * ((d + 15) / 16), d,
                     (const float*)tw[TW_FC_B], q, d, (float*)(ws + pl.gemm_ws), pl.gemm_ws_bytes, st, 0,
                     bn2 ? (const float*)tw[TW_BN2_S] : nullptr, bn2 ? (const float*)tw[TW_BN2_B] : nullptr, 1,
                     (float*)(ws + pl.q_hi), (float*)(ws + pl.q_lo));
    return e;
  }
  e = convtranse_features(first, second, triples, col0, col1, B, d, C, ksz, (const float*)tw[TW_BN0_S],
                              (const float*)tw[TW_BN0_B], (const float*)tw[TW_CONV_W], (const float*)tw[TW_CONV_B],
                              (const float*)tw[TW_BN1_S], (const float*)tw[TW_BN1_B], feat, nullptr, nullptr, st);
  if (e) return e;
  e = gemm_tf32_a32(feat, C * d, C * d, nullptr, nullptr, 0, 0, nullptr, (const float*)tw[TW_FC_HI],
                    (const float*)tw[TW_FC_LO], C * d, q, d, B, d, (const float*)tw[TW_FC_B], 0, 3, pl.split_k,
                    (float*)(ws + pl.gemm_ws), gemm_tf32_workspace_bytes(B, d, pl.split_k), nullptr, 0, st);
  if (e) return e;
  e = affine_relu(q, bn2 ? (const float*)tw[TW_BN2_S] : nullptr, bn2 ? (const float*)tw[TW_BN2_B] : nullptr, B, d, 1, st);
  if (e) return e;
  return split_tf32(q, (float*)(ws + pl.q_hi), (float*)(ws + pl.q_lo), (size_t)B * d, st);
}

int convtrans_decode_rank(const float* emb, const float* r_emb, const int64_t* triples, const void* const* tower_ent,
                          const void* const* tower_rel, const int* fe_ptr, const int* fe_idx, const int* fe_end,
                          const int* pair_a, const int* pair_e, int P, const int* fr_ptr, const int* fr_idx,
                          const int* fr_end, int N, int R2, int d, int B, int C, int ksz, int layer_norm, int* packed,
                          void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (!emb || !r_emb || !triples || !tower_ent || !tower_rel || !fe_ptr || !fe_idx || !pair_a || !pair_e || !fr_ptr ||
      !fr_idx || !packed || !workspace) { set_last_error("convtrans_decode_rank: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || d > 256 || B <= 0 || N <= 0 || R2 <= 0 || P < B) { set_last_error("convtrans_decode_rank: bad dims"); return REGCN_ERR_DIM; }
  const DecodePlan pl = plan_decode(N, R2, d, B, C, P);
  if (ws_bytes < pl.total) { set_last_error("convtrans_decode_rank: workspace too small"); return REGCN_ERR_WORKSPACE; }
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  int e;
  // [F.normalize] (src/rrgcn.py:190) and tanh of the entity table (src/decoder.py:30,79), with its TF32 split
  float* e_all = (float*)(ws + pl.e_all);
  float* e_hi = (float*)(ws + pl.e_hi);
  float* e_lo = (float*)(ws + pl.e_lo);
  // one pass: row_map mode 9 = tanh(normalize(x)) (the same two steps on the row held in registers), mode 1 = tanh
  if ((e = row_map(emb, e_all, N, d, layer_norm ? 9 : 1, 1.0, nullptr, e_hi, e_lo, st))) return e;
  // ---- entity head: query tower, pair scores, counting GEMM, filter correction ------------------------------------
  if ((e = run_tower(e_all, r_emb, triples, 0, 1, B, d, C, ksz, tower_ent, pl, ws, B > 1, st))) return e;
  const float* q_hi = (const float*)(ws + pl.q_hi);
  const float* q_lo = (const float*)(ws + pl.q_lo);
  int* target = (int*)(ws + pl.target);
  int* raw = (int*)(ws + pl.raw);
  int* filt = (int*)(ws + pl.filt);
  launch_k(extract_target_kernel, (unsigned)((B + 255) / 256), 256, 0, st, triples, B, 2, target);
  float* a_hi = (float*)(ws + pl.a_hi); float* a_lo = (float*)(ws + pl.a_lo);
  float* b_hi = (float*)(ws + pl.b_hi); float* b_lo = (float*)(ws + pl.b_lo);
  float* ps = (float*)(ws + pl.ps);
  if ((e = gather_rows2(q_hi, q_lo, pair_a, P, d, a_hi, a_lo, st))) return e;
  if ((e = gather_rows2(e_hi, e_lo, pair_e, P, d, b_hi, b_lo, st))) return e;
  if ((e = pair_scores_tf32(a_hi, a_lo, b_hi, b_lo, P, d, 0, nullptr, nullptr, nullptr, 1.0, nullptr, nullptr, ps, 3, st))) return e;
  cudaMemsetAsync(raw, 0, (size_t)B * sizeof(int), st);
  if ((e = score_count_tf32(q_hi, q_lo, e_hi, e_lo, B, N, d, ps, target, raw, 0, 0, nullptr, nullptr, nullptr, 1.0,
                            nullptr, nullptr, 3, st))) return e;
  if ((e = filter_correct(B, fe_ptr, fe_idx, target, ps, raw, 0, N, filt, fe_end, st))) return e;
  // ---- relation head: tower on [tanh E[s]; tanh E[o]], dense (B,2R) scores, count-based ranks -------------------------
  if ((e = run_tower(e_all, e_all, triples, 0, 2, B, d, C, ksz, tower_rel, pl, ws, 1, st))) return e;
  float* r_hi = (float*)(ws + pl.r_hi);
  float* r_lo = (float*)(ws + pl.r_lo);
  if ((e = split_tf32(r_emb, r_hi, r_lo, (size_t)R2 * d, st))) return e;
  float* score_rel = (float*)(ws + pl.score_rel);
  if ((e = gemm_tf32(q_hi, q_lo, d, r_hi, r_lo, d, score_rel, R2, B, R2, d, nullptr, 0, 3, 1, nullptr, 0, nullptr, 0, st))) return e;
  float* tscore_r = (float*)(ws + pl.tscore_r);
  int* raw_r = (int*)(ws + pl.raw_r);
  int* filt_r = (int*)(ws + pl.filt_r);
  cudaMemsetAsync(tscore_r, 0, (size_t)B * sizeof(float), st);
  if ((e = gather_target_score(score_rel, R2, B, R2, triples, 1, 0, tscore_r, st))) return e;
  if ((e = rank_count(score_rel, R2, B, R2, triples, 1, fr_ptr, fr_idx, 0, tscore_r, raw_r, filt_r, fr_end, st))) return e;
  launch_k(pack_ranks_kernel, (unsigned)((B + 255) / 256), 256, 0, st, (const int*)raw, (const int*)filt, (const int*)raw_r,
           (const int*)filt_r, B, packed);
  return check_launch("convtrans_decode_rank");
}

}  // namespace regcn

using namespace regcn;
extern "C" {
size_t regcn_convtrans_decode_rank_workspace_bytes(int N, int R2, int d, int B, int C, int P) {
  return convtrans_decode_rank_workspace_bytes(N, R2, d, B, C, P);
}
int regcn_convtrans_decode_rank(const float* emb, const float* r_emb, const int64_t* triples,
                                const void* const* tower_ent, const void* const* tower_rel, const int32_t* fe_ptr,
                                const int32_t* fe_idx, const int32_t* fe_end, const int32_t* pair_a,
                                const int32_t* pair_e, int P, const int32_t* fr_ptr, const int32_t* fr_idx,
                                const int32_t* fr_end, int N, int R2, int d, int B, int C, int ksz, int layer_norm,
                                int32_t* packed, void* workspace, size_t workspace_bytes, void* stream) {
  return convtrans_decode_rank(emb, r_emb, triples, tower_ent, tower_rel, fe_ptr, fe_idx, fe_end, pair_a, pair_e, P,
                               fr_ptr, fr_idx, fr_end, N, R2, d, B, C, ksz, layer_norm, packed, workspace,
                               workspace_bytes, (cudaStream_t)stream);
}
}
