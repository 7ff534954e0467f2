// extern "C" surface of libregcn_b200.so (declared in include/regcn_b200.h).
#include "../../include/regcn_b200.h"
#include "common.cuh"
#include "internal.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace regcn {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return (int)e;
  }
  return REGCN_OK;
}
}  // namespace regcn

namespace regcn {
static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("REGCN_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
void pdl_set(int on) { g_pdl = on ? 1 : 0; }
static long long g_kernel_launches = 0;
void count_kernel_launch() { __atomic_fetch_add(&g_kernel_launches, 1, __ATOMIC_RELAXED); }
static thread_local bool g_pdl_suppress = false;
void pdl_suppress(bool on) { g_pdl_suppress = on; }
bool pdl_suppressed() { return g_pdl_suppress; }
}  // namespace regcn

namespace regcn {
struct ProfRec { cudaEvent_t a, b; double work; };
static bool g_prof = false;
static std::vector<ProfRec> g_recs[PROF_NUM_SLOTS];
bool prof_on() { return g_prof; }
void prof_begin(int slot, cudaStream_t st) {
  if (!g_prof) return;
  ProfRec r;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b); r.work = 0;
  cudaEventRecord(r.a, st);
  g_recs[slot].push_back(r);
}
void prof_end(int slot, double work, cudaStream_t st) {
  if (!g_prof || g_recs[slot].empty()) return;
  ProfRec& r = g_recs[slot].back();
  r.work = work;
  cudaEventRecord(r.b, st);
}
}  // namespace regcn

using namespace regcn;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int regcn_version(void) { return 100; }
const char* regcn_last_error_string(void) { return g_err; }
int regcn_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? 1 : 0;
}

void regcn_prof_enable(int on) {
  g_prof = on != 0;
  if (on) for (int s = 0; s < PROF_NUM_SLOTS; ++s) {
    for (auto& r : g_recs[s]) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_recs[s].clear();
  }
}
int regcn_prof_read(int slot, double* total_ms, long long* launches, double* total_work) {
  if (slot < 0 || slot >= PROF_NUM_SLOTS || !total_ms || !launches || !total_work) return REGCN_ERR_DIM;
  cudaDeviceSynchronize();
  double ms = 0, work = 0;
  for (auto& r : g_recs[slot]) { float t = 0; cudaEventElapsedTime(&t, r.a, r.b); ms += t; work += r.work; }
  *total_ms = ms; *launches = (long long)g_recs[slot].size(); *total_work = work;
  return REGCN_OK;
}

size_t regcn_csr_build_workspace_bytes(int T, int N, int R) { return csr_build_workspace_bytes(T, N, R); }
int regcn_csr_build(const int64_t* triples, int T, int N, int R, int32_t* src, int32_t* dst, int32_t* etype,
                    int32_t* indeg, float* norm, int32_t* rowptr, int32_t* src_sorted, int32_t* etype_sorted,
                    int32_t* eperm, int32_t* vptr, int32_t* sptr, int32_t* vrow_row, int32_t* active_pos,
                    int32_t* active_rows, int32_t* rel_rowptr, int32_t* rel_ents, int32_t* counts, void* workspace,
                    size_t workspace_bytes, void* stream) {
  return csr_build(triples, T, N, R, src, dst, etype, indeg, norm, rowptr, src_sorted, etype_sorted, eperm, vptr, sptr,
                   vrow_row, active_pos, active_rows, rel_rowptr, rel_ents, counts, workspace, workspace_bytes, ST(stream));
}
size_t regcn_csr_build_batch_workspace_bytes(const int32_t* T, int L, int N, int R) {
  return T ? csr_build_batch_workspace_bytes(T, L, N, R) : 0;
}
int regcn_csr_build_batch(const regcn_csr_arrays* snaps, int L, int N, int R, void* workspace, size_t workspace_bytes,
                          void* stream) {
  return csr_build_batch(snaps, L, N, R, workspace, workspace_bytes, ST(stream));
}
int regcn_csr_concat(const regcn_csr_arrays* members, const int32_t* member_sizes, int G, int N, int R,
                     const regcn_csr_arrays* out, void* stream) {
  return csr_concat(members, member_sizes, G, N, R, out, ST(stream));
}
int regcn_rel_mean_pool(const float* h, const int32_t* rel_rowptr, const int32_t* rel_ents, int R, int d, int nsplit,
                        float* out, float* partial, void* stream) {
  return rel_mean_pool(h, rel_rowptr, rel_ents, R, d, nsplit, out, partial, nullptr, nullptr, ST(stream));
}
int regcn_union_aggregate(const float* h, const float* rel, const int32_t* rowptr, const int32_t* src_sorted,
                          const int32_t* etype_sorted, const float* norm, const int32_t* vptr, const int32_t* sptr,
                          const int32_t* vrow_row, int n_vrows, int n_split_chunks, const float* radius, float gamma,
                          int N, int d, float* out, float* partial, void* stream) {
  return union_aggregate(h, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, n_vrows, n_split_chunks,
                         radius, gamma, N, d, out, partial, nullptr, nullptr, nullptr, 0, 0, ST(stream));
}
int regcn_block_aggregate(const float* h, const float* W, const int32_t* rowptr, const int32_t* src_sorted,
                          const int32_t* etype_sorted, const float* norm, int N, int d_in, int d_out, int nb, float* out,
                          void* stream) {
  return block_aggregate(h, W, rowptr, src_sorted, etype_sorted, norm, N, d_in, d_out, nb, out, ST(stream));
}
int regcn_block_aggregate_radius(const float* h, const float* W, const float* radius, float gamma, const int32_t* rowptr,
                                 const int32_t* src_sorted, const int32_t* etype_sorted, const float* norm, int N, int d_in,
                                 int d_out, int nb, float* out, void* stream) {
  if (!radius) { set_last_error("block_aggregate_radius: null radius"); return REGCN_ERR_NULL; }
  return block_aggregate(h, W, rowptr, src_sorted, etype_sorted, norm, N, d_in, d_out, nb, out, ST(stream), radius, gamma);
}
int regcn_lorentz_aggregate(const float* ht, const float* W, const float* rel, const int32_t* rowptr,
                            const int32_t* src_sorted, const int32_t* etype_sorted, const float* norm,
                            const int32_t* vptr, const int32_t* sptr, const int32_t* vrow_row, int n_vrows,
                            int n_split_chunks, int N, int d, int nb, double c, float* out, float* partial,
                            void* stream) {
  return lorentz_aggregate(ht, W, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, n_vrows,
                           n_split_chunks, N, d, nb, c, out, partial, ST(stream));
}
size_t regcn_gemm_f32_workspace_bytes(int M, int N, int split_k) { return gemm_f32_workspace_bytes(M, N, split_k); }
int regcn_gemm_f32(const float* A, int lda, const float* B, int ldb, int transB, float* C, int ldc, int M, int N, int K,
                   const float* bias, int accumulate, int split_k, float* workspace, size_t workspace_bytes,
                   void* stream) {
  return gemm_f32(A, lda, B, ldb, transB, C, ldc, M, N, K, bias, accumulate, split_k, workspace, workspace_bytes,
                  ST(stream));
}
int regcn_split_tf32(const float* x, float* hi, float* lo, size_t n, void* stream) {
  return split_tf32(x, hi, lo, n, ST(stream));
}
int regcn_to_bf16(const float* x, void* out, size_t n, void* stream) { return to_bf16(x, out, n, ST(stream)); }
size_t regcn_gemm_tf32_workspace_bytes(int M, int N, int split_k) { return gemm_tf32_workspace_bytes(M, N, split_k); }
int regcn_gemm_tf32(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb,
                    float* C, int ldc, int M, int N, int K, const float* bias, int accumulate, int passes, int split_k,
                    float* workspace, size_t workspace_bytes, void* stream) {
  return gemm_tf32(a_hi, a_lo, lda, b_hi, b_lo, ldb, C, ldc, M, N, K, bias, accumulate, passes, split_k, workspace,
                   workspace_bytes, nullptr, 0, ST(stream));
}
void regcn_pdl_enable(int on) { regcn::pdl_set(on); }
void regcn_two_stream_enable(int on) { regcn::two_stream_set(on); }
void regcn_evolve_a32_mode(int mode) { regcn::evolve_a32_set(mode); }
long long regcn_kernel_launches(void) { return __atomic_load_n(&regcn::g_kernel_launches, __ATOMIC_RELAXED); }
void regcn_gemm_tf32_tune(int block_n, int stages) { gemm_tf32_tune(block_n, stages); }
void regcn_gemm_tf32_grid_cap(int ctas) { regcn::gemm_tf32_grid_cap(ctas); }
void regcn_aggregate_tune(int impl) { aggregate_tune(impl); }
int regcn_gemm_tf32_a32(const float* a0, int lda0, int k0, const int32_t* rows0, const float* a1, int lda1, int k1,
                        const int32_t* rows1, const float* b_hi, const float* b_lo, int ldb, float* C, int ldc, int M, int N,
                        const float* bias, int accumulate, int passes, int split_k, float* workspace, size_t workspace_bytes,
                        const float* addend, int ld_add, void* stream) {
  return gemm_tf32_a32(a0, lda0, k0, rows0, a1, lda1, k1, rows1, b_hi, b_lo, ldb, C, ldc, M, N, bias, accumulate, passes,
                       split_k, workspace, workspace_bytes, addend, ld_add, ST(stream));
}
int regcn_gemm_tf32_layer_a32(const float* a0, int lda0, int k0, const int32_t* rows0, const float* a1, int lda1, int k1,
                              const int32_t* rows1, const float* b_hi, const float* b_lo, int ldb, int M, int N, int d,
                              float* out_raw, float* out_hi, float* out_lo, float* gate_out, int ld_gate_out,
                              const int32_t* row_idx, const int32_t* skip_rows, const float* gate_G, int gate_ld,
                              const float* gate_bias, const float* gate_h, int gate_norm, void* stream) {
  return gemm_tf32_layer_a32(a0, lda0, k0, rows0, a1, lda1, k1, rows1, b_hi, b_lo, ldb, M, N, d, out_raw, out_hi, out_lo,
                             gate_out, ld_gate_out, row_idx, skip_rows, gate_G, gate_ld, gate_bias, gate_h, gate_norm,
                             ST(stream));
}
void regcn_gemm_tf32_trace(void* dev_buf) { gemm_tf32_trace(dev_buf); }
void regcn_gemm_tf32_trace_begin(void* dev_buf, size_t bytes) { gemm_tf32_trace_begin(dev_buf, bytes); }
int regcn_gemm_tf32_trace_count(void) { return gemm_tf32_trace_count(); }
int regcn_gemm_tf32_trace_read(int i, int* epi, int* M, int* N, int* K, int* grid, int* passes, double* flops) {
  return gemm_tf32_trace_read(i, epi, M, N, K, grid, passes, flops);
}
void regcn_score_count_poly(int on) { score_count_poly(on); }
int regcn_gemm_tf32_trace_slots(void) { return gemm_tf32_trace_slots(); }
int regcn_gemm_tf32_layer(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb,
                          int M, int N, int K, int d, float* out_raw, float* out_hi, float* out_lo, float* gate_out,
                          int ld_gate_out, const int32_t* row_idx, const int32_t* skip_rows, const float* gate_G,
                          int gate_ld, const float* gate_bias, const float* gate_h, int gate_norm, void* stream) {
  return gemm_tf32_layer(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, d, out_raw, out_hi, out_lo, gate_out, ld_gate_out,
                         row_idx, skip_rows, gate_G, gate_ld, gate_bias, gate_h, gate_norm, ST(stream));
}
int regcn_score_count_tf32(const float* q_hi, const float* q_lo, const float* e_hi, const float* e_lo, int B, int N, int K,
                           const float* tscore, const int32_t* target, int32_t* raw_count, int col_offset, int hyp,
                           const float* x2, const float* y2, const float* col_bias, double c,
                           const float* scale_margin, const float* row_c, int passes, void* stream) {
  return score_count_tf32(q_hi, q_lo, e_hi, e_lo, B, N, K, tscore, target, raw_count, col_offset, hyp, x2, y2, col_bias, c,
                          scale_margin, row_c, passes, ST(stream));
}
int regcn_pair_scores_tf32(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int P, int K,
                           int hyp, const float* x2, const float* y2, const float* col_bias, double c,
                           const float* scale_margin, const float* row_c, float* out, int passes, void* stream) {
  return pair_scores_tf32(a_hi, a_lo, b_hi, b_lo, P, K, hyp, x2, y2, col_bias, c, scale_margin, row_c, out, passes,
                          ST(stream));
}
int regcn_score_lse_num_parts(int N) { return score_lse_num_parts(N); }
int regcn_score_lse_tf32(const float* q_hi, const float* q_lo, const float* e_hi, const float* e_lo, int B, int N, int K,
                         int hyp, const float* x2, const float* y2, const float* col_bias, double c,
                         const float* scale_margin, const float* row_c, int passes, float* part_max, float* part_sum,
                         void* stream) {
  return score_lse_tf32(q_hi, q_lo, e_hi, e_lo, B, N, K, hyp, x2, y2, col_bias, c, scale_margin, row_c, passes, part_max,
                        part_sum, ST(stream));
}
int regcn_ce_from_lse(const float* part_max, const float* part_sum, int nparts, int B, const float* tscore, float* ce,
                      float* loss, void* stream) {
  return ce_from_lse(part_max, part_sum, nparts, B, tscore, ce, loss, ST(stream));
}
int regcn_gather_rows2(const float* src_hi, const float* src_lo, const int32_t* idx, int P, int d, float* out_hi,
                       float* out_lo, void* stream) {
  return gather_rows2(src_hi, src_lo, idx, P, d, out_hi, out_lo, ST(stream));
}
int regcn_gather_scalars(const float* a, const float* b, const float* c, const int32_t* ia, const int32_t* ib, int P,
                         float* oa, float* ob, float* oc, void* stream) {
  return gather_scalars(a, b, c, ia, ib, P, oa, ob, oc, ST(stream));
}
int regcn_filter_correct(int B, const int32_t* filt_ptr, const int32_t* filt_idx, const int32_t* target,
                         const float* pair_score, const int32_t* raw_count, int col_lo, int col_hi,
                         int32_t* filt_count, const int32_t* filt_end, void* stream) {
  return filter_correct(B, filt_ptr, filt_idx, target, pair_score, raw_count, col_lo, col_hi, filt_count, filt_end,
                        ST(stream));
}
int regcn_queries_prepare(const int64_t* triples, int T, int R, int64_t* all_t, int32_t* counts, int32_t* beg, int32_t* totals,
                          void* stream) {
  return queries_prepare(triples, T, R, all_t, counts, beg, totals, ST(stream));
}
int regcn_queries_prepare_batch(const int64_t* triples_cat, const int32_t* toff, int n, int R, int64_t* all_t_cat,
                                int32_t* counts_cat, int32_t* beg_cat, int32_t* totals, void* stream) {
  return queries_prepare_batch(triples_cat, toff, n, R, all_t_cat, counts_cat, beg_cat, totals, ST(stream));
}
int regcn_filter_count(const int64_t* triples, int B, int key_col, int32_t* counts, void* stream) {
  return filter_count(triples, B, key_col, counts, ST(stream));
}
int regcn_filter_fill(const int64_t* triples, int B, int key_col, int ans_col, const int32_t* beg, int32_t* idx,
                      int32_t* end, int32_t* pair_a, int32_t* pair_e, void* stream) {
  return filter_fill(triples, B, key_col, ans_col, beg, idx, end, pair_a, pair_e, ST(stream));
}
int regcn_filter_fill2(const int64_t* triples, int B, const int32_t* beg_e, int32_t* idx_e, int32_t* end_e, int32_t* pair_a_e,
                       int32_t* pair_e_e, const int32_t* beg_r, int32_t* idx_r, int32_t* end_r, int32_t* pair_a_r,
                       int32_t* pair_e_r, void* stream) {
  return filter_fill2(triples, B, beg_e, idx_e, end_e, pair_a_e, pair_e_e, beg_r, idx_r, end_r, pair_a_r, pair_e_r, ST(stream));
}
int regcn_row_map_split(const float* x, float* out, float* out_hi, float* out_lo, int M, int d, int mode, double c,
                        void* stream) {
  return row_map(x, out, M, d, mode, c, nullptr, out_hi, out_lo, ST(stream));
}
int regcn_row_map(const float* x, float* out, int M, int d, int mode, double c, float* sumsq, void* stream) {
  return row_map(x, out, M, d, mode, c, sumsq, nullptr, nullptr, ST(stream));
}
int regcn_gru_gate(const float* gi, const float* gh, const float* hprev, float* out, int M, int d, int normalize,
                   void* stream) {
  return gru_gate(gi, gh, hprev, out, M, d, normalize, nullptr, nullptr, ST(stream));
}
int regcn_union_combine(const float* P, const float* L, const int32_t* indeg, const float* S, const float* skip_bias,
                        const float* prev, int N, int d, int act, int hyper, double c, float* out, float* ht_next,
                        float* radius_next, void* stream) {
  return union_combine(P, L, indeg, S, skip_bias, prev, N, d, act, hyper, c, out, ht_next, radius_next, 0, nullptr,
                       nullptr, nullptr, nullptr, nullptr, ST(stream));
}
int regcn_time_gate(const float* G, const float* bias, const float* cur, const float* h, float* out, int N, int d,
                    int normalize_cur, void* stream) {
  return time_gate(G, bias, cur, h, out, N, d, normalize_cur, 0, nullptr, nullptr, ST(stream));
}
int regcn_hyp_init(const float* emb, const float* radius_static, int N, int d, int normalize, int on_manifold, double c,
                   float radius_min, float radius_max, float* out, void* stream) {
  return hyp_init(emb, radius_static, N, d, normalize, on_manifold, c, radius_min, radius_max, out, ST(stream));
}
int regcn_hyp_tangent(const float* h, int N, int d, double c, float* ht, float* pt, float* radius, void* stream) {
  return hyp_tangent(h, N, d, c, ht, pt, radius, nullptr, nullptr, nullptr, nullptr, ST(stream));
}
int regcn_hyp_time_gate(const float* h2, const float* pt, const float* G, const float* bias, const float* radius_static,
                        const float* radius_w, float radius_b, int N, int d, int layer_norm, int residual, double c,
                        float radius_min, float radius_max, float beta, float eps_r, float* out, void* stream) {
  return hyp_time_gate(h2, pt, G, bias, radius_static, radius_w, radius_b, N, d, layer_norm, residual, c, radius_min,
                       radius_max, beta, eps_r, out, ST(stream));
}
int regcn_convtranse_features(const float* ent, const float* second, const int64_t* triples, int col0, int col1, int B,
                              int d, int C, int ksz, const float* bn0_scale, const float* bn0_shift,
                              const float* conv_w, const float* conv_b, const float* bn1_scale, const float* bn1_shift,
                              float* F, float* F_hi, float* F_lo, void* stream) {
  return convtranse_features(ent, second, triples, col0, col1, B, d, C, ksz, bn0_scale, bn0_shift, conv_w, conv_b,
                             bn1_scale, bn1_shift, F, F_hi, F_lo, ST(stream));
}
int regcn_convtrans_fc_pack_weight(const float* fc_weight, int N, int C, int d, float* w_hi, float* w_lo, void* stream) {
  return convfc_pack_weight(fc_weight, N, C, d, w_hi, w_lo, ST(stream));
}
size_t regcn_convtrans_fc_workspace_bytes(int batch_total, int N) { return convtrans_fc_workspace_bytes(batch_total, N); }
int regcn_convtrans_fc(const float* x0, const float* x1, const int64_t* triples, int col0, int col1, int B, int batch_total,
                       int d, int C, int ksz, const float* bn0_scale, const float* bn0_shift, const float* conv_w,
                       const float* conv_b, const float* bn1_scale, const float* bn1_shift, const float* w_hi,
                       const float* w_lo, int ldw, int N, const float* bias, const float* bn2_scale, const float* bn2_shift,
                       int relu, float* out, int ldc, float* out_hi, float* out_lo, float* ws, size_t ws_bytes,
                       void* stream) {
  return convtrans_fc(x0, x1, triples, col0, col1, B, d, C, ksz, bn0_scale, bn0_shift, conv_w, conv_b, bn1_scale, bn1_shift,
                      w_hi, w_lo, ldw, N, bias, out, ldc, ws, ws_bytes, ST(stream), batch_total, bn2_scale, bn2_shift, relu,
                      out_hi, out_lo);
}
int regcn_affine_relu(float* x, const float* scale, const float* shift, int M, int d, int relu, void* stream) {
  return affine_relu(x, scale, shift, M, d, relu, ST(stream));
}
int regcn_gather_log0(const float* E, const int64_t* triples, int col, int B, int d, int project, double c, float* out,
                      void* stream) {
  return gather_log0(E, triples, col, B, d, project, c, out, ST(stream));
}
int regcn_hyp_query(const float* s_tan, const float* ang, const float* trans, const float* E, const int64_t* triples,
                    int B, int d, int kind, double c, float* Q, float* q_sumsq, void* stream) {
  return hyp_query(s_tan, ang, trans, E, triples, B, d, kind, c, Q, q_sumsq, ST(stream));
}
int regcn_hyp_score_epilogue(float* S, int ld, int B, int N, const float* q_sumsq, const float* e_sumsq,
                             const float* bias, const float* qbias, double c, const float* scale_margin,
                             const float* row_c, void* stream) {
  return hyp_score_epilogue(S, ld, B, N, q_sumsq, e_sumsq, bias, qbias, c, scale_margin, row_c, ST(stream));
}
int regcn_rel_curvature(const float* raw, const int64_t* triples, int B, int R, double c, double cmax, float* out,
                        void* stream) {
  return rel_curvature(raw, triples, B, R, c, cmax, out, ST(stream));
}
int regcn_gather_target_score(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                              int col_offset, float* target_score, void* stream) {
  return gather_target_score(S, ld, B, N, triples, target_col, col_offset, target_score, ST(stream));
}
int regcn_rank_count(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                     const int32_t* filt_ptr, const int32_t* filt_idx, int col_offset, const float* target_score,
                     int32_t* raw_count, int32_t* filt_count, const int32_t* filt_end, void* stream) {
  return rank_count(S, ld, B, N, triples, target_col, filt_ptr, filt_idx, col_offset, target_score, raw_count,
                    filt_count, filt_end, ST(stream));
}
int regcn_ce_rows(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, float* ce,
                  float* loss, void* stream) {
  int e = ce_rows(S, ld, B, N, triples, target_col, ce, ST(stream));
  if (e || !loss || B <= 0) return e;
  return mean_f32(ce, B, loss, ST(stream));
}
int regcn_counts_to_ranks(const int32_t* raw_count, const int32_t* filt_count, int B, int64_t* rank, int64_t* filt_rank,
                          void* stream) {
  return counts_to_ranks(raw_count, filt_count, B, rank, filt_rank, ST(stream));
}
int regcn_apply_filter(float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                       const int32_t* filt_ptr, const int32_t* filt_idx, int col_offset, const int32_t* filt_end,
                       void* stream) {
  return apply_filter(S, ld, B, N, triples, target_col, filt_ptr, filt_idx, col_offset, filt_end, ST(stream));
}

int regcn_topk_construct_snap(const float* S, int64_t ld, int B, int N, int K, const int64_t* triples, int R,
                              int rel_mode, int32_t* top_idx, int64_t* out, void* stream) {
  return topk_construct_snap(S, ld, B, N, K, triples, R, rel_mode, top_idx, out, ST(stream));
}
int regcn_atth_query(const float* s_tan, const float* rot, const float* ref, const float* attn, const float* rel,
                     const float* trans, const float* E, const int64_t* triples, int B, int d, int mode, double c,
                     float* Q, float* q_sumsq, void* stream) {
  return atth_query(s_tan, rot, ref, attn, rel, trans, E, triples, B, d, mode, c, Q, q_sumsq, ST(stream));
}
int regcn_gemm_tf32_mn(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb,
                       float* C, int ldc, int M, int N, int K, int a_mn, int b_mn, const float* bias, int accumulate,
                       int passes, int split_k, float* workspace, size_t workspace_bytes, void* stream) {
  return gemm_tf32_mn(a_hi, a_lo, lda, b_hi, b_lo, ldb, C, ldc, M, N, K, a_mn, b_mn, bias, accumulate, passes, split_k,
                      workspace, workspace_bytes, ST(stream));
}
size_t regcn_lorentz_aggregate_bwd_workspace_bytes(int N, int R2, int d) { return lorentz_aggregate_bwd_workspace_bytes(N, R2, d); }
int regcn_lorentz_bwd_splits(void) { return lorentz_bwd_splits(); }
int regcn_lorentz_aggregate_bwd(const float* ht, const float* W, const float* rel, const float* gout, const int32_t* rowptr,
                                const int32_t* src_sorted, const int32_t* etype_sorted, const float* norm,
                                const int32_t* type_rowptr, const int32_t* type_src, const int32_t* type_dst, int N, int R2,
                                int d, int num_bases, double c, float* dht, float* part_rel, float* part_w, float* workspace,
                                size_t workspace_bytes, void* stream) {
  return lorentz_aggregate_bwd(ht, W, rel, gout, rowptr, src_sorted, etype_sorted, norm, type_rowptr, type_src, type_dst, N,
                               R2, d, num_bases, c, dht, part_rel, part_w, workspace, workspace_bytes, ST(stream));
}
}  // extern "C"
