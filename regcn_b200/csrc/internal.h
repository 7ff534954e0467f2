// Internal C++ declarations of the kernel launchers (one per extern "C" entry in include/regcn_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/regcn_b200.h"

namespace regcn {
size_t csr_build_workspace_bytes(int T, int N, int R);
int csr_build(const int64_t* triples, int T, int N, int R, int* src, int* dst, int* etype, int* indeg, float* norm,
              int* rowptr, int* src_sorted, int* etype_sorted, int* eperm, int* vptr, int* sptr, int* vrow_row,
              int* active_pos, int* active_rows, int* rel_rowptr, int* rel_ents, int* counts, void* ws, size_t ws_bytes, cudaStream_t st);
size_t csr_build_batch_workspace_bytes(const int* T, int L, int N, int R);
int csr_build_batch(const regcn_csr_arrays* snaps, int L, int N, int R, void* ws, size_t ws_bytes, cudaStream_t st);
int csr_concat(const regcn_csr_arrays* members, const int* sizes, int G, int N, int R, const regcn_csr_arrays* out, cudaStream_t st);
int rel_mean_pool(const float* h, const int* rel_rowptr, const int* rel_ents, int R, int d, int nsplit, float* out,
                  float* partial, float* out_hi, float* out_lo, cudaStream_t st);
int union_aggregate(const float* h, const float* rel, const int* rowptr, const int* src_sorted, const int* etype_sorted,
                    const float* norm, const int* vptr, const int* sptr, const int* vrow_row, int nv, int nsplit,
                    const float* radius, float gamma, int N, int d, float* out, float* partial, float* out_hi,
                    float* out_lo, const int* active_pos, int ldo, int max_chunks, cudaStream_t st,
                    int* fold_count = nullptr);
int block_aggregate(const float* h, const float* W, const int* rowptr, const int* src_sorted, const int* etype_sorted,
                    const float* norm, int N, int d_in, int d_out, int nb, float* out, cudaStream_t st,
                    const float* radius = nullptr, float gamma = 0.f);
int lorentz_aggregate(const float* ht, const float* W, const float* rel, const int* rowptr, const int* src_sorted,
                      const int* etype_sorted, const float* norm, const int* vptr, const int* sptr, const int* vrow_row,
                      int nv_rows, int nsplit, int N, int d, int nb, double c, float* out, float* partial,
                      cudaStream_t st);
size_t gemm_f32_workspace_bytes(int M, int N, int split_k);
int gemm_f32(const float* A, int lda, const float* B, int ldb, int transB, float* C, int ldc, int M, int N, int K,
             const float* bias, int accumulate, int split_k, float* ws, size_t ws_bytes, cudaStream_t st);
int split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t st);
int to_bf16(const float* x, void* out, size_t n, cudaStream_t st);
size_t gemm_tf32_workspace_bytes(int M, int N, int split_k);
int gemm_tf32(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb, float* C,
              int ldc, int M, int N, int K, const float* bias, int accumulate, int passes, int split_k, float* ws,
              size_t ws_bytes, const float* addend, int ld_add, cudaStream_t st);
int gemm_tf32_layer(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb, int M,
                    int N, int K, int d, float* out_raw, float* out_hi, float* out_lo, float* gate_out, int ld_gate_out,
                    const int* row_idx, const int* skip_rows, const float* gate_G, int gate_ld, const float* gate_bias,
                    const float* gate_h, int gate_norm, cudaStream_t st);
int gemm_tf32_a32(const float* a0, int lda0, int k0, const int* rows0, const float* a1, int lda1, int k1, const int* rows1,
                  const float* b_hi, const float* b_lo, int ldb, float* C, int ldc, int M, int N, const float* bias,
                  int accumulate, int passes, int split_k, float* ws, size_t ws_bytes, const float* addend, int ld_add,
                  cudaStream_t st);
int gemm_tf32_layer_a32(const float* a0, int lda0, int k0, const int* rows0, const float* a1, int lda1, int k1,
                        const int* rows1, const float* b_hi, const float* b_lo, int ldb, int M, int N, int d,
                        float* out_raw, float* out_hi, float* out_lo, float* gate_out, int ld_gate_out, const int* row_idx,
                        const int* skip_rows, const float* gate_G, int gate_ld, const float* gate_bias, const float* gate_h,
                        int gate_norm, cudaStream_t st);
void pdl_set(int on);
void two_stream_set(int on);
void evolve_a32_set(int mode);
void gemm_tf32_tune(int block_n, int stages);
void gemm_tf32_trace(void* dev_buf);
void gemm_tf32_trace_begin(void* dev_buf, size_t bytes);
int gemm_tf32_trace_count();
int gemm_tf32_trace_read(int i, int* epi, int* M, int* N, int* K, int* grid, int* passes, double* flops);
void score_count_poly(int on);
int gemm_tf32_trace_slots();
void gemm_tf32_sm_hint(int sms);
void gemm_tf32_grid_cap(int ctas);
void aggregate_tune(int impl);
int score_count_tf32(const float* q_hi, const float* q_lo, const float* e_hi, const float* e_lo, int B, int N, int K,
                     const float* tscore, const int* target, int* raw_count, int col_offset, int hyp, const float* x2,
                     const float* y2, const float* col_bias, double c, const float* scale_margin, const float* row_c,
                     int passes, cudaStream_t st);
int pair_scores_tf32(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int P, int K, int hyp,
                     const float* x2, const float* y2, const float* col_bias, double c, const float* scale_margin,
                     const float* row_c, float* out, int passes, cudaStream_t st);
int score_lse_num_parts(int N);
int score_lse_tf32(const float* q_hi, const float* q_lo, const float* e_hi, const float* e_lo, int B, int N, int K, int hyp,
                   const float* x2, const float* y2, const float* col_bias, double c, const float* scale_margin,
                   const float* row_c, int passes, float* part_max, float* part_sum, cudaStream_t st);
int ce_from_lse(const float* part_max, const float* part_sum, int nparts, int B, const float* tscore, float* ce,
                float* loss, cudaStream_t st);
int gather_rows2(const float* src_hi, const float* src_lo, const int* idx, int P, int d, float* out_hi, float* out_lo,
                 cudaStream_t st);
int gather_scalars(const float* a, const float* b, const float* c, const int* ia, const int* ib, int P, float* oa,
                   float* ob, float* oc, cudaStream_t st);
int filter_correct(int B, const int* filt_ptr, const int* filt_idx, const int* target, const float* pair_score,
                   const int* raw_count, int col_lo, int col_hi, int* filt_count, const int* filt_end, cudaStream_t st);
int row_map(const float* x, float* out, int M, int d, int mode, double c, float* sumsq, float* out_hi, float* out_lo,
            cudaStream_t st);
int gru_gate(const float* gi, const float* gh, const float* hprev, float* out, int M, int d, int normalize,
             float* out_hi, float* out_lo, cudaStream_t st);
int union_combine(const float* P, const float* L, const int* indeg, const float* S, const float* skip_bias,
                  const float* prev, int N, int d, int act, int hyper, double c, float* out, float* ht_next,
                  float* radius_next, int ldL, float* out_hi, float* out_lo, float* ht_hi, float* ht_lo,
                  const int* active_pos, cudaStream_t st);
int time_gate(const float* G, const float* bias, const float* cur, const float* h, float* out, int N, int d,
              int normalize_cur, int ldg, float* out_hi, float* out_lo, cudaStream_t st, const int* row_idx = nullptr,
              int act = 0);
int shared_rows_update(const int* active_rows, int n_active, int N0, int d, int* cpos, int* count, float* h_c, float* x_full,
                       int* act_c, int* apos_c, cudaStream_t st);
int shared_rows_expand(const float* h_c, const int* cpos, int N, int N0, int d, float* out, cudaStream_t st);
int hyp_init(const float* emb, const float* radius_static, int N, int d, int normalize, int on_manifold, double c,
             float rmin, float rmax, float* out, cudaStream_t st);
int hyp_tangent(const float* h, int N, int d, double c, float* ht, float* pt, float* radius, float* ht_hi,
                float* ht_lo, float* pt_hi, float* pt_lo, cudaStream_t st);
int hyp_time_gate(const float* h2, const float* pt, const float* G, const float* bias, const float* radius_static,
                  const float* rw, float rb, int N, int d, int layer_norm, int residual, double c, float rmin,
                  float rmax, float beta, float eps_r, float* out, cudaStream_t st);
int convtranse_features(const float* ent, const float* second, const int64_t* triples, int col0, int col1, int B,
                        int d, int C, int ksz, const float* bn0_scale, const float* bn0_shift, const float* conv_w,
                        const float* conv_b, const float* bn1_scale, const float* bn1_shift, float* F, float* F_hi,
                        float* F_lo, cudaStream_t st);
int convfc_pack_weight(const float* w, int N, int C, int d, float* hi, float* lo, cudaStream_t st);
int convtrans_fc_splits(int B);
size_t convtrans_fc_workspace_bytes(int B, int N);
int convtrans_fc(const float* x0, const float* x1, const int64_t* triples, int col0, int col1, int B, int d, int C, int ksz,
                 const float* bn0_scale, const float* bn0_shift, const float* conv_w, const float* conv_b,
                 const float* bn1_scale, const float* bn1_shift, const float* w_hi, const float* w_lo, int ldw, int N,
                 const float* bias, float* out, int ldc, float* ws, size_t ws_bytes, cudaStream_t st, int batch_total = 0,
                 const float* act_scale = nullptr, const float* act_shift = nullptr, int relu = 0, float* out_hi = nullptr,
                 float* out_lo = nullptr);
int affine_relu(float* x, const float* scale, const float* shift, int M, int d, int relu, cudaStream_t st);
int gather_log0(const float* E, const int64_t* triples, int col, int B, int d, int project, double c, float* out,
                cudaStream_t st);
int hyp_query(const float* s_tan, const float* ang, const float* trans, const float* E, const int64_t* triples, int B,
              int d, int kind, double c, float* Q, float* q_sumsq, cudaStream_t st);
int hyp_score_epilogue(float* S, int ld, int B, int N, const float* q_sumsq, const float* e_sumsq, const float* bias,
                       const float* qbias, double c, const float* scale_margin, const float* row_c, cudaStream_t st);
int rel_curvature(const float* raw, const int64_t* triples, int B, int R, double c, double cmax, float* out,
                  cudaStream_t st);
int gather_target_score(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                        int col_offset, float* target_score, cudaStream_t st);
int rank_count(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, const int* filt_ptr,
               const int* filt_idx, int col_offset, const float* target_score, int* raw_count, int* filt_count,
               const int* filt_end, cudaStream_t st);
int filter_count(const int64_t* triples, int B, int key_col, int* counts, cudaStream_t st);
int queries_prepare(const int64_t* triples, int T, int R, int64_t* all_t, int* counts, int* beg, int* totals, cudaStream_t st);
int queries_prepare_batch(const int64_t* triples_cat, const int* toff, int n, int R, int64_t* all_t_cat, int* counts_cat,
                          int* beg_cat, int* totals, cudaStream_t st);
int filter_fill2(const int64_t* triples, int B, const int* beg_e, int* idx_e, int* end_e, int* pair_a_e, int* pair_e_e,
                 const int* beg_r, int* idx_r, int* end_r, int* pair_a_r, int* pair_e_r, cudaStream_t st);
int filter_fill(const int64_t* triples, int B, int key_col, int ans_col, const int* beg, int* idx, int* end, int* pair_a,
                int* pair_e, cudaStream_t st);
int mean_f32(const float* x, int n, float* out, cudaStream_t st);
int ce_rows(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, float* ce, cudaStream_t st);
int counts_to_ranks(const int* raw_count, const int* filt_count, int B, int64_t* rank, int64_t* filt_rank, cudaStream_t st);
int apply_filter(float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, const int* filt_ptr,
                 const int* filt_idx, int col_offset, const int* filt_end, cudaStream_t st);
int topk_construct_snap(const float* S, int64_t ld, int B, int N, int K, const int64_t* triples, int R, int rel_mode,
                        int* top_idx, int64_t* out, cudaStream_t st);
int atth_query(const float* s_tan, const float* rot, const float* ref, const float* attn, const float* rel,
               const float* trans, const float* E, const int64_t* triples, int B, int d, int mode, double c, float* Q,
               float* q_sumsq, cudaStream_t st);
int gemm_tf32_mn(const float* x_hi, const float* x_lo, int ldx, const float* y_hi, const float* y_lo, int ldy, float* C,
                 int ldc, int M, int N, int K, int a_mn, int b_mn, const float* bias, int accumulate, int passes,
                 int split_k, float* ws, size_t ws_bytes, cudaStream_t st);
size_t lorentz_aggregate_bwd_workspace_bytes(int N, int R2, int d);
int lorentz_aggregate_bwd(const float* ht, const float* W, const float* rel, const float* gout, const int* rowptr,
                          const int* src_sorted, const int* etype_sorted, const float* norm, const int* type_rowptr,
                          const int* type_src, const int* type_dst, int N, int R2, int d, int nb, double c, float* dht,
                          float* part_rel, float* part_w, float* ws, size_t ws_bytes, cudaStream_t st);
int lorentz_bwd_splits();
// training (backward.cu)
int csr_gather_sum(const float* X, int ldx, const float* col_w, const float* row_w, const int* rowptr, const int* col,
                   int nrows, int d, int col2_off, float* out, int ldo, int accumulate, cudaStream_t st,
                   const float* rho = nullptr, float gamma = 0.f, const int* partner = nullptr);
int col_sum(const float* X, int ld, int rows, int cols, int L, float* out, int accumulate, float* ws, size_t ws_bytes,
            cudaStream_t st);
}  // namespace regcn
