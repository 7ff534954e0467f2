/* regcn_b200 -- C ABI of the B200 (sm_100a) kernels behind RE-GCN's per-snapshot evolution and
 * all-entity scoring path.  The reference (sgxxyyds/RE-GCN) has no FFI layer: its boundary is the
 * Python nn.Module surface.  The Python mirror in regcn_b200/*.py keeps those
 * signatures and reaches this library through ctypes; every entry point below names the reference
 * code (file:line under the reference root) whose device work it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (16-byte aligned for float rows);
 *     the library never allocates, frees or retains device memory;
 *   - ids on the device are int32, triples are the reference's int64 (T,3) row-major layout;
 *   - float matrices are dense row-major fp32;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and asynchronous;
 *   - return 0 on success, <0 argument error (-1 null pointer, -2 bad dimension/alignment,
 *     -3 workspace too small, -4 unsupported shape), >0 a cudaError_t from the launch;
 *     regcn_last_error_string() describes the last failure on the calling thread.
 */
#ifndef REGCN_B200_H_
#define REGCN_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define REGCN_API __attribute__((visibility("default")))
#else
#define REGCN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

REGCN_API int regcn_version(void);
REGCN_API const char* regcn_last_error_string(void);
/* 1 if the calling process sees a CUDA device of compute capability 10.x, else 0 */
REGCN_API int regcn_device_ok(void);

/* ---- K1 edge index: rgcn/utils.py:100-134 (build_sub_graph), :78-97 (r2e) ---------------------
 * triples (T,3) int64 -> E = 2T edges [src;dst]->[dst;src], type [rel;rel+R]; in-degree; norm;
 * CSR by destination (stable in edge id): rowptr (N+1), src_sorted/etype_sorted/eperm (E);
 * virtual rows (chunks of 32 in-edges) of the ACTIVE destinations only: vptr/sptr (N+1), vrow_row (min(N,E)+E/32+1);
 * active_pos (N): position of a destination among the active ones (in-degree > 0), -1 otherwise;
 * active_rows (min(N,E)): the inverse map, the sorted ids of the active destinations;
 * relation->entity CSR shared by r and r+R: rel_rowptr (R+1), rel_ents (<= 2T, sorted per relation);
 * counts[8] = {n_virtual_rows, n_split_chunks, n_rel_ents, max_hub_degree, n_active, 0, 0, 0}.   */
REGCN_API size_t regcn_csr_build_workspace_bytes(int T, int N, int R);
REGCN_API int regcn_csr_build(const int64_t* triples, int T, int N, int R,
                    int32_t* src, int32_t* dst, int32_t* etype, int32_t* indeg, float* norm,
                    int32_t* rowptr, int32_t* src_sorted, int32_t* etype_sorted, int32_t* eperm,
                    int32_t* vptr, int32_t* sptr, int32_t* vrow_row, int32_t* active_pos, int32_t* active_rows,
                    int32_t* rel_rowptr, int32_t* rel_ents, int32_t* counts,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Batched form: the index of L history snapshots in one call (the reference rebuilds all L graphs for every
 * evaluated timestamp, src/main.py:68,233; hyperbolic_main.py:100-113).  `snaps` is a HOST array of L descriptors
 * holding the same device pointers regcn_csr_build takes.  Snapshots of at most 16384 edges are built by one CTA
 * each (block radix sorts + fused block scans in shared memory), all L concurrently in one launch; larger ones go
 * through the regcn_csr_build pipeline on the same stream and need the workspace the query below reports.       */
typedef struct regcn_csr_arrays {
  const int64_t* triples;   /* (T,3) int64, device */
  int32_t T;
  int32_t* src; int32_t* dst; int32_t* etype; int32_t* indeg; float* norm;
  int32_t* rowptr; int32_t* src_sorted; int32_t* etype_sorted; int32_t* eperm;
  int32_t* vptr; int32_t* sptr; int32_t* vrow_row; int32_t* active_pos; int32_t* active_rows;
  int32_t* rel_rowptr; int32_t* rel_ents; int32_t* counts;
} regcn_csr_arrays;
REGCN_API size_t regcn_csr_build_batch_workspace_bytes(const int32_t* T, int L, int N, int R);
REGCN_API int regcn_csr_build_batch(const regcn_csr_arrays* snaps, int L, int N, int R, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* Block-diagonal union of G (<= 16) snapshot indices, each over N entities / R relations, as ONE index over G*N entities
 * and G*R relations: the graph one step of a recurrence batched over G independent history windows runs on.  The
 * reference evaluates test timestamps one after the other, each over its own window (src/main.py:60-90,
 * hyperbolic_main.py:100-113); they do not depend on each other, so G windows are evolved by the same kernels at G times
 * the rows per launch.  Member g's entity v -> g*N + v, relation r < R -> g*R + r, inverse relation R + r ->
 * G*R + g*R + r.  member_sizes: (G,4) HOST int32 = counts[0], counts[1], counts[2], counts[4] of every member as
 * regcn_csr_build reported them; `out` must provide arrays for G*N entities, G*R relations and the summed edge counts
 * (out->triples / out->T are not read).  out->counts receives the combined counters.                              */
REGCN_API int regcn_csr_concat(const regcn_csr_arrays* members, const int32_t* member_sizes, int G, int N, int R,
                               const regcn_csr_arrays* out, void* stream);

/* ---- K2 relation mean-pool: src/rrgcn.py:161-166, hyperbolic_model.py:802-812 -----------------
 * out (2R,d): out[r] = out[r+R] = mean of h rows in ents(r); zero rows for absent relations.
 * nsplit > 1 splits every relation over nsplit CTAs (partial: R*nsplit*d floats).               */
REGCN_API int regcn_rel_mean_pool(const float* h, const int32_t* rel_rowptr, const int32_t* rel_ents, int R, int d,
                        int nsplit, float* out, float* partial, void* stream);

/* ---- K4 union aggregate: rgcn/layers.py:257-279; hyperbolic_layers.py:222-240 ------------------
 * out[v] = norm[v] * sum_{(u,r)->v} w_uv (h[u] + rel[r]);  w_uv = exp(-gamma |radius[u]-radius[v]|)
 * when radius != NULL else 1.  partial: n_split_chunks*d floats (NULL when n_split_chunks == 0). */
REGCN_API int regcn_union_aggregate(const float* h, const float* rel, const int32_t* rowptr, const int32_t* src_sorted,
                          const int32_t* etype_sorted, const float* norm, const int32_t* vptr,
                          const int32_t* sptr, const int32_t* vrow_row, int n_vrows, int n_split_chunks,
                          const float* radius, float gamma, int N, int d, float* out, float* partial,
                          void* stream);
/* ---- K6 block-diagonal aggregate: rgcn/layers.py:167-179; hyperbolic_layers.py:87-109 ----------
 * out[v] = norm[v] * sum_in blockdiag(W[type]) . h[src];  W (num_rels, nb*(d_in/nb)*(d_out/nb)). */
REGCN_API int regcn_block_aggregate(const float* h, const float* W, const int32_t* rowptr, const int32_t* src_sorted,
                          const int32_t* etype_sorted, const float* norm, int N, int d_in, int d_out, int nb,
                          float* out, void* stream);
/* HyperbolicRGCNLayer message (hyperbolic_layers.py:87-109): the block-diagonal transform of the source's tangent vector
 * weighted by exp(-gamma |radius[src] - radius[dst]|), summed per destination, times norm. */
REGCN_API int regcn_block_aggregate_radius(const float* h, const float* W, const float* radius, float gamma,
                                           const int32_t* rowptr, const int32_t* src_sorted, const int32_t* etype_sorted,
                                           const float* norm, int N, int d_in, int d_out, int nb, float* out, void* stream);

/* ---- K7 Lorentz centroid aggregate: hyperbolic_layers.py:589-625,665-672; hyperbolic_ops.py:477-518,563-581
 * ht tangent input; out = clamp(log_0(to_poincare(centroid)), +-10), zero rows for in-degree 0.
 * partial: n_split_chunks*(d+1) floats for destinations split into several 32-edge chunks (NULL if none). */
REGCN_API int regcn_lorentz_aggregate(const float* ht, const float* W, const float* rel, const int32_t* rowptr,
                            const int32_t* src_sorted, const int32_t* etype_sorted, const float* norm,
                            const int32_t* vptr, const int32_t* sptr, const int32_t* vrow_row, int n_vrows,
                            int n_split_chunks, int N, int d, int nb, double c, float* out, float* partial,
                            void* stream);

/* ---- dense contraction (fp32 CUDA cores): torch.mm / F.linear call sites on the path -----------
 * C[M,N] (+)= A[M,K] . op(B) (+ bias[N]); op(B) = B[K,N] if !transB else B[N,K]^T.
 * K, lda, ldb multiples of 4; split_k > 1 needs regcn_gemm_f32_workspace_bytes(M,N,split_k).     */
REGCN_API size_t regcn_gemm_f32_workspace_bytes(int M, int N, int split_k);
REGCN_API int regcn_gemm_f32(const float* A, int lda, const float* B, int ldb, int transB, float* C, int ldc, int M,
                   int N, int K, const float* bias, int accumulate, int split_k, float* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- dense contraction on the tensor cores (tcgen05.mma kind::tf32, TMA-fed, TMEM accumulator) ----
 * C[M,N] (+)= A[M,K] . B[N,K]^T (+ bias[N]), A and B both K-major.  passes == 3: error-compensated 3xTF32 on the
 * (hi, lo) splits produced by regcn_split_tf32 (fp32 parity, ~2^-21 operand precision); passes == 1: plain TF32
 * on a_hi / b_hi only (lo may be NULL).  lda, ldb multiples of 4, pointers 16-byte aligned.                    */
REGCN_API int regcn_split_tf32(const float* x, float* hi, float* lo, size_t n, void* stream);
REGCN_API size_t regcn_gemm_tf32_workspace_bytes(int M, int N, int split_k);
REGCN_API int regcn_gemm_tf32(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo,
                    int ldb, float* C, int ldc, int M, int N, int K, const float* bias, int accumulate,
                    int passes, int split_k, float* workspace, size_t workspace_bytes, void* stream);

/* ---- fused all-entity scoring + rank (K11/K13 + K14), no score matrix: src/decoder.py:96-99 and
 * hyperbolic_decoder.py:89-179 fused with rgcn/utils.py:21-25,51-75.
 * regcn_score_count_tf32: the scoring GEMM Q[B,K] . E[N,K]^T with a counting epilogue:
 *     raw_count[b] += #{n in shard, col_offset+n != target[b] : score(b,n) ranks ahead of tscore[b]}
 *   score = <q,e> (+col_bias[n]) or, with hyp != 0, scale*(margin - |(-q)(+)_c e|^2) (+col_bias[n]) from x2=|q|^2, y2=|e|^2;
 *   row_c != NULL (needs hyp != 0) selects the true-distance branch with a per-query curvature (hyperbolic_decoder.py:145-163):
 *   scale*(margin - 2/sqrt(c_q) * artanh(sqrt(c_q) |(-q)(+)_{c_q} e|)), c_q = row_c[b] (regcn_pair_scores_tf32: row_c[p]).
 * regcn_pair_scores_tf32: out[p] = score(A'[p], B'[p]) for gathered operand rows, through the same tensor-core
 *   arithmetic (bit-identical to the corresponding element of the scoring GEMM): target and filter-entry scores.
 * regcn_filter_correct: filt_count[b] = raw_count[b] - (filter entries that beat the target) + (-1e7 entries that do).
 * regcn_gather_rows2 / regcn_gather_scalars: operand gathers for the pair pass.                                */
/* bf16 scoring mode (reported separately from the fp32-parity mode): passes == 0 in regcn_score_count_tf32 /
 * regcn_pair_scores_tf32 means the "hi" operand pointers address bf16 rows of K elements (K % 8 == 0) produced by
 * regcn_to_bf16, the "lo" pointers are ignored, and the contraction runs as tcgen05.mma kind::f16 (fp32 accumulate). */
REGCN_API int regcn_to_bf16(const float* x, void* out_bf16, size_t n, void* stream);
REGCN_API int regcn_score_count_tf32(const float* q_hi, const float* q_lo, const float* e_hi, const float* e_lo, int B,
                           int N, int K, const float* tscore, const int32_t* target, int32_t* raw_count,
                           int col_offset, int hyp, const float* x2, const float* y2, const float* col_bias, double c,
                           const float* scale_margin, const float* row_c, int passes, void* stream);
REGCN_API int regcn_pair_scores_tf32(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int P,
                           int K, int hyp, const float* x2, const float* y2, const float* col_bias, double c,
                           const float* scale_margin, const float* row_c, float* out, int passes, void* stream);
/* ---- loss heads (forward): CrossEntropy over ALL candidates without the (B,N) logits -- src/rrgcn.py:205-223
 * (loss_e / loss_r), hyperbolic_decoder.py:182-307 (_chunked_hyperbolic_ce_loss).
 * regcn_score_lse_tf32: the scoring GEMM with a streaming log-sum-exp epilogue; every (256-candidate tile, warp share)
 *   writes max and sum-of-exp of its slice of each row into part_max / part_sum, regcn_score_lse_num_parts(N) slices of
 *   B floats each.  Score arguments as regcn_score_count_tf32.
 * regcn_ce_from_lse: ce[b] = logsumexp_n score(b,n) - tscore[b] (tscore from regcn_pair_scores_tf32: same arithmetic),
 *   loss[0] = mean(ce) when loss != NULL.  Deterministic (fixed slice order, single-CTA mean).                       */
REGCN_API int regcn_score_lse_num_parts(int N);
REGCN_API int regcn_score_lse_tf32(const float* q_hi, const float* q_lo, const float* e_hi, const float* e_lo, int B, int N,
                         int K, int hyp, const float* x2, const float* y2, const float* col_bias, double c,
                         const float* scale_margin, const float* row_c, int passes, float* part_max, float* part_sum,
                         void* stream);
REGCN_API int regcn_ce_from_lse(const float* part_max, const float* part_sum, int nparts, int B, const float* tscore,
                      float* ce, float* loss, void* stream);
/* dense variant for small candidate sets (relation prediction, N = 2R): ce[b] = logsumexp(S[b,:]) - S[b, triples[b,target_col]] */
REGCN_API int regcn_ce_rows(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, float* ce,
                  float* loss, void* stream);
REGCN_API int regcn_gather_rows2(const float* src_hi, const float* src_lo, const int32_t* idx, int P, int d,
                       float* out_hi, float* out_lo, void* stream);
REGCN_API int regcn_gather_scalars(const float* a, const float* b, const float* c, const int32_t* ia, const int32_t* ib,
                         int P, float* oa, float* ob, float* oc, void* stream);
REGCN_API int regcn_filter_correct(int B, const int32_t* filt_ptr, const int32_t* filt_idx, const int32_t* target,
                         const float* pair_score, const int32_t* raw_count, int col_lo, int col_hi,
                         int32_t* filt_count, const int32_t* filt_end, void* stream);

/* opt-in kernel timing with CUDA events on the launching stream: slot 0 = tcgen05 GEMM (work = 2MNK flops),
 * slot 1 = union aggregate.  enable(1) clears the records; read() synchronises the device and sums them.       */
REGCN_API void regcn_prof_enable(int on);
REGCN_API int regcn_prof_read(int slot, double* total_ms, long long* launches, double* total_work);

/* programmatic dependent launch between the kernels of the path (default on; REGCN_PDL=0 in the environment or
 * regcn_pdl_enable(0) falls back to plain stream-ordered launches) */
REGCN_API void regcn_pdl_enable(int on);
/* two-stream schedule of the evolve engines (default on; REGCN_TWO_STREAM=0 or regcn_two_stream_enable(0): every kernel
 * on the caller's stream, which is also how the per-kernel timings of bench.py's roofline block are taken) */
REGCN_API void regcn_two_stream_enable(int on);
/* Data flow of the all-entity GEMMs of regcn_regcn_evolve (sparse-snapshot form): 1 = the entity state is kept as ONE
 * fp32 copy and split to TF32 (hi, lo) on chip by the GEMM's converter warps (regcn_gemm_tf32_layer_a32), 0 = (hi, lo)
 * copies in HBM read by TMA, -1 (default) = by size: fp32 from 65 536 entity rows, where those GEMMs are HBM-bound
 * (REGCN_EVOLVE_A32 sets the initial mode).  Results are bit-identical either way.                                   */
REGCN_API void regcn_evolve_a32_mode(int mode);
/* number of kernels this library has launched in the calling process (the snapshot-index build adds ~8 CUB launches per
 * large snapshot that are not counted) */
REGCN_API long long regcn_kernel_launches(void);
/* tuning knob for experiments: force the N tile (multiple of 16, <= 256; 0 = automatic) and cap the pipeline depth */
REGCN_API void regcn_gemm_tf32_tune(int block_n, int stages);
/* Persistent-grid cap of the calling thread's next GEMM launches: at most `ctas` CTAs (0 = the whole machine).  The evolve
 * engine uses it to split the SMs between its two streams; profiles/ uses it to separate per-SM from chip-wide limits. */
REGCN_API void regcn_gemm_tf32_grid_cap(int ctas);
/* fp32-A variants of regcn_gemm_tf32 / regcn_gemm_tf32_layer: the A operand is ONE fp32 copy in memory,
 *   A[m, :] = [ a0[rows0 ? rows0[m] : m, 0:k0] | a1[rows1 ? rows1[m] : m, 0:k1] ]      (k1 = 0: one segment)
 * and is split into the TF32 (hi, lo) pair inside shared memory by converter warps of the GEMM kernel, bit-identically
 * to regcn_split_tf32.  B stays a pre-split weight (N, k0 + k1), K-major.  k0, k1, lda0, lda1 multiples of 4, 16-byte
 * aligned bases.  This is what the evolve engines use: activations exist once, as fp32 (torch.mm call sites of
 * rgcn/layers.py:222-255, src/rrgcn.py:168-178). */
REGCN_API int regcn_gemm_tf32_a32(const float* a0, int lda0, int k0, const int32_t* rows0, const float* a1, int lda1, int k1,
                                  const int32_t* rows1, const float* b_hi, const float* b_lo, int ldb, float* C, int ldc,
                                  int M, int N, const float* bias, int accumulate, int passes, int split_k, float* workspace,
                                  size_t workspace_bytes, const float* addend, int ld_add, void* stream);
REGCN_API int regcn_gemm_tf32_layer_a32(const float* a0, int lda0, int k0, const int32_t* rows0, const float* a1, int lda1,
                                        int k1, const int32_t* rows1, const float* b_hi, const float* b_lo, int ldb, int M,
                                        int N, int d, float* out_raw, float* out_hi, float* out_lo, float* gate_out,
                                        int ld_gate_out, const int32_t* row_idx, const int32_t* skip_rows,
                                        const float* gate_G, int gate_ld, const float* gate_bias, const float* gate_h,
                                        int gate_norm, void* stream);
/* in-kernel timeline of the tcgen05 GEMM (profiles/gemm_trace.py): while a device buffer of
 * gridDim.x * regcn_gemm_tf32_trace_slots() uint64 is attached, one thread per warp role stamps %globaltimer at the
 * pipeline hand-over points of its CTA (entry, dependency wait, per tile: loads issued, first operands landed, MMAs
 * issued, accumulator complete, epilogue done).  NULL detaches.  Not thread-safe; a measurement aid only. */
REGCN_API void regcn_gemm_tf32_trace(void* dev_buf);
/* regcn_score_count_tf32 with a hyperbolic (RotH / MuRP form) score and no candidate bias decides "beats the target"
 * with a division- and sqrt-free polynomial threshold test per candidate and evaluates the IEEE score
 * (hyperbolic_decoder.py:89-179) only inside its rounding band; counts are identical either way.  0 switches the test
 * off (every candidate takes the IEEE score): the yardstick of tests and bench.py. Default 1. */
REGCN_API void regcn_score_count_poly(int on);
REGCN_API int regcn_gemm_tf32_trace_slots(void);
/* multi-launch session (bench.py's roofline): record r of the buffer = 148 CTAs x regcn_gemm_tf32_trace_slots() uint64;
 * launch i of the session stamps into record i (launches beyond bytes / record size are not recorded) and the library
 * remembers what it computed.  Per launch: duration = max over CTAs of slot 40 (exit) - min over CTAs of slot 1 (past
 * the dependency wait), both %globaltimer nanoseconds -- measured inside the kernel, so streams, programmatic dependent
 * launch and warm caches are those of the timed region.  regcn_gemm_tf32_trace_begin(NULL, 0) ends the session. */
REGCN_API void regcn_gemm_tf32_trace_begin(void* dev_buf, size_t bytes);
REGCN_API int regcn_gemm_tf32_trace_count(void);
REGCN_API int regcn_gemm_tf32_trace_read(int i, int* epi, int* M, int* N, int* K, int* grid, int* passes, double* flops);
/* the layer GEMM of the evolve engine on its own (UnionRGCNLayer apply step, rgcn/layers.py:247-255, and for the last
 * layer the time gate, src/rrgcn.py:176-178, straight out of the accumulator):
 *   acc = A[M,K] . B[N,K]^T (3xTF32);  columns [0,d): out row (row_idx ? row_idx[m] : m) = rrelu(acc) unless
 *   skip_rows[m] >= 0, written as fp32 (out_raw) and / or as the TF32 split (out_hi, out_lo); columns [d,N) are stored
 *   unchanged to gate_out[:, col-d].  gate_G != NULL (needs N == d): out = s * [normalize](rrelu(acc)) + (1-s) * gate_h,
 *   s = sigmoid(gate_G + gate_bias). */
REGCN_API int regcn_gemm_tf32_layer(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo,
                                    int ldb, int M, int N, int K, int d, float* out_raw, float* out_hi, float* out_lo,
                                    float* gate_out, int ld_gate_out, const int32_t* row_idx, const int32_t* skip_rows,
                                    const float* gate_G, int gate_ld, const float* gate_bias, const float* gate_h,
                                    int gate_norm, void* stream);
/* edge kernel variant: 0 automatic, 1 register-staged gathers, 2 cp.async.bulk gathers staged in shared memory */
REGCN_API void regcn_aggregate_tune(int impl);

/* ---- row maps: F.normalize / tanh / log_0 / exp_0 / project (hyperbolic_ops.py:38-116) ---------
 * mode 0 normalize, 1 tanh, 2 0.9 tanh(log_0 x)+0.1 log_0 x, 3 log_0, 4 exp_0, 5 project,
 * 6 exp_0(normalize(log_0 x)), 7 identity, 8 exp_0(rrelu(x)); sumsq (optional, M): |out|^2 per row; out may be NULL
 * when only sumsq is wanted.                                                                    */
REGCN_API int regcn_row_map(const float* x, float* out, int M, int d, int mode, double c, float* sumsq, void* stream);
/* the same map, also emitting the TF32 split (out_hi, out_lo) of the result for a consumer GEMM (no separate split pass);
 * mode 9 = tanh(x / max(|x|, 1e-12)): predict-time F.normalize (src/rrgcn.py:190) and the ConvTransE activation
 * (src/decoder.py:79) in one pass */
REGCN_API int regcn_row_map_split(const float* x, float* out, float* out_hi, float* out_lo, int M, int d, int mode, double c,
                        void* stream);

/* ---- K3 GRU gates: nn.GRUCell at src/rrgcn.py:133,168-174; hyperbolic_model.py:408,815-824 ----- */
REGCN_API int regcn_gru_gate(const float* gi, const float* gh, const float* hprev, float* out, int M, int d,
                   int normalize, void* stream);

/* ---- K5 self-loop combine: rgcn/layers.py:226-255; hyperbolic_layers.py:273-323,649-694 -------- */
REGCN_API int regcn_union_combine(const float* P, const float* L, const int32_t* indeg, const float* S,
                        const float* skip_bias, const float* prev, int N, int d, int act, int hyper, double c,
                        float* out, float* ht_next, float* radius_next, void* stream);

/* ---- K9 time gate (Euclidean): src/rrgcn.py:176-178 -------------------------------------------- */
REGCN_API int regcn_time_gate(const float* G, const float* bias, const float* cur, const float* h, float* out, int N,
                    int d, int normalize_cur, void* stream);

/* ---- K8/K9 hyperbolic: hyperbolic_model.py:715-720,773-782,802,829-867; hyperbolic_ops.py:395-435 */
REGCN_API int regcn_hyp_init(const float* emb, const float* radius_static, int N, int d, int normalize, int on_manifold,
                   double c, float radius_min, float radius_max, float* out, void* stream);
REGCN_API int regcn_hyp_tangent(const float* h, int N, int d, double c, float* ht, float* pt, float* radius, void* stream);
REGCN_API int regcn_hyp_time_gate(const float* h2, const float* pt, const float* G, const float* bias,
                        const float* radius_static, const float* radius_w, float radius_b, int N, int d,
                        int layer_norm, int residual, double c, float radius_min, float radius_max, float beta,
                        float eps_r, float* out, void* stream);

/* ---- K10 ConvTransE/ConvTransR tower: src/decoder.py:29-52,78-95; hyperbolic_decoder.py:376-406 ---
 * F (B, C*d) raw features and/or their TF32 split (F_hi, F_lo) for the FC GEMM; either may be NULL.          */
REGCN_API int regcn_convtranse_features(const float* ent, const float* second, const int64_t* triples, int col0, int col1,
                              int B, int d, int C, int ksz, const float* bn0_scale, const float* bn0_shift,
                              const float* conv_w, const float* conv_b, const float* bn1_scale,
                              const float* bn1_shift, float* F, float* F_hi, float* F_lo, void* stream);
/* The tower up to and including the fully-connected layer in ONE GEMM (src/decoder.py:81-93: bn0 -> conv1d(2 -> C, k = 3) ->
 * bn1 -> relu -> fc): out (B, N) = W_fc . features + bias with the (B, C d) feature map computed inside the operand ring of
 * the tcgen05 GEMM instead of written and read back.  w_hi / w_lo (N, ldw >= 16 C ceil(d / 16)) = fc.weight in the order
 * the reduction is walked in, blocks of 16 positions outermost (regcn_convtrans_fc_pack_weight, once per weight).  Split-K over
 * equal shares of the reduction; batch_total >= B = size of the whole query batch when B rows are one slice of it (the
 * split follows the whole batch so that a row does not depend on the cut; 0 = B).  ws: regcn_convtrans_fc_workspace_bytes.
 * Tail (src/decoder.py:92-95), folded into the split-K reduction: out = [relu]([bn2_scale *] (fc + bias) [+ bn2_shift]),
 * bn2_* NULL = no BatchNorm (the B == 1 rule); out_hi / out_lo (B, N), optional = TF32 split of out for the scoring GEMM.
 * REGCN_ERR_UNSUPPORTED unless kernel size 3, d % 4 == 0, C <= 64, N % 4 == 0 (callers then use regcn_convtranse_features +
 * regcn_gemm_tf32_a32).                                                                                             */
REGCN_API int regcn_convtrans_fc_pack_weight(const float* fc_weight, int N, int C, int d, float* w_hi, float* w_lo, void* stream);
REGCN_API size_t regcn_convtrans_fc_workspace_bytes(int batch_total, int N);
REGCN_API int regcn_convtrans_fc(const float* x0, const float* x1, const int64_t* triples, int col0, int col1, int B,
                       int batch_total, int d, int C, int ksz, const float* bn0_scale, const float* bn0_shift,
                       const float* conv_w, const float* conv_b, const float* bn1_scale, const float* bn1_shift,
                       const float* w_hi, const float* w_lo, int ldw, int N, const float* bias, const float* bn2_scale,
                       const float* bn2_shift, int relu, float* out, int ldc, float* out_hi, float* out_lo, float* ws,
                       size_t ws_bytes, void* stream);
REGCN_API int regcn_affine_relu(float* x, const float* scale, const float* shift, int M, int d, int relu, void* stream);

/* ---- K12 RotH / MuRP / RotHRel query builder: hyperbolic_decoder.py:744-765,1064-1086,1243-1251 - */
REGCN_API int regcn_gather_log0(const float* E, const int64_t* triples, int col, int B, int d, int project, double c,
                      float* out, void* stream);
REGCN_API int regcn_hyp_query(const float* s_tan, const float* ang, const float* trans, const float* E,
                    const int64_t* triples, int B, int d, int kind, double c, float* Q, float* q_sumsq,
                    void* stream);

/* ---- K13 hyperbolic score epilogue on a dense <q,e> matrix: hyperbolic_decoder.py:89-179 --------
 * bias (N) candidate bias, qbias (B) per-query bias (entity_bias[subject], :1097-1098), row_c (B) per-query
 * curvature selecting the artanh true-distance branch (:145-163); each may be NULL.                 */
REGCN_API int regcn_hyp_score_epilogue(float* S, int ld, int B, int N, const float* q_sumsq, const float* e_sumsq,
                             const float* bias, const float* qbias, double c, const float* scale_margin,
                             const float* row_c, void* stream);
/* per-query curvature c_q[b] = max(1e-5, min(softplus(raw[r_b mod R]), 0.999 c, cmax)): hyperbolic_decoder.py:66-86,1020-1026
 * (cmax <= 0: no warm-up bound) */
REGCN_API int regcn_rel_curvature(const float* raw, const int64_t* triples, int B, int R, double c, double cmax,
                                  float* out, void* stream);

/* ---- K14 rank / filter: rgcn/utils.py:21-25,51-75,136-166 --------------------------------------
 * S (B, ld) scores of a shard of N candidate columns starting at global column col_offset.
 * filt_ptr (B+1)/filt_idx: per-query sorted global ids of the other true answers (all_ans).      */
REGCN_API int regcn_gather_target_score(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                              int col_offset, float* target_score, void* stream);
REGCN_API int regcn_rank_count(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                     const int32_t* filt_ptr, const int32_t* filt_idx, int col_offset, const float* target_score,
                     int32_t* raw_count, int32_t* filt_count, const int32_t* filt_end, void* stream);
REGCN_API int regcn_counts_to_ranks(const int32_t* raw_count, const int32_t* filt_count, int B, int64_t* rank,
                          int64_t* filt_rank, void* stream);
REGCN_API int regcn_apply_filter(float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                       const int32_t* filt_ptr, const int32_t* filt_idx, int col_offset, const int32_t* filt_end,
                       void* stream);
/* filter lists of query b are filt_idx[filt_ptr[b] .. filt_end[b]) (filt_end == NULL: .. filt_ptr[b+1]).
 * regcn_filter_count / regcn_filter_fill build them from the query triples themselves (time-aware filtering,
 * rgcn/utils.py:264-304): key = (col 0, key_col), answers = ans_col; counts -> exclusive scan (caller) -> fill,
 * which also emits the (query, candidate) pair lists of the fused rank path (B target pairs first).           */
/* One-call preparation of a test snapshot's queries (src/main.py:60-74, src/rrgcn.py:184-186): all_t (2T,3) = the triples
 * followed by their inverses (o, r + R, s); counts (2, 2T) = regcn_filter_count of all_t for key_col 1 (entity filter) and 2
 * (relation filter); beg (2, 2T) = their exclusive scans (the list offsets regcn_filter_fill takes); totals (2) = slot totals. */
REGCN_API int regcn_queries_prepare(const int64_t* triples, int T, int R, int64_t* all_t, int32_t* counts, int32_t* beg,
                                    int32_t* totals, void* stream);
/* The same for the n <= 32 test snapshots of a group of timestamps in three launches (regcn_b200.test() evolves consecutive
 * timestamps together and prepares a group at once).  triples_cat (toff[n], 3): the snapshots back to back; toff (n+1, HOST):
 * rows in front of snapshot g, toff[0] = 0, every snapshot non-empty.  Snapshot g's outputs: all_t_cat + 3 * 2 toff[g]
 * (2 T_g rows), counts_cat / beg_cat + 4 toff[g] (each (2, 2 T_g)), totals + 2 g -- array for array what
 * regcn_queries_prepare writes for that snapshot alone.                                                              */
REGCN_API int regcn_queries_prepare_batch(const int64_t* triples_cat, const int32_t* toff, int n, int R, int64_t* all_t_cat,
                                int32_t* counts_cat, int32_t* beg_cat, int32_t* totals, void* stream);
REGCN_API int regcn_filter_count(const int64_t* triples, int B, int key_col, int32_t* counts, void* stream);
REGCN_API int regcn_filter_fill(const int64_t* triples, int B, int key_col, int ans_col, const int32_t* beg, int32_t* idx,
                      int32_t* end, int32_t* pair_a, int32_t* pair_e, void* stream);
/* Both filter lists of a timestamp's queries in ONE launch (what regcn_b200.test() needs per timestamp, src/main.py:71-74:
 * entity ranks filtered by key (h, r) -> answers t, relation ranks by key (h, t) -> answers r): the *_e arguments are
 * regcn_filter_fill(key_col 1, ans_col 2), the *_r arguments regcn_filter_fill(key_col 2, ans_col 1); identical lists.  */
REGCN_API int regcn_filter_fill2(const int64_t* triples, int B, const int32_t* beg_e, int32_t* idx_e, int32_t* end_e,
                       int32_t* pair_a_e, int32_t* pair_e_e, const int32_t* beg_r, int32_t* idx_r, int32_t* end_r,
                       int32_t* pair_a_r, int32_t* pair_e_r, void* stream);

/* ---- whole-recurrence orchestration: RecurrentRGCN.forward, src/rrgcn.py:142-180 (uvrgcn, self_loop, no skip) ---
 * One call enqueues every kernel of the L-snapshot recurrence on `stream`.  Inputs are pointer / int tables:
 *   model_ptrs[RM_*]   device pointers to parameters; GEMM weights are passed K-major ([out, in]) and TF32-split
 *                      (regcn_split_tf32): W_ih[:, d:], W_hh, per layer W_n^T and [W_loop | W_evolve (| W_time)]^T;
 *                      RM_GI_STATIC = emb_rel . W_ih[:, :d]^T + b_ih  (2R, 3d), constant per model
 *   model_ints[RMI_*]  N, 2R, d, n_layers, layer_norm, self_loop
 *   graph_ptrs         L consecutive blocks of RG_NUM_PTRS pointers (outputs of regcn_csr_build)
 *   graph_ints         L consecutive blocks of RGI_NUM_INTS ints
 * Outputs: hist (L, N, d) = history_embs, h0_out (2R, d) = evolved relation embeddings.                        */
enum { RM_DYNAMIC_EMB = 0, RM_EMB_REL, RM_EMB_REL_HI, RM_EMB_REL_LO, RM_GI_STATIC, RM_WIH_R_HI, RM_WIH_R_LO,
       RM_WHH_HI, RM_WHH_LO, RM_B_HH, RM_GATE_BIAS, RM_LAYER0, RM_LAYER_STRIDE = 8
       /* RM_LAYER0 + 8*l: W_n^T hi,lo | [W_loop|W_evolve(|W_time)]^T hi,lo | [W_n;W_loop]^T hi,lo | [W_evolve(|W_time)]^T hi,lo */ };
enum { RMI_NUM_ENTS = 0, RMI_NUM_RELS2, RMI_DIM, RMI_NUM_LAYERS, RMI_LAYER_NORM, RMI_SELF_LOOP, RMI_NUM_INTS };
enum { RG_ROWPTR = 0, RG_SRC_SORTED, RG_ETYPE_SORTED, RG_INDEG, RG_NORM, RG_VPTR, RG_SPTR, RG_VROW_ROW,
       RG_REL_ROWPTR, RG_REL_ENTS, RG_ACTIVE_POS, RG_ACTIVE_ROWS, RG_NUM_PTRS };
enum { RGI_NUM_EDGES = 0, RGI_N_VROWS, RGI_N_SPLIT_CHUNKS, RGI_N_REL_ENTS, RGI_N_ACTIVE, RGI_MAX_CHUNKS, RGI_NUM_INTS };
REGCN_API size_t regcn_regcn_evolve_workspace_bytes(int N, int R2, int d, int max_split_chunks, int rel_nsplit);
REGCN_API int regcn_regcn_evolve(const void* const* model_ptrs, const int* model_ints, const void* const* graph_ptrs,
                       const int* graph_ints, int L, float* hist, float* h0_out, int rel_nsplit, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Shared-trajectory form of the same recurrence for G history windows of ONE model evolved together
 * (RecurrentRGCN.forward_batch; the evaluation loop of src/main.py:60-90 re-runs src/rrgcn.py:142-180 per test timestamp
 * over windows that do not depend on each other).  Tables and graphs are those of regcn_regcn_evolve in the numbering of
 * regcn_csr_concat (N = G N0 entity rows, entity (g, v) = row g N0 + v).  A row without in-edges is updated from its own
 * state only and all windows start from one table, so the all-entity products run over N0 shared rows plus the rows that
 * have been active in their window so far (compact, bit-identical per row) instead of G N0 rows.  Output: h_final (N, d) =
 * history_embs[-1] of every window, h0_out as above; intermediate history_embs are not produced.  sum_active / max_active
 * = sum / maximum of graph_ints[RGI_N_ACTIVE] over the L snapshots.  REGCN_ERR_UNSUPPORTED (nothing enqueued) unless
 * n_layers >= 2, self_loop and every snapshot has n_active <= N / 2: callers then use regcn_regcn_evolve.            */
REGCN_API size_t regcn_regcn_evolve_shared_workspace_bytes(int N0, int G, int R2, int d, int max_split_chunks, int rel_nsplit,
                                                           long long sum_active, int max_active);
REGCN_API int regcn_regcn_evolve_shared(const void* const* model_ptrs, const int* model_ints, const void* const* graph_ptrs,
                              const int* graph_ints, int L, int G, float* h_final, float* h0_out, int rel_nsplit,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---- whole-recurrence orchestration, hyperbolic: HyperbolicRecurrentRGCN.forward, hyperbolic_model.py:722-890
 * (encoders hyperbolic_uvrgcn = 0 and lgcn = 1, self_loop, no skip connection, fixed curvature).  Same calling
 * convention as regcn_regcn_evolve plus a double table for the scalars.                                         */
enum { HM_DYNAMIC_EMB = 0, HM_RADIUS_STATIC, HM_EMB_REL, HM_EMB_REL_HI, HM_EMB_REL_LO, HM_GI_STATIC, HM_WIH_R_HI,
       HM_WIH_R_LO, HM_WHH_HI, HM_WHH_LO, HM_B_HH, HM_GATE_W_HI, HM_GATE_W_LO, HM_GATE_BIAS, HM_RADIUS_W, HM_LAYER0,
       HM_LAYER_STRIDE = 4
       /* HM_LAYER0 + 4*l: uvrgcn: W_n^T hi, lo | lgcn: block weight (raw), unused ; then [W_loop|W_evolve]^T hi, lo */ };
enum { HMI_NUM_ENTS = 0, HMI_NUM_RELS2, HMI_DIM, HMI_NUM_LAYERS, HMI_LAYER_NORM, HMI_SELF_LOOP, HMI_ENCODER,
       HMI_NUM_BASES, HMI_RESIDUAL, HMI_NUM_INTS };
enum { HMD_C = 0, HMD_GAMMA, HMD_RMIN, HMD_RMAX, HMD_BETA, HMD_EPS_R, HMD_RADIUS_BIAS, HMD_NUM };
REGCN_API size_t regcn_hyp_evolve_workspace_bytes(int N, int R2, int d, int max_split_chunks, int rel_nsplit);
REGCN_API int regcn_hyp_evolve(const void* const* model_ptrs, const int* model_ints, const double* model_doubles,
                     const void* const* graph_ptrs, const int* graph_ints, int L, float* hist, float* h0_out,
                     int rel_nsplit, void* workspace, size_t workspace_bytes, void* stream);

/* =================================================================================================================
 * Training step: RecurrentRGCN.get_loss in train mode, src/rrgcn.py:197-223, and the
 * optimisation step of src/main.py:235-246 (backward, clip_grad_norm_(1.0), Adam(lr, weight_decay)).
 * All reductions are fixed-order (no float atomics): a training step is bit-reproducible.
 * ================================================================================================================= */

/* CSR gather-sum: out[row] (+)= row_w[row] * sum_{j in row} col_w[col_j] * (X[col_j] (+ X[col_j + col2_off])).
 * The backward of every gather on the path: UnionRGCNLayer message (rgcn/layers.py:257-279; the snapshot graph holds
 * each edge with its inverse, so the forward CSR-by-destination is also the CSR-by-source), the relation table
 * gather (same lines), the relation mean-pool (src/rrgcn.py:161-166, col2_off = R) and the decoder's E[s] / rel[r]
 * gathers (src/decoder.py:81-82).  col_w / row_w may be NULL (= 1).  rho != NULL multiplies every term by the
 * radius-difference edge weight exp(-gamma |rho[col_j] - rho[other]|) of the hyperbolic layers
 * (hyperbolic_layers.py:232-234), other = partner[j] (per CSR position) or the row itself when partner is NULL.  */
REGCN_API int regcn_csr_gather_sum(const float* X, int ldx, const float* col_w, const float* row_w, const int32_t* rowptr,
                         const int32_t* col, int nrows, int d, int col2_off, float* out, int ldo, int accumulate,
                         const float* rho, float gamma, const int32_t* partner, void* stream);
/* Stable grouping of n int32 keys in [0,nkeys): rowptr (nkeys+1), perm (n) original positions grouped by key,
 * vals_out[i] = vals[perm[i]] (vals may be NULL).  Builds the transposed indices the gathers above run on.      */
REGCN_API size_t regcn_group_by_key_workspace_bytes(int n);
REGCN_API int regcn_group_by_key(const int32_t* keys, int n, int nkeys, const int32_t* vals, int32_t* rowptr, int32_t* perm,
                       int32_t* vals_out, void* workspace, size_t workspace_bytes, void* stream);
/* rowid[i] = CSR row of position i (nnz entries); inv_len[r] = 1/len(row r) or 0 (either output may be NULL).   */
REGCN_API int regcn_expand_rowptr(const int32_t* rowptr, int nrows, int nnz, int32_t* rowid, float* inv_len, void* stream);

/* F.normalize backward (src/rrgcn.py:154,170,176,206): dx = (dy - y <y,dy>) / max(|x|,1e-12).                   */
REGCN_API int regcn_normalize_bwd(const float* x, const float* dy, float* dx, int M, int d, void* stream);
/* nn.GRUCell backward from the saved pre-activations gi, gh (M,3d) (src/rrgcn.py:133,168-174), optional F.normalize
 * of the output folded in: dgi, dgh (M,3d), dhprev (M,d).                                                        */
REGCN_API int regcn_gru_gate_bwd(const float* gi, const float* gh, const float* hprev, const float* dout, int M, int d,
                       int normalize, float* dgi, float* dgh, float* dhprev, void* stream);
/* UnionRGCNLayer apply-step backward (rgcn/layers.py:241-253): out = dropout_p(rrelu(P + where(indeg>0, L0, L1)));
 * masks recovered from `out` (out == NULL: no activation / dropout, dP = dout).  dP (N,d); dL (N,2d) =
 * [indeg>0 ? dP : 0 | indeg>0 ? 0 : dP] (may be NULL).                                                              */
REGCN_API int regcn_union_combine_bwd(const float* out, const float* dout, const int32_t* indeg, int N, int d, float p,
                            float* dP, float* dL, void* stream);
/* time gate backward (src/rrgcn.py:176-178): dG (pre-sigmoid), dcur (through the optional F.normalize), dh (direct). */
REGCN_API int regcn_time_gate_bwd(const float* G, const float* bias, const float* cur, const float* h, const float* dout,
                        int N, int d, int normalize_cur, float* dG, float* dcur, float* dh, void* stream);
/* dx = dy * (1 - y^2), y = tanh(x)  (src/decoder.py:30,79); n % 4 == 0                                           */
REGCN_API int regcn_tanh_bwd(const float* y, const float* dy, float* dx, size_t n, void* stream);
/* in-place nn.Dropout(p) in train mode (counter-based mask: element i of call `seed`)                            */
REGCN_API int regcn_dropout(float* x, size_t n, float p, uint32_t seed, void* stream);

/* BatchNorm1d, train mode (src/decoder.py:20-22,68-70): X viewed as (B,C,L).  Two fixed-order stages (row slabs, then
 * channels in double).  workspace: regcn_col_reduce_workspace_bytes(B, C*L).                                      */
REGCN_API size_t regcn_col_reduce_workspace_bytes(int rows, int cols);
REGCN_API int regcn_bn_stats(const float* X, int B, int C, int L, float eps, float momentum, float* mean, float* invstd,
                   float* running_mean, float* running_var, float* workspace, size_t workspace_bytes, void* stream);
/* sums of the BatchNorm backward: sum_dy (C) = dbeta, sum_dy_xhat (C) = dgamma; dy = mask(dZ; Z): mask_mode 0 none,
 * 1 Z>0 (relu + dropout), 2 Z!=0 (dropout), times mask_scale.  Y = the BatchNorm input.                           */
REGCN_API int regcn_bn_bwd_stats(const float* dZ, const float* Z, const float* Y, int B, int C, int L, int mask_mode,
                       float mask_scale, const float* mean, const float* invstd, float* sum_dy, float* sum_dy_xhat,
                       float* workspace, size_t workspace_bytes, void* stream);
/* out (cols) (+)= column sums of X (rows, cols; pitch ld): bias gradients                                        */
REGCN_API int regcn_col_sum(const float* X, int ld, int rows, int cols, float* out, int accumulate, float* workspace,
                  size_t workspace_bytes, void* stream);
/* out = dropout_p(relu?(gamma (x-mean) invstd + beta)); mean == NULL skips the normalisation                    */
REGCN_API int regcn_bn_act_drop(const float* X, int B, int C, int L, const float* mean, const float* invstd,
                      const float* gamma, const float* beta, int relu, float p, uint32_t seed, float* out, void* stream);
/* dX = gamma invstd (dy - sum_dy/n - xhat sum_dy_xhat/n) with dy = mask(dZ; Z); then mask(.; out_src) of the dropout
 * that precedes the BatchNorm input.  mean == NULL: masks only.                                                   */
REGCN_API int regcn_bn_bwd_apply(const float* dZ, const float* Z, const float* Y, int B, int C, int L, int mask_mode,
                       float mask_scale, const float* mean, const float* invstd, const float* gamma,
                       const float* sum_dy, const float* sum_dy_xhat, const float* out_src, int out_mode,
                       float out_scale, float* dX, void* stream);

/* ConvTransE / ConvTransR tower in train mode (src/decoder.py:29-52, 78-100): stacked inputs, bn0 + input dropout +
 * Conv1d(2,C,3,pad 1) with the dropped input kept, and the three convolution gradients.                          */
REGCN_API int regcn_dec_gather_stack(const float* first, const float* second, const int64_t* triples, int col0, int col1,
                           int B, int d, float* X0, void* stream);
REGCN_API int regcn_dec_conv_fwd(const float* X0, int B, int d, int C, int ksz, const float* mean0, const float* invstd0,
                       const float* gamma0, const float* beta0, float p, uint32_t seed, const float* W,
                       const float* bias, float* X1, float* Y, void* stream);
REGCN_API int regcn_dec_conv_bwd_input(const float* dY, const float* X1, int B, int d, int C, int ksz, const float* W,
                             float p, float* dX1, void* stream);
REGCN_API size_t regcn_dec_conv_bwd_weight_workspace_bytes(int B, int C);
/* dW_db: C*2*3 weight gradients followed by C bias gradients                                                     */
REGCN_API int regcn_dec_conv_bwd_weight(const float* dY, const float* X1, int B, int d, int C, int ksz, float* dW_db,
                              float* workspace, size_t workspace_bytes, void* stream);

/* nn.CrossEntropyLoss over materialised logits (src/rrgcn.py:87-88,218-223): ce (B), lse (B), loss (1) = mean(ce);
 * and its gradient in place: S <- (softmax(S) - onehot) * (*gscale) / B, pitch padding zeroed.                   */
REGCN_API int regcn_ce_lse_rows(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                      float* ce, float* lse, float* loss, void* stream);
REGCN_API int regcn_softmax_grad_rows(float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col,
                            const float* lse, const float* gscale, void* stream);
/* out[c][r] = X[r][c] for r < rows, 0 for rows <= r < ldo (a K-major GEMM operand with K = ldo); plain and/or TF32
 * (hi, lo) outputs: the weight-gradient GEMMs dW = x^T dy                                                        */
REGCN_API int regcn_transpose_split(const float* X, int rows, int cols, int ldx, float* out, float* out_hi, float* out_lo,
                          int ldo, void* stream);

/* clip_grad_norm_ + torch.optim.Adam over one flat fp32 buffer (src/main.py:194,243-246).  total_norm: device float.
 * step counts from 1.  max_norm <= 0 or total_norm == NULL: no clipping.                                          */
REGCN_API size_t regcn_adam_workspace_bytes(void);
REGCN_API int regcn_grad_norm(const float* g, size_t n, float* total_norm, void* workspace, size_t workspace_bytes,
                    void* stream);
REGCN_API int regcn_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float max_norm, const float* total_norm, void* stream);

/* Static-graph constraint in training (src/rrgcn.py:101-106,146-152,225-247): RGCNBlockLayer backward (dh over the
 * forward CSR with inverse relation types; dW over edges grouped by type: type_rowptr (R2+1), type_src / type_dst) and
 * the angle loss: term[row] = weight * max(cos_step - sim(row), 0) per history step, and its gradient.            */
REGCN_API size_t regcn_block_aggregate_bwd_workspace_bytes(int R2, int d_in, int d_out, int num_bases);
REGCN_API int regcn_block_aggregate_bwd(const float* h, const float* dAgg, const float* W, const int32_t* rowptr,
                              const int32_t* src_sorted, const int32_t* etype_sorted, const float* norm,
                              const int32_t* type_rowptr, const int32_t* type_src, const int32_t* type_dst, int N, int R2,
                              int d_in, int d_out, int num_bases, float* dh, float* dW, float* workspace,
                              size_t workspace_bytes, void* stream);
REGCN_API int regcn_static_angle_fwd(const float* S, const float* E, int N, int d, float cos_step, float weight,
                           int layer_norm, float* term, void* stream);
REGCN_API int regcn_static_angle_bwd(const float* S, const float* E, int N, int d, float cos_step, float weight,
                           int layer_norm, const float* gscale, float* dS, int accumulate_dS, float* dE, void* stream);

/* Multi-step inference (src/main.py:90-97; construct_snap / construct_snap_r, rgcn/utils.py:367-405): top_idx (B,K) =
 * the K best candidates of every score row in descending order, ties by ascending id (a stable descending sort);
 * out (B*K,3) int64 = the predicted snapshot (may be NULL).  rel_mode 0: entity scores, 1: relation scores.        */
REGCN_API int regcn_topk_construct_snap(const float* S, int64_t ld, int B, int N, int K, const int64_t* triples, int R,
                              int rel_mode, int32_t* top_idx, int64_t* out, void* stream);

/* AttH / AttHRel query builders (hyperbolic_decoder.py:1403-1480, 1593-1640): attention-weighted mix of a Givens
 * rotation and a Givens reflection of the subject's tangent vector.  mode 0: per-relation tables rot, ref (2R,d/2),
 * attn (2R,2d), rel (2R,d), trans (2R,d), query = project(exp_0(mix)) (+)_c project(exp_0(trans[r]));
 * mode 1: global rot, ref (d/2), attn (2d), query = (-exp_0(mix)) (+)_c E[o].  Q (B,d), q_sumsq (B) optional.        */
REGCN_API int regcn_atth_query(const float* s_tan, const float* rot, const float* ref, const float* attn, const float* rel,
                     const float* trans, const float* E, const int64_t* triples, int B, int d, int mode, double c,
                     float* Q, float* q_sumsq, void* stream);

/* regcn_gemm_tf32 with operands in their natural row-major layouts, fed to tcgen05 as MN-major tiles (no transposes):
 *   a_mn = 1: A is passed as X (K,M), the product uses X^T;   b_mn = 1: B is passed as Y (K,N), the product uses Y
 *   (b_mn = 0 keeps regcn_gemm_tf32's B (N,K), i.e. B^T).  So (a_mn, b_mn) = (1,1): C = X^T Y, the weight gradient
 *   dW = x^T dy of every torch.mm / F.linear on the path (src/rrgcn.py:168-178, rgcn/layers.py:229-233,257-276,
 *   src/decoder.py:90);  (0,1): C = A Y, i.e. torch.mm(x, W) and dX = dY W without transposing W.
 * TF32 (hi, lo) pairs (lo NULL with passes == 1); split-K workspace: regcn_gemm_tf32_workspace_bytes(M, N, split_k). */
REGCN_API int regcn_gemm_tf32_mn(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb,
                       float* C, int ldc, int M, int N, int K, int a_mn, int b_mn, const float* bias, int accumulate,
                       int passes, int split_k, float* workspace, size_t workspace_bytes, void* stream);

/* One-call decode + rank of an evaluated timestamp for the ConvTransE / ConvTransR pair (src/rrgcn.py:190-193,
 * src/decoder.py:29-52,78-100, rgcn/utils.py:136-166 as called from src/main.py:71-74): [F.normalize] -> tanh(E) ->
 * entity tower -> fused score/count over all N entities + filter correction -> relation tower -> (B,2R) scores ->
 * raw/filtered relation ranks.  packed (4*B int32) = [rank | filter_rank | rank_rel | filter_rank_rel], 1-based.
 * tower_ent / tower_rel: HOST arrays of 13 device pointers {bn0 scale, bn0 shift, conv weight (C,2,k), conv bias,
 * bn1 scale, bn1 shift, fc weight hi, fc weight lo (d, C*d), fc bias, bn2 scale, bn2 shift, fc weight hi, lo packed by
 * regcn_convtrans_fc_pack_weight} (eval-mode BatchNorm folded to scale/shift; the packed pair is read when the tower runs
 * as regcn_convtrans_fc: kernel size 3, d % 4 == 0, C <= 64, REGCN_FUSED_TOWER != 0 -- else it may repeat the plain pair).  fe_* / fr_*: the entity / relation filter lists (ptr, idx, end as built by regcn_filter_count /
 * regcn_filter_fill); pair_a / pair_e (P >= B): the (query, candidate) pairs regcn_filter_fill emits.
 * Same kernels and arithmetic as the per-op entry points, so ranks are bit-identical to calling those in sequence. */
REGCN_API size_t regcn_convtrans_decode_rank_workspace_bytes(int N, int R2, int d, int B, int C, int P);
REGCN_API int regcn_convtrans_decode_rank(const float* emb, const float* r_emb, const int64_t* triples,
                                const void* const* tower_ent, const void* const* tower_rel, const int32_t* fe_ptr,
                                const int32_t* fe_idx, const int32_t* fe_end, const int32_t* pair_a,
                                const int32_t* pair_e, int P, const int32_t* fr_ptr, const int32_t* fr_idx,
                                const int32_t* fr_end, int N, int R2, int d, int B, int C, int ksz, int layer_norm,
                                int32_t* packed, void* workspace, size_t workspace_bytes, void* stream);

/* =================================================================================================================
 * Hyperbolic training step: HyperbolicRecurrentRGCN.get_loss in train mode, hyperbolic_uvrgcn encoder +
 * hyperbolic_convtranse decoder (hyperbolic_model.py:722-890,941-1088; hyperbolic_layers.py:222-323;
 * hyperbolic_ops.py:38-233,395-435; hyperbolic_decoder.py:360-413).
 * ================================================================================================================= */
/* Backward of the radial Poincare row maps y = s(|x|) x: mode 0 log_0, 1 exp_0 (with its projection), 2 project_to_ball,
 * 3 exp_0(F.normalize(log_0 x)).  dx = s dy + s'(n)/n <x,dy> x with the reference's norm clamps (s' = 0 where active). */
REGCN_API int regcn_radial_bwd(const float* x, const float* dy, float* dx, int M, int d, int mode, double c, void* stream);
/* get_radius (hyperbolic_ops.py:206) and apply_radius (:222-233), forward and backward                             */
REGCN_API int regcn_row_radius(const float* x, int M, int d, float* rho, void* stream);
REGCN_API int regcn_row_radius_bwd(const float* x, const float* drho, int M, int d, float* dx, void* stream);
REGCN_API int regcn_apply_radius(const float* x, const float* r, int M, int d, double c, float* y, void* stream);
REGCN_API int regcn_apply_radius_bwd(const float* x, const float* r, const float* dy, int M, int d, double c, float* dx,
                           float* dr, void* stream);
/* elementwise op 0: clamp(x, -lim, lim) (the +-10 tangent clamps); op 1: 0.9 tanh(x) + 0.1 x (hyperbolic_decoder.py:378);
 * op 2: -x; op 3: relu(x)                                                                                          */
REGCN_API int regcn_eltwise_fwd(const float* x, float* y, size_t n, int op, float lim, void* stream);
REGCN_API int regcn_eltwise_bwd(const float* x, const float* dy, float* dx, size_t n, int op, float lim, void* stream);
/* _static_radius (hyperbolic_model.py:715-720) + TemporalRadiusEvolution's scalar part (hyperbolic_ops.py:408-424):
 * out = beta rs + (1-beta) dyn + clamp(delta, +-eps_r), rs = min(clamp(raw, rmin, rmax), 1/sqrt(c) - 1e-6); dyn = delta =
 * NULL gives out = rs.  Backward: draw, ddyn, ddelta.                                                               */
REGCN_API int regcn_radius_combine(const float* raw, const float* dyn, const float* delta, int M, float rmin, float rmax,
                         double c, float beta, float eps_r, float* out, void* stream);
REGCN_API int regcn_radius_combine_bwd(const float* raw, const float* delta, const float* g, int M, float rmin, float rmax,
                             double c, float beta, float eps_r, float* draw, float* ddyn, float* ddelta, void* stream);
/* radius_mlp = Linear(d,1): out[n] = <t[n], w> + b[0] (b: device scalar, may be NULL); backward: dt and the rows dout[n] t[n] (their column sum is dw)   */
REGCN_API int regcn_row_dot(const float* t, const float* w, const float* b, int M, int d, float* out, void* stream);
REGCN_API int regcn_row_dot_bwd(const float* t, const float* w, const float* dout, int M, int d, float* dt, float* scaled,
                      void* stream);
/* gradient of the radius-difference edge weights w.r.t. the radii: s_edge (per CSR position) and drho_dst (N);
 * regcn_edge_scalar_gather sums s_edge per source through a grouping of src_sorted (regcn_group_by_key)            */
REGCN_API int regcn_edge_radius_grad(const float* ht, const float* rel, const float* dagg, const int32_t* rowptr,
                           const int32_t* src_sorted, const int32_t* etype_sorted, const float* norm, const float* rho,
                           float gamma, int N, int d, float* s_edge, float* drho_dst, void* stream);
REGCN_API int regcn_edge_scalar_gather(const float* vals, const int32_t* rowptr, const int32_t* perm, int nrows, float* out,
                             int accumulate, void* stream);
/* loss_radius (hyperbolic_model.py:1066-1073): term[j] = lambda/n (rs[ids[j]] - target[ids[j]])^2 and d/d radius_static */
REGCN_API int regcn_radius_mse(const float* raw, const float* target, const int64_t* ids, int n, float rmin, float rmax,
                     double c, float lambda, float* term, void* stream);
REGCN_API int regcn_radius_mse_bwd(const float* raw, const float* target, const int64_t* ids, int n, float rmin, float rmax,
                         double c, float lambda, const float* gscale, float* draw, void* stream);

/* Distance decoders in training (HyperbolicMuRP / MuRPRel .loss, hyperbolic_decoder.py:647-928): mobius_add without its
 * projection, forward and backward (hyperbolic_ops.py:135-142); z = x*y; out[r] += alpha s[r] x[r]; and the gradient of
 * the proxy-distance score through its three scalars (dot, |q|^2, |e|^2): in place D <- dS dS/dD, H <- dS dS/d|e|^2,
 * per query gx = sum dS dS/d|q|^2, gs = sum dS (margin - n^2) (d scale), gm = sum dS scale (d margin).              */
REGCN_API int regcn_mobius_fwd(const float* x, const float* y, int M, int d, double c, float* z, void* stream);
REGCN_API int regcn_mobius_bwd(const float* x, const float* y, const float* dz, int M, int d, double c, float* dx, float* dy,
                     void* stream);
REGCN_API int regcn_eltwise_mul(const float* x, const float* y, float* z, size_t n, void* stream);
REGCN_API int regcn_row_axpy(const float* x, const float* s, float alpha, int M, int d, float* out, void* stream);
REGCN_API int regcn_hyp_dist_grad(float* D, const float* dS, float* H, int64_t ld, int B, int N, const float* x2,
                        const float* y2, double c, const float* scale_margin, float* gx, float* gs, float* gm,
                        void* stream);

/* Givens rotation (mode 0) / reflection (mode 1) of the tangent pairs (hyperbolic_decoder.py:1033-1051,1380-1401) and its
 * backward w.r.t. the vector and the angles; ang (B,d/2), or one (d/2) vector for all rows with ang_bcast (dang stays
 * per row: its column sum is the gradient of the shared angles).                                                    */
REGCN_API int regcn_givens_fwd(const float* x, const float* ang, int ang_bcast, int B, int d, int mode, float* y,
                     void* stream);
REGCN_API int regcn_givens_bwd(const float* x, const float* ang, const float* dy, int ang_bcast, int B, int d, int mode,
                     float* dx, float* dang, void* stream);

/* AttH / AttHRel attention mix in training (hyperbolic_decoder.py:1434-1445, 1617-1625): a[b] = sigmoid(<w[b], u[b]>) over
 * 2d features (w (B,2d), or one shared (2d) vector with w_bcast), mixed = a rot + (1-a) ref; and its backward (dw stays
 * per row: its column sum is the gradient of a shared w).                                                          */
REGCN_API int regcn_attn_mix_fwd(const float* w, int w_bcast, const float* u, const float* rot, const float* ref, int B, int d,
                       float* a, float* mixed, void* stream);
REGCN_API int regcn_attn_mix_bwd(const float* w, int w_bcast, const float* u, const float* rot, const float* ref,
                       const float* a, const float* g, int B, int d, float* dw, float* du, float* drot, float* dref,
                       void* stream);

/* LorentzRGCNLayer aggregate backward (lgcn encoder in training; hyperbolic_layers.py:589-625, hyperbolic_ops.py:492-518,
 * 563-581); num_bases relation blocks of sb x sb, sb = d / num_bases (2x2: chunk-local register kernels; any other size,
 * e.g. 10x10 where the reference clamps num_bases to 2R (:559-561): generic kernels with the (d, sb) gradient of a relation
 * accumulated in shared memory, REGCN_ERR_UNSUPPORTED above 200 KB).  gout (N,d) = dL/d(aggregate output).  dht (N,d);
 * part_rel (S, R2, d) and part_w (S, R2, d sb) with S = regcn_lorentz_bwd_splits(): per-split partial sums of drel / dW,
 * to be summed over S (regcn_col_sum).  type_* : edges grouped by relation type (regcn_group_by_key).
 * workspace: regcn_lorentz_aggregate_bwd_workspace_bytes(N, R2, d).                                                */
REGCN_API size_t regcn_lorentz_aggregate_bwd_workspace_bytes(int N, int R2, int d);
REGCN_API int regcn_lorentz_bwd_splits(void);
REGCN_API int regcn_lorentz_aggregate_bwd(const float* ht, const float* W, const float* rel, const float* gout,
                                const int32_t* rowptr, const int32_t* src_sorted, const int32_t* etype_sorted,
                                const float* norm, const int32_t* type_rowptr, const int32_t* type_src,
                                const int32_t* type_dst, int N, int R2, int d, int num_bases, double c, float* dht,
                                float* part_rel, float* part_w, float* workspace, size_t workspace_bytes, void* stream);

/* --plus-relation-specific-curvature in training (hyperbolic_decoder.py:66-86,145-163): gradient of the true-distance
 * score with per-query curvature row_c (as regcn_hyp_dist_grad, plus gc[b] = sum_n dS dS/dc_q), and the per-query
 * gradient w.r.t. rel_curvature_raw through softplus and the two clamps (draw_q (B), base_rel (B) = r mod R for the
 * per-relation sum).                                                                                                */
REGCN_API int regcn_hyp_truedist_grad(float* D, const float* dS, float* H, int64_t ld, int B, int N, const float* x2,
                            const float* y2, const float* row_c, const float* scale_margin, float* gx, float* gs,
                            float* gm, float* gc, void* stream);
REGCN_API int regcn_rel_curvature_bwd(const float* raw, const int64_t* triples, int B, int R, double c, double cmax,
                            const float* dcq, float* draw_q, int32_t* base_rel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REGCN_B200_H_ */
