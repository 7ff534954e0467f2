// K14: raw + time-filtered rank of the target among all candidates, without sorting.
// Restates rgcn/utils.py:21-25 (sort_and_rank), :51-75 (filter_score / filter_score_r), :136-166.
//   reference rank (0-based) = position of the target after a descending sort of the row
//   here                     = #{j : s_j > s_t} + #{j < t : s_j == s_t}   (the stable-sort position)
// which is identical whenever the target's score is unique in the row (ties only occur among the
// -1e7 filtered entries, which sit below any real target score).  Filtered variant: every other
// true answer j in filt(b) \ {t} is scored -10000000 (utils.py:60,74) before ranking.
// Output ranks are 1-based int64 like the reference (utils.py:162-163).
#include "common.cuh"

namespace regcn {

constexpr float kFilterScore = -10000000.0f;

__device__ __forceinline__ int rank_contrib(float s, int j, float st, int t) { return rank_beats(s, j, st, t); }

// One CTA per query row: counts over the dense row, then corrects for the (short, sorted) filter list.
__global__ void __launch_bounds__(256) rank_rows_kernel(
    const float* __restrict__ S, size_t ld, int B, int N, const int64_t* __restrict__ triples, int target_col,
    const int* __restrict__ filt_ptr, const int* __restrict__ filt_idx, int col_offset,
    int* __restrict__ raw_count, int* __restrict__ filt_count, float* __restrict__ target_score,
    const int* __restrict__ filt_end) {
  pdl_grid_sync();
  __shared__ int red[8];
  const int b = blockIdx.x;
  const float* row = S + (size_t)b * ld;
  const int t = (int)triples[3 * (size_t)b + target_col] - col_offset;  // local column of the target (may be outside this shard)
  const float st = target_score[b];
  int cnt = 0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    if (j != t) cnt += rank_contrib(row[j], j, st, t);
  }
  // filter correction: entries of filt(b) inside this shard, excluding the target itself
  int corr = 0;
  if (filt_ptr) {
    const int fb = filt_ptr[b], fe = filt_end ? filt_end[b] : filt_ptr[b + 1];
    for (int i = fb + threadIdx.x; i < fe; i += blockDim.x) {
      const int j = filt_idx[i] - col_offset;
      if (j < 0 || j >= N || j == t) continue;
      corr += rank_contrib(kFilterScore, j, st, t) - rank_contrib(row[j], j, st, t);
    }
  }
  cnt = warp_sum_i(cnt);
  corr = warp_sum_i(corr);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = cnt;
  __syncthreads();
  int tot = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i]; }
  __syncthreads();
  if (lane == 0) red[wid] = corr;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tc = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tc += red[i];
    raw_count[b] = tot;
    if (filt_count) filt_count[b] = tot + tc;
  }
}

// target_score[b] = S[b, t_b] read from the dense matrix (exactly the value the counts compare against)
__global__ void gather_target_score_kernel(const float* __restrict__ S, size_t ld, int B, int N,
                                           const int64_t* __restrict__ triples, int target_col, int col_offset,
                                           float* __restrict__ target_score) {
  pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int t = (int)triples[3 * (size_t)b + target_col] - col_offset;
  if (t >= 0 && t < N) target_score[b] = S[(size_t)b * ld + t];
}

__global__ void counts_to_ranks_kernel(const int* __restrict__ raw_count, const int* __restrict__ filt_count, int B,
                                       int64_t* __restrict__ rank, int64_t* __restrict__ filt_rank) {
  pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  rank[b] = (int64_t)raw_count[b] + 1;
  if (filt_rank) filt_rank[b] = (int64_t)(filt_count ? filt_count[b] : raw_count[b]) + 1;
}

// Optionally reproduce the reference's in-place side effect: score[b][ans \ {t}] = -1e7 (utils.py:60).
__global__ void apply_filter_kernel(float* __restrict__ S, size_t ld, int B, int N, const int64_t* __restrict__ triples,
                                    int target_col, const int* __restrict__ filt_ptr, const int* __restrict__ filt_idx,
                                    int col_offset, const int* __restrict__ filt_end) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  const int t = (int)triples[3 * (size_t)b + target_col] - col_offset;
  const int fe = filt_end ? filt_end[b] : filt_ptr[b + 1];
  for (int i = filt_ptr[b] + threadIdx.x; i < fe; i += blockDim.x) {
    const int j = filt_idx[i] - col_offset;
    if (j >= 0 && j < N && j != t) S[(size_t)b * ld + j] = kFilterScore;
  }
}

// Filtered counts for the fused (never materialised) scoring path.  pair_score[B + i] is the score of the i-th
// filter-CSR entry (query b, candidate f = filt_idx[i]) computed by the pair-score pass with the scoring GEMM's own
// arithmetic, so subtracting its contribution from the raw count is exact.  Only candidates of the shard
// [col_lo, col_hi) are corrected (entity-sharded scoring sums the shards afterwards).
__global__ void filter_correct_kernel(int B, const int* __restrict__ filt_ptr, const int* __restrict__ filt_idx,
                                      const int* __restrict__ target, const float* __restrict__ pair_score,
                                      const int* __restrict__ raw_count, int col_lo, int col_hi,
                                      int* __restrict__ filt_count, const int* __restrict__ filt_end) {
  pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float st = pair_score[b];
  const int t = target[b];
  int corr = 0;
  const int fe = filt_end ? filt_end[b] : filt_ptr[b + 1];
  for (int i = filt_ptr[b]; i < fe; ++i) {
    const int f = filt_idx[i];
    if (f == t || f < col_lo || f >= col_hi) continue;
    corr += rank_contrib(kFilterScore, f, st, t) - rank_contrib(pair_score[B + i], f, st, t);
  }
  filt_count[b] = raw_count[b] + corr;
}

int filter_correct(int B, const int* filt_ptr, const int* filt_idx, const int* target, const float* pair_score,
                   const int* raw_count, int col_lo, int col_hi, int* filt_count, const int* filt_end, cudaStream_t st) {
  if (!filt_ptr || !filt_idx || !target || !pair_score || !raw_count || !filt_count) { set_last_error("filter_correct: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0) return REGCN_OK;
  launch_k(filter_correct_kernel, (B + 127) / 128, 128, 0, st, B, filt_ptr, filt_idx, target, pair_score, raw_count, col_lo, col_hi, filt_count, filt_end);
  return check_launch("filter_correct");
}

// ---- time-aware filter lists straight from the query triples (rgcn/utils.py:264-304 on the test snapshot itself) ----
// Query b = (h, r, t) [all_triples incl. inverses].  Entity prediction: answers of key (h, r) = every t' among the
// queries with the same (h, r); relation prediction: key (h, t), answers r'.  B is a few thousand, so an all-pairs
// scan (B^2 key comparisons out of L1) beats a sort.  One WARP per query: the lanes stride over the B keys, matches
// are compacted with a ballot; pass 1 counts them, pass 2 (after an exclusive scan) collects them, sorts + uniques
// each short list in registers (rank sort over shuffles) and emits the (query, candidate) pair lists of the fused
// rank path.
constexpr int kFiltThreads = 256;

__device__ __forceinline__ long long filt_key(const int64_t* __restrict__ triples, int j, int key_col) {
  return (long long)((unsigned long long)triples[3 * (size_t)j] << 32) | (triples[3 * (size_t)j + key_col] & 0xffffffffLL);
}

constexpr int kFiltBlocks = 8;        // lists of up to 32 * kFiltBlocks answers are sorted across the warp
constexpr int kFiltTile = 2048;       // keys staged in shared memory per pass (16 KB)

// Every CTA scans all B keys: they are packed once per tile into shared memory (the triples are 24-byte records, a key
// takes two of their fields) and the warps of the CTA -- one query each -- compare against the staged copy.
__device__ __forceinline__ int filt_stage_keys(const int64_t* __restrict__ triples, int B, int key_col, int tile0,
                                               long long* __restrict__ skeys) {
  const int tn = min(kFiltTile, B - tile0);
  __syncthreads();                                        // the previous tile is no longer read
  for (int j = threadIdx.x; j < tn; j += blockDim.x) skeys[j] = filt_key(triples, tile0 + j, key_col);
  __syncthreads();
  return tn;
}

__global__ void __launch_bounds__(kFiltThreads) filter_count_kernel(const int64_t* __restrict__ triples, int B, int key_col,
                                                                    int* __restrict__ counts) {
  pdl_grid_sync();
  __shared__ long long skeys[kFiltTile];
  const int lane = threadIdx.x & 31;
  const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const bool live = b < B;
  const long long mykey = live ? filt_key(triples, b, key_col) : 0;
  int c = 0;
  for (int tile0 = 0; tile0 < B; tile0 += kFiltTile) {
    const int tn = filt_stage_keys(triples, B, key_col, tile0, skeys);
    if (live)
      for (int j = lane; j < tn; j += 32) c += skeys[j] == mykey ? 1 : 0;
  }
  c = warp_sum_i(c);
  if (live && lane == 0) counts[b] = c;
}

// One filter list set: key (col 0, key_col) -> answers ans_col, offsets beg, outputs idx / end / pairs.
struct FiltFill {
  int key_col, ans_col;
  const int* beg;
  int* idx;
  int* end;
  int* pair_a;
  int* pair_e;
};
constexpr int kFillTile = 3072;       // (key, answer) records staged per pass: 36 KB; one pass for a TKG timestamp

// grid.y selects the list set (the entity and the relation filter of a timestamp are filled by ONE launch: each was a
// ~17 us latency chain of its own).  Keys AND answers are staged, so a match costs no global load; the common short list
// (<= 32 matches) never touches global memory before its final stores: matches are compacted through a 32-slot
// shared-memory line per warp, rank-sorted + uniqued over shuffles, and idx / pair lists are written from registers.
__global__ void __launch_bounds__(kFiltThreads) filter_fill_kernel(const int64_t* __restrict__ triples, int B, FiltFill f0,
                                                                   FiltFill f1) {
  pdl_grid_sync();
  __shared__ long long skeys[kFillTile];
  __shared__ int sans[kFillTile];
  __shared__ int wl[kFiltThreads / 32][32];
  const FiltFill f = blockIdx.y ? f1 : f0;
  const int key_col = f.key_col, ans_col = f.ans_col;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const bool live = b < B;                        // (every warp of the CTA takes part in staging the keys)
  const long long mykey = live ? filt_key(triples, b, key_col) : 0;
  const int b0 = live ? __ldg(f.beg + b) : 0;
  int* lst = f.idx + b0;
  int* pair_a = f.pair_a;
  int* pair_e = f.pair_e;
  // collect the answers of the matching queries in query order
  int n = 0;
  for (int tile0 = 0; tile0 < B; tile0 += kFillTile) {
    const int tn = min(kFillTile, B - tile0);
    __syncthreads();                                        // the previous tile is no longer read
    for (int j = threadIdx.x; j < tn; j += blockDim.x) {
      skeys[j] = filt_key(triples, tile0 + j, key_col);
      sans[j] = (int)triples[3 * (size_t)(tile0 + j) + ans_col];
    }
    __syncthreads();
    if (!live) continue;
#pragma unroll 4
    for (int j0 = 0; j0 < tn; j0 += 32) {
      const int j = j0 + lane;
      const bool m = j < tn && skeys[j] == mykey;
      const unsigned bal = __ballot_sync(0xffffffffu, m);
      if (bal) {
        const int pos = n + __popc(bal & ((1u << lane) - 1u));
        if (m) {
          if (pos < 32) wl[wid][pos] = sans[j];
          else lst[pos] = sans[j];
        }
        n += __popc(bal);
      }
    }
  }
  if (!live) return;
  __syncwarp();
  int u;
  if (n <= 32) {
    const int mine = lane < n ? wl[wid][lane] : 0x7fffffff;   // lane i keeps the i-th match
    // duplicate = an equal value at a smaller lane; unique position = #{distinct values below mine}
    bool dup = false;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int o = __shfl_sync(0xffffffffu, mine, k);
      dup |= (k < lane) & (o == mine);
    }
    const bool keep = lane < n && !dup;
    int upos = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int o = __shfl_sync(0xffffffffu, mine, k);
      const int od = __shfl_sync(0xffffffffu, (int)keep, k);
      upos += od & (int)(o < mine);
    }
    u = __popc(__ballot_sync(0xffffffffu, keep));
    int first = mine;                                          // smallest answer: fills the unused tail slots
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    if (keep) lst[upos] = mine;
    if (lane >= u && lane < n) lst[lane] = first;              // unused tail slots stay valid candidate ids
    if (lane == 0) f.end[b] = b0 + u;
    if (pair_a) {
      if (lane == 0) { pair_a[b] = b; pair_e[b] = (int)triples[3 * (size_t)b + ans_col]; }
      if (lane < n) pair_a[B + b0 + lane] = b;
      if (keep) pair_e[B + b0 + upos] = mine;
      if (lane >= u && lane < n) pair_e[B + b0 + lane] = first;
    }
    return;
  }
  if (lane < 32) lst[lane] = wl[wid][lane];                    // longer lists are finished in place
  __syncwarp();
  if (n <= 32 * kFiltBlocks) {
    // medium list (a hub pair / hub entity of a skewed snapshot): the same rank sort with several elements per lane --
    // element blk*32 + lane lives in val[blk]; one serial insertion sort of 50 elements would cost the whole launch 40 us
    int val[kFiltBlocks], pos[kFiltBlocks];
    unsigned keepmask = 0;
    const int nb = (n + 31) >> 5;
#pragma unroll
    for (int blk = 0; blk < kFiltBlocks; ++blk) val[blk] = (blk * 32 + lane < n) ? lst[blk * 32 + lane] : 0x7fffffff;
#pragma unroll
    for (int blk = 0; blk < kFiltBlocks; ++blk) {
      if (blk < nb) {
        const int i = blk * 32 + lane;
        bool dup = false;
#pragma unroll
        for (int jb = 0; jb < kFiltBlocks; ++jb) {
          if (jb <= blk) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int o = __shfl_sync(0xffffffffu, val[jb], k);
              dup |= (jb * 32 + k < i) & (o == val[blk]);
            }
          }
        }
        if (i < n && !dup) keepmask |= 1u << blk;
      }
    }
    u = 0;
#pragma unroll
    for (int blk = 0; blk < kFiltBlocks; ++blk) {
      pos[blk] = 0;
      if (blk < nb) {
        int upos = 0;
#pragma unroll
        for (int jb = 0; jb < kFiltBlocks; ++jb) {
          if (jb < nb) {
            const unsigned kb = __ballot_sync(0xffffffffu, (keepmask >> jb) & 1u);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int o = __shfl_sync(0xffffffffu, val[jb], k);
              upos += (int)((kb >> k) & 1u) & (int)(o < val[blk]);
            }
          }
        }
        pos[blk] = upos;
        u += __popc(__ballot_sync(0xffffffffu, (keepmask >> blk) & 1u));
      }
    }
    __syncwarp();
#pragma unroll
    for (int blk = 0; blk < kFiltBlocks; ++blk)
      if ((keepmask >> blk) & 1u) lst[pos[blk]] = val[blk];
    __syncwarp();
    const int first = lst[0];
    for (int i = u + lane; i < n; i += 32) lst[i] = first;     // unused tail slots stay valid candidate ids
    __syncwarp();
  } else {
    // long list (hub query): insertion sort + unique by one lane, in place
    u = 0;
    if (lane == 0) {
      for (int i = 1; i < n; ++i) {
        const int a = lst[i];
        int p = i;
        while (p > 0 && lst[p - 1] > a) { lst[p] = lst[p - 1]; --p; }
        lst[p] = a;
      }
      for (int i = 0; i < n; ++i) if (i == 0 || lst[i] != lst[i - 1]) lst[u++] = lst[i];
      for (int i = u; i < n; ++i) lst[i] = lst[0];
    }
    u = __shfl_sync(0xffffffffu, u, 0);
    __syncwarp();
  }
  if (lane == 0) f.end[b] = b0 + u;
  if (pair_a) {
    if (lane == 0) { pair_a[b] = b; pair_e[b] = (int)triples[3 * (size_t)b + ans_col]; }
    for (int i = lane; i < n; i += 32) { pair_a[B + b0 + i] = b; pair_e[B + b0 + i] = lst[i]; }
  }
}

// ---- one-call preparation of a test snapshot's queries (src/main.py:60-74 before predict / get_total_rank) ----------
// all_t = [triples ; (o, r + R, s)] (src/rrgcn.py:184-186), the match counts of both filter keys and their exclusive
// scans: what regcn_b200.test() needs on the device before it can size the filter lists of a timestamp.  Three launches
// instead of ~14 framework operations (flip / add / cat / 2 x {count, cumsum, subtract, slice}).
__global__ void __launch_bounds__(256) queries_inverse_kernel(const int64_t* __restrict__ triples, int T, int R,
                                                              int64_t* __restrict__ all_t) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const int64_t s = triples[3 * (size_t)i], r = triples[3 * (size_t)i + 1], o = triples[3 * (size_t)i + 2];
  int64_t* f = all_t + 3 * (size_t)i;
  int64_t* b = all_t + 3 * ((size_t)T + i);
  f[0] = s; f[1] = r; f[2] = o;
  b[0] = o; b[1] = r + R; b[2] = s;
}

__global__ void __launch_bounds__(kFiltThreads) filter_count2_kernel(const int64_t* __restrict__ triples, int B,
                                                                     int* __restrict__ counts) {
  pdl_grid_sync();
  __shared__ long long skeys[kFiltTile];
  const int key_col = 1 + (int)blockIdx.y;                 // y = 0: entity filter (h, r); y = 1: relation filter (h, t)
  const int lane = threadIdx.x & 31;
  const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const bool live = b < B;
  const long long mykey = live ? filt_key(triples, b, key_col) : 0;
  int c = 0;
  for (int tile0 = 0; tile0 < B; tile0 += kFiltTile) {
    const int tn = filt_stage_keys(triples, B, key_col, tile0, skeys);
    if (live)
      for (int j = lane; j < tn; j += 32) c += skeys[j] == mykey ? 1 : 0;
  }
  c = warp_sum_i(c);
  if (live && lane == 0) counts[(size_t)blockIdx.y * B + b] = c;
}

// exclusive scan of counts[y][0..B) into beg[y][0..B), totals[y] = sum; one 1024-thread CTA per y
__global__ void __launch_bounds__(1024) filter_scan2_kernel(const int* __restrict__ counts, int B, int* __restrict__ beg,
                                                            int* __restrict__ totals) {
  pdl_grid_sync();
  __shared__ int wsum[32];
  const int* c = counts + (size_t)blockIdx.y * B;
  int* o = beg + (size_t)blockIdx.y * B;
  const int per = (B + 1023) / 1024;
  const int lo = min(B, (int)threadIdx.x * per), hi = min(B, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += c[i];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = wsum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
    wsum[lane] = v;
  }
  __syncthreads();
  int run = inc - s + (w > 0 ? wsum[w - 1] : 0);          // exclusive prefix of this thread's slice
  for (int i = lo; i < hi; ++i) { o[i] = run; run += c[i]; }
  if (threadIdx.x == 1023) totals[blockIdx.y] = wsum[31];
}

// ---- the same for the n test snapshots of a group in three launches (grid.z = snapshot) -------------------------------
// triples_cat (sum T, 3): the snapshots back to back, toff[g] = rows in front of snapshot g.  Snapshot g's outputs sit at
// all_t_cat + 3 * 2 toff[g] (2 T_g rows), counts_cat / beg_cat + 4 toff[g] ((2, 2 T_g) each), totals + 2 g.
constexpr int kPrepBatchMax = 32;
struct QPrepBatch {
  int n;
  int toff[kPrepBatchMax + 1];
};

__global__ void __launch_bounds__(256) queries_inverse_batch_kernel(const int64_t* __restrict__ triples_cat, int R,
                                                                    int64_t* __restrict__ all_t_cat, QPrepBatch qb) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= qb.toff[qb.n]) return;
  int g = 0;
  while (i >= qb.toff[g + 1]) ++g;
  const int T = qb.toff[g + 1] - qb.toff[g], li = i - qb.toff[g];
  const int64_t s = triples_cat[3 * (size_t)i], r = triples_cat[3 * (size_t)i + 1], o = triples_cat[3 * (size_t)i + 2];
  int64_t* f = all_t_cat + 3 * ((size_t)2 * qb.toff[g] + li);
  int64_t* b = f + 3 * (size_t)T;
  f[0] = s; f[1] = r; f[2] = o;
  b[0] = o; b[1] = r + R; b[2] = s;
}

// 1024 threads: every CTA stages all keys of its snapshot, so 32 queries per CTA read a quarter of what 8 would
__global__ void __launch_bounds__(1024) filter_count2_batch_kernel(const int64_t* __restrict__ all_t_cat,
                                                                           int* __restrict__ counts_cat, QPrepBatch qb) {
  pdl_grid_sync();
  __shared__ long long skeys[kFiltTile];
  const int g = blockIdx.z;
  const int B = 2 * (qb.toff[g + 1] - qb.toff[g]);
  if ((size_t)blockIdx.x * (blockDim.x / 32) >= (size_t)B) return;        // (whole CTA: the grid is sized for the largest snapshot)
  const int64_t* triples = all_t_cat + 3 * (size_t)2 * qb.toff[g];
  int* counts = counts_cat + (size_t)4 * qb.toff[g];
  const int key_col = 1 + (int)blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const bool live = b < B;
  const long long mykey = live ? filt_key(triples, b, key_col) : 0;
  int c = 0;
  for (int tile0 = 0; tile0 < B; tile0 += kFiltTile) {
    const int tn = filt_stage_keys(triples, B, key_col, tile0, skeys);
    if (live)
      for (int j = lane; j < tn; j += 32) c += skeys[j] == mykey ? 1 : 0;
  }
  c = warp_sum_i(c);
  if (live && lane == 0) counts[(size_t)blockIdx.y * B + b] = c;
}

__global__ void __launch_bounds__(1024) filter_scan2_batch_kernel(const int* __restrict__ counts_cat, int* __restrict__ beg_cat,
                                                                  int* __restrict__ totals, QPrepBatch qb) {
  pdl_grid_sync();
  __shared__ int wsum[32];
  const int g = blockIdx.z;
  const int B = 2 * (qb.toff[g + 1] - qb.toff[g]);
  const int* c = counts_cat + (size_t)4 * qb.toff[g] + (size_t)blockIdx.y * B;
  int* o = beg_cat + (size_t)4 * qb.toff[g] + (size_t)blockIdx.y * B;
  const int per = (B + 1023) / 1024;
  const int lo = min(B, (int)threadIdx.x * per), hi = min(B, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += c[i];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = wsum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
    wsum[lane] = v;
  }
  __syncthreads();
  int run = inc - s + (w > 0 ? wsum[w - 1] : 0);
  for (int i = lo; i < hi; ++i) { o[i] = run; run += c[i]; }
  if (threadIdx.x == 1023) totals[2 * g + blockIdx.y] = wsum[31];
}

int queries_prepare_batch(const int64_t* triples_cat, const int* toff, int n, int R, int64_t* all_t_cat, int* counts_cat,
                          int* beg_cat, int* totals, cudaStream_t st) {
  if (!triples_cat || !toff || !all_t_cat || !counts_cat || !beg_cat || !totals) { set_last_error("queries_prepare_batch: null pointer"); return REGCN_ERR_NULL; }
  if (n <= 0 || n > kPrepBatchMax || R <= 0) { set_last_error("queries_prepare_batch: 1..%d snapshots per call (n=%d), R=%d", kPrepBatchMax, n, R); return REGCN_ERR_DIM; }
  QPrepBatch qb;
  qb.n = n;
  int maxT = 0;
  for (int g = 0; g <= kPrepBatchMax; ++g) qb.toff[g] = toff[g <= n ? g : n];
  if (qb.toff[0] != 0) { set_last_error("queries_prepare_batch: toff[0] must be 0"); return REGCN_ERR_DIM; }
  for (int g = 0; g < n; ++g) {
    const int T = qb.toff[g + 1] - qb.toff[g];
    if (T <= 0 || T > (1 << 24)) { set_last_error("queries_prepare_batch: snapshot %d has %d triples", g, T); return REGCN_ERR_DIM; }
    maxT = T > maxT ? T : maxT;
  }
  const int total = qb.toff[n];
  launch_k(queries_inverse_batch_kernel, (unsigned)((total + 255) / 256), 256, 0, st, triples_cat, R, all_t_cat, qb);
  launch_k(filter_count2_batch_kernel, dim3((unsigned)(((size_t)2 * maxT * 32 + 1023) / 1024), 2, (unsigned)n),
           dim3(1024), 0, st, (const int64_t*)all_t_cat, counts_cat, qb);
  launch_k(filter_scan2_batch_kernel, dim3(1, 2, (unsigned)n), dim3(1024), 0, st, (const int*)counts_cat, beg_cat, totals, qb);
  return check_launch("queries_prepare_batch");
}

int queries_prepare(const int64_t* triples, int T, int R, int64_t* all_t, int* counts, int* beg, int* totals, cudaStream_t st) {
  if (!triples || !all_t || !counts || !beg || !totals) { set_last_error("queries_prepare: null pointer"); return REGCN_ERR_NULL; }
  if (T <= 0 || R <= 0 || T > (1 << 24)) { set_last_error("queries_prepare: bad dims T=%d R=%d", T, R); return REGCN_ERR_DIM; }
  const int B = 2 * T;
  launch_k(queries_inverse_kernel, (unsigned)((T + 255) / 256), 256, 0, st, triples, T, R, all_t);
  launch_k(filter_count2_kernel, dim3((unsigned)(((size_t)B * 32 + kFiltThreads - 1) / kFiltThreads), 2), dim3(kFiltThreads), 0, st,
           (const int64_t*)all_t, B, counts);
  launch_k(filter_scan2_kernel, dim3(1, 2), dim3(1024), 0, st, (const int*)counts, B, beg, totals);
  return check_launch("queries_prepare");
}

int filter_count(const int64_t* triples, int B, int key_col, int* counts, cudaStream_t st) {
  if (!triples || !counts) { set_last_error("filter_count: null pointer"); return REGCN_ERR_NULL; }
  if (key_col < 1 || key_col > 2) { set_last_error("filter_count: key_col must be 1 or 2"); return REGCN_ERR_DIM; }
  if (B <= 0) return REGCN_OK;
  launch_k(filter_count_kernel, (unsigned)(((size_t)B * 32 + kFiltThreads - 1) / kFiltThreads), kFiltThreads, 0, st, triples, B, key_col, counts);
  return check_launch("filter_count");
}

static bool filt_fill_ok(const FiltFill& f) {
  return f.beg && f.idx && f.end && !(f.pair_a && !f.pair_e) && f.key_col >= 1 && f.key_col <= 2 && f.ans_col >= 1 &&
         f.ans_col <= 2 && f.key_col != f.ans_col;
}

int filter_fill(const int64_t* triples, int B, int key_col, int ans_col, const int* beg, int* idx, int* end, int* pair_a,
                int* pair_e, cudaStream_t st) {
  const FiltFill f{key_col, ans_col, beg, idx, end, pair_a, pair_e};
  if (!triples || !filt_fill_ok(f)) { set_last_error("filter_fill: null pointer or bad columns (key_col / ans_col are 1 and 2)"); return triples && beg && idx && end ? REGCN_ERR_DIM : REGCN_ERR_NULL; }
  if (B <= 0) return REGCN_OK;
  launch_k(filter_fill_kernel, dim3((unsigned)(((size_t)B * 32 + kFiltThreads - 1) / kFiltThreads), 1), dim3(kFiltThreads), 0, st, triples, B, f, f);
  return check_launch("filter_fill");
}

// entity filter (key (h, r) -> answers t) and relation filter (key (h, t) -> answers r) of the same queries in one launch
int filter_fill2(const int64_t* triples, int B, const int* beg_e, int* idx_e, int* end_e, int* pair_a_e, int* pair_e_e,
                 const int* beg_r, int* idx_r, int* end_r, int* pair_a_r, int* pair_e_r, cudaStream_t st) {
  const FiltFill fe{1, 2, beg_e, idx_e, end_e, pair_a_e, pair_e_e}, fr{2, 1, beg_r, idx_r, end_r, pair_a_r, pair_e_r};
  if (!triples || !filt_fill_ok(fe) || !filt_fill_ok(fr)) { set_last_error("filter_fill2: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0) return REGCN_OK;
  launch_k(filter_fill_kernel, dim3((unsigned)(((size_t)B * 32 + kFiltThreads - 1) / kFiltThreads), 2), dim3(kFiltThreads), 0, st, triples, B, fe, fr);
  return check_launch("filter_fill2");
}

int gather_target_score(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, int col_offset,
                        float* target_score, cudaStream_t st) {
  if (!S || !triples || !target_score) { set_last_error("gather_target_score: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0) return REGCN_OK;
  launch_k(gather_target_score_kernel, (B + 255) / 256, 256, 0, st, S, (size_t)ld, B, N, triples, target_col, col_offset, target_score);
  return check_launch("gather_target_score");
}

int rank_count(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, const int* filt_ptr,
               const int* filt_idx, int col_offset, const float* target_score, int* raw_count, int* filt_count,
               const int* filt_end, cudaStream_t st) {
  if (!S || !triples || !target_score || !raw_count || (filt_ptr && !filt_idx)) { set_last_error("rank_count: null pointer"); return REGCN_ERR_NULL; }
  if (target_col < 0 || target_col > 2 || ld < N) { set_last_error("rank_count: bad target_col=%d or ld", target_col); return REGCN_ERR_DIM; }
  if (B <= 0) return REGCN_OK;
  launch_k(rank_rows_kernel, B, 256, 0, st, S, (size_t)ld, B, N, triples, target_col, filt_ptr, filt_idx, col_offset, raw_count,
                                      filt_count, const_cast<float*>(target_score), filt_end);
  return check_launch("rank_count");
}

int counts_to_ranks(const int* raw_count, const int* filt_count, int B, int64_t* rank, int64_t* filt_rank, cudaStream_t st) {
  if (!raw_count || !rank) { set_last_error("counts_to_ranks: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0) return REGCN_OK;
  launch_k(counts_to_ranks_kernel, (B + 255) / 256, 256, 0, st, raw_count, filt_count, B, rank, filt_rank);
  return check_launch("counts_to_ranks");
}

int apply_filter(float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, const int* filt_ptr,
                 const int* filt_idx, int col_offset, const int* filt_end, cudaStream_t st) {
  if (!S || !triples || !filt_ptr || !filt_idx) { set_last_error("apply_filter: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0) return REGCN_OK;
  launch_k(apply_filter_kernel, B, 128, 0, st, S, (size_t)ld, B, N, triples, target_col, filt_ptr, filt_idx, col_offset, filt_end);
  return check_launch("apply_filter");
}


// Cross entropy of a dense (B, N) score matrix against column triples[:, target_col] (the relation-prediction loss
// head, src/rrgcn.py:220-222: N = 2R is small enough to materialise).  One warp per row, fixed lane order.
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ S, size_t ld, int B, int N,
                                                      const int64_t* __restrict__ triples, int target_col,
                                                      float* __restrict__ ce) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (b >= B) return;
  const float* row = S + (size_t)b * ld;
  float m = -INFINITY;
  for (int j = lane; j < N; j += 32) m = fmaxf(m, row[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < N; j += 32) s += expf(row[j] - m);
  s = warp_sum(s);
  if (lane == 0) ce[b] = (m + logf(s)) - row[(int)triples[3 * (size_t)b + target_col]];
}

int ce_rows(const float* S, int64_t ld, int B, int N, const int64_t* triples, int target_col, float* ce, cudaStream_t st) {
  if (!S || !triples || !ce) { set_last_error("ce_rows: null pointer"); return REGCN_ERR_NULL; }
  if (target_col < 0 || target_col > 2 || ld < N || N <= 0) { set_last_error("ce_rows: bad dims"); return REGCN_ERR_DIM; }
  if (B <= 0) return REGCN_OK;
  launch_k(ce_rows_kernel, (unsigned)(((size_t)B * 32 + 255) / 256), 256, 0, st, S, (size_t)ld, B, N, triples, target_col, ce);
  return check_launch("ce_rows");
}


// ---------------------------------------------------------------------------------------------------------------
// Multi-step inference (src/main.py:90-97, rgcn/utils.py:367-405): the K best candidates of every score row, in
// descending order with ties by ascending id (the order of a stable descending sort), turned into the predicted
// snapshot's triples.  One CTA per row, K selection passes over the row (it stays in L1/L2): pass p takes the
// greatest (score, -id) strictly below the pass p-1 winner; the score matrix is not modified.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ S, int64_t ld, int B, int N, int K,
                                                        int* __restrict__ top_idx) {
  pdl_grid_sync();
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ float last_v_s;
  __shared__ int last_i_s;
  const int b = blockIdx.x;
  const float* s = S + (size_t)b * ld;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float last_v = INFINITY;
  int last_i = -1;
  for (int p = 0; p < K; ++p) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const float v = s[j];
      const bool below = p == 0 || v < last_v || (v == last_v && j > last_i);
      if (below && (v > bv || (v == bv && j < bi))) { bv = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[wid] = bv; si[wid] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = sv[0];
      int i = si[0];
      for (int w = 1; w < 8; ++w)
        if (sv[w] > v || (sv[w] == v && si[w] < i)) { v = sv[w]; i = si[w]; }
      last_v_s = v; last_i_s = i;
      top_idx[(size_t)b * K + p] = i == 0x7fffffff ? -1 : i;
    }
    __syncthreads();
    last_v = last_v_s; last_i = last_i_s;
  }
}
// rel_mode 0 (construct_snap): query (s, r, .) -> r < R ? (s, r, idx) : (idx, r-R, s)
// rel_mode 1 (construct_snap_r): query (h, ., t) -> idx < R ? (h, idx, t) : (t, idx-R, h)
__global__ void construct_snap_kernel(const int64_t* __restrict__ triples, const int* __restrict__ top_idx, int B, int K,
                                      int R, int rel_mode, int64_t* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K;
  const int64_t idx = top_idx[i];
  const int64_t h = triples[(size_t)b * 3], r = triples[(size_t)b * 3 + 1], t = triples[(size_t)b * 3 + 2];
  int64_t o0, o1, o2;
  if (!rel_mode) {
    if (r < R) { o0 = h; o1 = r; o2 = idx; } else { o0 = idx; o1 = r - R; o2 = h; }
  } else {
    if (idx < R) { o0 = h; o1 = idx; o2 = t; } else { o0 = t; o1 = idx - R; o2 = h; }
  }
  out[(size_t)i * 3] = o0; out[(size_t)i * 3 + 1] = o1; out[(size_t)i * 3 + 2] = o2;
}
int topk_construct_snap(const float* S, int64_t ld, int B, int N, int K, const int64_t* triples, int R, int rel_mode,
                        int* top_idx, int64_t* out, cudaStream_t st) {
  if (!S || !top_idx || (out && !triples)) { set_last_error("topk_construct_snap: null pointer"); return REGCN_ERR_NULL; }
  if (K <= 0 || K > N) { set_last_error("topk_construct_snap: need 0 < K <= N (K=%d, N=%d)", K, N); return REGCN_ERR_DIM; }
  if (B <= 0) return REGCN_OK;
  launch_k(topk_rows_kernel, (unsigned)B, 256, 0, st, S, ld, B, N, K, top_idx);
  if (out) launch_k(construct_snap_kernel, (unsigned)(((size_t)B * K + 255) / 256), 256, 0, st, triples, (const int*)top_idx, B, K, R, rel_mode, out);
  return check_launch("topk_construct_snap");
}
}  // namespace regcn
