// Dense contraction on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32, operands staged in
// shared memory by TMA (128-byte swizzle), fp32 accumulator in tensor memory, epilogue via tcgen05.ld.
//
//   C[M,N] (+)= A[M,K] . B[N,K]^T (+ bias[N])          A and B both K-major (row-major with K contiguous)
//
// fp32 parity ("3xTF32"): every operand is split on the fly-side into hi = rna_tf32(x) and lo = rna_tf32(x - hi)
// (split_tf32_kernel below); the kernel accumulates lo.hi + hi.lo + hi.hi into the same TMEM accumulator,
// which restores ~2^-21 relative operand precision (the dropped lo.lo term is 2^-22).  With passes == 1 only
// hi.hi is issued (plain TF32, ~1e-3 relative) -- reported separately, never used for the parity path.
//
// Warp roles (384 threads):  warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..3 = A-operand converters (fp32-A mode, below), warps 4..11 = epilogue (TMEM lane quarter = warp_id % 4, two
// warps per quarter splitting the 32-column chunks: one warp per scheduler cannot hide the ALU / memory latency of a
// non-trivial epilogue).  Persistent: each CTA walks a list of 128 x BLOCK_N output tiles (x split-K slices).
// K is walked in 32-float (128-byte) blocks; TMA zero-fills the K / M / N tails.
//
// fp32-A mode (Params::a_f32): the A operand stays ONE fp32 copy in global memory.  The two converter warps read the
// 128 x 32 k-block of the tile with coalesced 16-byte loads (optionally gathering rows through an index list, optionally
// from two K segments -- the compact [agg | x] operand of the active rows never exists in memory), split every value into
// hi / lo in registers and write both halves straight into the 128-byte-swizzled K-major stage buffers the tensor core
// reads (same bytes TMA would have written from pre-split copies), then fence.proxy.async + mbarrier arrive.  Producers
// of activations therefore write fp32 only, and a layer GEMM moves half the A bytes through L2.
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <vector>

namespace regcn {

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;            // fp32 elements = one 128-byte swizzle row
constexpr int UMMA_K = 8;              // tf32
constexpr int EPI_WARPS = 8;           // epilogue warps: 2 per TMEM lane quarter, each takes every other 32-column chunk
constexpr int CVT_WARPS = 2;           // A-operand converter warps (fp32-A mode)
constexpr int EPI_WARP0 = 2 + CVT_WARPS;   // first epilogue warp; a multiple of 4 so that warp % 4 is the TMEM lane quarter
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
// The time-gate epilogue (EPI 5) is a latency-bound instruction stream (profiles/r02_ncu_le1_batched.txt): it runs with THREE
// warps per TMEM lane quarter = per warp scheduler (512 threads, 128 registers each) instead of two.
template <int EPI> constexpr int epi_warps() { return EPI == 5 ? 12 : EPI_WARPS; }
template <int EPI> constexpr int num_threads() { return 32 * (EPI_WARP0 + epi_warps<EPI>()); }
constexpr uint32_t staging_bytes(int warps) { return (uint32_t)warps * 32 * 32 * 4; }       // one 32 x 32 fp32 tile per epilogue warp
// operand-ring budget: what is left of the 227 KB (232 448 B) per CTA after 2 KB of static shared memory, the alignment slack
// and the epilogue staging
constexpr uint32_t ring_budget(int warps) { return warps == EPI_WARPS ? 192u * 1024u : 232448u - 2048u - 1024u - staging_bytes(warps); }
constexpr uint32_t SMEM_BUDGET = 192 * 1024;   // operand ring; + 32 KB of epilogue staging + alignment slack <= 227 KB
constexpr int kStagePitch = 32;                                      // floats per row of an epilogue staging tile (XOR-swizzled)
constexpr uint32_t STAGING_BYTES = EPI_WARPS * 32 * kStagePitch * 4; // one 32 x 32 fp32 tile per epilogue warp
// staging tile addressing: float4 block q of row r lives at block (q ^ (r & 7)) -- conflict-free for the row-per-thread
// writes (a quarter-warp writes 8 different blocks) and for the transposed reads (8 lanes read the 8 blocks of one row)
__device__ __forceinline__ int stage_off(int r, int q) { return r * kStagePitch + 4 * (q ^ (r & 7)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
// K-major, 128B-swizzled tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1, layout type 2.
// sw64: rows of 64 bytes (k-blocks of 16 floats), 8-row groups 512 bytes apart, layout type 4 (SWIZZLE_64B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, bool sw64 = false) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), 16 B
  d |= (uint64_t)((sw64 ? 512 : 1024) >> 4) << 32;   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)(sw64 ? 4 : 2) << 61;       // SWIZZLE_64B / SWIZZLE_128B
  return d;
}
// MN-major tf32 tile.  For 32-bit operands tcgen05 accepts exactly one MN-major shared-memory layout: 128-byte rows
// swizzled in 32-byte chunks (layout type SWIZZLE_128B_BASE32B; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), atom =
// 32 MN elements (one 128-byte row) x 4 K rows.  As written by TMA boxes of 32 floats (MN) x 32 rows (K): 4-row groups
// follow each other along K every 512 bytes (SBO), the next 32 MN elements start a new 4096-byte box (LBO).  One tf32
// MMA (K = 8) consumes two groups, so the K step is +1024 bytes.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(4096 >> 4) << 16;          // leading byte offset: stride between MN atoms (boxes)
  d |= (uint64_t)(512 >> 4) << 32;           // stride byte offset: stride between 4-row K groups
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;                    // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float& v) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v = __uint_as_float(r);
}

struct Params {
  float* C;
  int ldc;
  int M, N, K;
  const float* bias;
  int accumulate;
  const float* addend;  // optional [M, N] matrix added in the epilogue (leading dimension ld_add)
  int ld_add;
  int block_n;        // multiple of 16, <= 256
  int tmem_cols;      // power of two >= block_n
  int stages;
  int passes;         // 3 = 3xTF32 (fp32 parity), 1 = plain TF32 (or bf16)
  int bf16;           // operands are bf16 (kind::f16 MMA, 64-element k-blocks); passes == 1
  int a_mn, b_mn;     // operand is MN-major: A given as X (K, M) row-major (A = X^T) / B given as Y (K, N) row-major
                      // (B^T = Y, i.e. C = A Y): dW = x^T dy and y = x W straight from the natural layouts
  int kb_per_split;   // k-blocks per split-K slice
  int m_tiles, n_tiles, splits;   // persistent work list: splits x n_tiles x m_tiles items
  float* ws;          // split-K partials [splits][M][N] (NULL when splits == 1)
  // ---- fused scoring epilogues (K11/K13 + K14) ----
  int epi;                    // 0 store C, 1 count candidates beating the target (no C), 2 diagonal (pair scores)
  const float* tscore;        // epi 1: [M] target score of each query row
  const int* target;          // epi 1: [M] global candidate id of each query's target
  int* raw_count;             // epi 1: [M] += #{candidates of this tile beating the target}
  int col_offset;             // epi 1: global id of candidate row 0 of B (entity shard)
  int hyp;                    // score = hyp_score_from_dot(dot, x2[row], y2[col]) instead of dot
  const float* x2;            // [M] |q|^2
  const float* y2;            // [N] |e|^2   (epi 2: indexed by pair)
  const float* col_bias;      // [N] candidate bias or NULL (epi 2: indexed by pair)
  float hc, hproj_max;        // curvature, projection bound
  float hyp_ymax;             // epi 1 polynomial test: bound on |e|^2 (and |q|^2) its rounding band is valid for
  const float* scale_margin;  // device [scale, margin]
  const float* row_c;         // hyp == 2: [M] per-query curvature (true-distance branch); epi 2: indexed by pair
  float* diag_out;            // epi 2: [M]
  float* lse_max;             // epi 4: [n_tiles * EPI_WARPS/4][M] running maximum of this (tile, warp share) slice
  float* lse_sum;             // epi 4: same shape, sum of exp(score - max) over the slice
  // ---- fused layer epilogue (epi 3): UnionRGCNLayer apply (rgcn/layers.py:247-255) and, for the last layer, the
  //      time gate (src/rrgcn.py:176-178) straight out of the accumulator ----
  int lay_d;                  // hidden size: columns [0, lay_d) are layer outputs rrelu(acc); columns >= lay_d are stored
                              // unchanged to C[row, col - lay_d] (gate pre-activations x.W_time riding in the same GEMM)
  float* lay_raw;             // [*, lay_d] fp32 output rows (NULL: not needed)
  float* lay_hi;              // [*, lay_d] TF32 split of the output for the consumer GEMM (NULL: not needed)
  float* lay_lo;
  const int* row_idx;         // tile row r writes output row row_idx[r] (compact active-row GEMM); NULL = identity
  const int* skip_rows;       // rows with skip_rows[row] >= 0 are left alone (they belong to the compact GEMM)
  const float* gate_G;        // time gate: pre-activation [*, gate_ld] (NULL = no gate); needs N == lay_d <= block_n
  int gate_ld;
  const float* gate_bias;     // [lay_d]
  const float* gate_h;        // [*, lay_d] previous entity state
  int gate_norm;              // F.normalize the layer output first (layer_norm)
  // ---- fp32-A mode: A = [seg0 | seg1] along K, each segment an fp32 row-major matrix, rows optionally gathered ----
  int kblk;                   // fp32 elements per k-block row: 32 (128-byte swizzle rows) or 16 (64-byte rows: twice the
                              // ring slots in the same shared memory -- 3xTF32 tiles of 208-256 columns get 4 instead of 2)
  int a_f32;                  // 1: the tensor maps of A are unused, converter warps build the hi / lo tiles;
                              // 2: TMA delivers the fp32 k-block into the stage and the converter warps split it in place
  const float* a_ptr[2];      // segment base pointers (seg1 NULL when a_k[1] == 0)
  int a_ld[2];                // leading dimensions (floats, multiples of 4)
  int a_k[2];                 // segment lengths along K (multiples of 4); p.K = k-block-padded total
  const int* a_rows[2];       // tile row m reads source row a_rows[s][m] (NULL: m)
  int a_kb0;                  // k-blocks of segment 0 = ceil(a_k[0] / 32)
  // ---- conv-producer mode (a_f32 == 3, EPI 0): the A operand of the ConvTransE / ConvTransR fully-connected layer is never
  //      materialised.  A[b, c*d + i] = relu(bn1(conv1d_k3(bn0([x0[idx0[b]]; x1[idx1[b]]]))))[c, i] (src/decoder.py:81-90,
  //      eval-mode BatchNorm folded to scale / shift) is computed by the converter warps straight into the operand ring.
  //      K is walked in units (z, c) = (block of 16 positions, channel), z-major: a producer thread keeps the inputs of its
  //      (two rows, 4-position chunk) of a position block in registers for the ~C units that use them.  The weight is given in the same order (convfc_pack_weight:
  //      W'[n, 16 (z C + c) + j] = W[n, c d + 16 z + j], zero where 16 z + j >= d), so the B stream is a plain contiguous
  //      K walk with 64-byte aligned boxes; a split-K slice is a contiguous range of units. ----
  const float* cv_x0;         // [*, cv_d] first input table (entity rows)
  const float* cv_x1;         // [*, cv_d] second input table (relation rows / entity rows)
  const int64_t* cv_triples;  // [M, 3] query triples: row b reads cv_x0[triples[b][cv_col0]] and cv_x1[triples[b][cv_col1]]
  int cv_col0, cv_col1;
  const float* cv_bn0_s;      // (2) bn0 scale, (2) shift
  const float* cv_bn0_t;
  const float* cv_w;          // (C, 2, 3) conv weight
  const float* cv_b;          // (C) conv bias
  const float* cv_bn1_s;      // (C) bn1 scale, shift
  const float* cv_bn1_t;
  int cv_C, cv_d, cv_zb;      // channels, positions per row, position blocks = ceil(d / 16)
  unsigned long long* trace;  // optional in-kernel timeline: [gridDim.x][kTraceSlots] %globaltimer stamps (NULL = off)
};
constexpr int kConvMaxC = 64;
constexpr int kConvWarps = 8;                                    // producer warps of the conv mode: warps 2 .. 9
// Extra shared memory of the conv mode: the per-channel constants (12 floats each).  The inputs of a position block live in
// the producer threads' registers (two rows x two tables x six positions).
constexpr uint32_t kConvBytes = kConvMaxC * 12 * 4 + 256;
constexpr int kTraceSlots = 48;
__device__ __forceinline__ void trace_stamp(const Params& p, int slot) {
  if (p.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");   // "memory": stays on its side of barriers
    p.trace[(size_t)blockIdx.x * kTraceSlots + slot] = t;
  }
}

// sigmoid on the SFU (ex2.approx + rcp.approx, ~2 ulp; the epilogue warps have no spare issue slots for the IEEE path),
// without the denormal / huge-divisor fix-ups __expf / __fdividef wrap around the two operations (8 instructions instead
// of 14): ex2 flushes to +0 for x > ~87 (sigmoid -> 1) and overflows to +inf for x < ~-88 (rcp(inf) = 0)
__device__ __forceinline__ float lean_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
// F.rrelu in eval mode as max(x, slope x) (0 < slope < 1): the same value as the select for every finite x, 2 instructions
__device__ __forceinline__ float rrelu_max(float x) { return fmaxf(x, x * kRReluSlope); }

__device__ __forceinline__ float finish_score(const Params& p, float dot, int row, int col, float scale, float margin) {
  float v = dot;
  if (p.hyp == 2) v = hyp_dist_score_from_dot(dot, __ldg(p.x2 + row), __ldg(p.y2 + col), __ldg(p.row_c + row), scale, margin);
  else if (p.hyp) v = hyp_score_from_dot(dot, __ldg(p.x2 + row), __ldg(p.y2 + col), p.hc, p.hproj_max, scale, margin);
  if (p.col_bias) v = __fadd_rn(v, __ldg(p.col_bias + col));
  return v;
}

// Persistent kernel: one CTA per SM walks the (split, n-tile, m-tile) work list.  Three pipelines:
//   smem ring   full[s] / empty[s]          TMA producer  <-> MMA issuer   (continues across tiles)
//   TMEM slots  tmem_full[a] / tmem_empty[a] MMA issuer   <-> epilogue     (2 accumulators of block_n columns)
// so the epilogue of tile i (TMEM -> registers -> global / counting) overlaps the main loop of tile i+1.
template <int EPI>
__global__ void __launch_bounds__(num_threads<EPI>(), 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                 const Params p) {
  pdl_trigger();          // the next kernel may be scheduled; its own wait keeps it off our outputs
  if (threadIdx.x == 0) trace_stamp(p, 0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t raw_bar[8];             // a_f32 == 2: TMA (raw fp32 A + B) -> converter warps
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  constexpr int EW = epi_warps<EPI>();                     // epilogue warps of this instantiation
  __shared__ float norm_xchg[EW / 4][BLOCK_M];             // EPI 5: row-norm partial sums of the warps of a lane quarter

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int block_k = p.bf16 ? 2 * BLOCK_K : p.kblk;       // elements per k-block row (128 bytes; 64 with kblk == 16)
  const int total_kb = p.a_f32 == 3 ? p.cv_C * p.cv_zb : (p.K + block_k - 1) / block_k;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_mn * p.splits;

  // dynamic smem carve-up (1024-byte aligned tiles): per stage [A_hi][A_lo?][B_hi][B_lo?]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = BLOCK_M * (uint32_t)p.kblk * 4;
  const uint32_t b_bytes = (uint32_t)p.block_n * (uint32_t)p.kblk * 4;
  const bool three = p.passes == 3;
  const uint32_t stage_bytes = (three ? 2u : 1u) * (a_bytes + b_bytes);
  const uint32_t smem_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      // a_f32 == 1: TMA (expect_tx, B only) + one arrival per converter warp;  a_f32 == 2: the converter warps alone (they
      // pass on what TMA delivered under raw_bar)
      mbar_init(smem_u32(&full_bar[s]), p.a_f32 == 3 ? 1 + kConvWarps : p.a_f32 == 2 ? CVT_WARPS : p.a_f32 ? 1 + CVT_WARPS : 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
      mbar_init(smem_u32(&raw_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), 32 * EW);     // every epilogue thread arrives
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();             // barrier init and the TMEM allocation above overlap the tail of the previous kernel
  if (threadIdx.x == 0) trace_stamp(p, 1);

  // work item -> (split z, n tile, m tile); m fastest so that concurrently running CTAs share the B tile in L2
  auto decode = [&](int t, int& m0, int& n0, int& kb_beg, int& kb_end) {
    const int z = t / tiles_mn;
    const int r = t - z * tiles_mn;
    // the dimension with fewer tiles runs fastest: CTAs working at the same time then share the other operand's tile in L2
    // (many m-tiles x 2 n-tiles: the A tile, fetched from HBM once; few query tiles x many candidate tiles: the B tile)
    const bool n_fast = p.n_tiles < p.m_tiles;
    const int nt = n_fast ? r % p.n_tiles : r / p.m_tiles;
    const int mt = n_fast ? r / p.n_tiles : r - nt * p.m_tiles;
    m0 = mt * BLOCK_M;
    n0 = EPI == 2 ? m0 : nt * p.block_n;   // pair scores: diagonal tiles only
    if (p.a_f32 == 3) {          // conv-producer mode: equal shares of the (position block, channel) units
      kb_beg = (int)((long long)z * total_kb / p.splits);
      kb_end = (int)((long long)(z + 1) * total_kb / p.splits);
    } else {
      kb_beg = z * p.kb_per_split;
      kb_end = min(total_kb, kb_beg + p.kb_per_split);
    }
  };

  // ---- conv-producer mode (a_f32 == 3, see Params): kConvWarps warps (the two converter warps and the first six epilogue
  //      warps, which have nothing to do while a tile's main loop runs) compute the A k-blocks of a tile's units.
  //      thread <-> (16-byte chunk q of the 64-byte k-block row, rows rb and rb + 64).  The arithmetic of an element is that
  //      of convtranse_features_kernel (same fused multiply-adds in the same order). ----
  int cv_stage = 0, cv_m0 = -1, cv_z = -1;
  uint32_t cv_phase = 0;
  float cv_in[2][2][6];                                       // [row j][table][positions 16 z + 4 q - 1 .. + 4], bn0 applied
  int cv_idx[2][2] = {{-1, -1}, {-1, -1}};                     // source rows of the thread's two tile rows in both tables (-1: past M)
  auto conv_produce = [&](int m0, int kb_beg, int kb_end) {
    if constexpr (EPI == 0) {
      const int tid = (warp - 2) * 32 + lane;                 // 0 .. 32 kConvWarps - 1
      float* cpar = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + staging_bytes(EW));
      const int C = p.cv_C, d = p.cv_d;
      if (cv_m0 < 0) {
        // per-channel constants, 12 floats per channel: {w0 w1 w2 w3} {w4 w5 conv_b bn1_s} {bn1_t - - -}
        for (int i = tid; i < C; i += 32 * kConvWarps) {
#pragma unroll
          for (int k = 0; k < 6; ++k) cpar[i * 12 + k] = __ldg(p.cv_w + i * 6 + k);
          cpar[i * 12 + 6] = __ldg(p.cv_b + i);
          cpar[i * 12 + 7] = __ldg(p.cv_bn1_s + i);
          cpar[i * 12 + 8] = __ldg(p.cv_bn1_t + i);
        }
        asm volatile("bar.sync 8, %0;" ::"r"(32 * kConvWarps) : "memory");
      }
      const int q = tid & 3, rb = tid >> 2;
      constexpr int kRowStep = 8 * kConvWarps;                 // rows covered by the producers per pass (64: two rows a thread)
      static_assert(BLOCK_M / kRowStep == 2, "a producer thread holds the inputs of two rows");
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        const int zz = kb / C, c = kb - zz * C;
        if (m0 != cv_m0) {
          // rows of this tile: source row ids in both input tables
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int gm = m0 + rb + kRowStep * j;
            const bool rv = gm < p.M;
            cv_idx[j][0] = rv ? (int)p.cv_triples[3 * (size_t)gm + p.cv_col0] : -1;
            cv_idx[j][1] = rv ? (int)p.cv_triples[3 * (size_t)gm + p.cv_col1] : -1;
          }
        }
        if (m0 != cv_m0 || zz != cv_z) {
          // bn0([x0; x1]) of positions 16 z + 4 q - 1 .. 16 z + 4 q + 4 of the thread's two rows, kept in registers for the C
          // units of this position block (zero outside the row: the conv's padding)
          const int pos = 16 * zz + 4 * q;
          const float s0 = __ldg(p.cv_bn0_s), t0 = __ldg(p.cv_bn0_t), s1 = __ldg(p.cv_bn0_s + 1), t1 = __ldg(p.cv_bn0_t + 1);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int tab = 0; tab < 2; ++tab) {
              const float* src = (tab ? p.cv_x1 : p.cv_x0) + (size_t)(cv_idx[j][tab] < 0 ? 0 : cv_idx[j][tab]) * d;
              const float sc = tab ? s1 : s0, sh = tab ? t1 : t0;
              const bool in = cv_idx[j][tab] >= 0 && pos < d;
              const float4 v = in ? __ldg(reinterpret_cast<const float4*>(src + pos)) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float lft = (in && pos > 0) ? __ldg(src + pos - 1) : 0.f;
              const float rgt = (in && pos + 4 < d) ? __ldg(src + pos + 4) : 0.f;
              cv_in[j][tab][0] = (in && pos > 0) ? fmaf(lft, sc, sh) : 0.f;
              cv_in[j][tab][1] = in ? fmaf(v.x, sc, sh) : 0.f;
              cv_in[j][tab][2] = in ? fmaf(v.y, sc, sh) : 0.f;
              cv_in[j][tab][3] = in ? fmaf(v.z, sc, sh) : 0.f;
              cv_in[j][tab][4] = in ? fmaf(v.w, sc, sh) : 0.f;
              cv_in[j][tab][5] = (in && pos + 4 < d) ? fmaf(rgt, sc, sh) : 0.f;
            }
          }
          cv_m0 = m0; cv_z = zz;
        }
        const float4 wa = *reinterpret_cast<const float4*>(cpar + c * 12);
        const float4 wb = *reinterpret_cast<const float4*>(cpar + c * 12 + 4);
        const float bt = cpar[c * 12 + 8];
        const float w0 = wa.x, w1 = wa.y, w2 = wa.z, w3 = wa.w, w4 = wb.x, w5 = wb.y, cb = wb.z, bs = wb.w;
        const bool chunk_in = 16 * zz + 4 * q < d;       // d % 4 == 0: a chunk lies inside the row or outside it
        mbar_wait(smem_u32(&empty_bar[cv_stage]), cv_phase ^ 1);
        const uint32_t a_hi = smem_base + cv_stage * stage_bytes;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int r = rb + kRowStep * j;
          float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
          if (chunk_in) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float acc = cb;
              acc = fmaf(w0, cv_in[j][0][u], acc); acc = fmaf(w1, cv_in[j][0][u + 1], acc); acc = fmaf(w2, cv_in[j][0][u + 2], acc);
              acc = fmaf(w3, cv_in[j][1][u], acc); acc = fmaf(w4, cv_in[j][1][u + 1], acc); acc = fmaf(w5, cv_in[j][1][u + 2], acc);
              o[u] = fmaxf(fmaf(acc, bs, bt), 0.f);
            }
            h.x = rna_tf32(o[0]); h.y = rna_tf32(o[1]); h.z = rna_tf32(o[2]); h.w = rna_tf32(o[3]);
            l.x = rna_tf32(o[0] - h.x); l.y = rna_tf32(o[1] - h.y); l.z = rna_tf32(o[2] - h.z); l.w = rna_tf32(o[3] - h.w);
          }
          // 64-byte swizzle: 16-byte chunk index XOR address bits [7, 9) = (row >> 1) & 3
          const uint32_t ph = a_hi + r * 64 + ((q ^ ((r >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ph), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
          if (three)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ph + a_bytes), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[cv_stage])) : "memory");
        if (++cv_stage == p.stages) { cv_stage = 0; cv_phase ^= 1; }
      }
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pit = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++pit) {
        int m0, n0, kb_beg, kb_end;
        decode(t, m0, n0, kb_beg, kb_end);
        if (pit < 4) trace_stamp(p, 2 + pit);          // producer starts issuing tile pit
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const bool raw_a = p.a_f32 == 2;
          const uint32_t fb = smem_u32(raw_a ? &raw_bar[stage] : &full_bar[stage]);
          mbar_expect_tx(fb, raw_a ? a_bytes + (three ? 2u : 1u) * b_bytes : p.a_f32 ? (three ? 2u : 1u) * b_bytes : stage_bytes);
          uint32_t dst = smem_base + stage * stage_bytes;
          // fp32-A mode with two K segments: k-blocks never straddle a segment, so the weight columns of segment 1
          // start at a_k[0] (the zero-filled A tail of segment 0 cancels the columns its last k-block overlaps)
          const int k0 = (p.a_f32 == 1 && kb >= p.a_kb0) ? p.a_k[0] + (kb - p.a_kb0) * BLOCK_K : kb * block_k;
          // K-major operand: one box of 32 k-floats x rows; MN-major operand: boxes of 32 MN-elements x 32 reduction
          // rows, 4096 bytes each (inner coordinate = MN offset)
          auto load_op = [&](const CUtensorMap* tm, int mn, int r0, int nrows, uint32_t bytes) {
            if (mn) { for (int j = 0; j < nrows / 32; ++j) tma_load_2d(dst + j * 4096, tm, fb, r0 + 32 * j, k0); }
            else tma_load_2d(dst, tm, fb, k0, r0);
            dst += bytes;
          };
          if (raw_a) {                                            // the fp32 tile lands where its hi part will be
            tma_load_2d(dst, &tm_a_hi, fb, k0, m0);
            dst += (three ? 2u : 1u) * a_bytes;
          } else if (p.a_f32) dst += (three ? 2u : 1u) * a_bytes;  // the converter warps fill the A half of the stage
          else {
            load_op(&tm_a_hi, p.a_mn, m0, BLOCK_M, a_bytes);
            if (three) load_op(&tm_a_lo, p.a_mn, m0, BLOCK_M, a_bytes);
          }
          load_op(&tm_b_hi, p.b_mn, n0, p.block_n, b_bytes);
          if (three) load_op(&tm_b_lo, p.b_mn, n0, p.block_n, b_bytes);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (pit < 4) trace_stamp(p, 6 + pit);          // all loads of tile pit issued
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, K-major both, N>>3, M>>4
      // (kind::f16 with bf16 operands: format code 1 for A and B, same layout of the other fields)
      const uint32_t fmt = p.bf16 ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                             ((uint32_t)(BLOCK_M >> 4) << 24) | (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);   // a_major / b_major
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        int m0, n0, kb_beg, kb_end;
        decode(t, m0, n0, kb_beg, kb_end);
        const int slot = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(smem_u32(&tmem_empty_bar[slot]), acc_phase ^ 1);   // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + (uint32_t)(slot * (p.tmem_cols >> 1));
        uint32_t acc = 0;
        if (it < 6) trace_stamp(p, 10 + 3 * it);       // accumulator slot free, waiting for operands
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          if (kb == kb_beg && it < 6) trace_stamp(p, 11 + 3 * it);   // first operands landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa_hi = smem_base + stage * stage_bytes;
          const uint32_t sa_lo = sa_hi + a_bytes;
          const uint32_t sb_hi = sa_hi + (three ? 2u : 1u) * a_bytes;
          const uint32_t sb_lo = sb_hi + b_bytes;
          if (p.a_mn | p.b_mn) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // K step: one 8-row group (1024 B) of an MN-major tile, 32 B inside the swizzle row of a K-major one
              const uint32_t ka = p.a_mn ? k * 1024 : k * UMMA_K * 4, kbo = p.b_mn ? k * 1024 : k * UMMA_K * 4;
              auto da = [&](uint32_t base) { return p.a_mn ? make_desc_mn(base + ka) : make_desc(base + ka); };
              auto db = [&](uint32_t base) { return p.b_mn ? make_desc_mn(base + kbo) : make_desc(base + kbo); };
              if (three) {
                umma_tf32(tmem_acc, da(sa_lo), db(sb_hi), idesc, acc);
                acc = 1;
                umma_tf32(tmem_acc, da(sa_hi), db(sb_lo), idesc, acc);
              }
              umma_tf32(tmem_acc, da(sa_hi), db(sb_hi), idesc, acc);
              acc = 1;
            }
          } else {
            const bool sw64 = p.kblk == 16;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              if (k * UMMA_K >= p.kblk) break;        // 64-byte rows hold two K steps
              const uint32_t koff = k * UMMA_K * 4;   // bytes inside the swizzle row
              if (three) {
                umma_tf32(tmem_acc, make_desc(sa_lo + koff, sw64), make_desc(sb_hi + koff, sw64), idesc, acc);
                acc = 1;
                umma_tf32(tmem_acc, make_desc(sa_hi + koff, sw64), make_desc(sb_lo + koff, sw64), idesc, acc);
              }
              if (p.bf16) umma_f16(tmem_acc, make_desc(sa_hi + koff), make_desc(sb_hi + koff), idesc, acc);   // 16 bf16 = 32 bytes per step
              else umma_tf32(tmem_acc, make_desc(sa_hi + koff, sw64), make_desc(sb_hi + koff, sw64), idesc, acc);
              acc = 1;
            }
          }
          umma_commit(smem_u32(&empty_bar[stage]));      // frees the smem slot once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tmem_full_bar[slot]));     // accumulator of this tile complete
        if (it < 6) trace_stamp(p, 12 + 3 * it);       // all MMAs of the tile issued
      }
    }
  } else if (warp < EPI_WARP0) {
    // ===================== A-operand converters (fp32-A mode) =====================
    // Each warp owns 64 rows of the 128-row tile; a lane owns ONE 16-byte chunk column (c16) of 16 rows, so a warp-wide
    // load instruction reads 4 rows x 128 contiguous bytes.  The raw fp32 chunks of k-block kb+1 are loaded into
    // registers BEFORE the warp waits for the stage of k-block kb to drain: a stage turns around in the time of the
    // split + 32 shared-memory stores, not in a global-memory round trip.  hi = rna_tf32(x) and lo = rna_tf32(x - hi) go
    // to the swizzled places TMA would have written them to; fence.proxy.async publishes them to the tensor core.
    if (p.a_f32 == 2) {
      // Raw-tile form (one K segment, rows not gathered): TMA has put the fp32 k-block where the hi tile belongs, in the
      // swizzled layout the tensor core reads.  The split is elementwise, so the layout does not matter here: every
      // thread rewrites its 16-byte chunks in place (hi) and writes the same offsets of the lo tile.  The loads are TMA's
      // (a whole stage in flight, like the pre-split operands) -- the register-staged form below keeps only 16 KB per SM
      // in flight, which is latency-bound once the fp32 rows come from HBM.
      const int tid = (warp - 2) * 32 + lane;
      constexpr int GRP = 8;                                  // 16-byte chunks a thread holds at a time
      const int groups = (int)(a_bytes / 16) / (32 * CVT_WARPS * GRP);          // 2 (k-blocks of 32 floats) or 1 (16)
      int stage_c = 0;
      uint32_t phase_c = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0, kb_beg, kb_end;
        decode(t, m0, n0, kb_beg, kb_end);
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          mbar_wait(smem_u32(&raw_bar[stage_c]), phase_c);
          const uint32_t base = smem_base + stage_c * stage_bytes + tid * 16;
          for (int half = 0; half < groups; ++half) {
            float4 v[GRP];
#pragma unroll
            for (int j = 0; j < GRP; ++j) {
              const uint32_t a = base + (half * GRP + j) * (32 * CVT_WARPS * 16);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w) : "r"(a) : "memory");
            }
#pragma unroll
            for (int j = 0; j < GRP; ++j) {
              const uint32_t a = base + (half * GRP + j) * (32 * CVT_WARPS * 16);
              float4 h, l;
              h.x = rna_tf32(v[j].x); h.y = rna_tf32(v[j].y); h.z = rna_tf32(v[j].z); h.w = rna_tf32(v[j].w);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
              if (three) {
                l.x = rna_tf32(v[j].x - h.x); l.y = rna_tf32(v[j].y - h.y); l.z = rna_tf32(v[j].z - h.z); l.w = rna_tf32(v[j].w - h.w);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a + a_bytes), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[stage_c])) : "memory");
          if (++stage_c == p.stages) { stage_c = 0; phase_c ^= 1; }
        }
      }
    } else if (EPI == 0 && p.a_f32 == 3) {
      // Conv-producer form: these two warps and the first kConvWarps - 2 epilogue warps build the A tiles (conv_produce)
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0, kb_beg, kb_end;
        decode(t, m0, n0, kb_beg, kb_end);
        conv_produce(m0, kb_beg, kb_end);
      }
    } else if (EPI != 5 && p.a_f32) {       // (the 128-register time-gate instantiation takes fp32 A in the raw-tile form only)
      const int cw = warp - 2;
      const int c16 = lane & 7, rsub = lane >> 3;
      constexpr int RPL = BLOCK_M / CVT_WARPS / 4;        // rows per lane (16)
      int ti = blockIdx.x, mi0 = 0, ni0 = 0, kbi = 0, kbi_end = 0;      // load cursor
      int tc_ = blockIdx.x, mc0 = 0, nc0 = 0, kbc = 0, kbc_end = 0;     // store cursor (same sequence, one k-block behind)
      int stage_c = 0;
      uint32_t phase_c = 0;
      if (ti < total_tiles) { decode(ti, mi0, ni0, kbi, kbi_end); decode(tc_, mc0, nc0, kbc, kbc_end); }
      auto load_kb = [&](float4 (&v)[RPL]) {
        const int s = kbi >= p.a_kb0 ? 1 : 0;
        const int* rows = p.a_rows[s];
        const int kloc = (kbi - (s ? p.a_kb0 : 0)) * BLOCK_K + 4 * c16;
        const bool kval = kloc < p.a_k[s];
        const float* base = p.a_ptr[s] + kloc;
        const size_t ld = (size_t)p.a_ld[s];
        int src[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          const int m = mi0 + cw * (BLOCK_M / CVT_WARPS) + 4 * j + rsub;
          src[j] = (kval && m < p.M) ? (rows ? __ldg(rows + m) : m) : -1;
        }
#pragma unroll
        for (int j = 0; j < RPL; ++j)
          v[j] = src[j] >= 0 ? __ldg(reinterpret_cast<const float4*>(base + (size_t)src[j] * ld)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (++kbi == kbi_end) {
          ti += gridDim.x;
          if (ti < total_tiles) decode(ti, mi0, ni0, kbi, kbi_end);
        }
      };
      // rna_tf32 on the bit pattern: add half an ulp of the 10-bit mantissa to the magnitude, drop the low 13 bits.
      // Bit-identical to cvt.rna.tf32.f32 for every finite value (the PTX instruction costs 4 SASS instructions for its
      // Inf/NaN handling, 12 per element for the split; this is 5).
      auto rna = [](float a) { return rna_tf32(a); };
      auto store_kb = [&](const float4 (&v)[RPL]) {
        mbar_wait(smem_u32(&empty_bar[stage_c]), phase_c ^ 1);
        const uint32_t a_hi = smem_base + stage_c * stage_bytes;
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          const int r = cw * (BLOCK_M / CVT_WARPS) + 4 * j + rsub;
          const uint32_t ph = a_hi + r * 128 + ((c16 ^ (r & 7)) << 4);
          float4 h, l;
          h.x = rna(v[j].x); h.y = rna(v[j].y); h.z = rna(v[j].z); h.w = rna(v[j].w);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ph), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
          if (three) {
            l.x = rna(v[j].x - h.x); l.y = rna(v[j].y - h.y); l.z = rna(v[j].z - h.z); l.w = rna(v[j].w - h.w);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ph + a_bytes), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[stage_c])) : "memory");
        if (++stage_c == p.stages) { stage_c = 0; phase_c ^= 1; }
        if (++kbc == kbc_end) {
          tc_ += gridDim.x;
          if (tc_ < total_tiles) decode(tc_, mc0, nc0, kbc, kbc_end);
        }
      };
      float4 b0[RPL], b1[RPL];
      if (ti < total_tiles) load_kb(b0);
      while (tc_ < total_tiles) {
        if (ti < total_tiles) load_kb(b1);
        store_kb(b0);
        if (tc_ >= total_tiles) break;
        if (ti < total_tiles) load_kb(b0);
        store_kb(b1);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global / counts =====================
    const int quarter = warp & 3;                      // TMEM lanes [32*quarter, 32*quarter+32)
    const int part = (warp - EPI_WARP0) >> 2;                  // which share of the 32-column chunks this warp takes
    constexpr int kParts = EW / 4;
    float* stage = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes) + (warp - EPI_WARP0) * (32 * kStagePitch);
    float scale = 1.f, margin = 0.f;
    if (EPI != 0 && EPI != 3 && EPI != 5 && p.hyp) { scale = __ldg(p.scale_margin); margin = __ldg(p.scale_margin + 1); }
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      int m0, n0, kb_beg, kb_end;
      decode(t, m0, n0, kb_beg, kb_end);
      if (EPI == 0 && p.a_f32 == 3 && warp < 2 + kConvWarps) conv_produce(m0, kb_beg, kb_end);   // idle until the tile is done
      const int slot = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int row = m0 + quarter * 32 + lane;
      mbar_wait(smem_u32(&tmem_full_bar[slot]), acc_phase);
      if (warp == EPI_WARP0 && lane == 0 && it < 6) trace_stamp(p, 28 + 2 * it);     // accumulator complete
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_acc = tmem_base + (uint32_t)(slot * (p.tmem_cols >> 1)) + ((uint32_t)(quarter * 32) << 16);
      if constexpr (EPI == 0) {
        // Store epilogue.  The accumulator arrives one ROW per thread (TMEM lane = row); written like that, every
        // store instruction would touch 32 rows x 16 bytes.  Each warp therefore transposes its 32 x 32 chunk through a
        // private shared-memory tile (pitch 36 floats: conflict-free float4 writes and reads) so that a store
        // instruction covers 4 rows x 128 contiguous bytes -- full sectors, full lines.
        const bool split = p.splits > 1;
        const int z = t / tiles_mn;
        float* out = split ? p.ws + (size_t)z * (size_t)p.M * (size_t)p.N : p.C;
        const int ldo = split ? p.N : p.ldc;
        const bool vec_ok = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                            (!p.addend || (((p.ld_add & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.addend) & 15) == 0)));
        const int sub_r = lane >> 3, sub_q = lane & 7, sub_c = sub_q * 4;
        for (int c0 = 32 * part; c0 < p.block_n; c0 += 32 * kParts) {
          float v[32];
          tmem_ld32(tmem_acc + (uint32_t)c0, v);
          const int gn0 = n0 + c0;
          const int ncols = min(32, min(p.block_n - c0, p.N - gn0));   // columns of this chunk owned by this tile
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stage + stage_off(lane, q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          __syncwarp();
          if (ncols > 0) {
            const int col = gn0 + sub_c;
            const int nc = min(4, ncols - sub_c);              // valid columns of this lane's float4 (<= 0: none)
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!split && p.bias && nc > 0) {
              b4.x = __ldg(p.bias + col);
              if (nc > 1) b4.y = __ldg(p.bias + col + 1);
              if (nc > 2) b4.z = __ldg(p.bias + col + 2);
              if (nc > 3) b4.w = __ldg(p.bias + col + 3);
            }
            // every global read of the chunk (accumulate / addend rows) is issued before the first store: the stores
            // may alias them as far as the compiler knows, and one exposed L2 round trip per row group is 8 per chunk
            float4 acc4[8], add4[8];
            const bool fast = vec_ok && nc == 4 && !split;
            if (fast && (p.accumulate || p.addend)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int grow = m0 + quarter * 32 + 4 * i + sub_r;
                acc4[i] = add4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (grow < p.M) {
                  if (p.accumulate) acc4[i] = *reinterpret_cast<const float4*>(out + (size_t)grow * ldo + col);
                  if (p.addend) add4[i] = *reinterpret_cast<const float4*>(p.addend + (size_t)grow * p.ld_add + col);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + sub_r;
              const int grow = m0 + quarter * 32 + r;
              if (grow < p.M && nc > 0) {
                float4 a = *reinterpret_cast<const float4*>(stage + stage_off(r, sub_q));
                float* dst = out + (size_t)grow * ldo + col;
                if (vec_ok && nc == 4) {
                  if (!split) {
                    a = f4_add(a, b4);
                    if (p.accumulate) a = f4_add(a, acc4[i]);
                    if (p.addend) a = f4_add(a, add4[i]);
                  }
                  st4(dst, a);
                } else {
                  float e[4] = {a.x, a.y, a.z, a.w};
                  const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                  for (int j = 0; j < nc; ++j) {
                    float x = e[j];
                    if (!split) {
                      x += bb[j];
                      if (p.accumulate) x += dst[j];
                      if (p.addend) x += p.addend[(size_t)grow * p.ld_add + col + j];
                    }
                    dst[j] = x;
                  }
                }
              }
            }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == 1) {
        // fused K14: count the candidates of this tile that rank ahead of the row's target; nothing is stored
        const bool rv = row < p.M;
        const float st_ = rv ? __ldg(p.tscore + row) : 0.f;
        const int tg = rv ? __ldg(p.target + row) - p.col_offset : -1;
        int cnt = 0;
        const bool plain = !p.hyp && !p.col_bias;          // hoisted: no per-element uniform branches in the hot loop
        // ---- hyperbolic (RotH / MuRP form) score without a bias: threshold test instead of the score -------------------
        // score(dot) > st  <=>  |(-q) (+) e|^2 < T,  T = margin - st / scale  (scale > 0, below the projection clamp)
        //                  <=>  num - T den^2 < 0  with the Moebius numerator / denominator of hyp_score_from_dot.
        // For a fixed query row, num - T den^2 is a quadratic polynomial in (dot, y = |e|^2):
        //   diff = dot (k1 dot + k2 y + k4) + (y (k3 y + k5) + k6)                         5 FMAs, no division, no sqrt
        // with six per-row coefficients formed in double.  The sign of diff decides the comparison whenever |diff| exceeds
        // `band`, a bound on the rounding of BOTH evaluations (this polynomial and the IEEE sequence of
        // hyp_score_from_dot), evaluated per 32-candidate chunk from the largest |e|^2 of the chunk; candidates inside
        // the band -- ties included -- and rows whose target sits at the projection clamp take the exact score, so every
        // count equals the dense path's bit for bit.
        const bool poly = p.hyp == 1 && !p.col_bias && p.hyp_ymax > 0.f;
        // band_on: 0 = every candidate beats (target below the clamped minimum; k6 = -1), 1 = polynomial, 2 = exact only
        float k1 = 0.f, k2 = 0.f, k3 = 0.f, k4 = 0.f, k5 = 0.f, k6 = -1.f, x2r = 0.f, sx = 0.f, e0 = 0.f;
        int band_on = rv ? 2 : 0;
        if (poly && rv) {
          x2r = __ldg(p.x2 + row);
          const double c = (double)p.hc, x2 = (double)x2r, pm2 = (double)p.hproj_max * (double)p.hproj_max;
          const double T = (double)margin - (double)st_ / (double)scale;
          if (scale > 0.f && x2 >= 0.0 && c * x2 <= 1.0 && T == T && fabs(T) < 1e30) {
            if (T > pm2 * (1.0 + 1e-5)) {
              band_on = 0;
            } else if (T < pm2 * (1.0 - 1e-5)) {
              const double b = 1.0 - c * x2, kap = c * c * x2, del = 1.0 + (double)kEps;
              // s = -dot:  num = x2 a^2 + 2 a b s + b^2 y,  a = 1 + 2 c s + c y;  den = del + 2 c s + kap y
              const double q1 = 4.0 * c * c * x2 + 4.0 * b * c - 4.0 * c * c * T;          // s^2
              const double q2 = 4.0 * c * c * x2 + 2.0 * b * c - 4.0 * c * kap * T;        // s y
              const double q3 = c * c * x2 - kap * kap * T;                                // y^2
              const double q4 = 4.0 * c * x2 + 2.0 * b - 4.0 * c * del * T;                // s
              const double q5 = 2.0 * c * x2 + b * b - 2.0 * del * kap * T;                // y
              const double q6 = x2 - del * del * T;                                        // 1
              k1 = (float)q1; k2 = (float)-q2; k3 = (float)q3; k4 = (float)-q4; k5 = (float)q5; k6 = (float)q6;
              sx = (float)(sqrt(x2) * 1.0001);
              // weight of a^2 / den^2 in the magnitudes the IEEE sequence rounds: |q|^2 and |T|, |margin|, |st / scale|
              e0 = (float)((x2 + fabs(T) + fabs((double)margin) + fabs((double)st_ / (double)scale)) * 1.0001);
              band_on = 1;
            }
          }
        }
        for (int c0 = 32 * part; c0 < p.block_n; c0 += 32 * kParts) {
          float v[32];
          tmem_ld32(tmem_acc + (uint32_t)c0, v);
          const int gn0 = n0 + c0;
          const int ncols = min(32, min(p.block_n - c0, p.N - gn0));
          if (poly) {
            if (ncols > 0) {                                 // uniform across the warp
              const float ycol = lane < ncols ? __ldg(p.y2 + gn0 + lane) : 0.f;      // lane j holds |e_j|^2 of the chunk
              float ymx = ycol;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) ymx = fmaxf(ymx, __shfl_xor_sync(0xffffffffu, ymx, o));
              // rounding band of this (row, chunk): 64 ulp of the summed magnitudes of every monomial either evaluation
              // rounds, with |dot| <= u = |q| sqrt(ymax) (Cauchy-Schwarz) and |e|^2 <= ymax, a, den <= A = 1 + 2cu + c ymax
              float bnd = INFINITY;
              if (band_on == 0) bnd = 0.f;
              else if (band_on == 1 && ymx >= 0.f && p.hc * ymx <= 1.0001f) {
                const float u = sx * (sqrtf(ymx) * 1.0001f);
                const float A = fmaf(2.0f * p.hc, u, fmaf(p.hc, ymx, 1.0001f));
                float mag = fmaf(fabsf(k1), u * u, fmaf(fabsf(k2), u * ymx, fmaf(fabsf(k3), ymx * ymx,
                            fmaf(fabsf(k4), u, fmaf(fabsf(k5), ymx, fabsf(k6))))));
                mag += fmaf(A * A, e0, fmaf(2.0f * A, u, ymx));
                bnd = mag * (64.0f * 5.9604644775390625e-08f * 1.01f);
              }
              const int tj = tg - gn0;
              // pass 1, branch-free: decide every candidate outside its band, remember the others in a bit mask
              uint32_t amb = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float y = __shfl_sync(0xffffffffu, ycol, j);
                const float dotv = v[j];
                const float diff = fmaf(dotv, fmaf(k1, dotv, fmaf(k2, y, k4)), fmaf(y, fmaf(k3, y, k5), k6));
                const int inband = (int)!(fabsf(diff) > bnd);          // inside the rounding band, or NaN
                const int valid = (int)(j < ncols) & (int)(j != tj) & (int)rv;
                amb |= (uint32_t)(inband & valid) << j;
                cnt += (int)(diff < 0.f) & (inband ^ 1) & valid;
              }
              // pass 2, rare: the exact IEEE score for the in-band candidates.  One rolled loop over the columns any lane
              // needs (the accumulator column is re-read from TMEM, warp-uniformly), so the division / sqrt sequence
              // exists once in the instruction stream instead of 32 times inside the hot loop
              uint32_t any = __reduce_or_sync(0xffffffffu, amb);
              while (any) {
                const int j = __ffs(any) - 1;
                any &= any - 1;
                float dotv;
                tmem_ld1(tmem_acc + (uint32_t)(c0 + j), dotv);
                const float y = __shfl_sync(0xffffffffu, ycol, j);
                if ((amb >> j) & 1u) {
                  const float sc = hyp_score_from_dot(dotv, x2r, y, p.hc, p.hproj_max, scale, margin);
                  cnt += (int)(sc > st_) | ((int)(sc == st_) & (int)(j < tj));
                }
              }
            }
            continue;
          }
          if (rv && ncols > 0) {
            if (!plain) {
              if (p.hyp == 2) {
                const float x2q = __ldg(p.x2 + row), cq = __ldg(p.row_c + row);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncols) v[j] = hyp_dist_score_from_dot(v[j], x2q, __ldg(p.y2 + gn0 + j), cq, scale, margin);
              } else if (p.hyp) {
                const float x2q = __ldg(p.x2 + row);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncols) v[j] = hyp_score_from_dot(v[j], x2q, __ldg(p.y2 + gn0 + j), p.hc, p.hproj_max, scale, margin);
              }
              if (p.col_bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (j < ncols) v[j] = __fadd_rn(v[j], __ldg(p.col_bias + gn0 + j));
              }
            }
            // branch-free stable-rank contribution: (s > st) | (s == st & col < t), masked by validity and col != t
            const int tj = tg - gn0;                        // target position inside this chunk (any int)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int beats = (int)(v[j] > st_) | ((int)(v[j] == st_) & (int)(j < tj));
              cnt += beats & (int)(j < ncols) & (int)(j != tj);
            }
          }
        }
        if (rv && cnt) atomicAdd(p.raw_count + row, cnt);
      } else if constexpr (EPI == 3) {
        // Fused layer epilogue, same transposed staging as the store epilogue: columns < lay_d are layer outputs
        // rrelu(acc) written as fp32 and / or TF32 split (rows optionally scattered through row_idx, rows owned by the
        // compact GEMM skipped), columns >= lay_d are stored unchanged to C (gate pre-activations).  Per-row quantities
        // are computed by the thread that owns the row in TMEM and fetched by shuffle.  (The time-gate form is EPI 5.)
        const bool rv = row < p.M;
        int my_orow = rv ? row : 0;
        int my_skip = rv ? 0 : 1;
        if (rv && p.row_idx) my_orow = __ldg(p.row_idx + row);
        if (rv && p.skip_rows && __ldg(p.skip_rows + row) >= 0) my_skip = 1;
        const int sub_r = lane >> 3, sub_q = lane & 7, sub_c = sub_q * 4;
        {
          // No gate: nothing is read from global memory, and the per-row facts of the 8 rows a lane serves in the
          // transposed layout (output row, write flags) are fetched once per tile instead of once per 32-column chunk.
          const unsigned rows_in = __ballot_sync(0xffffffffu, rv) >> sub_r;              // bit 4 i: row 4 i + sub_r < M
          const unsigned rows_out = __ballot_sync(0xffffffffu, !my_skip) >> sub_r;       // ... and not left to the compact GEMM
          int orow8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) orow8[i] = __shfl_sync(0xffffffffu, my_orow, 4 * i + sub_r);
          for (int c0 = 32 * part; c0 < p.block_n; c0 += 32 * kParts) {
            float v[32];
            tmem_ld32(tmem_acc + (uint32_t)c0, v);
            const int gn0 = n0 + c0;
            const int ncols = min(32, min(p.block_n - c0, p.N - gn0));
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(stage + stage_off(lane, q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            __syncwarp();
            const int col = gn0 + sub_c;
            if (sub_c < ncols) {                               // N and lay_d are multiples of 4: whole float4 or nothing
              if (col < p.lay_d) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if ((rows_out >> (4 * i)) & 1u) {
                    float4 a = *reinterpret_cast<const float4*>(stage + stage_off(4 * i + sub_r, sub_q));
                    a.x = rrelu_max(a.x); a.y = rrelu_max(a.y); a.z = rrelu_max(a.z); a.w = rrelu_max(a.w);
                    const size_t o = (size_t)orow8[i] * p.lay_d + col;
                    if (p.lay_raw) st4(p.lay_raw + o, a);
                    if (p.lay_hi) {
                      float4 h, l;
                      split_tf32_1(a.x, h.x, l.x); split_tf32_1(a.y, h.y, l.y);
                      split_tf32_1(a.z, h.z, l.z); split_tf32_1(a.w, h.w, l.w);
                      st4(p.lay_hi + o, h);
                      st4(p.lay_lo + o, l);
                    }
                  }
                }
              } else {
                const int cc = col - p.lay_d;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if ((rows_in >> (4 * i)) & 1u)
                    st4(p.C + (size_t)orow8[i] * p.ldc + cc, *reinterpret_cast<const float4*>(stage + stage_off(4 * i + sub_r, sub_q)));
                }
              }
            }
            __syncwarp();
          }
        }
      } else if constexpr (EPI == 5) {
        // Time-gate epilogue of a snapshot's last layer, specialised (gemm_tf32_layer picks it when the tile spans the
        // whole row and all three outputs exist):  h' = h + s (normalize(rrelu(acc)) - h),  s = sigmoid(G + b)
        // (src/rrgcn.py:176-178 with the UnionRGCNLayer activation, rgcn/layers.py:253-255), written as fp32 + TF32 split.
        // Same staging as the general layer epilogue; what is different is the instruction count: the per-row facts
        // (output row, 1/norm, write flag) of the 8 rows a lane serves in the transposed layout are fetched once per
        // tile, not per 32-column chunk, the sigmoid is two bare SFU operations, the blend one subtract and one FMA.
        const bool rv = row < p.M;
        int my_orow = rv ? row : 0;
        bool my_wr = rv;
        if (rv && p.row_idx) my_orow = __ldg(p.row_idx + row);
        if (rv && p.skip_rows && __ldg(p.skip_rows + row) >= 0) my_wr = false;
        float my_nrm = 1.f;
        if (p.gate_norm) {
          float ss = 0.f;
          for (int c0 = 32 * part; c0 < p.block_n; c0 += 32 * kParts) {
            float v[32];
            tmem_ld32(tmem_acc + (uint32_t)c0, v);
            const int ncols = p.lay_d - c0;
            if (ncols >= 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { const float r = rrelu_max(v[j]); ss = fmaf(r, r, ss); }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) { const float r = rrelu_max(v[j]); ss = j < ncols ? fmaf(r, r, ss) : ss; }
            }
          }
          float* slot = &norm_xchg[0][0];
          slot[part * BLOCK_M + quarter * 32 + lane] = ss;
          asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(32 * kParts) : "memory");
          float tot = 0.f;
#pragma unroll
          for (int q = 0; q < kParts; ++q) tot += slot[q * BLOCK_M + quarter * 32 + lane];
          asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(32 * kParts) : "memory");   // slots free for the next tile
          my_nrm = 1.0f / fmaxf(sqrtf(tot), 1e-12f);    // F.normalize: x / max(|x|, 1e-12), applied as a multiply
        }
        const int sub_r = lane >> 3, sub_q = lane & 7, sub_c = sub_q * 4;
        const unsigned wr_all = __ballot_sync(0xffffffffu, my_wr);
        int orow8[8];
        float nrm8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          orow8[i] = __shfl_sync(0xffffffffu, my_orow, 4 * i + sub_r);
          nrm8[i] = __shfl_sync(0xffffffffu, my_nrm, 4 * i + sub_r);
        }
        const unsigned wr8 = wr_all >> sub_r;               // bit 4 i: row 4 i + sub_r is written
        const int ld_h = p.lay_d, ld_g = p.gate_ld;
        for (int c0 = 32 * part; c0 < p.block_n; c0 += 32 * kParts) {
          float v[32];
          tmem_ld32(tmem_acc + (uint32_t)c0, v);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stage + stage_off(lane, q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          __syncwarp();
          const int col = c0 + sub_c;
          if (col < ld_h) {                                  // lay_d is a multiple of 4: a whole float4 or nothing
            const float4 b4 = ldg4(p.gate_bias + col);
            // two groups of 4 rows: 8 float4 loads in flight per thread (the instantiation lives in 128 registers)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float4 g4[4], h4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i = 4 * half + j;
                if ((wr8 >> (4 * i)) & 1u) {
                  g4[j] = *reinterpret_cast<const float4*>(p.gate_G + (size_t)orow8[i] * ld_g + col);
                  h4[j] = *reinterpret_cast<const float4*>(p.gate_h + (size_t)orow8[i] * ld_h + col);
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i = 4 * half + j;
                if ((wr8 >> (4 * i)) & 1u) {
                  float4 a = *reinterpret_cast<const float4*>(stage + stage_off(4 * i + sub_r, sub_q));
                  const float nrm = nrm8[i];
                  a.x = rrelu_max(a.x) * nrm; a.y = rrelu_max(a.y) * nrm; a.z = rrelu_max(a.z) * nrm; a.w = rrelu_max(a.w) * nrm;
                  const float sx = lean_sigmoid(g4[j].x + b4.x), sy = lean_sigmoid(g4[j].y + b4.y);
                  const float sz = lean_sigmoid(g4[j].z + b4.z), sw = lean_sigmoid(g4[j].w + b4.w);
                  a.x = fmaf(sx, a.x - h4[j].x, h4[j].x); a.y = fmaf(sy, a.y - h4[j].y, h4[j].y);
                  a.z = fmaf(sz, a.z - h4[j].z, h4[j].z); a.w = fmaf(sw, a.w - h4[j].w, h4[j].w);
                  const size_t o = (size_t)orow8[i] * ld_h + col;
                  st4(p.lay_raw + o, a);
                  if (p.lay_hi) {                              // uniform: the fp32-A consumers need no split copy
                    float4 h, l;
                    split_tf32_1(a.x, h.x, l.x); split_tf32_1(a.y, h.y, l.y);
                    split_tf32_1(a.z, h.z, l.z); split_tf32_1(a.w, h.w, l.w);
                    st4(p.lay_hi + o, h);
                    st4(p.lay_lo + o, l);
                  }
                }
              }
            }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == 4) {
        // Streaming log-sum-exp of the score row (the training loss heads: CrossEntropy over all entities without the
        // (B,N) logits, src/rrgcn.py:217-218, hyperbolic_decoder.py:182-307).  Each (n-tile, warp share) emits the
        // maximum and the sum of exp(score - max) of its slice of the row; ce_from_lse folds the slices.
        const bool rv = row < p.M;
        float m_run = -INFINITY, s_run = 0.f;
        for (int c0 = 32 * part; c0 < p.block_n; c0 += 32 * kParts) {
          float v[32];
          tmem_ld32(tmem_acc + (uint32_t)c0, v);
          const int gn0 = n0 + c0;
          const int ncols = min(32, min(p.block_n - c0, p.N - gn0));
          if (rv && ncols > 0) {
            if (p.hyp == 2) {
              const float x2r = __ldg(p.x2 + row), cq = __ldg(p.row_c + row);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] = hyp_dist_score_from_dot(v[j], x2r, __ldg(p.y2 + gn0 + j), cq, scale, margin);
            } else if (p.hyp) {
              const float x2r = __ldg(p.x2 + row);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] = hyp_score_from_dot(v[j], x2r, __ldg(p.y2 + gn0 + j), p.hc, p.hproj_max, scale, margin);
            }
            if (p.col_bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < ncols) v[j] = __fadd_rn(v[j], __ldg(p.col_bias + gn0 + j));
            }
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = j < ncols ? fmaxf(cm, v[j]) : cm;
            if (cm > m_run) { s_run *= expf(m_run - cm); m_run = cm; }     // exp(-inf) = 0 on the first chunk
#pragma unroll
            for (int j = 0; j < 32; ++j) s_run += j < ncols ? expf(v[j] - m_run) : 0.f;
          }
        }
        if (rv) {
          const int nt = n0 / p.block_n;
          const size_t slot = ((size_t)(nt * kParts + part)) * (size_t)p.M + row;
          p.lse_max[slot] = m_run;
          p.lse_sum[slot] = s_run;
        }
      } else {
        // pair scores: the tile is a diagonal block of A' . B'^T (pair p = row), keep acc[r][r]
        float v[32];
        tmem_ld32(tmem_acc + (uint32_t)(quarter * 32), v);
        float dsel = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j == lane) dsel = v[j];
        if (row < p.M && part == 0) p.diag_out[row] = finish_score(p, dsel, row, row, scale, margin);
      }
      // this thread's TMEM reads of the slot are complete (tcgen05.wait::ld inside tmem_ld32): hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[slot])) : "memory");
      if (warp == EPI_WARP0 && lane == 0 && it < 6) trace_stamp(p, 29 + 2 * it);     // epilogue of the tile done (this warp)
    }
  }
  // exit stamp of the CTA = the latest of its epilogue warps (the producer / MMA lanes reach the barrier below long
  // before the last accumulator has been drained, so a stamp by thread 0 would not mean "done")
  if (p.trace && warp >= EPI_WARP0 && lane == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
    atomicMax(p.trace + (size_t)blockIdx.x * kTraceSlots + 40, t);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// hi = rna_tf32(x), lo = rna_tf32(x - hi): both exactly representable in tf32, hi + lo == x to 2^-22 relative.
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n4) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  float4 h, l;
  auto split1 = [](float a, float& hh, float& ll) { split_tf32_1(a, hh, ll); };
  split1(v.x, h.x, l.x); split1(v.y, h.y, l.y); split1(v.z, h.z, l.z); split1(v.w, h.w, l.w);
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread; the bf16 scoring mode's operand conversion
__global__ void to_bf16_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, size_t n8) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
  auto pack = [](float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  };
  uint4 o;
  o.x = pack(a.x, a.y); o.y = pack(a.z, a.w); o.z = pack(b.x, b.y); o.w = pack(b.z, b.w);
  reinterpret_cast<uint4*>(out)[i] = o;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major [rows, cols] with leading dimension ld; box = 32 floats x box_rows, 128B swizzle, zero OOB fill.
static int make_map(CUtensorMap* m, const void* ptr, int rows, int cols, int ld, int box_rows, bool bf16 = false,
                    bool atom32 = false, int kblk = BLOCK_K) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("gemm_tf32: cuTensorMapEncodeTiled entry point unavailable"); return REGCN_ERR_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(bf16 ? 2 * BLOCK_K : kblk), (cuuint32_t)box_rows};   // kblk 16: 64-byte rows
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (!bf16 && kblk == 16) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("gemm_tf32: cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld); return REGCN_ERR_DIM; }
  return REGCN_OK;
}

}  // namespace tc

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, float* __restrict__ C, int ldc,
                                     int M, int N, const float* __restrict__ bias, int accumulate);

int split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t st) {
  if (!x || !hi || !lo) { set_last_error("split_tf32: null pointer"); return REGCN_ERR_NULL; }
  if (n & 3) { set_last_error("split_tf32: element count must be a multiple of 4"); return REGCN_ERR_DIM; }
  if (!n) return REGCN_OK;
  const size_t n4 = n / 4;
  launch_k(tc::split_tf32_kernel, (unsigned)((n4 + 255) / 256), 256, 0, st, x, hi, lo, n4);
  return check_launch("split_tf32");
}

int to_bf16(const float* x, void* out, size_t n, cudaStream_t st) {
  if (!x || !out) { set_last_error("to_bf16: null pointer"); return REGCN_ERR_NULL; }
  if (n & 7) { set_last_error("to_bf16: element count must be a multiple of 8"); return REGCN_ERR_DIM; }
  if (!n) return REGCN_OK;
  const size_t n8 = n / 8;
  launch_k(tc::to_bf16_kernel, (unsigned)((n8 + 255) / 256), 256, 0, st, x, (uint16_t*)out, n8);
  return check_launch("to_bf16");
}

static int g_force_block_n = 0, g_force_stages = 0;
// REGCN_A32_TMA=0: fp32-A tiles always through the register-staged converter (A/B switch for profiles/)
// REGCN_GEMM_KBLK16=0: k-blocks of 32 floats everywhere (A/B switch)
static const bool g_kblk16 = [] { const char* e = getenv("REGCN_GEMM_KBLK16"); return !(e && e[0] == '0'); }();
static const bool g_a32_tma = [] { const char* e = getenv("REGCN_A32_TMA"); return !(e && e[0] == '0'); }();
static int g_hyp_poly = 1;      // 0: the counting epilogue evaluates the IEEE score of every candidate (yardstick / tests)
void score_count_poly(int on) { g_hyp_poly = on ? 1 : 0; }
void gemm_tf32_tune(int block_n, int stages) { g_force_block_n = block_n; g_force_stages = stages; }
// In-kernel timeline: launch i of an attached session stamps into record (i % capacity) of the device buffer; the host
// keeps what each recorded launch computed (epilogue kind, shape, grid, algorithmic flops).
struct TraceRec { int epi, M, N, K, grid, passes; double flops; };
static unsigned long long* g_trace = nullptr;
static size_t g_trace_cap = 0;
static std::vector<TraceRec> g_trace_recs;
static std::mutex g_trace_mu;
constexpr int kTraceCtas = 148;
void gemm_tf32_trace(void* dev_buf) {                   // single-record session (profiles/gemm_trace.py)
  std::lock_guard<std::mutex> g(g_trace_mu);
  g_trace = (unsigned long long*)dev_buf; g_trace_cap = dev_buf ? 1 : 0; g_trace_recs.clear();
}
void gemm_tf32_trace_begin(void* dev_buf, size_t bytes) {
  std::lock_guard<std::mutex> g(g_trace_mu);
  g_trace = (unsigned long long*)dev_buf;
  g_trace_cap = dev_buf ? bytes / ((size_t)kTraceCtas * tc::kTraceSlots * sizeof(unsigned long long)) : 0;
  if (!g_trace_cap) g_trace = nullptr;
  g_trace_recs.clear();
}
int gemm_tf32_trace_count() { std::lock_guard<std::mutex> g(g_trace_mu); return (int)g_trace_recs.size(); }
int gemm_tf32_trace_read(int i, int* epi, int* M, int* N, int* K, int* grid, int* passes, double* flops) {
  std::lock_guard<std::mutex> g(g_trace_mu);
  if (i < 0 || i >= (int)g_trace_recs.size()) return REGCN_ERR_DIM;
  const TraceRec& r = g_trace_recs[i];
  if (epi) *epi = r.epi; if (M) *M = r.M; if (N) *N = r.N; if (K) *K = r.K; if (grid) *grid = r.grid;
  if (passes) *passes = r.passes; if (flops) *flops = r.flops;
  return REGCN_OK;
}
int gemm_tf32_trace_slots() { return tc::kTraceSlots; }

// N-tile selection.  Large problems (>= one wave of 128-row tiles): the widest tile with the least padding, i.e.
// the fewest operand re-reads.  Small problems (relation GRU, compact active-row GEMMs, pair scores): a tile costs
// ~12 us of pure latency however little it computes, so narrow the tile until the grid covers the machine --
// every CTA then walks the same K loop on a quarter of the operand bytes and of the MMA work.
// SM hint: the evolve engine runs the small relation / active-row GEMMs on a side stream next to a large GEMM that
// owns most of the machine; they then pick the narrowest tile whose grid still fits the SMs that are left (one wave)
static thread_local int g_sm_hint = 0;
void gemm_tf32_sm_hint(int sms) { g_sm_hint = sms; }
// Grid cap: the next persistent grids use at most this many CTAs (0 = the whole machine).  The evolve engine splits the SMs
// between its two streams with it when the all-entity GEMMs run many rounds (a 144-CTA persistent grid would otherwise hold
// every SM for its whole duration and the other stream's chain would run behind it, not beside it).
static thread_local int g_grid_cap = 0;
void gemm_tf32_grid_cap(int ctas) { g_grid_cap = ctas > 0 ? ctas : 0; }

static int pick_block_n(int M, int N, int K, int split_k) {
  if (g_force_block_n > 0) return g_force_block_n;
  if (N <= 64) return (N + 15) / 16 * 16;
  const int m_tiles = (M + tc::BLOCK_M - 1) / tc::BLOCK_M;
  if (g_sm_hint > 0) {
    const int opts[5] = {32, 64, 128, 208, 256};
    for (int i = 0; i < 5; ++i)
      if (m_tiles * ((N + opts[i] - 1) / opts[i]) <= g_sm_hint) return opts[i];
    return N <= 208 ? 208 : 256;
  }
  // Wide N (scoring against a large candidate table): the widest tile -- the padding of the last tile is noise next to
  // the A-operand re-reads of a narrow one (N = 1 000 000 is a multiple of 64: "least padding" picked 64-column tiles and
  // the counting GEMM ran 2.35x slower per output than at N = 125 000).  Narrow N: least padding among the wide tiles.
  int best = 256, best_waste = 1 << 30;
  if (N < 8 * 256) {
    const int cands[4] = {256, 208, 128, 64};
    for (int i = 0; i < 4; ++i) {
      const int bn = cands[i];
      const int waste = (N + bn - 1) / bn * bn - N;
      if (waste < best_waste) { best = bn; best_waste = waste; }
    }
  } else if ((N + 207) / 208 * 208 - N < ((N + 255) / 256 * 256 - N) / 4 && (N + 207) / 208 * 208 - N < N / 256) {
    best = 208;
  }
  const int kb_per_cta = ((K + tc::BLOCK_K - 1) / tc::BLOCK_K + (split_k > 1 ? split_k : 1) - 1) / (split_k > 1 ? split_k : 1);
  if (kb_per_cta > 16) return best;                 // long K loops amortise the per-tile latency: keep the wide tile
  if (m_tiles * ((N + best - 1) / best) >= 148) return best;
  const int small[3] = {128, 64, 32};
  for (int i = 0; i < 3; ++i) {
    const int bn = small[i];
    if (bn >= best) continue;
    best = bn;
    if (m_tiles * ((N + bn - 1) / bn) >= 120) break;
  }
  return best;
}

size_t gemm_tf32_workspace_bytes(int M, int N, int split_k) {
  return split_k > 1 ? (size_t)split_k * (size_t)M * (size_t)N * sizeof(float) : 0;
}

static void clear_epi(tc::Params& p) {
  p.epi = 0; p.tscore = nullptr; p.target = nullptr; p.raw_count = nullptr; p.col_offset = 0; p.hyp = 0; p.x2 = nullptr;
  p.y2 = nullptr; p.row_c = nullptr; p.col_bias = nullptr; p.hc = 0.f; p.hproj_max = 0.f; p.hyp_ymax = 0.f; p.scale_margin = nullptr; p.diag_out = nullptr;
  p.addend = nullptr; p.ld_add = 0; p.bias = nullptr; p.accumulate = 0; p.ws = nullptr; p.C = nullptr; p.ldc = 0;
  p.lse_max = nullptr; p.lse_sum = nullptr;
  p.lay_d = 0; p.lay_raw = nullptr; p.lay_hi = nullptr; p.lay_lo = nullptr; p.row_idx = nullptr; p.skip_rows = nullptr;
  p.gate_G = nullptr; p.gate_ld = 0; p.gate_bias = nullptr; p.gate_h = nullptr; p.gate_norm = 0;
  p.bf16 = 0; p.a_mn = 0; p.b_mn = 0;
  p.kblk = tc::BLOCK_K;
  p.a_f32 = 0; p.a_ptr[0] = p.a_ptr[1] = nullptr; p.a_ld[0] = p.a_ld[1] = 0; p.a_k[0] = p.a_k[1] = 0;
  p.a_rows[0] = p.a_rows[1] = nullptr; p.a_kb0 = 0;
  p.cv_x0 = p.cv_x1 = nullptr; p.cv_triples = nullptr; p.cv_col0 = p.cv_col1 = 0; p.cv_bn0_s = p.cv_bn0_t = nullptr;
  p.cv_w = p.cv_b = p.cv_bn1_s = p.cv_bn1_t = nullptr; p.cv_C = p.cv_d = p.cv_zb = 0;
  p.trace = nullptr;
}

// fp32-A operand description (see the header comment): up to two K segments, rows optionally gathered.
static int set_a_f32(tc::Params& p, const float* a0, int lda0, int k0, const int* rows0, const float* a1, int lda1, int k1,
                     const int* rows1, const char* who) {
  if (!a0 || k0 <= 0 || (k0 & 3) || (lda0 & 3) || lda0 < k0 || ((uintptr_t)a0 & 15) ||
      (k1 > 0 && (!a1 || (k1 & 3) || (lda1 & 3) || lda1 < k1 || ((uintptr_t)a1 & 15))) || k1 < 0) {
    set_last_error("%s: bad fp32 A operand (k0=%d lda0=%d k1=%d lda1=%d; 16-byte alignment, multiples of 4)", who, k0, lda0, k1, lda1);
    return REGCN_ERR_DIM;
  }
  p.a_f32 = 1;
  p.a_ptr[0] = a0; p.a_ld[0] = lda0; p.a_k[0] = k0; p.a_rows[0] = rows0;
  p.a_ptr[1] = k1 > 0 ? a1 : nullptr; p.a_ld[1] = lda1; p.a_k[1] = k1; p.a_rows[1] = rows1;
  p.a_kb0 = (k0 + tc::BLOCK_K - 1) / tc::BLOCK_K;
  return REGCN_OK;
}

// Common launcher: validates operands, builds the tensor maps, sizes the pipeline, launches.
static int launch_tc(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb,
                     tc::Params p, int passes, int split_k, int force_block_n, const char* who, cudaStream_t st) {
  using namespace tc;
  const int M = p.M, N = p.N;
  const int Ktrue = p.K;                 // reduction length of B (the weights); fp32-A mode pads every segment to k-blocks
  const bool conv = p.a_f32 == 3;          // conv-producer mode: the converter warps compute the A tiles (see Params)
  if (conv) {
    if (p.epi != 0 || passes != 3) { set_last_error("%s: the conv-producer mode takes the store epilogue and 3 passes", who); return REGCN_ERR_UNSUPPORTED; }
    a_hi = a_lo = b_hi;
    lda = ldb;
  } else if (p.a_f32) {
    if (p.a_k[0] + p.a_k[1] != Ktrue) { set_last_error("%s: fp32 A segments %d + %d != K = %d", who, p.a_k[0], p.a_k[1], Ktrue); return REGCN_ERR_DIM; }
    if (p.bf16 || p.a_mn || p.b_mn) { set_last_error("%s: fp32 A takes K-major tf32 operands", who); return REGCN_ERR_UNSUPPORTED; }
    // one K segment, rows in place: TMA delivers the fp32 k-blocks, the converter warps split them inside the stage
    if ((g_a32_tma || p.epi == 5) && p.a_k[1] == 0 && !p.a_rows[0] && !(p.a_ld[0] & 3) && !((uintptr_t)p.a_ptr[0] & 15) && p.a_ld[0] >= p.a_k[0]) p.a_f32 = 2;
    if (p.epi == 5 && p.a_f32 != 2) { set_last_error("%s: the fused time gate takes fp32 A as one K segment with its rows in place", who); return REGCN_ERR_UNSUPPORTED; }
    a_hi = a_lo = b_hi;                  // placeholders for the checks / unused tensor maps below
    lda = ldb;
  }
  // tile width first: it decides how many ring slots fit, and with it the k-block width
  p.block_n = force_block_n > 0 ? force_block_n : pick_block_n(M, N, Ktrue, split_k);
  if (p.b_mn) {                                         // MN-major B tiles are whole 32-element atoms
    const int nt = (N + 255) / 256;
    p.block_n = ((N + nt - 1) / nt + 31) / 32 * 32;
  }
  p.kblk = BLOCK_K;
  {
    const uint32_t stage32 = (passes == 3 ? 2u : 1u) * (BLOCK_M * BLOCK_K * 4 + (uint32_t)p.block_n * BLOCK_K * 4);
    // k-blocks of 16 floats (64-byte swizzle rows) when fewer than four 128-byte-row stages fit: the ring is what hides the
    // L2 / HBM latency of the operand stream, and two slots of 86 KB do not (8.5 us per 128 x 208 tile against 4.6 us of MMA)
    if (g_kblk16 && !p.bf16 && !p.a_mn && !p.b_mn && p.a_f32 != 1 && SMEM_BUDGET / stage32 < 4) {
      p.kblk = 16;
      if (split_k > 1) {
        // split-K: the caller sized its workspace / reduction for `split_k` parts -- keep the narrow k-blocks only if the
        // work list they give has exactly that many
        const int tkb = (Ktrue + 15) / 16, per = (tkb + split_k - 1) / split_k;
        if ((tkb + per - 1) / per != split_k) p.kblk = BLOCK_K;
      }
    }
  }
  if (conv) p.kblk = 16;
  if (p.a_f32 == 2) p.K = (Ktrue + p.kblk - 1) / p.kblk * p.kblk;
  else if (p.a_f32 == 1) p.K = BLOCK_K * (p.a_kb0 + (p.a_k[1] + BLOCK_K - 1) / BLOCK_K);
  const int K = p.K;
  if (!a_hi || !b_hi || (passes == 3 && (!a_lo || !b_lo))) { set_last_error("%s: null operand", who); return REGCN_ERR_NULL; }
  if (passes != 1 && passes != 3) { set_last_error("%s: passes must be 1 or 3", who); return REGCN_ERR_DIM; }
  const int ld_mask = p.bf16 ? 7 : 3;                 // row pitch must be a multiple of 16 bytes
  if (p.bf16 && passes != 1) { set_last_error("%s: bf16 operands take one pass", who); return REGCN_ERR_DIM; }
  if (M < 0 || N <= 0 || K <= 0 || (lda & ld_mask) || (ldb & ld_mask) || (!p.a_f32 && lda < (p.a_mn ? M : K)) || ldb < (p.b_mn ? N : Ktrue) ||
      (((uintptr_t)a_hi | (uintptr_t)b_hi | (uintptr_t)a_lo | (uintptr_t)b_lo) & 15)) {
    set_last_error("%s: bad dims/alignment M=%d N=%d K=%d lda=%d ldb=%d", who, M, N, K, lda, ldb);
    return REGCN_ERR_DIM;
  }
  if (M == 0) return REGCN_OK;
  p.passes = passes;
  if (p.a_mn || p.b_mn) {
    if (p.bf16 || p.epi != 0) { set_last_error("%s: MN-major operands take fp32 data and the store epilogue", who); return REGCN_ERR_UNSUPPORTED; }
  }
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.block_n) p.tmem_cols <<= 1;     // two accumulator slots
  const uint32_t stage_bytes = (passes == 3 ? 2u : 1u) * (BLOCK_M * (uint32_t)p.kblk * 4 + (uint32_t)p.block_n * (uint32_t)p.kblk * 4);
  const int block_k = p.bf16 ? 2 * BLOCK_K : p.kblk;
  const int total_kb = conv ? p.cv_C * p.cv_zb : (K + block_k - 1) / block_k;
  if (split_k < 1) split_k = 1;
  if (split_k > total_kb) split_k = total_kb;
  p.kb_per_split = (total_kb + split_k - 1) / split_k;
  if (!conv) split_k = (total_kb + p.kb_per_split - 1) / p.kb_per_split;     // (conv mode: equal shares of the units, any count)
  const int ew = p.epi == 5 ? epi_warps<5>() : EPI_WARPS;                 // epilogue warps of the instantiation that will run
  p.stages = (int)((ring_budget(ew) - (conv ? kConvBytes : 0u)) / stage_bytes);
  if (p.stages > 8) p.stages = 8;
  if (g_force_stages > 0 && g_force_stages < p.stages) p.stages = g_force_stages;
  if (p.stages < (p.a_f32 ? 2 : 1)) { set_last_error("%s: tile does not fit in shared memory", who); return REGCN_ERR_UNSUPPORTED; }
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int e;
  // K-major operand (rows, K): box of 32 k-floats x tile rows; MN-major operand (K, rows): boxes of 32 columns x 32 rows
  if ((e = p.b_mn ? make_map(&tb_hi, b_hi, K, N, ldb, BLOCK_K, false, true) : make_map(&tb_hi, b_hi, N, Ktrue, ldb, p.block_n, p.bf16 != 0, false, p.kblk))) return e;
  if (p.a_f32 == 2) {
    if ((e = make_map(&ta_hi, p.a_ptr[0], M, p.a_k[0], p.a_ld[0], BLOCK_M, false, false, p.kblk))) return e;
  } else if (p.a_f32) ta_hi = tb_hi;    // never dereferenced: the converter warps build the A tiles
  else if ((e = p.a_mn ? make_map(&ta_hi, a_hi, K, M, lda, BLOCK_K, false, true) : make_map(&ta_hi, a_hi, M, K, lda, BLOCK_M, p.bf16 != 0, false, p.kblk))) return e;
  if (passes == 3) {
    if (p.a_f32) ta_lo = tb_hi;
    else if ((e = p.a_mn ? make_map(&ta_lo, a_lo, K, M, lda, BLOCK_K, false, true) : make_map(&ta_lo, a_lo, M, K, lda, BLOCK_M, false, false, p.kblk))) return e;
    if ((e = p.b_mn ? make_map(&tb_lo, b_lo, K, N, ldb, BLOCK_K, false, true) : make_map(&tb_lo, b_lo, N, Ktrue, ldb, p.block_n, false, false, p.kblk))) return e;
  } else {
    ta_lo = ta_hi; tb_lo = tb_hi;
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + staging_bytes(ew) + (conv ? kConvBytes : 0u);
  static bool attr_set = false;
  if (!attr_set) {
    const int mx = (int)(SMEM_BUDGET + 1024 + STAGING_BYTES);   // + 2 KB of static shared memory = the 227 KB limit
    cudaError_t ce = cudaFuncSetAttribute(gemm_tf32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(gemm_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(gemm_tf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(gemm_tf32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(gemm_tf32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(gemm_tf32_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)(ring_budget(epi_warps<5>()) + 1024 + staging_bytes(epi_warps<5>())));
    if (ce != cudaSuccess) { set_last_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = p.epi == 2 ? 1 : (N + p.block_n - 1) / p.block_n;
  p.splits = split_k;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const long long total_tiles = (long long)p.m_tiles * p.n_tiles * p.splits;
  // balanced persistent grid: the smallest grid that needs no more rounds than the whole machine would (360 tiles on
  // 148 SMs take 3 rounds; so do 120 CTAs -- and 28 SMs stay free for whatever runs next to this kernel)
  const int sms_eff = (g_grid_cap > 0 && g_grid_cap < sms) ? g_grid_cap : sms;
  const long long rounds = (total_tiles + sms_eff - 1) / sms_eff;
  dim3 grid((unsigned)((total_tiles + rounds - 1) / (rounds > 0 ? rounds : 1)));
  const int Kalg = conv ? p.cv_C * p.cv_d : Ktrue;        // (the packed weight of the conv mode pads every row to 16 positions)
  const double alg_flops = p.epi == 2 ? 2.0 * M * (double)Kalg : 2.0 * M * (double)N * Kalg;
  if (g_trace) {
    std::lock_guard<std::mutex> g(g_trace_mu);
    if (g_trace && grid.x <= (unsigned)kTraceCtas) {
      const size_t rec = g_trace_recs.size() % g_trace_cap;
      p.trace = g_trace + rec * (size_t)kTraceCtas * kTraceSlots;
      if (g_trace_recs.size() < g_trace_cap || g_trace_cap == 1) {
        if (g_trace_cap == 1) g_trace_recs.clear();
        g_trace_recs.push_back(TraceRec{p.epi, M, N, Kalg, (int)grid.x, passes, alg_flops});
      } else {
        p.trace = nullptr;                               // session full: later launches are not recorded
      }
    }
  }
  prof_begin(PROF_GEMM_TC, st);
  switch (p.epi) {
    case 0: launch_k(gemm_tf32_kernel<0>, grid, NUM_THREADS, smem, st, ta_hi, ta_lo, tb_hi, tb_lo, p); break;
    case 1: launch_k(gemm_tf32_kernel<1>, grid, NUM_THREADS, smem, st, ta_hi, ta_lo, tb_hi, tb_lo, p); break;
    case 2: launch_k(gemm_tf32_kernel<2>, grid, NUM_THREADS, smem, st, ta_hi, ta_lo, tb_hi, tb_lo, p); break;
    case 3: launch_k(gemm_tf32_kernel<3>, grid, NUM_THREADS, smem, st, ta_hi, ta_lo, tb_hi, tb_lo, p); break;
    case 5: launch_k(gemm_tf32_kernel<5>, grid, num_threads<5>(), smem, st, ta_hi, ta_lo, tb_hi, tb_lo, p); break;
    default: launch_k(gemm_tf32_kernel<4>, grid, NUM_THREADS, smem, st, ta_hi, ta_lo, tb_hi, tb_lo, p); break;
  }
  prof_end(PROF_GEMM_TC, alg_flops, st);
  return REGCN_OK;
}

// A_hi/A_lo [M,K] (lda), B_hi/B_lo [N,K] (ldb); passes==1 ignores the lo operands (may be NULL).
int gemm_tf32(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb, float* C,
              int ldc, int M, int N, int K, const float* bias, int accumulate, int passes, int split_k, float* ws,
              size_t ws_bytes, const float* addend, int ld_add, cudaStream_t st) {
  if (!C) { set_last_error("gemm_tf32: null output"); return REGCN_ERR_NULL; }
  if (ldc < N) { set_last_error("gemm_tf32: ldc=%d < N=%d", ldc, N); return REGCN_ERR_DIM; }
  if (addend && split_k > 1) { set_last_error("gemm_tf32: addend is not supported together with split-K"); return REGCN_ERR_UNSUPPORTED; }
  tc::Params p;
  clear_epi(p);
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.accumulate = accumulate;
  p.addend = addend; p.ld_add = ld_add;
  const int total_kb = (K + tc::BLOCK_K - 1) / tc::BLOCK_K;
  int sk = split_k < 1 ? 1 : (split_k > total_kb ? total_kb : split_k);
  if (sk > 1) {
    const int kb_per = (total_kb + sk - 1) / sk;
    sk = (total_kb + kb_per - 1) / kb_per;
  }
  if (sk > 1) {
    if (!ws || ws_bytes < gemm_tf32_workspace_bytes(M, N, sk)) { set_last_error("gemm_tf32: split-K workspace too small"); return REGCN_ERR_WORKSPACE; }
    p.ws = ws;
  }
  int e = launch_tc(a_hi, a_lo, lda, b_hi, b_lo, ldb, p, passes, sk, 0, "gemm_tf32", st);
  if (e) return e;
  if (sk > 1 && M > 0) {
    const size_t total = (size_t)M * N;
    launch_k(splitk_reduce_kernel, (unsigned)((total + 255) / 256), 256, 0, st, ws, sk, C, ldc, M, N, bias, accumulate);
  }
  return check_launch("gemm_tf32");
}

// C (M, N) = op(A) op(B) with either operand given in its natural row-major layout as an MN-major tcgen05 operand:
//   a_mn: A passed as X (K, M) (C = X^T ...), else A (M, K);  b_mn: B passed as Y (K, N) (C = ... Y), else B (N, K) (C = ... B^T)
int gemm_tf32_mn(const float* x_hi, const float* x_lo, int ldx, const float* y_hi, const float* y_lo, int ldy, float* C,
                 int ldc, int M, int N, int K, int a_mn, int b_mn, const float* bias, int accumulate, int passes,
                 int split_k, float* ws, size_t ws_bytes, cudaStream_t st) {
  if (!C) { set_last_error("gemm_tf32_mn: null output"); return REGCN_ERR_NULL; }
  if (ldc < N) { set_last_error("gemm_tf32_mn: ldc=%d < N=%d", ldc, N); return REGCN_ERR_DIM; }
  tc::Params p;
  clear_epi(p);
  p.a_mn = a_mn != 0; p.b_mn = b_mn != 0;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.accumulate = accumulate; p.bias = bias;
  const int total_kb = (K + tc::BLOCK_K - 1) / tc::BLOCK_K;
  int sk = split_k < 1 ? 1 : (split_k > total_kb ? total_kb : split_k);
  if (sk > 1) {
    const int kb_per = (total_kb + sk - 1) / sk;
    sk = (total_kb + kb_per - 1) / kb_per;
  }
  if (sk > 1) {
    if (!ws || ws_bytes < gemm_tf32_workspace_bytes(M, N, sk)) { set_last_error("gemm_tf32_mn: split-K workspace too small"); return REGCN_ERR_WORKSPACE; }
    p.ws = ws;
  }
  int e = launch_tc(x_hi, x_lo, ldx, y_hi, y_lo, ldy, p, passes, sk, 0, "gemm_tf32_mn", st);
  if (e) return e;
  if (sk > 1 && M > 0) {
    const size_t total = (size_t)M * N;
    launch_k(splitk_reduce_kernel, (unsigned)((total + 255) / 256), 256, 0, st, ws, sk, C, ldc, M, N, bias, accumulate);
  }
  return check_launch("gemm_tf32_mn");
}

// GEMM with the fused layer epilogue (see Params): out rows = rrelu(A . B^T)[:, :d] (+ time gate), gate columns -> C.
int gemm_tf32_layer(const float* a_hi, const float* a_lo, int lda, const float* b_hi, const float* b_lo, int ldb, int M,
                    int N, int K, int d, float* out_raw, float* out_hi, float* out_lo, float* gate_out, int ld_gate_out,
                    const int* row_idx, const int* skip_rows, const float* gate_G, int gate_ld, const float* gate_bias,
                    const float* gate_h, int gate_norm, cudaStream_t st) {
  if ((!out_raw && !out_hi) || (out_hi && !out_lo)) { set_last_error("gemm_tf32_layer: no output"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || (N & 3) || N < d || (N > d && (!gate_out || ld_gate_out < N - d || (ld_gate_out & 3)))) {
    set_last_error("gemm_tf32_layer: bad dims N=%d d=%d", N, d); return REGCN_ERR_DIM;
  }
  if (gate_G && (N != d || d > 256 || !gate_bias || !gate_h || (gate_ld & 3) || !out_raw)) {
    set_last_error("gemm_tf32_layer: the fused time gate needs N == d <= 256 and an fp32 output"); return REGCN_ERR_DIM;
  }
  tc::Params p;
  clear_epi(p);
  p.M = M; p.N = N; p.K = K; p.epi = 3; p.lay_d = d; p.lay_raw = out_raw; p.lay_hi = out_hi; p.lay_lo = out_lo;
  p.C = gate_out; p.ldc = ld_gate_out; p.row_idx = row_idx; p.skip_rows = skip_rows;
  p.gate_G = gate_G; p.gate_ld = gate_ld; p.gate_bias = gate_bias; p.gate_h = gate_h; p.gate_norm = gate_norm;
  const int force_bn = gate_G ? (d + 15) / 16 * 16 : 0;      // one tile must span the whole row for the norm
  if (gate_G) p.epi = 5;                                     // the time-gate epilogue
  int e = launch_tc(a_hi, a_lo, lda, b_hi, b_lo, ldb, p, 3, 1, force_bn, "gemm_tf32_layer", st);
  if (e) return e;
  return check_launch("gemm_tf32_layer");
}

// ---- fp32-A variants: the A operand is one fp32 copy (two K segments, optional row gather), split to TF32 on chip ----
int gemm_tf32_a32(const float* a0, int lda0, int k0, const int* rows0, const float* a1, int lda1, int k1, const int* rows1,
                  const float* b_hi, const float* b_lo, int ldb, float* C, int ldc, int M, int N, const float* bias,
                  int accumulate, int passes, int split_k, float* ws, size_t ws_bytes, const float* addend, int ld_add,
                  cudaStream_t st) {
  if (!C) { set_last_error("gemm_tf32_a32: null output"); return REGCN_ERR_NULL; }
  if (ldc < N) { set_last_error("gemm_tf32_a32: ldc=%d < N=%d", ldc, N); return REGCN_ERR_DIM; }
  if (addend && split_k > 1) { set_last_error("gemm_tf32_a32: addend is not supported together with split-K"); return REGCN_ERR_UNSUPPORTED; }
  tc::Params p;
  clear_epi(p);
  int e = set_a_f32(p, a0, lda0, k0, rows0, a1, lda1, k1, rows1, "gemm_tf32_a32");
  if (e) return e;
  const int K = k0 + (k1 > 0 ? k1 : 0);
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.accumulate = accumulate;
  p.addend = addend; p.ld_add = ld_add;
  const int total_kb = p.a_kb0 + (p.a_k[1] + tc::BLOCK_K - 1) / tc::BLOCK_K;
  int sk = split_k < 1 ? 1 : (split_k > total_kb ? total_kb : split_k);
  if (sk > 1) {
    const int kb_per = (total_kb + sk - 1) / sk;
    sk = (total_kb + kb_per - 1) / kb_per;
  }
  if (sk > 1) {
    if (!ws || ws_bytes < gemm_tf32_workspace_bytes(M, N, sk)) { set_last_error("gemm_tf32_a32: split-K workspace too small"); return REGCN_ERR_WORKSPACE; }
    p.ws = ws;
  }
  e = launch_tc(nullptr, nullptr, 0, b_hi, b_lo, ldb, p, passes, sk, 0, "gemm_tf32_a32", st);
  if (e) return e;
  if (sk > 1 && M > 0) {
    const size_t total = (size_t)M * N;
    launch_k(splitk_reduce_kernel, (unsigned)((total + 255) / 256), 256, 0, st, ws, sk, C, ldc, M, N, bias, accumulate);
  }
  return check_launch("gemm_tf32_a32");
}

// ---- ConvTransE / ConvTransR query tower up to the fully-connected layer in ONE GEMM (src/decoder.py:78-93):
//   out[b, :] = W_fc . vec(relu(bn1(conv1d_k3(bn0([x0[t[b][col0]]; x1[t[b][col1]]]))))) + bias
// The (B, C d) feature map is computed inside the operand ring (Params: conv-producer mode) instead of being written
// and read back (117 MB per tower at the ICEWS18 shape).  Split-K over equal shares of the (position block, channel)
// units; partials are folded in split order by splitk_reduce_kernel (deterministic).
// fc.weight (N, C d) -> the unit order of the conv-producer GEMM, split to TF32:
//   out[n, 16 (z C + c) + j] = W[n, c d + 16 z + j]   (zero where 16 z + j >= d),   row length zb C 16, zb = ceil(d / 16)
__global__ void convfc_pack_weight_kernel(const float* __restrict__ w, int N, int C, int d, int zb, float* __restrict__ hi,
                                          float* __restrict__ lo) {
  pdl_grid_sync();
  const size_t total = (size_t)N * zb * C * 16;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i & 15);
  const size_t u = i >> 4;
  const int c = (int)(u % C);
  const size_t u2 = u / C;
  const int z = (int)(u2 % zb);
  const size_t n = u2 / zb;
  const int pos = 16 * z + j;
  const float v = pos < d ? w[n * (size_t)C * d + (size_t)c * d + pos] : 0.f;
  float h, l;
  split_tf32_1(v, h, l);
  hi[i] = h;
  lo[i] = l;
}
int convfc_pack_weight(const float* w, int N, int C, int d, float* hi, float* lo, cudaStream_t st) {
  if (!w || !hi || !lo) { set_last_error("convfc_pack_weight: null pointer"); return REGCN_ERR_NULL; }
  if (N <= 0 || C <= 0 || d <= 0) { set_last_error("convfc_pack_weight: bad dims"); return REGCN_ERR_DIM; }
  const size_t total = (size_t)N * ((d + 15) / 16) * C * 16;
  launch_k(convfc_pack_weight_kernel, (unsigned)((total + 255) / 256), 256, 0, st, w, N, C, d, (d + 15) / 16, hi, lo);
  return check_launch("convfc_pack_weight");
}
int affine_relu(float* x, const float* scale, const float* shift, int M, int d, int relu, cudaStream_t st);   // decoder.cu
// split-K fold of the fused tower with its tail: out = [relu]([scale *] (sum_z ws[z] + bias) [+ shift]), optional TF32 split
// of the result for the scoring GEMM (one pass instead of splitk_reduce + affine_relu + split_tf32); split order fixed
__global__ void convfc_finish_kernel(const float* __restrict__ ws, int splits, int M, int N, const float* __restrict__ bias,
                                     const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                     float* __restrict__ out, int ldc, float* __restrict__ out_hi, float* __restrict__ out_lo) {
  pdl_grid_sync();
  const size_t total4 = (size_t)M * N / 4;
  const size_t i4 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const size_t i = i4 * 4;
  const int col = (int)(i % N);
  const size_t row = i / N;
  float4 a = *reinterpret_cast<const float4*>(ws + i);
  for (int z = 1; z < splits; ++z) a = f4_add(a, *reinterpret_cast<const float4*>(ws + (size_t)z * M * N + i));
  if (bias) a = f4_add(a, ldg4(bias + col));
  if (scale) {
    const float4 s4 = ldg4(scale + col), t4 = ldg4(shift + col);
    a.x = fmaf(a.x, s4.x, t4.x); a.y = fmaf(a.y, s4.y, t4.y); a.z = fmaf(a.z, s4.z, t4.z); a.w = fmaf(a.w, s4.w, t4.w);
  }
  if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
  st4(out + row * ldc + col, a);
  if (out_hi) {
    float4 h, l;
    split_tf32_1(a.x, h.x, l.x); split_tf32_1(a.y, h.y, l.y); split_tf32_1(a.z, h.z, l.z); split_tf32_1(a.w, h.w, l.w);
    st4(out_hi + i, h);
    st4(out_lo + i, l);
  }
}
int convtrans_fc_splits(int B) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int m_tiles = (B + tc::BLOCK_M - 1) / tc::BLOCK_M;
  int s = sms / (m_tiles > 0 ? m_tiles : 1);
  return s < 1 ? 1 : (s > 16 ? 16 : s);
}
size_t convtrans_fc_workspace_bytes(int B, int N) { return gemm_tf32_workspace_bytes(B, N, convtrans_fc_splits(B)) + 256; }
int convtrans_fc(const float* x0, const float* x1, const int64_t* triples, int col0, int col1, int B, int d, int C, int ksz,
                 const float* bn0_scale, const float* bn0_shift, const float* conv_w, const float* conv_b,
                 const float* bn1_scale, const float* bn1_shift, const float* w_hi, const float* w_lo, int ldw, int N,
                 const float* bias, float* out, int ldc, float* ws, size_t ws_bytes, cudaStream_t st, int batch_total,
                 const float* act_scale, const float* act_shift, int relu, float* out_hi, float* out_lo) {
  if ((act_scale && !act_shift) || (out_hi && !out_lo) || (ldc & 3)) { set_last_error("convtrans_fc: bad tail arguments"); return REGCN_ERR_NULL; }
  if (!x0 || !x1 || !triples || !bn0_scale || !bn0_shift || !conv_w || !conv_b || !bn1_scale || !bn1_shift || !w_hi || !w_lo || !out) {
    set_last_error("convtrans_fc: null pointer"); return REGCN_ERR_NULL;
  }
  if (B <= 0) return REGCN_OK;
  const int zb = (d + 15) / 16;
  if (ksz != 3 || d <= 0 || (d & 3) || C <= 0 || C > tc::kConvMaxC || N <= 0 || (N & 3) || ldc < N || ldw < zb * C * 16 ||
      (((uintptr_t)x0 | (uintptr_t)x1) & 15) || col0 < 0 || col0 > 2 || col1 < 0 || col1 > 2) {
    set_last_error("convtrans_fc: unsupported shape d=%d C=%d k=%d N=%d (kernel 3, d %% 4 == 0, C <= %d)", d, C, ksz, N, tc::kConvMaxC);
    return REGCN_ERR_UNSUPPORTED;
  }
  tc::Params p;
  clear_epi(p);
  p.a_f32 = 3;
  p.cv_x0 = x0; p.cv_x1 = x1; p.cv_triples = triples; p.cv_col0 = col0; p.cv_col1 = col1;
  p.cv_bn0_s = bn0_scale; p.cv_bn0_t = bn0_shift; p.cv_w = conv_w; p.cv_b = conv_b; p.cv_bn1_s = bn1_scale; p.cv_bn1_t = bn1_shift;
  p.cv_C = C; p.cv_d = d; p.cv_zb = (d + 15) / 16;
  p.C = out; p.ldc = ldc; p.M = B; p.N = N; p.K = zb * C * 16; p.bias = bias;
  int sk = convtrans_fc_splits(batch_total > B ? batch_total : B);   // a slice of a sharded batch splits like the whole batch
  if (sk > C * p.cv_zb) sk = C * p.cv_zb;
  if (sk > 1) {
    if (!ws || ws_bytes < gemm_tf32_workspace_bytes(B, N, sk)) { set_last_error("convtrans_fc: split-K workspace too small"); return REGCN_ERR_WORKSPACE; }
    p.ws = ws;
  }
  const int bn = N <= 256 ? (N + 15) / 16 * 16 : 256;
  int e = launch_tc(nullptr, nullptr, 0, w_hi, w_lo, ldw, p, 3, sk, bn, "convtrans_fc", st);
  if (e) return e;
  if (sk > 1) {
    const size_t total4 = (size_t)B * N / 4;
    launch_k(convfc_finish_kernel, (unsigned)((total4 + 255) / 256), 256, 0, st, (const float*)ws, sk, B, N, bias, act_scale, act_shift,
             relu, out, ldc, out_hi, out_lo);
    return check_launch("convtrans_fc");
  }
  // one slice: the GEMM wrote out (+ bias) itself
  if ((e = check_launch("convtrans_fc"))) return e;
  if (act_scale || relu) {
    if (ldc != N) { set_last_error("convtrans_fc: the tail needs a dense output when the GEMM runs unsplit"); return REGCN_ERR_UNSUPPORTED; }
    if ((e = affine_relu(out, act_scale, act_shift, B, N, relu, st))) return e;
  }
  if (out_hi) {
    if (ldc != N) { set_last_error("convtrans_fc: the tail needs a dense output when the GEMM runs unsplit"); return REGCN_ERR_UNSUPPORTED; }
    return split_tf32(out, out_hi, out_lo, (size_t)B * N, st);
  }
  return REGCN_OK;
}

int gemm_tf32_layer_a32(const float* a0, int lda0, int k0, const int* rows0, const float* a1, int lda1, int k1,
                        const int* rows1, const float* b_hi, const float* b_lo, int ldb, int M, int N, int d,
                        float* out_raw, float* out_hi, float* out_lo, float* gate_out, int ld_gate_out, const int* row_idx,
                        const int* skip_rows, const float* gate_G, int gate_ld, const float* gate_bias, const float* gate_h,
                        int gate_norm, cudaStream_t st) {
  if ((!out_raw && !out_hi) || (out_hi && !out_lo)) { set_last_error("gemm_tf32_layer_a32: no output"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || (N & 3) || N < d || (N > d && (!gate_out || ld_gate_out < N - d || (ld_gate_out & 3)))) {
    set_last_error("gemm_tf32_layer_a32: bad dims N=%d d=%d", N, d); return REGCN_ERR_DIM;
  }
  if (gate_G && (N != d || d > 256 || !gate_bias || !gate_h || (gate_ld & 3) || !out_raw)) {
    set_last_error("gemm_tf32_layer_a32: the fused time gate needs N == d <= 256 and an fp32 output"); return REGCN_ERR_DIM;
  }
  tc::Params p;
  clear_epi(p);
  int e = set_a_f32(p, a0, lda0, k0, rows0, a1, lda1, k1, rows1, "gemm_tf32_layer_a32");
  if (e) return e;
  p.M = M; p.N = N; p.K = k0 + (k1 > 0 ? k1 : 0); p.epi = 3; p.lay_d = d; p.lay_raw = out_raw; p.lay_hi = out_hi; p.lay_lo = out_lo;
  p.C = gate_out; p.ldc = ld_gate_out; p.row_idx = row_idx; p.skip_rows = skip_rows;
  p.gate_G = gate_G; p.gate_ld = gate_ld; p.gate_bias = gate_bias; p.gate_h = gate_h; p.gate_norm = gate_norm;
  const int force_bn = gate_G ? (d + 15) / 16 * 16 : 0;
  if (gate_G) p.epi = 5;
  e = launch_tc(nullptr, nullptr, 0, b_hi, b_lo, ldb, p, 3, 1, force_bn, "gemm_tf32_layer_a32", st);
  if (e) return e;
  return check_launch("gemm_tf32_layer_a32");
}

// Fused K11/K13 + K14: raw_count[b] += #{candidate n of this shard, n != target[b] : score(b,n) beats tscore[b]}.
// Candidates are rows of E (hi/lo, [N,K]); global id of row n is col_offset + n.  Scores are never written.
int score_count_tf32(const float* q_hi, const float* q_lo, const float* e_hi, const float* e_lo, int B, int N, int K,
                     const float* tscore, const int* target, int* raw_count, int col_offset, int hyp, const float* x2,
                     const float* y2, const float* col_bias, double c, const float* scale_margin, const float* row_c,
                     int passes, cudaStream_t st) {
  if (!tscore || !target || !raw_count || (hyp && (!x2 || !y2 || !scale_margin))) { set_last_error("score_count_tf32: null pointer"); return REGCN_ERR_NULL; }
  if (row_c && !hyp) { set_last_error("score_count_tf32: row_c needs hyp != 0"); return REGCN_ERR_DIM; }
  tc::Params p;
  clear_epi(p);
  p.M = B; p.N = N; p.K = K; p.epi = 1; p.tscore = tscore; p.target = target; p.raw_count = raw_count;
  p.col_offset = col_offset; p.hyp = hyp ? (row_c ? 2 : 1) : 0; p.x2 = x2; p.y2 = y2; p.col_bias = col_bias;
  p.scale_margin = scale_margin; p.row_c = row_c;
  if (hyp) {
    Curv cv = make_curv(c); p.hc = cv.c; p.hproj_max = cv.proj_max;
    // points of the ball satisfy |e|^2 <= proj_max^2; the polynomial threshold test (EPI 1) bounds its rounding over
    // that range and sends anything outside it to the exact score
    p.hyp_ymax = g_hyp_poly ? (float)((double)cv.proj_max * (double)cv.proj_max * 1.0001) : -1.f;
  }
  if (passes == 0) { p.bf16 = 1; passes = 1; }       // bf16 operands (q_hi / e_hi point at bf16 rows of K elements)
  int e = launch_tc(q_hi, q_lo, K, e_hi, e_lo, K, p, passes, 1, 0, "score_count_tf32", st);
  if (e) return e;
  return check_launch("score_count_tf32");
}

// Fused K11/K13 + loss head: per-row streaming log-sum-exp over all candidates (no (B,N) logits).  The N tile is fixed
// at 256 so that the caller can size the slice buffers: score_lse_num_parts(N) slices of B floats each.
constexpr int kLseBlockN = 256;
int score_lse_num_parts(int N) { return ((N + kLseBlockN - 1) / kLseBlockN) * (tc::EPI_WARPS / 4); }

int score_lse_tf32(const float* q_hi, const float* q_lo, const float* e_hi, const float* e_lo, int B, int N, int K, int hyp,
                   const float* x2, const float* y2, const float* col_bias, double c, const float* scale_margin,
                   const float* row_c, int passes, float* part_max, float* part_sum, cudaStream_t st) {
  if (!part_max || !part_sum || (hyp && (!x2 || !y2 || !scale_margin))) { set_last_error("score_lse_tf32: null pointer"); return REGCN_ERR_NULL; }
  if (row_c && !hyp) { set_last_error("score_lse_tf32: row_c needs hyp != 0"); return REGCN_ERR_DIM; }
  tc::Params p;
  clear_epi(p);
  p.M = B; p.N = N; p.K = K; p.epi = 4; p.hyp = hyp ? (row_c ? 2 : 1) : 0; p.x2 = x2; p.y2 = y2; p.col_bias = col_bias;
  p.scale_margin = scale_margin; p.row_c = row_c; p.lse_max = part_max; p.lse_sum = part_sum;
  if (hyp) { Curv cv = make_curv(c); p.hc = cv.c; p.hproj_max = cv.proj_max; }
  if (passes == 0) { p.bf16 = 1; passes = 1; }
  int e = launch_tc(q_hi, q_lo, K, e_hi, e_lo, K, p, passes, 1, kLseBlockN, "score_lse_tf32", st);
  if (e) return e;
  return check_launch("score_lse_tf32");
}

// ce[b] = logsumexp_n score(b,n) - tscore[b] from the slices of score_lse_tf32 (fixed slice order: deterministic)
__global__ void ce_from_lse_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum, int nparts,
                                   int B, const float* __restrict__ tscore, float* __restrict__ ce) {
  pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = -INFINITY;
  for (int t = 0; t < nparts; ++t) m = fmaxf(m, part_max[(size_t)t * B + b]);
  float s = 0.f;
  for (int t = 0; t < nparts; ++t) {
    const float pm = part_max[(size_t)t * B + b];
    if (pm > -INFINITY) s += part_sum[(size_t)t * B + b] * expf(pm - m);
  }
  ce[b] = (m + logf(s)) - tscore[b];
}
// loss = mean(ce): one CTA, fixed reduction tree (deterministic)
__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t / (float)n;
  }
}
int mean_f32(const float* x, int n, float* out, cudaStream_t st) {
  if (!x || !out || n <= 0) { set_last_error("mean_f32: bad arguments"); return REGCN_ERR_NULL; }
  launch_k(mean_kernel, 1, 1024, 0, st, x, n, out);
  return check_launch("mean_f32");
}
int ce_from_lse(const float* part_max, const float* part_sum, int nparts, int B, const float* tscore, float* ce,
                float* loss, cudaStream_t st) {
  if (!part_max || !part_sum || !tscore || !ce) { set_last_error("ce_from_lse: null pointer"); return REGCN_ERR_NULL; }
  if (B <= 0 || nparts <= 0) { set_last_error("ce_from_lse: empty batch"); return REGCN_ERR_DIM; }
  launch_k(ce_from_lse_kernel, (unsigned)((B + 127) / 128), 128, 0, st, part_max, part_sum, nparts, B, tscore, ce);
  if (loss) launch_k(mean_kernel, 1, 1024, 0, st, (const float*)ce, B, loss);
  return check_launch("ce_from_lse");
}

// Pair scores through the same MMA arithmetic as the scoring GEMM: out[p] = score(A'[p], B'[p]) with A', B' the
// gathered (hi, lo) operand rows of the P pairs; x2 / y2 / col_bias are gathered per pair as well.
int pair_scores_tf32(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int P, int K, int hyp,
                     const float* x2, const float* y2, const float* col_bias, double c, const float* scale_margin,
                     const float* row_c, float* out, int passes, cudaStream_t st) {
  if (!out || (hyp && (!x2 || !y2 || !scale_margin))) { set_last_error("pair_scores_tf32: null pointer"); return REGCN_ERR_NULL; }
  if (row_c && !hyp) { set_last_error("pair_scores_tf32: row_c needs hyp != 0"); return REGCN_ERR_DIM; }
  tc::Params p;
  clear_epi(p);
  p.M = P; p.N = P; p.K = K; p.epi = 2; p.hyp = hyp ? (row_c ? 2 : 1) : 0; p.x2 = x2; p.y2 = y2; p.col_bias = col_bias;
  p.scale_margin = scale_margin; p.diag_out = out; p.row_c = row_c;
  if (hyp) { Curv cv = make_curv(c); p.hc = cv.c; p.hproj_max = cv.proj_max; }
  if (passes == 0) { p.bf16 = 1; passes = 1; }
  int e = launch_tc(a_hi, a_lo, K, b_hi, b_lo, K, p, passes, 1, 128, "pair_scores_tf32", st);
  if (e) return e;
  return check_launch("pair_scores_tf32");
}

// out_hi/out_lo[p] = src_hi/src_lo[idx[p]]  (operand gather for the pair-score pass)
__global__ void gather_rows2_kernel(const float* __restrict__ src_hi, const float* __restrict__ src_lo,
                                    const int* __restrict__ idx, int P, int d, float* __restrict__ out_hi,
                                    float* __restrict__ out_lo) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= P) return;
  const size_t s = (size_t)__ldg(idx + row) * d, o = (size_t)row * d;
  for (int c = lane * 4; c < d; c += 128) {
    st4(out_hi + o + c, ldg4(src_hi + s + c));
    if (src_lo) st4(out_lo + o + c, ldg4(src_lo + s + c));
  }
}
__global__ void gather_scalars_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                      const int* __restrict__ ia, const int* __restrict__ ib, int P, float* __restrict__ oa,
                                      float* __restrict__ ob, float* __restrict__ oc) {
  pdl_grid_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  if (a) oa[p] = __ldg(a + __ldg(ia + p));
  if (b) ob[p] = __ldg(b + __ldg(ib + p));
  if (c) oc[p] = __ldg(c + __ldg(ib + p));
}

int gather_rows2(const float* src_hi, const float* src_lo, const int* idx, int P, int d, float* out_hi, float* out_lo,
                 cudaStream_t st) {
  if (!src_hi || !idx || !out_hi || (src_lo && !out_lo)) { set_last_error("gather_rows2: null pointer"); return REGCN_ERR_NULL; }
  if (d & 3) { set_last_error("gather_rows2: d %% 4 != 0"); return REGCN_ERR_DIM; }
  if (P <= 0) return REGCN_OK;
  launch_k(gather_rows2_kernel, (unsigned)(((size_t)P * 32 + 255) / 256), 256, 0, st, src_hi, src_lo, idx, P, d, out_hi, out_lo);
  return check_launch("gather_rows2");
}
int gather_scalars(const float* a, const float* b, const float* c, const int* ia, const int* ib, int P, float* oa,
                   float* ob, float* oc, cudaStream_t st) {
  if (P <= 0) return REGCN_OK;
  if (!ia || !ib) { set_last_error("gather_scalars: null index"); return REGCN_ERR_NULL; }
  launch_k(gather_scalars_kernel, (P + 255) / 256, 256, 0, st, a, b, c, ia, ib, P, oa, ob, oc);
  return check_launch("gather_scalars");
}

}  // namespace regcn
