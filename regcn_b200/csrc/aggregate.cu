// Memory-bound edge path: CSR-by-destination gather / segmented sum kernels.
//   K4  union aggregate      rgcn/layers.py:257-279 (msg_func/apply_func), hyperbolic_layers.py:222-240
//   K6  block-diag aggregate rgcn/layers.py:167-179, hyperbolic_layers.py:87-109
//   K7  Lorentz centroid     hyperbolic_layers.py:589-625, hyperbolic_ops.py:477-518,563-581
//   K2  relation mean-pool   src/rrgcn.py:161-166, hyperbolic_model.py:802-812
// One warp owns one (virtual) destination row: lanes cover the row's float4 chunks (coalesced
// 128-bit loads of whole 800-byte entity rows), edges of the row are walked in CSR order and
// summed in registers -- a warp-segmented reduction, no atomics, deterministic.
#include "common.cuh"

namespace regcn {

constexpr int kAggChunk = 32;  // must match graph_build.cu
static int g_agg_impl = 0;       // 0 auto, 1 register-staged, 2 bulk-copy per row, 3 streaming bulk-copy ring
void aggregate_tune(int impl) { g_agg_impl = impl; }

// ---------------------------------------------------------------------------
// K4: agg[v] = norm[v] * sum_{(u,r)->v} w_uv * (h[u] + rel[r]),  w_uv = exp(-gamma*|rad[u]-rad[v]|) or 1.
// The d x d neighbour transform is applied afterwards by the node GEMM (the message is linear in
// (h[u] + rel[r]), so aggregate-then-transform is exact up to fp32 re-association).
// ---------------------------------------------------------------------------
template <int RV, bool RADIUS>
__global__ void __launch_bounds__(256) union_aggregate_kernel(
    const float* __restrict__ h, const float* __restrict__ rel, const int* __restrict__ rowptr,
    const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted, const float* __restrict__ norm,
    const int* __restrict__ vptr, const int* __restrict__ sptr, const int* __restrict__ vrow_row, int nv,
    const float* __restrict__ radius, float gamma, int d, float* __restrict__ out, float* __restrict__ partial,
    float* __restrict__ out_hi, float* __restrict__ out_lo, const int* __restrict__ active_pos, int ldo,
    int* __restrict__ fold_count) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (w >= nv) return;
  const int nvec = d >> 2;
  const int row = __ldg(vrow_row + w);
  const int v0 = __ldg(vptr + row), v1 = __ldg(vptr + row + 1);
  const int k = w - v0;
  const int rbeg = __ldg(rowptr + row), rend = __ldg(rowptr + row + 1);
  const int beg = rbeg + k * kAggChunk;
  const int end = min(beg + kAggChunk, rend);
  float r_dst = 0.f;
  if (RADIUS) r_dst = __ldg(radius + row);

  WarpRow<RV> acc;
  acc.zero();
  for (int base = beg; base < end; base += kWarp) {
    const int e = base + lane;
    int s = 0, t = 0;
    float wgt = 1.f;
    if (e < end) {
      s = __ldg(src_sorted + e);
      t = __ldg(etype_sorted + e);
      if (RADIUS) wgt = expf(-gamma * fabsf(__ldg(radius + s) - r_dst));
    }
    const int cnt = min(kWarp, end - base);
    int j = 0;
    for (; j + 8 <= cnt; j += 8) {  // 8 edges in flight per lane: 16 independent 16-byte row loads
      int sj[8], tj[8];
      float wj[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        sj[u] = __shfl_sync(0xffffffffu, s, j + u);
        tj[u] = __shfl_sync(0xffffffffu, t, j + u);
        wj[u] = RADIUS ? __shfl_sync(0xffffffffu, wgt, j + u) : 1.f;
      }
      float4 hv[8][RV];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int i = 0; i < RV; ++i) {
          int c = lane + i * kWarp;
          hv[u][i] = c < nvec ? ldg4(h + (size_t)sj[u] * d + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {      // relation rows come from a cache-resident table: 4 at a time
        float4 rv[4][RV];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int i = 0; i < RV; ++i) {
            int c = lane + i * kWarp;
            rv[u][i] = c < nvec ? ldg4(rel + (size_t)tj[half * 4 + u] * d + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int i = 0; i < RV; ++i) {
            float4 m = f4_add(hv[half * 4 + u][i], rv[u][i]);
            acc.v[i] = RADIUS ? f4_fma(wj[half * 4 + u], m, acc.v[i]) : f4_add(acc.v[i], m);
          }
        }
      }
    }
    for (; j + 2 <= cnt; j += 2) {
      int sj[2], tj[2];
      float wj[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        sj[u] = __shfl_sync(0xffffffffu, s, j + u);
        tj[u] = __shfl_sync(0xffffffffu, t, j + u);
        wj[u] = RADIUS ? __shfl_sync(0xffffffffu, wgt, j + u) : 1.f;
      }
      float4 hv[2][RV], rv[2][RV];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int i = 0; i < RV; ++i) {
          int c = lane + i * kWarp;
          if (c < nvec) {
            hv[u][i] = ldg4(h + (size_t)sj[u] * d + 4 * c);
            rv[u][i] = ldg4(rel + (size_t)tj[u] * d + 4 * c);
          } else {
            hv[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            rv[u][i] = hv[u][i];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int i = 0; i < RV; ++i) {
          float4 m = f4_add(hv[u][i], rv[u][i]);
          acc.v[i] = RADIUS ? f4_fma(wj[u], m, acc.v[i]) : f4_add(acc.v[i], m);
        }
      }
    }
    for (; j < cnt; ++j) {
      int sj = __shfl_sync(0xffffffffu, s, j);
      int tj = __shfl_sync(0xffffffffu, t, j);
      float wj = RADIUS ? __shfl_sync(0xffffffffu, wgt, j) : 1.f;
#pragma unroll
      for (int i = 0; i < RV; ++i) {
        int c = lane + i * kWarp;
        if (c < nvec) {
          float4 m = f4_add(ldg4(h + (size_t)sj * d + 4 * c), ldg4(rel + (size_t)tj * d + 4 * c));
          acc.v[i] = RADIUS ? f4_fma(wj, m, acc.v[i]) : f4_add(acc.v[i], m);
        }
      }
    }
  }
  // compact mode (active_pos != NULL): output row = position among the active destinations, row stride ldo; the
  // row's own features are copied (TF32-split) into columns [d, 2d) so that [agg | h] is one K = 2d GEMM operand
  const size_t orow = active_pos ? (size_t)__ldg(active_pos + row) : (size_t)row;
  if (active_pos && k == 0 && out_hi && ldo >= 2 * d) {
    WarpRow<RV> self;
    self.load(h + (size_t)row * d, nvec, lane);
    self.store_split(out_hi + orow * ldo + d, out_lo + orow * ldo + d, nvec, lane);
  }
  if (v1 - v0 == 1) {
    acc.scale(__ldg(norm + row));
    if (out) acc.store(out + orow * ldo, nvec, lane);
    if (out_hi) acc.store_split(out_hi + orow * ldo, out_lo + orow * ldo, nvec, lane);
  } else {
    const int s0 = __ldg(sptr + row);
    acc.store(partial + (size_t)(s0 + k) * d, nvec, lane);
    if (fold_count) {
      // In-kernel fold of a row split into <= 32 chunks (the launcher only passes fold_count then): the chunk warp that
      // arrives LAST sums the partials in chunk order -- the arithmetic of aggregate_fixup_kernel at stride 1, so the
      // result does not depend on which warp that is -- and re-arms the counter for the next launch.
      const int nch = v1 - v0;
      __threadfence();
      int prev = 0;
      if (lane == 0) prev = atomicAdd(fold_count + s0, 1);
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (prev == nch - 1) {
        __threadfence();
        WarpRow<RV> sum, p[4];
        sum.zero();
        int j = 0;
        for (; j + 4 <= nch; j += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < RV; ++i) {
              const int c = lane + i * kWarp;
              p[u].v[i] = c < nvec ? __ldcg(reinterpret_cast<const float4*>(partial + (size_t)(s0 + j + u) * d) + c)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < RV; ++i) sum.v[i] = f4_add(sum.v[i], p[u].v[i]);
        }
        for (; j < nch; ++j) {
#pragma unroll
          for (int i = 0; i < RV; ++i) {
            const int c = lane + i * kWarp;
            if (c < nvec) sum.v[i] = f4_add(sum.v[i], __ldcg(reinterpret_cast<const float4*>(partial + (size_t)(s0 + j) * d) + c));
          }
        }
        sum.scale(__ldg(norm + row));
        if (out) sum.store(out + orow * ldo, nvec, lane);
        if (out_hi) sum.store_split(out_hi + orow * ldo, out_lo + orow * ldo, nvec, lane);
        if (lane == 0) fold_count[s0] = 0;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K4, bulk-copy variant for HBM-bound sizes.  Same math and outputs as union_aggregate_kernel, but the gathered
// entity rows travel global -> shared memory with cp.async.bulk (one 4*d-byte bulk copy per edge, issued by the
// lane that owns the edge, completion counted on a per-warp mbarrier), so the bytes in flight per SM are bounded
// by shared memory (16 warps x 16 rows x 800 B = 205 KB) instead of by the register file.  Persistent: warps
// stride over the virtual rows.
// ---------------------------------------------------------------------------
constexpr int kBulkWarps = 16;
constexpr int kBulkRows = 16;

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RV, bool RADIUS>
__global__ void __launch_bounds__(kBulkWarps * 32, 1) union_aggregate_bulk_kernel(
    const float* __restrict__ h, const float* __restrict__ rel, const int* __restrict__ rowptr,
    const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted, const float* __restrict__ norm,
    const int* __restrict__ vptr, const int* __restrict__ sptr, const int* __restrict__ vrow_row, int nv,
    const float* __restrict__ radius, float gamma, int d, float* __restrict__ out, float* __restrict__ partial,
    float* __restrict__ out_hi, float* __restrict__ out_lo, const int* __restrict__ active_pos, int ldo) {
  pdl_grid_sync();
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  __shared__ __align__(8) unsigned long long bars[kBulkWarps];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int nvec = d >> 2;
  const uint32_t row_bytes = (uint32_t)d * 4u;
  float* stage = reinterpret_cast<float*>(bulk_smem) + (size_t)wid * kBulkRows * d;
  const uint32_t stage_u32 = smem_addr_u32(stage);
  const uint32_t bar = smem_addr_u32(&bars[wid]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase = 0;
  const int total_warps = gridDim.x * kBulkWarps;
  for (int w = blockIdx.x * kBulkWarps + wid; w < nv; w += total_warps) {
    const int row = __ldg(vrow_row + w);
    const int v0 = __ldg(vptr + row), v1 = __ldg(vptr + row + 1);
    const int k = w - v0;
    const int rbeg = __ldg(rowptr + row), rend = __ldg(rowptr + row + 1);
    const int beg = rbeg + k * kAggChunk;
    const int end = min(beg + kAggChunk, rend);
    float r_dst = 0.f;
    if (RADIUS) r_dst = __ldg(radius + row);
    WarpRow<RV> acc;
    acc.zero();
    for (int base = beg; base < end; base += kBulkRows) {
      const int cnt = min(kBulkRows, end - base);
      int s = 0, t = 0;
      float wgt = 1.f;
      if (lane < cnt) {
        s = __ldg(src_sorted + base + lane);
        t = __ldg(etype_sorted + base + lane);
        if (RADIUS) wgt = expf(-gamma * fabsf(__ldg(radius + s) - r_dst));
      }
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)cnt * row_bytes) : "memory");
      }
      __syncwarp();
      if (lane < cnt) {
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(stage_u32 + (uint32_t)lane * row_bytes), "l"(h + (size_t)s * d), "r"(row_bytes), "r"(bar) : "memory");
      }
      // relation rows come from a cache-resident table: fetch them while the bulk copies fly
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      }
      phase ^= 1;
      for (int j = 0; j < cnt; ++j) {
        const int tj = __shfl_sync(0xffffffffu, t, j);
        const float wj = RADIUS ? __shfl_sync(0xffffffffu, wgt, j) : 1.f;
#pragma unroll
        for (int i = 0; i < RV; ++i) {
          const int c = lane + i * kWarp;
          if (c < nvec) {
            const float4 hv = *reinterpret_cast<const float4*>(stage + (size_t)j * d + 4 * c);
            const float4 m = f4_add(hv, ldg4(rel + (size_t)tj * d + 4 * c));
            acc.v[i] = RADIUS ? f4_fma(wj, m, acc.v[i]) : f4_add(acc.v[i], m);
          }
        }
      }
      __syncwarp();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows are dead before the next bulk writes
    }
    const size_t orow = active_pos ? (size_t)__ldg(active_pos + row) : (size_t)row;
    if (active_pos && k == 0 && out_hi && ldo >= 2 * d) {
      WarpRow<RV> self;
      self.load(h + (size_t)row * d, nvec, lane);
      self.store_split(out_hi + orow * ldo + d, out_lo + orow * ldo + d, nvec, lane);
    }
    if (v1 - v0 == 1) {
      acc.scale(__ldg(norm + row));
      if (out) acc.store(out + orow * ldo, nvec, lane);
      if (out_hi) acc.store_split(out_hi + orow * ldo, out_lo + orow * ldo, nvec, lane);
    } else {
      acc.store(partial + (size_t)(__ldg(sptr + row) + k) * d, nvec, lane);
    }
  }
}

// ---------------------------------------------------------------------------
// K4, streaming variant for HBM-bound sizes (impl 3).  The per-row variants above issue a batch of gathers, wait for
// it, and then walk the index chain of the NEXT row (vrow_row -> vptr/rowptr -> src ids) with nothing in flight:
// on a 10-edges-per-row graph the gathers are outstanding a third of the time (ncu: DRAM 42 % busy, L2 31 %,
// long-scoreboard stalls).  Here a warp owns BLOCKS of 32 consecutive virtual rows, whose edges are one contiguous
// CSR range, and treats them as an edge STREAM:
//   * the row metadata of a block is loaded 32 rows at a time (one lane per row), one block ahead;
//   * gathers are issued in windows of 8 edges (cp.async.bulk, one 4*d-byte copy per edge, mbarrier completion)
//     into a two-slot ring, so a window is always in flight while the previous one is being summed;
//   * relation rows of a window are fetched into registers before the wait;
//   * row boundaries are handled while consuming (flush = scale by the degree norm, store, zero).
// Same outputs (bit for bit: same summation order inside a virtual row) as the other variants.
// ---------------------------------------------------------------------------
constexpr int kStreamWarps = 16;
constexpr int kStreamWin = 8;

template <int RV, bool RADIUS>
__global__ void __launch_bounds__(kStreamWarps * 32, 1) union_aggregate_stream_kernel(
    const float* __restrict__ h, const float* __restrict__ rel, const int* __restrict__ rowptr,
    const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted, const float* __restrict__ norm,
    const int* __restrict__ vptr, const int* __restrict__ sptr, const int* __restrict__ vrow_row, int nv,
    const float* __restrict__ radius, float gamma, int d, float* __restrict__ out, float* __restrict__ partial,
    float* __restrict__ out_hi, float* __restrict__ out_lo, const int* __restrict__ active_pos, int ldo) {
  pdl_grid_sync();
  extern __shared__ __align__(128) unsigned char stream_smem[];
  __shared__ __align__(8) unsigned long long bars[kStreamWarps][2];
  __shared__ int meta_t[kStreamWarps][2][kStreamWin];
  __shared__ float meta_rs[kStreamWarps][2][kStreamWin];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int nvec = d >> 2;
  const uint32_t row_bytes = (uint32_t)d * 4u;
  float* stage = reinterpret_cast<float*>(stream_smem) + (size_t)wid * 2 * kStreamWin * d;
  const uint32_t stage_u32 = smem_addr_u32(stage);
  const uint32_t bar0 = smem_addr_u32(&bars[wid][0]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int G = gridDim.x * kStreamWarps;                 // warps in the grid
  const int nblk = (nv + 31) >> 5;
  // per-lane metadata of one block: lane i describes virtual row 32*blk + i
  struct Meta { int row, beg, end, part, k0; float nrm, rdst; int orow; };
  auto load_meta = [&](int blk) {
    Meta m;
    m.row = 0; m.beg = 0; m.end = 0; m.part = -1; m.k0 = 0; m.nrm = 0.f; m.rdst = 0.f; m.orow = 0;
    const int w = blk * 32 + lane;
    if (blk < nblk && w < nv) {
      const int row = __ldg(vrow_row + w);
      const int v0 = __ldg(vptr + row), v1 = __ldg(vptr + row + 1);
      const int rb = __ldg(rowptr + row), re = __ldg(rowptr + row + 1);
      const int k = w - v0;
      m.row = row;
      m.beg = rb + k * kAggChunk;
      m.end = min(m.beg + kAggChunk, re);
      m.part = (v1 - v0 == 1) ? -1 : __ldg(sptr + row) + k;
      m.k0 = k == 0;
      m.nrm = __ldg(norm + row);
      m.orow = active_pos ? __ldg(active_pos + row) : row;
      if (RADIUS) m.rdst = __ldg(radius + row);
    }
    return m;
  };
  auto block_rows = [&](int blk) { return min(32, nv - blk * 32); };

  int cblk = blockIdx.x * kStreamWarps + wid;             // consumer's block; blocks cblk, cblk + G, ... belong to this warp
  if (cblk >= nblk) return;
  Meta cur = load_meta(cblk);
  Meta nxt = load_meta(cblk + G);
  // producer cursor: edge range [pe, p_hi) of block pblk
  int pblk = cblk;
  int pe = __shfl_sync(0xffffffffu, cur.beg, 0);
  int p_hi = __shfl_sync(0xffffffffu, cur.end, block_rows(cblk) - 1);
  // consumer cursor
  int ce = pe, c_hi = p_hi;
  int cv = 0;                                             // current virtual row inside the consumer's block
  int cv_end = __shfl_sync(0xffffffffu, cur.end, 0);
  float cv_rdst = RADIUS ? __shfl_sync(0xffffffffu, cur.rdst, 0) : 0.f;
  uint32_t prod_j = 0, cons_j = 0;
  uint32_t phase_bits = 0;                                // bit s = phase of ring slot s

  auto produce = [&]() {
    // one window of the producer's block into ring slot prod_j & 1 (the slot was drained two windows ago)
    const int cnt = min(kStreamWin, p_hi - pe);
    const int slot = (int)(prod_j & 1u);
    const uint32_t bar = bar0 + 8u * (uint32_t)slot;
    int s = 0, t = 0;
    if (lane < cnt) {
      s = __ldg(src_sorted + pe + lane);
      t = __ldg(etype_sorted + pe + lane);
    }
    if (lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)cnt * row_bytes) : "memory");
    }
    __syncwarp();
    if (lane < cnt) {
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(stage_u32 + (uint32_t)(slot * kStreamWin + lane) * row_bytes), "l"(h + (size_t)s * d), "r"(row_bytes), "r"(bar) : "memory");
      meta_t[wid][slot][lane] = t;
      if (RADIUS) meta_rs[wid][slot][lane] = __ldg(radius + s);
    }
    __syncwarp();
    pe += cnt;
    ++prod_j;
    if (pe == p_hi) {                                     // producer moves on to this warp's next block
      pblk += G;
      if (pblk < nblk) {
        // the producer is never more than two windows (16 edges) ahead and every block but the very last holds
        // >= 32 edges, so the next block is always the one described by `nxt`
        pe = __shfl_sync(0xffffffffu, nxt.beg, 0);
        p_hi = __shfl_sync(0xffffffffu, nxt.end, block_rows(pblk) - 1);
      }
    }
  };

  WarpRow<RV> acc;
  acc.zero();
  auto flush = [&]() {
    // virtual row cv of the consumer's block is complete
    const int row = __shfl_sync(0xffffffffu, cur.row, cv);
    const int part = __shfl_sync(0xffffffffu, cur.part, cv);
    const int k0 = __shfl_sync(0xffffffffu, cur.k0, cv);
    const float nrm = __shfl_sync(0xffffffffu, cur.nrm, cv);
    const size_t orow = (size_t)__shfl_sync(0xffffffffu, cur.orow, cv);
    if (active_pos && k0 && out_hi && ldo >= 2 * d) {
      WarpRow<RV> self;
      self.load(h + (size_t)row * d, nvec, lane);
      self.store_split(out_hi + orow * ldo + d, out_lo + orow * ldo + d, nvec, lane);
    }
    if (part < 0) {
      acc.scale(nrm);
      if (out) acc.store(out + orow * ldo, nvec, lane);
      if (out_hi) acc.store_split(out_hi + orow * ldo, out_lo + orow * ldo, nvec, lane);
    } else {
      acc.store(partial + (size_t)part * d, nvec, lane);
    }
    acc.zero();
  };

  produce();
  while (true) {
    if (pblk < nblk) produce();                           // keep one window in flight behind the one being consumed
    // ---- consume window cons_j ----
    const int cnt = min(kStreamWin, c_hi - ce);
    const int slot = (int)(cons_j & 1u);
    const uint32_t bar = bar0 + 8u * (uint32_t)slot;
    // relation rows of the window (cache-resident table) travel while the bulk copies land
    float4 rv[kStreamWin][RV];
    float wj[kStreamWin];
#pragma unroll
    for (int i = 0; i < kStreamWin; ++i) {
      wj[i] = 1.f;
      if (i < cnt) {
        const int t = meta_t[wid][slot][i];
#pragma unroll
        for (int q = 0; q < RV; ++q) {
          const int c = lane + q * kWarp;
          rv[i][q] = c < nvec ? ldg4(rel + (size_t)t * d + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    {
      const uint32_t ph = (phase_bits >> slot) & 1u;
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(ph) : "memory");
      }
      phase_bits ^= 1u << slot;
    }
    const float* srow = stage + (size_t)slot * kStreamWin * d;
#pragma unroll
    for (int i = 0; i < kStreamWin; ++i) {
      if (i < cnt) {
        const int e = ce + i;
        if (e >= cv_end) {                                // first edge of the next virtual row
          flush();
          ++cv;
          cv_end = __shfl_sync(0xffffffffu, cur.end, cv);
          if (RADIUS) cv_rdst = __shfl_sync(0xffffffffu, cur.rdst, cv);
        }
        if (RADIUS) wj[i] = expf(-gamma * fabsf(meta_rs[wid][slot][i] - cv_rdst));
#pragma unroll
        for (int q = 0; q < RV; ++q) {
          const int c = lane + q * kWarp;
          if (c < nvec) {
            const float4 hv = *reinterpret_cast<const float4*>(srow + (size_t)i * d + 4 * c);
            const float4 m = f4_add(hv, rv[i][q]);
            acc.v[q] = RADIUS ? f4_fma(wj[i], m, acc.v[q]) : f4_add(acc.v[q], m);
          }
        }
      }
    }
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot's rows are dead before the next bulk writes
    ce += cnt;
    ++cons_j;
    if (ce == c_hi) {                                     // block finished: flush its last row, move to the next block
      flush();
      cblk += G;
      if (cblk >= nblk) break;
      cur = nxt;
      nxt = load_meta(cblk + G);
      ce = __shfl_sync(0xffffffffu, cur.beg, 0);
      c_hi = __shfl_sync(0xffffffffu, cur.end, block_rows(cblk) - 1);
      cv = 0;
      cv_end = __shfl_sync(0xffffffffu, cur.end, 0);
      if (RADIUS) cv_rdst = __shfl_sync(0xffffffffu, cur.rdst, 0);
    }
  }
}

// Rows that were split into several chunks: fold the chunk partials with a radix-32 tree, one level per launch.
// Level with stride s: the warp of chunk k (k % (32 s) == 0) sums partial[k], partial[k+s], ..., partial[k+31 s]
// (fixed order: deterministic) back into partial[k]; the level that covers the whole row applies the degree norm
// and writes the output row.  A hub with 350k in-edges (11k chunks) is folded in 3 levels instead of one serial walk.
template <int RV>
__global__ void __launch_bounds__(256) aggregate_fixup_kernel(
    const int* __restrict__ vptr, const int* __restrict__ sptr, const int* __restrict__ vrow_row,
    const float* __restrict__ norm, int nv, int d, float* __restrict__ partial, float* __restrict__ out,
    float* __restrict__ out_hi, float* __restrict__ out_lo, const int* __restrict__ active_pos, int ldo, int stride) {
  pdl_grid_sync();
  // Persistent grid: every lane tests one virtual row per sweep (three dependent index loads), the warp then folds the
  // few that have work at this level one after the other.  A launch over all ~10^6 virtual rows of a 10 M-edge graph with
  // a warp per row cost 155-215 us per level just to start and retire 128 k CTAs that had nothing to do.
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < nv; base += warps * 32) {
    const long long wl = base + lane;
    int row = 0, nch = 0, k = 0;
    bool work = false;
    if (wl < nv) {
      row = __ldg(vrow_row + wl);
      const int v0 = __ldg(vptr + row);
      nch = __ldg(vptr + row + 1) - v0;
      k = (int)wl - v0;
      work = nch > 1 && (k % (32 * stride)) == 0;
      const bool covers = (long long)32 * stride >= nch;
      if (work && !covers && stride > 1 && k + stride >= nch) work = false;   // a lone slot: nothing to fold at this level
      if (work && stride > 1 && (long long)stride >= nch) work = false;       // row already finished at a lower level
    }
    unsigned todo = __ballot_sync(0xffffffffu, work);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int r_ = __shfl_sync(0xffffffffu, row, src);
      const int n_ = __shfl_sync(0xffffffffu, nch, src);
      const int k_ = __shfl_sync(0xffffffffu, k, src);
      const bool covers_row = (long long)32 * stride >= n_;                   // only true for k == 0
      const int s0 = __ldg(sptr + r_);
      WarpRow<RV> acc, p[4];
      acc.zero();
      int j = 0;
      for (; j + 4 <= 32; j += 4) {
        if (k_ + (j + 3) * stride >= n_) break;
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u].load_plain(partial + (size_t)(s0 + k_ + (j + u) * stride) * d, nvec, lane);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < RV; ++i) acc.v[i] = f4_add(acc.v[i], p[u].v[i]);
      }
      for (; j < 32 && k_ + j * stride < n_; ++j) {
        p[0].load_plain(partial + (size_t)(s0 + k_ + j * stride) * d, nvec, lane);
#pragma unroll
        for (int i = 0; i < RV; ++i) acc.v[i] = f4_add(acc.v[i], p[0].v[i]);
      }
      if (covers_row) {
        acc.scale(__ldg(norm + r_));
        const size_t orow = active_pos ? (size_t)__ldg(active_pos + r_) : (size_t)r_;
        if (out) acc.store(out + orow * ldo, nvec, lane);
        if (out_hi) acc.store_split(out_hi + orow * ldo, out_lo + orow * ldo, nvec, lane);
      } else {
        acc.store(partial + (size_t)(s0 + k_) * d, nvec, lane);
      }
    }
  }
}

// Dense-output mode: rows without in-edges receive exact zeros (DGL zero fill).  Persistent grid: a lane tests one row per
// sweep, the warp writes the rows that need it (a thread per float4 of the whole table spent 0.2 ms at N = 1M finding out that
// nothing had to be written).
__global__ void __launch_bounds__(256) zero_inactive_rows_kernel(const int* __restrict__ rowptr, int N, int d,
                                                                 float* __restrict__ out, float* __restrict__ out_hi,
                                                                 float* __restrict__ out_lo) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < N; base += warps * 32) {
    const long long r = base + lane;
    const bool empty = r < N && __ldg(rowptr + r + 1) == __ldg(rowptr + r);
    unsigned todo = __ballot_sync(0xffffffffu, empty);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const size_t o = (size_t)(base + src) * d;
      for (int c = lane; c < nvec; c += 32) {
        if (out) reinterpret_cast<float4*>(out + o)[c] = z;
        if (out_hi) { reinterpret_cast<float4*>(out_hi + o)[c] = z; reinterpret_cast<float4*>(out_lo + o)[c] = z; }
      }
    }
  }
}

// grid of the two persistent index-sweep kernels above: enough warps to keep every SM busy, never more than the work
static unsigned sweep_grid(long long items) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const long long need = (items + 255) / 256;            // one lane per item, 8 warps per CTA
  const long long cap = (long long)sms * 8;
  return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}

int union_aggregate(const float* h, const float* rel, const int* rowptr, const int* src_sorted,
                    const int* etype_sorted, const float* norm, const int* vptr, const int* sptr,
                    const int* vrow_row, int nv, int nsplit, const float* radius, float gamma, int N, int d,
                    float* out, float* partial, float* out_hi, float* out_lo, const int* active_pos, int ldo,
                    int max_chunks, cudaStream_t st, int* fold_count) {
  if (!h || !rel || !rowptr || !src_sorted || !etype_sorted || !norm || !vptr || !sptr || !vrow_row ||
      (!out && !out_hi) || (out_hi && !out_lo)) {
    set_last_error("union_aggregate: null pointer"); return REGCN_ERR_NULL;
  }
  if (d <= 0 || (d & 3) || d > 256) { set_last_error("union_aggregate: d=%d unsupported (need d%%4==0, d<=256)", d); return REGCN_ERR_UNSUPPORTED; }
  if (nsplit > 0 && !partial) { set_last_error("union_aggregate: split rows need a partial buffer"); return REGCN_ERR_WORKSPACE; }
  const int TB = 256;
  if (ldo <= 0) ldo = d;
  prof_begin(PROF_AGGREGATE, st);
  if (!active_pos) {
    const size_t total = (size_t)N * (d >> 2);
    (void)total;
    launch_k(zero_inactive_rows_kernel, sweep_grid(N), TB, 0, st, rowptr, N, d, out, out_hi, out_lo);
  }
  if (nv <= 0) { prof_end(PROF_AGGREGATE, 0.0, st); return check_launch("union_aggregate"); }
  const unsigned grid = (unsigned)(((size_t)nv * 32 + TB - 1) / TB);
  const bool small = d <= 128;
  const int impl = g_agg_impl;
  const bool bulk = impl == 2;   // opt-in: measured within +-10% of the register-staged variant (profiles/README.md)
  // streaming variant: automatic for HBM-bound sizes (the gathered rows no longer fit in L2), opt-in otherwise
  // (hub-heavy graphs stay on the register variant: their hot source rows hit in L1, which bulk copies bypass --
  //  measured on B200, N = 1M, E = 10M: uniform endpoints 2.14 ms stream vs 2.89 ms registers, Zipf 3.58 vs 2.91)
  // split rows of at most 32 chunks (1024 in-edges) are folded by the last chunk warp of the register kernel itself when
  // the caller provides zeroed arrival counters (one per partial slot): no fix-up launch
  if (max_chunks <= 0 || max_chunks > nsplit) max_chunks = nsplit;
  int* fold = (fold_count && nsplit > 0 && max_chunks <= 32) ? fold_count : nullptr;
  const bool stream = impl == 3 || (impl == 0 && (size_t)nv >= 65536 && (size_t)N * d * 4 > (size_t)96 * 1024 * 1024 &&
                                    (size_t)nsplit * 8 < (size_t)nv);
  if (stream || bulk) fold = nullptr;          // only the register kernel folds in place
  if (stream) {
    const size_t smem = (size_t)kStreamWarps * 2 * kStreamWin * d * sizeof(float);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned sgrid = (unsigned)sms;
#define LAUNCH_STREAM(RVV, RAD)                                                                                        \
    do {                                                                                                               \
      static bool attr_done = false;                                                                                   \
      if (!attr_done) {                                                                                                \
        cudaFuncSetAttribute(union_aggregate_stream_kernel<RVV, RAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
        attr_done = true;                                                                                              \
      }                                                                                                                \
      launch_k(union_aggregate_stream_kernel<RVV, RAD>, sgrid, kStreamWarps * 32, smem, st,                            \
          h, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv, radius, gamma, d, out, partial,   \
          out_hi, out_lo, active_pos, ldo);                                                                            \
    } while (0)
    if (radius) { if (small) LAUNCH_STREAM(1, true); else LAUNCH_STREAM(2, true); }
    else { if (small) LAUNCH_STREAM(1, false); else LAUNCH_STREAM(2, false); }
#undef LAUNCH_STREAM
  } else if (bulk) {
    const size_t smem = (size_t)kBulkWarps * kBulkRows * d * sizeof(float);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned bgrid = (unsigned)sms;
#define LAUNCH_BULK(RVV, RAD)                                                                                          \
    do {                                                                                                               \
      static bool attr_done = false;                                                                                   \
      if (!attr_done) {                                                                                                \
        cudaFuncSetAttribute(union_aggregate_bulk_kernel<RVV, RAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
        attr_done = true;                                                                                              \
      }                                                                                                                \
      launch_k(union_aggregate_bulk_kernel<RVV, RAD>, bgrid, kBulkWarps * 32, smem, st,                                      \
          h, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv, radius, gamma, d, out, partial,   \
          out_hi, out_lo, active_pos, ldo);                                                                            \
    } while (0)
    if (radius) { if (small) LAUNCH_BULK(1, true); else LAUNCH_BULK(2, true); }
    else { if (small) LAUNCH_BULK(1, false); else LAUNCH_BULK(2, false); }
#undef LAUNCH_BULK
  } else if (radius) {
    if (small) launch_k(union_aggregate_kernel<1, true>, grid, TB, 0, st, h, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv, radius, gamma, d, out, partial, out_hi, out_lo, active_pos, ldo, fold);
    else launch_k(union_aggregate_kernel<2, true>, grid, TB, 0, st, h, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv, radius, gamma, d, out, partial, out_hi, out_lo, active_pos, ldo, fold);
  } else {
    if (small) launch_k(union_aggregate_kernel<1, false>, grid, TB, 0, st, h, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv, nullptr, 0.f, d, out, partial, out_hi, out_lo, active_pos, ldo, fold);
    else launch_k(union_aggregate_kernel<2, false>, grid, TB, 0, st, h, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv, nullptr, 0.f, d, out, partial, out_hi, out_lo, active_pos, ldo, fold);
  }
  if (nsplit > 0 && !fold) {
    // max_chunks = chunk count of the largest hub row (<= nsplit); one launch per radix-32 level
    for (long long stride = 1; stride < (long long)max_chunks; stride *= 32) {
      if (small) launch_k(aggregate_fixup_kernel<1>, sweep_grid(nv), TB, 0, st, vptr, sptr, vrow_row, norm, nv, d, partial, out, out_hi, out_lo, active_pos, ldo, (int)stride);
      else launch_k(aggregate_fixup_kernel<2>, sweep_grid(nv), TB, 0, st, vptr, sptr, vrow_row, norm, nv, d, partial, out, out_hi, out_lo, active_pos, ldo, (int)stride);
    }
  }
  prof_end(PROF_AGGREGATE, 0.0, st);   // bytes are filled in by the caller-side formula (needs E, R); see bench.py
  return check_launch("union_aggregate");
}

// ---------------------------------------------------------------------------
// K6: block-diagonal relation transform, agg[v] = norm[v] * sum_in blockdiag(W[type]) . h[src].
// W is (num_rels, nb*si*so); output column j = b*so+o reads inputs b*si .. b*si+si-1.
// One warp per destination row, lanes stride over output columns (coalesced).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) block_aggregate_kernel(
    const float* __restrict__ h, const float* __restrict__ W, const float* __restrict__ rel_add,
    const int* __restrict__ rowptr, const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted,
    const float* __restrict__ norm, int N, int d_in, int d_out, int nb, float* __restrict__ out,
    const float* __restrict__ radius, float gamma) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= N) return;
  const int si = d_in / nb, so = d_out / nb;
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  const float nrm = __ldg(norm + row);
  const float r_dst = radius ? __ldg(radius + row) : 0.f;
  for (int j0 = 0; j0 < d_out; j0 += kWarp) {
    const int j = j0 + lane;
    float acc = 0.f;
    if (j < d_out) {
      const int b = j / so, o = j - b * so;
      for (int e = beg; e < end; ++e) {
        const int s = __ldg(src_sorted + e), t = __ldg(etype_sorted + e);
        const float* hp = h + (size_t)s * d_in + b * si;
        const float* wp = W + (size_t)t * ((size_t)nb * si * so) + (size_t)b * si * so + o;
        float m = 0.f;
        for (int i = 0; i < si; ++i) m = fmaf(__ldg(hp + i), __ldg(wp + i * so), m);
        if (rel_add) m += __ldg(rel_add + (size_t)t * d_out + j);
        // radius-difference message weight exp(-gamma |r_src - r_dst|)   (hyperbolic_layers.py:100-103)
        if (radius) m *= expf(-gamma * fabsf(__ldg(radius + s) - r_dst));
        acc += m;
      }
      out[(size_t)row * d_out + j] = acc * nrm;
    }
  }
}

int block_aggregate(const float* h, const float* W, const int* rowptr, const int* src_sorted,
                    const int* etype_sorted, const float* norm, int N, int d_in, int d_out, int nb,
                    float* out, cudaStream_t st, const float* radius, float gamma) {
  if (!h || !W || !rowptr || !src_sorted || !etype_sorted || !norm || !out) { set_last_error("block_aggregate: null pointer"); return REGCN_ERR_NULL; }
  if (nb <= 0 || d_in % nb || d_out % nb) { set_last_error("block_aggregate: num_bases=%d must divide d_in=%d and d_out=%d", nb, d_in, d_out); return REGCN_ERR_UNSUPPORTED; }
  const int TB = 256;
  const unsigned grid = (unsigned)(((size_t)N * 32 + TB - 1) / TB);
  launch_k(block_aggregate_kernel, grid, TB, 0, st, h, W, nullptr, rowptr, src_sorted, etype_sorted, norm, N, d_in, d_out, nb, out,
           radius, gamma);
  return check_launch("block_aggregate");
}

// ---------------------------------------------------------------------------
// K7: Lorentz-centroid aggregation (lgcn encoder).  Per edge (tangent space input ht):
//   m  = blockdiag(W[type]) . ht[src] + rel[type]            hyperbolic_layers.py:593-606
//   p  = exp_0(m);  mL = to_lorentz(p) = [(1+c|p|^2)/(sqrt_c*D), 2p/D], D = max(1-c|p|^2, eps)   :609-610, ops:492-499
// Per node v with K = indeg(v) > 0:
//   w_i = norm_v / (K*norm_v + 1e-6)  (all equal), then lorentz_centroid: w <- w/(sum w + eps),
//   cbar = sum w_i mL_i ; cbar /= sqrt(max(-<cbar,cbar>_L * c, eps))                              :613-625, ops:576-581
//   out = clamp(log_0(to_poincare(cbar)), +-10)                                                   :670-672
// indeg 0 -> exact zero row (DGL zero fill; to_poincare(0)=0, log_0(0)=0).
// Requires si == so == 2 or generic; lanes own float4 chunks of the d-vector like K4.
// ---------------------------------------------------------------------------
// Message of one edge for the float4 chunks owned by this lane: m = blockdiag(W[type]) . x[src] + rel[type].
// SB == 2 (the reference's default: 100 bases of 2x2 at d = 200): a float4 chunk holds exactly two blocks, so the
// transform is chunk-local -- one float4 of x, two float4 of weights, no index arithmetic.  SB == 0: generic block size.
template <int RV, int SB>
__device__ __forceinline__ void lorentz_message(WarpRow<RV>& m, const float* __restrict__ hp, const float* __restrict__ wp,
                                                const float* __restrict__ rp, int nvec, int lane, int sb) {
#pragma unroll
  for (int i = 0; i < RV; ++i) {
    const int c = lane + i * kWarp;
    if (c < nvec) {
      float4 r4 = rp ? ldg4(rp + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (SB == 2) {
        const float4 x = ldg4(hp + 4 * c);
        const float4 w0 = ldg4(wp + 8 * c);        // block 2c   : [i][o] = (00, 01, 10, 11)
        const float4 w1 = ldg4(wp + 8 * c + 4);    // block 2c+1
        m.v[i] = make_float4(fmaf(x.y, w0.z, x.x * w0.x) + r4.x, fmaf(x.y, w0.w, x.x * w0.y) + r4.y,
                             fmaf(x.w, w1.z, x.z * w1.x) + r4.z, fmaf(x.w, w1.w, x.z * w1.y) + r4.w);
      } else {
        float o4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = 4 * c + q;
          const int b = j / sb, o = j - b * sb;
          float acc = 0.f;
          for (int ii = 0; ii < sb; ++ii) acc = fmaf(__ldg(hp + b * sb + ii), __ldg(wp + ((size_t)b * sb + ii) * sb + o), acc);
          o4[q] = acc;
        }
        m.v[i] = make_float4(o4[0] + r4.x, o4[1] + r4.y, o4[2] + r4.z, o4[3] + r4.w);
      }
    } else {
      m.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Centroid tail: cbar = wgt * (sum0, sum) -> normalise to <x,x>_L = -1/c -> to_poincare -> log_0 -> clamp +-10.
template <int RV>
__device__ __forceinline__ void lorentz_finish(WarpRow<RV>& acc, float acc0, int K, float nv, const Curv& cv) {
  // reduce_func weights (:620): norms / (norms.sum() + 1e-6), all K equal; lorentz_centroid (:576): w / (sum w + eps)
  const float w0 = nv / ((float)K * nv + 1e-6f);
  const float wgt = w0 / ((float)K * w0 + kEps);
  acc.scale(wgt);
  acc0 *= wgt;
  const float ip = -acc0 * acc0 + acc.sumsq();
  const float scale = sqrtf(fmaxf(-ip * cv.c, kEps));
  const float y0 = acc0 / scale;
  const float den = fmaxf(1.0f + y0 * cv.sqrt_c, kEps);
  acc.map([=](float a) { return (a / scale) / den; });
  row_log0(acc, cv);
  acc.map([](float a) { return clampf_(a, -10.f, 10.f); });
}

// One warp per 32-edge virtual row of an active destination.  Per edge: message (above), exp_0 (one warp reduction;
// the norm after the projection follows analytically), to_lorentz, accumulate (time, space).  The centroid weights
// are equal inside a node, so chunks accumulate unweighted sums; single-chunk rows finish in place, hub rows leave
// (sum, sum0) partials that lorentz_fixup_kernel folds.
template <int RV, int SB>
__global__ void __launch_bounds__(256) lorentz_aggregate_kernel(
    const float* __restrict__ ht, const float* __restrict__ W, const float* __restrict__ rel,
    const int* __restrict__ rowptr, const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted,
    const float* __restrict__ norm, const int* __restrict__ vptr, const int* __restrict__ sptr,
    const int* __restrict__ vrow_row, int nv_rows, int d, int nb, Curv cv, float* __restrict__ out,
    float* __restrict__ partial, float* __restrict__ partial0) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (w >= nv_rows) return;
  const int nvec = d >> 2;
  const int row = __ldg(vrow_row + w);
  const int v0 = __ldg(vptr + row), v1 = __ldg(vptr + row + 1);
  const int k = w - v0;
  const int rbeg = __ldg(rowptr + row), rend = __ldg(rowptr + row + 1);
  const int beg = rbeg + k * kAggChunk;
  const int end = min(beg + kAggChunk, rend);
  const int sb = d / nb;
  const size_t wstride = (size_t)nb * sb * sb;
  WarpRow<RV> acc;
  acc.zero();
  float acc0 = 0.f;
  const int e0 = beg + lane;
  const int s_l = e0 < end ? __ldg(src_sorted + e0) : 0;
  const int t_l = e0 < end ? __ldg(etype_sorted + e0) : 0;
  const int cnt = end - beg;
  for (int j = 0; j < cnt; ++j) {
    const int s = __shfl_sync(0xffffffffu, s_l, j), t = __shfl_sync(0xffffffffu, t_l, j);
    WarpRow<RV> m;
    lorentz_message<RV, SB>(m, ht + (size_t)s * d, W + (size_t)t * wstride, rel ? rel + (size_t)t * d : nullptr, nvec, lane, sb);
    // exp_0 + projection as one scale factor (hyperbolic_ops.py:91-95, :51-53)
    const float mn = sqrtf(m.sumsq());
    const float n = fmaxf(mn, kEps);
    const float th = tanhf(cv.sqrt_c * n);
    const float f1 = th / (n * cv.sqrt_c);           // exp_0 scale: p = m * f1
    const float pn = fmaxf(mn * f1, kEps);
    const float f2 = fminf(pn, cv.proj_max) / pn;    // projection scale
    const float f = f1 * f2;
    const float nsq = (mn * f) * (mn * f);
    const float D = fmaxf(1.0f - cv.c * nsq, kEps);  // to_lorentz (:492-499)
    acc0 += (1.0f + cv.c * nsq) / (cv.sqrt_c * D);
    const float g = 2.0f * f / D;
#pragma unroll
    for (int i = 0; i < RV; ++i) acc.v[i] = f4_fma(g, m.v[i], acc.v[i]);
  }
  if (v1 - v0 == 1) {
    lorentz_finish(acc, acc0, rend - rbeg, __ldg(norm + row), cv);
    acc.store(out + (size_t)row * d, nvec, lane);
  } else {
    const size_t slot = (size_t)(__ldg(sptr + row) + k);
    acc.store(partial + slot * d, nvec, lane);
    if (lane == 0) partial0[slot] = acc0;
  }
}

template <int RV>
__global__ void __launch_bounds__(256) lorentz_fixup_kernel(
    const int* __restrict__ rowptr, const int* __restrict__ vptr, const int* __restrict__ sptr,
    const int* __restrict__ vrow_row, const float* __restrict__ norm, int nv_rows, int d, Curv cv,
    const float* __restrict__ partial, const float* __restrict__ partial0, float* __restrict__ out) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (w >= nv_rows) return;
  const int row = __ldg(vrow_row + w);
  const int v0 = __ldg(vptr + row);
  const int nch = __ldg(vptr + row + 1) - v0;
  if (w != v0 || nch <= 1) return;
  const int nvec = d >> 2;
  const int s0 = __ldg(sptr + row);
  WarpRow<RV> acc, p;
  acc.zero();
  float acc0 = 0.f;
  for (int k = 0; k < nch; ++k) {
    p.load_plain(partial + (size_t)(s0 + k) * d, nvec, lane);
#pragma unroll
    for (int i = 0; i < RV; ++i) acc.v[i] = f4_add(acc.v[i], p.v[i]);
    acc0 += partial0[s0 + k];
  }
  lorentz_finish(acc, acc0, __ldg(rowptr + row + 1) - __ldg(rowptr + row), __ldg(norm + row), cv);
  acc.store(out + (size_t)row * d, nvec, lane);
}

int lorentz_aggregate(const float* ht, const float* W, const float* rel, const int* rowptr,
                      const int* src_sorted, const int* etype_sorted, const float* norm, const int* vptr,
                      const int* sptr, const int* vrow_row, int nv_rows, int nsplit, int N, int d, int nb, double c,
                      float* out, float* partial, cudaStream_t st) {
  if (!ht || !W || !rowptr || !src_sorted || !etype_sorted || !norm || !vptr || !sptr || !vrow_row || !out) { set_last_error("lorentz_aggregate: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || d > 256 || nb <= 0 || d % nb) { set_last_error("lorentz_aggregate: d=%d nb=%d unsupported", d, nb); return REGCN_ERR_UNSUPPORTED; }
  if (nsplit > 0 && !partial) { set_last_error("lorentz_aggregate: split rows need a partial buffer of n_split_chunks*(d+1) floats"); return REGCN_ERR_WORKSPACE; }
  const int TB = 256;
  Curv cv = make_curv(c);
  // isolated destinations: exact zero rows (DGL zero fill; to_poincare(0) = 0, log_0(0) = 0)
  const size_t total = (size_t)N * (d >> 2);
  (void)total;
  launch_k(zero_inactive_rows_kernel, sweep_grid(N), TB, 0, st, rowptr, N, d, out, nullptr, nullptr);
  if (nv_rows > 0) {
    const unsigned grid = (unsigned)(((size_t)nv_rows * 32 + TB - 1) / TB);
    float* partial0 = partial ? partial + (size_t)nsplit * d : nullptr;
    const bool sb2 = (d / nb) == 2;
    if (d <= 128) {
      if (sb2) launch_k(lorentz_aggregate_kernel<1, 2>, grid, TB, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv_rows, d, nb, cv, out, partial, partial0);
      else launch_k(lorentz_aggregate_kernel<1, 0>, grid, TB, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv_rows, d, nb, cv, out, partial, partial0);
      if (nsplit > 0) launch_k(lorentz_fixup_kernel<1>, grid, TB, 0, st, rowptr, vptr, sptr, vrow_row, norm, nv_rows, d, cv, partial, partial0, out);
    } else {
      if (sb2) launch_k(lorentz_aggregate_kernel<2, 2>, grid, TB, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv_rows, d, nb, cv, out, partial, partial0);
      else launch_k(lorentz_aggregate_kernel<2, 0>, grid, TB, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, norm, vptr, sptr, vrow_row, nv_rows, d, nb, cv, out, partial, partial0);
      if (nsplit > 0) launch_k(lorentz_fixup_kernel<2>, grid, TB, 0, st, rowptr, vptr, sptr, vrow_row, norm, nv_rows, d, cv, partial, partial0, out);
    }
  }
  return check_launch("lorentz_aggregate");
}

// ---------------------------------------------------------------------------
// K2: x_input[r] = x_input[r+R] = mean_{e in ents(r)} h[e]; absent relations stay zero.
// grid = (R, nsplit): each block reduces a slice of the relation's entity list.
// ---------------------------------------------------------------------------
template <int RV>
__global__ void __launch_bounds__(256) rel_mean_pool_kernel(
    const float* __restrict__ h, const int* __restrict__ rel_rowptr, const int* __restrict__ rel_ents,
    int R, int d, int nsplit, float* __restrict__ out, float* __restrict__ partial,
    float* __restrict__ out_hi, float* __restrict__ out_lo) {
  pdl_grid_sync();
  __shared__ float4 red[8][RV * 32];
  const int r = blockIdx.x, sp = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nvec = d >> 2;
  const int beg = __ldg(rel_rowptr + r), end = __ldg(rel_rowptr + r + 1);
  const int n = end - beg;
  const int per = (n + nsplit - 1) / nsplit;
  const int b = beg + sp * per, e = min(b + per, end);
  WarpRow<RV> acc, row;
  acc.zero();
  for (int i = b + wid; i < e; i += nw) {
    row.load(h + (size_t)__ldg(rel_ents + i) * d, nvec, lane);
#pragma unroll
    for (int q = 0; q < RV; ++q) acc.v[q] = f4_add(acc.v[q], row.v[q]);
  }
#pragma unroll
  for (int q = 0; q < RV; ++q) red[wid][q * 32 + lane] = acc.v[q];
  __syncthreads();
  if (wid == 0) {
    for (int w2 = 1; w2 < nw; ++w2) {
#pragma unroll
      for (int q = 0; q < RV; ++q) acc.v[q] = f4_add(acc.v[q], red[w2][q * 32 + lane]);
    }
    if (nsplit == 1) {
      if (n > 0) {
        const float fn = (float)n;
        acc.map([=](float a) { return a / fn; });
      }
      if (out) {
        acc.store(out + (size_t)r * d, nvec, lane);
        acc.store(out + (size_t)(r + R) * d, nvec, lane);
      }
      if (out_hi) {
        acc.store_split(out_hi + (size_t)r * d, out_lo + (size_t)r * d, nvec, lane);
        acc.store_split(out_hi + (size_t)(r + R) * d, out_lo + (size_t)(r + R) * d, nvec, lane);
      }
    } else {
      acc.store(partial + ((size_t)r * nsplit + sp) * d, nvec, lane);
    }
  }
}

template <int RV>
__global__ void rel_mean_finalize_kernel(const float* __restrict__ partial, const int* __restrict__ rel_rowptr,
                                         int R, int d, int nsplit, float* __restrict__ out,
                                         float* __restrict__ out_hi, float* __restrict__ out_lo) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int r = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (r >= R) return;
  const int nvec = d >> 2;
  const int n = __ldg(rel_rowptr + r + 1) - __ldg(rel_rowptr + r);
  WarpRow<RV> acc, p;
  acc.zero();
  for (int s = 0; s < nsplit; ++s) {
    p.load_plain(partial + ((size_t)r * nsplit + s) * d, nvec, lane);
#pragma unroll
    for (int q = 0; q < RV; ++q) acc.v[q] = f4_add(acc.v[q], p.v[q]);
  }
  if (n > 0) {
    const float fn = (float)n;
    acc.map([=](float a) { return a / fn; });
  }
  if (out) {
    acc.store(out + (size_t)r * d, nvec, lane);
    acc.store(out + (size_t)(r + R) * d, nvec, lane);
  }
  if (out_hi) {
    acc.store_split(out_hi + (size_t)r * d, out_lo + (size_t)r * d, nvec, lane);
    acc.store_split(out_hi + (size_t)(r + R) * d, out_lo + (size_t)(r + R) * d, nvec, lane);
  }
}

int rel_mean_pool(const float* h, const int* rel_rowptr, const int* rel_ents, int R, int d, int nsplit,
                  float* out, float* partial, float* out_hi, float* out_lo, cudaStream_t st) {
  if (!h || !rel_rowptr || !rel_ents || (!out && !out_hi) || (out_hi && !out_lo)) { set_last_error("rel_mean_pool: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || d > 256) { set_last_error("rel_mean_pool: d=%d unsupported", d); return REGCN_ERR_UNSUPPORTED; }
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 1 && !partial) { set_last_error("rel_mean_pool: nsplit>1 needs a partial buffer"); return REGCN_ERR_WORKSPACE; }
  dim3 grid(R, nsplit);
  if (d <= 128) launch_k(rel_mean_pool_kernel<1>, grid, 256, 0, st, h, rel_rowptr, rel_ents, R, d, nsplit, out, partial, out_hi, out_lo);
  else launch_k(rel_mean_pool_kernel<2>, grid, 256, 0, st, h, rel_rowptr, rel_ents, R, d, nsplit, out, partial, out_hi, out_lo);
  if (nsplit > 1) {
    const unsigned g2 = (unsigned)(((size_t)R * 32 + 255) / 256);
    if (d <= 128) launch_k(rel_mean_finalize_kernel<1>, g2, 256, 0, st, partial, rel_rowptr, R, d, nsplit, out, out_hi, out_lo);
    else launch_k(rel_mean_finalize_kernel<2>, g2, 256, 0, st, partial, rel_rowptr, R, d, nsplit, out, out_hi, out_lo);
  }
  return check_launch("rel_mean_pool");
}


// ---------------------------------------------------------------------------------------------------------------------
// K7 backward (lgcn encoder in training; hyperbolic_layers.py:589-625, ops:492-518,563-581).  The centroid weights are
// equal inside a node, so d(out_v)/d(mL_e) is ONE (d+1)-vector G_v per destination: a node kernel produces it, the edge
// kernels recompute every message and push G through to_lorentz and exp_0.  2x2 relation blocks (the reference's 100
// bases at d = 200) stay chunk-local in registers; any other block size takes the generic variants (SB == 0).
// ---------------------------------------------------------------------------------------------------------------------
template <int RV>
__device__ __forceinline__ void lorentz_edge_forward(const WarpRow<RV>& m, const Curv& cv, float& f, float& D, float& mn) {
  mn = sqrtf(m.sumsq());
  const float n = fmaxf(mn, kEps);
  const float th = tanhf(cv.sqrt_c * n);
  const float f1 = th / (n * cv.sqrt_c);
  const float pn = fmaxf(mn * f1, kEps);
  const float f2 = fminf(pn, cv.proj_max) / pn;
  f = f1 * f2;
  const float nsq = (mn * f) * (mn * f);
  D = fmaxf(1.0f - cv.c * nsq, kEps);
}
// gradient of the loss w.r.t. the message m_e given G = dL/d(mL_e) = (g0, g):  p = f m;
//   dp = (4 sqrt_c g0 / D^2) p + 2 g / D + (4 c <p,g> / D^2) p;   dm = s dp + (s'/n) <m,dp> m   (exp_0 with projection)
template <int RV>
__device__ __forceinline__ void lorentz_message_grad(WarpRow<RV>& m, float g0, const WarpRow<RV>& g, const Curv& cv) {
  float f, D, mn;
  lorentz_edge_forward(m, cv, f, D, mn);
  const float pg = f * m.dot(g);
  const float kp = (4.0f * cv.sqrt_c * g0 + 4.0f * cv.c * pg) / (D * D) * f;     // coefficient of m in dp
  const float kg = 2.0f / D;
  WarpRow<RV> dp = m;
  dp.zip(g, [=](float mm, float gg) { return kp * mm + kg * gg; });
  // radial backward of exp_0 (+ projection) at m
  const float n = fmaxf(mn, kEps);
  const float t = tanhf(cv.sqrt_c * n);
  float s, sp;
  if (t / cv.sqrt_c > cv.proj_max) { s = cv.proj_max / n; sp = -cv.proj_max / (n * n); }
  else { s = t / (cv.sqrt_c * n); sp = (1.0f - t * t) / n - t / (cv.sqrt_c * n * n); }
  if (!(mn > kEps)) sp = 0.f;
  const float coef = sp / n * m.dot(dp);
  m.zip(dp, [=](float mm, float dd) { return s * dd + coef * mm; });
}

template <int RV, int SB>
__global__ void __launch_bounds__(256) lorentz_node_grad_kernel(
    const float* __restrict__ ht, const float* __restrict__ W, const float* __restrict__ rel,
    const int* __restrict__ rowptr, const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted,
    const float* __restrict__ norm, const float* __restrict__ gout, int N, int d, int nb, Curv cv,
    float* __restrict__ G0, float* __restrict__ G) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= N) return;
  const int nvec = d >> 2;
  const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  WarpRow<RV> acc;
  acc.zero();
  if (e == b) {
    acc.store(G + (size_t)row * d, nvec, lane);
    if (lane == 0) G0[row] = 0.f;
    return;
  }
  const int sb = d / nb;
  const size_t wstride = (size_t)d * sb;
  float acc0 = 0.f;
  for (int p = b; p < e; ++p) {
    const int s = __ldg(src_sorted + p), t = __ldg(etype_sorted + p);
    WarpRow<RV> m;
    lorentz_message<RV, SB>(m, ht + (size_t)s * d, W + (size_t)t * wstride, rel + (size_t)t * d, nvec, lane, sb);
    float f, D, mn;
    lorentz_edge_forward(m, cv, f, D, mn);
    const float nsq = (mn * f) * (mn * f);
    acc0 += (1.0f + cv.c * nsq) / (cv.sqrt_c * D);
    const float g = 2.0f * f / D;
#pragma unroll
    for (int i = 0; i < RV; ++i) acc.v[i] = f4_fma(g, m.v[i], acc.v[i]);
  }
  // forward tail (lorentz_finish) with its intermediates
  const int K = e - b;
  const float nv = __ldg(norm + row);
  const float w0 = nv / ((float)K * nv + 1e-6f);
  const float wgt = w0 / ((float)K * w0 + kEps);
  acc.scale(wgt);                                        // a
  const float a0 = acc0 * wgt;
  const float ip = -a0 * a0 + acc.sumsq();
  const bool sc_free = -ip * cv.c > kEps;
  const float sc = sqrtf(fmaxf(-ip * cv.c, kEps));
  const float Q = sc + cv.sqrt_c * a0;                   // sc * (1 + sqrt_c a0 / sc)
  const bool q_free = Q / sc > kEps;
  const float Qc = q_free ? Q : kEps * sc;
  WarpRow<RV> y = acc;
  y.scale(1.0f / Qc);
  // log_0 and the +-10 clamp, backward
  WarpRow<RV> t = y;
  row_log0(t, cv);
  WarpRow<RV> g;
  g.load_plain(gout + (size_t)row * d, nvec, lane);
  g.zip(t, [](float gg, float tt) { return (tt >= -10.f && tt <= 10.f) ? gg : 0.f; });
  {
    const float nraw = sqrtf(y.sumsq());
    const float n = fmaxf(nraw, kEps);
    const float u = cv.sqrt_c * n;
    const bool clamped = u >= 1.0f - kEps;
    const float a = atanhf(fminf(u, 1.0f - kEps));
    const float s = a / (cv.sqrt_c * n);
    float sp = (clamped ? 0.f : 1.0f / (n * (1.0f - cv.c * n * n))) - a / (cv.sqrt_c * n * n);
    if (!(nraw > kEps)) sp = 0.f;
    const float coef = sp / n * y.dot(g);
    g.zip(y, [=](float gg, float yy) { return s * gg + coef * yy; });       // g = dL/dy
  }
  // y = a / Q
  const float dQ = q_free ? -g.dot(acc) / (Qc * Qc) : 0.f;
  float da0 = cv.sqrt_c * dQ;
  const float dip = sc_free ? dQ * (-cv.c / (2.0f * sc)) : 0.f;
  da0 += dip * (-2.0f * a0);
  const float k1 = 1.0f / Qc, k2 = 2.0f * dip;
  g.zip(acc, [=](float gg, float aa) { return (k1 * gg + k2 * aa) * wgt; });
  g.store(G + (size_t)row * d, nvec, lane);
  if (lane == 0) G0[row] = da0 * wgt;
}

// dht[u] = sum over the in-edges (w -> u, r') of row u of W[inv r']^T dm(u -> w, inv r')   (the graph holds every edge with
// its inverse, so the forward CSR enumerates u's out-edges too)
template <int RV, int SB>
__global__ void __launch_bounds__(256) lorentz_grad_src_kernel(
    const float* __restrict__ ht, const float* __restrict__ W, const float* __restrict__ rel,
    const int* __restrict__ rowptr, const int* __restrict__ src_sorted, const int* __restrict__ etype_sorted,
    const float* __restrict__ G0, const float* __restrict__ G, int N, int d, int nb, int R, Curv cv,
    float* __restrict__ dht) {
  pdl_grid_sync();
  __shared__ __align__(16) float dms[SB == 2 ? 4 : 8 * 256];     // generic blocks: a warp's dm, read across lanes
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= N) return;
  const int nvec = d >> 2;
  const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  const int sb = d / nb;
  const size_t wstride = (size_t)d * sb;
  float* dmw = dms + (SB == 2 ? 0 : (threadIdx.x >> 5) * 256);
  WarpRow<RV> acc;
  acc.zero();
  for (int p = b; p < e; ++p) {
    const int w = __ldg(src_sorted + p), t = __ldg(etype_sorted + p);
    const int ti = t < R ? t + R : t - R;
    const float* wp = W + (size_t)ti * wstride;
    WarpRow<RV> m, g;
    lorentz_message<RV, SB>(m, ht + (size_t)row * d, wp, rel + (size_t)ti * d, nvec, lane, sb);
    g.load_plain(G + (size_t)w * d, nvec, lane);
    lorentz_message_grad(m, __ldg(G0 + w), g, cv);         // m <- dm
    if (SB == 2) {
#pragma unroll
      for (int i = 0; i < RV; ++i) {
        const int c = lane + i * kWarp;
        if (c < nvec) {
          const float4 w0 = ldg4(wp + 8 * c), w1 = ldg4(wp + 8 * c + 4);
          const float4 dm = m.v[i];
          acc.v[i].x += w0.x * dm.x + w0.y * dm.y;
          acc.v[i].y += w0.z * dm.x + w0.w * dm.y;
          acc.v[i].z += w1.x * dm.z + w1.y * dm.w;
          acc.v[i].w += w1.z * dm.z + w1.w * dm.w;
        }
      }
    } else {
      // dx[j] = sum_o W[j][o] dm[block(j) * sb + o]  (row j of the (d, sb) weight of this relation)
      __syncwarp();
      m.store(dmw, nvec, lane);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < RV; ++i) {
        const int c = lane + i * kWarp;
        if (c < nvec) {
          float o4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = 4 * c + q;
            const float* dmb = dmw + (j / sb) * sb;
            const float* wr = wp + (size_t)j * sb;
            float a = 0.f;
            for (int o = 0; o < sb; ++o) a = fmaf(__ldg(wr + o), dmb[o], a);
            o4[q] = a;
          }
          acc.v[i].x += o4[0]; acc.v[i].y += o4[1]; acc.v[i].z += o4[2]; acc.v[i].w += o4[3];
        }
      }
    }
  }
  acc.store(dht + (size_t)row * d, nvec, lane);
}

// per relation type (edges grouped by type: type_src, type_dst): drel[r] = sum dm, dW[r] = sum ht[src] (x) dm per 2x2 block.
// CTA = (type, split); its 8 warps stride over the edges, partial sums folded through shared memory in warp order.
constexpr int kLorentzSplit = 8;
template <int RV>
__global__ void __launch_bounds__(256) lorentz_grad_type_kernel(
    const float* __restrict__ ht, const float* __restrict__ W, const float* __restrict__ rel,
    const int* __restrict__ type_rowptr, const int* __restrict__ type_src, const int* __restrict__ type_dst,
    const float* __restrict__ G0, const float* __restrict__ G, int d, int nb, int R2, Curv cv,
    float* __restrict__ part_rel, float* __restrict__ part_w) {
  pdl_grid_sync();
  extern __shared__ float4 lsm[];                 // [8 warps][3][32*RV]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x, sp = blockIdx.y;
  const int nvec = d >> 2;
  const int tb = __ldg(type_rowptr + r), te = __ldg(type_rowptr + r + 1);
  const int per = (te - tb + kLorentzSplit - 1) / kLorentzSplit;
  const int e0 = tb + sp * per, e1 = min(te, e0 + per);
  const float* wp = W + (size_t)r * nb * 4;
  WarpRow<RV> ar, aw0, aw1;
  ar.zero(); aw0.zero(); aw1.zero();
  for (int e = e0 + wid; e < e1; e += 8) {
    const int u = __ldg(type_src + e), v = __ldg(type_dst + e);
    WarpRow<RV> m, g, x;
    x.load_plain(ht + (size_t)u * d, nvec, lane);
    lorentz_message<RV, 2>(m, ht + (size_t)u * d, wp, rel + (size_t)r * d, nvec, lane, 2);
    g.load_plain(G + (size_t)v * d, nvec, lane);
    lorentz_message_grad(m, __ldg(G0 + v), g, cv);
#pragma unroll
    for (int i = 0; i < RV; ++i) {
      const float4 dm = m.v[i], xx = x.v[i];
      ar.v[i] = f4_add(ar.v[i], dm);
      aw0.v[i].x += xx.x * dm.x; aw0.v[i].y += xx.x * dm.y; aw0.v[i].z += xx.y * dm.x; aw0.v[i].w += xx.y * dm.y;
      aw1.v[i].x += xx.z * dm.z; aw1.v[i].y += xx.z * dm.w; aw1.v[i].z += xx.w * dm.z; aw1.v[i].w += xx.w * dm.w;
    }
  }
  const int W3 = 32 * RV;
#pragma unroll
  for (int i = 0; i < RV; ++i) {
    lsm[(wid * 3 + 0) * W3 + lane + i * kWarp] = ar.v[i];
    lsm[(wid * 3 + 1) * W3 + lane + i * kWarp] = aw0.v[i];
    lsm[(wid * 3 + 2) * W3 + lane + i * kWarp] = aw1.v[i];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < RV; ++i) {
      const int c = lane + i * kWarp;
      if (c < nvec) {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0;
        for (int w = 0; w < 8; ++w) {
          s0 = f4_add(s0, lsm[(w * 3 + 0) * W3 + c]);
          s1 = f4_add(s1, lsm[(w * 3 + 1) * W3 + c]);
          s2 = f4_add(s2, lsm[(w * 3 + 2) * W3 + c]);
        }
        const size_t pr = (size_t)sp * R2 + r;
        st4(part_rel + pr * d + 4 * c, s0);
        st4(part_w + pr * (size_t)nb * 4 + 8 * c, s1);
        st4(part_w + pr * (size_t)nb * 4 + 8 * c + 4, s2);
      }
    }
  }
}

// Generic block size sb = d / nb (the reference clamps num_bases to 2R, hyperbolic_layers.py:559-561: 10x10 blocks on
// a 10-relation dataset): dW[r] is a (d, sb) matrix, dW[j][o] = sum_e x_e[j] dm_e[block(j) sb + o].  CTA = (type, split);
// per round its 8 warps recompute dm of 8 edges into shared memory, then all threads add the round's outer products to
// the shared (d, sb) accumulator in warp order (fixed summation order: bit-reproducible).
template <int RV>
__global__ void __launch_bounds__(256) lorentz_grad_type_generic_kernel(
    const float* __restrict__ ht, const float* __restrict__ W, const float* __restrict__ rel,
    const int* __restrict__ type_rowptr, const int* __restrict__ type_src, const int* __restrict__ type_dst,
    const float* __restrict__ G0, const float* __restrict__ G, int d, int nb, int R2, Curv cv,
    float* __restrict__ part_rel, float* __restrict__ part_w) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float gsm[];                  // xs [8][d] | dm [8][d] | accW [d * sb] | accR [d]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x, sp = blockIdx.y;
  const int nvec = d >> 2;
  const int sb = d / nb;
  const int nW = d * sb;
  float* xs = gsm;
  float* dmsm = gsm + 8 * d;
  float* accW = gsm + 16 * d;
  float* accR = accW + nW;
  for (int i = threadIdx.x; i < nW + d; i += 256) accW[i] = 0.f;
  const int tb = __ldg(type_rowptr + r), te = __ldg(type_rowptr + r + 1);
  const int per = (te - tb + kLorentzSplit - 1) / kLorentzSplit;
  const int e0 = tb + sp * per, e1 = min(te, e0 + per);
  const float* wp = W + (size_t)r * nW;
  for (int base = e0; base < e1; base += 8) {
    const int e = base + wid;
    WarpRow<RV> m, x;
    if (e < e1) {
      const int u = __ldg(type_src + e), v = __ldg(type_dst + e);
      WarpRow<RV> g;
      x.load_plain(ht + (size_t)u * d, nvec, lane);
      lorentz_message<RV, 0>(m, ht + (size_t)u * d, wp, rel + (size_t)r * d, nvec, lane, sb);
      g.load_plain(G + (size_t)v * d, nvec, lane);
      lorentz_message_grad(m, __ldg(G0 + v), g, cv);
    } else {
      m.zero(); x.zero();
    }
    __syncthreads();                              // the previous round's outer products are done
    x.store(xs + wid * d, nvec, lane);
    m.store(dmsm + wid * d, nvec, lane);
    __syncthreads();
    for (int i = threadIdx.x; i < nW; i += 256) {
      const int j = i / sb, o = i - j * sb;
      const int bo = (j / sb) * sb + o;
      float a = accW[i];
#pragma unroll
      for (int w = 0; w < 8; ++w) a = fmaf(xs[w * d + j], dmsm[w * d + bo], a);
      accW[i] = a;
    }
    for (int i = threadIdx.x; i < d; i += 256) {
      float a = accR[i];
#pragma unroll
      for (int w = 0; w < 8; ++w) a += dmsm[w * d + i];
      accR[i] = a;
    }
  }
  __syncthreads();
  const size_t pr = (size_t)sp * R2 + r;
  for (int i = threadIdx.x; i < nW; i += 256) part_w[pr * nW + i] = accW[i];
  for (int i = threadIdx.x; i < d; i += 256) part_rel[pr * d + i] = accR[i];
}

size_t lorentz_aggregate_bwd_workspace_bytes(int N, int R2, int d) {
  return ((size_t)N * (d + 1) + (size_t)kLorentzSplit * R2 * 3 * d) * sizeof(float) + 1024;   // G, G0 (+ slack)
}
int lorentz_aggregate_bwd(const float* ht, const float* W, const float* rel, const float* gout, const int* rowptr,
                          const int* src_sorted, const int* etype_sorted, const float* norm, const int* type_rowptr,
                          const int* type_src, const int* type_dst, int N, int R2, int d, int nb, double c, float* dht,
                          float* part_rel, float* part_w, float* ws, size_t ws_bytes, cudaStream_t st) {
  if (!ht || !W || !rel || !gout || !rowptr || !src_sorted || !etype_sorted || !norm || !type_rowptr || !type_src ||
      !type_dst || !dht || !part_rel || !part_w || !ws) { set_last_error("lorentz_aggregate_bwd: null pointer"); return REGCN_ERR_NULL; }
  if (d <= 0 || (d & 3) || d > 256 || nb <= 0 || d % nb || (R2 & 1)) { set_last_error("lorentz_aggregate_bwd: needs d%%4==0, d<=256, num_bases dividing d (d=%d nb=%d)", d, nb); return REGCN_ERR_UNSUPPORTED; }
  if (ws_bytes < (size_t)N * (d + 1) * sizeof(float)) { set_last_error("lorentz_aggregate_bwd: workspace too small"); return REGCN_ERR_WORKSPACE; }
  if (N <= 0) return REGCN_OK;
  Curv cv = make_curv(c);
  float* G = ws;
  float* G0 = ws + (size_t)N * d;
  const unsigned grid = (unsigned)(((size_t)N * 32 + 255) / 256);
  dim3 tgrid((unsigned)R2, (unsigned)kLorentzSplit);
  const int sb = d / nb;
  if (sb != 2) {
    // generic relation blocks: (d, sb) gradient accumulator of a relation in shared memory
    const size_t smem = ((size_t)17 * d + (size_t)d * sb) * sizeof(float);
    if (smem > 200 * 1024) { set_last_error("lorentz_aggregate_bwd: relation blocks of %dx%d at d=%d need %zu bytes of shared memory", sb, sb, d, smem); return REGCN_ERR_UNSUPPORTED; }
    if (d <= 128) {
      cudaFuncSetAttribute(lorentz_grad_type_generic_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_k(lorentz_node_grad_kernel<1, 0>, grid, 256, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, norm, gout, N, d, nb, cv, G0, G);
      launch_k(lorentz_grad_src_kernel<1, 0>, grid, 256, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, (const float*)G0, (const float*)G, N, d, nb, R2 / 2, cv, dht);
      launch_k(lorentz_grad_type_generic_kernel<1>, tgrid, 256, smem, st, ht, W, rel, type_rowptr, type_src, type_dst, (const float*)G0, (const float*)G, d, nb, R2, cv, part_rel, part_w);
    } else {
      cudaFuncSetAttribute(lorentz_grad_type_generic_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_k(lorentz_node_grad_kernel<2, 0>, grid, 256, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, norm, gout, N, d, nb, cv, G0, G);
      launch_k(lorentz_grad_src_kernel<2, 0>, grid, 256, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, (const float*)G0, (const float*)G, N, d, nb, R2 / 2, cv, dht);
      launch_k(lorentz_grad_type_generic_kernel<2>, tgrid, 256, smem, st, ht, W, rel, type_rowptr, type_src, type_dst, (const float*)G0, (const float*)G, d, nb, R2, cv, part_rel, part_w);
    }
    return check_launch("lorentz_aggregate_bwd");
  }
  if (d <= 128) {
    launch_k(lorentz_node_grad_kernel<1, 2>, grid, 256, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, norm, gout, N, d, nb, cv, G0, G);
    launch_k(lorentz_grad_src_kernel<1, 2>, grid, 256, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, (const float*)G0, (const float*)G, N, d, nb, R2 / 2, cv, dht);
    launch_k(lorentz_grad_type_kernel<1>, tgrid, 256, (size_t)8 * 3 * 32 * sizeof(float4), st, ht, W, rel, type_rowptr, type_src, type_dst, (const float*)G0, (const float*)G, d, nb, R2, cv, part_rel, part_w);
  } else {
    launch_k(lorentz_node_grad_kernel<2, 2>, grid, 256, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, norm, gout, N, d, nb, cv, G0, G);
    launch_k(lorentz_grad_src_kernel<2, 2>, grid, 256, 0, st, ht, W, rel, rowptr, src_sorted, etype_sorted, (const float*)G0, (const float*)G, N, d, nb, R2 / 2, cv, dht);
    launch_k(lorentz_grad_type_kernel<2>, tgrid, 256, (size_t)8 * 3 * 64 * sizeof(float4), st, ht, W, rel, type_rowptr, type_src, type_dst, (const float*)G0, (const float*)G, d, nb, R2, cv, part_rel, part_w);
  }
  return check_launch("lorentz_aggregate_bwd");
}
int lorentz_bwd_splits() { return kLorentzSplit; }
}  // namespace regcn
