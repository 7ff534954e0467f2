// K1: snapshot edge-index build on the device.
// Restates rgcn/utils.py:100-134 (build_sub_graph: inverse-edge doubling, in-degree,
// norm = 1/max(indeg,1)) and rgcn/utils.py:78-97 (r2e: per-relation entity sets) as a
// CSR-by-destination index plus a relation->entity CSR.  Integer work: bit-exact.
#include "../../include/regcn_b200.h"
#include "common.cuh"
#include "internal.h"
#include <cub/cub.cuh>
#include <algorithm>

namespace regcn {

constexpr int kAggChunk = 32;  // edges per virtual row (hub rows are split into chunks of this many)

// E = 2T edges in the reference's order: [src;dst] -> [dst;src], type [rel; rel+R]  (utils.py:116-118)
__global__ void expand_edges_kernel(const int64_t* __restrict__ triples, int T, int R,
                                    int* __restrict__ src, int* __restrict__ dst, int* __restrict__ etype,
                                    int* __restrict__ indeg, int* __restrict__ eid,
                                    unsigned long long* __restrict__ rel_keys, int N) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int s = (int)triples[3 * (size_t)t + 0];
  int r = (int)triples[3 * (size_t)t + 1];
  int o = (int)triples[3 * (size_t)t + 2];
  src[t] = s; dst[t] = o; etype[t] = r;
  src[T + t] = o; dst[T + t] = s; etype[T + t] = r + R;
  eid[t] = t; eid[T + t] = T + t;
  atomicAdd(&indeg[o], 1);
  atomicAdd(&indeg[s], 1);
  rel_keys[t] = (unsigned long long)r * (unsigned long long)N + (unsigned long long)s;
  rel_keys[T + t] = (unsigned long long)r * (unsigned long long)N + (unsigned long long)o;
}

// Virtual rows exist only for ACTIVE destinations (in-degree > 0): real snapshots touch ~10% of the entities, and
// the aggregate kernels walk the active rows only.
__global__ void norm_chunks_kernel(const int* __restrict__ indeg, int N, float* __restrict__ norm,
                                   int* __restrict__ nchunk, int* __restrict__ nsplit, int* __restrict__ active_flag,
                                   int* __restrict__ max_deg) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  int d = indeg[v];
  norm[v] = 1.0f / (float)(d == 0 ? 1 : d);  // comp_deg_norm, utils.py:110-114
  int nc = (d + kAggChunk - 1) / kAggChunk;  // 0 for isolated nodes
  nchunk[v] = nc;
  nsplit[v] = nc > 1 ? nc : 0;
  active_flag[v] = d > 0 ? 1 : 0;
  if (d > kAggChunk) atomicMax(max_deg, d);
}

__global__ void active_pos_kernel(const int* __restrict__ indeg, const int* __restrict__ scan, int N,
                                  int* __restrict__ active_pos, int* __restrict__ active_rows) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const bool act = indeg[v] > 0;
  active_pos[v] = act ? scan[v] : -1;
  if (act) active_rows[scan[v]] = v;            // inverse map: the sorted list of active destinations
}

__global__ void gather_sorted_kernel(const int* __restrict__ eperm, const int* __restrict__ src,
                                     const int* __restrict__ etype, int E,
                                     int* __restrict__ src_sorted, int* __restrict__ etype_sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  int e = eperm[i];
  src_sorted[i] = src[e];
  etype_sorted[i] = etype[e];
}

__global__ void fill_vrows_kernel(const int* __restrict__ vptr, int N, int* __restrict__ vrow_row) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  int b = vptr[v], e = vptr[v + 1];
  for (int k = b; k < e; ++k) vrow_row[k] = v;
}

__global__ void rel_heads_kernel(const unsigned long long* __restrict__ keys, int n, int N,
                                 int* __restrict__ head, int* __restrict__ rel_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int h = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
  head[i] = h;
  if (h) atomicAdd(&rel_count[(int)(keys[i] / (unsigned long long)N)], 1);
}

__global__ void rel_compact_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ head,
                                   const int* __restrict__ pos, int n, int N, int* __restrict__ rel_ents) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (head[i]) rel_ents[pos[i]] = (int)(keys[i] % (unsigned long long)N);
}

__global__ void write_counts_kernel(const int* __restrict__ vptr, const int* __restrict__ sptr, int N,
                                    const int* __restrict__ rel_rowptr, int R, const int* __restrict__ max_deg,
                                    int* __restrict__ rowptr, int E, const int* __restrict__ active_scan,
                                    int* __restrict__ counts) {
  rowptr[N] = E;
  counts[0] = vptr[N];
  counts[1] = sptr[N];
  counts[2] = rel_rowptr[R];
  counts[3] = *max_deg;
  counts[4] = active_scan[N];
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BuildWs {
  size_t eid, eid_alt, dstkey_alt, rel_keys, rel_keys_alt, head, pos, nchunk, nsplit, aflag, ascan, rel_count, max_deg, cub, total;
  size_t cub_bytes;
};

static BuildWs plan_build_ws(int T, int N, int R) {
  size_t E = 2 * (size_t)T;
  BuildWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  w.eid = take(E * 4); w.eid_alt = take(E * 4); w.dstkey_alt = take(E * 4);
  w.rel_keys = take(E * 8); w.rel_keys_alt = take(E * 8);
  w.head = take(E * 4); w.pos = take(E * 4);
  w.nchunk = take(((size_t)N + 1) * 4); w.nsplit = take(((size_t)N + 1) * 4);
  w.aflag = take(((size_t)N + 1) * 4); w.ascan = take(((size_t)N + 1) * 4);
  w.rel_count = take(((size_t)R + 1) * 4); w.max_deg = take(4);
  size_t b1 = 0, b2 = 0, b3 = 0, b4 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, (int)E);
  cub::DeviceRadixSort::SortKeys(nullptr, b2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)E);
  cub::DeviceScan::ExclusiveSum(nullptr, b3, (const int*)nullptr, (int*)nullptr, N + 1);
  cub::DeviceScan::ExclusiveSum(nullptr, b4, (const int*)nullptr, (int*)nullptr, (int)E);
  w.cub_bytes = b1;
  if (b2 > w.cub_bytes) w.cub_bytes = b2;
  if (b3 > w.cub_bytes) w.cub_bytes = b3;
  if (b4 > w.cub_bytes) w.cub_bytes = b4;
  w.cub_bytes += 256;
  w.cub = take(w.cub_bytes);
  w.total = off;
  return w;
}

size_t csr_build_workspace_bytes(int T, int N, int R) { return plan_build_ws(T, N, R).total; }

int csr_build(const int64_t* triples, int T, int N, int R,
              int* src, int* dst, int* etype, int* indeg, float* norm,
              int* rowptr, int* src_sorted, int* etype_sorted, int* eperm,
              int* vptr, int* sptr, int* vrow_row, int* active_pos, int* active_rows,
              int* rel_rowptr, int* rel_ents, int* counts,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  if (T < 0 || N <= 0 || R <= 0) { set_last_error("csr_build: bad dims T=%d N=%d R=%d", T, N, R); return REGCN_ERR_DIM; }
  BuildWs w = plan_build_ws(T, N, R);
  if (ws_bytes < w.total) { set_last_error("csr_build: workspace %zu < %zu", ws_bytes, w.total); return REGCN_ERR_WORKSPACE; }
  if (!src || !dst || !etype || !indeg || !norm || !rowptr || !src_sorted || !etype_sorted || !eperm || !vptr ||
      !sptr || !vrow_row || !active_pos || !active_rows || !rel_rowptr || !rel_ents || !counts || !ws || (T > 0 && !triples)) {
    set_last_error("csr_build: null pointer"); return REGCN_ERR_NULL;
  }
  char* base = (char*)ws;
  int E = 2 * T;
  int* eid = (int*)(base + w.eid);
  int* dst_alt = (int*)(base + w.dstkey_alt);
  unsigned long long* rk = (unsigned long long*)(base + w.rel_keys);
  unsigned long long* rk_alt = (unsigned long long*)(base + w.rel_keys_alt);
  int* head = (int*)(base + w.head);
  int* pos = (int*)(base + w.pos);
  int* nchunk = (int*)(base + w.nchunk);
  int* nsplit = (int*)(base + w.nsplit);
  int* aflag = (int*)(base + w.aflag);
  int* ascan = (int*)(base + w.ascan);
  int* rel_count = (int*)(base + w.rel_count);
  int* max_deg = (int*)(base + w.max_deg);
  void* cubtmp = base + w.cub;
  size_t cb = w.cub_bytes;
  const int TB = 256;

  cudaMemsetAsync(indeg, 0, (size_t)N * 4, st);
  cudaMemsetAsync(rel_count, 0, ((size_t)R + 1) * 4, st);
  cudaMemsetAsync(max_deg, 0, 4, st);
  cudaMemsetAsync(nchunk + N, 0, 4, st);
  cudaMemsetAsync(nsplit + N, 0, 4, st);
  cudaMemsetAsync(aflag + N, 0, 4, st);
  if (T > 0) {
    expand_edges_kernel<<<(T + TB - 1) / TB, TB, 0, st>>>(triples, T, R, src, dst, etype, indeg, eid, rk, N);
  }
  norm_chunks_kernel<<<(N + TB - 1) / TB, TB, 0, st>>>(indeg, N, norm, nchunk, nsplit, aflag, max_deg);
  // rowptr[0..N-1] = exclusive scan of indeg; rowptr[N] = E is written by write_counts_kernel
  // (every edge has exactly one destination).
  cub::DeviceScan::ExclusiveSum(cubtmp, cb, indeg, rowptr, N, st);
  cub::DeviceScan::ExclusiveSum(cubtmp, cb, nchunk, vptr, N + 1, st);
  cub::DeviceScan::ExclusiveSum(cubtmp, cb, nsplit, sptr, N + 1, st);
  cub::DeviceScan::ExclusiveSum(cubtmp, cb, aflag, ascan, N + 1, st);
  active_pos_kernel<<<(N + TB - 1) / TB, TB, 0, st>>>(indeg, ascan, N, active_pos, active_rows);
  fill_vrows_kernel<<<(N + TB - 1) / TB, TB, 0, st>>>(vptr, N, vrow_row);
  if (T > 0) {
    // stable sort of edge ids by destination: eperm[i] = original edge id of the i-th CSR slot
    int end_bit = 1; while ((1LL << end_bit) < (long long)N) ++end_bit;
    cub::DeviceRadixSort::SortPairs(cubtmp, cb, (const int*)dst, dst_alt, (const int*)eid, eperm, E, 0, end_bit, st);
    gather_sorted_kernel<<<(E + TB - 1) / TB, TB, 0, st>>>(eperm, src, etype, E, src_sorted, etype_sorted);
    // relation -> entity sets
    int rbits = 1; while ((1ULL << rbits) < (unsigned long long)R * (unsigned long long)N) ++rbits;
    cub::DeviceRadixSort::SortKeys(cubtmp, cb, (const unsigned long long*)rk, rk_alt, E, 0, rbits, st);
    rel_heads_kernel<<<(E + TB - 1) / TB, TB, 0, st>>>(rk_alt, E, N, head, rel_count);
    cub::DeviceScan::ExclusiveSum(cubtmp, cb, head, pos, E, st);
    rel_compact_kernel<<<(E + TB - 1) / TB, TB, 0, st>>>(rk_alt, head, pos, E, N, rel_ents);
  }
  cub::DeviceScan::ExclusiveSum(cubtmp, cb, rel_count, rel_rowptr, R + 1, st);
  write_counts_kernel<<<1, 1, 0, st>>>(vptr, sptr, N, rel_rowptr, R, max_deg, rowptr, E, ascan, counts);
  return check_launch("csr_build");
}


// =====================================================================================================================
// Small-snapshot path: ONE CTA builds the whole index of one snapshot, L snapshots per launch (grid = L).
// Real TKG snapshots have a few hundred to a few thousand triples (ICEWS14s ~250, ICEWS18 ~1.5k, GDELT ~0.75k): the
// ~30-launch CUB pipeline above costs ~160 us of pure launch latency per snapshot and the reference rebuilds L graphs
// for every evaluated timestamp (src/main.py:68,233).  Here the two sorts run as block radix sorts in shared memory,
// the four per-node scans as one fused block scan, and every snapshot of the history is built concurrently.
// Same outputs, bit for bit, as csr_build (tests compare the two).
// =====================================================================================================================
constexpr int kSmallThreads = 1024;
constexpr int kSmallBatch = 16;          // snapshots per launch (kernel-parameter space: 16 x 176 bytes)
constexpr int kSmallMaxRels = 8191;      // rel_count lives in shared memory

struct SmallSnap {
  const int64_t* triples;
  int* src; int* dst; int* etype; int* indeg; float* norm; int* rowptr; int* src_sorted; int* etype_sorted; int* eperm;
  int* vptr; int* sptr; int* vrow_row; int* active_pos; int* active_rows; int* rel_rowptr; int* rel_ents; int* counts;
  int T;
};
struct SmallBatch { SmallSnap g[kSmallBatch]; };

struct Quad { int a, b, c, d; };
struct QuadSum {
  __device__ __forceinline__ Quad operator()(const Quad& x, const Quad& y) const { return Quad{x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; }
};

template <int ITEMS>
struct SmallTypes {
  using SortKV = cub::BlockRadixSort<unsigned, kSmallThreads, ITEMS, int>;
  using SortK = cub::BlockRadixSort<unsigned, kSmallThreads, ITEMS>;
  using Scan4 = cub::BlockScan<Quad, kSmallThreads>;
  using ScanI = cub::BlockScan<int, kSmallThreads>;
  using RedI = cub::BlockReduce<int, kSmallThreads>;
  union Temp {
    typename SortKV::TempStorage kv;
    typename SortK::TempStorage k;
    typename Scan4::TempStorage s4;
    typename ScanI::TempStorage si;
    typename RedI::TempStorage ri;
  };
  static constexpr size_t temp_bytes = (sizeof(Temp) + 15) & ~(size_t)15;
  static constexpr size_t smem_bytes = temp_bytes + (size_t)kSmallThreads * ITEMS * 4 + ((size_t)kSmallMaxRels + 1) * 4 + 16;
};

template <int ITEMS>
__global__ void __launch_bounds__(kSmallThreads, 1)
csr_build_small_kernel(const __grid_constant__ SmallBatch batch, int N, int R) {
  using TY = SmallTypes<ITEMS>;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  typename TY::Temp& tmp = *reinterpret_cast<typename TY::Temp*>(sm_raw);
  unsigned* skeys = reinterpret_cast<unsigned*>(sm_raw + TY::temp_bytes);
  int* rel_count = reinterpret_cast<int*>(skeys + kSmallThreads * ITEMS);
  __shared__ Quad run;          // running prefix of the node scans
  __shared__ int s_maxdeg;

  const SmallSnap& g = batch.g[blockIdx.x];
  const int tid = threadIdx.x;
  const int T = g.T, E = 2 * T;

  // ---- 1. clear the counters ----
  for (int v = tid; v < N; v += kSmallThreads) g.indeg[v] = 0;
  for (int r = tid; r <= R; r += kSmallThreads) rel_count[r] = 0;
  if (tid == 0) { run = Quad{0, 0, 0, 0}; s_maxdeg = 0; }
  __syncthreads();

  // ---- 2. inverse-edge doubling (utils.py:116-118) + in-degrees ----
  for (int t = tid; t < T; t += kSmallThreads) {
    const int s = (int)g.triples[3 * (size_t)t + 0];
    const int r = (int)g.triples[3 * (size_t)t + 1];
    const int o = (int)g.triples[3 * (size_t)t + 2];
    g.src[t] = s; g.dst[t] = o; g.etype[t] = r;
    g.src[T + t] = o; g.dst[T + t] = s; g.etype[T + t] = r + R;
    atomicAdd(&g.indeg[o], 1);
    atomicAdd(&g.indeg[s], 1);
  }
  __syncthreads();

  // ---- 3. per-node pass: norm, rowptr / vptr / sptr / active_pos (four exclusive scans fused), virtual rows ----
  constexpr int NI = 4;
  int my_max = 0;
  for (int base = 0; base < N; base += kSmallThreads * NI) {
    Quad in[NI], out[NI];
    int deg[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int v = base + tid * NI + i;
      const int d = v < N ? __ldcg(g.indeg + v) : 0;
      deg[i] = d;
      const int nc = (d + kAggChunk - 1) / kAggChunk;
      in[i] = Quad{d, nc, nc > 1 ? nc : 0, d > 0 ? 1 : 0};
      if (d > kAggChunk) my_max = max(my_max, d);
    }
    Quad total;
    typename TY::Scan4(tmp.s4).ExclusiveScan(in, out, Quad{0, 0, 0, 0}, QuadSum(), total);
    const Quad pre = run;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int v = base + tid * NI + i;
      if (v < N) {
        const int d = deg[i];
        g.norm[v] = 1.0f / (float)(d == 0 ? 1 : d);       // comp_deg_norm, utils.py:110-114
        g.rowptr[v] = pre.a + out[i].a;
        const int vb = pre.b + out[i].b;
        g.vptr[v] = vb;
        g.sptr[v] = pre.c + out[i].c;
        g.active_pos[v] = d > 0 ? pre.d + out[i].d : -1;
        if (d > 0) g.active_rows[pre.d + out[i].d] = v;
        for (int k = 0; k < in[i].b; ++k) g.vrow_row[vb + k] = v;
      }
    }
    __syncthreads();
    if (tid == 0) run = Quad{pre.a + total.a, pre.b + total.b, pre.c + total.c, pre.d + total.d};
    __syncthreads();
  }
  if (my_max) atomicMax(&s_maxdeg, my_max);
  __syncthreads();
  if (tid == 0) {
    g.rowptr[N] = E;
    g.vptr[N] = run.b;
    g.sptr[N] = run.c;
    g.counts[0] = run.b; g.counts[1] = run.c; g.counts[3] = s_maxdeg; g.counts[4] = run.d;
    g.counts[5] = 0; g.counts[6] = 0; g.counts[7] = 0;
  }

  // ---- 4. stable sort of the edge ids by destination (blocked arrangement: item j of thread t is slot t*ITEMS+j) ----
  // padding slots carry the key 2^end_bit (above every real key), so only end_bit + 1 bits are sorted
  int end_bit = 1;
  while ((1LL << end_bit) < (long long)N) ++end_bit;
  {
    const unsigned pad = 1u << end_bit;
    unsigned keys[ITEMS];
    int vals[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int e = tid * ITEMS + i;
      vals[i] = e;
      keys[i] = e < E ? (unsigned)(e < T ? (int)g.triples[3 * (size_t)e + 2] : (int)g.triples[3 * (size_t)(e - T) + 0]) : pad;
    }
    typename TY::SortKV(tmp.kv).Sort(keys, vals, 0, end_bit + 1);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int pos = tid * ITEMS + i;
      if (pos < E) {
        const int e = vals[i];
        const size_t t3 = 3 * (size_t)(e < T ? e : e - T);
        const int s = (int)g.triples[t3 + 0], r = (int)g.triples[t3 + 1], o = (int)g.triples[t3 + 2];
        g.eperm[pos] = e;
        g.src_sorted[pos] = e < T ? s : o;
        g.etype_sorted[pos] = e < T ? r : r + R;
      }
    }
  }
  __syncthreads();

  // ---- 5. relation -> entity sets: sort (r, entity) keys, keep the first of every run ----
  {
    int rbits = 1;
    while ((1ULL << rbits) < (unsigned long long)R * (unsigned long long)N) ++rbits;
    const unsigned rpad = rbits >= 32 ? 0xffffffffu : 1u << rbits;     // small_ok() keeps R*N below 2^32 - 1
    unsigned keys[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int e = tid * ITEMS + i;
      if (e < E) {
        const size_t t3 = 3 * (size_t)(e < T ? e : e - T);
        const unsigned r = (unsigned)g.triples[t3 + 1];
        const unsigned ent = (unsigned)(e < T ? g.triples[t3 + 0] : g.triples[t3 + 2]);
        keys[i] = r * (unsigned)N + ent;
      } else {
        keys[i] = rpad;
      }
    }
    typename TY::SortK(tmp.k).Sort(keys, 0, rbits >= 32 ? 32 : rbits + 1);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) skeys[tid * ITEMS + i] = keys[i];
    __syncthreads();
    int heads = 0;
    bool head[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int pos = tid * ITEMS + i;
      head[i] = pos < E && (pos == 0 || skeys[pos - 1] != keys[i]);
      heads += head[i] ? 1 : 0;
    }
    int pos0, total;
    typename TY::ScanI(tmp.si).ExclusiveSum(heads, pos0, total);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (head[i]) {
        const unsigned r = keys[i] / (unsigned)N;
        g.rel_ents[pos0++] = (int)(keys[i] - r * (unsigned)N);
        atomicAdd(&rel_count[r], 1);
      }
    }
    if (tid == 0) g.counts[2] = total;
  }
  __syncthreads();
  // rel_rowptr = exclusive scan of rel_count over R + 1 entries
  {
    __shared__ int rrun;
    if (tid == 0) rrun = 0;
    __syncthreads();
    for (int base = 0; base <= R; base += kSmallThreads) {
      const int r = base + tid;
      const int c = r <= R ? rel_count[r] : 0;
      int ex, total;
      typename TY::ScanI(tmp.si).ExclusiveSum(c, ex, total);
      const int pre = rrun;
      if (r <= R) g.rel_rowptr[r] = pre + ex;
      __syncthreads();
      if (tid == 0) rrun = pre + total;
      __syncthreads();
    }
  }
}

template <int ITEMS>
static int launch_small(const SmallBatch& b, int n, int N, int R, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t ce = cudaFuncSetAttribute(csr_build_small_kernel<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)SmallTypes<ITEMS>::smem_bytes);
    if (ce != cudaSuccess) { set_last_error("csr_build_batch: cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce)); return (int)ce; }
    attr = true;
  }
  count_kernel_launch();
  csr_build_small_kernel<ITEMS><<<n, kSmallThreads, SmallTypes<ITEMS>::smem_bytes, st>>>(b, N, R);
  return check_launch("csr_build_batch");
}

static bool small_ok(int T, int N, int R) {
  // the per-node pass walks N with one CTA: beyond ~256k entities the multi-CTA pipeline wins again
  return 2 * (long long)T <= (long long)kSmallThreads * 16 && R <= kSmallMaxRels && N <= 262144 &&
         (unsigned long long)R * (unsigned long long)N < 0xffffffffULL;
}

size_t csr_build_batch_workspace_bytes(const int* T, int L, int N, int R) {
  size_t need = 0;
  for (int i = 0; i < L; ++i)
    if (!small_ok(T[i], N, R)) { size_t b = csr_build_workspace_bytes(T[i], N, R); need = b > need ? b : need; }
  return need;
}

// Builds the index of L snapshots: small ones (E <= 16384 edges) by the one-CTA-per-snapshot kernel, all of them
// concurrently; larger ones by the CUB pipeline, one after the other on the same stream (shared workspace).
int csr_build_batch(const regcn_csr_arrays* snaps, int L, int N, int R, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (L < 0 || N <= 0 || R <= 0) { set_last_error("csr_build_batch: bad dims L=%d N=%d R=%d", L, N, R); return REGCN_ERR_DIM; }
  if (L > 0 && !snaps) { set_last_error("csr_build_batch: null pointer"); return REGCN_ERR_NULL; }
  for (int i = 0; i < L; ++i) {
    const regcn_csr_arrays& a = snaps[i];
    if (a.T < 0) { set_last_error("csr_build_batch: snapshot %d has T=%d", i, a.T); return REGCN_ERR_DIM; }
    if (!a.src || !a.dst || !a.etype || !a.indeg || !a.norm || !a.rowptr || !a.src_sorted || !a.etype_sorted || !a.eperm ||
        !a.vptr || !a.sptr || !a.vrow_row || !a.active_pos || !a.active_rows || !a.rel_rowptr || !a.rel_ents || !a.counts || (a.T > 0 && !a.triples)) {
      set_last_error("csr_build_batch: null pointer in snapshot %d", i); return REGCN_ERR_NULL;
    }
  }
  // group the small snapshots by sort width so that one launch serves each group
  const int widths[3] = {4, 8, 16};
  for (int wi = 0; wi < 3; ++wi) {
    const long long cap = (long long)kSmallThreads * widths[wi];
    const long long lo = wi == 0 ? -1 : (long long)kSmallThreads * widths[wi - 1];
    SmallBatch b;
    int n = 0;
    for (int i = 0; i <= L; ++i) {
      bool take = false;
      if (i < L) {
        const long long E = 2LL * snaps[i].T;
        take = small_ok(snaps[i].T, N, R) && E <= cap && E > lo;
      }
      if (take) {
        const regcn_csr_arrays& a = snaps[i];
        SmallSnap& s = b.g[n++];
        s.triples = a.triples; s.src = a.src; s.dst = a.dst; s.etype = a.etype; s.indeg = a.indeg; s.norm = a.norm;
        s.rowptr = a.rowptr; s.src_sorted = a.src_sorted; s.etype_sorted = a.etype_sorted; s.eperm = a.eperm;
        s.vptr = a.vptr; s.sptr = a.sptr; s.vrow_row = a.vrow_row; s.active_pos = a.active_pos; s.active_rows = a.active_rows;
        s.rel_rowptr = a.rel_rowptr; s.rel_ents = a.rel_ents; s.counts = a.counts; s.T = a.T;
      }
      if (n == kSmallBatch || (i == L && n > 0)) {
        int e = wi == 0 ? launch_small<4>(b, n, N, R, st) : wi == 1 ? launch_small<8>(b, n, N, R, st) : launch_small<16>(b, n, N, R, st);
        if (e) return e;
        n = 0;
      }
    }
  }
  for (int i = 0; i < L; ++i) {
    const regcn_csr_arrays& a = snaps[i];
    if (small_ok(a.T, N, R)) continue;
    int e = csr_build(a.triples, a.T, N, R, a.src, a.dst, a.etype, a.indeg, a.norm, a.rowptr, a.src_sorted, a.etype_sorted,
                      a.eperm, a.vptr, a.sptr, a.vrow_row, a.active_pos, a.active_rows, a.rel_rowptr, a.rel_ents, a.counts, ws,
                      ws_bytes, st);
    if (e) return e;
  }
  return REGCN_OK;
}

// =====================================================================================================================
// Index concatenation: the indices of G snapshots over (N entities, R relations) become ONE index over G*N entities and
// G*R relations -- the block-diagonal union graph.  The windows of consecutive test timestamps are independent
// (src/main.py:60-90: every timestamp re-runs the recurrence over its own history), so G of them can be evolved by the
// same kernels with G times the rows per launch; step i of that batched recurrence needs the union of G snapshots, and
// every one of them is already indexed (SnapshotCache).  Member g's entity v becomes g*N + v, its relation r < R becomes
// g*R + r and the inverse relation R + r becomes G*R + g*R + r; CSR slots, virtual rows, split chunks, active rows and
// relation->entity lists keep their member-local order and are shifted by the totals of the members before them.
// One launch; pure copies with offsets (a few hundred KB per member).
// =====================================================================================================================
constexpr int kConcatMax = 32;     // (ConcatArgs = 5.8 KB of kernel parameters: CUDA >= 12.1 / sm_70+ take up to 32 KB)

struct ConcatMember {
  const int* src; const int* dst; const int* etype; const int* indeg; const float* norm; const int* rowptr;
  const int* src_sorted; const int* etype_sorted; const int* eperm; const int* vptr; const int* sptr; const int* vrow_row;
  const int* active_pos; const int* active_rows; const int* rel_rowptr; const int* rel_ents; const int* counts;
  int E, n_vrows, n_split, n_rel_ents, n_active;          // member sizes
  int eoff, voff, soff, roff, aoff;                       // totals of the members before this one
};
struct ConcatArgs {
  ConcatMember m[kConcatMax];
  int* src; int* dst; int* etype; int* indeg; float* norm; int* rowptr; int* src_sorted; int* etype_sorted; int* eperm;
  int* vptr; int* sptr; int* vrow_row; int* active_pos; int* active_rows; int* rel_rowptr; int* rel_ents; int* counts;
  int G, N, R;
};

__global__ void __launch_bounds__(256) csr_concat_kernel(const __grid_constant__ ConcatArgs a) {
  pdl_grid_sync();
  const int g = blockIdx.y;
  const ConcatMember& m = a.m[g];
  const int N = a.N, R = a.R, G = a.G;
  const int noff = g * N, reloff = g * R, GR = G * R;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const bool last = g == G - 1;
  for (int v = tid; v < N + (last ? 1 : 0); v += nth) {          // per-entity arrays (+ the closing pointer entries)
    a.rowptr[noff + v] = m.rowptr[v] + m.eoff;
    a.vptr[noff + v] = m.vptr[v] + m.voff;
    a.sptr[noff + v] = m.sptr[v] + m.soff;
    if (v < N) {
      a.indeg[noff + v] = m.indeg[v];
      a.norm[noff + v] = m.norm[v];
      const int ap = m.active_pos[v];
      a.active_pos[noff + v] = ap < 0 ? -1 : ap + m.aoff;
    }
  }
  for (int j = tid; j < m.E; j += nth) {                          // per-edge arrays
    const int o = m.eoff + j;
    const int t0 = m.etype[j], t1 = m.etype_sorted[j];
    a.src[o] = m.src[j] + noff;
    a.dst[o] = m.dst[j] + noff;
    a.etype[o] = t0 < R ? t0 + reloff : t0 - R + GR + reloff;
    a.src_sorted[o] = m.src_sorted[j] + noff;
    a.etype_sorted[o] = t1 < R ? t1 + reloff : t1 - R + GR + reloff;
    a.eperm[o] = m.eperm[j] + m.eoff;
  }
  for (int k = tid; k < m.n_vrows; k += nth) a.vrow_row[m.voff + k] = m.vrow_row[k] + noff;
  for (int k = tid; k < m.n_active; k += nth) a.active_rows[m.aoff + k] = m.active_rows[k] + noff;
  for (int r = tid; r < R + (last ? 1 : 0); r += nth) a.rel_rowptr[reloff + r] = m.rel_rowptr[r] + m.roff;
  for (int k = tid; k < m.n_rel_ents; k += nth) a.rel_ents[m.roff + k] = m.rel_ents[k] + noff;
  if (g == 0 && tid == 0) {
    int maxdeg = 0;
    for (int i = 0; i < G; ++i) maxdeg = max(maxdeg, a.m[i].counts[3]);
    const ConcatMember& z = a.m[G - 1];
    a.counts[0] = z.voff + z.n_vrows; a.counts[1] = z.soff + z.n_split; a.counts[2] = z.roff + z.n_rel_ents;
    a.counts[3] = maxdeg; a.counts[4] = z.aoff + z.n_active; a.counts[5] = 0; a.counts[6] = 0; a.counts[7] = 0;
  }
}

// sizes: (G,4) host ints per member: n_vrows, n_split_chunks, n_rel_ents, n_active (the counters regcn_csr_build reports)
int csr_concat(const regcn_csr_arrays* members, const int* sizes, int G, int N, int R, const regcn_csr_arrays* out,
               cudaStream_t st) {
  if (G <= 0 || G > kConcatMax || N <= 0 || R <= 0) { set_last_error("csr_concat: bad dims G=%d (max %d) N=%d R=%d", G, kConcatMax, N, R); return REGCN_ERR_DIM; }
  if (!members || !sizes || !out) { set_last_error("csr_concat: null pointer"); return REGCN_ERR_NULL; }
  if ((long long)G * N > 0x7fffffffLL || (long long)G * R * 2 > 0x7fffffffLL) { set_last_error("csr_concat: G*N overflows int32"); return REGCN_ERR_DIM; }
  ConcatArgs a;
  long long eo = 0, vo = 0, so = 0, ro = 0, ao = 0;
  for (int g = 0; g < G; ++g) {
    const regcn_csr_arrays& s = members[g];
    if (s.T < 0 || !s.src || !s.dst || !s.etype || !s.indeg || !s.norm || !s.rowptr || !s.src_sorted || !s.etype_sorted || !s.eperm ||
        !s.vptr || !s.sptr || !s.vrow_row || !s.active_pos || !s.active_rows || !s.rel_rowptr || !s.rel_ents || !s.counts) {
      set_last_error("csr_concat: null pointer / bad T in member %d", g); return REGCN_ERR_NULL;
    }
    ConcatMember& m = a.m[g];
    m.src = s.src; m.dst = s.dst; m.etype = s.etype; m.indeg = s.indeg; m.norm = s.norm; m.rowptr = s.rowptr;
    m.src_sorted = s.src_sorted; m.etype_sorted = s.etype_sorted; m.eperm = s.eperm; m.vptr = s.vptr; m.sptr = s.sptr;
    m.vrow_row = s.vrow_row; m.active_pos = s.active_pos; m.active_rows = s.active_rows; m.rel_rowptr = s.rel_rowptr;
    m.rel_ents = s.rel_ents; m.counts = s.counts;
    m.E = 2 * s.T; m.n_vrows = sizes[4 * g]; m.n_split = sizes[4 * g + 1]; m.n_rel_ents = sizes[4 * g + 2]; m.n_active = sizes[4 * g + 3];
    if (m.n_vrows < 0 || m.n_split < 0 || m.n_rel_ents < 0 || m.n_active < 0 || m.n_active > N) { set_last_error("csr_concat: bad sizes of member %d", g); return REGCN_ERR_DIM; }
    m.eoff = (int)eo; m.voff = (int)vo; m.soff = (int)so; m.roff = (int)ro; m.aoff = (int)ao;
    eo += m.E; vo += m.n_vrows; so += m.n_split; ro += m.n_rel_ents; ao += m.n_active;
    if (eo > 0x7fffffffLL || ro > 0x7fffffffLL) { set_last_error("csr_concat: edge total overflows int32"); return REGCN_ERR_DIM; }
  }
  const regcn_csr_arrays& o = *out;
  if (!o.src || !o.dst || !o.etype || !o.indeg || !o.norm || !o.rowptr || !o.src_sorted || !o.etype_sorted || !o.eperm || !o.vptr ||
      !o.sptr || !o.vrow_row || !o.active_pos || !o.active_rows || !o.rel_rowptr || !o.rel_ents || !o.counts) {
    set_last_error("csr_concat: null pointer in the output index"); return REGCN_ERR_NULL;
  }
  a.src = o.src; a.dst = o.dst; a.etype = o.etype; a.indeg = o.indeg; a.norm = o.norm; a.rowptr = o.rowptr;
  a.src_sorted = o.src_sorted; a.etype_sorted = o.etype_sorted; a.eperm = o.eperm; a.vptr = o.vptr; a.sptr = o.sptr;
  a.vrow_row = o.vrow_row; a.active_pos = o.active_pos; a.active_rows = o.active_rows; a.rel_rowptr = o.rel_rowptr;
  a.rel_ents = o.rel_ents; a.counts = o.counts;
  a.G = G; a.N = N; a.R = R;
  const int blocks = std::max(1, std::min(64, (N + 255) / 256));
  launch_k(csr_concat_kernel, dim3(blocks, G), dim3(256), 0, st, a);
  return check_launch("csr_concat");
}

}  // namespace regcn
