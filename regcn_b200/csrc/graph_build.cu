// K1: snapshot edge-index build on the device.
// Restates rgcn/utils.py:100-134 (build_sub_graph: inverse-edge doubling, in-degree,
// norm = 1/max(indeg,1)) and rgcn/utils.py:78-97 (r2e: per-relation entity sets) as a
// CSR-by-destination index plus a relation->entity CSR.  Integer work: bit-exact.
#include "common.cuh"
#include <cub/cub.cuh>

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
                                  int* __restrict__ active_pos) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  active_pos[v] = indeg[v] > 0 ? scan[v] : -1;
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
              int* vptr, int* sptr, int* vrow_row, int* active_pos,
              int* rel_rowptr, int* rel_ents, int* counts,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  if (T < 0 || N <= 0 || R <= 0) { set_last_error("csr_build: bad dims T=%d N=%d R=%d", T, N, R); return REGCN_ERR_DIM; }
  BuildWs w = plan_build_ws(T, N, R);
  if (ws_bytes < w.total) { set_last_error("csr_build: workspace %zu < %zu", ws_bytes, w.total); return REGCN_ERR_WORKSPACE; }
  if (!src || !dst || !etype || !indeg || !norm || !rowptr || !src_sorted || !etype_sorted || !eperm || !vptr ||
      !sptr || !vrow_row || !active_pos || !rel_rowptr || !rel_ents || !counts || !ws || (T > 0 && !triples)) {
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
  active_pos_kernel<<<(N + TB - 1) / TB, TB, 0, st>>>(indeg, ascan, N, active_pos);
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

}  // namespace regcn
