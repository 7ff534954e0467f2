// Stream-ordered orchestration of one whole history recurrence in a single C-ABI call.
//
// The per-snapshot loop of RecurrentRGCN.forward (src/rrgcn.py:159-179) is ~13 kernels on tiny-to-medium
// operands; driven one launch at a time from Python it is launch-bound.  Here the host side of the loop is C++:
// every kernel of the L-snapshot recurrence is enqueued back to back on the caller's stream, all intermediates
// live in one caller-provided workspace, and every producer kernel emits the (hi, lo) TF32 split its consumer
// GEMM needs, so no stand-alone conversion pass runs inside the loop.
//
// Per snapshot (Euclidean RE-GCN, uvrgcn encoder, no skip connection):
//   K2  rel_mean_pool(h)                       -> x_mean (hi, lo)
//   GEMM gi = x_mean . W_ih[:, d:]^T + gi_static        gi_static = emb_rel . W_ih[:, :d]^T + b_ih (constant per model)
//   GEMM gh = h0 . W_hh^T + b_hh
//   K3  gru_gate                               -> h0 (raw, hi, lo)
//   per layer l:  K4 aggregate(x, h0)          -> agg (hi, lo)
//                 GEMM Lm = x . [W_loop | W_evolve (| W_time for l = 0)]
//                 GEMM P  = agg . W_n
//                 K5 combine(P, Lm, indeg)     -> x' (raw, hi, lo)
//   K9  time_gate(Lm0[:, 2d:3d], cur, h)       -> h (raw into hist[i], hi, lo)
#include "../../include/regcn_b200.h"
#include "common.cuh"
#include "internal.h"
#include <stdlib.h>
#include <mutex>

namespace regcn {

static inline size_t al(size_t n) { return (n + 63) & ~(size_t)63; }   // 256-byte aligned float counts

// Side streams of the evolve schedule (library-owned, per device, created on first use).  The caller's stream carries
// the large all-entity GEMMs of a snapshot; side stream B carries the chain of small kernels (relation mean-pool, GRU,
// aggregates, active-row GEMMs) that would otherwise sit between them, each costing a launch ramp and a memory-latency
// chain on an otherwise idle machine; side stream C runs the one GEMM of that chain that only depends on the PREVIOUS
// snapshot (gh = h0 . W_hh^T) ahead of time.
// The streams and their events are shared by every call on the device, so an evolve call holds the device's mutex
// from its first fork to its last join: two host threads (or two caller streams) cannot interleave their event
// records.  StreamScope joins whatever is still outstanding on the side streams back into the caller's stream on
// EVERY exit path, and puts the thread-local launch knobs (PDL suppression, SM hint) back.
struct AuxStream {
  cudaStream_t sb = nullptr, sc = nullptr;
  cudaEvent_t fork = nullptr, b_done = nullptr, c_done = nullptr, h0_ready = nullptr;
  bool ok = false, failed = false;
  std::mutex mu;
};
static AuxStream* aux_stream() {
  static AuxStream aux[16];
  static std::mutex init_mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  AuxStream& a = aux[dev];
  std::lock_guard<std::mutex> g(init_mu);
  if (!a.ok && !a.failed) {
    bool good = cudaStreamCreateWithFlags(&a.sb, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&a.sc, cudaStreamNonBlocking) == cudaSuccess;
    cudaEvent_t* evs[4] = {&a.fork, &a.b_done, &a.c_done, &a.h0_ready};
    for (int i = 0; good && i < 4; ++i) good = cudaEventCreateWithFlags(evs[i], cudaEventDisableTiming) == cudaSuccess;
    if (good) a.ok = true; else { a.failed = true; cudaGetLastError(); }
  }
  return a.ok ? &a : nullptr;
}
// orders `waiter` after everything enqueued on `src` so far; false (and a cleared error) when the event API fails
static bool stream_after(cudaStream_t waiter, cudaStream_t src, cudaEvent_t ev) {
  if (cudaEventRecord(ev, src) != cudaSuccess || cudaStreamWaitEvent(waiter, ev, 0) != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}
void gemm_tf32_sm_hint(int sms);
void gemm_tf32_grid_cap(int ctas);
struct StreamScope {
  AuxStream* aux;
  cudaStream_t st;
  bool b_open = false, c_open = false;     // work enqueued on a side stream that the caller's stream has not waited for
  StreamScope(AuxStream* a, cudaStream_t s) : aux(a), st(s) { if (aux) aux->mu.lock(); }
  void join() {
    if (aux && b_open) { if (!stream_after(st, aux->sb, aux->b_done)) cudaStreamSynchronize(aux->sb); b_open = false; }
    if (aux && c_open) { if (!stream_after(st, aux->sc, aux->c_done)) cudaStreamSynchronize(aux->sc); c_open = false; }
  }
  ~StreamScope() {
    join();
    pdl_suppress(false);
    gemm_tf32_sm_hint(0);
    gemm_tf32_grid_cap(0);
    if (aux) aux->mu.unlock();
  }
};
// SMs a persistent all-entity GEMM of `ncol` output columns leaves free (see the balanced grid in gemm_tc.cu)
static int side_hint(int N, int d, int ncol, int sms) {
  (void)d;
  const long long t = (long long)((N + 127) / 128) * ((ncol + 207) / 208);
  const long long rounds = (t + sms - 1) / sms;
  const int grid_a = (int)((t + rounds - 1) / (rounds > 0 ? rounds : 1));
  return sms - grid_a >= 16 ? sms - grid_a : 16;
}
// Programmatic dependent launch on side stream B: off by default.  A side kernel launched early parks its CTAs (the
// aggregates: ~200 CTAs, one or two per SM) at their dependency wait on exactly the SMs the next all-entity GEMM of the
// caller's stream needs -- a 384-thread GEMM CTA takes a whole SM's registers -- and that GEMM then runs its tiles in
// two waves (measured at C3, B200: 1.13 ms per step with plain stream order on the side stream, 1.18 ms with PDL).
static bool side_pdl_keep() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("REGCN_SIDE_PDL"); v = (e && e[0] == '1') ? 1 : 0; }
  return v != 0;
}
// fp32-A dataflow for the all-entity GEMMs of the sparse-snapshot form (see regcn_regcn_evolve): on from 64 k entity rows,
// REGCN_EVOLVE_A32=0 / 1 or regcn_evolve_a32_mode(0 / 1) force it off / on, -1 = by size
// SMs the side stream gets in the many-round regime (REGCN_SIDE_SMS; 0 = no split).  Measured at 8 ICEWS18-shaped windows
// per recurrence (profiles/timeline.py c3x8, us per batched recurrence): 0: 3155, 12: 4250, 16: 3942, 24: 3238, 32: 3007,
// 40: 3041, 48: 3167, 64: 3576 -- the side chain is latency-bound and needs ~a fifth of the machine to keep pace
static int side_sm_share() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("REGCN_SIDE_SMS"); v = e ? atoi(e) : 32; if (v < 0 || v > 64) v = 32; }
  return v;
}
static int g_evolve_a32 = -2;          // -2: read REGCN_EVOLVE_A32 on first use; -1 auto, 0 off, 1 on
void evolve_a32_set(int mode) { g_evolve_a32 = mode < 0 ? -1 : (mode ? 1 : 0); }
static bool evolve_a32(int N) {
  if (g_evolve_a32 == -2) { const char* e = getenv("REGCN_EVOLVE_A32"); g_evolve_a32 = e ? (e[0] == '1' ? 1 : e[0] == '0' ? 0 : -1) : -1; }
  return g_evolve_a32 >= 0 ? g_evolve_a32 != 0 : N >= 65536;
}
static int g_two_stream = -1;
void two_stream_set(int on) { g_two_stream = on ? 1 : 0; }
static bool two_stream_enabled() {
  if (g_two_stream < 0) {
    const char* e = getenv("REGCN_TWO_STREAM");
    g_two_stream = (e && e[0] == '0') ? 0 : 1;
  }
  return g_two_stream != 0;
}

struct EvolveWs {
  size_t xm_hi, xm_lo, gi, gh, h0_hi, h0_lo, agg_hi, agg_lo, Lm, L2, P, set[2][3], h_hi, h_lo, h_init, partial, rel_partial, fold, fold_n, total;
  // agg_hi/agg_lo double as the compact [agg | h] operand of the sparse-snapshot path (n_active <= N/2 rows of 2d)
};

static EvolveWs plan_evolve(int N, int R2, int d, int max_split_chunks, int rel_nsplit) {
  EvolveWs w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += al(n); return o; };
  const size_t nd = (size_t)N * d, rd = (size_t)R2 * d;
  w.xm_hi = take(rd); w.xm_lo = take(rd);
  w.gi = take(rd * 3); w.gh = take(rd * 3);
  w.h0_hi = take(rd); w.h0_lo = take(rd);
  w.agg_hi = take(nd); w.agg_lo = take(nd);
  w.Lm = take(nd * 3); w.L2 = take(nd * 2); w.P = take(nd);
  for (int s = 0; s < 2; ++s) for (int k = 0; k < 3; ++k) w.set[s][k] = take(nd);
  w.h_hi = take(nd); w.h_lo = take(nd); w.h_init = take(nd);
  w.partial = take((size_t)(max_split_chunks > 0 ? max_split_chunks : 1) * (d + 1));
  w.rel_partial = take(rel_nsplit > 1 ? (size_t)(R2 / 2) * rel_nsplit * d : 1);
  w.fold_n = (size_t)(max_split_chunks > 0 ? max_split_chunks : 1);      // arrival counters of the in-kernel fold (int32)
  w.fold = take(w.fold_n);
  w.total = off * sizeof(float);
  return w;
}

}  // namespace regcn

using namespace regcn;

extern "C" {

size_t regcn_regcn_evolve_workspace_bytes(int N, int R2, int d, int max_split_chunks, int rel_nsplit) {
  return plan_evolve(N, R2, d, max_split_chunks, rel_nsplit).total;
}

int regcn_regcn_evolve(const void* const* mp, const int* mi, const void* const* gp, const int* gi_, int L,
                       float* hist, float* h0_out, int rel_nsplit, void* workspace, size_t workspace_bytes,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!mp || !mi || !gp || !gi_ || !hist || !h0_out || !workspace) { set_last_error("regcn_evolve: null pointer"); return REGCN_ERR_NULL; }
  const int N = mi[RMI_NUM_ENTS], R2 = mi[RMI_NUM_RELS2], d = mi[RMI_DIM], nl = mi[RMI_NUM_LAYERS];
  const int layer_norm = mi[RMI_LAYER_NORM], self_loop = mi[RMI_SELF_LOOP];
  if (N <= 0 || R2 <= 0 || (R2 & 1) || d <= 0 || (d & 3) || d > 256 || nl < 1 || nl > 8 || L < 0) {
    set_last_error("regcn_evolve: bad model dims N=%d R2=%d d=%d layers=%d", N, R2, d, nl); return REGCN_ERR_DIM;
  }
  if (!self_loop) { set_last_error("regcn_evolve: self_loop=False is served by the layer-level path"); return REGCN_ERR_UNSUPPORTED; }
  int max_split = 0;
  for (int i = 0; i < L; ++i) max_split = gi_[i * RGI_NUM_INTS + RGI_N_SPLIT_CHUNKS] > max_split ? gi_[i * RGI_NUM_INTS + RGI_N_SPLIT_CHUNKS] : max_split;
  if (rel_nsplit < 1) rel_nsplit = 1;
  const EvolveWs w = plan_evolve(N, R2, d, max_split, rel_nsplit);
  if (workspace_bytes < w.total) { set_last_error("regcn_evolve: workspace %zu < %zu", workspace_bytes, w.total); return REGCN_ERR_WORKSPACE; }
  float* ws = (float*)workspace;
  auto F = [&](int k) { return (const float*)mp[k]; };
  const size_t nd = (size_t)N * d;
  int e;

  // ---- initial entity state: F.normalize(dynamic_emb) if layer_norm (src/rrgcn.py:154) ----
  const float* h_raw;
  if (layer_norm) {
    if ((e = row_map(F(RM_DYNAMIC_EMB), ws + w.h_init, N, d, 0, 1.0, nullptr, ws + w.h_hi, ws + w.h_lo, st))) return e;
    h_raw = ws + w.h_init;
  } else {
    if ((e = split_tf32(F(RM_DYNAMIC_EMB), ws + w.h_hi, ws + w.h_lo, nd, st))) return e;
    h_raw = F(RM_DYNAMIC_EMB);
  }
  const float* h0_raw = F(RM_EMB_REL);
  const float* h0_hi = F(RM_EMB_REL_HI);
  const float* h0_lo = F(RM_EMB_REL_LO);
  int* fold = reinterpret_cast<int*>(ws + w.fold);
  if (max_split > 0 && cudaMemsetAsync(fold, 0, w.fold_n * sizeof(int), st) != cudaSuccess) { cudaGetLastError(); fold = nullptr; }
  if (max_split <= 0) fold = nullptr;
  // Large tables (a recurrence batched over several history windows, or a big graph): the all-entity GEMMs are
  // HBM-bound, and two thirds of their bytes are the (hi, lo) copies of activations.  There the entity state stays ONE
  // fp32 copy and the GEMM's converter warps split it on chip (gemm_tf32_layer_a32): 7 instead of 12 row-sized
  // transfers per snapshot.  Small tables keep the pre-split operands (the TMA-fed main loop has the lower latency).
  const bool a32 = evolve_a32(N);
  StreamScope scope(two_stream_enabled() ? aux_stream() : nullptr, st);
  AuxStream* aux = scope.aux;
  bool gh_ahead = false;        // gh of the current snapshot was already enqueued on side stream C
  int sm_count = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }

  // Many-round regime (a batched recurrence: >= 4 tiles per SM in the all-entity GEMMs): the SMs are split between the two
  // streams -- the all-entity GEMMs get sm_count - side_sms persistent CTAs, the side stream's GEMMs side_sms, and the
  // side stream's other kernels run on whatever the capped grids leave free.  Small tables keep the balanced whole-machine
  // grids (their one- or two-round GEMMs leave SMs free by themselves).
  const int side_sms = side_sm_share();
  const bool split_sms = side_sms > 0 && (long long)((N + 127) / 128) >= 4LL * sm_count;
  auto main_knobs = [&]() { gemm_tf32_sm_hint(0); gemm_tf32_grid_cap(split_sms ? sm_count - side_sms : 0); };
  auto side_knobs = [&](int ncol) {
    gemm_tf32_sm_hint(split_sms ? side_sms : side_hint(N, d, ncol, sm_count));
    gemm_tf32_grid_cap(split_sms ? side_sms : 0);
  };
  for (int i = 0; i < L; ++i) {
    const void* const* g = gp + (size_t)i * RG_NUM_PTRS;
    const int* gn = gi_ + (size_t)i * RGI_NUM_INTS;
    auto GI = [&](int k) { return (const int*)g[k]; };
    // ---- schedule of this snapshot ----
    // Two-stream form (sparse snapshot, >= 2 layers, a table large enough for the all-entity GEMMs to matter): the
    // caller's stream runs the all-entity GEMMs  x . [W_evolve | W_time]  of every layer back to back -- they depend
    // on nothing but the previous snapshot -- while the side stream runs the chain they do not depend on: relation
    // mean-pool -> GRU GEMMs -> gate -> per layer {aggregate, fix-up, active-row GEMM}.  The persistent GEMMs leave
    // whole SMs free (balanced grids), which is where the side stream's kernels run.
    bool side_pdl_off = false;
    const int n_active_ = gn[RGI_N_ACTIVE];
    bool two = aux && n_active_ * 2 <= N && nl >= 2 && N >= 4096;
    if (two) {
      // everything the side streams read (h of the previous snapshot) is ordered before this fork
      two = stream_after(aux->sb, st, aux->fork);
      if (!two) scope.join();
    }
    cudaStream_t sB = two ? aux->sb : st;
    if (two) {
      scope.b_open = true;
      side_pdl_off = !side_pdl_keep();
      side_knobs(2 * d);
      // kernels launched early park their CTAs on the SMs the other stream needs: the side stream stays plainly ordered
      pdl_suppress(side_pdl_off);
    }
    // ---- relation evolution (K2, K3) ----
    // gh = h0 . W_hh^T + b_hh depends on the previous snapshot only: in the two-stream schedule it was enqueued on
    // side stream C right after the previous snapshot's GRU gate (below) and is joined in front of this one's
    if (!gh_ahead) {
      cudaStream_t sG = sB;
      if (two && cudaStreamWaitEvent(aux->sc, aux->fork, 0) == cudaSuccess) { sG = aux->sc; scope.c_open = true; }
      if ((e = gemm_tf32(h0_hi, h0_lo, d, F(RM_WHH_HI), F(RM_WHH_LO), d, ws + w.gh, 3 * d, R2, 3 * d, d, F(RM_B_HH), 0, 3, 1,
                         nullptr, 0, nullptr, 0, sG))) return e;
      gh_ahead = sG != sB;
    }
    if ((e = rel_mean_pool(h_raw, GI(RG_REL_ROWPTR), GI(RG_REL_ENTS), R2 / 2, d, rel_nsplit, nullptr, ws + w.rel_partial,
                           ws + w.xm_hi, ws + w.xm_lo, sB))) return e;
    if ((e = gemm_tf32(ws + w.xm_hi, ws + w.xm_lo, d, F(RM_WIH_R_HI), F(RM_WIH_R_LO), d, ws + w.gi, 3 * d, R2, 3 * d, d,
                       nullptr, 0, 3, 1, nullptr, 0, F(RM_GI_STATIC), 3 * d, sB))) return e;
    if (gh_ahead) {
      if (!stream_after(sB, aux->sc, aux->c_done)) cudaStreamSynchronize(aux->sc);
      scope.c_open = false;
      gh_ahead = false;
    }
    if ((e = gru_gate(ws + w.gi, ws + w.gh, h0_raw, h0_out, R2, d, layer_norm, ws + w.h0_hi, ws + w.h0_lo, sB))) return e;
    h0_raw = h0_out; h0_hi = ws + w.h0_hi; h0_lo = ws + w.h0_lo;
    if (two && i + 1 < L) {
      // next snapshot's gh: needs nothing but the relation state just written
      const int n_next = gi_[(size_t)(i + 1) * RGI_NUM_INTS + RGI_N_ACTIVE];
      if (n_next * 2 <= N && stream_after(aux->sc, sB, aux->h0_ready)) {
        scope.c_open = true;
        if ((e = gemm_tf32(h0_hi, h0_lo, d, F(RM_WHH_HI), F(RM_WHH_LO), d, ws + w.gh, 3 * d, R2, 3 * d, d, F(RM_B_HH), 0, 3,
                           1, nullptr, 0, nullptr, 0, aux->sc))) return e;
        gh_ahead = true;
      }
    }
    // ---- entity evolution: n_layers x UnionRGCNLayer (K4, GEMMs, K5) ----
    const float* x_raw = h_raw;
    const float* x_hi = ws + w.h_hi;
    const float* x_lo = ws + w.h_lo;
    // Sparse snapshots (real TKG data: ~10% of the entities have in-edges) use the row-partitioned form
    //   active rows   : rrelu([agg | x] . [W_n ; W_loop])     one K = 2d GEMM over the n_active compact rows
    //   inactive rows : rrelu(x . W_evolve)                    one K = d GEMM over all rows (also yields the gate)
    // dense snapshots keep  rrelu(agg . W_n + where(indeg>0, x.W_loop, x.W_evolve)).
    const int n_active = gn[RGI_N_ACTIVE];
    const bool sparse = n_active * 2 <= N;
    const float* gate_G = nullptr;
    int gate_ld = 0;
    bool gate_done = false;
    for (int l = 0; l < nl; ++l) {
      const int base = RM_LAYER0 + RM_LAYER_STRIDE * l;
      const bool last = l == nl - 1;
      float* o_raw = ws + w.set[l & 1][0];
      float* o_hi = last ? nullptr : ws + w.set[l & 1][1];
      float* o_lo = last ? nullptr : ws + w.set[l & 1][2];
      if (sparse && nl >= 2) {
        // Row-partitioned layer with the elementwise tail fused into the GEMM epilogues (gemm_tf32_layer):
        //   inactive rows : x' = rrelu(x . W_evolve) written as the TF32 split only (nobody reads their fp32 rows: every
        //                   edge source is also a destination, so the next aggregate gathers active rows only);
        //                   layer 0 also emits the gate pre-activations x . W_time, the last layer applies the time gate
        //   active rows   : x' = rrelu([agg | x] . [W_n ; W_loop]) scattered back through active_rows
        const int ncol = (l == 0 ? 2 : 1) * d;           // [W_evolve (| W_time)]
        const int* arows = GI(RG_ACTIVE_ROWS);
        if ((e = union_aggregate(x_raw, h0_raw, GI(RG_ROWPTR), GI(RG_SRC_SORTED), GI(RG_ETYPE_SORTED), (const float*)g[RG_NORM],
                                 GI(RG_VPTR), GI(RG_SPTR), GI(RG_VROW_ROW), gn[RGI_N_VROWS], gn[RGI_N_SPLIT_CHUNKS], nullptr,
                                 0.f, N, d, nullptr, ws + w.partial, ws + w.agg_hi, ws + w.agg_lo, GI(RG_ACTIVE_POS), 2 * d,
                                 gn[RGI_MAX_CHUNKS], sB, fold))) return e;
        // a32: the inactive rows' input is the fp32 state itself (split on chip), their output one fp32 copy; the split
        // copies of the entity state are then never read.  The NEXT snapshot decides whether it needs them (it may be a
        // dense one), so the last layer still writes them unless that snapshot takes this form too.
        const bool next_a32 = a32 && (i + 1 >= L || (gi_[(size_t)(i + 1) * RGI_NUM_INTS + RGI_N_ACTIVE] * 2 <= N));
        if (!last) {
          if (two) { main_knobs(); pdl_suppress(l > 0); }   // later layers must not park CTAs on the free SMs early
          if (a32)
            e = gemm_tf32_layer_a32(x_raw, d, d, nullptr, nullptr, 0, 0, nullptr, F(base + 6), F(base + 7), d, N, ncol, d, o_raw,
                                    nullptr, nullptr, l == 0 ? ws + w.Lm : nullptr, d, nullptr, GI(RG_ACTIVE_POS), nullptr, 0,
                                    nullptr, nullptr, 0, st);
          else
            e = gemm_tf32_layer(x_hi, x_lo, d, F(base + 6), F(base + 7), d, N, ncol, d, d, nullptr, o_hi, o_lo,
                                l == 0 ? ws + w.Lm : nullptr, d, nullptr, GI(RG_ACTIVE_POS), nullptr, 0, nullptr, nullptr, 0, st);
          pdl_suppress(two && side_pdl_off);
          if (e) return e;
          if (two) side_knobs(ncol);
          if (n_active > 0 &&
              (e = gemm_tf32_layer(ws + w.agg_hi, ws + w.agg_lo, 2 * d, F(base + 4), F(base + 5), 2 * d, n_active, d, 2 * d, d,
                                   o_raw, a32 ? nullptr : o_hi, a32 ? nullptr : o_lo, nullptr, 0, arows, nullptr, nullptr, 0,
                                   nullptr, nullptr, 0, sB))) return e;
        } else {
          float* h_new = hist + (size_t)i * nd;
          float* n_hi = next_a32 ? nullptr : ws + w.h_hi;
          float* n_lo = next_a32 ? nullptr : ws + w.h_lo;
          if (two) side_knobs(d);
          if (n_active > 0 &&
              (e = gemm_tf32(ws + w.agg_hi, ws + w.agg_lo, 2 * d, F(base + 4), F(base + 5), 2 * d, ws + w.P, d, n_active, d,
                             2 * d, nullptr, 0, 3, 1, nullptr, 0, nullptr, 0, sB))) return e;
          if (two) { main_knobs(); pdl_suppress(true); }
          if (a32)
            e = gemm_tf32_layer_a32(x_raw, d, d, nullptr, nullptr, 0, 0, nullptr, F(base + 6), F(base + 7), d, N, d, d, h_new,
                                    n_hi, n_lo, nullptr, 0, nullptr, GI(RG_ACTIVE_POS), ws + w.Lm, d, F(RM_GATE_BIAS), h_raw,
                                    layer_norm, st);
          else
            e = gemm_tf32_layer(x_hi, x_lo, d, F(base + 6), F(base + 7), d, N, d, d, d, h_new, n_hi, n_lo,
                                nullptr, 0, nullptr, GI(RG_ACTIVE_POS), ws + w.Lm, d, F(RM_GATE_BIAS), h_raw, layer_norm, st);
          pdl_suppress(false);
          if (e) return e;
          if (two) {                                   // join: the active rows need the gate columns (caller's stream) and P (side stream)
            if (!stream_after(st, sB, aux->b_done)) cudaStreamSynchronize(sB);
            scope.b_open = false;
          }
          if (n_active > 0 &&
              (e = time_gate(ws + w.Lm, F(RM_GATE_BIAS), ws + w.P, h_raw, h_new, n_active, d, layer_norm, d, n_hi, n_lo, st,
                             arows, 1))) return e;
          gate_done = true;
        }
      } else if (sparse) {
        const int ncol = (l == 0 ? 2 : 1) * d;           // [W_evolve (| W_time)]
        float* Le = (l == 0) ? ws + w.Lm : ws + w.L2;
        if ((e = union_aggregate(x_raw, h0_raw, GI(RG_ROWPTR), GI(RG_SRC_SORTED), GI(RG_ETYPE_SORTED), (const float*)g[RG_NORM],
                                 GI(RG_VPTR), GI(RG_SPTR), GI(RG_VROW_ROW), gn[RGI_N_VROWS], gn[RGI_N_SPLIT_CHUNKS], nullptr,
                                 0.f, N, d, nullptr, ws + w.partial, ws + w.agg_hi, ws + w.agg_lo, GI(RG_ACTIVE_POS), 2 * d,
                                 gn[RGI_MAX_CHUNKS], st, fold))) return e;
        if (n_active > 0 &&
            (e = gemm_tf32(ws + w.agg_hi, ws + w.agg_lo, 2 * d, F(base + 4), F(base + 5), 2 * d, ws + w.P, d, n_active, d,
                           2 * d, nullptr, 0, 3, 1, nullptr, 0, nullptr, 0, st))) return e;
        if ((e = gemm_tf32(x_hi, x_lo, d, F(base + 6), F(base + 7), d, Le, ncol, N, ncol, d, nullptr, 0, 3, 1, nullptr, 0,
                           nullptr, 0, st))) return e;
        if ((e = union_combine(ws + w.P, Le, GI(RG_INDEG), nullptr, nullptr, nullptr, N, d, 1, 0, 1.0, o_raw, nullptr, nullptr,
                               ncol, o_hi, o_lo, nullptr, nullptr, GI(RG_ACTIVE_POS), st))) return e;
        if (l == 0) { gate_G = Le + d; gate_ld = ncol; }
      } else {
        const int ncol = (l == 0 ? 3 : 2) * d;           // [W_loop | W_evolve (| W_time)]
        float* Lbuf = (l == 0) ? ws + w.Lm : ws + w.L2;  // layer 0's result carries the gate columns and must survive
        if ((e = union_aggregate(x_raw, h0_raw, GI(RG_ROWPTR), GI(RG_SRC_SORTED), GI(RG_ETYPE_SORTED), (const float*)g[RG_NORM],
                                 GI(RG_VPTR), GI(RG_SPTR), GI(RG_VROW_ROW), gn[RGI_N_VROWS], gn[RGI_N_SPLIT_CHUNKS], nullptr,
                                 0.f, N, d, nullptr, ws + w.partial, ws + w.agg_hi, ws + w.agg_lo, nullptr, d,
                                 gn[RGI_MAX_CHUNKS], st))) return e;
        if ((e = gemm_tf32(x_hi, x_lo, d, F(base + 2), F(base + 3), d, Lbuf, ncol, N, ncol, d, nullptr, 0, 3, 1, nullptr, 0,
                           nullptr, 0, st))) return e;
        if ((e = gemm_tf32(ws + w.agg_hi, ws + w.agg_lo, d, F(base + 0), F(base + 1), d, ws + w.P, d, N, d, d, nullptr, 0, 3,
                           1, nullptr, 0, nullptr, 0, st))) return e;
        if ((e = union_combine(ws + w.P, Lbuf, GI(RG_INDEG), nullptr, nullptr, nullptr, N, d, 1, 0, 1.0, o_raw, nullptr,
                               nullptr, ncol, o_hi, o_lo, nullptr, nullptr, nullptr, st))) return e;
        if (l == 0) { gate_G = Lbuf + 2 * d; gate_ld = ncol; }
      }
      x_raw = o_raw; x_hi = o_hi; x_lo = o_lo;
    }
    // ---- time gate (K9): h = s(h W_t + b) * [normalize](cur) + (1 - s) * h ----
    gemm_tf32_sm_hint(0);
    gemm_tf32_grid_cap(0);
    pdl_suppress(false);
    float* h_new = hist + (size_t)i * nd;
    if (!gate_done &&
        (e = time_gate(gate_G, F(RM_GATE_BIAS), x_raw, h_raw, h_new, N, d, layer_norm, gate_ld, ws + w.h_hi,
                       ws + w.h_lo, st))) return e;
    h_raw = h_new;
  }
  return REGCN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Shared-trajectory form of the batched recurrence (RecurrentRGCN.forward_batch: G history windows of one model as ONE
// recurrence over the block-diagonal union graph, N = G N0 entity rows).
//
// The update of a row WITHOUT in-edges is row-local -- x' = rrelu(x W_evolve) per layer, then the time gate against its
// own previous state (rgcn/layers.py:240-255, src/rrgcn.py:176-178) -- and every window starts from the same table
// (src/rrgcn.py:154).  So until entity v first receives an edge in window g, row (g, v) holds exactly the state U_i[v]
// that v has in every other window where it has not been touched either: the all-entity products of a step are needed
// for N0 shared rows plus the rows that HAVE been active in their window, not for G N0 rows (ICEWS18 shape, 8 windows of
// 6 snapshots: 36 % of the rows on average).  The entity state is therefore kept compact (rowops.cu
// shared_rows_update): rows [0, N0) = the shared trajectory, one further row per (window, entity) from its first
// activity on, in order of arrival.  The all-entity GEMMs run over the compact rows (contiguous: same TMA-fed kernels,
// same per-row arithmetic, so every row is bit-identical to the full computation); the edge kernels keep the union
// numbering and gather from a union-numbered table that only ever holds the rows of the CURRENT step's active
// entities (every gathered row is an active one: each edge source is also a destination).  Only the last step's state
// is expanded to union numbering (h_final); intermediate history_embs are not produced.
// Preconditions (else REGCN_ERR_UNSUPPORTED, callers fall back to regcn_regcn_evolve): >= 2 layers, self_loop, every
// snapshot in the sparse form (n_active <= N / 2).
// ---------------------------------------------------------------------------------------------------------
namespace regcn {
struct SharedWs {
  size_t xm_hi, xm_lo, gi, gh, h0_hi, h0_lo, agg_hi, agg_lo, P, Lm, o_c[2], h_c[2], x_full, o_full[2], partial, rel_partial, fold, fold_n;
  size_t cpos, apos, act, count, total;
  long long cap;
};
static SharedWs plan_shared(int N0, int G, int R2, int d, int max_split_chunks, int rel_nsplit, long long sum_active, int max_active) {
  SharedWs w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += al(n); return o; };
  const long long N = (long long)N0 * G;
  w.cap = N0 + (sum_active < N ? sum_active : N);
  const size_t rd = (size_t)R2 * d, cd = (size_t)w.cap * d, nd = (size_t)N * d;
  const size_t ma = (size_t)(max_active > 0 ? max_active : 1);
  w.xm_hi = take(rd); w.xm_lo = take(rd);
  w.gi = take(rd * 3); w.gh = take(rd * 3);
  w.h0_hi = take(rd); w.h0_lo = take(rd);
  w.agg_hi = take(ma * 2 * d); w.agg_lo = take(ma * 2 * d);
  w.P = take(ma * d);
  w.Lm = take(cd);
  for (int k = 0; k < 2; ++k) { w.o_c[k] = take(cd); w.h_c[k] = take(cd); }
  w.x_full = take(nd);
  for (int k = 0; k < 2; ++k) w.o_full[k] = take(nd);
  w.partial = take((size_t)(max_split_chunks > 0 ? max_split_chunks : 1) * (d + 1));
  w.rel_partial = take(rel_nsplit > 1 ? (size_t)(R2 / 2) * rel_nsplit * d : 1);
  w.fold_n = (size_t)(max_split_chunks > 0 ? max_split_chunks : 1);
  w.fold = take(w.fold_n);
  w.cpos = take((size_t)N); w.apos = take((size_t)w.cap); w.act = take(ma); w.count = take(64);
  w.total = off * sizeof(float);
  return w;
}
// SMs for the side stream in the shared-trajectory form (REGCN_SHARED_SIDE_SMS; 0 = balanced whole-machine grids).  With
// a third of the rows in the all-entity GEMMs the chain of small kernels is the longer one and wants more of the machine
// than in regcn_regcn_evolve.  Measured (B200, ICEWS18 shape, us per timestamp at 8 / 12 / 16 windows per recurrence;
// profiles/time_batched_forward.py): 24 SMs 349 / - / -, 32: 305, 40: 285 / - / 225, 48: 279 / 235 / 220, 56: 261 / 225 / 210,
// 64: 265 / 228 / 208, 72: 271 / 229 / 207, 80: 279; balanced grids: 271.  Re-measured with up to 32 windows (16 / 30 windows):
// 48: 236 / 200, 60: 224 / 191, 72: 218 / 191, 84: 231 / 199, 96: 255 / 221, balanced grids: 223 / 185 -- from ~24 windows
// on both chains are many-round GEMM work and whole-machine grids on both streams win.
static int shared_side_sms(int G) {
  static int v = -2;
  if (v == -2) { const char* e = getenv("REGCN_SHARED_SIDE_SMS"); v = e ? atoi(e) : -1; if (v < -1 || v > 96) v = -1; }
  return v >= 0 ? v : (G >= 24 ? 0 : G > 12 ? 72 : 60);
}
}  // namespace regcn

extern "C" {

size_t regcn_regcn_evolve_shared_workspace_bytes(int N0, int G, int R2, int d, int max_split_chunks, int rel_nsplit,
                                                 long long sum_active, int max_active) {
  return plan_shared(N0, G, R2, d, max_split_chunks, rel_nsplit < 1 ? 1 : rel_nsplit, sum_active, max_active).total;
}

int regcn_regcn_evolve_shared(const void* const* mp, const int* mi, const void* const* gp, const int* gi_, int L, int G,
                              float* h_final, float* h0_out, int rel_nsplit, void* workspace, size_t workspace_bytes,
                              void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!mp || !mi || !gp || !gi_ || !h_final || !h0_out || !workspace) { set_last_error("regcn_evolve_shared: null pointer"); return REGCN_ERR_NULL; }
  const int N = mi[RMI_NUM_ENTS], R2 = mi[RMI_NUM_RELS2], d = mi[RMI_DIM], nl = mi[RMI_NUM_LAYERS];
  const int layer_norm = mi[RMI_LAYER_NORM], self_loop = mi[RMI_SELF_LOOP];
  if (N <= 0 || G < 1 || N % G || R2 <= 0 || (R2 & 1) || d <= 0 || (d & 3) || d > 256 || nl < 1 || nl > 8 || L < 1) {
    set_last_error("regcn_evolve_shared: bad dims N=%d G=%d R2=%d d=%d layers=%d L=%d", N, G, R2, d, nl, L); return REGCN_ERR_DIM;
  }
  if (!self_loop || nl < 2) { set_last_error("regcn_evolve_shared: needs self_loop and >= 2 layers"); return REGCN_ERR_UNSUPPORTED; }
  const int N0 = N / G;
  int max_split = 0, max_active = 0;
  long long sum_active = 0;
  for (int i = 0; i < L; ++i) {
    const int* gn = gi_ + (size_t)i * RGI_NUM_INTS;
    if (gn[RGI_N_SPLIT_CHUNKS] > max_split) max_split = gn[RGI_N_SPLIT_CHUNKS];
    if (gn[RGI_N_ACTIVE] > max_active) max_active = gn[RGI_N_ACTIVE];
    if ((long long)gn[RGI_N_ACTIVE] * 2 > N) { set_last_error("regcn_evolve_shared: snapshot %d is not sparse (%d active of %d rows)", i, gn[RGI_N_ACTIVE], N); return REGCN_ERR_UNSUPPORTED; }
    sum_active += gn[RGI_N_ACTIVE];
  }
  if (rel_nsplit < 1) rel_nsplit = 1;
  const SharedWs w = plan_shared(N0, G, R2, d, max_split, rel_nsplit, sum_active, max_active);
  if (workspace_bytes < w.total) { set_last_error("regcn_evolve_shared: workspace %zu < %zu", workspace_bytes, w.total); return REGCN_ERR_WORKSPACE; }
  float* ws = (float*)workspace;
  auto F = [&](int k) { return (const float*)mp[k]; };
  int* cpos = reinterpret_cast<int*>(ws + w.cpos);
  int* apos = reinterpret_cast<int*>(ws + w.apos);
  int* act_c = reinterpret_cast<int*>(ws + w.act);
  int* count = reinterpret_cast<int*>(ws + w.count);
  int* fold = reinterpret_cast<int*>(ws + w.fold);
  int e;

  // ---- shared rows of step 0: F.normalize(dynamic_emb) if layer_norm (src/rrgcn.py:154); no window row exists yet ----
  float* h_cur = ws + w.h_c[0];
  float* h_nxt = ws + w.h_c[1];
  if (layer_norm) {
    if ((e = row_map(F(RM_DYNAMIC_EMB), h_cur, N0, d, 0, 1.0, nullptr, nullptr, nullptr, st))) return e;
  } else if (cudaMemcpyAsync(h_cur, F(RM_DYNAMIC_EMB), (size_t)N0 * d * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    cudaGetLastError(); set_last_error("regcn_evolve_shared: copy of the initial table failed"); return REGCN_ERR_UNSUPPORTED;
  }
  if (cudaMemsetAsync(cpos, 0xFF, (size_t)N * sizeof(int), st) != cudaSuccess ||
      cudaMemsetAsync(count, 0, sizeof(int), st) != cudaSuccess) { cudaGetLastError(); set_last_error("regcn_evolve_shared: memset failed"); return REGCN_ERR_UNSUPPORTED; }
  if (max_split > 0 && cudaMemsetAsync(fold, 0, w.fold_n * sizeof(int), st) != cudaSuccess) { cudaGetLastError(); fold = nullptr; }
  if (max_split <= 0) fold = nullptr;
  const float* h0_raw = F(RM_EMB_REL);
  const float* h0_hi = F(RM_EMB_REL_HI);
  const float* h0_lo = F(RM_EMB_REL_LO);

  StreamScope scope(two_stream_enabled() ? aux_stream() : nullptr, st);
  AuxStream* aux = scope.aux;
  bool gh_ahead = false;
  int sm_count = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int side_sms = shared_side_sms(G);
  long long rows_bound = N0;              // compact rows that can exist after this step's update (host-side bound)
  for (int i = 0; i < L; ++i) {
    const void* const* g = gp + (size_t)i * RG_NUM_PTRS;
    const int* gn = gi_ + (size_t)i * RGI_NUM_INTS;
    auto GI = [&](int k) { return (const int*)g[k]; };
    const int n_active = gn[RGI_N_ACTIVE];
    rows_bound += n_active;
    const int Mc = (int)(rows_bound < w.cap ? rows_bound : w.cap);
    const int* arows = GI(RG_ACTIVE_ROWS);
    // ---- compact rows of this step's active entities; their state published in union numbering ----
    if (cudaMemsetAsync(apos, 0xFF, (size_t)Mc * sizeof(int), st) != cudaSuccess) { cudaGetLastError(); set_last_error("regcn_evolve_shared: memset failed"); return REGCN_ERR_UNSUPPORTED; }
    if ((e = shared_rows_update(arows, n_active, N0, d, cpos, count, h_cur, ws + w.x_full, act_c, apos, st))) return e;

    bool two = aux != nullptr && N >= 4096;
    bool side_pdl_off = false;
    if (two) {
      two = stream_after(aux->sb, st, aux->fork);
      if (!two) scope.join();
    }
    cudaStream_t sB = two ? aux->sb : st;
    // SM split between the streams: the compact all-entity GEMMs get sm_count - side_sms persistent CTAs when they run
    // at least two rounds on them, the side chain the rest
    const bool split_sms = two && side_sms > 0 && (long long)((Mc + 127) / 128) * 2 >= 2LL * (sm_count - side_sms);
    auto main_knobs = [&]() { gemm_tf32_sm_hint(0); gemm_tf32_grid_cap(split_sms ? sm_count - side_sms : 0); };
    auto side_knobs = [&](int ncol) {
      gemm_tf32_sm_hint(split_sms ? side_sms : side_hint(Mc, d, ncol, sm_count));
      gemm_tf32_grid_cap(split_sms ? side_sms : 0);
    };
    if (two) {
      scope.b_open = true;
      side_pdl_off = !side_pdl_keep();
      side_knobs(2 * d);
      pdl_suppress(side_pdl_off);
    }
    // ---- relation evolution (K2, K3): as in regcn_regcn_evolve, entity rows read from the published table ----
    if (!gh_ahead) {
      cudaStream_t sG = sB;
      if (two && cudaStreamWaitEvent(aux->sc, aux->fork, 0) == cudaSuccess) { sG = aux->sc; scope.c_open = true; }
      if ((e = gemm_tf32(h0_hi, h0_lo, d, F(RM_WHH_HI), F(RM_WHH_LO), d, ws + w.gh, 3 * d, R2, 3 * d, d, F(RM_B_HH), 0, 3, 1,
                         nullptr, 0, nullptr, 0, sG))) return e;
      gh_ahead = sG != sB;
    }
    if ((e = rel_mean_pool(ws + w.x_full, GI(RG_REL_ROWPTR), GI(RG_REL_ENTS), R2 / 2, d, rel_nsplit, nullptr, ws + w.rel_partial,
                           ws + w.xm_hi, ws + w.xm_lo, sB))) return e;
    if ((e = gemm_tf32(ws + w.xm_hi, ws + w.xm_lo, d, F(RM_WIH_R_HI), F(RM_WIH_R_LO), d, ws + w.gi, 3 * d, R2, 3 * d, d,
                       nullptr, 0, 3, 1, nullptr, 0, F(RM_GI_STATIC), 3 * d, sB))) return e;
    if (gh_ahead) {
      if (!stream_after(sB, aux->sc, aux->c_done)) cudaStreamSynchronize(aux->sc);
      scope.c_open = false;
      gh_ahead = false;
    }
    if ((e = gru_gate(ws + w.gi, ws + w.gh, h0_raw, h0_out, R2, d, layer_norm, ws + w.h0_hi, ws + w.h0_lo, sB))) return e;
    h0_raw = h0_out; h0_hi = ws + w.h0_hi; h0_lo = ws + w.h0_lo;
    if (two && i + 1 < L && stream_after(aux->sc, sB, aux->h0_ready)) {
      scope.c_open = true;
      if ((e = gemm_tf32(h0_hi, h0_lo, d, F(RM_WHH_HI), F(RM_WHH_LO), d, ws + w.gh, 3 * d, R2, 3 * d, d, F(RM_B_HH), 0, 3,
                         1, nullptr, 0, nullptr, 0, aux->sc))) return e;
      gh_ahead = true;
    }
    // ---- entity evolution ----
    const float* xa_full = ws + w.x_full;        // union-numbered input of the layer (active rows only)
    const float* xa_c = h_cur;                   // compact input of the layer (all compact rows)
    for (int l = 0; l < nl; ++l) {
      const int base = RM_LAYER0 + RM_LAYER_STRIDE * l;
      const bool last = l == nl - 1;
      const int ncol = (l == 0 ? 2 : 1) * d;           // [W_evolve (| W_time)]
      if ((e = union_aggregate(xa_full, h0_raw, GI(RG_ROWPTR), GI(RG_SRC_SORTED), GI(RG_ETYPE_SORTED), (const float*)g[RG_NORM],
                               GI(RG_VPTR), GI(RG_SPTR), GI(RG_VROW_ROW), gn[RGI_N_VROWS], gn[RGI_N_SPLIT_CHUNKS], nullptr,
                               0.f, N, d, nullptr, ws + w.partial, ws + w.agg_hi, ws + w.agg_lo, GI(RG_ACTIVE_POS), 2 * d,
                               gn[RGI_MAX_CHUNKS], sB, fold))) return e;
      if (!last) {
        float* oc = ws + w.o_c[l & 1];
        float* of = ws + w.o_full[l & 1];
        if (two) { main_knobs(); pdl_suppress(l > 0); }
        e = gemm_tf32_layer_a32(xa_c, d, d, nullptr, nullptr, 0, 0, nullptr, F(base + 6), F(base + 7), d, Mc, ncol, d, oc,
                                nullptr, nullptr, l == 0 ? ws + w.Lm : nullptr, d, nullptr, apos, nullptr, 0, nullptr, nullptr,
                                0, st);
        pdl_suppress(two && side_pdl_off);
        if (e) return e;
        if (two) side_knobs(ncol);
        if (n_active > 0 &&
            (e = gemm_tf32_layer(ws + w.agg_hi, ws + w.agg_lo, 2 * d, F(base + 4), F(base + 5), 2 * d, n_active, d, 2 * d, d,
                                 of, nullptr, nullptr, nullptr, 0, arows, nullptr, nullptr, 0, nullptr, nullptr, 0, sB))) return e;
        xa_full = of; xa_c = oc;
      } else {
        if (two) side_knobs(d);
        if (n_active > 0 &&
            (e = gemm_tf32(ws + w.agg_hi, ws + w.agg_lo, 2 * d, F(base + 4), F(base + 5), 2 * d, ws + w.P, d, n_active, d,
                           2 * d, nullptr, 0, 3, 1, nullptr, 0, nullptr, 0, sB))) return e;
        if (two) { main_knobs(); pdl_suppress(true); }
        e = gemm_tf32_layer_a32(xa_c, d, d, nullptr, nullptr, 0, 0, nullptr, F(base + 6), F(base + 7), d, Mc, d, d, h_nxt,
                                nullptr, nullptr, nullptr, 0, nullptr, apos, ws + w.Lm, d, F(RM_GATE_BIAS), h_cur, layer_norm, st);
        pdl_suppress(false);
        if (e) return e;
        if (two) {
          if (!stream_after(st, sB, aux->b_done)) cudaStreamSynchronize(sB);
          scope.b_open = false;
        }
        // active rows: gate columns and previous state at their compact rows, P in the order of active_rows
        if (n_active > 0 &&
            (e = time_gate(ws + w.Lm, F(RM_GATE_BIAS), ws + w.P, h_cur, h_nxt, n_active, d, layer_norm, d, nullptr, nullptr, st,
                           act_c, 1))) return e;
      }
    }
    gemm_tf32_sm_hint(0);
    gemm_tf32_grid_cap(0);
    pdl_suppress(false);
    float* t = h_cur; h_cur = h_nxt; h_nxt = t;
  }
  return shared_rows_expand(h_cur, cpos, N, N0, d, h_final, st);
}

}  // extern "C"

extern "C" {

// ---------------------------------------------------------------------------------------------------------
// Hyperbolic recurrence (HyperbolicRecurrentRGCN.forward, hyperbolic_model.py:722-890) in one call.
// Encoder 0 = hyperbolic_uvrgcn (radius-weighted union aggregate + W_n GEMM), 1 = lgcn (Lorentz centroid).
// Per snapshot: tangent prep (ht, clamp(ht), |h|, TF32 splits) -> K2 -> GRU GEMMs -> K3 -> per layer
// {K4 or K7, loop GEMM, [W_n GEMM], K5 with clamps / exp_0 / next tangent} -> gate GEMM -> fused time gate +
// projection + residual radius evolution (K9 + K8).
// ---------------------------------------------------------------------------------------------------------
size_t regcn_hyp_evolve_workspace_bytes(int N, int R2, int d, int max_split_chunks, int rel_nsplit) {
  return plan_evolve(N, R2, d, max_split_chunks, rel_nsplit).total + (size_t)8 * al((size_t)N * d) * sizeof(float) +
         al((size_t)2 * N) * sizeof(float) + 1024;
}

int regcn_hyp_evolve(const void* const* mp, const int* mi, const double* md, const void* const* gp, const int* gi_,
                     int L, float* hist, float* h0_out, int rel_nsplit, void* workspace, size_t workspace_bytes,
                     void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!mp || !mi || !md || !gp || !gi_ || !hist || !h0_out || !workspace) { set_last_error("hyp_evolve: null pointer"); return REGCN_ERR_NULL; }
  const int N = mi[HMI_NUM_ENTS], R2 = mi[HMI_NUM_RELS2], d = mi[HMI_DIM], nl = mi[HMI_NUM_LAYERS];
  const int layer_norm = mi[HMI_LAYER_NORM], enc = mi[HMI_ENCODER], nb = mi[HMI_NUM_BASES], residual = mi[HMI_RESIDUAL];
  const double c = md[HMD_C];
  const float gamma = (float)md[HMD_GAMMA], rmin = (float)md[HMD_RMIN], rmax = (float)md[HMD_RMAX];
  const float beta = (float)md[HMD_BETA], eps_r = (float)md[HMD_EPS_R], rb = (float)md[HMD_RADIUS_BIAS];
  if (N <= 0 || R2 <= 0 || (R2 & 1) || d <= 0 || (d & 3) || d > 256 || nl < 1 || nl > 8 || L < 0 || !(c > 0) ||
      enc < 0 || enc > 1 || !mi[HMI_SELF_LOOP]) {
    set_last_error("hyp_evolve: bad model configuration N=%d R2=%d d=%d layers=%d enc=%d", N, R2, d, nl, enc); return REGCN_ERR_DIM;
  }
  int max_split = 0;
  for (int i = 0; i < L; ++i) max_split = gi_[i * RGI_NUM_INTS + RGI_N_SPLIT_CHUNKS] > max_split ? gi_[i * RGI_NUM_INTS + RGI_N_SPLIT_CHUNKS] : max_split;
  if (rel_nsplit < 1) rel_nsplit = 1;
  const EvolveWs w = plan_evolve(N, R2, d, max_split, rel_nsplit);
  if (workspace_bytes < regcn_hyp_evolve_workspace_bytes(N, R2, d, max_split, rel_nsplit)) { set_last_error("hyp_evolve: workspace too small"); return REGCN_ERR_WORKSPACE; }
  float* ws = (float*)workspace;
  const size_t nd = (size_t)N * d;
  float* extra = ws + w.total / sizeof(float);
  float* ht_raw = extra;                  extra += al(nd);
  float* ht_hi = extra;                   extra += al(nd);
  float* ht_lo = extra;                   extra += al(nd);
  float* pt_raw = extra;                  extra += al(nd);
  float* pt_hi = extra;                   extra += al(nd);
  float* pt_lo = extra;                   extra += al(nd);
  float* nx_raw = extra;                  extra += al(nd);   // tangent of a layer output (raw)
  float* h_init = extra;                  extra += al(nd);
  float* rad0 = extra;                    extra += al(N);
  float* rad1 = extra;
  auto F = [&](int k) { return (const float*)mp[k]; };
  int e;

  // h = apply_radius(exp_0([normalize](dynamic_emb)), static_radius)          hyperbolic_model.py:773-782
  if ((e = hyp_init(F(HM_DYNAMIC_EMB), F(HM_RADIUS_STATIC), N, d, layer_norm, 0, c, rmin, rmax, h_init, st))) return e;
  const float* h_raw = h_init;
  const float* h0_raw = F(HM_EMB_REL);
  const float* h0_hi = F(HM_EMB_REL_HI);
  const float* h0_lo = F(HM_EMB_REL_LO);

  for (int i = 0; i < L; ++i) {
    const void* const* g = gp + (size_t)i * RG_NUM_PTRS;
    const int* gn = gi_ + (size_t)i * RGI_NUM_INTS;
    auto GI = [&](int k) { return (const int*)g[k]; };
    if ((e = hyp_tangent(h_raw, N, d, c, ht_raw, pt_raw, rad0, ht_hi, ht_lo, pt_hi, pt_lo, st))) return e;
    // ---- relation evolution on the tangent vectors ----
    if ((e = rel_mean_pool(ht_raw, GI(RG_REL_ROWPTR), GI(RG_REL_ENTS), R2 / 2, d, rel_nsplit, nullptr, ws + w.rel_partial,
                           ws + w.xm_hi, ws + w.xm_lo, st))) return e;
    if ((e = gemm_tf32(ws + w.xm_hi, ws + w.xm_lo, d, F(HM_WIH_R_HI), F(HM_WIH_R_LO), d, ws + w.gi, 3 * d, R2, 3 * d, d,
                       nullptr, 0, 3, 1, nullptr, 0, F(HM_GI_STATIC), 3 * d, st))) return e;
    if ((e = gemm_tf32(h0_hi, h0_lo, d, F(HM_WHH_HI), F(HM_WHH_LO), d, ws + w.gh, 3 * d, R2, 3 * d, d, F(HM_B_HH), 0, 3, 1,
                       nullptr, 0, nullptr, 0, st))) return e;
    if ((e = gru_gate(ws + w.gi, ws + w.gh, h0_raw, h0_out, R2, d, layer_norm, ws + w.h0_hi, ws + w.h0_lo, st))) return e;
    h0_raw = h0_out; h0_hi = ws + w.h0_hi; h0_lo = ws + w.h0_lo;
    // ---- gate pre-activation from the clamped previous tangent (independent of the layers) ----
    if ((e = gemm_tf32(pt_hi, pt_lo, d, F(HM_GATE_W_HI), F(HM_GATE_W_LO), d, ws + w.Lm, d, N, d, d, nullptr, 0, 3, 1, nullptr,
                       0, nullptr, 0, st))) return e;
    // ---- layers ----
    const float* x_t = ht_raw;     // tangent of the layer input
    const float* x_hi = ht_hi;
    const float* x_lo = ht_lo;
    const float* x_rad = rad0;
    const float* out_h = nullptr;
    for (int l = 0; l < nl; ++l) {
      const int base = HM_LAYER0 + HM_LAYER_STRIDE * l;
      const bool last = l == nl - 1;
      float* Lbuf = ws + w.L2;                                     // x_t . [W_loop | W_evolve]
      if ((e = gemm_tf32(x_hi, x_lo, d, F(base + 2), F(base + 3), d, Lbuf, 2 * d, N, 2 * d, d, nullptr, 0, 3, 1, nullptr, 0,
                         nullptr, 0, st))) return e;
      const float* P;
      if (enc == 0) {
        if ((e = union_aggregate(x_t, h0_raw, GI(RG_ROWPTR), GI(RG_SRC_SORTED), GI(RG_ETYPE_SORTED), (const float*)g[RG_NORM],
                                 GI(RG_VPTR), GI(RG_SPTR), GI(RG_VROW_ROW), gn[RGI_N_VROWS], gn[RGI_N_SPLIT_CHUNKS], x_rad,
                                 gamma, N, d, nullptr, ws + w.partial, ws + w.agg_hi, ws + w.agg_lo, nullptr, d,
                                 gn[RGI_MAX_CHUNKS], st))) return e;
        if ((e = gemm_tf32(ws + w.agg_hi, ws + w.agg_lo, d, F(base + 0), F(base + 1), d, ws + w.P, d, N, d, d, nullptr, 0, 3,
                           1, nullptr, 0, nullptr, 0, st))) return e;
        P = ws + w.P;
      } else {
        if ((e = lorentz_aggregate(x_t, F(base + 0), h0_raw, GI(RG_ROWPTR), GI(RG_SRC_SORTED), GI(RG_ETYPE_SORTED),
                                   (const float*)g[RG_NORM], GI(RG_VPTR), GI(RG_SPTR), GI(RG_VROW_ROW), gn[RGI_N_VROWS],
                                   gn[RGI_N_SPLIT_CHUNKS], N, d, nb, c, ws + w.P, ws + w.partial, st))) return e;
        P = ws + w.P;
      }
      float* o_raw = ws + w.set[l & 1][0];
      // next layer consumes the tangent (raw + split) and the radius of this layer's output
      if ((e = union_combine(P, Lbuf, GI(RG_INDEG), nullptr, nullptr, nullptr, N, d, 1, 1, c, o_raw, last ? nullptr : nx_raw,
                             last ? nullptr : rad1, 2 * d, nullptr, nullptr, last ? nullptr : ws + w.set[l & 1][1],
                             last ? nullptr : ws + w.set[l & 1][2], nullptr, st))) return e;
      out_h = o_raw;
      if (!last) {
        // the next layer reads these before its own combine overwrites nx_raw / rad1 (stream order)
        x_t = nx_raw; x_hi = ws + w.set[l & 1][1]; x_lo = ws + w.set[l & 1][2]; x_rad = rad1;
      }
    }
    // ---- time gate + projection + radius evolution ----
    float* h_new = hist + (size_t)i * nd;
    if ((e = hyp_time_gate(out_h, pt_raw, ws + w.Lm, F(HM_GATE_BIAS), F(HM_RADIUS_STATIC), F(HM_RADIUS_W), rb, N, d, layer_norm,
                           residual, c, rmin, rmax, beta, eps_r, h_new, st))) return e;
    h_raw = h_new;
  }
  return REGCN_OK;
}

}  // extern "C"
