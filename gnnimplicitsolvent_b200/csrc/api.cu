// api.cu -- the C ABI of libgnnis.so (include/gnnis.h): handle management, weight packing, workspace and the
// launch sequence of one batched energy/force evaluation.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace gnnis {
int launch_born_count(Ctx*, const float*, int, cudaStream_t);
int launch_fill_edges(Ctx*, const float*, int, int*, int*, uint32_t*, float*, int64_t, cudaStream_t);
int launch_cell_neighbors(Ctx*, const float*, int, int*, int*, uint32_t*, float*, int64_t, cudaStream_t);
int launch_gb_energy(Ctx*, const float*, const float*, int, int, float*, cudaStream_t);
int launch_born_adjoint(Ctx*, const float*, int, int, float*, cudaStream_t);
int launch_reduce_energy(Ctx*, const float*, float*, cudaStream_t);
int launch_edge_fwd(Ctx*, int, cudaStream_t);
int launch_tile_perm(Ctx*, cudaStream_t);
int launch_edge_fwd_ws(Ctx*, int, int, int, cudaStream_t);
int launch_edge_bwd_ws(Ctx*, int, int, int, int, cudaStream_t);
int st_fwd_tables(const Ctx*);
int st_bwd_tables(const Ctx*);
int launch_umma_selftest(Ctx*, const float*, const float*, float*, int, int, cudaStream_t);
bool tc_path_fits(const Ctx*);
int tc_table_mode(const Ctx*);
int launch_edge_bwd(Ctx*, int, int, cudaStream_t);
int launch_node_pre1(Ctx*, cudaStream_t);
int launch_node_fwd(Ctx*, int, cudaStream_t);
int launch_node_bwd(Ctx*, int, cudaStream_t);
int launch_node_fwd_tc(Ctx*, int, cudaStream_t);
int build_edge_images(Ctx*, cudaStream_t);
int build_node_images(Ctx*, cudaStream_t);
int launch_node_bwd_tc(Ctx*, int, cudaStream_t);
int launch_node_pre1_bwd(Ctx*, cudaStream_t);
int launch_gb_only_head(Ctx*, cudaStream_t);
int launch_add_vec(Ctx*, int, const float*, float*, cudaStream_t);
int launch_bond_forces(Ctx*, const float*, float*, float*, cudaStream_t);
int launch_mm_forces(Ctx*, const float*, float*, float*, int, cudaStream_t);

static std::string g_create_err;

enum Stage { kStNeighborBorn = 0, kStEdgeFwd, kStNodeFwd, kStGbEnergy, kStNodeBwd, kStEdgeBwd, kStBornAdjoint, kNumStages };
static const char* kStageNames[kNumStages] = {"neighbor_born", "edge_fwd", "node_fwd", "gb_energy",
                                              "node_bwd", "edge_bwd", "born_adjoint"};

struct StageTimer {
  std::vector<cudaEvent_t> ev;
  std::vector<int> stage;
  cudaStream_t st;
  void mark(int s) {   // called BEFORE the launches of stage s; a final mark(-1) closes the last one
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    stage.push_back(s);
  }
};

template <typename T>
static int dev_alloc(Ctx* c, T** p, size_t count) {
  c->graph_gen++;   // a buffer moves: captured evaluations are stale
  if (*p) {
    cudaFree(*p);
    *p = nullptr;
  }
  if (count == 0) count = 1;
  GNNIS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return 0;
}
template <typename T>
static void dev_free(T** p) {
  if (*p) cudaFree(*p);
  *p = nullptr;
}

static int ensure_workspace(Ctx* c) {
  if (c->N <= 0 || c->R <= 0) return 0;
  if (c->N > kMaxAtoms) {
    c->err = "n_atoms per replica must be <= 1024: the pair kernels keep all pairs of a replica in shared memory (DESIGN.md 4)";
    return -1;
  }
  if (c->N <= kUnitAtomsMax) {
    // replicas per unit: as many as the shared-memory tables hold (full 128-edge tiles), but never so many that
    // fewer than two units per SM remain -- with few replicas the length of one group's serial tile chain, not the
    // tile fill, decides the latency of an evaluation
    c->G = kUnitAtomsMax / c->N;
    const int spread = c->R / (2 * (c->num_sms > 0 ? c->num_sms : 148));
    if (c->G > spread) c->G = spread;
    if (c->G < 1) c->G = 1;
    c->smem_tables = true;
    c->tab_mode = 2;
  } else {
    c->G = 1;
    c->smem_tables = false;
    c->tab_mode = tc_table_mode(c);   // up to ~128 atoms one table per kernel still fits (forward Q, adjoint Qbar)
  }
  c->n_units = (c->R + c->G - 1) / c->G;
  // sub-units per replica for small batches (Ctx::split): as many as it takes to give every warp group (2 per SM) a
  // unit, at least ~6 target atoms (one to two edge tiles) per part, at most split_max parts.  Splitting further (to
  // even out the waves of mid-size batches) does not pay: every part loads the replica's tables, ends on a partly
  // filled tile and writes a partial Qbar table (measured: R = 1024 1.81 -> 1.96 ms, R = 2496 4.06 -> 4.55 ms with 2 parts).
  // What does pay for mid-size batches is splitting only the replicas of the LAST, partly filled wave (Ctx::split_from):
  // R = 1024 on 296 groups = 3 full waves + 136 replicas, which as 272 half-units take ~0.6 instead of 1.0 wave.
  c->split = 1;
  c->split_from = 0;
  if (c->G == 1) {
    const int groups = 2 * (c->num_sms > 0 ? c->num_sms : 148);
    const int rem = c->R % groups;
    int s_ = c->R < groups ? groups / c->R : ((rem > 0 && c->tail_split) ? groups / rem : 1);
    if (s_ > c->N / 6) s_ = c->N / 6;
    if (s_ > c->split_max) s_ = c->split_max;
    if (s_ > 1) {
      c->split = s_;
      if (c->R >= groups) c->split_from = c->R - rem;
    }
  }
  size_t n = (size_t)c->R * c->N;
  int64_t cap = (int64_t)n * (c->N - 1);
  if (cap > 2147483000LL) {
    c->err = "R*N*(N-1) exceeds the int32 CSR range; shard the replicas";
    return -1;
  }
  if (cap < 1) cap = 1;
  if (n == c->ws_atoms && cap == c->edge_cap && c->split <= c->ws_split) return 0;
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  int nblk = (int)((n + kPairThreads - 1) / kPairThreads);
  if (dev_alloc(c, &c->deg, n) || dev_alloc(c, &c->row_ptr, n + 1) || dev_alloc(c, &c->blk_sum, (size_t)nblk + 1) ||
      dev_alloc(c, &c->nbr_mask, n * (size_t)((c->N + 31) / 32)) || dev_alloc(c, &c->col, (size_t)cap) || dev_alloc(c, &c->pair, (size_t)cap) || dev_alloc(c, &c->perm, (size_t)cap + 128) || dev_alloc(c, &c->r_e, (size_t)cap))
    return -1;
  float** scal[] = {&c->born, &c->dBdI, &c->born_s, &c->dE_dBs, &c->Bbar, &c->e_atom, &c->sa_atom};
  for (float** p : scal)
    if (dev_alloc(c, p, n)) return -1;
  for (int l = 0; l < 3; ++l) {
    if (dev_alloc(c, &c->P[l], n * 64) || dev_alloc(c, &c->Q[l], n * 64) || dev_alloc(c, &c->a[l], n * 64) ||
        dev_alloc(c, &c->o[l], n * (l == 2 ? 2 : 64)))
      return -1;
  }
  if (dev_alloc(c, &c->gbar, n * 128)) return -1;
  if (dev_alloc(c, &c->abar, n * 64) || dev_alloc(c, &c->Pbar, n * 64) || dev_alloc(c, &c->Qbar, n * 64 * (size_t)c->split) ||
      dev_alloc(c, &c->rbar, n * (size_t)c->N))
    return -1;
  if (dev_alloc(c, &c->h_pos_dev, n * 3) || dev_alloc(c, &c->h_f_dev, n * 3) || dev_alloc(c, &c->h_e_dev, (size_t)c->R))
    return -1;
  // activation stash: 2 x 128 bytes per edge and layer (+ one tile of slack: partial tiles are addressed, never written,
  // past their last row).  Sized for the edge CAPACITY when that fits the budget; otherwise for as many edges as the
  // budget holds, provided that is at least half of the capacity (real molecules fill 15-60 % of it) -- the kernels then
  // compare the edge count of each evaluation with Ctx::stash_cap on the device (launch_edge_*_ws).
  {
    dev_free(&c->stash);
    c->stash_on = false;
    c->stash_cap = 0;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    double budget = 0.45 * (double)free_b;
    if (c->stash_max_gb >= 0 && c->stash_max_gb * 1e9 < budget) budget = c->stash_max_gb * 1e9;
    int64_t hold = cap;
    if ((double)(6 * ((size_t)cap + kTileE) * 128) > budget) {
      hold = (int64_t)(budget / (6.0 * 128.0)) - kTileE;
      hold -= hold % kTileE;
      if (2 * hold < cap) hold = 0;
    }
    if (c->use_stash && hold > 0 && st_fwd_tables(c) >= 0 && st_bwd_tables(c) >= 0) {
      const size_t per_arr = ((size_t)hold + kTileE) * 128;
      if (dev_alloc(c, &c->stash, 6 * per_arr)) return -1;
      for (int l = 0; l < 3; ++l) {
        c->st_u[l] = c->stash + (size_t)(2 * l) * per_arr;
        c->st_w[l] = c->stash + (size_t)(2 * l + 1) * per_arr;
      }
      c->stash_on = true;
      c->stash_cap = hold;
    }
  }
  if (!c->n_edges_dev && dev_alloc(c, &c->n_edges_dev, 1)) return -1;
  if (!c->unit_ctr && dev_alloc(c, &c->unit_ctr, 8)) return -1;
  GNNIS_CUDA_OK(cudaMemset(c->sa_atom, 0, n * sizeof(float)));
  c->ws_atoms = n;
  c->ws_split = c->split;
  c->edge_cap = cap;
  return 0;
}

static int check_ready(Ctx* c, bool need_weights) {
  if (c->N <= 0) {
    c->err = "gnnis_set_system has not been called";
    return -1;
  }
  if (c->R <= 0) {
    c->err = "gnnis_set_replicas has not been called";
    return -1;
  }
  if (need_weights && !c->have_weights) {
    c->err = "gnnis_load_weights has not been called";
    return -1;
  }
  return 0;
}

static int energy_force_direct(Ctx* c, const float* pos, float* e_rep, float* force, uint32_t flags, cudaStream_t st,
                               StageTimer* tm);

// neighbour back-end of a call: request 1 = cell list, 0 = all pairs, -1 = the handle's default (GNNIS_NEIGHBORS, else by size)
static bool use_cell_list(const Ctx* c, int request) {
  if (request >= 0) return request == 1;
  if (c->nbr_mode >= 0) return c->nbr_mode == 1;
  return c->N > c->nbr_cell_min;
}

static void graph_entry_clear(Ctx::GraphEntry& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  g = Ctx::GraphEntry();
}

// One evaluation.  The launches of an evaluation depend only on (positions / energy / force pointers, flags, the
// minimiser's skip list, the handle state counted by graph_gen), so a combination seen for the second time is captured
// into a CUDA graph (on the handle's own stream; nothing is enqueued by the capture) and from then on replayed with one
// cudaGraphLaunch on the caller's stream: the ~2 us launch gaps between the 23 dependent kernels go away (0.23 -> 0.20 ms
// at N = 13 x R = 128, 6.43 -> 6.38 ms at C2).  Calls inside a caller's own stream capture, profiling calls and
// first sightings use plain stream launches.
int energy_force_impl(Ctx* c, const float* pos, float* e_rep, float* force, uint32_t flags, cudaStream_t st,
                      StageTimer* tm) {
  if (tm || !c->use_graphs) return energy_force_direct(c, pos, e_rep, force, flags, st, tm);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return energy_force_direct(c, pos, e_rep, force, flags, st, tm);
  }
  const uint64_t tick = ++c->graph_tick;
  Ctx::GraphEntry* hit = nullptr;
  Ctx::GraphEntry* victim = &c->gcache[0];
  for (Ctx::GraphEntry& g : c->gcache) {
    if (g.used && g.gen != c->graph_gen) graph_entry_clear(g);
    if (g.used && g.pos == pos && g.e == e_rep && g.f == force && g.flags == flags && g.skip == c->skip_state) hit = &g;
    if (!g.used || (victim->used && g.last_use < victim->last_use)) victim = &g;
  }
  if (!hit) {   // first sighting: remember the combination, launch directly
    graph_entry_clear(*victim);
    victim->used = true;
    victim->pos = pos;
    victim->e = e_rep;
    victim->f = force;
    victim->flags = flags;
    victim->skip = c->skip_state;
    victim->gen = c->graph_gen;
    victim->last_use = tick;
    return energy_force_direct(c, pos, e_rep, force, flags, st, tm);
  }
  hit->last_use = tick;
  if (!hit->exec) {   // second sighting: capture
    GNNIS_CUDA_OK(cudaSetDevice(c->device));
    const int64_t l0 = c->launches;
    if (cudaStreamBeginCapture(c->own_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      return energy_force_direct(c, pos, e_rep, force, flags, st, tm);
    }
    const int rc = energy_force_direct(c, pos, e_rep, force, flags, c->own_stream, nullptr);
    cudaGraph_t graph = nullptr;
    const cudaError_t ec = cudaStreamEndCapture(c->own_stream, &graph);
    const int n_launch = (int)(c->launches - l0);
    c->launches = l0;
    if (rc || ec != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc) return rc;      // argument / state error: reported as without graphs
      graph_entry_clear(*hit);
      return energy_force_direct(c, pos, e_rep, force, flags, st, tm);
    }
    const cudaError_t ei = cudaGraphInstantiate(&hit->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
      cudaGetLastError();
      graph_entry_clear(*hit);
      return energy_force_direct(c, pos, e_rep, force, flags, st, tm);
    }
    hit->launches = n_launch;
  }
  GNNIS_CUDA_OK(cudaGraphLaunch(hit->exec, st));
  c->launches += hit->launches;
  return 0;
}

static int energy_force_direct(Ctx* c, const float* pos, float* e_rep, float* force, uint32_t flags, cudaStream_t st,
                               StageTimer* tm) {
  const bool gb_only = flags & GNNIS_FLAG_GB_ONLY;
  const bool want_grad = !(flags & GNNIS_FLAG_NO_FORCES);
  if (flags & GNNIS_FLAG_VACUUM_ONLY) {
    if (c->N <= 0 || c->R <= 0 || !c->have_ff) {
      c->err = "GNNIS_FLAG_VACUUM_ONLY needs gnnis_set_system, gnnis_set_replicas and gnnis_set_forcefield";
      return -1;
    }
    if (want_grad && !force) {
      c->err = "force_dev is NULL but GNNIS_FLAG_NO_FORCES is not set";
      return -1;
    }
    GNNIS_CUDA_OK(cudaSetDevice(c->device));
    if (launch_mm_forces(c, pos, e_rep, want_grad ? force : nullptr, 0, st)) return -1;
    if (c->n_bonds > 0 && launch_bond_forces(c, pos, e_rep, want_grad ? force : nullptr, st)) return -1;
    return 0;
  }
  const int delta = (flags & GNNIS_FLAG_DELTA) ? 1 : 0;
  const int predicate = (flags & GNNIS_FLAG_DATA_PREDICATE) ? 1 : 0;
  if (check_ready(c, !gb_only)) return -1;
  if (want_grad && !force) {
    c->err = "force_dev is NULL but GNNIS_FLAG_NO_FORCES is not set";
    return -1;
  }
  // arithmetic of the edge MLPs: 3 = 3xTF32 (default, fp32-accurate), 1 = one TF32 product, 2 = bf16 operands + tanh activations
  const int terms = (flags & GNNIS_FLAG_MLP_BF16) ? 2 : ((flags & GNNIS_FLAG_MLP_FP16) ? 4 : ((flags & GNNIS_FLAG_MLP_TF32X1) ? 1 : 3));
  // tensor-core (tcgen05, 3xTF32) edge kernels unless the caller forces the CUDA-core path or the unit's tables
  // do not fit next to the staged weight operands
  const bool tc_ok = !(flags & GNNIS_FLAG_MLP_CUDA_CORES) && tc_path_fits(c);
  const bool tc_fwd = tc_ok && (c->tc_mask & 1), tc_bwd = tc_ok && (c->tc_mask & 2);
  const bool node_tc = !(flags & GNNIS_FLAG_MLP_CUDA_CORES) && (c->tc_mask & 4);
  if ((terms == 2 || terms == 4) && !(tc_fwd && tc_bwd)) {
    c->err = "GNNIS_FLAG_MLP_BF16 / _FP16 need the tensor-core edge kernels (not with GNNIS_FLAG_MLP_CUDA_CORES, GNNIS_TC masks, or "
             "unit tables that do not fit next to the operand images)";
    return -3;
  }
  c->split_eff = (tc_fwd && tc_bwd && c->G == 1) ? c->split : 1;   // the CUDA-core edge kernels do not split replicas
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  const int n = c->R * c->N;
  // the adjoint reads the activations that the forward leaves behind when both edge stages run on tensor cores and the
  // stash fits the memory budget (Ctx::stash_on); otherwise it recomputes them (edge_bwd_ws_kernel)
  const int stash = (want_grad && tc_fwd && tc_bwd && c->stash_on) ? 1 : 0;
  if (want_grad) GNNIS_CUDA_OK(cudaMemsetAsync(c->unit_ctr, 0, 8 * sizeof(unsigned), st));
  if (tm) tm->mark(kStNeighborBorn);
  const bool cells = !gb_only && use_cell_list(c, (flags & GNNIS_FLAG_CELL_LIST) ? 1 : ((flags & GNNIS_FLAG_ALL_PAIRS) ? 0 : -1));
  if (launch_born_count(c, pos, cells ? -1 : predicate, st)) return -1;
  if (!gb_only) {
    if (cells ? launch_cell_neighbors(c, pos, predicate, c->row_ptr, nullptr, c->pair, c->r_e, c->edge_cap, st)
              : launch_fill_edges(c, pos, predicate, c->row_ptr, nullptr, c->pair, c->r_e, c->edge_cap, st))
      return -1;
    if (want_grad && launch_tile_perm(c, st)) return -1;
    if (tm) tm->mark(kStNodeFwd);
    if (launch_node_pre1(c, st)) return -1;
    for (int l = 0; l < 3; ++l) {
      if (tm) tm->mark(kStEdgeFwd);
      if (tc_fwd ? launch_edge_fwd_ws(c, l, (terms == 3 && c->act_tanh) ? 7 : terms, stash, st) : launch_edge_fwd(c, l, st)) return -1;
      if (tm) tm->mark(kStNodeFwd);
      if ((node_tc ? launch_node_fwd_tc(c, l, st) : launch_node_fwd(c, l, st))) return -1;
    }
  } else {
    if (launch_gb_only_head(c, st)) return -1;
  }
  if (tm) tm->mark(kStGbEnergy);
  if (launch_gb_energy(c, pos, c->born_s, delta, want_grad ? 1 : 0, force, st)) return -1;
  if (launch_reduce_energy(c, gb_only ? nullptr : c->sa_atom, e_rep, st)) return -1;
  if (want_grad) {
    if (gb_only) {
      if (launch_add_vec(c, n, c->dE_dBs, c->Bbar, st)) return -1;
    } else {
      for (int l = 2; l >= 0; --l) {
        if (tm) tm->mark(kStNodeBwd);
        if ((node_tc ? launch_node_bwd_tc(c, l, st) : launch_node_bwd(c, l, st))) return -1;
        if (tm) tm->mark(kStEdgeBwd);
        if (tc_bwd ? launch_edge_bwd_ws(c, l, l == 2 ? 0 : 1, (terms == 3 && stash && c->adj_terms) ? c->adj_terms : terms, stash, st)
                   : launch_edge_bwd(c, l, l == 2 ? 0 : 1, st))
          return -1;
      }
      if (tm) tm->mark(kStNodeBwd);
      if (launch_node_pre1_bwd(c, st)) return -1;
    }
    if (tm) tm->mark(kStBornAdjoint);
    if (launch_born_adjoint(c, pos, gb_only ? 0 : 1, predicate, force, st)) return -1;
    if (c->n_bonds > 0 && launch_bond_forces(c, pos, e_rep, force, st)) return -1;
    if (c->have_ff && launch_mm_forces(c, pos, e_rep, force, 1, st)) return -1;
  } else {
    if (c->n_bonds > 0 && launch_bond_forces(c, pos, e_rep, nullptr, st)) return -1;
    if (c->have_ff && launch_mm_forces(c, pos, e_rep, nullptr, 1, st)) return -1;
  }
  if (tm) tm->mark(-1);
  return 0;
}

}  // namespace gnnis

using namespace gnnis;

extern "C" {

int gnnis_create(gnnis_handle* out, int device) {
  if (!out) return -1;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_err = std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                   " (libgnnis has no CPU fallback by design)";
    return -1;
  }
  if (device < 0 || device >= count) {
    g_create_err = "device index out of range";
    return -1;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_err = "libgnnis is built for sm_100a (B200) only; found compute capability " +
                   std::to_string(prop.major) + "." + std::to_string(prop.minor);
    return -1;
  }
  Ctx* c = new Ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  cudaSetDevice(device);
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_err = "cudaStreamCreate failed";
    delete c;
    return -1;
  }
  if (const char* e = getenv("GNNIS_GRAPHS")) c->use_graphs = atoi(e);   // debug: 0 = no CUDA-graph replay of evaluations
  if (const char* e = getenv("GNNIS_NEIGHBORS")) c->nbr_mode = !strcmp(e, "cell") ? 1 : (!strcmp(e, "allpairs") ? 0 : -1);
  if (const char* e = getenv("GNNIS_CELL_MIN")) c->nbr_cell_min = atoi(e);
  if (const char* e = getenv("GNNIS_TAIL_SPLIT")) c->tail_split = atoi(e) != 0;
  if (const char* e = getenv("GNNIS_SPLIT")) c->split_max = atoi(e) < 1 ? 1 : (atoi(e) > 16 ? 16 : atoi(e));
  if (const char* e = getenv("GNNIS_STASH")) c->use_stash = atoi(e);      // 0: the edge adjoint always recomputes its activations
  if (const char* e = getenv("GNNIS_STASH_GB")) c->stash_max_gb = atof(e);
  if (const char* e = getenv("GNNIS_ONEBUF")) c->one_buf = atoi(e);
  if (const char* e = getenv("GNNIS_PDL")) c->pdl = atoi(e);
  if (const char* e = getenv("GNNIS_LB_STALL")) c->lb_stall_limit = atoi(e) < 1 ? 1 : atoi(e);
  if (const char* e = getenv("GNNIS_LB_MAXDISP")) c->lb_max_disp = (float)atof(e);
  if (const char* e = getenv("GNNIS_ACT")) c->act_tanh = atoi(e);    // 0: exponential-based sigmoid in the forward edge kernels
  if (const char* e = getenv("GNNIS_ADJ")) c->adj_terms = atoi(e);   // experiment: arithmetic of the stash adjoint's products (1, 3, 4)
  if (const char* e = getenv("GNNIS_TC")) c->tc_mask = atoi(e);   // debug: bit0 edge fwd, bit1 edge adjoint, bit2 node kernels on tensor cores
  *out = reinterpret_cast<gnnis_handle>(c);
  return 0;
}

int gnnis_destroy(gnnis_handle h) {
  if (!h) return 0;
  Ctx* c = reinterpret_cast<Ctx*>(h);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (Ctx::GraphEntry& g : c->gcache) graph_entry_clear(g);
  dev_free(&c->wblob);
  dev_free(&c->img);
  float** f[] = {&c->q, &c->rho, &c->s, &c->alpha, &c->beta, &c->gamma, &c->d0, &c->m0, &c->pref, &c->r_e, &c->born,
                 &c->dBdI, &c->born_s, &c->dE_dBs, &c->Bbar, &c->abar, &c->Pbar, &c->Qbar, &c->rbar, &c->e_atom,
                 &c->sa_atom, &c->h_pos_dev, &c->h_e_dev, &c->h_f_dev, &c->bond_r0, &c->bond_k, &c->lb_S, &c->lb_Y,
                 &c->lb_rho, &c->lb_alpha, &c->lb_g, &c->lb_gprev, &c->lb_xprev, &c->lb_d, &c->lb_e, &c->lb_eprev,
                 &c->lb_step, &c->lb_ftmp, &c->lb_etmp, &c->masses, &c->gamma_emb, &c->gbar, &c->lb_eref, &c->lb_grms};
  for (float** p : f) dev_free(p);
  for (int l = 0; l < 3; ++l) {
    dev_free(&c->P[l]);
    dev_free(&c->Q[l]);
    dev_free(&c->a[l]);
    dev_free(&c->o[l]);
  }
  cudaFree(c->cl_atom_ptr);
  cudaFree(c->cl_atoms);
  cudaFree(c->cl_con_ptr);
  cudaFree(c->cl_con);
  cudaFree(c->cl_con_d);
  cudaFree(c->ff_excl);
  cudaFree(c->ff_bond);
  cudaFree(c->ff_angle);
  cudaFree(c->ff_tors);
  cudaFree(c->ff_exc);
  cudaFree(c->ff_bond_p);
  cudaFree(c->ff_angle_p);
  cudaFree(c->ff_tors_p);
  cudaFree(c->ff_exc_p);
  float** fff[] = {&c->ff_charge, &c->ff_sigma, &c->ff_epsilon};
  for (float** p : fff) dev_free(p);
  dev_free(&c->ff_term_ptr);
  dev_free(&c->ff_term_ref);
  int** ip[] = {&c->radidx, &c->solvent_id, &c->deg, &c->row_ptr, &c->blk_sum, &c->col, &c->bond_idx, &c->lb_state,
                &c->lb_count, &c->lb_active, &c->lb_stall, &c->lb_nevals};
  for (int** p : ip) dev_free(p);
  dev_free(&c->pair);
  dev_free(&c->nbr_mask);
  dev_free(&c->perm);
  dev_free(&c->n_edges_dev);
  dev_free(&c->unit_ctr);
  dev_free(&c->stash);
  for (Ctx::HostSlot& s : c->hs) {
    dev_free(&s.pos);
    dev_free(&s.e);
    dev_free(&s.f);
    if (s.in) cudaEventDestroy(s.in);
    if (s.done) cudaEventDestroy(s.done);
    if (s.out) cudaEventDestroy(s.out);
  }
  if (c->hs_in) cudaStreamDestroy(c->hs_in);
  if (c->hs_out) cudaStreamDestroy(c->hs_out);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return 0;
}

const char* gnnis_last_error(gnnis_handle h) {
  if (!h) return g_create_err.c_str();
  return reinterpret_cast<Ctx*>(h)->err.c_str();
}

int gnnis_load_weights(gnnis_handle h, const gnnis_weights* w) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (c) c->graph_gen++;
  if (!c || !w) return -1;
  if (w->hidden != kH || w->hidden_token != kT || w->n_rbf != kRbf || w->num_solvents < 1) {
    c->err = "unsupported architecture: this build is specialised for hidden=64, hidden_token=128, 20 RBFs "
             "(the shipped GNN.pt architecture, train_config.yml)";
    return -1;
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  const int S = w->num_solvents;
  std::vector<float> blob;
  auto reserve = [&](size_t count) {
    size_t off = blob.size();
    size_t padded = (count + 63) / 64 * 64;
    blob.resize(off + padded, 0.0f);
    return off;
  };
  struct Off {
    size_t Wi_t, Wj_t, Wr_t, b1, W2_t, b2, W3a_t, sbias, W4_t, b4, Wi, Wj, Wr, W2, W3a, W4;
  } off[3];
  for (int l = 0; l < 3; ++l) {
    const int cin = l == 0 ? 3 : kH, cout = l == 2 ? 2 : kH, ld1 = 2 * cin + kRbf;
    const float *W1 = w->message1_w[l], *W2 = w->message2_w[l], *W3 = w->lin1_w[l], *W4 = w->lin2_w[l];
    Off& o = off[l];
    o.Wi_t = reserve((size_t)cin * 64);
    o.Wj_t = reserve((size_t)cin * 64);
    o.Wr_t = reserve(kRbf * 64);
    o.b1 = reserve(64);
    o.W2_t = reserve(64 * 64);
    o.b2 = reserve(64);
    o.W3a_t = reserve(64 * 128);
    o.sbias = reserve((size_t)S * 128);
    o.W4_t = reserve((size_t)128 * cout);
    o.b4 = reserve(cout);
    o.Wi = reserve((size_t)64 * cin);
    o.Wj = reserve((size_t)64 * cin);
    o.Wr = reserve(64 * kRbf);
    o.W2 = reserve(64 * 64);
    o.W3a = reserve(128 * 64);
    o.W4 = reserve((size_t)cout * 128);
    for (int cc = 0; cc < 64; ++cc) {
      for (int k = 0; k < cin; ++k) {
        blob[o.Wi_t + (size_t)k * 64 + cc] = W1[(size_t)cc * ld1 + k];
        blob[o.Wj_t + (size_t)k * 64 + cc] = W1[(size_t)cc * ld1 + cin + k];
        blob[o.Wi + (size_t)cc * cin + k] = W1[(size_t)cc * ld1 + k];
        blob[o.Wj + (size_t)cc * cin + k] = W1[(size_t)cc * ld1 + cin + k];
      }
      for (int k = 0; k < kRbf; ++k) {
        blob[o.Wr_t + (size_t)k * 64 + cc] = W1[(size_t)cc * ld1 + 2 * cin + k];
        blob[o.Wr + (size_t)cc * kRbf + k] = W1[(size_t)cc * ld1 + 2 * cin + k];
      }
      blob[o.b1 + cc] = w->message1_b[l][cc];
      blob[o.b2 + cc] = w->message2_b[l][cc];
      for (int k = 0; k < 64; ++k) {
        blob[o.W2_t + (size_t)k * 64 + cc] = W2[(size_t)cc * 64 + k];
        blob[o.W2 + (size_t)cc * 64 + k] = W2[(size_t)cc * 64 + k];
      }
    }
    for (int m = 0; m < 128; ++m) {
      for (int k = 0; k < 64; ++k) {
        blob[o.W3a_t + (size_t)k * 128 + m] = W3[(size_t)m * 128 + k];
        blob[o.W3a + (size_t)m * 64 + k] = W3[(size_t)m * 128 + k];
      }
      for (int s = 0; s < S; ++s) {
        double acc = 0.0;
        for (int k = 0; k < 64; ++k) acc += (double)W3[(size_t)m * 128 + 64 + k] * (double)w->solvent_embedding[(size_t)s * 64 + k];
        blob[o.sbias + (size_t)s * 128 + m] = (float)(acc + (double)w->lin1_b[l][m]);
      }
      for (int cc = 0; cc < cout; ++cc) {
        blob[o.W4_t + (size_t)m * cout + cc] = W4[(size_t)cc * 128 + m];
        blob[o.W4 + (size_t)cc * 128 + m] = W4[(size_t)cc * 128 + m];
      }
    }
    for (int cc = 0; cc < cout; ++cc) blob[o.b4 + cc] = w->lin2_b[l][cc];
  }
  size_t gamma_off = reserve(S);
  for (int s = 0; s < S; ++s) blob[gamma_off + s] = w->gamma_embedding[s];
  if (dev_alloc(c, &c->wblob, blob.size())) return -1;
  GNNIS_CUDA_OK(cudaMemcpy(c->wblob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  for (int l = 0; l < 3; ++l) {
    LayerW& L = c->L[l];
    const Off& o = off[l];
    L.cin = l == 0 ? 3 : kH;
    L.cout = l == 2 ? 2 : kH;
    L.Wi_t = c->wblob + o.Wi_t;
    L.Wj_t = c->wblob + o.Wj_t;
    L.Wr_t = c->wblob + o.Wr_t;
    L.b1 = c->wblob + o.b1;
    L.W2_t = c->wblob + o.W2_t;
    L.b2 = c->wblob + o.b2;
    L.W3a_t = c->wblob + o.W3a_t;
    L.sbias = c->wblob + o.sbias;
    L.W4_t = c->wblob + o.W4_t;
    L.b4 = c->wblob + o.b4;
    L.Wi = c->wblob + o.Wi;
    L.Wj = c->wblob + o.Wj;
    L.Wr = c->wblob + o.Wr;
    L.W2 = c->wblob + o.W2;
    L.W3a = c->wblob + o.W3a;
    L.W4 = c->wblob + o.W4;
  }
  // tensor-core operand images (zero-filled: the padded K columns of Wr stay zero)
  {
    const size_t per_layer = 81920 + 196608 + 131072 + 131072 + 2 * 24576;
    if (dev_alloc(c, &c->img, 3 * per_layer)) return -1;
    GNNIS_CUDA_OK(cudaMemset(c->img, 0, 3 * per_layer));
    for (int l = 0; l < 3; ++l) {
      uint8_t* base = c->img + (size_t)l * per_layer;
      c->img_edge[l] = base;
      c->img_nf[l] = base + 81920;
      c->img_nba[l] = base + 81920 + 196608;
      c->img_nbb[l] = base + 81920 + 196608 + 131072;
      c->img_edge_bf[l] = base + 81920 + 196608 + 131072 + 131072;
    }
    if (build_edge_images(c, 0) || build_node_images(c, 0)) return -1;
    GNNIS_CUDA_OK(cudaDeviceSynchronize());
  }
  if (dev_alloc(c, &c->gamma_emb, (size_t)S)) return -1;   // (dev_alloc releases the table of a previous load)
  GNNIS_CUDA_OK(cudaMemcpy(c->gamma_emb, w->gamma_embedding, S * sizeof(float), cudaMemcpyHostToDevice));
  c->S = S;
  c->have_weights = true;
  for (int sid : c->solvent_host)
    if (sid >= S) {   // replicas configured before the weights were (re)loaded must fit the new embedding table
      c->err = "gnnis_load_weights: a replica uses solvent id " + std::to_string(sid) + " but the checkpoint has " +
               std::to_string(S) + " solvents; call gnnis_set_replicas again";
      c->R = 0;
      return -1;
    }
  return 0;
}

int gnnis_set_system(gnnis_handle h, int32_t n_atoms, const float* params7, const float* d0, const float* m0,
                     int32_t n_unique) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (c) c->graph_gen++;
  if (!c) return -1;
  if (n_atoms < 1 || !params7 || !d0 || !m0 || n_unique < 1) {
    c->err = "gnnis_set_system: bad arguments";
    return -1;
  }
  if (n_atoms > kMaxAtoms) {
    c->err = "gnnis_set_system: " + std::to_string(n_atoms) + " atoms per replica; this library supports at most " +
             std::to_string(kMaxAtoms) + " (all pairs of a replica are processed in shared memory; there is no cell-list "
             "path for larger systems, see INTEGRATION.md 'Size limits')";
    return -1;
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  std::vector<float> col[6];
  std::vector<int> ridx(n_atoms);
  for (auto& v : col) v.resize(n_atoms);
  for (int i = 0; i < n_atoms; ++i) {
    for (int k = 0; k < 6; ++k) col[k][i] = params7[(size_t)i * 7 + k];
    int r = (int)params7[(size_t)i * 7 + 6];   // .to(torch.int) truncation, GNN_Layers.py:1023
    if (r < 0 || r >= n_unique) {
      c->err = "gnnis_set_system: radindex out of range of the neck tables";
      return -1;
    }
    ridx[i] = r;
  }
  float** dst[6] = {&c->q, &c->rho, &c->s, &c->alpha, &c->beta, &c->gamma};
  for (int k = 0; k < 6; ++k) {
    if (dev_alloc(c, dst[k], (size_t)n_atoms)) return -1;
    GNNIS_CUDA_OK(cudaMemcpy(*dst[k], col[k].data(), n_atoms * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (dev_alloc(c, &c->radidx, (size_t)n_atoms)) return -1;
  GNNIS_CUDA_OK(cudaMemcpy(c->radidx, ridx.data(), n_atoms * sizeof(int), cudaMemcpyHostToDevice));
  size_t uu = (size_t)n_unique * n_unique;
  if (dev_alloc(c, &c->d0, uu) || dev_alloc(c, &c->m0, uu)) return -1;
  GNNIS_CUDA_OK(cudaMemcpy(c->d0, d0, uu * sizeof(float), cudaMemcpyHostToDevice));
  GNNIS_CUDA_OK(cudaMemcpy(c->m0, m0, uu * sizeof(float), cudaMemcpyHostToDevice));
  c->N = n_atoms;
  c->U = n_unique;
  c->n_bonds = 0;        // per-molecule add-ons belong to the previous system
  c->have_ff = false;
  c->n_clusters = 0;
  return ensure_workspace(c);
}

int gnnis_set_replicas(gnnis_handle h, int32_t n_rep, const int32_t* solvent_id, const float* dielectric) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (c) c->graph_gen++;
  if (!c) return -1;
  if (n_rep < 1 || !solvent_id || !dielectric) {
    c->err = "gnnis_set_replicas: bad arguments";
    return -1;
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  std::vector<float> pref(n_rep);
  for (int r = 0; r < n_rep; ++r) {
    if (solvent_id[r] < 0 || (c->have_weights && solvent_id[r] >= c->S)) {
      c->err = "gnnis_set_replicas: solvent id outside the embedding table";
      return -1;
    }
    // fp32 evaluation order of GNN_Layers.py:2133-2139: (-0.5*138.935485) * (1/solute - 1/eps)
    float inv = 1.0f / dielectric[r];
    pref[r] = (float)(-0.5 * 138.935485) * (1.0f - inv);
  }
  if (dev_alloc(c, &c->solvent_id, (size_t)n_rep) || dev_alloc(c, &c->pref, (size_t)n_rep)) return -1;
  GNNIS_CUDA_OK(cudaMemcpy(c->solvent_id, solvent_id, n_rep * sizeof(int), cudaMemcpyHostToDevice));
  GNNIS_CUDA_OK(cudaMemcpy(c->pref, pref.data(), n_rep * sizeof(float), cudaMemcpyHostToDevice));
  c->R = n_rep;
  c->solvent_host.assign(solvent_id, solvent_id + n_rep);
  return ensure_workspace(c);
}

int gnnis_set_model_constants(gnnis_handle h, float fraction, float scaling_factor, float cutoff) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (c) c->graph_gen++;
  if (!c) return -1;
  if (!(cutoff > 0.0f)) {
    c->err = "cutoff must be positive";
    return -1;
  }
  c->fraction = fraction;
  c->scaling = scaling_factor;
  c->cutoff = cutoff;
  return 0;
}

int gnnis_build_neighbors(gnnis_handle h, const float* pos_dev, int32_t predicate, int32_t* row_ptr_dev,
                          int32_t* col_dev, float* r_dev, int64_t capacity, int64_t* n_edges_host, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if (check_ready(c, false)) return -1;
  if (!pos_dev || !row_ptr_dev || !n_edges_host) {
    c->err = "gnnis_build_neighbors: NULL argument";
    return -1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  const int n = c->R * c->N;
  const bool cells = use_cell_list(c, (predicate & GNNIS_PREDICATE_CELL_LIST) ? 1 : ((predicate & GNNIS_PREDICATE_ALL_PAIRS) ? 0 : -1));
  predicate &= 1;
  if (launch_born_count(c, pos_dev, cells ? -1 : predicate, st)) return -1;
  if (cells ? launch_cell_neighbors(c, pos_dev, predicate, row_ptr_dev, col_dev, nullptr, r_dev, col_dev ? capacity : 0, st)
            : launch_fill_edges(c, pos_dev, predicate, row_ptr_dev, col_dev, nullptr, r_dev, col_dev ? capacity : 0, st))
    return -1;
  GNNIS_CUDA_OK(cudaMemcpyAsync(row_ptr_dev + n, c->row_ptr + n, sizeof(int), cudaMemcpyDeviceToDevice, st));
  GNNIS_CUDA_OK(cudaMemcpyAsync(n_edges_host, c->n_edges_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  GNNIS_CUDA_OK(cudaStreamSynchronize(st));
  if (*n_edges_host > capacity) {
    c->err = "gnnis_build_neighbors: capacity too small";
    return -2;
  }
  return 0;
}

int gnnis_energy_force(gnnis_handle h, const float* pos_dev, float* e_rep_dev, float* force_dev, uint32_t flags,
                       void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if (!pos_dev || !e_rep_dev) {
    c->err = "gnnis_energy_force: NULL argument";
    return -1;
  }
  return energy_force_impl(c, pos_dev, e_rep_dev, force_dev, flags, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int gnnis_energy_force_host(gnnis_handle h, const float* pos_host, float* e_rep_host, float* force_host, uint32_t flags) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if (check_ready(c, !(flags & (GNNIS_FLAG_GB_ONLY | GNNIS_FLAG_VACUUM_ONLY)))) return -1;
  if (!pos_host || !e_rep_host) {
    c->err = "gnnis_energy_force_host: NULL argument";
    return -1;
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->own_stream;
  size_t n = (size_t)c->R * c->N;
  GNNIS_CUDA_OK(cudaMemcpyAsync(c->h_pos_dev, pos_host, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
  bool want = !(flags & GNNIS_FLAG_NO_FORCES);
  int rc = energy_force_impl(c, c->h_pos_dev, c->h_e_dev, want ? c->h_f_dev : nullptr, flags, st, nullptr);
  if (rc) return rc;
  GNNIS_CUDA_OK(cudaMemcpyAsync(e_rep_host, c->h_e_dev, c->R * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (want && force_host)
    GNNIS_CUDA_OK(cudaMemcpyAsync(force_host, c->h_f_dev, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  GNNIS_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

// ---- pipelined variant: up to two calls in flight; copies on their own streams, evaluations in submission order
int gnnis_energy_force_host_submit(gnnis_handle h, const float* pos_host, float* e_rep_host, float* force_host,
                                   uint32_t flags) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if (check_ready(c, !(flags & (GNNIS_FLAG_GB_ONLY | GNNIS_FLAG_VACUUM_ONLY)))) return -1;
  if (!pos_host || !e_rep_host) {
    c->err = "gnnis_energy_force_host_submit: NULL argument";
    return -1;
  }
  if (c->hs_inflight >= 2) {
    c->err = "gnnis_energy_force_host_submit: two calls are already in flight; call gnnis_energy_force_host_wait first";
    return -4;
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  if (!c->hs_in) {
    GNNIS_CUDA_OK(cudaStreamCreateWithFlags(&c->hs_in, cudaStreamNonBlocking));
    GNNIS_CUDA_OK(cudaStreamCreateWithFlags(&c->hs_out, cudaStreamNonBlocking));
  }
  const size_t n = (size_t)c->R * c->N;
  Ctx::HostSlot& s = c->hs[c->hs_next];
  if (s.atoms != n || s.R != c->R) {   // the slot is idle here: its previous call has been waited for
    if (dev_alloc(c, &s.pos, n * 3) || dev_alloc(c, &s.f, n * 3) || dev_alloc(c, &s.e, (size_t)c->R)) return -1;
    s.atoms = n;
    s.R = c->R;
  }
  if (!s.in) {
    GNNIS_CUDA_OK(cudaEventCreateWithFlags(&s.in, cudaEventDisableTiming));
    GNNIS_CUDA_OK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    GNNIS_CUDA_OK(cudaEventCreateWithFlags(&s.out, cudaEventDisableTiming));
  }
  const bool want = !(flags & GNNIS_FLAG_NO_FORCES);
  GNNIS_CUDA_OK(cudaMemcpyAsync(s.pos, pos_host, n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->hs_in));
  GNNIS_CUDA_OK(cudaEventRecord(s.in, c->hs_in));
  GNNIS_CUDA_OK(cudaStreamWaitEvent(c->own_stream, s.in, 0));
  int rc = energy_force_impl(c, s.pos, s.e, want ? s.f : nullptr, flags, c->own_stream, nullptr);
  if (rc) return rc;
  GNNIS_CUDA_OK(cudaEventRecord(s.done, c->own_stream));
  GNNIS_CUDA_OK(cudaStreamWaitEvent(c->hs_out, s.done, 0));
  GNNIS_CUDA_OK(cudaMemcpyAsync(e_rep_host, s.e, c->R * sizeof(float), cudaMemcpyDeviceToHost, c->hs_out));
  if (want && force_host)
    GNNIS_CUDA_OK(cudaMemcpyAsync(force_host, s.f, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->hs_out));
  GNNIS_CUDA_OK(cudaEventRecord(s.out, c->hs_out));
  c->hs_next ^= 1u;
  c->hs_inflight += 1;
  return 0;
}

int gnnis_energy_force_host_wait(gnnis_handle h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if (c->hs_inflight <= 0) {
    c->err = "gnnis_energy_force_host_wait: nothing in flight";
    return -4;
  }
  // oldest call in flight: with two in flight it sits in the slot the next submission would take
  const unsigned k = c->hs_inflight == 2 ? c->hs_next : (c->hs_next ^ 1u);
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  GNNIS_CUDA_OK(cudaEventSynchronize(c->hs[k].out));
  c->hs_inflight -= 1;
  return 0;
}

int gnnis_get_atom_energies(gnnis_handle h, float* e_atom_dev, float* sa_atom_dev, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || check_ready(c, false)) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  size_t n = (size_t)c->R * c->N;
  if (e_atom_dev) GNNIS_CUDA_OK(cudaMemcpyAsync(e_atom_dev, c->e_atom, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (sa_atom_dev) GNNIS_CUDA_OK(cudaMemcpyAsync(sa_atom_dev, c->sa_atom, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int gnnis_get_born_radii(gnnis_handle h, float* born_dev, float* born_scaled_dev, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || check_ready(c, false)) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  size_t n = (size_t)c->R * c->N;
  if (born_dev) GNNIS_CUDA_OK(cudaMemcpyAsync(born_dev, c->born, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (born_scaled_dev) GNNIS_CUDA_OK(cudaMemcpyAsync(born_scaled_dev, c->born_s, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int gnnis_last_edge_count(gnnis_handle h, int64_t* n_edges_host, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || !n_edges_host || check_ready(c, false)) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GNNIS_CUDA_OK(cudaMemcpyAsync(n_edges_host, c->n_edges_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  GNNIS_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int gnnis_umma_selftest(gnnis_handle h, const float* a_dev, const float* b_dev, float* d_dev, int32_t k, int32_t n,
                        void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if ((k != 32 && k != 64) || (n != 32 && n != 64)) {
    c->err = "gnnis_umma_selftest: K and N must be 32 or 64";
    return -1;
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  return launch_umma_selftest(c, a_dev, b_dev, d_dev, k, n, reinterpret_cast<cudaStream_t>(stream));
}

int64_t gnnis_launch_count(gnnis_handle h) { return h ? reinterpret_cast<Ctx*>(h)->launches : 0; }

const char* gnnis_describe(gnnis_handle h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return "";
  const bool tc = tc_path_fits(c);
  std::string d = "N=" + std::to_string(c->N) + " R=" + std::to_string(c->R) + " replicas_per_unit=" + std::to_string(c->G) +
                  " units=" + std::to_string(c->n_units) + " split=" + std::to_string(c->split) +
                  (c->split > 1 && c->split_from > 0 ? " (replicas " + std::to_string(c->split_from) + ".." + std::to_string(c->R - 1) + " only)" : "");
  d += std::string(" edge_kernels=") + (tc ? "tcgen05" : "cuda-cores (unit tables do not fit next to the operand images)");
  if (tc) {
    d += std::string(" adjoint=") + (c->stash_on ? (c->stash_cap < c->edge_cap ? "stash up to " + std::to_string(c->stash_cap) + " edges, recompute beyond (decided per evaluation on the device)" : std::string("stash"))
                                                 : (c->use_stash ? "recompute (stash exceeds the memory budget)" : "recompute (GNNIS_STASH=0)"));
    d += " tables=" + std::to_string(c->stash_on ? st_fwd_tables(c) : c->tab_mode);
  }
  d += " stash_bytes=" + std::to_string(c->stash_on ? (size_t)6 * ((size_t)c->stash_cap + kTileE) * 128 : (size_t)0);
  d += std::string(" neighbors=") + (use_cell_list(c, -1) ? "cell-list" : "all-pairs");
  c->desc = d;
  return c->desc.c_str();
}

int32_t gnnis_num_stages(void) { return kNumStages; }
const char* gnnis_stage_name(int32_t i) { return (i >= 0 && i < kNumStages) ? kStageNames[i] : ""; }

int gnnis_profile_eval(gnnis_handle h, const float* pos_dev, float* e_rep_dev, float* force_dev, uint32_t flags,
                       void* stream, float* stage_ms_host, int32_t n_stages) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || !stage_ms_host || n_stages < kNumStages) return -1;
  StageTimer tm;
  tm.st = reinterpret_cast<cudaStream_t>(stream);
  int rc = energy_force_impl(c, pos_dev, e_rep_dev, force_dev, flags, tm.st, &tm);
  cudaStreamSynchronize(tm.st);
  for (int i = 0; i < n_stages; ++i) stage_ms_host[i] = 0.0f;
  for (size_t i = 0; i + 1 < tm.ev.size(); ++i) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, tm.ev[i], tm.ev[i + 1]);
    if (tm.stage[i] >= 0) stage_ms_host[tm.stage[i]] += ms;
  }
  for (cudaEvent_t e : tm.ev) cudaEventDestroy(e);
  return rc;
}

}  // extern "C"
