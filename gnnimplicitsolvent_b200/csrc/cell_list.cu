// cell_list.cu -- cell-list back-end of the sorted-CSR radius-graph builder (sm_100a).
//
// Replaces torch_cluster.radius_graph (data path, GNN_Models.py:345-357 / 887-905) and the `edge_attributes < 0.6`
// compaction of the run path (GNN_Models.py:794-804) for replicas that are much wider than the cutoff; produces the
// same bytes as the all-pairs builder of pair_kernels.cu (sorted by target, then source; same fp32 eps-distances).
//
//   cell_mask_kernel   one CTA per replica and chunk of 128 targets (every CTA bins the whole replica -- O(N), cheap next
//                      to the candidate tests -- so that large molecules in few replicas still fill the GPU): bounding
//                      box -> uniform grid whose cells are at least 1.001 x cutoff wide (at most 8 x 8 x 8 cells; wider
//                      boxes get wider cells) -> counting sort of the atoms by cell in shared memory -> one warp per
//                      target walks the 3 x 3 runs of x-adjacent neighbour cells (their atoms are contiguous in the
//                      sorted list), applies the SAME predicate function as the all-pairs builder to every candidate and
//                      records the accepted sources in a bit mask of N bits.  The mask makes the result independent of
//                      the order in which cells and lanes deliver the candidates, so no sort is needed and nothing
//                      depends on the order of the shared-memory atomics (integers only).
//                      Output: mask rows [n][ceil(N/32)] and the degree of every target.
//   deg_block_sum_kernel, scan_blocks_kernel   CSR offsets in the block layout the fill kernels share
//   mask_fill_kernel   one warp per target: lane l owns mask word l, a warp scan of the pop counts gives every lane its
//                      offset, the set bits are emitted in ascending order = ascending source; the eps-distance is
//                      evaluated only for accepted pairs.
//
// A pair accepted by either predicate is closer than cutoff * (1 + 1e-5) in every coordinate, cells are at least
// 1.001 x cutoff wide, so its two atoms sit in the same or in adjacent cells: the candidate set is a superset of the edge
// set and the edge set itself is decided by in_cutoff() alone -- bit-identical to the all-pairs builder by construction
// (tests/test_cell_list.py).
#include "common.cuh"

namespace gnnis {

int launch_scan_blocks(Ctx* c, cudaStream_t st);

constexpr int kCellDim = 8;                              // cells per dimension, at most
constexpr int kCellMax = kCellDim * kCellDim * kCellDim;
constexpr int kCellThreads = 512;                       // 16 warps: 8 targets per warp and chunk
constexpr int kCellWarps = kCellThreads / 32;
constexpr int kCellChunk = 128;                          // targets per CTA (blockIdx.y)

// dynamic smem: pos[3N] f32 | order[N] u16 | cell_of[N] u16  (+ static: cell starts, fill counters, warp masks, box)
__global__ void __launch_bounds__(kCellThreads)
cell_mask_kernel(const float* __restrict__ pos, int N, int W, float cutoff, int predicate, const int* __restrict__ skip,
                 uint32_t* __restrict__ mask, int* __restrict__ deg) {
  extern __shared__ float sm[];
  float* spos = sm;
  uint16_t* order = reinterpret_cast<uint16_t*>(spos + 3 * N);
  uint16_t* cell_of = order + N;
  __shared__ int cstart[kCellMax + 1];
  __shared__ int cfill[kCellMax];
  __shared__ uint32_t wmask[kCellWarps][32];
  __shared__ float red[2][3][kCellWarps];
  __shared__ float box[2][3];
  const int rep = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t base = (size_t)rep * N;
  const int t_begin = blockIdx.y * kCellChunk, t_end = min(N, t_begin + kCellChunk);   // this CTA's targets
  if (skip && skip[rep] == 2) {   // finished replica of the batched minimiser: no edges
    for (int a = t_begin + tid; a < t_end; a += kCellThreads) deg[base + a] = 0;
    for (int t = t_begin * W + tid; t < t_end * W; t += kCellThreads) mask[base * W + t] = 0u;
    return;
  }
  // ---- positions and bounding box
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int t = tid; t < 3 * N; t += kCellThreads) spos[t] = pos[base * 3 + t];
  __syncthreads();
  for (int a = tid; a < N; a += kCellThreads) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float x = spos[3 * a + d];
      lo[d] = fminf(lo[d], x);
      hi[d] = fmaxf(hi[d], x);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) {
      red[0][d][warp] = lo[d];
      red[1][d][warp] = hi[d];
    }
  }
  for (int t = tid; t < kCellMax; t += kCellThreads) cfill[t] = 0;
  __syncthreads();
  if (tid < 3) {
    float l = red[0][tid][0], h = red[1][tid][0];
    for (int w = 1; w < kCellWarps; ++w) {
      l = fminf(l, red[0][tid][w]);
      h = fmaxf(h, red[1][tid][w]);
    }
    box[0][tid] = l;
    box[1][tid] = h;
  }
  __syncthreads();
  // ---- grid: g_d cells of width ext_d / g_d >= 1.001 cutoff
  int g[3];
  float scale[3], org[3];
  const float h0 = 1.001f * cutoff;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    org[d] = box[0][d];
    const float ext = box[1][d] - box[0][d];
    int gd = (ext > 0.0f && ext < 1.0e30f) ? (int)(ext / h0) : 1;   // non-finite boxes: one cell (the predicate decides)
    gd = gd < 1 ? 1 : (gd > kCellDim ? kCellDim : gd);
    g[d] = gd;
    scale[d] = gd > 1 ? (float)gd / ext : 0.0f;
  }
  const int ncell = g[0] * g[1] * g[2];
  auto cell_coord = [&](float x, int d) {
    int cc = (int)((x - org[d]) * scale[d]);
    return cc < 0 ? 0 : (cc > g[d] - 1 ? g[d] - 1 : cc);
  };
  // ---- counting sort of the atoms by cell (x fastest)
  for (int a = tid; a < N; a += kCellThreads) {
    const int cid = (cell_coord(spos[3 * a + 2], 2) * g[1] + cell_coord(spos[3 * a + 1], 1)) * g[0] + cell_coord(spos[3 * a], 0);
    cell_of[a] = (uint16_t)cid;
    atomicAdd(&cfill[cid], 1);
  }
  __syncthreads();
  if (warp == 0) {   // exclusive scan of <= 512 counts: 16 per lane
    int loc[kCellMax / 32];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kCellMax / 32; ++k) {
      const int cidx = lane * (kCellMax / 32) + k;
      loc[k] = cidx < ncell ? cfill[cidx] : 0;
      sum += loc[k];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int run = incl - sum;
#pragma unroll
    for (int k = 0; k < kCellMax / 32; ++k) {
      const int cidx = lane * (kCellMax / 32) + k;
      if (cidx <= ncell) cstart[cidx] = run;
      run += loc[k];
      if (cidx < kCellMax) cfill[cidx] = 0;
    }
    if (lane == 31 && ncell == kCellMax) cstart[kCellMax] = run;
  }
  __syncthreads();
  for (int a = tid; a < N; a += kCellThreads) {
    const int cid = cell_of[a];
    order[cstart[cid] + atomicAdd(&cfill[cid], 1)] = (uint16_t)a;   // order inside a cell is irrelevant (bit mask below)
  }
  __syncthreads();
  // ---- one warp per target
  for (int a = t_begin + warp; a < t_end; a += kCellWarps) {
    wmask[warp][lane] = 0u;
    __syncwarp();
    const float tx = spos[3 * a], ty = spos[3 * a + 1], tz = spos[3 * a + 2];
    const int cx = cell_coord(tx, 0), cy = cell_coord(ty, 1), cz = cell_coord(tz, 2);
    const int x0 = cx > 0 ? cx - 1 : 0, x1 = cx < g[0] - 1 ? cx + 1 : g[0] - 1;
    for (int zz = (cz > 0 ? cz - 1 : 0); zz <= (cz < g[2] - 1 ? cz + 1 : g[2] - 1); ++zz) {
      for (int yy = (cy > 0 ? cy - 1 : 0); yy <= (cy < g[1] - 1 ? cy + 1 : g[1] - 1); ++yy) {
        const int row = (zz * g[1] + yy) * g[0];
        const int k0 = cstart[row + x0], k1 = cstart[row + x1 + 1];   // x-adjacent cells are contiguous in `order`
        for (int k = k0 + lane; k < k1; k += 32) {
          const int j = order[k];
          if (j == a) continue;
          float dx, dy, dz;
          const float sx = spos[3 * j], sy = spos[3 * j + 1], sz = spos[3 * j + 2];
          const float r = eps_dist(sx, sy, sz, tx, ty, tz, dx, dy, dz);
          if (in_cutoff(predicate, r, sx, sy, sz, tx, ty, tz, cutoff)) atomicOr(&wmask[warp][j >> 5], 1u << (j & 31));
        }
      }
    }
    __syncwarp();
    const uint32_t word = wmask[warp][lane];
    int cnt = __popc(word);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane < W) mask[(base + a) * W + lane] = word;
    if (lane == 0) deg[base + a] = cnt;
    __syncwarp();
  }
}

// blk_sum[b] = sum of deg over the kPairThreads atoms of block b (the layout fill_edges_kernel / mask_fill_kernel scan)
__global__ void __launch_bounds__(kPairThreads) deg_block_sum_kernel(const int* __restrict__ deg, int n, int* __restrict__ blk_sum) {
  __shared__ int part[kPairThreads / 32];
  const int i = blockIdx.x * kPairThreads + threadIdx.x;
  int v = i < n ? deg[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kPairThreads / 32; ++w) s += part[w];
    blk_sum[blockIdx.x] = s;
  }
}

// dynamic smem: pos of the replicas this block's atoms belong to.  One block = kPairThreads consecutive targets (the layout
// of blk_sum) worked on by kFillThreads threads: the first kPairThreads compute the row offsets, then the 16 warps take the
// block's targets round-robin (a target is one mask row per warp: short, latency-bound work that needs many warps in flight)
constexpr int kFillThreads = 512;
__global__ void __launch_bounds__(kFillThreads)
mask_fill_kernel(const float* __restrict__ pos, int n, int N, int G, int W, const uint32_t* __restrict__ mask,
                 const int* __restrict__ deg, const int* __restrict__ blk_off, int64_t capacity, int* __restrict__ row_ptr,
                 int* __restrict__ col, uint32_t* __restrict__ pair, float* __restrict__ r_e) {
  extern __shared__ float sm[];
  __shared__ int sdeg[kPairThreads];
  __shared__ int soff[kPairThreads];
  const int a0 = blockIdx.x * kPairThreads, a1 = min(n, a0 + kPairThreads) - 1;
  const int first_rep = a0 / N, last_rep = a1 / N;
  const int n_stage = (last_rep - first_rep + 1) * N;
  float* spos = sm;
  for (int t = threadIdx.x; t < 3 * n_stage; t += kFillThreads) spos[t] = pos[(size_t)first_rep * N * 3 + t];
  const int i = a0 + threadIdx.x;
  const int mydeg = (threadIdx.x < kPairThreads && i < n) ? deg[i] : 0;
  if (threadIdx.x < kPairThreads) sdeg[threadIdx.x] = mydeg;
  __syncthreads();
  for (int off = 1; off < kPairThreads; off <<= 1) {
    int v = (threadIdx.x < kPairThreads && threadIdx.x >= off) ? sdeg[threadIdx.x - off] : 0;
    __syncthreads();
    if (threadIdx.x < kPairThreads) sdeg[threadIdx.x] += v;
    __syncthreads();
  }
  if (threadIdx.x < kPairThreads) {
    int w = -1;   // -1: nothing to emit (past the end, empty row, or the row does not fit the caller's capacity)
    if (i < n) {
      const int64_t w64 = (int64_t)blk_off[blockIdx.x] + sdeg[threadIdx.x] - mydeg;
      row_ptr[i] = (int)w64;
      if (mydeg > 0 && w64 + mydeg <= capacity) w = (int)w64;
    }
    soff[threadIdx.x] = w;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < kPairThreads; t += kFillThreads / 32) {
    const int wt = soff[t];
    if (wt < 0) continue;
    const int it = a0 + t;
    const int rep = it / N, a = it - rep * N;
    const float* rp = spos + (size_t)(rep - first_rep) * N * 3;
    const float tx = rp[3 * a], ty = rp[3 * a + 1], tz = rp[3 * a + 2];
    const uint32_t ul_base = (uint32_t)(rep % G) * (uint32_t)N;
    const uint32_t tgt_ul = (ul_base + (uint32_t)a) << 16;
    uint32_t word = lane < W ? mask[(size_t)it * W + lane] : 0u;
    const int cnt = __popc(word);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int64_t at = (int64_t)wt + (incl - cnt);
    for (; word; word &= word - 1, ++at) {
      const int j = 32 * lane + __ffs(word) - 1;
      float dx, dy, dz;
      const float r = eps_dist(rp[3 * j], rp[3 * j + 1], rp[3 * j + 2], tx, ty, tz, dx, dy, dz);
      if (col) col[at] = rep * N + j;
      if (pair) pair[at] = tgt_ul | (ul_base + (uint32_t)j);
      if (r_e) r_e[at] = r;
    }
  }
}

// degrees + CSR offsets + sorted fill through the cell list; same outputs as launch_born_count's degree part followed by
// launch_fill_edges (pair_kernels.cu)
int launch_cell_neighbors(Ctx* c, const float* pos, int predicate, int* row_ptr, int* col, uint32_t* pair, float* r_e,
                          int64_t capacity, cudaStream_t st) {
  const int n = c->R * c->N, N = c->N, W = (N + 31) / 32;
  const int nblk = (n + kPairThreads - 1) / kPairThreads;
  const size_t smem_mask = (size_t)N * (3 * sizeof(float) + 2 * sizeof(uint16_t));
  cell_mask_kernel<<<dim3(c->R, (N + kCellChunk - 1) / kCellChunk), kCellThreads, smem_mask, st>>>(pos, N, W, c->cutoff, predicate, c->skip_state, c->nbr_mask, c->deg);
  GNNIS_LAUNCH_CHECK();
  deg_block_sum_kernel<<<nblk, kPairThreads, 0, st>>>(c->deg, n, c->blk_sum);
  GNNIS_LAUNCH_CHECK();
  if (launch_scan_blocks(c, st)) return -1;
  const size_t smem_fill = 3 * ((size_t)kPairThreads + 2 * (size_t)N) * sizeof(float);
  if (smem_fill > 48 * 1024) cudaFuncSetAttribute(mask_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fill);
  mask_fill_kernel<<<nblk, kFillThreads, smem_fill, st>>>(pos, n, N, c->G, W, c->nbr_mask, c->deg, c->blk_sum, capacity, row_ptr,
                                                         col, pair, r_e);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

}  // namespace gnnis
