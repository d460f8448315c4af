// edge_mlp_tc.cu -- the edge-message MLP and its adjoint on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Formulation (per directed edge e = (j -> i), IN_layer_all_swish_2pass_tokens.message, GNN_Layers.py:2186-2192):
//     u = P[i] + Q[j] + rbf(r_e) Wr^T        P = h W1[:, :cin]^T + b1, Q = h W1[:, cin:2cin]^T  (node kernels)
//     v = silu(u);  w = v W2^T + b2;  z = silu(w);  a[i] += z
// and the adjoint, recomputing u, w tile by tile instead of storing per-edge activations:
//     wbar = abar[i] * silu'(w);  vbar = wbar W2;  ubar = vbar * silu'(u)
//     Pbar[i] += ubar;  Qbar[j] += ubar;  rbar_e = sum_c ubar_c (d rbf/dr Wr^T)_c
//
// Work decomposition: one persistent CTA per SM made of TWO independent groups of 8 warps.  Each group owns its
// own units (a unit = G replicas whose tables fit in shared memory), its own 128-edge tile, its own 256 TMEM
// columns, its own mbarriers / arrival counters and named barrier -- the groups never synchronise with each other,
// so while one group waits for its tcgen05.mma the other one runs its SiLU epilogue and vice versa.  Inside a group
// a thread is (edge row = TMEM lane, 32-column half): warps 0-3 take channels 0..31, warps 4-7 channels 32..63, and
// each walks its channels in two chunks of 16 columns (tcgen05.ld/st.32x32b.x16).
//
// Every dense contraction of a tile is tcgen05.mma.kind::tf32 with the accumulator in TMEM and the A operand
// written back to TMEM by the previous epilogue (tcgen05.st), i.e. A never touches shared memory:
//     u    = rbf  . Wr^T   [128 x 24] x [24 x 64]
//     du   = drbf . Wr^T   [128 x 24] x [24 x 64]   (adjoint only: d u / d r, dotted with ubar per row)
//     w    = v    . W2^T   [128 x 64] x [64 x 64]
//     vbar = wbar . W2     [128 x 64] x [64 x 64]   (W2^T staged as a second K-major operand: kind::tf32 only
//                                                    reads MN-major operands in the 128B_BASE32B layout)
// Weights are staged once per CTA in the canonical K-major 128B-swizzle layout: prebuilt hi / lo images
// (build_edge_images, once per weight load) fetched with one cp.async.bulk per 32 KB.
//
// Precision: fp32-accurate.  Each product is three TF32 MMAs, A_hi B_hi + A_hi B_lo + A_lo B_hi with x_hi = top 19
// bits of x and x_lo = x - x_hi ("3xTF32"), accumulated in fp32 in TMEM; the dropped term is O(2^-22).
//
// Reductions are deterministic (no atomics).  Rows of a tile with equal target are contiguous (CSR order), rows
// with equal source are contiguous in the precomputed source-sorted permutation of the tile; one warp builds the
// list of segment starts while the first MMA runs, then every segment is summed in a fixed order by a fixed set of
// threads.  Target sums go straight to global memory (a row is complete unless it straddles a tile boundary: then
// the next tile adds to it), source sums into a shared-memory table of the unit.  The reductions of tile t run
// while the tensor pipe works on tile t+1.
#include "common.cuh"
#include "umma.cuh"
#include "fastmath.cuh"

namespace gnnis {

using namespace umma;

// values of the TERMS template parameter that select the half-precision modes: 16-bit tensor-core operands
// (tcgen05.mma.kind::f16, fp32 accumulate) and tanh-based activations.  kBf16: GNNIS_FLAG_MLP_BF16; kFp16: the same
// kernels with fp16 operands (GNNIS_FLAG_MLP_FP16: 3 more mantissa bits; every operand of these MLPs fits fp16's range)
constexpr int kBf16 = 2, kFp16 = 4;
constexpr int kTanh3 = 7;   // forward only: 3xTF32 products with the tanh-based activations (the default, Ctx::act_tanh)
__host__ __device__ constexpr bool is_half(int terms) { return terms == kBf16 || terms == kFp16; }
template <int TERMS>
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  return TERMS == kFp16 ? pack_f16x2(a, b) : pack_bf16x2(a, b);
}
constexpr int kTabLd = 68;        // padded row of the shared-memory node tables (bank spread for row-per-lane gathers)
constexpr int kGroups = 2;

struct EdgeTcArgs {
  int n_units, G, N, R;
  const int* row_ptr;
  const uint32_t* pair;
  const float* r_e;
  float inv_cutoff;
  const uint8_t* img;          // operand image: WrH 8K | WrL 8K | W2H 16K | W2L 16K | W2tH 16K | W2tL 16K
  const float* b2;
  const float *P, *Q;
  float* a_out;
  const float* abar;
  float *Pbar, *Qbar, *rbar;
  int rbar_accumulate;
  const uint8_t* perm;
  int unit_atoms;
  unsigned* unit_counter; // dynamic unit hand-out of the adjoint, zeroed at the start of every evaluation
  int S;                  // sub-units per replica (Ctx::split_eff)
  int U0;                 // first split replica (Ctx::split_from): units [0, U0) are whole replicas
  size_t qbar_stride;     // floats between the partial Qbar tables of consecutive parts
  // activation stash (forward with forces -> adjoint): 16-bit codes of silu'(u) and silu'(w), [E][64] each, row e =
  // 128 bytes whose 16-byte chunks are XOR-swizzled with (row in tile & 7) -- the shared-memory image of a tile
  uint8_t *st_u, *st_w;
  // capacity-limited stash (Ctx::stash_cap < edge_cap): edge count of this evaluation and the edges the stash holds.
  // n_edges == nullptr: the stash holds every possible edge list.  stash_role: 0 = not a stash-dependent launch,
  // 1 = runs only when the edge list fits the stash (stash adjoint), 2 = runs only when it does not (recomputing adjoint)
  const int64_t* n_edges;
  int64_t stash_cap;
  int stash_role;
};
// does the edge list of this evaluation fit the stash?  (uniform over the grid: one 8-byte load per thread)
__device__ __forceinline__ bool stash_fits(const EdgeTcArgs& g) { return !g.n_edges || *g.n_edges <= g.stash_cap; }

// ---------------------------------------------------------------------------------------------- helpers
// stage a K-major B operand [rows][kdim] (element (n,k) = src[n*ld_n + k*ld_k], zero outside n_valid x k_valid)
// as hi / lo TF32 parts in the SW128 layout
__device__ __forceinline__ void stage_b_operand(uint8_t* hi, uint8_t* lo, const float* __restrict__ src, int rows,
                                                int kdim, int n_valid, int k_valid, int ld_n, int ld_k) {
  for (int t = threadIdx.x; t < rows * kdim; t += blockDim.x) {
    int n = t / kdim, k = t - n * kdim;
    float v = (n < n_valid && k < k_valid) ? src[(size_t)n * ld_n + (size_t)k * ld_k] : 0.0f;
    uint32_t h, l;
    split_tf32(v, h, l);
    uint32_t off = sw128_offset(n, k, rows);
    *reinterpret_cast<uint32_t*>(hi + off) = h;
    *reinterpret_cast<uint32_t*>(lo + off) = l;
  }
}

// D[128 x N] (+)= A[128 x k in [8 S0, 8 S1)] * B^T, three TF32 products; A (hi, lo) in TMEM, B (hi, lo) K-major SW128
// in smem.  `accumulate` = 0 overwrites D with the first product.
template <int S0, int S1, int TERMS = 3>
__device__ __forceinline__ void issue_gemm_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_addr,
                                                  uint32_t b_lo_addr, int b_rows, uint32_t idesc, uint32_t accumulate) {
  uint32_t acc = accumulate;
  if (is_half(TERMS)) {   // 16-bit operands (kind::f16): K = 8 (S1 - S0) in steps of 16 = 8 TMEM columns / 32 bytes of the B row
#pragma unroll
    for (int s = 0; s < (S1 - S0 + 1) / 2; ++s) {
      mma_bf16_ts(d_tmem, a_hi + 8u * (uint32_t)s, smem_desc_k_sw128(b_hi_addr + 32u * (uint32_t)s), idesc, acc);
      acc = 1;
    }
    return;
  }
#pragma unroll
  for (int term = 0; term < TERMS; ++term) {   // TERMS = 1: single TF32 product (opt-in reduced-precision mode)
    const uint32_t a = term == 2 ? a_lo : a_hi;
    const uint32_t b = term == 1 ? b_lo_addr : b_hi_addr;
#pragma unroll
    for (int s = S0; s < S1; ++s) {
      uint32_t baddr = b + (uint32_t)(s >> 2) * (uint32_t)b_rows * 128u + (uint32_t)(s & 3) * 32u;
      mma_tf32_ts(d_tmem, a + 8u * (uint32_t)s, smem_desc_k_sw128(baddr), idesc, acc);
      acc = 1;
    }
  }
}
// the same staging into a GLOBAL image (built once per weight load): block (n0.., k0..) of an operand with rows_total
// rows, element (n0 + n, k0 + k) = src[n * ld_n + k]
__global__ void build_operand_image_kernel(uint8_t* __restrict__ hi, uint8_t* __restrict__ lo,
                                           const float* __restrict__ src, int rows_total, int n0, int nrows, int k0,
                                           int kcount, int ld_n) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nrows * kcount; t += gridDim.x * blockDim.x) {
    int n = t / kcount, k = t - n * kcount;
    uint32_t h, l;
    split_tf32(src[(size_t)n * ld_n + k], h, l);
    uint32_t off = sw128_offset(n0 + n, k0 + k, rows_total);
    *reinterpret_cast<uint32_t*>(hi + off) = h;
    *reinterpret_cast<uint32_t*>(lo + off) = l;
  }
}
int launch_build_operand(Ctx* c, uint8_t* hi, uint8_t* lo, const float* src, int rows_total, int n0, int nrows, int k0,
                         int kcount, int ld_n, cudaStream_t st) {
  build_operand_image_kernel<<<32, 256, 0, st>>>(hi, lo, src, rows_total, n0, nrows, k0, kcount, ld_n);
  GNNIS_LAUNCH_CHECK();
  return 0;
}
__device__ __forceinline__ void st_rbar(float* p, float v) {
#if GNNIS_L1_POLICY >= 2
  __stcg(p, v);
#else
  *p = v;
#endif
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// radial basis row in registers (GNN_Layers.py:2194-2208): d = r/cutoff, env = 1/d - 28 d^5 + 48 d^6 - 21 d^7,
// rbf_k = env sin(k pi d); sin/cos(k pi d) from one sincospi and the Chebyshev recurrence.  o[20..31] = 0.
template <bool DERIV>
__device__ __forceinline__ void rbf_regs(float r, float inv_cutoff, float (&o)[32], float (&od)[32]) {
  const float d = r * inv_cutoff;
  const float d2 = d * d, d4 = d2 * d2, d5 = d4 * d;
  const float env = 1.0f / d - 28.0f * d5 + 48.0f * (d5 * d) - 21.0f * (d5 * d2);
  float denv = 0.0f;
  if (DERIV) denv = -1.0f / d2 - 140.0f * d4 + 288.0f * d5 - 147.0f * (d5 * d);
  float s1, c1;
#ifdef GNNIS_FAST_SINCOS   // A/B: MUFU.SIN / MUFU.COS (absolute error ~5e-7 on [0, pi]) instead of the library sincospif
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(3.14159265358979323846f * d));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(3.14159265358979323846f * d));
#else
  sincospif(d, &s1, &c1);
#endif
  const float twoc = 2.0f * c1;
  float sp = 0.0f, sc = s1, cp = 1.0f, cc = c1;
#pragma unroll
  for (int k = 0; k < kRbf; ++k) {
    o[k] = env * sc;
    if (DERIV) od[k] = (denv * sc + env * (3.14159265358979323846f * (float)(k + 1)) * cc) * inv_cutoff;
    float sn = fmaf(twoc, sc, -sp);
    sp = sc;
    sc = sn;
    if (DERIV) {
      float cn = fmaf(twoc, cc, -cp);
      cp = cc;
      cc = cn;
    }
  }
#pragma unroll
  for (int k = kRbf; k < 32; ++k) {
    o[k] = 0.0f;
    if (DERIV) od[k] = 0.0f;
  }
}

__device__ __forceinline__ void st_split(uint32_t t_hi, uint32_t t_lo, const float (&x)[32]) {
  uint32_t h[32], l[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) split_tf32(x[j], h[j], l[j]);
  tmem_st32(t_hi, h);
  tmem_st32(t_lo, l);
}

// radial operand rows: only the 24 columns that the K = 24 products read (20 basis functions + 4 zeros)
template <int TERMS = 3>
__device__ __forceinline__ void st_split24(uint32_t t_hi, uint32_t t_lo, const float (&x)[32]) {
  if (is_half(TERMS)) {   // 24 values -> 12 words of two bf16 / fp16, padded with zeros to K = 32
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 12; ++j) w[j] = pack_half2<TERMS>(x[2 * j], x[2 * j + 1]);
#pragma unroll
    for (int j = 12; j < 16; ++j) w[j] = 0u;
    tmem_st16(t_hi, w);
    return;
  }
  if (TERMS == 1) {   // single-product mode: the raw fp32 words (kind::tf32 reads their top 19 bits), no lo part
    uint32_t h[16], h8[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = __float_as_uint(x[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) h8[j] = __float_as_uint(x[16 + j]);
    tmem_st16(t_hi, h);
    tmem_st8(t_hi + 16, h8);
    return;
  }
  {
    uint32_t h[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) split_tf32(x[j], h[j], l[j]);
    tmem_st16(t_hi, h);
    tmem_st16(t_lo, l);
  }
  {
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_tf32(x[16 + j], h[j], l[j]);
    tmem_st8(t_hi + 16, h);
    tmem_st8(t_lo + 16, l);
  }
}

// Per-tile keys (double-buffered by tile parity so that the reduction of tile t may overlap the set-up of t+1)
struct TileKeys {
  int tgt[kTileE];        // unit-local target of row e (-1 past the end of the unit)
  int src[kTileE];        // unit-local source
  uint8_t perm[kTileE];   // row at source-sorted position p
  uint8_t seg_t[kTileE + 8];   // starts of the target segments (row order), then the number of valid rows
  uint8_t seg_s[kTileE + 8];   // starts of the source segments (permuted order), then the number of valid rows
  int n_seg_t, n_seg_s, cont_prev, pad;
};

// One warp: list the starts of the runs of equal keys among the valid rows (valid rows come first).
template <bool PERM>
__device__ __forceinline__ int build_segments(const int* __restrict__ key, const uint8_t* __restrict__ perm,
                                              uint8_t* __restrict__ seg, int n_valid, int lane) {
  int base = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int p = 32 * c + lane;
    const int k = key[PERM ? (int)perm[p] : p];
    const int kprev = p > 0 ? key[PERM ? (int)perm[p - 1] : p - 1] : -2;
    const bool flag = p < n_valid && k != kprev;
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    if (flag) seg[base + __popc(b & ((1u << lane) - 1u))] = (uint8_t)p;
    base += __popc(b);
  }
  if (lane == 0) seg[base] = (uint8_t)n_valid;   // <= 128
  return base;
}

// ---------------------------------------------------------------------------------------------- self test
// D[128][N] = A[128][K] * B[N][K]^T through the exact building blocks used below (TMEM-resident A, SW128 B,
// 3xTF32).  K in {32, 64}, N in {32, 64}.  One CTA of 256 threads.
__global__ void __launch_bounds__(256, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                              float* __restrict__ D, int K, int N) {
  extern __shared__ uint8_t smraw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t base = smem_u32(smraw);
  uint8_t* sm = smraw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t* bh = sm;
  uint8_t* bl = sm + 16384;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, ch = warp >> 2;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) mbar_init(&bar, 1);
  stage_b_operand(bh, bl, B, N, K, N, K, K, 1);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_slot;
  const uint32_t lane_addr = tm + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t ACC = 0, AH = 64, AL = 128;
  if (32 * ch < K) {
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = A[(size_t)row * K + 32 * ch + j];
    st_split(lane_addr + AH + 32 * ch, lane_addr + AL + 32 * ch, x);
  }
  tmem_st_wait();
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    if (K == 64) issue_gemm_3xtf32<0, 8>(tm + ACC, tm + AH, tm + AL, smem_u32(bh), smem_u32(bl), N, idesc_tf32(128, N), 0u);
    else issue_gemm_3xtf32<0, 4>(tm + ACC, tm + AH, tm + AL, smem_u32(bh), smem_u32(bl), N, idesc_tf32(128, N), 0u);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  __syncwarp();
  fence_after_sync();
  if (32 * ch < N) {
    uint32_t r[32];
    tmem_ld32(lane_addr + ACC + 32 * ch, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + 32 * ch + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------- shared memory map
// common (1024-aligned): WrH 8K | WrL 8K | W2H 16K | W2L 16K | [adjoint: W2tH 16K | W2tL 16K] | b2[64]
// per group:             X[128][68] | TileKeys[2] | tables
//   forward tables  : Q          [unit_atoms][68]
//   adjoint tables  : Q, Qbar    [unit_atoms][68] each
constexpr uint32_t kFwdWeights = 49152, kBwdWeights = 81920;
constexpr uint32_t kGrpX = 0;
constexpr uint32_t kGrpStage = kGrpX + kTileE * kTileLd * 4;              // stash kernels: two 16 KB tile images
constexpr uint32_t kStageBytes = 2 * kTileE * 128;
constexpr uint32_t kGrpKeys = kGrpStage;                                  // (+ kStageBytes in the stash kernels)
constexpr uint32_t kGrpTab = kGrpKeys + 3 * (uint32_t)sizeof(TileKeys);   // ws kernels rotate three key buffers
static_assert(sizeof(TileKeys) % 16 == 0, "TileKeys must keep the tables 16-byte aligned");
static_assert(kGrpStage % 128 == 0, "tile images must be 128-byte aligned");

// shared-memory tables per group (Ctx::tab_mode): 2 = both tables of the kernel (forward Q, P; adjoint Q, Qbar),
// 1 = one (forward Q; adjoint Qbar) with the other rows gathered from global memory / L1, 0 = none
static size_t group_bytes_mode(const Ctx* c, int mode) {
  size_t b = kGrpTab + (size_t)mode * c->G * c->N * kTabLd * 4;
  return (b + 15) / 16 * 16;
}
static size_t group_bytes(const Ctx* c) { return group_bytes_mode(c, c->tab_mode); }
static size_t fwd_tc_smem(const Ctx* c) { return kFwdWeights + 256 + kGroups * group_bytes(c) + 1024; }
static size_t bwd_tc_smem(const Ctx* c) { return kBwdWeights + 256 + kGroups * group_bytes(c) + 1024; }
// largest table mode whose adjoint kernel (the bigger one) fits into 227 KB
int tc_table_mode(const Ctx* c) {
  for (int mode = 2; mode >= 1; --mode)
    if (kBwdWeights + 256 + kGroups * group_bytes_mode(c, mode) + 1024 + 2048 <= 227 * 1024) return mode;   // 2 KB: static smem
  return 0;
}

// stash kernels: X | two tile images | keys | tables (forward: Q [, P]; adjoint: Qbar); group size a multiple of 128.
// Forward table mode 3 = Q table + ONE image buffer that the silu'(u) and silu'(w) codes of a tile use in turn (the
// second image is what keeps the Q table of a 67 ... 126-atom replica out of shared memory; edge_fwd_wsi_kernel).
static size_t st_group_bytes(const Ctx* c, int tables) {
  const size_t images = tables == 3 ? kStageBytes / 2 : kStageBytes;
  size_t b = kGrpTab + images + (size_t)(tables == 3 ? 1 : tables) * c->G * c->N * kTabLd * 4;
  return (b + 127) / 128 * 128;
}
constexpr size_t kDynSmemMax = 227 * 1024 - 2048;   // 2 KB: static shared memory of the kernels
static size_t st_fwd_smem(const Ctx* c, int tables) { return kFwdWeights + 256 + kGroups * st_group_bytes(c, tables) + 1024; }
static size_t st_bwd_smem(const Ctx* c, int tables) { return 49152 + 256 + kGroups * st_group_bytes(c, tables) + 1024; }
int st_fwd_tables(const Ctx* c) {
  const int order[4] = {2, 1, 3, 0};
  for (int t : order)
    if ((t != 3 || c->one_buf) && st_fwd_smem(c, t) <= kDynSmemMax) return t;
  return -1;
}
int st_bwd_tables(const Ctx* c) {
  for (int t = 1; t >= 0; --t)
    if (st_bwd_smem(c, t) <= kDynSmemMax) return t;
  return -1;
}

// ---------------------------------------------------------------------------------------------- launchers
static EdgeTcArgs make_tc_args(Ctx* c, int l, int terms = 3) {
  EdgeTcArgs g{};
  g.n_units = c->eff_units();
  g.S = c->split_eff;
  g.U0 = c->split_eff > 1 ? c->split_from : 0;
  g.unit_counter = c->unit_ctr ? c->unit_ctr + l : nullptr;
  g.qbar_stride = (size_t)c->R * c->N * 64;
  g.G = c->G;
  g.N = c->N;
  g.R = c->R;
  g.row_ptr = c->row_ptr;
  g.pair = c->pair;
  g.r_e = c->r_e;
  g.inv_cutoff = (float)(1.0 / (double)c->cutoff);
  g.img = terms == kBf16 ? c->img_edge_bf[l] : (terms == kFp16 ? c->img_edge_bf[l] + 24576 : c->img_edge[l]);
  g.b2 = c->L[l].b2;
  g.P = c->P[l];
  g.Q = c->Q[l];
  g.a_out = c->a[l];
  g.abar = c->abar;
  g.Pbar = c->Pbar;
  g.Qbar = c->Qbar;
  g.rbar = c->rbar;
  g.perm = c->perm;
  g.unit_atoms = c->G * c->N;
  g.st_u = c->st_u[l];
  g.st_w = c->st_w[l];
  g.n_edges = (c->stash_on && c->stash_cap < c->edge_cap) ? c->n_edges_dev : nullptr;
  g.stash_cap = c->stash_cap;
  g.stash_role = 0;
  return g;
}

// bf16 image of a K-major operand block (element (n, k) = src[n * ld_n + k], round to nearest even; the image is
// zero-filled beforehand, so padded K columns stay zero)
__global__ void build_operand_image_bf16_kernel(uint8_t* __restrict__ img, const float* __restrict__ src, int rows_total,
                                                int nrows, int kcount, int ld_n, int fp16) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nrows * kcount; t += gridDim.x * blockDim.x) {
    int n = t / kcount, k = t - n * kcount;
    const float v = src[(size_t)n * ld_n + k];
    const uint32_t w = fp16 ? pack_f16x2(v, 0.0f) : pack_bf16x2(v, 0.0f);
    *reinterpret_cast<uint16_t*>(img + sw128_offset_bf16(n, k, rows_total)) = (uint16_t)(w & 0xFFFFu);
  }
}

int build_edge_images(Ctx* c, cudaStream_t st) {
  for (int l = 0; l < 3; ++l) {
    for (int fp16 = 0; fp16 < 2; ++fp16) {   // half-precision modes: Wr | W2 | W2^T, 8 KB each, as bf16 and as fp16
      uint8_t* bf = c->img_edge_bf[l] + (size_t)fp16 * 24576;
      build_operand_image_bf16_kernel<<<32, 256, 0, st>>>(bf, c->L[l].Wr, 64, 64, kRbf, kRbf, fp16);
      build_operand_image_bf16_kernel<<<32, 256, 0, st>>>(bf + 8192, c->L[l].W2, 64, 64, 64, 64, fp16);
      build_operand_image_bf16_kernel<<<32, 256, 0, st>>>(bf + 16384, c->L[l].W2_t, 64, 64, 64, 64, fp16);
      GNNIS_LAUNCH_CHECK();
    }
    uint8_t* im = c->img_edge[l];
    if (launch_build_operand(c, im, im + 8192, c->L[l].Wr, 64, 0, 64, 0, kRbf, kRbf, st) ||
        launch_build_operand(c, im + 16384, im + 32768, c->L[l].W2, 64, 0, 64, 0, 64, 64, st) ||
        launch_build_operand(c, im + 49152, im + 65536, c->L[l].W2_t, 64, 0, 64, 0, 64, 64, st))
      return -1;
  }
  return 0;
}

bool tc_path_fits(const Ctx* c) { return fwd_tc_smem(c) <= 227 * 1024 && bwd_tc_smem(c) <= 227 * 1024; }

// Units are dealt out CTA-first (unit u -> CTA u % grid, group (u / grid) % 2), so that with fewer units than
// 2 x #SM every unit gets an SM of its own before any SM runs two groups.
static int tc_grid(const Ctx* c) {
  const int nu = c->eff_units();
  return nu < c->num_sms ? nu : c->num_sms;
}

// ================================================================================================================
// 16-warp variant (default): 2 groups x 8 warps per CTA, no group-wide barrier on the MMA hand-off.
//
// A thread is (edge row, 32-column half): warps 0-3 of a group take channels 0..31, warps 4-7 channels 32..63 of
// the same 128 rows, so 16 warps share the four schedulers of the SM instead of 8.  Operands are handed to the
// tensor pipe by arrival counters in shared memory: every warp arrives once its part of an A operand is in TMEM
// (or once it has drained an accumulator), and the warp whose arrival completes the count issues the tcgen05.mma
// batch itself (one elected lane, warp converged) -- nobody waits for a dedicated issuer or for the slowest warp
// before carrying on with independent work:
//     cnt_u  (8 arrivals)  : accumulators drained (8 warps), rbf operand of the next tile written (4 of them)
//     cnt_a  ( 8 arrivals) : A operand (v, wbar) written by all warps
//     d_full ( tcgen05.commit -> mbarrier ) : products of the last issue are in TMEM
// Only two group-wide named barriers per tile remain, both around the shared-memory tile of the reduction.
constexpr int kWsGroupThreads = 256;
constexpr int kWsThreads = kGroups * kWsGroupThreads;

struct WsBars {
  uint64_t d_full[kGroups], a_full[kGroups];
  int cnt_a[kGroups], cnt_u[kGroups];
};

__device__ __forceinline__ void ws_group_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }
// One arrival per warp, after the warp's outstanding TMEM traffic has been waited for by the caller.  Returns true
// (warp-uniform) in the warp that completes the count; that warp then issues the MMAs that depended on it.
__device__ __forceinline__ bool warp_arrive_last(int* cnt, int expected, int lane) {
  fence_before_sync();
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    // acq_rel at CTA scope orders the counter protocol; TMEM traffic is ordered by the tcgen05 fences around it
    int old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
    if (old == expected - 1) {
      asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(cnt)), "r"(0) : "memory");
      last = 1;
    }
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) fence_after_sync();
  return last != 0;
}
__device__ __forceinline__ void warp_arrive_mbar(uint64_t* bar, int lane) {
  if (lane == 0) mbar_arrive(bar);
}
template <int TERMS = 3>
__device__ __forceinline__ void st_split16(uint32_t t_hi, uint32_t t_lo, const float (&x)[16]) {
  if (TERMS == 1) {
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(x[j]);
    tmem_st16(t_hi, w);
    return;
  }
  uint32_t h[16], l[16];
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    h[j] = __float_as_uint(x[j]) & 0xFFFFE000u;
    h[j + 1] = __float_as_uint(x[j + 1]) & 0xFFFFE000u;
    float l0, l1;
    upk(sub2(pk(x[j], x[j + 1]), pk(__uint_as_float(h[j]), __uint_as_float(h[j + 1]))), l0, l1);
    l[j] = __float_as_uint(l0);
    l[j + 1] = __float_as_uint(l1);
  }
  tmem_st16(t_hi, h);
  tmem_st16(t_lo, l);
}
// 16 channels [col0, col0 + 16) of an A operand row: hi / lo TF32 words at base + col0 (one TMEM column per channel),
// or 8 words of two bf16 at base + col0 / 2 in the bf16 mode
template <int TERMS>
__device__ __forceinline__ void st_operand16(uint32_t lane_addr, uint32_t base_hi, uint32_t base_lo, int col0, const float (&x)[16]) {
  if (is_half(TERMS)) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = pack_half2<TERMS>(x[2 * j], x[2 * j + 1]);
    tmem_st8(lane_addr + base_hi + (uint32_t)(col0 >> 1), w);
  } else {
    st_split16<TERMS>(lane_addr + base_hi + col0, lane_addr + base_lo + col0, x);
  }
}
__device__ __forceinline__ void copy_rows_n(float* __restrict__ dst, int ld_dst, const float* __restrict__ src,
                                            int ld_src, int nrows, int t0, int nthreads) {
  for (int t = t0; t < nrows * 16; t += nthreads) {
    int r = t >> 4, c4 = t & 15;
    *reinterpret_cast<float4*>(dst + (size_t)r * ld_dst + c4 * 4) =
        *reinterpret_cast<const float4*>(src + (size_t)r * ld_src + c4 * 4);
  }
}
__device__ __forceinline__ void zero_rows_n(float* __restrict__ dst, int ld_dst, int nrows, int t0, int nthreads) {
  for (int t = t0; t < nrows * 16; t += nthreads) {
    int r = t >> 4, c4 = t & 15;
    *reinterpret_cast<float4*>(dst + (size_t)r * ld_dst + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// target-segment sums to global rows, 256 threads = 8 warps.  Warp w takes segments w, w+8, ...; its lanes are
// 16 column quads x 2 row halves (first / second half of the segment), combined in fixed order through one shuffle.
__device__ __forceinline__ void reduce_targets_global_256(const float* __restrict__ sX, const TileKeys* __restrict__ tk,
                                                          float* __restrict__ out, int gt) {
  // the warps of column half 1 (which do not prepare radial operands) take the first four segments -- a tile has
  // five to six -- : forward 2.27 -> 2.23 ms, adjoint 2.40 -> 2.31 ms at C2
  const int lane = gt & 31, cq = lane & 15, half = lane >> 4, w = ((gt >> 5) + 4) & 7;
  const int nseg = tk->n_seg_t;
  for (int s = w; s < nseg; s += 8) {
    const int p0 = tk->seg_t[s], p1 = tk->seg_t[s + 1];
    const int mid = p0 + ((p1 - p0 + 1) >> 1);
    const int a0 = half ? mid : p0, a1 = half ? p1 : mid;
    float4 run = make_float4(0.f, 0.f, 0.f, 0.f), run2 = make_float4(0.f, 0.f, 0.f, 0.f);
    int p = a0;
#pragma unroll 2
    for (; p + 1 < a1; p += 2) {   // two interleaved partial sums, fixed order
      const float4 v = *reinterpret_cast<const float4*>(sX + p * kTileLd + 4 * cq);
      const float4 u = *reinterpret_cast<const float4*>(sX + (p + 1) * kTileLd + 4 * cq);
      run.x += v.x;
      run.y += v.y;
      run.z += v.z;
      run.w += v.w;
      run2.x += u.x;
      run2.y += u.y;
      run2.z += u.z;
      run2.w += u.w;
    }
    if (p < a1) {
      const float4 v = *reinterpret_cast<const float4*>(sX + p * kTileLd + 4 * cq);
      run.x += v.x;
      run.y += v.y;
      run.z += v.z;
      run.w += v.w;
    }
    run.x += run2.x;
    run.y += run2.y;
    run.z += run2.z;
    run.w += run2.w;
    // first half + second half
    const float ox = __shfl_xor_sync(0xffffffffu, run.x, 16), oy = __shfl_xor_sync(0xffffffffu, run.y, 16);
    const float oz = __shfl_xor_sync(0xffffffffu, run.z, 16), ow = __shfl_xor_sync(0xffffffffu, run.w, 16);
    if (half == 0) {
      run.x += ox;
      run.y += oy;
      run.z += oz;
      run.w += ow;
      float4* dst = reinterpret_cast<float4*>(out + (size_t)tk->tgt[p0] * 64 + 4 * cq);
      if (s == 0 && tk->cont_prev) {
        float4 t = *dst;
        run.x += t.x;
        run.y += t.y;
        run.z += t.z;
        run.w += t.w;
      }
      *dst = run;
    }
  }
}
// source-segment sums into table rows, 256 threads: 16 column quads x 16 segment slots; rows of a segment are
// fetched four at a time (independent loads) and added in order
__device__ __forceinline__ void reduce_sources_table_256(const float* __restrict__ sX, const TileKeys* __restrict__ tk,
                                                         float* __restrict__ table, int ld, int gt) {
  const int cq = gt & 15;
  const int nseg = tk->n_seg_s;
  constexpr int kSlots = 16;
  for (int s = gt >> 4; s < nseg; s += kSlots) {
    const int p0 = tk->seg_s[s], p1 = tk->seg_s[s + 1];
    const int row0 = tk->perm[p0];
    float4* dst = reinterpret_cast<float4*>(table + (size_t)tk->src[row0] * ld + 4 * cq);
    float4 t = *dst;
    for (int p = p0; p < p1; p += 4) {
      const int r0 = tk->perm[p], r1 = tk->perm[min(p + 1, p1 - 1)], r2 = tk->perm[min(p + 2, p1 - 1)],
                r3 = tk->perm[min(p + 3, p1 - 1)];
      const float4 v0 = *reinterpret_cast<const float4*>(sX + r0 * kTileLd + 4 * cq);
      const float4 v1 = *reinterpret_cast<const float4*>(sX + r1 * kTileLd + 4 * cq);
      const float4 v2 = *reinterpret_cast<const float4*>(sX + r2 * kTileLd + 4 * cq);
      const float4 v3 = *reinterpret_cast<const float4*>(sX + r3 * kTileLd + 4 * cq);
      const bool b1 = p + 1 < p1, b2 = p + 2 < p1, b3 = p + 3 < p1;
      t.x += v0.x;
      t.y += v0.y;
      t.z += v0.z;
      t.w += v0.w;
      if (b1) {
        t.x += v1.x;
        t.y += v1.y;
        t.z += v1.z;
        t.w += v1.w;
      }
      if (b2) {
        t.x += v2.x;
        t.y += v2.y;
        t.z += v2.z;
        t.w += v2.w;
      }
      if (b3) {
        t.x += v3.x;
        t.y += v3.y;
        t.z += v3.z;
        t.w += v3.w;
      }
    }
    *dst = t;
  }
}

// ---- variant with a dedicated MMA-issuer warp (17 warps -> 96 registers per thread; used for the forward kernel,
// where the issue work would otherwise sit on the critical path of the slowest epilogue warp): lane-elected issue
// by warp 16, which polls (mbarrier.test_wait) the two groups in turn:
//     u_ok (12 arrivals), a_full (8 arrivals), d_full (tcgen05.commit)
constexpr int kWsiThreads = kGroups * kWsGroupThreads + 32;
struct WsiBars {
  uint64_t d_full[kGroups], a_full[kGroups], u_ok[kGroups], tile_b[kGroups];
  uint64_t img_free[kGroups];   // table mode 3: the silu'(w) image of the previous tile has left the shared image buffer
};
__device__ __forceinline__ void warp_arrive_bar(uint64_t* bar, int lane);
__device__ __forceinline__ void warp_arrive_i(uint64_t* bar, int lane) {
  fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
// issuer-side walk over the units of one group (must visit exactly the tiles the epilogue warps visit)
struct IssueState {
  int unit, t0, e_end, stage;
  uint32_t ph_a, ph_u, cur;
  bool done;
};
template <bool SPLIT>
__device__ __forceinline__ void issue_next_unit(IssueState& x, const EdgeTcArgs& g) {
  for (;;) {
    if (x.unit >= g.n_units) {
      x.done = true;
      return;
    }
    const UnitSpan us = unit_span(x.unit, g.G, g.N, g.R, SPLIT ? g.S : 1, g.U0);
    const int e_begin = g.row_ptr[us.atom_base + us.a0], e_end = g.row_ptr[us.atom_base + us.a1];
    if (e_begin < e_end) {
      x.t0 = e_begin;
      x.e_end = e_end;
      return;
    }
    x.unit += gridDim.x * kGroups;
  }
}

// ---------------------------------------------------------------------------------------------- forward (ws, dedicated issuer warp)
// STASH: the evaluation wants forces and the activation stash is in use -- the two epilogues also produce silu'(u) and
// silu'(w), pack them as 16-bit codes into a swizzled 16 KB image of the tile in shared memory, and one thread sends
// each image to global memory with one bulk store (TMA engine); the adjoint (edge_bwd_st_kernel) then neither
// recomputes u and w nor evaluates a single sigmoid.
template <int TAB, int TM, bool SPLIT, bool STASH>
__global__ void __launch_bounds__(kWsiThreads, 1) edge_fwd_wsi_kernel(EdgeTcArgs g, uint32_t grp_bytes) {
  constexpr int TERMS = TM == kTanh3 ? 3 : TM;                 // arithmetic of the products
  constexpr bool FAST = is_half(TM) || TM == kTanh3;           // tanh-based activations
  constexpr bool Q_SM = TAB >= 1, P_SM = TAB == 2;   // which node tables of the unit live in shared memory
  // TAB == 3 (stash only): Q table as in mode 1, and ONE 16 KB image buffer instead of two.  The silu'(u) image of a tile
  // has left it when the second epilogue starts (thread 0 waits for that before barrier (A) in every mode); what is new is
  // that the first epilogue of the NEXT tile may only write its codes once the silu'(w) image has left: thread 128 sends
  // that image after barrier (B), waits until the TMA engine has read it and arrives on img_free, on which every thread
  // waits just before its first code store (by then it has evaluated 16 activations).  Measured (edge forward, ms; results
  // bit-identical): N = 100, R = 4096 7.59 -> 5.57; N = 80 4.87 -> 3.63; N = 116, R = 1024 2.03 -> 1.48.  The same trade for
  // 50 atoms (P and Q tables + one buffer instead of Q table + two buffers) loses: 1.988 -> 2.064.
  constexpr bool ONEBUF = STASH && TAB == 3;
  static_assert(TAB != 3 || STASH, "table mode 3 exists for the stash kernels only");
  // How a tile is closed.  With the unit's tables in shared memory: a group barrier (B) after the second epilogue.  Without
  // tables (more than 64 atoms with the stash: both row gathers go through L1 / L2 and the warps drift apart): an mbarrier
  // arrival -- a warp goes on to the next tile's first epilogue at once, warp 4 (no radial work) sends the tile's silu'(w)
  // image as soon as all eight arrivals are in, and the tile is reduced under the next w product by whoever has waited.
  // Measured (edge forward, ms): N = 100, R = 1024 1.967 -> 1.925; N = 50, R = 4096 1.986 -> 2.000 (hence the barrier there)
  constexpr bool TB = !Q_SM;
  constexpr uint32_t kStg = STASH ? (ONEBUF ? kStageBytes / 2 : kStageBytes) : 0u;
  extern __shared__ uint8_t smraw[];
  __shared__ WsiBars bars;
  __shared__ uint32_t tmem_slot;
  uint32_t base = smem_u32(smraw);
  uint8_t* sm = smraw + ((1024u - (base & 1023u)) & 1023u);
  // operand images: 3xTF32 / TF32: Wr hi | lo | W2 hi | lo (48 KB); bf16 mode: Wr | W2 (8 KB each, no lo parts)
  constexpr bool BF = is_half(TERMS);
  uint8_t *sWrH = sm, *sWrL = sm + 8192, *sW2H = sm + (BF ? 8192 : 16384), *sW2L = sm + 32768;
  float* sb2 = reinterpret_cast<float*>(sm + kFwdWeights);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) {
      mbar_init(&bars.d_full[i], 1);
      mbar_init(&bars.a_full[i], 8);
      mbar_init(&bars.u_ok[i], 12);
      mbar_init(&bars.tile_b[i], 8);
      mbar_init(&bars.img_free[i], 1);
      if (ONEBUF) mbar_arrive(&bars.img_free[i]);   // the buffer starts out free (phase 0 complete)
    }
  }
  __shared__ uint64_t ibar;
  if (tid == 0) {
    mbar_init(&ibar, 1);
    bulk_copy_image(sm, g.img, BF ? 16384u : kFwdWeights, &ibar);   // B[n = c][k] = Wr[c][k] and W2[c][k], hi | lo
  }
  if (tid < 64) sb2[tid] = g.b2[tid];
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  mbar_wait(&ibar, 0);
  __syncwarp();
  pdl_wait();   // everything above touched only the weight images; the evaluation's data from here on
  // capacity-limited stash: an edge list that does not fit is not stashed (the adjoint of this evaluation recomputes)
  const bool st_ok = STASH && stash_fits(g);
  constexpr uint32_t VH = 128, VL = 192;
  constexpr uint32_t idesc64 = TERMS == kFp16 ? idesc_fp16(128, 64) : (BF ? idesc_bf16(128, 64) : idesc_tf32(128, 64));

  if (warp == 2 * 8) {
    // ===================================== MMA issuer =====================================
    // the whole warp walks the state machine (converged); one elected lane issues
    {
      const uint32_t aWrH = smem_u32(sWrH), aWrL = smem_u32(sWrL), aW2H = smem_u32(sW2H), aW2L = smem_u32(sW2L);
      IssueState st[kGroups];
      for (int i = 0; i < kGroups; ++i) {
        st[i] = IssueState{(int)blockIdx.x + i * (int)gridDim.x, 0, 0, 0, 0u, 0u, 0u, false};
        issue_next_unit<SPLIT>(st[i], g);
      }
      while (!(st[0].done && st[1].done)) {
#pragma unroll
        for (int i = 0; i < kGroups; ++i) {
          IssueState& x = st[i];
          if (x.done) continue;
          const uint32_t tm = tmem_slot + (uint32_t)i * 256u;
          if (x.stage == 0) {
            if (!mbar_test(&bars.u_ok[i], x.ph_u)) continue;
            x.ph_u ^= 1u;
            fence_after_sync();
            if (elect_one()) {
              issue_gemm_3xtf32<0, 3, TERMS>(tm + x.cur, tm + (x.cur ^ 64u), tm + (x.cur ^ 64u) + 32, aWrH, aWrL, 64, idesc64, 0u);
              mma_commit(&bars.d_full[i]);
            }
            __syncwarp();
            x.stage = 1;
          } else {
            if (!mbar_test(&bars.a_full[i], x.ph_a)) continue;
            x.ph_a ^= 1u;
            fence_after_sync();
            if (elect_one()) {
              issue_gemm_3xtf32<0, 8, TERMS>(tm + (x.cur ^ 64u), tm + VH, tm + VL, aW2H, aW2L, 64, idesc64, 0u);
              mma_commit(&bars.d_full[i]);
            }
            __syncwarp();
            x.stage = 0;
            x.cur ^= 64u;
            x.t0 += kTileE;
            if (x.t0 >= x.e_end) {
              x.unit += gridDim.x * kGroups;
              issue_next_unit<SPLIT>(x, g);
            }
          }
        }
      }
    }
  } else {
    // ===================================== epilogue warps =====================================
    const int grp = tid >> 8, gt = tid & 255, w8 = gt >> 5, wq = w8 & 3, ch = w8 >> 2, row = 32 * wq + lane;
    uint8_t* gsm = sm + kFwdWeights + 256 + (size_t)grp * grp_bytes;
    float* sZ = reinterpret_cast<float*>(gsm + kGrpX);
    uint8_t* sStU = gsm + kGrpStage;                  // STASH: tile images of the silu'(u) / silu'(w) codes
    uint8_t* sStW = ONEBUF ? sStU : gsm + kGrpStage + kStageBytes / 2;
    TileKeys* keys = reinterpret_cast<TileKeys*>(gsm + kGrpKeys + kStg);
    float* sTab = reinterpret_cast<float*>(gsm + kGrpTab + kStg);
    const uint32_t swz = (uint32_t)(row & 7);         // chunk swizzle of this thread's image row
    const uint32_t tm = tmem_slot + (uint32_t)grp * 256u;
    const uint32_t lane_addr = tm + ((uint32_t)(32 * wq) << 16);
    uint64_t *d_full = &bars.d_full[grp], *a_full = &bars.a_full[grp], *u_ok = &bars.u_ok[grp], *tile_b = &bars.tile_b[grp];
    uint64_t* img_free = &bars.img_free[grp];
    uint32_t ph_d = 0, ph_a = 0, ph_b = 0, ph_f = 0, cur = 0;
    int kidx = 0;   // TileKeys buffer of the current tile; (kidx + 1) % 3 = next tile, (kidx + 2) % 3 = previous tile

    for (int unit = blockIdx.x + grp * gridDim.x; unit < g.n_units; unit += gridDim.x * kGroups) {
      const UnitSpan us = unit_span(unit, g.G, g.N, g.R, SPLIT ? g.S : 1, g.U0);
      const int atom0 = us.atom_base, natoms = us.n_tab;
      const int e_begin = g.row_ptr[atom0 + us.a0], e_end = g.row_ptr[atom0 + us.a1];
      float* gA = g.a_out + (size_t)atom0 * 64;
      const float *tQ, *gP;   // gP: the P rows of the unit (shared-memory copy when the tables fit)
      constexpr int ldq = Q_SM ? kTabLd : 64, ldp = P_SM ? kTabLd : 64;
      if (Q_SM) {
        copy_rows_n(sTab, kTabLd, g.Q + (size_t)atom0 * 64, 64, natoms, gt, kWsGroupThreads);
        tQ = sTab;
      } else {
        tQ = g.Q + (size_t)atom0 * 64;
      }
      if (P_SM) {
        copy_rows_n(sTab + (size_t)g.unit_atoms * kTabLd, kTabLd, g.P + (size_t)atom0 * 64, 64, natoms, gt, kWsGroupThreads);
        gP = sTab + (size_t)g.unit_atoms * kTabLd;
      } else {
        gP = g.P + (size_t)atom0 * 64;
      }
      if (SPLIT) zero_rows_n(gA + (size_t)us.a0 * 64, 64, us.a1 - us.a0, gt, kWsGroupThreads);   // the targets this unit owns
      else zero_rows_n(gA, 64, natoms, gt, kWsGroupThreads);
      if (e_begin >= e_end) {
        ws_group_sync(grp);
        continue;
      }
      // ---- prologue: keys + radial basis of the first tile; edge data of the second tile in flight
      uint32_t pr = 0u, pr_n = 0u;
      float r = 0.3f, r_n = 0.3f;
      bool valid = e_begin + row < e_end;
      if (valid) {
        pr = ldg_stream(g.pair + e_begin + row);
        if (ch == 0) r = ldg_stream(g.r_e + e_begin + row);
      }
      if (e_begin + kTileE + row < e_end) {
        pr_n = ldg_stream(g.pair + e_begin + kTileE + row);
        if (ch == 0) r_n = ldg_stream(g.r_e + e_begin + kTileE + row);
      }
      TileKeys* tk = keys + kidx;
      if (ch == 0) tk->tgt[row] = valid ? (int)(pr >> 16) : -1;
      if (!P_SM) prefetch_l1(gP + (size_t)(valid ? (pr >> 16) : 0u) * 64 + 32 * ch);
      warp_arrive_i(u_ok, lane);                       // accumulators are free
      if (ch == 0) {
        float o[32], od[32];
        rbf_regs<false>(r, g.inv_cutoff, o, od);
        st_split24<TERMS>(lane_addr + (cur ^ 64u), lane_addr + (cur ^ 64u) + 32, o);
        tmem_st_wait();
        warp_arrive_i(u_ok, lane);
      }
      ws_group_sync(grp);                            // tables + keys of the first tile visible

      for (int t0 = e_begin; t0 < e_end; t0 += kTileE) {
        const bool has_next = t0 + kTileE < e_end;
        const int tgc = valid ? (int)(pr >> 16) : 0, sr = valid ? (int)(pr & 0xFFFFu) : 0;
        if (w8 == 4) {
          const int ns = build_segments<false>(tk->tgt, nullptr, tk->seg_t, min(kTileE, e_end - t0), lane);
          if (lane == 0) {
            tk->n_seg_t = ns;
            tk->cont_prev = (t0 > e_begin && keys[(kidx + 2) % 3].tgt[kTileE - 1] == tk->tgt[0]) ? 1 : 0;
          }
          if (TB && t0 > e_begin) {                    // this warp sends the previous tile's silu'(w) image as soon as it is complete
            mbar_wait(tile_b, ph_b);
            if (st_ok && lane == 0) bulk_store(g.st_w + (size_t)(t0 - kTileE) * 128, sStW, (uint32_t)kTileE * 128u);
          }
        }
        mbar_wait(d_full, ph_d);
        __syncwarp();
        ph_d ^= 1u;
        fence_after_sync();
        // ---- v = silu(u + P[tgt] + Q[src]) -> TMEM (this thread's 32 channels, two chunks of 16)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = 32 * ch + 16 * c;
          uint32_t ur[16];
          tmem_ld16(lane_addr + cur + col0, ur);
          const float4* pp = reinterpret_cast<const float4*>(gP + (size_t)tgc * ldp + col0);
          const float4* qq = reinterpret_cast<const float4*>(tQ + (size_t)sr * ldq + col0);
          float4 p[4], q[4];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            p[j4] = P_SM ? pp[j4] : ldg_keep4(pp + j4);
            q[j4] = Q_SM ? qq[j4] : ldg_keep4(qq + j4);
          }
          tmem_ld_wait();
          float v[16];
          if (STASH) {
            uint32_t code[8];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              float d[4];
              if (FAST) silu_fd4_sum3_tanh(&ur[4 * j4], p[j4], q[j4], &v[4 * j4], d);
              else silu_fd4_sum3(&ur[4 * j4], p[j4], q[j4], &v[4 * j4], d);
              code[2 * j4] = stash_enc2(d[0], d[1]);
              code[2 * j4 + 1] = stash_enc2(d[2], d[3]);
            }
            const uint32_t q0 = (uint32_t)(4 * ch + 2 * c);
            uint8_t* img = sStU + row * 128;
            if (ONEBUF && c == 0) {   // the silu'(w) image of the previous tile has left the buffer
              mbar_wait(img_free, ph_f);
              ph_f ^= 1u;
            }
            *reinterpret_cast<uint4*>(img + ((q0 ^ swz) << 4)) = make_uint4(code[0], code[1], code[2], code[3]);
            *reinterpret_cast<uint4*>(img + (((q0 + 1u) ^ swz) << 4)) = make_uint4(code[4], code[5], code[6], code[7]);
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 y = FAST ? silu4_sum3_tanh(&ur[4 * j4], p[j4], q[j4]) : silu4_sum3(&ur[4 * j4], p[j4], q[j4]);
              v[4 * j4 + 0] = y.x;
              v[4 * j4 + 1] = y.y;
              v[4 * j4 + 2] = y.z;
              v[4 * j4 + 3] = y.w;
            }
          }
          st_operand16<TERMS>(lane_addr, VH, VL, col0, v);
        }
        tmem_st_wait();
        if (STASH) fence_proxy_async_smem();           // the image rows are read by the bulk store below
        warp_arrive_i(a_full, lane);
        // ---- look ahead: keys / radial basis of the next tile while w runs
        TileKeys* tkn = keys + (kidx + 1) % 3;
        bool valid_n = false;
        uint32_t pr_nn = 0u;
        float r_nn = 0.3f;
        if (has_next) {
          valid_n = t0 + kTileE + row < e_end;
          if (ch == 0) tkn->tgt[row] = valid_n ? (int)(pr_n >> 16) : -1;
          if (!P_SM) prefetch_l1(gP + (size_t)(valid_n ? (pr_n >> 16) : 0u) * 64 + 32 * ch);
          if (t0 + 2 * kTileE + row < e_end) {
            pr_nn = ldg_stream(g.pair + t0 + 2 * kTileE + row);
            if (ch == 0) r_nn = ldg_stream(g.r_e + t0 + 2 * kTileE + row);
          }
          if (ch == 0) {
            float o[32], od[32];
            rbf_regs<false>(r_n, g.inv_cutoff, o, od);
            mbar_wait(a_full, ph_a);                 // every warp has finished reading u of this tile
            __syncwarp();
            if (st_ok && gt == 0) bulk_store(g.st_u + (size_t)t0 * 128, sStU, (uint32_t)min(kTileE, e_end - t0) * 128u);
            st_split24<TERMS>(lane_addr + cur, lane_addr + cur + 32, o);
            tmem_st_wait();
            warp_arrive_i(u_ok, lane);
          }
        } else if (STASH && w8 == 0) {
          mbar_wait(a_full, ph_a);                   // all eight warps have written their rows of the silu'(u) image
          __syncwarp();
          if (st_ok && gt == 0) bulk_store(g.st_u + (size_t)t0 * 128, sStU, (uint32_t)min(kTileE, e_end - t0) * 128u);
        }
        ph_a ^= 1u;
        // ---- a[tgt] += z of the PREVIOUS tile while the tensor pipe works on w of this one
        if (t0 > e_begin) {
          if (TB) {
            mbar_wait(tile_b, ph_b);                 // every warp has written its rows of the previous tile
            ph_b ^= 1u;
          }
          reduce_targets_global_256(sZ, keys + (kidx + 2) % 3, gA, gt);
        }
        mbar_wait(d_full, ph_d);
        __syncwarp();
        ph_d ^= 1u;
        fence_after_sync();
        // ---- w -> registers; the accumulator may then be overwritten by u of the next tile
        uint32_t wr[32];
        tmem_ld32(lane_addr + (cur ^ 64u) + 32 * ch, wr);
        tmem_ld_wait();
        if (has_next) warp_arrive_i(u_ok, lane);
        if (STASH && (gt == 0 || ((TB || ONEBUF) && gt == 128))) bulk_store_wait_read();   // the images sent so far have left shared memory
        ws_group_sync(grp);                          // (A) the reduction of the previous tile is complete
        {
          const float4* bb = reinterpret_cast<const float4*>(sb2 + 32 * ch);
          float* zrow = sZ + row * kTileLd + 32 * ch;
          if (STASH) {
            uint8_t* img = sStW + row * 128;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              float z[8], d[8];
              if (FAST) {
                silu_fd4_sum2_tanh(&wr[8 * j8], bb[2 * j8], z, d);
                silu_fd4_sum2_tanh(&wr[8 * j8 + 4], bb[2 * j8 + 1], z + 4, d + 4);
              } else {
                silu_fd4_sum2(&wr[8 * j8], bb[2 * j8], z, d);
                silu_fd4_sum2(&wr[8 * j8 + 4], bb[2 * j8 + 1], z + 4, d + 4);
              }
              *reinterpret_cast<float4*>(zrow + 8 * j8) = make_float4(z[0], z[1], z[2], z[3]);
              *reinterpret_cast<float4*>(zrow + 8 * j8 + 4) = make_float4(z[4], z[5], z[6], z[7]);
              *reinterpret_cast<uint4*>(img + ((((uint32_t)(4 * ch + j8)) ^ swz) << 4)) =
                  make_uint4(stash_enc2(d[0], d[1]), stash_enc2(d[2], d[3]), stash_enc2(d[4], d[5]), stash_enc2(d[6], d[7]));
            }
            fence_proxy_async_smem();
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b = bb[j4];
              const float4 z = FAST ? silu4_sum2_tanh(&wr[4 * j4], b) : silu4_sum2(&wr[4 * j4], b);
              *reinterpret_cast<float4*>(zrow + 4 * j4) = z;
            }
          }
        }
        if (TB) {
          // (B) tile complete: an arrival, not a barrier.  The silu'(u) image of this tile is known to have left shared
          // memory at (A) of this tile (wait_read of thread 0)
          warp_arrive_bar(tile_b, lane);
        } else {
          if (STASH && gt == 0) bulk_store_wait_read();   // the silu'(u) image of this tile has left shared memory
          ws_group_sync(grp);                          // (B) tile complete; it is reduced under the next tile's w GEMM
          if (ONEBUF) {
            if (gt == 128) {
              if (st_ok) bulk_store(g.st_w + (size_t)t0 * 128, sStW, (uint32_t)min(kTileE, e_end - t0) * 128u);
              bulk_store_wait_read();
              mbar_arrive(img_free);
            }
          } else if (st_ok && gt == 0) bulk_store(g.st_w + (size_t)t0 * 128, sStW, (uint32_t)min(kTileE, e_end - t0) * 128u);
        }
        pr = pr_n;
        r = r_n;
        valid = valid_n;
        pr_n = pr_nn;
        r_n = r_nn;
        tk = tkn;
        kidx = (kidx + 1) % 3;
        cur ^= 64u;
      }
      if (TB) {
        mbar_wait(tile_b, ph_b);
        ph_b ^= 1u;
        if (st_ok && gt == 128) {
          const int t_last = e_begin + ((e_end - e_begin - 1) / kTileE) * kTileE;
          bulk_store(g.st_w + (size_t)t_last * 128, sStW, (uint32_t)(e_end - t_last) * 128u);
        }
      }
      reduce_targets_global_256(sZ, keys + (kidx + 2) % 3, gA, gt);   // last tile of the unit
      ws_group_sync(grp);   // all rows of the unit written before its tables are replaced
    }
    if (STASH && (gt == 0 || ((TB || ONEBUF) && gt == 128))) bulk_store_wait_all();   // every image has reached global memory before the CTA retires
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

// ---------------------------------------------------------------------------------------------- adjoint (ws)
template <int TAB, int TERMS, bool SPLIT>
__global__ void __launch_bounds__(kWsThreads, 1) edge_bwd_ws_kernel(EdgeTcArgs g, uint32_t grp_bytes) {
  constexpr bool QB_SM = TAB >= 1, Q_SM = TAB == 2;   // Qbar accumulator / Q rows of the unit in shared memory
  extern __shared__ uint8_t smraw[];
  __shared__ WsBars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ float sPart[kGroups][kTileE];
  uint32_t base = smem_u32(smraw);
  uint8_t* sm = smraw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t *sWrH = sm, *sWrL = sm + 8192, *sW2H = sm + 16384, *sW2L = sm + 32768;
  uint8_t *sW2tH = sm + 49152, *sW2tL = sm + 65536;
  float* sb2 = reinterpret_cast<float*>(sm + kBwdWeights);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // capacity-limited stash: this launch is the alternative to edge_bwd_st_kernel and runs only when the edge list of the
  // evaluation did not fit the stash (uniform over the grid, before anything is allocated)
  pdl_trigger();
  if (g.stash_role == 2 && stash_fits(g)) {
    pdl_wait();
    return;
  }

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) {
      mbar_init(&bars.d_full[i], 1);
      mbar_init(&bars.a_full[i], 8);
      bars.cnt_a[i] = 0;
      bars.cnt_u[i] = 0;
    }
  }
  __shared__ uint64_t ibar;
  if (tid == 0) {
    mbar_init(&ibar, 1);
    bulk_copy_image(sm, g.img, kBwdWeights, &ibar);   // Wr, W2 as above and B[n = k'][k = c] = W2[c][k'] for vbar
  }
  if (tid < 64) sb2[tid] = g.b2[tid];
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  mbar_wait(&ibar, 0);
  __syncwarp();
  pdl_wait();
  constexpr uint32_t ACC = 0, DU = 64, XH = 128, XL = 192;
  constexpr uint32_t idesc64 = idesc_tf32(128, 64);

  {
    const int grp = tid >> 8, gt = tid & 255, w8 = gt >> 5, wq = w8 & 3, ch = w8 >> 2, row = 32 * wq + lane;
    uint8_t* gsm = sm + kBwdWeights + 256 + (size_t)grp * grp_bytes;
    float* sU = reinterpret_cast<float*>(gsm + kGrpX);
    TileKeys* keys = reinterpret_cast<TileKeys*>(gsm + kGrpKeys);
    float* sTab = reinterpret_cast<float*>(gsm + kGrpTab);
    const uint32_t tm = tmem_slot + (uint32_t)grp * 256u;
    const uint32_t lane_addr = tm + ((uint32_t)(32 * wq) << 16);
    uint64_t* d_full = &bars.d_full[grp];
    int *cnt_a = &bars.cnt_a[grp], *cnt_u = &bars.cnt_u[grp];
    const uint32_t aWrH = smem_u32(sWrH), aWrL = smem_u32(sWrL), aW2H = smem_u32(sW2H), aW2L = smem_u32(sW2L);
    const uint32_t aW2tH = smem_u32(sW2tH), aW2tL = smem_u32(sW2tL);
    auto issue_u_du = [&]() {      // u = rbf . Wr^T ; du = drbf . Wr^T
      if (elect_one()) {
        issue_gemm_3xtf32<0, 3, TERMS>(tm + ACC, tm + XH, tm + XL, aWrH, aWrL, 64, idesc64, 0u);
        issue_gemm_3xtf32<0, 3, TERMS>(tm + DU, tm + XH + 32, tm + XL + 32, aWrH, aWrL, 64, idesc64, 0u);
        mma_commit(d_full);
      }
      __syncwarp();
    };
    auto issue_x = [&](uint32_t bH, uint32_t bL) {   // ACC = X . B^T  (w with W2, vbar with W2^T)
      if (elect_one()) {
        issue_gemm_3xtf32<0, 8, TERMS>(tm + ACC, tm + XH, tm + XL, bH, bL, 64, idesc64, 0u);
        mma_commit(d_full);
      }
      __syncwarp();
    };
    uint32_t ph_d = 0;
    int kidx = 0;   // TileKeys buffer of the current tile; (kidx + 1) % 3 = next tile, (kidx + 2) % 3 = previous tile

    // the first unit of a group is fixed, the following ones are handed out by a global counter: groups that got
    // lighter units take more of them (units are independent, so results do not depend on who computes them;
    // adjoint 3.59 -> 3.57 ms at C2)
    __shared__ int s_next[kGroups];
    for (int unit = blockIdx.x + grp * gridDim.x; unit < g.n_units; unit = s_next[grp]) {
      const UnitSpan us = unit_span(unit, g.G, g.N, g.R, SPLIT ? g.S : 1, g.U0);
      const int atom0 = us.atom_base, natoms = us.n_tab;
      const int e_begin = g.row_ptr[atom0 + us.a0], e_end = g.row_ptr[atom0 + us.a1];
      const float* gP = g.P + (size_t)atom0 * 64;
      const float* gAb = g.abar + (size_t)atom0 * 64;
      float* gPb = g.Pbar + (size_t)atom0 * 64;
      // this part's (partial) Qbar rows; recomputed where it is needed instead of kept in registers across the tiles
      auto qbar_rows = [&]() { return g.Qbar + (SPLIT ? (size_t)us.part * g.qbar_stride : (size_t)0) + (size_t)atom0 * 64; };
      const float* tQ;
      float* tQb;
      constexpr int ldq = Q_SM ? kTabLd : 64, ld = QB_SM ? kTabLd : 64;   // ld: row pitch of the Qbar accumulator
      if (Q_SM) {                      // two tables: Q, then Qbar
        tQb = sTab + g.unit_atoms * kTabLd;
        copy_rows_n(sTab, kTabLd, g.Q + (size_t)atom0 * 64, 64, natoms, gt, kWsGroupThreads);
        tQ = sTab;
      } else {
        tQ = g.Q + (size_t)atom0 * 64;
        tQb = QB_SM ? sTab : qbar_rows();
      }
      if (SPLIT) zero_rows_n(gPb + (size_t)us.a0 * 64, 64, us.a1 - us.a0, gt, kWsGroupThreads);   // the targets this unit owns
      else zero_rows_n(gPb, 64, natoms, gt, kWsGroupThreads);
      zero_rows_n(tQb, ld, natoms, gt, kWsGroupThreads);
      if (e_begin >= e_end) {
        ws_group_sync(grp);
        if (gt == 0) s_next[grp] = (int)atomicAdd(g.unit_counter, 1u) + (int)(gridDim.x * kGroups);
        ws_group_sync(grp);
        if (QB_SM) {
          copy_rows_n(qbar_rows(), 64, tQb, ld, natoms, gt, kWsGroupThreads);
          ws_group_sync(grp);
        }
        continue;
      }
      uint32_t pr = 0u, pr_n = 0u;
      float r = 0.3f, r_n = 0.3f;
      uint8_t pm = (uint8_t)row, pm_n = (uint8_t)row;
      bool valid = e_begin + row < e_end;
      if (valid) {
        pr = g.pair[e_begin + row];
        if (ch == 0) {
          r = g.r_e[e_begin + row];
          pm = g.perm[e_begin + row];
        }
      }
      if (e_begin + kTileE + row < e_end) {
        pr_n = g.pair[e_begin + kTileE + row];
        if (ch == 0) {
          r_n = g.r_e[e_begin + kTileE + row];
          pm_n = g.perm[e_begin + kTileE + row];
        }
      }
      TileKeys* tk = keys + kidx;
      if (ch == 0) {
        tk->tgt[row] = valid ? (int)(pr >> 16) : -1;
        tk->src[row] = valid ? (int)(pr & 0xFFFFu) : -1;
        tk->perm[row] = pm;
      }
      {
        const size_t t_ = (size_t)(valid ? (pr >> 16) : 0u) * 64 + 32 * ch;
        prefetch_l1(gP + t_);
        prefetch_l1(gAb + t_);
      }
      if (ch == 0) {
        float o[32], od[32];
        rbf_regs<true>(r, g.inv_cutoff, o, od);
        st_split24<TERMS>(lane_addr + XH, lane_addr + XL, o);
        st_split24<TERMS>(lane_addr + XH + 32, lane_addr + XL + 32, od);
        tmem_st_wait();
      }
      if (warp_arrive_last(cnt_u, 8, lane)) issue_u_du();
      ws_group_sync(grp);
      // every thread of the group has read the previous hand-out before this barrier; the new one is read after the
      // barriers that close the unit
      if (gt == 0) s_next[grp] = (int)atomicAdd(g.unit_counter, 1u) + (int)(gridDim.x * kGroups);

      for (int t0 = e_begin; t0 < e_end; t0 += kTileE) {
        const bool has_next = t0 + kTileE < e_end;
        const int tg = valid ? (int)(pr >> 16) : -1, tgc = valid ? tg : 0, sr = valid ? (int)(pr & 0xFFFFu) : 0;
        if (w8 == 4) {
          const int ns = build_segments<false>(tk->tgt, nullptr, tk->seg_t, min(kTileE, e_end - t0), lane);
          if (lane == 0) {
            tk->n_seg_t = ns;
            tk->cont_prev = (t0 > e_begin && keys[(kidx + 2) % 3].tgt[kTileE - 1] == tk->tgt[0]) ? 1 : 0;
          }
        } else if (w8 == 5) {
          const int ns = build_segments<true>(tk->src, tk->perm, tk->seg_s, min(kTileE, e_end - t0), lane);
          if (lane == 0) tk->n_seg_s = ns;
        }
        // the reductions of the PREVIOUS tile run while the tensor pipe works on u, du (targets) and w (sources)
        const TileKeys* tkp = keys + (kidx + 2) % 3;
        if (t0 > e_begin) reduce_targets_global_256(sU, tkp, gPb, gt);          // Pbar[tgt] += ubar
        mbar_wait(d_full, ph_d);
        __syncwarp();
        ph_d ^= 1u;
        fence_after_sync();
        // ---- v = silu(u + P[tgt] + Q[src]) -> TMEM ; keep silu'(u)
        float dsu[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = 32 * ch + 16 * c;
          uint32_t ur[16];
          tmem_ld16(lane_addr + ACC + col0, ur);
          const float4* pp = reinterpret_cast<const float4*>(gP + (size_t)tgc * 64 + col0);
          const float4* qq = reinterpret_cast<const float4*>(tQ + (size_t)sr * ldq + col0);
          float4 p[4], q[4];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            p[j4] = __ldg(pp + j4);
            q[j4] = Q_SM ? qq[j4] : __ldg(qq + j4);
          }
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            silu_fd4_sum3(&ur[4 * j4], p[j4], q[j4], &v[4 * j4], &dsu[16 * c + 4 * j4]);
          }
          st_operand16<TERMS>(lane_addr, XH, XL, col0, v);
        }
        tmem_st_wait();
        if (warp_arrive_last(cnt_a, 8, lane)) issue_x(aW2H, aW2L);              // w = v . W2^T
        if (t0 > e_begin) reduce_sources_table_256(sU, tkp, tQb, ld, gt);       // Qbar[src] += ubar
        mbar_wait(d_full, ph_d);
        __syncwarp();
        ph_d ^= 1u;
        fence_after_sync();
        // ---- wbar = abar[tgt] * silu'(w + b2) -> TMEM
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = 32 * ch + 16 * c;
          uint32_t wr[16];
          tmem_ld16(lane_addr + ACC + col0, wr);
          const float4* ab = reinterpret_cast<const float4*>(gAb + (size_t)tgc * 64 + col0);
          const float4* bb = reinterpret_cast<const float4*>(sb2 + col0);
          float4 a[4], b[4];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            a[j4] = __ldg(ab + j4);
            b[j4] = bb[j4];
          }
          tmem_ld_wait();
          float wb[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 sd = silu_d4_sum2_scaled(&wr[4 * j4], b[j4], a[j4]);
            wb[4 * j4 + 0] = sd.x;
            wb[4 * j4 + 1] = sd.y;
            wb[4 * j4 + 2] = sd.z;
            wb[4 * j4 + 3] = sd.w;
          }
          st_operand16<TERMS>(lane_addr, XH, XL, col0, wb);
        }
        tmem_st_wait();
        if (warp_arrive_last(cnt_a, 8, lane)) issue_x(aW2tH, aW2tL);            // vbar = wbar . W2
        // ---- edge data of the following tiles while vbar runs
        TileKeys* tkn = keys + (kidx + 1) % 3;
        bool valid_n = false;
        uint32_t pr_nn = 0u;
        float r_nn = 0.3f;
        uint8_t pm_nn = (uint8_t)row;
        if (has_next) {
          valid_n = t0 + kTileE + row < e_end;
          if (ch == 0) {
            tkn->tgt[row] = valid_n ? (int)(pr_n >> 16) : -1;
            tkn->src[row] = valid_n ? (int)(pr_n & 0xFFFFu) : -1;
            tkn->perm[row] = valid_n ? pm_n : (uint8_t)row;
          }
          const size_t t_ = (size_t)(valid_n ? (pr_n >> 16) : 0u) * 64 + 32 * ch;
          prefetch_l1(gP + t_);
          prefetch_l1(gAb + t_);
          if (t0 + 2 * kTileE + row < e_end) {
            pr_nn = g.pair[t0 + 2 * kTileE + row];
            if (ch == 0) {
              r_nn = g.r_e[t0 + 2 * kTileE + row];
              pm_nn = g.perm[t0 + 2 * kTileE + row];
            }
          }
        }
        // radial operands of the next tile are computed now (registers), while vbar runs, and stored after epi3
        float o_n[32], od_n[32];
        if (has_next && ch == 0) {
          rbf_regs<true>(r_n, g.inv_cutoff, o_n, od_n);
#pragma unroll
          for (int k = 0; k < kRbf; ++k) asm volatile("" : "+f"(o_n[k]), "+f"(od_n[k]));   // keep the work above the wait
        }
        mbar_wait(d_full, ph_d);
        __syncwarp();
        ph_d ^= 1u;
        fence_after_sync();
        // vbar is complete, so the X operand columns are free: the radial operands of the next tile leave the
        // registers before the last epilogue needs them
        if (has_next && ch == 0) {
          st_split24<TERMS>(lane_addr + XH, lane_addr + XL, o_n);
          st_split24<TERMS>(lane_addr + XH + 32, lane_addr + XL + 32, od_n);
        }
        // slot of this edge in the dense pair-adjoint array; its previous value is fetched early (layers 2, 1 add)
        size_t rslot = 0;
        float rold = 0.0f;
        if (ch == 0 && valid) {
          rslot = ((size_t)atom0 + tg) * g.N + (sr - (tg / g.N) * g.N);
          if (g.rbar_accumulate) rold = g.rbar[rslot];
        }
        // No barrier is needed before the shared tile is overwritten: this warp is past the vbar wait, vbar was issued
        // by the last of the eight warps to finish the wbar epilogue, and every warp runs both reductions of the
        // previous tile before that epilogue -- so all reads of the tile are complete.
        // ---- ubar = vbar * silu'(u) -> smem tile ; rbar = ubar . du (this thread's 32 channels)
        float racc = 0.0f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = 32 * ch + 16 * c;
          uint32_t vr[16], dr[16];
          tmem_ld16(lane_addr + ACC + col0, vr);
          tmem_ld16(lane_addr + DU + col0, dr);
          tmem_ld_wait();
          float* urow = sU + row * kTileLd + col0;
          f2_t rab = pk(0.0f, 0.0f);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float* ds = &dsu[16 * c + 4 * j4];
            const f2_t u01 = mul2(pk(__uint_as_float(vr[4 * j4 + 0]), __uint_as_float(vr[4 * j4 + 1])), pk(ds[0], ds[1]));
            const f2_t u23 = mul2(pk(__uint_as_float(vr[4 * j4 + 2]), __uint_as_float(vr[4 * j4 + 3])), pk(ds[2], ds[3]));
            rab = fma2(u01, pk(__uint_as_float(dr[4 * j4 + 0]), __uint_as_float(dr[4 * j4 + 1])), rab);
            rab = fma2(u23, pk(__uint_as_float(dr[4 * j4 + 2]), __uint_as_float(dr[4 * j4 + 3])), rab);
            float4 ub;
            upk(u01, ub.x, ub.y);
            upk(u23, ub.z, ub.w);
            *reinterpret_cast<float4*>(urow + 4 * j4) = ub;
          }
          float ra, rb;
          upk(rab, ra, rb);
          racc += ra + rb;
        }
        if (has_next) {
          if (ch == 0) tmem_st_wait();
          if (warp_arrive_last(cnt_u, 8, lane)) issue_u_du();   // ACC / DU drained, radial operands in place
        }
        if (ch == 1) sPart[grp][row] = racc;
        ws_group_sync(grp);                          // (B) tile complete
        if (ch == 0 && valid) g.rbar[rslot] = rold + (racc + sPart[grp][row]);   // channels 0..31 + 32..63
        pr = pr_n;
        r = r_n;
        pm = pm_n;
        valid = valid_n;
        pr_n = pr_nn;
        r_n = r_nn;
        pm_n = pm_nn;
        tk = tkn;
        kidx = (kidx + 1) % 3;
      }
      reduce_targets_global_256(sU, keys + (kidx + 2) % 3, gPb, gt);          // last tile of the unit
      reduce_sources_table_256(sU, keys + (kidx + 2) % 3, tQb, ld, gt);
      ws_group_sync(grp);
      if (QB_SM) {
        copy_rows_n(qbar_rows(), 64, tQb, ld, natoms, gt, kWsGroupThreads);
        ws_group_sync(grp);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

// ---------------------------------------------------------------------------------------------- adjoint from the stash
// The adjoint when the forward kernels have left silu'(u), silu'(w) behind (EdgeTcArgs::st_u / st_w): per tile
//     wbar = abar[i] * silu'(w);  vbar = wbar W2 (tcgen05);  ubar = vbar * silu'(u)
//     Pbar[i] += ubar;  Qbar[j] += ubar;  rbar_e = ubar . (drbf Wr^T)      (du = drbf Wr^T: second tcgen05 product)
// i.e. two of the four products of edge_bwd_ws_kernel, no P / Q gathers and no sigmoid.  The 16 KB code images of a
// tile arrive by bulk copies (TMA engine, mbarrier completion) that are issued one tile ahead by whichever warp is
// the last to have consumed the previous image; all hand-offs inside a group are arrival counters / mbarriers:
//     d_full  tcgen05.commit            cnt_a / cnt_u  operand written (last arriving warp issues the product)
//     ldw/ldu image landed              cnt_w / cnt_s  image consumed  (last arriving warp requests the next one)
//     tileB   ubar tile of this tile complete in shared memory (8 warp arrivals; waited for before it is reduced)
//     redR    reductions of the previous tile finished         (8 warp arrivals; waited for before the tile buffer is rewritten)
// so a warp only ever waits for data it is about to touch.  TMEM per group: vbar 64 | du 64 | X hi 64 | X lo 64 columns;
// X holds drbf (24 columns) for the du product and wbar for the vbar product in turn.
constexpr uint32_t kStWeights = 49152;   // WrH 8K | WrL 8K | W2tH 16K | W2tL 16K
struct StBars {
  uint64_t d_full[kGroups], ldw[kGroups], ldu[kGroups], tileB[kGroups], redR[kGroups];
  int cnt_a[kGroups], cnt_u[kGroups], cnt_w[kGroups], cnt_s[kGroups];
};
// one arrival per warp on a counter in shared memory; true (warp-uniform) in the warp that completes the count
__device__ __forceinline__ bool warp_arrive_last_plain(int* cnt, int expected, int lane) {
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    int old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
    if (old == expected - 1) {
      asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(cnt)), "r"(0) : "memory");
      last = 1;
    }
  }
  return __shfl_sync(0xffffffffu, last, 0) != 0;
}
__device__ __forceinline__ void warp_arrive_bar(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// GATED: the instantiation launched for a capacity-limited stash (the default one stays free of the test)
template <int TAB, int TERMS, bool SPLIT, bool GATED>
__global__ void __launch_bounds__(kWsThreads, 1) edge_bwd_st_kernel(EdgeTcArgs g, uint32_t grp_bytes) {
  constexpr bool QB_SM = TAB >= 1;   // Qbar accumulator of the unit in shared memory
  extern __shared__ uint8_t smraw[];
  __shared__ StBars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ float sPart[kGroups][kTileE];
  __shared__ int s_next[kGroups];
  __shared__ uint64_t ibar;
  uint32_t base = smem_u32(smraw);
  uint8_t* sm = smraw + ((1024u - (base & 1023u)) & 1023u);
  // operand images: Wr hi | lo | W2^T hi | lo (48 KB), or in the bf16 mode Wr | W2^T (8 KB each)
  constexpr bool BF = is_half(TERMS);
  uint8_t *sWrH = sm, *sWrL = sm + 8192, *sW2tH = sm + (BF ? 8192 : 16384), *sW2tL = sm + 32768;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (GATED && !stash_fits(g)) {   // nothing was stashed: edge_bwd_ws_kernel (stash_role 2) takes over
    pdl_wait();
    return;
  }

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) {
      mbar_init(&bars.d_full[i], 1);
      mbar_init(&bars.ldw[i], 1);
      mbar_init(&bars.ldu[i], 1);
      mbar_init(&bars.tileB[i], 8);
      mbar_init(&bars.redR[i], 8);
      bars.cnt_a[i] = 0;
      bars.cnt_u[i] = 0;
      bars.cnt_w[i] = 0;
      bars.cnt_s[i] = 0;
    }
    mbar_init(&ibar, 1);
    // operand image (build_edge_images): Wr hi | lo at 0, W2^T hi | lo at 48 KB
    // (build_edge_images) 3xTF32 image: Wr at 0 (16 KB), W2^T at 48 KB (32 KB); bf16 image: Wr at 0, W2^T at 16 KB (8 KB each)
    const uint32_t n_wr = BF ? 8192u : 16384u, n_w2t = BF ? 8192u : 32768u, off_w2t = BF ? 16384u : 49152u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&ibar)), "r"(n_wr + n_w2t) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)),
                 "l"(g.img), "r"(n_wr), "r"(smem_u32(&ibar))
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sm + n_wr)),
                 "l"(g.img + off_w2t), "r"(n_w2t), "r"(smem_u32(&ibar))
                 : "memory");
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  mbar_wait(&ibar, 0);
  __syncwarp();
  pdl_wait();
  constexpr uint32_t ACC = 0, DU = 64, XH = 128, XL = 192;
  constexpr uint32_t idesc64 = TERMS == kFp16 ? idesc_fp16(128, 64) : (BF ? idesc_bf16(128, 64) : idesc_tf32(128, 64));

  const int grp = tid >> 8, gt = tid & 255, w8 = gt >> 5, wq = w8 & 3, ch = w8 >> 2, row = 32 * wq + lane;
  uint8_t* gsm = sm + kStWeights + 256 + (size_t)grp * grp_bytes;
  float* sU = reinterpret_cast<float*>(gsm + kGrpX);
  uint8_t* sLW = gsm + kGrpStage;                       // image of the silu'(w) codes of the current tile
  uint8_t* sLU = gsm + kGrpStage + kStageBytes / 2;     // ... silu'(u)
  TileKeys* keys = reinterpret_cast<TileKeys*>(gsm + kGrpKeys + kStageBytes);
  float* sTab = reinterpret_cast<float*>(gsm + kGrpTab + kStageBytes);
  const uint32_t swz = (uint32_t)(row & 7);
  const uint32_t tm = tmem_slot + (uint32_t)grp * 256u;
  const uint32_t lane_addr = tm + ((uint32_t)(32 * wq) << 16);
  uint64_t *d_full = &bars.d_full[grp], *ldw = &bars.ldw[grp], *ldu = &bars.ldu[grp], *tileB = &bars.tileB[grp],
           *redR = &bars.redR[grp];
  int *cnt_a = &bars.cnt_a[grp], *cnt_u = &bars.cnt_u[grp], *cnt_w = &bars.cnt_w[grp], *cnt_s = &bars.cnt_s[grp];
  const uint32_t aWrH = smem_u32(sWrH), aWrL = smem_u32(sWrL), aW2tH = smem_u32(sW2tH), aW2tL = smem_u32(sW2tL);
  auto issue_du = [&]() {      // du = drbf . Wr^T
    if (elect_one()) {
      issue_gemm_3xtf32<0, 3, TERMS>(tm + DU, tm + XH, tm + XL, aWrH, aWrL, 64, idesc64, 0u);
      mma_commit(d_full);
    }
    __syncwarp();
  };
  auto issue_vbar = [&]() {    // vbar = wbar . W2
    if (elect_one()) {
      issue_gemm_3xtf32<0, 8, TERMS>(tm + ACC, tm + XH, tm + XL, aW2tH, aW2tL, 64, idesc64, 0u);
      mma_commit(d_full);
    }
    __syncwarp();
  };
  uint32_t ph_d = 0, ph_w = 0, ph_u = 0, ph_b = 0, ph_r = 0;
  int kidx = 0;   // TileKeys buffer of the current tile; (kidx + 1) % 3 = next tile, (kidx + 2) % 3 = previous tile

  // the first unit of a group is fixed, the following ones come from a global counter (see edge_bwd_ws_kernel)
  for (int unit = blockIdx.x + grp * gridDim.x; unit < g.n_units; unit = s_next[grp]) {
    const UnitSpan us = unit_span(unit, g.G, g.N, g.R, SPLIT ? g.S : 1, g.U0);
    const int atom0 = us.atom_base, natoms = us.n_tab;
    const int e_begin = g.row_ptr[atom0 + us.a0], e_end = g.row_ptr[atom0 + us.a1];
    const float* gAb = g.abar + (size_t)atom0 * 64;
    float* gPb = g.Pbar + (size_t)atom0 * 64;
    auto qbar_rows = [&]() { return g.Qbar + (SPLIT ? (size_t)us.part * g.qbar_stride : (size_t)0) + (size_t)atom0 * 64; };
    constexpr int ld = QB_SM ? kTabLd : 64;   // row pitch of the Qbar accumulator
    float* tQb = QB_SM ? sTab : qbar_rows();
    if (SPLIT) zero_rows_n(gPb + (size_t)us.a0 * 64, 64, us.a1 - us.a0, gt, kWsGroupThreads);   // the targets this unit owns
    else zero_rows_n(gPb, 64, natoms, gt, kWsGroupThreads);
    zero_rows_n(tQb, ld, natoms, gt, kWsGroupThreads);
    if (e_begin >= e_end) {
      ws_group_sync(grp);
      if (gt == 0) s_next[grp] = (int)atomicAdd(g.unit_counter, 1u) + (int)(gridDim.x * kGroups);
      ws_group_sync(grp);
      if (QB_SM) {
        copy_rows_n(qbar_rows(), 64, tQb, ld, natoms, gt, kWsGroupThreads);
        ws_group_sync(grp);
      }
      continue;
    }
    // ---- prologue: images and keys of the first tile, edge data of the second tile, first du product
    if (gt == 0) {
      const uint32_t nb = (uint32_t)min(kTileE, e_end - e_begin) * 128u;
      bulk_load(sLW, g.st_w + (size_t)e_begin * 128, nb, ldw);
      bulk_load(sLU, g.st_u + (size_t)e_begin * 128, nb, ldu);
    }
    uint32_t pr = 0u, pr_n = 0u;
    float r = 0.3f, r_n = 0.3f;
    uint8_t pm = (uint8_t)row, pm_n = (uint8_t)row;
    bool valid = e_begin + row < e_end;
    if (valid) {
      pr = ldg_stream(g.pair + e_begin + row);
      if (ch == 0) {
        r = ldg_stream(g.r_e + e_begin + row);
        pm = ldg_stream(g.perm + e_begin + row);
      }
    }
    if (e_begin + kTileE + row < e_end) {
      pr_n = ldg_stream(g.pair + e_begin + kTileE + row);
      if (ch == 0) {
        r_n = ldg_stream(g.r_e + e_begin + kTileE + row);
        pm_n = ldg_stream(g.perm + e_begin + kTileE + row);
      }
    }
    TileKeys* tk = keys + kidx;
    if (ch == 0) {
      tk->tgt[row] = valid ? (int)(pr >> 16) : -1;
      tk->src[row] = valid ? (int)(pr & 0xFFFFu) : -1;
      tk->perm[row] = pm;
    }
    prefetch_l1(gAb + (size_t)(valid ? (pr >> 16) : 0u) * 64 + 32 * ch);
    if (ch == 0) {
      float o[32], od[32];
      rbf_regs<true>(r, g.inv_cutoff, o, od);
      st_split24<TERMS>(lane_addr + XH, lane_addr + XL, od);
      tmem_st_wait();
    }
    if (warp_arrive_last(cnt_u, 8, lane)) issue_du();
    ws_group_sync(grp);
    // every thread of the group has read the previous hand-out before this barrier; the new one is read after the
    // barriers that close the unit
    if (gt == 0) s_next[grp] = (int)atomicAdd(g.unit_counter, 1u) + (int)(gridDim.x * kGroups);

    // pair-adjoint slot / previous value / partial sum of this thread's row in the tile that is reduced next
    size_t rslot_p = 0;
    float rold_p = 0.0f, racc_p = 0.0f;
    bool valid_p = false;

    for (int t0 = e_begin; t0 < e_end; t0 += kTileE) {
      const bool has_next = t0 + kTileE < e_end;
      const int tg = valid ? (int)(pr >> 16) : -1, tgc = valid ? tg : 0, sr = valid ? (int)(pr & 0xFFFFu) : 0;
      // ---- the previous tile's pair adjoints and target sums (Pbar[tgt] += ubar) run while the du product of this tile is
      // in flight; its source sums follow under the vbar product below (adjoint 2.250 -> 2.236 ms at C2; moving both
      // reductions here: 2.30 ms, only the source sums: 2.247 ms)
      if (t0 > e_begin) {
        mbar_wait(tileB, ph_b);
        ph_b ^= 1u;
        if (ch == 0 && valid_p) st_rbar(g.rbar + rslot_p, rold_p + (racc_p + sPart[grp][row]));   // channels 0..31 + 32..63
        reduce_targets_global_256(sU, keys + (kidx + 2) % 3, gPb, gt);
      }
      // ---- du of this tile is complete: the X columns are free
      mbar_wait(d_full, ph_d);
      __syncwarp();
      ph_d ^= 1u;
      fence_after_sync();
      // ---- wbar = abar[tgt] * silu'(w) -> TMEM
      mbar_wait(ldw, ph_w);
      ph_w ^= 1u;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = 32 * ch + 16 * c;
        const uint32_t q0 = (uint32_t)(4 * ch + 2 * c);
        const uint4 c0 = *reinterpret_cast<const uint4*>(sLW + row * 128 + ((q0 ^ swz) << 4));
        const uint4 c1 = *reinterpret_cast<const uint4*>(sLW + row * 128 + (((q0 + 1u) ^ swz) << 4));
        const float4* ab = reinterpret_cast<const float4*>(gAb + (size_t)tgc * 64 + col0);
        const float4 a0 = ldg_keep4(ab), a1 = ldg_keep4(ab + 1), a2 = ldg_keep4(ab + 2), a3 = ldg_keep4(ab + 3);
        float wb[16];
        upk(mul2(stash_dec2(c0.x), pk(a0.x, a0.y)), wb[0], wb[1]);
        upk(mul2(stash_dec2(c0.y), pk(a0.z, a0.w)), wb[2], wb[3]);
        upk(mul2(stash_dec2(c0.z), pk(a1.x, a1.y)), wb[4], wb[5]);
        upk(mul2(stash_dec2(c0.w), pk(a1.z, a1.w)), wb[6], wb[7]);
        upk(mul2(stash_dec2(c1.x), pk(a2.x, a2.y)), wb[8], wb[9]);
        upk(mul2(stash_dec2(c1.y), pk(a2.z, a2.w)), wb[10], wb[11]);
        upk(mul2(stash_dec2(c1.z), pk(a3.x, a3.y)), wb[12], wb[13]);
        upk(mul2(stash_dec2(c1.w), pk(a3.z, a3.w)), wb[14], wb[15]);
        st_operand16<TERMS>(lane_addr, XH, XL, col0, wb);
      }
      tmem_st_wait();
      if (warp_arrive_last(cnt_a, 8, lane)) issue_vbar();
      if (has_next && warp_arrive_last_plain(cnt_w, 8, lane) && lane == 0)   // the image is consumed: fetch the next one
        bulk_load(sLW, g.st_w + (size_t)(t0 + kTileE) * 128, (uint32_t)min(kTileE, e_end - t0 - kTileE) * 128u, ldw);
      // ---- the previous tile: Qbar[src] += ubar (under the vbar product)
      if (t0 > e_begin) reduce_sources_table_256(sU, keys + (kidx + 2) % 3, tQb, ld, gt);
      warp_arrive_bar(redR, lane);
      // ---- segment lists of this tile (its keys are visible: written before the previous tile was closed)
      if (w8 == 4) {
        const int ns = build_segments<false>(tk->tgt, nullptr, tk->seg_t, min(kTileE, e_end - t0), lane);
        if (lane == 0) {
          tk->n_seg_t = ns;
          tk->cont_prev = (t0 > e_begin && keys[(kidx + 2) % 3].tgt[kTileE - 1] == tk->tgt[0]) ? 1 : 0;
        }
      } else if (w8 == 5) {
        const int ns = build_segments<true>(tk->src, tk->perm, tk->seg_s, min(kTileE, e_end - t0), lane);
        if (lane == 0) tk->n_seg_s = ns;
      }
      // ---- edge data of the following tiles
      TileKeys* tkn = keys + (kidx + 1) % 3;
      bool valid_n = false;
      uint32_t pr_nn = 0u;
      float r_nn = 0.3f;
      uint8_t pm_nn = (uint8_t)row;
      if (has_next) {
        valid_n = t0 + kTileE + row < e_end;
        if (ch == 0) {
          tkn->tgt[row] = valid_n ? (int)(pr_n >> 16) : -1;
          tkn->src[row] = valid_n ? (int)(pr_n & 0xFFFFu) : -1;
          tkn->perm[row] = valid_n ? pm_n : (uint8_t)row;
        }
        prefetch_l1(gAb + (size_t)(valid_n ? (pr_n >> 16) : 0u) * 64 + 32 * ch);
        if (t0 + 2 * kTileE + row < e_end) {
          pr_nn = ldg_stream(g.pair + t0 + 2 * kTileE + row);
          if (ch == 0) {
            r_nn = ldg_stream(g.r_e + t0 + 2 * kTileE + row);
            pm_nn = ldg_stream(g.perm + t0 + 2 * kTileE + row);
          }
        }
      }
      // radial derivative operand of the next tile: computed now (registers), stored once vbar has left the X columns
      float o_n[32], od_n[32];
      if (has_next && ch == 0) {
        rbf_regs<true>(r_n, g.inv_cutoff, o_n, od_n);
#pragma unroll
        for (int k = 0; k < kRbf; ++k) asm volatile("" : "+f"(od_n[k]));   // keep the work above the wait
      }
      // slot of this edge in the dense pair-adjoint array; its previous value is fetched early (layers 2, 1 add)
      size_t rslot = 0;
      float rold = 0.0f;
      if (ch == 0 && valid) {
        rslot = ((size_t)atom0 + tg) * g.N + (sr - (tg / g.N) * g.N);
        if (g.rbar_accumulate) rold = GNNIS_L1_POLICY >= 2 ? __ldcg(g.rbar + rslot) : g.rbar[rslot];
      }
      mbar_wait(d_full, ph_d);
      __syncwarp();
      ph_d ^= 1u;
      fence_after_sync();
      if (has_next && ch == 0) st_split24<TERMS>(lane_addr + XH, lane_addr + XL, od_n);
      // ---- ubar = vbar * silu'(u) -> shared tile ; rbar = ubar . du (this thread's 32 channels)
      mbar_wait(ldu, ph_u);
      ph_u ^= 1u;
      mbar_wait(redR, ph_r);       // every warp has finished reducing the previous tile: the tile buffer is free
      ph_r ^= 1u;
      float racc = 0.0f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = 32 * ch + 16 * c;
        const uint32_t q0 = (uint32_t)(4 * ch + 2 * c);
        uint32_t vr[16], dr[16];
        tmem_ld16(lane_addr + ACC + col0, vr);
        tmem_ld16(lane_addr + DU + col0, dr);
        const uint4 c0 = *reinterpret_cast<const uint4*>(sLU + row * 128 + ((q0 ^ swz) << 4));
        const uint4 c1 = *reinterpret_cast<const uint4*>(sLU + row * 128 + (((q0 + 1u) ^ swz) << 4));
        const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        tmem_ld_wait();
        float* urow = sU + row * kTileLd + col0;
        f2_t rab = pk(0.0f, 0.0f);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const f2_t u01 = mul2(pk(__uint_as_float(vr[4 * j4 + 0]), __uint_as_float(vr[4 * j4 + 1])), stash_dec2(cw[2 * j4]));
          const f2_t u23 = mul2(pk(__uint_as_float(vr[4 * j4 + 2]), __uint_as_float(vr[4 * j4 + 3])), stash_dec2(cw[2 * j4 + 1]));
          rab = fma2(u01, pk(__uint_as_float(dr[4 * j4 + 0]), __uint_as_float(dr[4 * j4 + 1])), rab);
          rab = fma2(u23, pk(__uint_as_float(dr[4 * j4 + 2]), __uint_as_float(dr[4 * j4 + 3])), rab);
          float4 ub;
          upk(u01, ub.x, ub.y);
          upk(u23, ub.z, ub.w);
          *reinterpret_cast<float4*>(urow + 4 * j4) = ub;
        }
        float ra, rb;
        upk(rab, ra, rb);
        racc += ra + rb;
      }
      if (has_next) {
        if (ch == 0) tmem_st_wait();
        if (warp_arrive_last(cnt_u, 8, lane)) issue_du();   // vbar / du drained, next radial operand in place
        if (warp_arrive_last_plain(cnt_s, 8, lane) && lane == 0)
          bulk_load(sLU, g.st_u + (size_t)(t0 + kTileE) * 128, (uint32_t)min(kTileE, e_end - t0 - kTileE) * 128u, ldu);
      }
      if (ch == 1) sPart[grp][row] = racc;
      warp_arrive_bar(tileB, lane);                  // this warp's part of the ubar tile is in shared memory
      rslot_p = rslot;
      rold_p = rold;
      racc_p = racc;
      valid_p = valid;
      pr = pr_n;
      r = r_n;
      pm = pm_n;
      valid = valid_n;
      pr_n = pr_nn;
      r_n = r_nn;
      pm_n = pm_nn;
      tk = tkn;
      kidx = (kidx + 1) % 3;
    }
    // ---- last tile of the unit
    mbar_wait(tileB, ph_b);
    ph_b ^= 1u;
    if (ch == 0 && valid_p) st_rbar(g.rbar + rslot_p, rold_p + (racc_p + sPart[grp][row]));
    reduce_targets_global_256(sU, keys + (kidx + 2) % 3, gPb, gt);
    reduce_sources_table_256(sU, keys + (kidx + 2) % 3, tQb, ld, gt);
    ws_group_sync(grp);
    if (QB_SM) {
      copy_rows_n(qbar_rows(), 64, tQb, ld, natoms, gt, kWsGroupThreads);
      ws_group_sync(grp);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

// Qbar = sum over the parts of a split replica, in part order (deterministic)
__global__ void sum_qbar_parts_kernel(float4* __restrict__ qbar, size_t first4, size_t n4, size_t stride4, int S) {
  pdl_trigger();
  for (size_t i = first4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = qbar[i];
    for (int p = 1; p < S; ++p) {
      const float4 b = qbar[i + (size_t)p * stride4];
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
    }
    qbar[i] = a;
  }
}

template <int T, int TERMS, bool STASH>
static int launch_fwd_t(Ctx* c, const EdgeTcArgs& g, int grid, size_t smem, uint32_t gb, cudaStream_t st) {
  if (g.S > 1) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_fwd_wsi_kernel<T, TERMS, true, STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), edge_fwd_wsi_kernel<T, TERMS, true, STASH>, grid, kWsiThreads, smem, st, g, gb));
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_fwd_wsi_kernel<T, TERMS, false, STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), edge_fwd_wsi_kernel<T, TERMS, false, STASH>, grid, kWsiThreads, smem, st, g, gb));
  }
  return 0;
}
template <int T, int TERMS>
static int launch_bwd_t(Ctx* c, const EdgeTcArgs& g, int grid, size_t smem, uint32_t gb, cudaStream_t st) {
  if (g.S > 1) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_bwd_ws_kernel<T, TERMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), edge_bwd_ws_kernel<T, TERMS, true>, grid, kWsThreads, smem, st, g, gb));
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_bwd_ws_kernel<T, TERMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), edge_bwd_ws_kernel<T, TERMS, false>, grid, kWsThreads, smem, st, g, gb));
  }
  return 0;
}
template <int T, int TERMS, bool SPLIT, bool GATED>
static int launch_bwd_st_k(Ctx* c, const EdgeTcArgs& g, int grid, size_t smem, uint32_t gb, cudaStream_t st) {
  GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_bwd_st_kernel<T, TERMS, SPLIT, GATED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GNNIS_CUDA_OK(launch_k(c->pdl_now(), edge_bwd_st_kernel<T, TERMS, SPLIT, GATED>, grid, kWsThreads, smem, st, g, gb));
  return 0;
}
template <int T, int TERMS>
static int launch_bwd_st_t(Ctx* c, const EdgeTcArgs& g, int grid, size_t smem, uint32_t gb, cudaStream_t st) {
  if (g.n_edges) return g.S > 1 ? launch_bwd_st_k<T, TERMS, true, true>(c, g, grid, smem, gb, st) : launch_bwd_st_k<T, TERMS, false, true>(c, g, grid, smem, gb, st);
  return g.S > 1 ? launch_bwd_st_k<T, TERMS, true, false>(c, g, grid, smem, gb, st) : launch_bwd_st_k<T, TERMS, false, false>(c, g, grid, smem, gb, st);
}

// terms = 3: 3xTF32 (fp32-accurate, default); terms = 1: one TF32 product per contraction (GNNIS_FLAG_MLP_TF32X1).
// stash = 1: also leave silu'(u), silu'(w) for edge_bwd_st_kernel (evaluations with forces, Ctx::stash_on).
// dispatch on the tensor-core arithmetic of the forward kernels: 3 = 3xTF32, 1 = TF32, kBf16 / kFp16 = 16-bit operands + tanh activations
template <int T, bool STASH>
static int launch_fwd_terms(Ctx* c, const EdgeTcArgs& g, int terms, int grid, size_t smem, uint32_t gb, cudaStream_t st) {
  if (terms == kBf16) return launch_fwd_t<T, kBf16, STASH>(c, g, grid, smem, gb, st);
  if (terms == kFp16) return launch_fwd_t<T, kFp16, STASH>(c, g, grid, smem, gb, st);
  if (terms == kTanh3) return launch_fwd_t<T, kTanh3, STASH>(c, g, grid, smem, gb, st);
  if (terms == 1) return launch_fwd_t<T, 1, STASH>(c, g, grid, smem, gb, st);
  return launch_fwd_t<T, 3, STASH>(c, g, grid, smem, gb, st);
}
template <int T>
static int launch_bwd_st_terms(Ctx* c, const EdgeTcArgs& g, int terms, int grid, size_t smem, uint32_t gb, cudaStream_t st) {
  if (terms == kBf16) return launch_bwd_st_t<T, kBf16>(c, g, grid, smem, gb, st);
  if (terms == kFp16) return launch_bwd_st_t<T, kFp16>(c, g, grid, smem, gb, st);
  if (terms == 1) return launch_bwd_st_t<T, 1>(c, g, grid, smem, gb, st);
  return launch_bwd_st_t<T, 3>(c, g, grid, smem, gb, st);
}

int launch_edge_fwd_ws(Ctx* c, int l, int terms, int stash, cudaStream_t st) {
  EdgeTcArgs g = make_tc_args(c, l, terms);
  const int grid = tc_grid(c);
  int rc;
  if (stash) {
    const int tab = st_fwd_tables(c);
    if (tab < 0) {
      c->err = "edge forward (stash): shared memory does not fit";
      return -1;
    }
    const size_t smem = st_fwd_smem(c, tab);
    const uint32_t gb = (uint32_t)st_group_bytes(c, tab);
    if (tab == 2) rc = launch_fwd_terms<2, true>(c, g, terms, grid, smem, gb, st);
    else if (tab == 1) rc = launch_fwd_terms<1, true>(c, g, terms, grid, smem, gb, st);
    else if (tab == 3) rc = launch_fwd_terms<3, true>(c, g, terms, grid, smem, gb, st);
    else rc = launch_fwd_terms<0, true>(c, g, terms, grid, smem, gb, st);
  } else {
    const size_t smem = fwd_tc_smem(c);
    const uint32_t gb = (uint32_t)group_bytes(c);
    if (c->tab_mode == 2) rc = launch_fwd_terms<2, false>(c, g, terms, grid, smem, gb, st);
    else if (c->tab_mode == 1) rc = launch_fwd_terms<1, false>(c, g, terms, grid, smem, gb, st);
    else rc = launch_fwd_terms<0, false>(c, g, terms, grid, smem, gb, st);
  }
  if (rc) return rc;
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_edge_bwd_ws(Ctx* c, int l, int rbar_accumulate, int terms, int stash, cudaStream_t st) {
  const int terms_in = terms;
  if (!stash && is_half(terms)) terms = 1;   // the recomputing adjoint has no 16-bit form: single-pass TF32 products instead
  EdgeTcArgs g = make_tc_args(c, l, terms);
  g.rbar_accumulate = rbar_accumulate;
  const int grid = tc_grid(c);
  int rc;
  if (stash) {
    const int tab = st_bwd_tables(c);
    if (tab < 0) {
      c->err = "edge adjoint (stash): shared memory does not fit";
      return -1;
    }
    const size_t smem = st_bwd_smem(c, tab);
    const uint32_t gb = (uint32_t)st_group_bytes(c, tab);
    g.stash_role = 1;
    if (tab == 1) rc = launch_bwd_st_terms<1>(c, g, terms, grid, smem, gb, st);
    else rc = launch_bwd_st_terms<0>(c, g, terms, grid, smem, gb, st);
    if (rc) return rc;
    GNNIS_LAUNCH_CHECK();
    if (!g.n_edges) stash = -1;   // the stash holds every edge list: no alternative launch
  }
  if (stash >= 0) {
    // the recomputing adjoint: the only launch without a stash, the alternative launch of a capacity-limited one (it
    // returns at once when the stash adjoint above has run; the products then follow the default arithmetic)
    if (stash) {
      terms = is_half(terms_in) ? 1 : (terms_in == 1 ? 1 : 3);
      g = make_tc_args(c, l, terms);
      g.rbar_accumulate = rbar_accumulate;
      g.stash_role = 2;
    }
    const size_t smem = bwd_tc_smem(c);
    const uint32_t gb = (uint32_t)group_bytes(c);
    if (c->tab_mode == 2) rc = terms == 1 ? launch_bwd_t<2, 1>(c, g, grid, smem, gb, st) : launch_bwd_t<2, 3>(c, g, grid, smem, gb, st);
    else if (c->tab_mode == 1) rc = terms == 1 ? launch_bwd_t<1, 1>(c, g, grid, smem, gb, st) : launch_bwd_t<1, 3>(c, g, grid, smem, gb, st);
    else rc = terms == 1 ? launch_bwd_t<0, 1>(c, g, grid, smem, gb, st) : launch_bwd_t<0, 3>(c, g, grid, smem, gb, st);
    if (rc) return rc;
    GNNIS_LAUNCH_CHECK();
  }
  if (c->split_eff > 1) {
    const size_t n4 = (size_t)c->R * c->N * 16, first4 = (size_t)c->split_from * c->N * 16;   // only split replicas have parts
    const size_t w4 = n4 - first4;
    const int blocks = (int)((w4 + 255) / 256 < 4 * (size_t)c->num_sms ? (w4 + 255) / 256 : 4 * (size_t)c->num_sms);
    sum_qbar_parts_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<float4*>(c->Qbar), first4, n4, n4, c->split_eff);
    GNNIS_LAUNCH_CHECK();
  }
  return 0;
}

int launch_umma_selftest(Ctx* c, const float* A, const float* B, float* D, int K, int N, cudaStream_t st) {
  size_t smem = 32768 + 1024;
  GNNIS_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 256, smem, st>>>(A, B, D, K, N);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

}  // namespace gnnis
