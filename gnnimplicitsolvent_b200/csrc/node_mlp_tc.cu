// node_mlp_tc.cu -- the per-atom ("nodewise") MLPs of IN_layer_all_swish_2pass_tokens (GNN_Layers.py:2231-2246)
// and their adjoints on the tensor cores (tcgen05, 3xTF32, sm_100a).  Same arithmetic as node_mlp.cu (which stays
// as the CUDA-core cross-check).
//
// Rows are atoms, 128 per tile.  One persistent CTA per SM with TWO independent groups of 8 warps: each group walks
// its own tiles (thread = row = TMEM lane x column half) and owns 256 of the 512 TMEM columns, so the products of one
// group run under the epilogue (global loads, SiLU, hi / lo split) of the other.  A operands are written to TMEM by
// the epilogue of the previous product; weights are staged once per CTA as K-major 128B-swizzled hi / lo images.
// Per group the TMEM columns are R0 = [0,128) (A operand: hi | lo of at most 64 features) and R1 = [128,256)
// (accumulators); products over 128 features are issued as two K halves so that the operand fits R0:
//
//   forward  (layer l):  p  = a W3a^T                              R0 = a        -> R1 = p        (N 128, K 64)
//                        o  = silu(p + sbias[s]) W4^T + b4         R0 = g[:, h]  -> R1[0,64) = o  (N 64, K 64 + 64)
//                        [P|Q]_{l+1} = silu(o) [Wi;Wj]^T (+ b1)    R0 = silu(o)  -> R1 = [P|Q]    (N 128, K 64)
//            last layer: o3 = silu(p) W4^T + b4 (2 outputs, CUDA cores) and the scale / SA head
//   adjoint A (l < 2):   hbar = Pbar Wi + Qbar Wj                  R0 = Pbar, Qbar -> R1[0,64)    (N 64, K 64 + 64)
//                        gbar = (hbar silu'(o)) W4                 R0 = obar     -> R1 = gbar     (N 128, K 64)
//   adjoint B:           p recomputed as in the forward pass;  pbar = gbar silu'(p + sbias)
//                        abar = pbar W3a                           R0 = pbar[:, h] -> R1[0,64)    (N 64, K 64 + 64)
#include "common.cuh"
#include "umma.cuh"
#include "fastmath.cuh"

namespace gnnis {

using namespace umma;

int launch_build_operand(Ctx*, uint8_t*, uint8_t*, const float*, int, int, int, int, int, int, cudaStream_t);

constexpr int kNtThreads = 512;   // 2 groups x 8 warps: thread = (row, column half)
constexpr int kNtRows = 128;

struct NodeTcArgs {
  int n, N;
  const int* solvent_id;
  const int* skip;   // per replica: 2 = finished (minimiser), tiles made only of such replicas are skipped
  const float *a, *sbias, *W4, *b4, *b1_next;
  const uint8_t* img;   // operand image of this kernel (see Ctx::img_*)
  float *o, *P_next, *Q_next;
  // head
  const float *born, *rho, *gamma_emb;
  float fraction, kappa;
  float *born_s, *sa_atom;
  // adjoint
  const float *Pbar, *Qbar, *dE_dBs;
  float *gbar, *abar, *Bbar;
};

__device__ __forceinline__ void nt_group_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

// D[128 x N] (+)= A[128 x 8 (S1 - S0)] * B[:, 8 S0 .. 8 S1)^T (3xTF32): A (hi, lo) in TMEM starting at column 0 of
// its region, B K-major SW128 with b_rows rows
template <int S0, int S1>
__device__ __forceinline__ void nt_issue(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                         int b_rows, uint32_t idesc, uint32_t accumulate) {
  uint32_t acc = accumulate;
#pragma unroll
  for (int term = 0; term < 3; ++term) {
    const uint32_t a = term == 2 ? a_lo : a_hi;
    const uint32_t b = term == 1 ? b_lo : b_hi;
#pragma unroll
    for (int s = S0; s < S1; ++s) {
      uint32_t baddr = b + (uint32_t)(s >> 2) * (uint32_t)b_rows * 128u + (uint32_t)(s & 3) * 32u;
      mma_tf32_ts(d_tmem, a + 8u * (uint32_t)(s - S0), smem_desc_k_sw128(baddr), idesc, acc);
      acc = 1;
    }
  }
}

__device__ __forceinline__ void nt_store_a16(uint32_t t_hi, uint32_t t_lo, const float4 (&x)[4]) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[4 * j + 0] = __float_as_uint(x[j].x) & 0xFFFFE000u;
    h[4 * j + 1] = __float_as_uint(x[j].y) & 0xFFFFE000u;
    h[4 * j + 2] = __float_as_uint(x[j].z) & 0xFFFFE000u;
    h[4 * j + 3] = __float_as_uint(x[j].w) & 0xFFFFE000u;
    float l0, l1, l2, l3;
    upk(sub2(pk(x[j].x, x[j].y), pk(__uint_as_float(h[4 * j]), __uint_as_float(h[4 * j + 1]))), l0, l1);
    upk(sub2(pk(x[j].z, x[j].w), pk(__uint_as_float(h[4 * j + 2]), __uint_as_float(h[4 * j + 3]))), l2, l3);
    l[4 * j + 0] = __float_as_uint(l0);
    l[4 * j + 1] = __float_as_uint(l1);
    l[4 * j + 2] = __float_as_uint(l2);
    l[4 * j + 3] = __float_as_uint(l3);
  }
  tmem_st16(t_hi, h);
  tmem_st16(t_lo, l);
}

// 16 consecutive floats of a global row (zero when the row is past the end)
__device__ __forceinline__ void nt_load16(const float* __restrict__ p, bool ok, float4 (&x)[4]) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4)
    x[j4] = ok ? __ldg(reinterpret_cast<const float4*>(p) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void nt_store16(float* __restrict__ p, const float4 (&x)[4]) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) *(reinterpret_cast<float4*>(p) + j4) = x[j4];
}
// Row-per-lane accesses to global memory cost one L1 wavefront per lane (32 different 128-byte lines per
// instruction), which made the plain row loads / stores the longest phases of these kernels.  The 32 rows x 16 floats
// block of a warp therefore goes through a 2 KB shared-memory slice: global side = 4 lanes per row (64 contiguous
// bytes, 8 rows per instruction), register side = row per lane; quads are XOR-swizzled so both sides are free of
// bank conflicts.  g0 = first row of the warp's block (already offset to the chunk's first column).
struct NtRows {
  float* stage;   // this warp's [32][16] slice
  int nvalid;     // rows of the warp's block that exist (0..32)
  int lane;
};
__device__ __forceinline__ int nt_swz(int row, int q) { return row * 16 + 4 * (q ^ ((row >> 1) & 3)); }
__device__ __forceinline__ void nt_warp_load16(const NtRows& w, const float* __restrict__ g0, int ld, float4 (&x)[4]) {
  const int q = w.lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = 8 * i + (w.lane >> 2);
    const float4 v = row < w.nvalid ? __ldg(reinterpret_cast<const float4*>(g0 + (size_t)row * ld) + q)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(w.stage + nt_swz(row, q)) = v;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j] = *reinterpret_cast<const float4*>(w.stage + nt_swz(w.lane, j));
  __syncwarp();
}
__device__ __forceinline__ void nt_warp_store16(const NtRows& w, float* __restrict__ g0, int ld, const float4 (&x)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(w.stage + nt_swz(w.lane, j)) = x[j];
  __syncwarp();
  const int q = w.lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = 8 * i + (w.lane >> 2);
    const float4 v = *reinterpret_cast<const float4*>(w.stage + nt_swz(row, q));
    if (row < w.nvalid) *(reinterpret_cast<float4*>(g0 + (size_t)row * ld) + q) = v;
  }
  __syncwarp();
}
__device__ __forceinline__ float4 nt_acc4(const uint32_t* r) {
  return make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
}
__device__ __forceinline__ float4 nt_add4(float4 a, float4 b) {
  float4 y;
  upk(add2(pk(a.x, a.y), pk(b.x, b.y)), y.x, y.y);
  upk(add2(pk(a.z, a.w), pk(b.z, b.w)), y.z, y.w);
  return y;
}

#define NT_PROLOGUE()                                                              \
  extern __shared__ uint8_t smraw[];                                               \
  __shared__ uint64_t bar[2], ibar;                                                \
  __shared__ uint32_t tmem_slot;                                                   \
  uint32_t base_ = smem_u32(smraw);                                                \
  uint8_t* sm = smraw + ((1024u - (base_ & 1023u)) & 1023u);                       \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                   \
  const int grp = warp >> 3, ch = (warp >> 2) & 1, wq = warp & 3, row = 32 * wq + lane; \
  pdl_trigger();                                                                   \
  if (warp == 0) tmem_alloc(&tmem_slot, 512);                                      \
  if (tid == 0) {                                                                  \
    mbar_init(&bar[0], 1);                                                         \
    mbar_init(&bar[1], 1);                                                         \
    mbar_init(&ibar, 1);                                                           \
  }

#define NT_STAGED()                                                                \
  fence_proxy_async_smem();                                                        \
  fence_before_sync();                                                             \
  __syncthreads();                                                                 \
  fence_after_sync();                                                              \
  mbar_wait(&ibar, 0);   /* operand images have landed */                          \
  __syncwarp();                                                                    \
  pdl_wait();            /* the evaluation's data from here on */                  \
  const uint32_t tm = tmem_slot + 256u * (uint32_t)grp;                            \
  const uint32_t lane_addr = tm + ((uint32_t)(32 * wq) << 16);                     \
  uint32_t phase = 0;                                                              \
  (void)lane;

// this group: A operand stored -> one elected lane of the group's first warp issues
#define NT_ISSUE(...)                                                              \
  tmem_st_wait();                                                                  \
  fence_before_sync();                                                             \
  nt_group_sync(grp);                                                              \
  if ((warp & 7) == 0) {                                                           \
    fence_after_sync();                                                            \
    if (elect_one()) {                                                             \
      __VA_ARGS__;                                                                 \
      mma_commit(&bar[grp]);                                                       \
    }                                                                              \
    __syncwarp();                                                                  \
  }
// ... and every thread of the group waits for the product
#define NT_WAIT()                                                                  \
  mbar_wait(&bar[grp], phase);                                                     \
  __syncwarp();                                                                    \
  phase ^= 1;                                                                      \
  fence_after_sync();
#define NT_GEMM(...)    \
  NT_ISSUE(__VA_ARGS__) \
  NT_WAIT()

// true when every replica touched by the 128-atom tile is finished (uniform over the group)
__device__ __forceinline__ bool nt_tile_skipped(const NodeTcArgs& g, int tile) {
  if (!g.skip) return false;
  const int r0 = (tile * kNtRows) / g.N, r1 = (min(g.n, tile * kNtRows + kNtRows) - 1) / g.N;
  for (int r = r0; r <= r1; ++r)
    if (g.skip[r] != 2) return false;
  return true;
}

#define NT_EPILOGUE()                                                              \
  fence_before_sync();                                                             \
  __syncthreads();                                                                 \
  if (warp == 0) tmem_dealloc(tmem_slot, 512);

// one 64-feature global row -> hi | lo operand in R0
__device__ __forceinline__ void nt_row64_to_r0(uint32_t lane_addr, const NtRows& w, const float* __restrict__ src0, int ch) {
#pragma unroll
  for (int c = 2 * ch; c < 2 * ch + 2; ++c) {
    float4 x[4];
    nt_warp_load16(w, src0 + 16 * c, 64, x);
    nt_store_a16(lane_addr + 16 * c, lane_addr + 64 + 16 * c, x);
  }
}

// ------------------------------------------------------------------------------------------ forward
// smem: W3a (128 x 64) hi|lo 64 KB | W4 (64 x 128) hi|lo 64 KB | [Wi;Wj]_next (128 x 64) hi|lo 64 KB
template <bool LAST>
__global__ void __launch_bounds__(kNtThreads, 1) node_fwd_tc_kernel(NodeTcArgs g) {
  NT_PROLOGUE();
  uint8_t *sW3H = sm, *sW3L = sm + 32768, *sW4H = sm + 65536, *sW4L = sm + 98304;
  uint8_t *sWpH = sm + 131072, *sWpL = sm + 163840;
  float* sStage = reinterpret_cast<float*>(sm + (LAST ? 65536 : 196608));   // [16 warps][32][16] row staging
  if (tid == 0) bulk_copy_image(sm, g.img, LAST ? 65536 : 196608, &ibar);
  NT_STAGED();
  constexpr uint32_t idesc128 = idesc_tf32(128, 128), idesc64 = idesc_tf32(128, 64);
  const uint32_t aW3H = smem_u32(sW3H), aW3L = smem_u32(sW3L), aW4H = smem_u32(sW4H), aW4L = smem_u32(sW4L);
  const uint32_t aWpH = smem_u32(sWpH), aWpL = smem_u32(sWpL);
  (void)aW4H; (void)aW4L; (void)aWpH; (void)aWpL; (void)idesc64;
  const int ntiles = (g.n + kNtRows - 1) / kNtRows;

  for (int tile = blockIdx.x + grp * gridDim.x; tile < ntiles; tile += 2 * gridDim.x) {
    if (nt_tile_skipped(g, tile)) continue;
    const int r = tile * kNtRows + row;
    const bool ok = r < g.n;
    const int rr = ok ? r : g.n - 1;
    const int r0w = tile * kNtRows + 32 * wq;   // first row of this warp's 32-row block
    const NtRows wr{sStage + 512 * warp, max(0, min(32, g.n - r0w)), lane};
    nt_row64_to_r0(lane_addr, wr, g.a + (size_t)r0w * 64, ch);
    NT_GEMM(nt_issue<0, 8>(tm + 128, tm + 0, tm + 64, aW3H, aW3L, 128, idesc128, 0u));
    const float* sb = g.sbias + (size_t)g.solvent_id[rr / g.N] * 128;
    if constexpr (LAST) {
      // o3 = W4 silu(p + sbias) + b4 (2 outputs) and the head: B' = B (f + kappa sig(c)), SA = gamma_s sig(s) (rho + off + 0.14)^2
      float h0 = 0.0f, h1 = 0.0f;
#pragma unroll
      for (int c = 4 * ch; c < 4 * ch + 4; ++c) {
        uint32_t pr[16];
        tmem_ld16(lane_addr + 128 + 16 * c, pr);
        float4 bs[4];
        nt_load16(sb + 16 * c, true, bs);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 gv = silu4_sum2(&pr[4 * j], bs[j]);
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(g.W4 + 16 * c) + j);
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(g.W4 + 128 + 16 * c) + j);
          h0 = fmaf(gv.x, w0.x, h0);
          h0 = fmaf(gv.y, w0.y, h0);
          h0 = fmaf(gv.z, w0.z, h0);
          h0 = fmaf(gv.w, w0.w, h0);
          h1 = fmaf(gv.x, w1.x, h1);
          h1 = fmaf(gv.y, w1.y, h1);
          h1 = fmaf(gv.z, w1.z, h1);
          h1 = fmaf(gv.w, w1.w, h1);
        }
      }
      float* sHead = reinterpret_cast<float*>(sm + 65536 + 32768) + 256 * grp;   // [128][2] partial sums of column half 1
      if (ch == 1) {
        sHead[2 * row] = h0;
        sHead[2 * row + 1] = h1;
      }
      nt_group_sync(grp);
      if (ch == 0 && ok) {
        const float c0 = (h0 + sHead[2 * row]) + g.b4[0], c1 = (h1 + sHead[2 * row + 1]) + g.b4[1];
        g.o[2 * (size_t)r] = c0;
        g.o[2 * (size_t)r + 1] = c1;
        const int a = r % g.N;
        const float rp = g.rho[a] + kOffset + kProbe;
        g.sa_atom[r] = g.gamma_emb[g.solvent_id[r / g.N]] * (sigmoidf_(c1) * (rp * rp));
        g.born_s[r] = g.born[r] * (g.fraction + sigmoidf_(c0) * g.kappa);
      }
      continue;
    }
    if constexpr (!LAST) {
    // ---- g = silu(p + sbias[s]), first 64 features -> R0
#pragma unroll
    for (int c = 2 * ch; c < 2 * ch + 2; ++c) {
      uint32_t pr[16];
      tmem_ld16(lane_addr + 128 + 16 * c, pr);
      float4 bs[4], gv[4];
      nt_load16(sb + 16 * c, true, bs);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] = silu4_sum2(&pr[4 * j], bs[j]);
      nt_store_a16(lane_addr + 16 * c, lane_addr + 64 + 16 * c, gv);
    }
    NT_ISSUE(nt_issue<0, 8>(tm + 128, tm + 0, tm + 64, aW4H, aW4L, 64, idesc64, 0u));
    // ---- second 64 features are evaluated while the first half product runs (it writes R1[0,64), read above)
    float4 g2[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t pr[16];
      tmem_ld16(lane_addr + 192 + 16 * (2 * ch + c), pr);
      float4 bs[4];
      nt_load16(sb + 64 + 16 * (2 * ch + c), true, bs);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) g2[4 * c + j] = silu4_sum2(&pr[4 * j], bs[j]);
    }
    NT_WAIT();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float4(&gv)[4] = *reinterpret_cast<const float4(*)[4]>(&g2[4 * c]);
      nt_store_a16(lane_addr + 16 * (2 * ch + c), lane_addr + 64 + 16 * (2 * ch + c), gv);
    }
    NT_GEMM(nt_issue<8, 16>(tm + 128, tm + 0, tm + 64, aW4H, aW4L, 64, idesc64, 1u));
    // ---- o = acc + b4 -> global ; h = silu(o) -> R0
#pragma unroll
    for (int c = 2 * ch; c < 2 * ch + 2; ++c) {
      uint32_t orr[16];
      tmem_ld16(lane_addr + 128 + 16 * c, orr);
      float4 bs[4];
      nt_load16(g.b4 + 16 * c, true, bs);
      tmem_ld_wait();
      float4 ov[4], hv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ov[j] = nt_add4(nt_acc4(&orr[4 * j]), bs[j]);
        hv[j] = silu4(ov[j]);
      }
      nt_warp_store16(wr, g.o + (size_t)r0w * 64 + 16 * c, 64, ov);
      nt_store_a16(lane_addr + 16 * c, lane_addr + 64 + 16 * c, hv);
    }
    NT_GEMM(nt_issue<0, 8>(tm + 128, tm + 0, tm + 64, aWpH, aWpL, 128, idesc128, 0u));
    // ---- tables of the next layer: columns [0,64) -> P = . + b1 (column half 0), [64,128) -> Q (half 1)
#pragma unroll
    for (int c = 4 * ch; c < 4 * ch + 4; ++c) {
      uint32_t ar[16];
      tmem_ld16(lane_addr + 128 + 16 * c, ar);
      float4 bs[4];
      nt_load16(g.b1_next + 16 * (c & 3), c < 4, bs);
      tmem_ld_wait();
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = nt_add4(nt_acc4(&ar[4 * j]), bs[j]);
      nt_warp_store16(wr, (c < 4 ? g.P_next : g.Q_next) + (size_t)r0w * 64 + 16 * (c & 3), 64, v);
    }
    }   // !LAST
  }
  NT_EPILOGUE();
}

// ------------------------------------------------------------------------------------------ adjoint A (layers 1, 2)
// smem: [Wi;Wj]^T as B[n = c_in (64)][k = (which, c_out) (128)] hi|lo 64 KB | W4^T as B[n = m (128)][k = c (64)] hi|lo 64 KB
__global__ void __launch_bounds__(kNtThreads, 1) node_bwd_a_tc_kernel(NodeTcArgs g) {
  NT_PROLOGUE();
  uint8_t *sWpH = sm, *sWpL = sm + 32768, *sW4H = sm + 65536, *sW4L = sm + 98304;
  float* sStage = reinterpret_cast<float*>(sm + 131072);
  if (tid == 0) bulk_copy_image(sm, g.img, 131072, &ibar);   // B[n = c_in][k = c_out] = Wi|Wj[c_out][c_in] ; B[n = m][k = c] = W4[c][m]
  NT_STAGED();
  constexpr uint32_t idesc128 = idesc_tf32(128, 128), idesc64 = idesc_tf32(128, 64);
  const uint32_t aWpH = smem_u32(sWpH), aWpL = smem_u32(sWpL), aW4H = smem_u32(sW4H), aW4L = smem_u32(sW4L);
  const int ntiles = (g.n + kNtRows - 1) / kNtRows;

  for (int tile = blockIdx.x + grp * gridDim.x; tile < ntiles; tile += 2 * gridDim.x) {
    if (nt_tile_skipped(g, tile)) continue;
    const int r0w = tile * kNtRows + 32 * wq;   // first row of this warp's 32-row block
    const NtRows wr{sStage + 512 * warp, max(0, min(32, g.n - r0w)), lane};
    nt_row64_to_r0(lane_addr, wr, g.Pbar + (size_t)r0w * 64, ch);
    NT_ISSUE(nt_issue<0, 8>(tm + 128, tm + 0, tm + 64, aWpH, aWpL, 64, idesc64, 0u));
    float4 q[8];   // the Qbar row is fetched while the Pbar half runs
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float4(&qc)[4] = *reinterpret_cast<float4(*)[4]>(&q[4 * c]);
      nt_warp_load16(wr, g.Qbar + (size_t)r0w * 64 + 16 * (2 * ch + c), 64, qc);
    }
    NT_WAIT();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float4(&qc)[4] = *reinterpret_cast<const float4(*)[4]>(&q[4 * c]);
      nt_store_a16(lane_addr + 16 * (2 * ch + c), lane_addr + 64 + 16 * (2 * ch + c), qc);
    }
    NT_GEMM(nt_issue<8, 16>(tm + 128, tm + 0, tm + 64, aWpH, aWpL, 64, idesc64, 1u));
    // ---- obar = hbar * silu'(o) -> R0
#pragma unroll
    for (int c = 2 * ch; c < 2 * ch + 2; ++c) {
      uint32_t hr[16];
      tmem_ld16(lane_addr + 128 + 16 * c, hr);
      float4 ov[4], v[4];
      nt_warp_load16(wr, g.o + (size_t)r0w * 64 + 16 * c, 64, ov);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = silu_d4_scaled(ov[j], nt_acc4(&hr[4 * j]));
      nt_store_a16(lane_addr + 16 * c, lane_addr + 64 + 16 * c, v);
    }
    NT_GEMM(nt_issue<0, 8>(tm + 128, tm + 0, tm + 64, aW4H, aW4L, 128, idesc128, 0u));
    // ---- gbar -> global scratch
#pragma unroll
    for (int c = 4 * ch; c < 4 * ch + 4; ++c) {
      uint32_t gr[16];
      tmem_ld16(lane_addr + 128 + 16 * c, gr);
      tmem_ld_wait();
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = nt_acc4(&gr[4 * j]);
      nt_warp_store16(wr, g.gbar + (size_t)r0w * 128 + 16 * c, 128, v);
    }
  }
  NT_EPILOGUE();
}

// ------------------------------------------------------------------------------------------ adjoint B
// smem: W3a (128 x 64) hi|lo 64 KB | W3a^T as B[n = k_in (64)][k = m (128)] hi|lo 64 KB
template <bool LAST>
__global__ void __launch_bounds__(kNtThreads, 1) node_bwd_b_tc_kernel(NodeTcArgs g) {
  NT_PROLOGUE();
  uint8_t *sW3H = sm, *sW3L = sm + 32768, *sWtH = sm + 65536, *sWtL = sm + 98304;
  float* sStage = reinterpret_cast<float*>(sm + 131072);
  if (tid == 0) bulk_copy_image(sm, g.img, 131072, &ibar);   // B[n = m][k] = W3a[m][k] ; B[n = k_in][k = m] = W3a[m][k_in]
  NT_STAGED();
  constexpr uint32_t idesc128 = idesc_tf32(128, 128), idesc64 = idesc_tf32(128, 64);
  const uint32_t aW3H = smem_u32(sW3H), aW3L = smem_u32(sW3L), aWtH = smem_u32(sWtH), aWtL = smem_u32(sWtL);
  const int ntiles = (g.n + kNtRows - 1) / kNtRows;

  for (int tile = blockIdx.x + grp * gridDim.x; tile < ntiles; tile += 2 * gridDim.x) {
    if (nt_tile_skipped(g, tile)) continue;
    const int r = tile * kNtRows + row;
    const bool ok = r < g.n;
    const int rr = ok ? r : g.n - 1;
    const int r0w = tile * kNtRows + 32 * wq;   // first row of this warp's 32-row block
    const NtRows wr{sStage + 512 * warp, max(0, min(32, g.n - r0w)), lane};
    nt_row64_to_r0(lane_addr, wr, g.a + (size_t)r0w * 64, ch);
    NT_ISSUE(nt_issue<0, 8>(tm + 128, tm + 0, tm + 64, aW3H, aW3L, 128, idesc128, 0u));
    float cb = 0.0f, sbv = 0.0f;
    if (LAST && ok) {
      // head adjoint: cbar, sbar from dE/dB' and the SA term; Bbar += dE/dB' dB'/dB
      const float c0 = g.o[2 * (size_t)r], c1 = g.o[2 * (size_t)r + 1];
      const float sc = sigmoidf_(c0), ss = sigmoidf_(c1);
      const float dEdBs = g.dE_dBs[r];
      cb = dEdBs * g.born[r] * g.kappa * sc * (1.0f - sc);
      const int a = r % g.N;
      const float rp = g.rho[a] + kOffset + kProbe;
      sbv = g.gamma_emb[g.solvent_id[r / g.N]] * (rp * rp) * ss * (1.0f - ss);
      if (ch == 0) g.Bbar[r] += dEdBs * (g.fraction + g.kappa * sc);
    }
    NT_WAIT();
    const float* sb = g.sbias + (size_t)g.solvent_id[rr / g.N] * 128;
    // pbar = gbar * silu'(p + sbias[s]) for 16 features starting at col0
    auto pbar16 = [&](int col0, float4(&v)[4]) {
      uint32_t pr[16];
      tmem_ld16(lane_addr + 128 + col0, pr);
      float4 bs[4], gb[4];
      nt_load16(sb + col0, true, bs);
      if (LAST) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(g.W4 + col0) + j);
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(g.W4 + 128 + col0) + j);
          gb[j] = make_float4(cb * w0.x + sbv * w1.x, cb * w0.y + sbv * w1.y, cb * w0.z + sbv * w1.z,
                              cb * w0.w + sbv * w1.w);
        }
      } else {
        nt_warp_load16(wr, g.gbar + (size_t)r0w * 128 + col0, 128, gb);
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = silu_d4_sum2_scaled(&pr[4 * j], bs[j], gb[j]);
    };
#pragma unroll
    for (int c = 2 * ch; c < 2 * ch + 2; ++c) {
      float4 v[4];
      pbar16(16 * c, v);
      nt_store_a16(lane_addr + 16 * c, lane_addr + 64 + 16 * c, v);
    }
    NT_ISSUE(nt_issue<0, 8>(tm + 128, tm + 0, tm + 64, aWtH, aWtL, 64, idesc64, 0u));
    float4 v2[8];   // second 64 features, evaluated while the first half product runs
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float4(&vc)[4] = *reinterpret_cast<float4(*)[4]>(&v2[4 * c]);
      pbar16(64 + 16 * (2 * ch + c), vc);
    }
    NT_WAIT();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float4(&vc)[4] = *reinterpret_cast<const float4(*)[4]>(&v2[4 * c]);
      nt_store_a16(lane_addr + 16 * (2 * ch + c), lane_addr + 64 + 16 * (2 * ch + c), vc);
    }
    NT_GEMM(nt_issue<8, 16>(tm + 128, tm + 0, tm + 64, aWtH, aWtL, 64, idesc64, 1u));
    // ---- abar -> global
#pragma unroll
    for (int c = 2 * ch; c < 2 * ch + 2; ++c) {
      uint32_t ar[16];
      tmem_ld16(lane_addr + 128 + 16 * c, ar);
      tmem_ld_wait();
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = nt_acc4(&ar[4 * j]);
      nt_warp_store16(wr, g.abar + (size_t)r0w * 64 + 16 * c, 64, v);
    }
  }
  NT_EPILOGUE();
}

// ------------------------------------------------------------------------------------------ launchers
static NodeTcArgs nt_args(Ctx* c, int l) {
  NodeTcArgs g{};
  g.n = c->R * c->N;
  g.N = c->N;
  g.solvent_id = c->solvent_id;
  g.skip = c->skip_state;
  g.a = c->a[l];
  g.o = c->o[l];
  g.sbias = c->L[l].sbias;
  g.W4 = c->L[l].W4;
  g.b4 = c->L[l].b4;
  if (l < 2) {
    g.P_next = c->P[l + 1];
    g.Q_next = c->Q[l + 1];
    g.b1_next = c->L[l + 1].b1;
  }
  g.born = c->born;
  g.rho = c->rho;
  g.gamma_emb = c->gamma_emb;
  g.fraction = c->fraction;
  g.kappa = (1.0f - c->fraction) * c->scaling;
  g.born_s = c->born_s;
  g.sa_atom = c->sa_atom;
  g.Pbar = c->Pbar;
  g.Qbar = c->Qbar;
  g.dE_dBs = c->dE_dBs;
  g.gbar = c->gbar;
  g.abar = c->abar;
  g.Bbar = c->Bbar;
  return g;
}

static int nt_grid(const Ctx* c) {   // tiles are dealt out CTA-first; the second warp group of a CTA takes tile blockIdx + grid
  int ntiles = (c->R * c->N + kNtRows - 1) / kNtRows;
  return ntiles < c->num_sms ? ntiles : c->num_sms;
}

int build_node_images(Ctx* c, cudaStream_t st) {
  for (int l = 0; l < 3; ++l) {
    const LayerW& L = c->L[l];
    uint8_t* f = c->img_nf[l];
    if (launch_build_operand(c, f, f + 32768, L.W3a, 128, 0, 128, 0, 64, 64, st)) return -1;
    if (l < 2) {
      const LayerW& Ln = c->L[l + 1];
      uint8_t* a = c->img_nba[l];
      if (launch_build_operand(c, f + 65536, f + 98304, L.W4, 64, 0, 64, 0, 128, 128, st) ||
          launch_build_operand(c, f + 131072, f + 163840, Ln.Wi, 128, 0, 64, 0, 64, 64, st) ||
          launch_build_operand(c, f + 131072, f + 163840, Ln.Wj, 128, 64, 64, 0, 64, 64, st) ||
          launch_build_operand(c, a, a + 32768, Ln.Wi_t, 64, 0, 64, 0, 64, 64, st) ||
          launch_build_operand(c, a, a + 32768, Ln.Wj_t, 64, 0, 64, 64, 64, 64, st) ||
          launch_build_operand(c, a + 65536, a + 98304, L.W4_t, 128, 0, 128, 0, 64, 64, st))
        return -1;
    }
    uint8_t* b = c->img_nbb[l];
    if (launch_build_operand(c, b, b + 32768, L.W3a, 128, 0, 128, 0, 64, 64, st) ||
        launch_build_operand(c, b + 65536, b + 98304, L.W3a_t, 64, 0, 64, 0, 128, 128, st))
      return -1;
  }
  return 0;
}

int launch_node_fwd_tc(Ctx* c, int l, cudaStream_t st) {
  NodeTcArgs g = nt_args(c, l);
  g.img = c->img_nf[l];
  if (l == 2) {
    size_t smem = 65536 + 32768 + 2048 + 1024;
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), node_fwd_tc_kernel<true>, nt_grid(c), kNtThreads, smem, st, g));
  } else {
    size_t smem = 196608 + 32768 + 1024;
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), node_fwd_tc_kernel<false>, nt_grid(c), kNtThreads, smem, st, g));
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_node_bwd_tc(Ctx* c, int l, cudaStream_t st) {
  NodeTcArgs g = nt_args(c, l);
  if (l < 2) {
    g.img = c->img_nba[l];
    size_t smem = 131072 + 32768 + 1024;
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_bwd_a_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), node_bwd_a_tc_kernel, nt_grid(c), kNtThreads, smem, st, g));
    GNNIS_LAUNCH_CHECK();
  }
  g.img = c->img_nbb[l];
  size_t smem = 131072 + 32768 + 1024;
  if (l == 2) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_bwd_b_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), node_bwd_b_tc_kernel<true>, nt_grid(c), kNtThreads, smem, st, g));
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_bwd_b_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNNIS_CUDA_OK(launch_k(c->pdl_now(), node_bwd_b_tc_kernel<false>, nt_grid(c), kNtThreads, smem, st, g));
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

}  // namespace gnnis
