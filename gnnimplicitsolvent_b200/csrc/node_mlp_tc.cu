// node_mlp_tc.cu -- the per-atom ("nodewise") MLPs of IN_layer_all_swish_2pass_tokens (GNN_Layers.py:2231-2246)
// and their adjoints on the tensor cores (tcgen05, 3xTF32, sm_100a).  Same arithmetic as node_mlp.cu (which stays
// as the CUDA-core cross-check); rows are atoms, 128 per tile, one persistent CTA per SM, 256 threads:
// thread = (row = TMEM lane, column half).  A operands are written to TMEM by the epilogue of the previous
// product; weights are staged once per CTA as K-major 128B-swizzled hi / lo operands.
//
//   forward  (layer l):  p = a W3a^T + sbias[s]      [128 x 64] x [64 x 128]
//                        o = silu(p) W4^T + b4       [128 x 128] x [128 x 64]   (-> global)
//                        [P|Q]_{l+1} = silu(o) [Wi;Wj]^T (+ b1)   [128 x 64] x [64 x 128]
//            last layer: o3 = silu(p) W4^T + b4 (2 outputs, CUDA cores) and the scale / SA head
//   adjoint A (l < 2):   hbar = [Pbar|Qbar] [Wi;Wj]  [128 x 128] x [128 x 64];  obar = hbar silu'(o)
//                        gbar = obar W4              [128 x 64] x [64 x 128]    (-> global scratch [n][128])
//   adjoint B:           p recomputed; pbar = gbar silu'(p);  abar = pbar W3a   [128 x 128] x [128 x 64]
#include "common.cuh"
#include "umma.cuh"

namespace gnnis {

using namespace umma;

int launch_build_operand(Ctx*, uint8_t*, uint8_t*, const float*, int, int, int, int, int, int, cudaStream_t);

constexpr int kNtThreads = 256;
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

__device__ __forceinline__ float nt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float nt_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float nt_sig(float x) { return nt_rcp(1.0f + nt_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float nt_silu(float x) { return x * nt_sig(x); }
__device__ __forceinline__ float nt_silu_d(float x) {
  const float sg = nt_sig(x);
  return fmaf(sg, fmaf(-x, sg, x), sg);
}

__device__ __forceinline__ void nt_copy_image(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int bytes) {
  for (int t = threadIdx.x * 16; t < bytes; t += blockDim.x * 16)
    *reinterpret_cast<uint4*>(dst + t) = __ldg(reinterpret_cast<const uint4*>(src + t));
}

// D[128 x N] = A[128 x 8 KSTEPS] * B^T (3xTF32), A in TMEM, B K-major SW128 with b_rows rows
template <int KSTEPS>
__device__ __forceinline__ void nt_issue(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                         int b_rows, uint32_t idesc) {
  uint32_t acc = 0;
#pragma unroll
  for (int term = 0; term < 3; ++term) {
    const uint32_t a = term == 2 ? a_lo : a_hi;
    const uint32_t b = term == 1 ? b_lo : b_hi;
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) {
      uint32_t baddr = b + (uint32_t)(s >> 2) * (uint32_t)b_rows * 128u + (uint32_t)(s & 3) * 32u;
      mma_tf32_ts(d_tmem, a + 8u * (uint32_t)s, smem_desc_k_sw128(baddr), idesc, acc);
      acc = 1;
    }
  }
}

__device__ __forceinline__ void nt_store_a16(uint32_t t_hi, uint32_t t_lo, const float (&x)[16]) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split_tf32(x[j], h[j], l[j]);
  tmem_st16(t_hi, h);
  tmem_st16(t_lo, l);
}

// 16 consecutive floats of a global row (zero when the row is past the end)
__device__ __forceinline__ void nt_load16(const float* __restrict__ p, bool ok, float (&x)[16]) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    float4 v = ok ? __ldg(reinterpret_cast<const float4*>(p) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[4 * j4 + 0] = v.x;
    x[4 * j4 + 1] = v.y;
    x[4 * j4 + 2] = v.z;
    x[4 * j4 + 3] = v.w;
  }
}
__device__ __forceinline__ void nt_store16(float* __restrict__ p, const float (&x)[16]) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4)
    *(reinterpret_cast<float4*>(p) + j4) = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
}

#define NT_PROLOGUE()                                                              \
  extern __shared__ uint8_t smraw[];                                               \
  __shared__ uint64_t bar;                                                         \
  __shared__ uint32_t tmem_slot;                                                   \
  uint32_t base_ = smem_u32(smraw);                                                \
  uint8_t* sm = smraw + ((1024u - (base_ & 1023u)) & 1023u);                       \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                   \
  const int wq = warp & 3, ch = warp >> 2, row = 32 * wq + lane;                   \
  if (warp == 0) tmem_alloc(&tmem_slot, 512);                                      \
  if (tid == 0) mbar_init(&bar, 1);

#define NT_STAGED()                                                                \
  fence_proxy_async_smem();                                                        \
  fence_before_sync();                                                             \
  __syncthreads();                                                                 \
  fence_after_sync();                                                              \
  const uint32_t tm = tmem_slot;                                                   \
  const uint32_t lane_addr = tm + ((uint32_t)(32 * wq) << 16);                     \
  uint32_t phase = 0;

// all threads: A operand stored -> one elected lane of warp 0 issues -> everybody waits for the product
#define NT_GEMM(ISSUE_STMT)                                                        \
  tmem_st_wait();                                                                  \
  fence_before_sync();                                                             \
  __syncthreads();                                                                 \
  if (warp == 0) {                                                                 \
    fence_after_sync();                                                            \
    if (elect_one()) {                                                             \
      ISSUE_STMT;                                                                  \
      mma_commit(&bar);                                                            \
    }                                                                              \
    __syncwarp();                                                                  \
  }                                                                                \
  mbar_wait(&bar, phase);                                                          \
  __syncwarp();                                                                    \
  phase ^= 1;                                                                      \
  fence_after_sync();

// true when every replica touched by the 128-atom tile is finished (uniform over the CTA)
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

// ------------------------------------------------------------------------------------------ forward
// smem: W3a (128 x 64) hi|lo 64 KB | W4 (64 x 128) hi|lo 64 KB | [Wi;Wj]_next (128 x 64) hi|lo 64 KB | head scratch
// TMEM: [0,128) a hi|lo -> o accumulator ; [128,256) p accumulator -> silu(o) hi|lo ; [256,512) silu(p) hi|lo -> [P|Q]
template <bool LAST>
__global__ void __launch_bounds__(kNtThreads, 1) node_fwd_tc_kernel(NodeTcArgs g) {
  NT_PROLOGUE();
  uint8_t *sW3H = sm, *sW3L = sm + 32768, *sW4H = sm + 65536, *sW4L = sm + 98304;
  uint8_t *sWpH = sm + 131072, *sWpL = sm + 163840;
  float* sHead = reinterpret_cast<float*>(sm + (LAST ? 65536 : 196608));   // [128][2] partial head sums
  nt_copy_image(sm, g.img, LAST ? 65536 : 196608);
  NT_STAGED();
  constexpr uint32_t idesc128 = idesc_tf32(128, 128), idesc64 = idesc_tf32(128, 64);
  const uint32_t aW3H = smem_u32(sW3H), aW3L = smem_u32(sW3L), aW4H = smem_u32(sW4H), aW4L = smem_u32(sW4L);
  const uint32_t aWpH = smem_u32(sWpH), aWpL = smem_u32(sWpL);
  const int ntiles = (g.n + kNtRows - 1) / kNtRows;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (nt_tile_skipped(g, tile)) continue;
    const int r = tile * kNtRows + row;
    const bool ok = r < g.n;
    const int rr = ok ? r : g.n - 1;
    // ---- a rows -> TMEM (this thread: 32 of the 64 columns)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = 32 * ch + 16 * c;
      float x[16];
      nt_load16(g.a + (size_t)rr * 64 + col0, ok, x);
      nt_store_a16(lane_addr + col0, lane_addr + 64 + col0, x);
    }
    NT_GEMM(nt_issue<8>(tm + 128, tm + 0, tm + 64, aW3H, aW3L, 128, idesc128));
    // ---- g = silu(p + sbias[s])  (this thread: 64 of the 128 columns)
    const float* sb = g.sbias + (size_t)g.solvent_id[rr / g.N] * 128 + 64 * ch;
    float h0 = 0.0f, h1 = 0.0f;   // last layer: partial dot products with the two rows of W4
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col0 = 64 * ch + 16 * c;
      uint32_t pr[16];
      tmem_ld16(lane_addr + 128 + col0, pr);
      float bs[16];
      nt_load16(sb + 16 * c, true, bs);
      tmem_ld_wait();
      float gv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) gv[j] = nt_silu(__uint_as_float(pr[j]) + bs[j]);
      if (LAST) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          h0 = fmaf(gv[j], __ldg(g.W4 + col0 + j), h0);
          h1 = fmaf(gv[j], __ldg(g.W4 + 128 + col0 + j), h1);
        }
      } else {
        nt_store_a16(lane_addr + 256 + col0, lane_addr + 384 + col0, gv);
      }
    }
    if (LAST) {
      // o3 = W4 g + b4 (2 outputs) and the head: B' = B (f + kappa sig(c)), SA = gamma_s sig(s) (rho + off + 0.14)^2
      if (ch == 1) {
        sHead[2 * row] = h0;
        sHead[2 * row + 1] = h1;
      }
      __syncthreads();
      if (ch == 0 && ok) {
        const float c0 = (h0 + sHead[2 * row]) + g.b4[0], c1 = (h1 + sHead[2 * row + 1]) + g.b4[1];
        g.o[2 * (size_t)r] = c0;
        g.o[2 * (size_t)r + 1] = c1;
        const int a = r % g.N;
        const float rp = g.rho[a] + kOffset + kProbe;
        g.sa_atom[r] = g.gamma_emb[g.solvent_id[r / g.N]] * (sigmoidf_(c1) * (rp * rp));
        g.born_s[r] = g.born[r] * (g.fraction + sigmoidf_(c0) * g.kappa);
      }
      fence_before_sync();
      __syncthreads();   // sHead and the TMEM regions are reused by the next tile
      continue;
    }
    NT_GEMM(nt_issue<16>(tm + 0, tm + 256, tm + 384, aW4H, aW4L, 64, idesc64));
    // ---- o = acc + b4 -> global ; h = silu(o) -> TMEM  (this thread: 32 of the 64 columns)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = 32 * ch + 16 * c;
      uint32_t orr[16];
      tmem_ld16(lane_addr + col0, orr);
      float bs[16];
      nt_load16(g.b4 + col0, true, bs);
      tmem_ld_wait();
      float ov[16], hv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        ov[j] = __uint_as_float(orr[j]) + bs[j];
        hv[j] = nt_silu(ov[j]);
      }
      if (ok) nt_store16(g.o + (size_t)r * 64 + col0, ov);
      nt_store_a16(lane_addr + 128 + col0, lane_addr + 192 + col0, hv);
    }
    NT_GEMM(nt_issue<8>(tm + 256, tm + 128, tm + 192, aWpH, aWpL, 128, idesc128));
    // ---- tables of the next layer: ch 0 -> P = . + b1, ch 1 -> Q
    {
      float* dst = (ch == 0 ? g.P_next : g.Q_next) + (size_t)rr * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t ar[16];
        tmem_ld16(lane_addr + 256 + 64 * ch + 16 * c, ar);
        float bs[16];
        nt_load16(g.b1_next + 16 * c, ch == 0, bs);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(ar[j]) + bs[j];
        if (ok) nt_store16(dst + 16 * c, v);
      }
    }
    fence_before_sync();
    __syncthreads();
  }
  NT_EPILOGUE();
}

// ------------------------------------------------------------------------------------------ adjoint A (layers 1, 2)
// smem: [Wi;Wj]^T as B[n = c_in (64)][k = (which, c_out) (128)] hi|lo 64 KB | W4^T as B[n = m (128)][k = c (64)] hi|lo 64 KB
// TMEM: [0,256) [Pbar|Qbar] hi (128) | lo (128) -> gbar accumulator [0,128) ; [256,320) hbar ; [320,448) obar hi|lo
__global__ void __launch_bounds__(kNtThreads, 1) node_bwd_a_tc_kernel(NodeTcArgs g) {
  NT_PROLOGUE();
  uint8_t *sWpH = sm, *sWpL = sm + 32768, *sW4H = sm + 65536, *sW4L = sm + 98304;
  nt_copy_image(sm, g.img, 131072);   // B[n = c_in][k = c_out] = Wi|Wj[c_out][c_in] ; B[n = m][k = c] = W4[c][m]
  NT_STAGED();
  constexpr uint32_t idesc128 = idesc_tf32(128, 128), idesc64 = idesc_tf32(128, 64);
  const uint32_t aWpH = smem_u32(sWpH), aWpL = smem_u32(sWpL), aW4H = smem_u32(sW4H), aW4L = smem_u32(sW4L);
  const int ntiles = (g.n + kNtRows - 1) / kNtRows;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (nt_tile_skipped(g, tile)) continue;
    const int r = tile * kNtRows + row;
    const bool ok = r < g.n;
    const int rr = ok ? r : g.n - 1;
    {
      const float* src = (ch == 0 ? g.Pbar : g.Qbar) + (size_t)rr * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float x[16];
        nt_load16(src + 16 * c, ok, x);
        nt_store_a16(lane_addr + 64 * ch + 16 * c, lane_addr + 128 + 64 * ch + 16 * c, x);
      }
    }
    NT_GEMM(nt_issue<16>(tm + 256, tm + 0, tm + 128, aWpH, aWpL, 64, idesc64));
    // ---- obar = hbar * silu'(o)  (this thread: 32 of the 64 columns)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = 32 * ch + 16 * c;
      uint32_t hr[16];
      tmem_ld16(lane_addr + 256 + col0, hr);
      float ov[16];
      nt_load16(g.o + (size_t)rr * 64 + col0, ok, ov);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(hr[j]) * nt_silu_d(ov[j]);
      nt_store_a16(lane_addr + 320 + col0, lane_addr + 384 + col0, v);
    }
    NT_GEMM(nt_issue<8>(tm + 0, tm + 320, tm + 384, aW4H, aW4L, 128, idesc128));
    // ---- gbar -> global scratch (this thread: 64 of the 128 columns)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col0 = 64 * ch + 16 * c;
      uint32_t gr[16];
      tmem_ld16(lane_addr + col0, gr);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(gr[j]);
      if (ok) nt_store16(g.gbar + (size_t)r * 128 + col0, v);
    }
    fence_before_sync();
    __syncthreads();
  }
  NT_EPILOGUE();
}

// ------------------------------------------------------------------------------------------ adjoint B
// smem: W3a (128 x 64) hi|lo 64 KB | W3a^T as B[n = k_in (64)][k = m (128)] hi|lo 64 KB | head scratch
// TMEM: [0,128) a hi|lo -> abar accumulator [0,64) ; [128,256) p accumulator ; [256,512) pbar hi (128) | lo (128)
template <bool LAST>
__global__ void __launch_bounds__(kNtThreads, 1) node_bwd_b_tc_kernel(NodeTcArgs g) {
  NT_PROLOGUE();
  uint8_t *sW3H = sm, *sW3L = sm + 32768, *sWtH = sm + 65536, *sWtL = sm + 98304;
  float* sHead = reinterpret_cast<float*>(sm + 131072);   // [128][2] (cbar, sbar)
  nt_copy_image(sm, g.img, 131072);   // B[n = m][k] = W3a[m][k] ; B[n = k_in][k = m] = W3a[m][k_in]
  NT_STAGED();
  constexpr uint32_t idesc128 = idesc_tf32(128, 128), idesc64 = idesc_tf32(128, 64);
  const uint32_t aW3H = smem_u32(sW3H), aW3L = smem_u32(sW3L), aWtH = smem_u32(sWtH), aWtL = smem_u32(sWtL);
  const int ntiles = (g.n + kNtRows - 1) / kNtRows;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (nt_tile_skipped(g, tile)) continue;
    const int r = tile * kNtRows + row;
    const bool ok = r < g.n;
    const int rr = ok ? r : g.n - 1;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = 32 * ch + 16 * c;
      float x[16];
      nt_load16(g.a + (size_t)rr * 64 + col0, ok, x);
      nt_store_a16(lane_addr + col0, lane_addr + 64 + col0, x);
    }
    if (LAST) {
      // head adjoint: cbar, sbar from dE/dB' and the SA term; Bbar += dE/dB' dB'/dB
      if (ch == 0) {
        float cb = 0.0f, sbv = 0.0f;
        if (ok) {
          const float c0 = g.o[2 * (size_t)r], c1 = g.o[2 * (size_t)r + 1];
          const float sc = sigmoidf_(c0), ss = sigmoidf_(c1);
          const float dEdBs = g.dE_dBs[r];
          cb = dEdBs * g.born[r] * g.kappa * sc * (1.0f - sc);
          const int a = r % g.N;
          const float rp = g.rho[a] + kOffset + kProbe;
          sbv = g.gamma_emb[g.solvent_id[r / g.N]] * (rp * rp) * ss * (1.0f - ss);
          g.Bbar[r] += dEdBs * (g.fraction + g.kappa * sc);
        }
        sHead[2 * row] = cb;
        sHead[2 * row + 1] = sbv;
      }
    }
    NT_GEMM(nt_issue<8>(tm + 128, tm + 0, tm + 64, aW3H, aW3L, 128, idesc128));
    // ---- pbar = gbar * silu'(p + sbias[s])  (this thread: 64 of the 128 columns)
    const float* sb = g.sbias + (size_t)g.solvent_id[rr / g.N] * 128 + 64 * ch;
    float cb = 0.0f, sbv = 0.0f;
    if (LAST) {
      cb = sHead[2 * row];
      sbv = sHead[2 * row + 1];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col0 = 64 * ch + 16 * c;
      uint32_t pr[16];
      tmem_ld16(lane_addr + 128 + col0, pr);
      float bs[16], gb[16];
      nt_load16(sb + 16 * c, true, bs);
      if (LAST) {
#pragma unroll
        for (int j = 0; j < 16; ++j) gb[j] = cb * __ldg(g.W4 + col0 + j) + sbv * __ldg(g.W4 + 128 + col0 + j);
      } else {
        nt_load16(g.gbar + (size_t)rr * 128 + col0, ok, gb);
      }
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = gb[j] * nt_silu_d(__uint_as_float(pr[j]) + bs[j]);
      nt_store_a16(lane_addr + 256 + col0, lane_addr + 384 + col0, v);
    }
    NT_GEMM(nt_issue<16>(tm + 0, tm + 256, tm + 384, aWtH, aWtL, 64, idesc64));
    // ---- abar -> global (this thread: 32 of the 64 columns)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = 32 * ch + 16 * c;
      uint32_t ar[16];
      tmem_ld16(lane_addr + col0, ar);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(ar[j]);
      if (ok) nt_store16(g.abar + (size_t)r * 64 + col0, v);
    }
    fence_before_sync();
    __syncthreads();
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

static int nt_grid(const Ctx* c) {
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
    size_t smem = 65536 + 1024 + 1024;
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_fwd_tc_kernel<true><<<nt_grid(c), kNtThreads, smem, st>>>(g);
  } else {
    size_t smem = 196608 + 1024 + 1024;
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_fwd_tc_kernel<false><<<nt_grid(c), kNtThreads, smem, st>>>(g);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_node_bwd_tc(Ctx* c, int l, cudaStream_t st) {
  NodeTcArgs g = nt_args(c, l);
  if (l < 2) {
    g.img = c->img_nba[l];
    size_t smem = 131072 + 1024;
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_bwd_a_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_bwd_a_tc_kernel<<<nt_grid(c), kNtThreads, smem, st>>>(g);
    GNNIS_LAUNCH_CHECK();
  }
  g.img = c->img_nbb[l];
  size_t smem = 131072 + 1024 + 1024;
  if (l == 2) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_bwd_b_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_bwd_b_tc_kernel<true><<<nt_grid(c), kNtThreads, smem, st>>>(g);
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_bwd_b_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_bwd_b_tc_kernel<false><<<nt_grid(c), kNtThreads, smem, st>>>(g);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

}  // namespace gnnis
