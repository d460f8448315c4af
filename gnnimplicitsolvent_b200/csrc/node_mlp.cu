// node_mlp.cu -- per-atom ("nodewise") parts of IN_layer_all_swish_2pass_tokens (GNN_Layers.py:2231-2246),
// the inter-layer SiLU (GNN_Models.py:824,831), the scale/SA head (GNN_Models.py:840-858) and their adjoints.
//
// Rows are atoms (64 per CTA tile, persistent grid); matrices are read reduction-major from global memory
// through L1 (they are shared by all CTAs and total < 130 KB per layer).  The node kernel of layer l also emits
// the P/Q tables of layer l+1 (see edge_mlp.cu), so the edge kernels never touch the 2*cin gathered columns.
// The solvent half of lin1 multiplies a per-replica constant: it is folded into a per-solvent bias table
// sbias[s] = W3[:, 64:] emb[s] + b3 when the weights are loaded (exact up to fp32 summation order).
#include "common.cuh"

namespace gnnis {

struct NodeArgs {
  int n, N;
  const int* solvent_id;   // [R]
  // forward
  const float* a;          // [n][64]
  float* o;                // [n][cout]
  float *P_next, *Q_next;  // [n][64]
  const float *W3a_t, *sbias, *W4_t, *b4;
  const float *Wi_t_next, *Wj_t_next, *b1_next;
  // head (last layer)
  const float *born, *rho, *gamma_emb;
  float fraction, kappa;   // kappa = (1 - fraction) * scaling
  float *born_s, *sa_atom;
  int gb_only;
  // adjoint
  const float *Pbar, *Qbar;     // [n][64] adjoints of the NEXT layer's tables
  const float *Wi_next, *Wj_next;  // [64][64] torch layout
  const float *W4, *W3a;
  const float* dE_dBs;
  float* Bbar;
  float* abar;
};

// acc[4][NOUT/16] += A_s[64][K] (smem, ld LDA) * B[K][NOUT] (global, reduction-major)
// thread (tr, tcn): rows tr+16i; cols tcn*4..+3 (and 64+tcn*4..+3 when NOUT == 128)
template <int K, int LDA, int NOUT>
__device__ __forceinline__ void node_gemm(const float* __restrict__ A, const float* __restrict__ B, int tr, int tcn,
                                          float (&acc)[4][NOUT / 16]) {
#pragma unroll 2
  for (int k4 = 0; k4 < K / 4; ++k4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (tr + 16 * i) * LDA + k4 * 4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float* brow = B + (size_t)(k4 * 4 + kk) * NOUT + tcn * 4;
      float4 b0 = __ldg(reinterpret_cast<const float4*>(brow));
      float4 b1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (NOUT == 128) b1 = __ldg(reinterpret_cast<const float4*>(brow + 64));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, b0.x, acc[i][0]);
        acc[i][1] = fmaf(av, b0.y, acc[i][1]);
        acc[i][2] = fmaf(av, b0.z, acc[i][2]);
        acc[i][3] = fmaf(av, b0.w, acc[i][3]);
        if (NOUT == 128) {
          acc[i][4] = fmaf(av, b1.x, acc[i][4]);
          acc[i][5] = fmaf(av, b1.y, acc[i][5]);
          acc[i][6] = fmaf(av, b1.z, acc[i][6]);
          acc[i][7] = fmaf(av, b1.w, acc[i][7]);
        }
      }
    }
  }
}

constexpr int kLd64 = 68, kLd128 = 132;

__device__ __forceinline__ void load_tile64(float* __restrict__ dst, const float* __restrict__ src, int row0, int n) {
  // [64][64] rows row0.. from global [n][64] into smem [64][kLd64]; rows >= n are zero
  for (int t = threadIdx.x; t < 64 * 16; t += kNodeThreads) {
    int r = t >> 4, c4 = t & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n) v = *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * 64 + c4 * 4);
    *reinterpret_cast<float4*>(dst + r * kLd64 + c4 * 4) = v;
  }
}

// p = W3a a + sbias[solvent]; returns g = silu(p) (and optionally silu'(p)) in registers for (rows tr+16i, 8 cols)
template <bool DERIV>
__device__ __forceinline__ void lin1_tile(const NodeArgs& g, const float* sA, int row0, int tr, int tcn,
                                          float (&gv)[4][8], float (&dv)[4][8]) {
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  node_gemm<64, kLd64, 128>(sA, g.W3a_t, tr, tcn, acc);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = row0 + tr + 16 * i;
    int rr = row < g.n ? row : g.n - 1;
    const float* sb = g.sbias + (size_t)g.solvent_id[rr / g.N] * 128;
    float4 s0 = __ldg(reinterpret_cast<const float4*>(sb + tcn * 4));
    float4 s1 = __ldg(reinterpret_cast<const float4*>(sb + 64 + tcn * 4));
    float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (DERIV) silu_fd(acc[i][j] + sv[j], gv[i][j], dv[i][j]);
      else gv[i][j] = silu_f(acc[i][j] + sv[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------ forward
// P1 = W1[:, 0:3] [B q rho] + b1, Q1 = W1[:, 3:6] [B q rho]   (layer-1 input, GNN_Models.py:817)
__global__ void node_pre1_kernel(int n, int N, const float* __restrict__ born, const float* __restrict__ q,
                                 const float* __restrict__ rho, const float* __restrict__ Wi_t,
                                 const float* __restrict__ Wj_t, const float* __restrict__ b1, float* __restrict__ P,
                                 float* __restrict__ Q) {
  pdl_trigger();   // the first edge forward kernel may set itself up while this one runs (common.cuh)
  // one thread per (atom, four channels): 16-byte stores, the per-channel arithmetic is unchanged
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * 16) return;
  const int i = (int)(t >> 4), c = (int)(t & 15) * 4, a = i % N;
  const float h0 = born[i], h1 = q[a], h2 = rho[a];
  const float4 wi0 = *reinterpret_cast<const float4*>(Wi_t + c), wi1 = *reinterpret_cast<const float4*>(Wi_t + 64 + c),
               wi2 = *reinterpret_cast<const float4*>(Wi_t + 128 + c), bb = *reinterpret_cast<const float4*>(b1 + c);
  const float4 wj0 = *reinterpret_cast<const float4*>(Wj_t + c), wj1 = *reinterpret_cast<const float4*>(Wj_t + 64 + c),
               wj2 = *reinterpret_cast<const float4*>(Wj_t + 128 + c);
  float4 p, qv;
  p.x = fmaf(h2, wi2.x, fmaf(h1, wi1.x, h0 * wi0.x)) + bb.x;
  p.y = fmaf(h2, wi2.y, fmaf(h1, wi1.y, h0 * wi0.y)) + bb.y;
  p.z = fmaf(h2, wi2.z, fmaf(h1, wi1.z, h0 * wi0.z)) + bb.z;
  p.w = fmaf(h2, wi2.w, fmaf(h1, wi1.w, h0 * wi0.w)) + bb.w;
  qv.x = fmaf(h2, wj2.x, fmaf(h1, wj1.x, h0 * wj0.x));
  qv.y = fmaf(h2, wj2.y, fmaf(h1, wj1.y, h0 * wj0.y));
  qv.z = fmaf(h2, wj2.z, fmaf(h1, wj1.z, h0 * wj0.z));
  qv.w = fmaf(h2, wj2.w, fmaf(h1, wj1.w, h0 * wj0.w));
  *reinterpret_cast<float4*>(P + (size_t)i * 64 + c) = p;
  *reinterpret_cast<float4*>(Q + (size_t)i * 64 + c) = qv;
}

// smem: sA[64][68] | sG[64][132]
template <bool LAST>
__global__ void __launch_bounds__(kNodeThreads) node_fwd_kernel(NodeArgs g) {
  extern __shared__ __align__(16) float sm[];
  float* sA = sm;
  float* sG = sA + 64 * kLd64;
  const int tid = threadIdx.x, tcn = tid & 15, tr = tid >> 4;
  const int ntiles = (g.n + 63) / 64;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * 64;
    load_tile64(sA, g.a, row0, g.n);
    __syncthreads();
    {
      float gv[4][8], dv[4][8];
      lin1_tile<false>(g, sA, row0, tr, tcn, gv, dv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float* d = sG + (tr + 16 * i) * kLd128;
        *reinterpret_cast<float4*>(d + tcn * 4) = make_float4(gv[i][0], gv[i][1], gv[i][2], gv[i][3]);
        *reinterpret_cast<float4*>(d + 64 + tcn * 4) = make_float4(gv[i][4], gv[i][5], gv[i][6], gv[i][7]);
      }
    }
    __syncthreads();
    if (!LAST) {
      // o = W4 g + b4 -> global; h = silu(o) -> sA
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
      node_gemm<128, kLd128, 64>(sG, g.W4_t, tr, tcn, acc);
      float4 b = __ldg(reinterpret_cast<const float4*>(g.b4 + tcn * 4));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = tr + 16 * i;
        float4 ov = make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
        if (row0 + r < g.n) *reinterpret_cast<float4*>(g.o + (size_t)(row0 + r) * 64 + tcn * 4) = ov;
        *reinterpret_cast<float4*>(sA + r * kLd64 + tcn * 4) =
            make_float4(silu_f(ov.x), silu_f(ov.y), silu_f(ov.z), silu_f(ov.w));
      }
      __syncthreads();
      // tables of the next layer
      float ap[4][4], aq[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ap[i][j] = aq[i][j] = 0.0f;
      node_gemm<64, kLd64, 64>(sA, g.Wi_t_next, tr, tcn, ap);
      node_gemm<64, kLd64, 64>(sA, g.Wj_t_next, tr, tcn, aq);
      float4 b1 = __ldg(reinterpret_cast<const float4*>(g.b1_next + tcn * 4));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = row0 + tr + 16 * i;
        if (r < g.n) {
          *reinterpret_cast<float4*>(g.P_next + (size_t)r * 64 + tcn * 4) =
              make_float4(ap[i][0] + b1.x, ap[i][1] + b1.y, ap[i][2] + b1.z, ap[i][3] + b1.w);
          *reinterpret_cast<float4*>(g.Q_next + (size_t)r * 64 + tcn * 4) =
              make_float4(aq[i][0], aq[i][1], aq[i][2], aq[i][3]);
        }
      }
    } else {
      // o3 = W4 g + b4 (2 outputs) and the head: B' = B (f + kappa sig(c)), SA = gamma_s sig(s) (rho+off+0.14)^2
      if (tid < 64) {
        int r = row0 + tid;
        if (r < g.n) {
          const float* gr = sG + tid * kLd128;
          float c0 = 0.0f, c1 = 0.0f;
          for (int m = 0; m < 128; ++m) {
            float gm = gr[m];
            c0 = fmaf(gm, __ldg(g.W4_t + 2 * m), c0);
            c1 = fmaf(gm, __ldg(g.W4_t + 2 * m + 1), c1);
          }
          c0 += g.b4[0];
          c1 += g.b4[1];
          g.o[2 * (size_t)r] = c0;
          g.o[2 * (size_t)r + 1] = c1;
          int a = r % g.N;
          float radius = g.rho[a] + kOffset;
          float rp = radius + kProbe;
          float sasa = sigmoidf_(c1) * (rp * rp);
          g.sa_atom[r] = g.gamma_emb[g.solvent_id[r / g.N]] * sasa;
          g.born_s[r] = g.born[r] * (g.fraction + sigmoidf_(c0) * g.kappa);
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ adjoint
// smem: sA[64][68] | sG[64][132] | sX[64][68] | sY[64][68]
template <bool LAST>
__global__ void __launch_bounds__(kNodeThreads) node_bwd_kernel(NodeArgs g) {
  extern __shared__ __align__(16) float sm[];
  float* sA = sm;
  float* sG = sA + 64 * kLd64;
  float* sX = sG + 64 * kLd128;
  float* sY = sX + 64 * kLd64;
  const int tid = threadIdx.x, tcn = tid & 15, tr = tid >> 4;
  const int ntiles = (g.n + 63) / 64;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * 64;
    load_tile64(sA, g.a, row0, g.n);
    float gbar[4][8];
    if (!LAST) {
      // hbar = Wi_next^T Pbar + Wj_next^T Qbar ; obar = hbar * silu'(o)
      load_tile64(sX, g.Pbar, row0, g.n);
      load_tile64(sY, g.Qbar, row0, g.n);
      __syncthreads();
      float hb[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) hb[i][j] = 0.0f;
      node_gemm<64, kLd64, 64>(sX, g.Wi_next, tr, tcn, hb);
      node_gemm<64, kLd64, 64>(sY, g.Wj_next, tr, tcn, hb);
      __syncthreads();   // done reading sX before it is reused for obar
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = tr + 16 * i;
        float4 ov = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < g.n) ov = *reinterpret_cast<const float4*>(g.o + (size_t)(row0 + r) * 64 + tcn * 4);
        *reinterpret_cast<float4*>(sX + r * kLd64 + tcn * 4) =
            make_float4(hb[i][0] * silu_d(ov.x), hb[i][1] * silu_d(ov.y), hb[i][2] * silu_d(ov.z), hb[i][3] * silu_d(ov.w));
      }
      __syncthreads();
      // gbar = W4^T obar
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) gbar[i][j] = 0.0f;
      node_gemm<64, kLd64, 128>(sX, g.W4, tr, tcn, gbar);
    } else {
      // head adjoint: cbar, sbar from dE/dB' and the SA term; Bbar += dE/dB' * dB'/dB
      if (tid < 64) {
        int r = row0 + tid;
        float cb = 0.0f, sb = 0.0f;
        if (r < g.n) {
          float c0 = g.o[2 * (size_t)r], c1 = g.o[2 * (size_t)r + 1];
          float sc = sigmoidf_(c0), ss = sigmoidf_(c1);
          float dEdBs = g.dE_dBs[r];
          cb = dEdBs * g.born[r] * g.kappa * sc * (1.0f - sc);
          int a = r % g.N;
          float rp = g.rho[a] + kOffset + kProbe;
          sb = g.gamma_emb[g.solvent_id[r / g.N]] * (rp * rp) * ss * (1.0f - ss);
          g.Bbar[r] += dEdBs * (g.fraction + g.kappa * sc);
        }
        sX[tid * kLd64] = cb;
        sX[tid * kLd64 + 1] = sb;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = tr + 16 * i;
        float cb = sX[r * kLd64], sb = sX[r * kLd64 + 1];
        float4 w0a = __ldg(reinterpret_cast<const float4*>(g.W4 + tcn * 4));
        float4 w0b = __ldg(reinterpret_cast<const float4*>(g.W4 + 64 + tcn * 4));
        float4 w1a = __ldg(reinterpret_cast<const float4*>(g.W4 + 128 + tcn * 4));
        float4 w1b = __ldg(reinterpret_cast<const float4*>(g.W4 + 192 + tcn * 4));
        gbar[i][0] = cb * w0a.x + sb * w1a.x;
        gbar[i][1] = cb * w0a.y + sb * w1a.y;
        gbar[i][2] = cb * w0a.z + sb * w1a.z;
        gbar[i][3] = cb * w0a.w + sb * w1a.w;
        gbar[i][4] = cb * w0b.x + sb * w1b.x;
        gbar[i][5] = cb * w0b.y + sb * w1b.y;
        gbar[i][6] = cb * w0b.z + sb * w1b.z;
        gbar[i][7] = cb * w0b.w + sb * w1b.w;
      }
    }
    // pbar = gbar * silu'(p), p recomputed from a
    {
      float gv[4][8], dv[4][8];
      lin1_tile<true>(g, sA, row0, tr, tcn, gv, dv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float* d = sG + (tr + 16 * i) * kLd128;
        *reinterpret_cast<float4*>(d + tcn * 4) =
            make_float4(gbar[i][0] * dv[i][0], gbar[i][1] * dv[i][1], gbar[i][2] * dv[i][2], gbar[i][3] * dv[i][3]);
        *reinterpret_cast<float4*>(d + 64 + tcn * 4) =
            make_float4(gbar[i][4] * dv[i][4], gbar[i][5] * dv[i][5], gbar[i][6] * dv[i][6], gbar[i][7] * dv[i][7]);
      }
    }
    __syncthreads();
    // abar = W3a^T pbar
    {
      float ab[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ab[i][j] = 0.0f;
      node_gemm<128, kLd128, 64>(sG, g.W3a, tr, tcn, ab);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = row0 + tr + 16 * i;
        if (r < g.n)
          *reinterpret_cast<float4*>(g.abar + (size_t)r * 64 + tcn * 4) = make_float4(ab[i][0], ab[i][1], ab[i][2], ab[i][3]);
      }
    }
    __syncthreads();
  }
}

// Bbar_i += (W1[:,0]^T Pbar1_i + W1[:,3]^T Qbar1_i): adjoint of the layer-1 input feature B (App. D.3 last line).
// One warp per atom, fixed shuffle tree.
__global__ void node_pre1_bwd_kernel(int n, const float* __restrict__ Pbar, const float* __restrict__ Qbar,
                                     const float* __restrict__ Wi_t, const float* __restrict__ Wj_t,
                                     float* __restrict__ Bbar) {
  int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* pb = Pbar + (size_t)warp * 64;
  const float* qb = Qbar + (size_t)warp * 64;
  float acc = pb[lane] * Wi_t[lane] + pb[lane + 32] * Wi_t[lane + 32] + qb[lane] * Wj_t[lane] + qb[lane + 32] * Wj_t[lane + 32];
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) Bbar[warp] += acc;
}

// GB-only mode (GNN_GBNeck_2): B' = B, no SA
__global__ void copy_born_kernel(int n, const float* __restrict__ born, float* __restrict__ born_s,
                                 float* __restrict__ sa_atom) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    born_s[i] = born[i];
    sa_atom[i] = 0.0f;
  }
}
__global__ void add_vec_kernel(int n, const float* __restrict__ x, float* __restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

// ------------------------------------------------------------------------------------------ launchers
static NodeArgs base_args(Ctx* c, int l) {
  NodeArgs g{};
  g.n = c->R * c->N;
  g.N = c->N;
  g.solvent_id = c->solvent_id;
  g.a = c->a[l];
  g.o = c->o[l];
  g.W3a_t = c->L[l].W3a_t;
  g.sbias = c->L[l].sbias;
  g.W4_t = c->L[l].W4_t;
  g.b4 = c->L[l].b4;
  g.W4 = c->L[l].W4;
  g.W3a = c->L[l].W3a;
  if (l < 2) {
    g.P_next = c->P[l + 1];
    g.Q_next = c->Q[l + 1];
    g.Wi_t_next = c->L[l + 1].Wi_t;
    g.Wj_t_next = c->L[l + 1].Wj_t;
    g.b1_next = c->L[l + 1].b1;
    g.Wi_next = c->L[l + 1].Wi;
    g.Wj_next = c->L[l + 1].Wj;
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
  g.Bbar = c->Bbar;
  g.abar = c->abar;
  return g;
}

int launch_node_pre1(Ctx* c, cudaStream_t st) {
  int n = c->R * c->N;
  size_t total = (size_t)n * 16;
  node_pre1_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(n, c->N, c->born, c->q, c->rho, c->L[0].Wi_t,
                                                                    c->L[0].Wj_t, c->L[0].b1, c->P[0], c->Q[0]);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_node_fwd(Ctx* c, int l, cudaStream_t st) {
  NodeArgs g = base_args(c, l);
  int ntiles = (g.n + 63) / 64;
  int grid = ntiles < 2 * c->num_sms ? ntiles : 2 * c->num_sms;
  size_t smem = (64 * kLd64 + 64 * kLd128) * sizeof(float);
  if (l == 2) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_fwd_kernel<true><<<grid, kNodeThreads, smem, st>>>(g);
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_fwd_kernel<false><<<grid, kNodeThreads, smem, st>>>(g);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_node_bwd(Ctx* c, int l, cudaStream_t st) {
  NodeArgs g = base_args(c, l);
  int ntiles = (g.n + 63) / 64;
  int grid = ntiles < 2 * c->num_sms ? ntiles : 2 * c->num_sms;
  size_t smem = (3 * 64 * kLd64 + 64 * kLd128) * sizeof(float);
  if (l == 2) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_bwd_kernel<true><<<grid, kNodeThreads, smem, st>>>(g);
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(node_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    node_bwd_kernel<false><<<grid, kNodeThreads, smem, st>>>(g);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_node_pre1_bwd(Ctx* c, cudaStream_t st) {
  int n = c->R * c->N;
  size_t threads = (size_t)n * 32;
  node_pre1_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(n, c->Pbar, c->Qbar, c->L[0].Wi_t,
                                                                         c->L[0].Wj_t, c->Bbar);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_gb_only_head(Ctx* c, cudaStream_t st) {
  int n = c->R * c->N;
  copy_born_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, c->born, c->born_s, c->sa_atom);
  GNNIS_LAUNCH_CHECK();
  return 0;
}
int launch_add_vec(Ctx* c, int n, const float* x, float* y, cudaStream_t st) {
  add_vec_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, x, y);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

}  // namespace gnnis
