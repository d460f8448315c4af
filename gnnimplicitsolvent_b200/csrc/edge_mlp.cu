// edge_mlp.cu -- per-edge message MLP of IN_layer_all_swish_2pass_tokens (GNN_Layers.py:2186-2208) and its
// adjoint, fp32 CUDA-core path (parity mode).
//
// Formulation.  message1 acts on [h_i | h_j | rbf]; it is split exactly as
//     u_e = P[tgt] + Q[src] + Wr . rbf(r_e),   P = W1[:, :cin] h + b1,   Q = W1[:, cin:2cin] h
// so the 2*cin gathered columns become two per-node tables (computed by the node kernels) and only the
// 20 radial columns remain per-edge work (SURVEY.md §7 hard part 2; changes fp32 summation order only).
//
// Work decomposition.  A *unit* = G consecutive replicas (<= kUnitAtomsMax atoms).  One persistent CTA owns a
// unit at a time: its P/Q tables (and, in the adjoint, abar / Pbar / Qbar) live in shared memory, its edges are
// the contiguous CSR range of its atoms and are processed in tiles of 128 edges.  Edges never leave the unit, so
// every reduction is CTA-local: the segmented add into the target rows walks the tile in CSR order, the
// source-side reduction of the adjoint walks a precomputed source-sorted permutation of the tile; both use
// exclusive ownership + a fixed-order carry fix-up (seg_reduce_tile) -- no atomics, bit-reproducible.
// Per-edge activations are never written to HBM; the adjoint recomputes u, v, w per tile.
#include "common.cuh"

namespace gnnis {

struct EdgeArgs {
  int n_units, G, N, R;
  const int* row_ptr;
  const uint32_t* pair;
  const float* r_e;
  float inv_cutoff;
  // weights
  const float *Wr_t, *W2_t, *b2, *Wr, *W2;
  // node tensors
  const float *P, *Q;   // [n][64]
  float* a_out;         // fwd: [n][64]
  const float* abar;    // bwd: [n][64]
  float *Pbar, *Qbar;   // bwd out: [n][64]
  float* rbar;          // bwd: dense [n][N]
  int rbar_accumulate;
  const uint8_t* perm;  // [E] tile-local edge index at source-sorted position (tile_perm_kernel)
  int unit_atoms;       // G*N: rows of the shared-memory tables
};

// C[128][64] += A[128][K] * B[K][64]; thread (te, tc) owns rows te+32i (i<4), cols tc*4..+3 and 32+tc*4..+3
template <int K, int LDA>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A, const float* __restrict__ B, int te, int tc,
                                          float (&acc)[4][8]) {
#pragma unroll 2
  for (int k4 = 0; k4 < K / 4; ++k4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (te + 32 * i) * LDA + k4 * 4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float* brow = B + (k4 * 4 + kk) * 64 + tc * 4;
      float4 b0 = *reinterpret_cast<const float4*>(brow);
      float4 b1 = *reinterpret_cast<const float4*>(brow + 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, b0.x, acc[i][0]);
        acc[i][1] = fmaf(av, b0.y, acc[i][1]);
        acc[i][2] = fmaf(av, b0.z, acc[i][2]);
        acc[i][3] = fmaf(av, b0.w, acc[i][3]);
        acc[i][4] = fmaf(av, b1.x, acc[i][4]);
        acc[i][5] = fmaf(av, b1.y, acc[i][5]);
        acc[i][6] = fmaf(av, b1.z, acc[i][6]);
        acc[i][7] = fmaf(av, b1.w, acc[i][7]);
      }
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
}

// radial basis of one edge (GNN_Layers.py:2194-2208): d = r/cutoff, env = 1/d - 28 d^5 + 48 d^6 - 21 d^7,
// rbf_k = env sin(k pi d); optionally d rbf_k / dr.  sin/cos(k pi d) come from one sincospi and the Chebyshev
// recurrence s_{k+1} = 2 cos(x) s_k - s_{k-1} (2 transcendentals + 40 FMA instead of 20-40 sinf calls; the
// recurrence error is damped by the envelope, which vanishes cubically at d -> 1).  Rows padded to kRbfLd.
template <bool DERIV>
__device__ __forceinline__ void rbf_row(float r, float inv_cutoff, float* __restrict__ out, float* __restrict__ dout) {
  const float d = r * inv_cutoff;
  const float d2 = d * d, d4 = d2 * d2, d5 = d4 * d;
  const float env = 1.0f / d - 28.0f * d5 + 48.0f * (d5 * d) - 21.0f * (d5 * d2);
  float denv = 0.0f;
  if (DERIV) denv = -1.0f / d2 - 140.0f * d4 + 288.0f * d5 - 147.0f * (d5 * d);
  float s1, c1;
  sincospif(d, &s1, &c1);
  const float twoc = 2.0f * c1;
  float sp = 0.0f, sc = s1, cp = 1.0f, cc = c1;
  float o[kRbfLd], od[kRbfLd];
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
  for (int k = kRbf; k < kRbfLd; ++k) {
    o[k] = 0.0f;
    od[k] = 0.0f;
  }
#pragma unroll
  for (int k4 = 0; k4 < kRbfLd / 4; ++k4) {
    *reinterpret_cast<float4*>(out + 4 * k4) = make_float4(o[4 * k4], o[4 * k4 + 1], o[4 * k4 + 2], o[4 * k4 + 3]);
    if (DERIV)
      *reinterpret_cast<float4*>(dout + 4 * k4) = make_float4(od[4 * k4], od[4 * k4 + 1], od[4 * k4 + 2], od[4 * k4 + 3]);
  }
}

// Segmented add of the rows of a [128][kTileLd] tile into table rows selected by key[]: rows with equal keys are
// contiguous in walk order (CSR order for targets, the precomputed source-sorted tile permutation for sources).
// Thread (c, q) walks positions [32q, 32q+32); a run that continues the previous quarter's key is parked in
// sCarry and added by a fixed-order fix-up pass, so no two threads ever update the same table entry
// concurrently: deterministic, no atomics.  Ends with the table consistent (trailing __syncthreads).
template <bool PERM>
__device__ __forceinline__ void seg_reduce_tile(const float* __restrict__ sX, const int* __restrict__ sKey,
                                                const uint8_t* __restrict__ sPerm, float* __restrict__ table,
                                                float* __restrict__ sCarry, int* __restrict__ sCarryKey, int tid) {
  const int c = tid & 63, q = tid >> 6, p0 = 32 * q;
  int cur = sKey[PERM ? (int)sPerm[p0] : p0];
  bool cont = false;
  if (q > 0 && cur >= 0) cont = sKey[PERM ? (int)sPerm[p0 - 1] : p0 - 1] == cur;
  if (c == 0) sCarryKey[q] = cont ? cur : -1;
  bool first = true;
  float run = 0.0f;
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    const int e = PERM ? (int)sPerm[p0 + j] : p0 + j;
    const int k = sKey[e];
    const float v = sX[e * kTileLd + c];
    if (k != cur) {
      if (cur >= 0) {
        if (first && cont) sCarry[q * 64 + c] = run;
        else table[cur * 64 + c] += run;
      }
      cur = k;
      run = 0.0f;
      first = false;
    }
    if (k >= 0) run += v;
  }
  if (cur >= 0) {
    if (first && cont) sCarry[q * 64 + c] = run;
    else table[cur * 64 + c] += run;
  }
  __syncthreads();
  if (tid < 64) {
#pragma unroll
    for (int qq = 1; qq < 4; ++qq) {
      int k = sCarryKey[qq];
      if (k >= 0) table[k * 64 + tid] += sCarry[qq * 64 + tid];
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void copy_rows(float* __restrict__ dst, const float* __restrict__ src, int nfloat) {
  for (int t = threadIdx.x * 4; t < nfloat; t += kEdgeThreads * 4)
    *reinterpret_cast<float4*>(dst + t) = *reinterpret_cast<const float4*>(src + t);
}
__device__ __forceinline__ void zero_rows(float* __restrict__ dst, int nfloat) {
  for (int t = threadIdx.x * 4; t < nfloat; t += kEdgeThreads * 4)
    *reinterpret_cast<float4*>(dst + t) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------ tile permutation
// For every 128-edge tile of every unit: positions sorted by (source, edge) -> tile-local edge index (uint8).
// Computed once per evaluation (after the CSR fill), reused by the three adjoint launches.
__global__ void __launch_bounds__(kTileE) tile_perm_kernel(int n_units, int G, int N, int R, int S, int U0,
                                                           const int* __restrict__ row_ptr,
                                                           const uint32_t* __restrict__ pair,
                                                           uint8_t* __restrict__ perm) {
  __shared__ int sKey[kTileE];
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const UnitSpan us = unit_span(unit, G, N, R, S, U0);
    const int e_begin = row_ptr[us.atom_base + us.a0], e_end = row_ptr[us.atom_base + us.a1];
    for (int t0 = e_begin; t0 < e_end; t0 += kTileE) {
      const int e = t0 + threadIdx.x;
      const int key = e < e_end ? (int)(pair[e] & 0xFFFFu) : 0x7FFFFFFF;
      sKey[threadIdx.x] = key;
      __syncthreads();
      int rank = 0;
#pragma unroll 8
      for (int j = 0; j < kTileE; ++j) {
        int kj = sKey[j];
        rank += (kj < key || (kj == key && j < (int)threadIdx.x)) ? 1 : 0;
      }
      if (t0 + rank < e_end) perm[t0 + rank] = (uint8_t)threadIdx.x;
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------ forward
// smem: Wr_t[20*64] W2_t[64*64] b2[64] | rbf[128*24] | V[128*68] | carry[4*64] | tgt[128] src[128] ckey[4] |
//       (P,Q,acc)[3][unit_atoms*64]
template <bool SMEM_TABLES>
__global__ void __launch_bounds__(kEdgeThreads, 2) edge_fwd_kernel(EdgeArgs g) {
  extern __shared__ __align__(16) float sm[];
  float* sWr = sm;
  float* sW2 = sWr + kRbf * 64;
  float* sb2 = sW2 + 64 * 64;
  float* sRbf = sb2 + 64;
  float* sV = sRbf + kTileE * kRbfLd;
  float* sCarry = sV + kTileE * kTileLd;
  int* sTgt = reinterpret_cast<int*>(sCarry + 4 * 64);
  int* sSrc = sTgt + kTileE;
  int* sCKey = sSrc + kTileE;
  float* sTab = reinterpret_cast<float*>(sCKey + 4);

  const int tid = threadIdx.x, tc = tid & 7, te = tid >> 3;
  for (int t = tid; t < kRbf * 64; t += kEdgeThreads) sWr[t] = g.Wr_t[t];
  for (int t = tid; t < 64 * 64; t += kEdgeThreads) sW2[t] = g.W2_t[t];
  if (tid < 64) sb2[tid] = g.b2[tid];
  __syncthreads();

  for (int unit = blockIdx.x; unit < g.n_units; unit += gridDim.x) {
    const int rep0 = unit * g.G, rep1 = min(g.R, rep0 + g.G);
    const int atom0 = rep0 * g.N, natoms = (rep1 - rep0) * g.N;
    const int e_begin = g.row_ptr[atom0], e_end = g.row_ptr[atom0 + natoms];
    float *tP, *tQ, *tA;
    if (SMEM_TABLES) {
      tP = sTab;
      tQ = tP + g.unit_atoms * 64;
      tA = tQ + g.unit_atoms * 64;
      copy_rows(tP, g.P + (size_t)atom0 * 64, natoms * 64);
      copy_rows(tQ, g.Q + (size_t)atom0 * 64, natoms * 64);
      zero_rows(tA, natoms * 64);
    } else {
      tP = const_cast<float*>(g.P) + (size_t)atom0 * 64;
      tQ = const_cast<float*>(g.Q) + (size_t)atom0 * 64;
      tA = g.a_out + (size_t)atom0 * 64;
      zero_rows(tA, natoms * 64);
    }
    __syncthreads();

    for (int t0 = e_begin; t0 < e_end; t0 += kTileE) {
      if (tid < kTileE) {
        int e = t0 + tid;
        bool valid = e < e_end;
        uint32_t pr = valid ? g.pair[e] : 0u;
        sTgt[tid] = valid ? (int)(pr >> 16) : -1;
        sSrc[tid] = (int)(pr & 0xFFFFu);
        rbf_row<false>(valid ? g.r_e[e] : 0.3f, g.inv_cutoff, sRbf + tid * kRbfLd, nullptr);
      }
      __syncthreads();
      float acc[4][8];
      zero_acc(acc);
      gemm_tile<kRbf, kRbfLd>(sRbf, sWr, te, tc, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int row = te + 32 * i;
        int tg = sTgt[row], sr = sSrc[row];
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (tg >= 0) {
          const float4 p0 = *reinterpret_cast<const float4*>(tP + tg * 64 + tc * 4);
          const float4 p1 = *reinterpret_cast<const float4*>(tP + tg * 64 + 32 + tc * 4);
          const float4 q0 = *reinterpret_cast<const float4*>(tQ + sr * 64 + tc * 4);
          const float4 q1 = *reinterpret_cast<const float4*>(tQ + sr * 64 + 32 + tc * 4);
          v0.x = silu_f(acc[i][0] + (p0.x + q0.x));
          v0.y = silu_f(acc[i][1] + (p0.y + q0.y));
          v0.z = silu_f(acc[i][2] + (p0.z + q0.z));
          v0.w = silu_f(acc[i][3] + (p0.w + q0.w));
          v1.x = silu_f(acc[i][4] + (p1.x + q1.x));
          v1.y = silu_f(acc[i][5] + (p1.y + q1.y));
          v1.z = silu_f(acc[i][6] + (p1.z + q1.z));
          v1.w = silu_f(acc[i][7] + (p1.w + q1.w));
        }
        *reinterpret_cast<float4*>(sV + row * kTileLd + tc * 4) = v0;
        *reinterpret_cast<float4*>(sV + row * kTileLd + 32 + tc * 4) = v1;
      }
      __syncthreads();
      zero_acc(acc);
      gemm_tile<64, kTileLd>(sV, sW2, te, tc, acc);
      __syncthreads();   // every thread finished reading V before it is overwritten with Z
      {
        const float4 b0 = *reinterpret_cast<const float4*>(sb2 + tc * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(sb2 + 32 + tc * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int row = te + 32 * i;
          float4 z0, z1;
          z0.x = silu_f(acc[i][0] + b0.x);
          z0.y = silu_f(acc[i][1] + b0.y);
          z0.z = silu_f(acc[i][2] + b0.z);
          z0.w = silu_f(acc[i][3] + b0.w);
          z1.x = silu_f(acc[i][4] + b1.x);
          z1.y = silu_f(acc[i][5] + b1.y);
          z1.z = silu_f(acc[i][6] + b1.z);
          z1.w = silu_f(acc[i][7] + b1.w);
          *reinterpret_cast<float4*>(sV + row * kTileLd + tc * 4) = z0;
          *reinterpret_cast<float4*>(sV + row * kTileLd + 32 + tc * 4) = z1;
        }
      }
      __syncthreads();
      seg_reduce_tile<false>(sV, sTgt, nullptr, tA, sCarry, sCKey, tid);   // a[tgt] += z
    }
    if (SMEM_TABLES) {
      copy_rows(g.a_out + (size_t)atom0 * 64, tA, natoms * 64);
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------ adjoint
// smem: Wr_t[20*64] W2_t[64*64] W2[64*64] WrP[64*24] b2[64] | rbf[128*24] drbf[128*24] | V[128*68] | carry[4*64] |
//       tgt[128] src[128] part[256] ckey[4] perm[128 B] | (P,Q,abar,Pbar,Qbar)[5][unit_atoms*64]
template <bool SMEM_TABLES>
__global__ void __launch_bounds__(kEdgeThreads, 1) edge_bwd_kernel(EdgeArgs g) {
  extern __shared__ __align__(16) float sm[];
  float* sWr = sm;
  float* sW2t = sWr + kRbf * 64;
  float* sW2 = sW2t + 64 * 64;
  float* sWrP = sW2 + 64 * 64;
  float* sb2 = sWrP + 64 * kRbfLd;
  float* sRbf = sb2 + 64;
  float* sDrbf = sRbf + kTileE * kRbfLd;
  float* sV = sDrbf + kTileE * kRbfLd;
  float* sCarry = sV + kTileE * kTileLd;
  int* sTgt = reinterpret_cast<int*>(sCarry + 4 * 64);
  int* sSrc = sTgt + kTileE;
  float* sPart = reinterpret_cast<float*>(sSrc + kTileE);
  int* sCKey = reinterpret_cast<int*>(sPart + kEdgeThreads);
  uint8_t* sPerm = reinterpret_cast<uint8_t*>(sCKey + 4);
  float* sTab = reinterpret_cast<float*>(sPerm + kTileE);

  const int tid = threadIdx.x, tc = tid & 7, te = tid >> 3;
  for (int t = tid; t < kRbf * 64; t += kEdgeThreads) sWr[t] = g.Wr_t[t];
  for (int t = tid; t < 64 * 64; t += kEdgeThreads) {
    sW2t[t] = g.W2_t[t];
    sW2[t] = g.W2[t];
  }
  for (int t = tid; t < 64 * kRbfLd; t += kEdgeThreads) {
    int cc = t / kRbfLd, k = t - cc * kRbfLd;
    sWrP[t] = k < kRbf ? g.Wr[cc * kRbf + k] : 0.0f;
  }
  if (tid < 64) sb2[tid] = g.b2[tid];
  __syncthreads();

  for (int unit = blockIdx.x; unit < g.n_units; unit += gridDim.x) {
    const int rep0 = unit * g.G, rep1 = min(g.R, rep0 + g.G);
    const int atom0 = rep0 * g.N, natoms = (rep1 - rep0) * g.N;
    const int e_begin = g.row_ptr[atom0], e_end = g.row_ptr[atom0 + natoms];
    float *tP, *tQ, *tAb, *tPb, *tQb;
    if (SMEM_TABLES) {
      tP = sTab;
      tQ = tP + g.unit_atoms * 64;
      tAb = tQ + g.unit_atoms * 64;
      tPb = tAb + g.unit_atoms * 64;
      tQb = tPb + g.unit_atoms * 64;
      copy_rows(tP, g.P + (size_t)atom0 * 64, natoms * 64);
      copy_rows(tQ, g.Q + (size_t)atom0 * 64, natoms * 64);
      copy_rows(tAb, g.abar + (size_t)atom0 * 64, natoms * 64);
    } else {
      tP = const_cast<float*>(g.P) + (size_t)atom0 * 64;
      tQ = const_cast<float*>(g.Q) + (size_t)atom0 * 64;
      tAb = const_cast<float*>(g.abar) + (size_t)atom0 * 64;
      tPb = g.Pbar + (size_t)atom0 * 64;
      tQb = g.Qbar + (size_t)atom0 * 64;
    }
    zero_rows(tPb, natoms * 64);
    zero_rows(tQb, natoms * 64);
    __syncthreads();

    for (int t0 = e_begin; t0 < e_end; t0 += kTileE) {
      if (tid < kTileE) {
        int e = t0 + tid;
        bool valid = e < e_end;
        uint32_t pr = valid ? g.pair[e] : 0u;
        sTgt[tid] = valid ? (int)(pr >> 16) : -1;
        sSrc[tid] = valid ? (int)(pr & 0xFFFFu) : -1;
        sPerm[tid] = valid ? g.perm[e] : (uint8_t)tid;   // invalid rows sit at the tail in both orders
        rbf_row<true>(valid ? g.r_e[e] : 0.3f, g.inv_cutoff, sRbf + tid * kRbfLd, sDrbf + tid * kRbfLd);
      }
      __syncthreads();
      // recompute u, v = silu(u), keep silu'(u)
      float acc[4][8], dsu[4][8];
      zero_acc(acc);
      gemm_tile<kRbf, kRbfLd>(sRbf, sWr, te, tc, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int row = te + 32 * i;
        int tg = sTgt[row], sr = sSrc[row];
        float v[8];
        if (tg >= 0) {
          const float4 p0 = *reinterpret_cast<const float4*>(tP + tg * 64 + tc * 4);
          const float4 p1 = *reinterpret_cast<const float4*>(tP + tg * 64 + 32 + tc * 4);
          const float4 q0 = *reinterpret_cast<const float4*>(tQ + sr * 64 + tc * 4);
          const float4 q1 = *reinterpret_cast<const float4*>(tQ + sr * 64 + 32 + tc * 4);
          silu_fd(acc[i][0] + (p0.x + q0.x), v[0], dsu[i][0]);
          silu_fd(acc[i][1] + (p0.y + q0.y), v[1], dsu[i][1]);
          silu_fd(acc[i][2] + (p0.z + q0.z), v[2], dsu[i][2]);
          silu_fd(acc[i][3] + (p0.w + q0.w), v[3], dsu[i][3]);
          silu_fd(acc[i][4] + (p1.x + q1.x), v[4], dsu[i][4]);
          silu_fd(acc[i][5] + (p1.y + q1.y), v[5], dsu[i][5]);
          silu_fd(acc[i][6] + (p1.z + q1.z), v[6], dsu[i][6]);
          silu_fd(acc[i][7] + (p1.w + q1.w), v[7], dsu[i][7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j] = 0.0f;
            dsu[i][j] = 0.0f;
          }
        }
        *reinterpret_cast<float4*>(sV + row * kTileLd + tc * 4) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(sV + row * kTileLd + 32 + tc * 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      __syncthreads();
      // w = W2 v + b2;  wbar = abar[tgt] * silu'(w)
      zero_acc(acc);
      gemm_tile<64, kTileLd>(sV, sW2t, te, tc, acc);
      __syncthreads();
      {
        const float4 b0 = *reinterpret_cast<const float4*>(sb2 + tc * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(sb2 + 32 + tc * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int row = te + 32 * i;
          int tg = sTgt[row];
          float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
          if (tg >= 0) {
            const float4 a0 = *reinterpret_cast<const float4*>(tAb + tg * 64 + tc * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(tAb + tg * 64 + 32 + tc * 4);
            w0.x = a0.x * silu_d(acc[i][0] + b0.x);
            w0.y = a0.y * silu_d(acc[i][1] + b0.y);
            w0.z = a0.z * silu_d(acc[i][2] + b0.z);
            w0.w = a0.w * silu_d(acc[i][3] + b0.w);
            w1.x = a1.x * silu_d(acc[i][4] + b1.x);
            w1.y = a1.y * silu_d(acc[i][5] + b1.y);
            w1.z = a1.z * silu_d(acc[i][6] + b1.z);
            w1.w = a1.w * silu_d(acc[i][7] + b1.w);
          }
          *reinterpret_cast<float4*>(sV + row * kTileLd + tc * 4) = w0;
          *reinterpret_cast<float4*>(sV + row * kTileLd + 32 + tc * 4) = w1;
        }
      }
      __syncthreads();
      // vbar = W2^T wbar;  ubar = vbar * silu'(u)
      zero_acc(acc);
      gemm_tile<64, kTileLd>(sV, sW2, te, tc, acc);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int row = te + 32 * i;
        *reinterpret_cast<float4*>(sV + row * kTileLd + tc * 4) =
            make_float4(acc[i][0] * dsu[i][0], acc[i][1] * dsu[i][1], acc[i][2] * dsu[i][2], acc[i][3] * dsu[i][3]);
        *reinterpret_cast<float4*>(sV + row * kTileLd + 32 + tc * 4) =
            make_float4(acc[i][4] * dsu[i][4], acc[i][5] * dsu[i][5], acc[i][6] * dsu[i][6], acc[i][7] * dsu[i][7]);
      }
      __syncthreads();
      // rbar_e = sum_k (Wr^T ubar)_k * d rbf_k/dr : thread (edge, half) handles 12 of the 24 padded k
      {
        const int e = tid & (kTileE - 1), half = tid >> 7;
        float ak[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) ak[k] = 0.0f;
        const float* urow = sV + e * kTileLd;
        const float* wbase = sWrP + half * 12;
#pragma unroll 4
        for (int c4 = 0; c4 < 16; ++c4) {
          float4 u4 = *reinterpret_cast<const float4*>(urow + c4 * 4);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            float uv = cc == 0 ? u4.x : cc == 1 ? u4.y : cc == 2 ? u4.z : u4.w;
            const float* wr = wbase + (c4 * 4 + cc) * kRbfLd;
            float4 w0 = *reinterpret_cast<const float4*>(wr);
            float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
            float4 w2 = *reinterpret_cast<const float4*>(wr + 8);
            ak[0] = fmaf(uv, w0.x, ak[0]);
            ak[1] = fmaf(uv, w0.y, ak[1]);
            ak[2] = fmaf(uv, w0.z, ak[2]);
            ak[3] = fmaf(uv, w0.w, ak[3]);
            ak[4] = fmaf(uv, w1.x, ak[4]);
            ak[5] = fmaf(uv, w1.y, ak[5]);
            ak[6] = fmaf(uv, w1.z, ak[6]);
            ak[7] = fmaf(uv, w1.w, ak[7]);
            ak[8] = fmaf(uv, w2.x, ak[8]);
            ak[9] = fmaf(uv, w2.y, ak[9]);
            ak[10] = fmaf(uv, w2.z, ak[10]);
            ak[11] = fmaf(uv, w2.w, ak[11]);
          }
        }
        const float* dr = sDrbf + e * kRbfLd + half * 12;
        float part = 0.0f;
#pragma unroll
        for (int k = 0; k < 12; ++k) part = fmaf(ak[k], dr[k], part);
        sPart[tid] = part;
      }
      __syncthreads();
      if (tid < kTileE) {
        int tg = sTgt[tid];
        if (tg >= 0) {
          int sr = sSrc[tid];
          int repl = tg / g.N;
          size_t slot = ((size_t)atom0 + tg) * g.N + (sr - repl * g.N);
          float v = sPart[tid] + sPart[tid + kTileE];
          g.rbar[slot] = g.rbar_accumulate ? g.rbar[slot] + v : v;
        }
      }
      seg_reduce_tile<false>(sV, sTgt, nullptr, tPb, sCarry, sCKey, tid);   // Pbar[tgt] += ubar
      seg_reduce_tile<true>(sV, sSrc, sPerm, tQb, sCarry, sCKey, tid);      // Qbar[src] += ubar
    }
    if (SMEM_TABLES) {
      copy_rows(g.Pbar + (size_t)atom0 * 64, tPb, natoms * 64);
      copy_rows(g.Qbar + (size_t)atom0 * 64, tQb, natoms * 64);
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------ launchers
static size_t edge_fwd_smem(const Ctx* c) {
  size_t f = kRbf * 64 + 64 * 64 + 64 + kTileE * kRbfLd + kTileE * kTileLd + 4 * 64 + 2 * kTileE + 4;
  if (c->smem_tables) f += 3 * (size_t)c->G * c->N * 64;
  return f * sizeof(float);
}
static size_t edge_bwd_smem(const Ctx* c) {
  size_t f = kRbf * 64 + 2 * 64 * 64 + 64 * kRbfLd + 64 + 2 * kTileE * kRbfLd + kTileE * kTileLd + 4 * 64 +
             2 * kTileE + kEdgeThreads + 4 + kTileE / 4;
  if (c->smem_tables) f += 5 * (size_t)c->G * c->N * 64;
  return f * sizeof(float);
}

static EdgeArgs make_args(Ctx* c, int l) {
  EdgeArgs g{};
  g.n_units = c->n_units;
  g.G = c->G;
  g.N = c->N;
  g.R = c->R;
  g.row_ptr = c->row_ptr;
  g.pair = c->pair;
  g.r_e = c->r_e;
  g.inv_cutoff = (float)(1.0 / (double)c->cutoff);
  g.Wr_t = c->L[l].Wr_t;
  g.W2_t = c->L[l].W2_t;
  g.b2 = c->L[l].b2;
  g.Wr = c->L[l].Wr;
  g.W2 = c->L[l].W2;
  g.P = c->P[l];
  g.Q = c->Q[l];
  g.a_out = c->a[l];
  g.abar = c->abar;
  g.Pbar = c->Pbar;
  g.Qbar = c->Qbar;
  g.rbar = c->rbar;
  g.perm = c->perm;
  g.unit_atoms = c->G * c->N;
  return g;
}

// Same permutation by a counting sort over the (<= 128) unit-local source indices.  One WARP per tile (no block
// barriers; the warps of a CTA share a unit and take its tiles round-robin): the warp walks its rows 32 at a time in
// order, lanes with equal sources are grouped (ballots), so the rank inside a bin follows the edge order; the bins live in
// the warp's own shared-memory slice.
__global__ void __launch_bounds__(kTileE) tile_perm_count_kernel(int n_units, int G, int N, int R, int S, int U0,
                                                                 const int* __restrict__ row_ptr,
                                                                 const uint32_t* __restrict__ pair,
                                                                 uint8_t* __restrict__ perm) {
  __shared__ int sbins[kTileE / 32][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* bins = sbins[warp];
  constexpr int kWarps = kTileE / 32;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const UnitSpan us = unit_span(unit, G, N, R, S, U0);
    const int e_begin = row_ptr[us.atom_base + us.a0], e_end = row_ptr[us.atom_base + us.a1];
    for (int t0 = e_begin + warp * kTileE; t0 < e_end; t0 += kWarps * kTileE) {
      int key[4], in_bin[4];   // in_bin: rows of this tile with the same source that come before this one
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        bins[32 * c + lane] = 0;
        const int e = t0 + 32 * c + lane;
        key[c] = e < e_end ? (int)(__ldg(pair + e) & 0xFFFFu) : -1;
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool valid = key[c] >= 0;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        // lanes with the same 7-bit key, from one ballot per key bit (MATCH.ANY costs ~130 cycles per call and SM here:
        // the kernel spent 90 us of a C2 evaluation waiting for it)
        unsigned peers = act;
#pragma unroll
        for (int b = 0; b < 7; ++b) {
          const unsigned m = __ballot_sync(0xffffffffu, (key[c] >> b) & 1);
          peers &= ((key[c] >> b) & 1) ? m : ~m;
        }
        in_bin[c] = 0;
        if (valid) {
          const int leader = __ffs(peers) - 1;
          int base = 0;
          if (lane == leader) {
            base = bins[key[c]];
            bins[key[c]] = base + __popc(peers);
          }
          base = __shfl_sync(peers, base, leader);
          in_bin[c] = base + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
      }
      // exclusive scan of the 128 bin counts (4 bins per lane)
      {
        const int c0 = bins[4 * lane], c1 = bins[4 * lane + 1], c2 = bins[4 * lane + 2], c3 = bins[4 * lane + 3];
        const int tot = c0 + c1 + c2 + c3;
        int inc = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, inc, off);
          if (lane >= off) inc += v;
        }
        const int ex = inc - tot;
        __syncwarp();
        bins[4 * lane] = ex;
        bins[4 * lane + 1] = ex + c0;
        bins[4 * lane + 2] = ex + c0 + c1;
        bins[4 * lane + 3] = ex + c0 + c1 + c2;
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (key[c] >= 0) perm[t0 + bins[key[c]] + in_bin[c]] = (uint8_t)(32 * c + lane);
      __syncwarp();
    }
  }
}

int launch_tile_perm(Ctx* c, cudaStream_t st) {
  const int nu = c->eff_units();
  if (c->G * c->N <= 128)
    tile_perm_count_kernel<<<(nu < 16 * c->num_sms ? nu : 16 * c->num_sms), kTileE, 0, st>>>(nu, c->G, c->N, c->R, c->split_eff, c->split_eff > 1 ? c->split_from : 0, c->row_ptr, c->pair, c->perm);
  else
    tile_perm_kernel<<<(nu < 8 * c->num_sms ? nu : 8 * c->num_sms), kTileE, 0, st>>>(nu, c->G, c->N, c->R, c->split_eff, c->split_eff > 1 ? c->split_from : 0, c->row_ptr, c->pair, c->perm);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_edge_fwd(Ctx* c, int l, cudaStream_t st) {
  EdgeArgs g = make_args(c, l);
  size_t smem = edge_fwd_smem(c);
  int per_sm = smem <= 113 * 1024 ? 2 : 1;
  int grid = c->n_units < per_sm * c->num_sms ? c->n_units : per_sm * c->num_sms;
  if (c->smem_tables) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_fwd_kernel<true><<<grid, kEdgeThreads, smem, st>>>(g);
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_fwd_kernel<false><<<grid, kEdgeThreads, smem, st>>>(g);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_edge_bwd(Ctx* c, int l, int rbar_accumulate, cudaStream_t st) {
  EdgeArgs g = make_args(c, l);
  g.rbar_accumulate = rbar_accumulate;
  int grid = c->n_units < c->num_sms ? c->n_units : c->num_sms;
  size_t smem = edge_bwd_smem(c);
  if (c->smem_tables) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_bwd_kernel<true><<<grid, kEdgeThreads, smem, st>>>(g);
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_bwd_kernel<false><<<grid, kEdgeThreads, smem, st>>>(g);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

}  // namespace gnnis
