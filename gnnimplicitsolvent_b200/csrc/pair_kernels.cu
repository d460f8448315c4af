// pair_kernels.cu -- all-pairs GB-Neck2 kernels and the sorted-CSR neighbour builder (sm_100a).
//
// One thread per TARGET atom, sources walked in ascending order from a shared-memory copy of the replica
// (positions + per-molecule parameters), so every reduction has the reference's CPU summation order
// (index_add over the [rep][node][con] edge layout, GNN_Models.py:919-944) and no float atomics exist.
//   born_count_kernel : Born integral I_i, Born radius B_i, dB/dI, GNN degree   (GNN_Layers.py:982-1058)
//   scan_blocks_kernel: exclusive scan of per-CTA degree sums                     (replaces nonzero())
//   fill_edges_kernel : CSR fill, sorted by target then source                    (GNN_Models.py:794-804)
//   gb_energy_kernel  : GB pair energy + its adjoint w.r.t. r and B'              (GNN_Layers.py:2055-2074, 2123-2140)
//   born_adjoint_kernel: chain rule through the Born integral + GNN-edge r-adjoints -> forces
//   reduce_energy_kernel: per-replica energies
#include "common.cuh"

namespace gnnis {

struct Stage {   // which replicas a CTA of kPairThreads consecutive atoms touches
  int first_rep, n_stage_atoms, base_atom;
};

__device__ __forceinline__ Stage cta_stage(int N, int n) {
  Stage st;
  int a0 = blockIdx.x * kPairThreads;
  int a1 = min(n, a0 + kPairThreads) - 1;
  st.first_rep = a0 / N;
  int last_rep = a1 / N;
  st.base_atom = st.first_rep * N;
  st.n_stage_atoms = (last_rep - st.first_rep + 1) * N;
  return st;
}

// Born integral of source (s_j, rho_j, radidx_j) on target (rho_i, radidx_i) at distance r  (GNN_Layers.py:1010-1058)
__device__ __forceinline__ float born_integral(float r, float rho_i, float rho_j, float s_j, int k,
                                               const float* __restrict__ d0, const float* __restrict__ m0) {
  float D = fabsf(r - s_j);
  float L = fmaxf(rho_i, D);
  float U = r + s_j;
  float ivdw = 0.0f;
  if ((r + s_j - rho_i) > 0.0f) {
    const float iL = rcp_fast(L), iU = rcp_fast(U), ir = rcp_fast(r);
    ivdw = 0.5f * (iL - iU + 0.25f * (r - s_j * s_j * ir) * (iU * iU - iL * iL) + 0.5f * ln_fast(L * iU) * ir);
  }
  float ineck = 0.0f;
  float radius1 = rho_i + kOffset, radius2 = rho_j + kOffset;
  if ((radius1 + radius2 + kNeckCut - r) > 0.0f) {
    float u = r - d0[k];
    float u2 = u * u;
    ineck = m0[k] * rcp_fast(1.0f + 100.0f * u2 + 0.3f * 1000000.0f * (u2 * u2 * u2));
  }
  return ivdw + kNeckScale * ineck;
}

// d(Born integral)/dr, SURVEY.md Appendix D.5
__device__ __forceinline__ float born_integral_dr(float r, float rho_i, float rho_j, float s_j, int k,
                                                  const float* __restrict__ d0, const float* __restrict__ m0) {
  float g = 0.0f;
  if ((r + s_j - rho_i) > 0.0f) {
    float D = fabsf(r - s_j);
    float L = fmaxf(rho_i, D);
    float U = r + s_j;
    float Lp = (D > rho_i) ? ((r - s_j) >= 0.0f ? 1.0f : -1.0f) : 0.0f;
    float iL = rcp_fast(L), iU = rcp_fast(U), ir = rcp_fast(r);
    float s2 = s_j * s_j;
    g = 0.5f * (-Lp * iL * iL + iU * iU + 0.25f * (1.0f + s2 * ir * ir) * (iU * iU - iL * iL) +
                0.25f * (r - s2 * ir) * (-2.0f * iU * iU * iU + 2.0f * Lp * iL * iL * iL) +
                0.5f * ((Lp * iL - iU) * ir - ln_fast(L * iU) * ir * ir));
  }
  float radius1 = rho_i + kOffset, radius2 = rho_j + kOffset;
  if ((radius1 + radius2 + kNeckCut - r) > 0.0f) {
    float u = r - d0[k];
    float u2 = u * u, u4 = u2 * u2;
    float den = 1.0f + 100.0f * u2 + 0.3f * 1000000.0f * (u4 * u2);
    const float iden = rcp_fast(den);
    g += kNeckScale * (-m0[k] * (200.0f * u + 1.8f * 1000000.0f * (u4 * u)) * (iden * iden));
  }
  return g;
}

// The all-pairs kernels come in two shapes (template LANES): one thread per target (LANES = 1: large batches, where
// they are bound by instruction issue) or four lanes per target that share the source loop and combine their partial
// sums with two shuffles in a fixed order (LANES = 4: small batches, where one thread's 50-iteration dependent chain
// is the latency of the whole launch).  A CTA covers kPairThreads consecutive atoms in both shapes.
template <int LANES>
__device__ __forceinline__ float lanes_sum(float x) {
#pragma unroll
  for (int o = 1; o < LANES; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
template <int LANES>
__device__ __forceinline__ int lanes_sum_int(int x) {
#pragma unroll
  for (int o = 1; o < LANES; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
static int pair_lanes(const Ctx* c) {   // few CTAs per SM: latency matters more than issue slots
  const int nblk = (c->R * c->N + kPairThreads - 1) / kPairThreads;
  return nblk < 4 * c->num_sms ? 4 : 1;
}

// dynamic smem: pos[3*nst] | rho[N] | s[N] | radidx[N] | d0[UU] | m0[UU]
template <int LANES>
__global__ void __launch_bounds__(kPairThreads * LANES)
born_count_kernel(const float* __restrict__ pos, int n, int N, int U, const float* __restrict__ rho,
                  const float* __restrict__ s, const int* __restrict__ radidx, const float* __restrict__ alpha,
                  const float* __restrict__ beta, const float* __restrict__ gamma, const float* __restrict__ d0,
                  const float* __restrict__ m0, float cutoff, int predicate, float* __restrict__ born,
                  float* __restrict__ dBdI, int* __restrict__ deg, int* __restrict__ blk_sum,
                  const int* __restrict__ skip) {
  extern __shared__ float sm[];
  Stage st = cta_stage(N, n);
  float* spos = sm;
  float* srho = spos + 3 * st.n_stage_atoms;
  float* ss = srho + N;
  int* sidx = reinterpret_cast<int*>(ss + N);
  float* sd0 = reinterpret_cast<float*>(sidx + N);
  float* sm0 = sd0 + U * U;
  __shared__ int s_blk;
  if (threadIdx.x == 0) s_blk = 0;
  for (int t = threadIdx.x; t < 3 * st.n_stage_atoms; t += blockDim.x) spos[t] = pos[(size_t)st.base_atom * 3 + t];
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    srho[t] = rho[t];
    ss[t] = s[t];
    sidx[t] = radidx[t];
  }
  for (int t = threadIdx.x; t < U * U; t += blockDim.x) {
    sd0[t] = d0[t];
    sm0[t] = m0[t];
  }
  __syncthreads();
  const int sub = threadIdx.x % LANES;
  int i = blockIdx.x * kPairThreads + threadIdx.x / LANES;
  const bool skipped = i < n && skip && skip[i / N] == 2;
  const bool act = i < n && !skipped;
  int mydeg = 0;
  float I = 0.0f;
  int rep = 0, a = 0;
  if (act) {
    rep = i / N;
    a = i - rep * N;
    const float* rp = spos + (size_t)(rep - st.first_rep) * N * 3;
    float tx = rp[3 * a], ty = rp[3 * a + 1], tz = rp[3 * a + 2];
    float rho_i = srho[a];
    int ri = sidx[a];
    for (int j = sub; j < N; j += LANES) {
      if (j == a) continue;
      float dx, dy, dz;
      float sx = rp[3 * j], sy = rp[3 * j + 1], sz = rp[3 * j + 2];
      float r = eps_dist(sx, sy, sz, tx, ty, tz, dx, dy, dz);
      mydeg += in_cutoff(predicate, r, sx, sy, sz, tx, ty, tz, cutoff) ? 1 : 0;
      I += born_integral(r, rho_i, srho[j], ss[j], U * sidx[j] + ri, sd0, sm0);
    }
  }
  I = lanes_sum<LANES>(I);
  mydeg = lanes_sum_int<LANES>(mydeg);
  if (sub != 0) mydeg = 0;           // one lane per target carries the degree into the block sum
  const bool count = predicate >= 0;   // negative: the cell-list back-end (cell_list.cu) produces deg / blk_sum
  if (skipped && sub == 0) {
    if (count) deg[i] = 0;   // finished replica: no edges, nothing downstream touches it
  } else if (act && sub == 0) {
    float rho_i = srho[a];
    float radius = rho_i + kOffset;
    float psi = I * rho_i;
    float al = alpha[a], be = beta[a], ga = gamma[a];
    float t = al * psi - be * (psi * psi) + ga * (psi * psi * psi);
    float T = tanhf(t);
    float B = 1.0f / (1.0f / rho_i - T / radius);
    born[i] = B;
    dBdI[i] = B * B * (1.0f - T * T) * (al - 2.0f * be * psi + 3.0f * ga * psi * psi) * rho_i / radius;
    if (count) deg[i] = mydeg;
  }
  if (!count) return;
  atomicAdd(&s_blk, mydeg);   // integer: order-independent
  __syncthreads();
  if (threadIdx.x == 0) blk_sum[blockIdx.x] = s_blk;
}

// single CTA: exclusive scan of blk_sum[nblk] in place; total -> *n_edges and row_ptr[n]
__global__ void __launch_bounds__(1024)
scan_blocks_kernel(int* __restrict__ blk_sum, int nblk, int64_t* __restrict__ n_edges, int* __restrict__ row_ptr_last) {
  __shared__ int part[1024];
  int per = (nblk + 1023) / 1024;
  int b0 = threadIdx.x * per, b1 = min(nblk, b0 + per);
  int sum = 0;
  for (int b = b0; b < b1; ++b) sum += blk_sum[b];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    int v = (threadIdx.x >= off) ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - sum;
  for (int b = b0; b < b1; ++b) {
    int v = blk_sum[b];
    blk_sum[b] = run;
    run += v;
  }
  if (threadIdx.x == 1023) {
    *n_edges = part[1023];
    *row_ptr_last = part[1023];
  }
}

// dynamic smem: pos[3*nst]
__global__ void __launch_bounds__(kPairThreads)
fill_edges_kernel(const float* __restrict__ pos, int n, int N, int G, float cutoff, int predicate,
                  const int* __restrict__ deg, const int* __restrict__ blk_off, int64_t capacity,
                  int* __restrict__ row_ptr, int* __restrict__ col, uint32_t* __restrict__ pair,
                  float* __restrict__ r_e) {
  extern __shared__ float sm[];
  __shared__ int sdeg[kPairThreads];
  Stage st = cta_stage(N, n);
  float* spos = sm;
  for (int t = threadIdx.x; t < 3 * st.n_stage_atoms; t += kPairThreads) spos[t] = pos[(size_t)st.base_atom * 3 + t];
  int i = blockIdx.x * kPairThreads + threadIdx.x;
  int mydeg = (i < n) ? deg[i] : 0;
  sdeg[threadIdx.x] = mydeg;
  __syncthreads();
  for (int off = 1; off < kPairThreads; off <<= 1) {
    int v = (threadIdx.x >= off) ? sdeg[threadIdx.x - off] : 0;
    __syncthreads();
    sdeg[threadIdx.x] += v;
    __syncthreads();
  }
  int64_t w = 0;
  int rep_i = 0, key_i = 0;   // replica of this thread's target; its atom | offset of the replica in its unit << 16
  if (i < n) {
    w = (int64_t)blk_off[blockIdx.x] + sdeg[threadIdx.x] - mydeg;
    row_ptr[i] = (int)w;
    rep_i = i / N;
    key_i = (i - rep_i * N) | (((rep_i % G) * N) << 16);
  }
  // The rows are filled by whole warps, one target at a time: the lanes take the sources j = lane, lane + 32, ...,
  // a ballot compacts the accepted ones, and the row is written with consecutive addresses (ascending source).
  // (replica / atom / unit offset are computed once per thread above and handed round by shuffles: the integer divisions
  // were a third of the instructions of a target.)
  const int lane = threadIdx.x & 31;
  const unsigned todo = __ballot_sync(0xffffffffu, i < n && mydeg > 0 && w + mydeg <= capacity);   // E > capacity: caller
  const int w32 = (int)w;   // capacity < 2^31 (ensure_workspace)
  for (unsigned m = todo; m; m &= m - 1) {
    const int t = __ffs(m) - 1;
    const int rep = __shfl_sync(0xffffffffu, rep_i, t), key = __shfl_sync(0xffffffffu, key_i, t);
    int64_t wt = __shfl_sync(0xffffffffu, w32, t);
    const int a = key & 0xFFFF;
    const float* rp = spos + (size_t)(rep - st.first_rep) * N * 3;
    const float tx = rp[3 * a], ty = rp[3 * a + 1], tz = rp[3 * a + 2];
    const uint32_t ul_base = (uint32_t)key >> 16;
    const uint32_t tgt_ul = (ul_base + (uint32_t)a) << 16;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      bool take = false;
      float r = 0.0f;
      if (j < N && j != a) {
        float dx, dy, dz;
        const float sx = rp[3 * j], sy = rp[3 * j + 1], sz = rp[3 * j + 2];
        r = eps_dist(sx, sy, sz, tx, ty, tz, dx, dy, dz);
        take = in_cutoff(predicate, r, sx, sy, sz, tx, ty, tz, cutoff);
      }
      const unsigned acc = __ballot_sync(0xffffffffu, take);
      if (take) {
        const int64_t at = wt + __popc(acc & ((1u << lane) - 1u));
        if (col) col[at] = rep * N + j;
        if (pair) pair[at] = tgt_ul | (ul_base + (uint32_t)j);
        if (r_e) r_e[at] = r;
      }
      wt += __popc(acc);
    }
  }
}

struct GbTerm {
  float e, dr_r, dP;   // 1/f, (d(1/f)/dr) / r, d(1/f)/dP
};
// pair term q_i q_j / f, f = sqrt(r^2 + P exp(-r^2/(4P)))  (GNN_Layers.py:2064-2074) and d(1/f)/dr, d(1/f)/dP
__device__ __forceinline__ GbTerm gb_term(float r, float P) {
  const float r2 = r * r;
  const float x = r2 * rcp_fast(4.0f * P);
  const float ex = exp_fast(-x);
  const float f2 = r2 + P * ex;
  GbTerm t;
  t.e = rsqrt_fast(f2);
  const float if3 = t.e * t.e * t.e;
  t.dr_r = -(1.0f - 0.25f * ex) * if3;
  t.dP = -0.5f * ex * (1.0f + x) * if3;
  return t;
}

// dynamic smem: pos[3*nst] | Bs[nst] | B[nst] | q[N]
// mode bit0: delta (subtract GB energy evaluated with the unscaled B)
template <int LANES>
__global__ void __launch_bounds__(kPairThreads * LANES)
gb_energy_kernel(const int* __restrict__ skip, const float* __restrict__ pos, int n, int N, const float* __restrict__ q,
                 const float* __restrict__ born_s, const float* __restrict__ born, const float* __restrict__ pref,
                 int delta, int want_grad, float* __restrict__ e_atom, float* __restrict__ dE_dBs,
                 float* __restrict__ dE_dB_direct, float* __restrict__ force) {
  extern __shared__ float sm[];
  Stage st = cta_stage(N, n);
  float* spos = sm;
  float* sBs = spos + 3 * st.n_stage_atoms;
  float* sB = sBs + st.n_stage_atoms;
  float* sq = sB + st.n_stage_atoms;
  for (int t = threadIdx.x; t < 3 * st.n_stage_atoms; t += blockDim.x) spos[t] = pos[(size_t)st.base_atom * 3 + t];
  for (int t = threadIdx.x; t < st.n_stage_atoms; t += blockDim.x) {
    sBs[t] = born_s[st.base_atom + t];
    sB[t] = delta ? born[st.base_atom + t] : 0.0f;
  }
  for (int t = threadIdx.x; t < N; t += blockDim.x) sq[t] = q[t];
  __syncthreads();
  const int sub = threadIdx.x % LANES;
  int i = blockIdx.x * kPairThreads + threadIdx.x / LANES;
  const bool act = i < n && !(skip && skip[i / N] == 2);   // uniform over the LANES lanes of a target
  if (LANES == 1 && !act) return;
  if (!act) i = 0;                                          // idle lane groups still take part in the shuffles
  int rep = i / N, a = i - rep * N;
  int lo = act ? (rep - st.first_rep) * N : 0;
  const float* rp = spos + (size_t)lo * 3;
  float tx = rp[3 * a], ty = rp[3 * a + 1], tz = rp[3 * a + 2];
  float qi = sq[a], Bi = sBs[lo + a], B0i = sB[lo + a];
  float pf = pref[rep];
  float pair_sum = 0.0f, pair_sum0 = 0.0f;
  float gx = 0.0f, gy = 0.0f, gz = 0.0f, dBs = 0.0f, dB0 = 0.0f;
  for (int j = sub; j < (act ? N : 0); j += LANES) {
    if (j == a) continue;
    float sx = rp[3 * j], sy = rp[3 * j + 1], sz = rp[3 * j + 2];
    float d1x, d1y, d1z;
    float r1 = eps_dist(sx, sy, sz, tx, ty, tz, d1x, d1y, d1z);   // edge (j -> i): i is the target
    float qq = qi * sq[j];
    float Bj = sBs[lo + j];
    GbTerm t1 = gb_term(r1, Bi * Bj);
    pair_sum += qq * t1.e;
    if (want_grad) {
      // edge (i -> j), i the source: its separation vector is kept exactly, but its pair term is evaluated at r1
      // (the two lengths differ by the 1e-6 eps only: relative force change ~1e-6, three orders below the gate)
      const float d2x = __fadd_rn(__fsub_rn(tx, sx), kEps), d2y = __fadd_rn(__fsub_rn(ty, sy), kEps),
                  d2z = __fadd_rn(__fsub_rn(tz, sz), kEps);
      float c = pf * qq;
      float g1 = c * t1.dr_r;
      gx += g1 * (d2x - d1x);
      gy += g1 * (d2y - d1y);
      gz += g1 * (d2z - d1z);
      dBs += 2.0f * c * t1.dP * Bj;
      if (delta) {
        float B0j = sB[lo + j];
        GbTerm u1 = gb_term(r1, B0i * B0j);
        pair_sum0 += qq * u1.e;
        float h1 = c * u1.dr_r;
        gx -= h1 * (d2x - d1x);
        gy -= h1 * (d2y - d1y);
        gz -= h1 * (d2z - d1z);
        dB0 -= 2.0f * c * u1.dP * B0j;
      }
    } else if (delta) {
      pair_sum0 += qq * gb_term(r1, B0i * sB[lo + j]).e;
    }
  }
  if (LANES > 1) {
    pair_sum = lanes_sum<LANES>(pair_sum);
    pair_sum0 = lanes_sum<LANES>(pair_sum0);
    gx = lanes_sum<LANES>(gx);
    gy = lanes_sum<LANES>(gy);
    gz = lanes_sum<LANES>(gz);
    dBs = lanes_sum<LANES>(dBs);
    dB0 = lanes_sum<LANES>(dB0);
    if (!act || sub != 0) return;
  }
  float e = pf * (pair_sum + qi * qi / Bi);
  if (delta) e -= pf * (pair_sum0 + qi * qi / B0i);
  e_atom[i] = e;
  if (want_grad) {
    dBs += -pf * qi * qi / (Bi * Bi);
    dE_dBs[i] = dBs;
    if (delta) dB0 -= -pf * qi * qi / (B0i * B0i);
    dE_dB_direct[i] = dB0;
    force[3 * (size_t)i] = -gx;
    force[3 * (size_t)i + 1] = -gy;
    force[3 * (size_t)i + 2] = -gz;
  }
}

// dynamic smem: pos[3*nst] | Ibar[nst] | rho[N] | s[N] | radidx[N] | d0[UU] | m0[UU]
// (12 CTAs of 128 threads per SM = 42 registers: 204 800 targets of the C2 batch are then resident in ONE wave of 227 k
// thread slots instead of 1.2 waves of 170 k at the 52 registers the compiler would take)
template <int LANES>
__global__ void __launch_bounds__(kPairThreads * LANES, LANES == 1 ? 12 : 1)
born_adjoint_kernel(const int* __restrict__ skip, const float* __restrict__ pos, int n, int N, int U, const float* __restrict__ rho,
                    const float* __restrict__ s, const int* __restrict__ radidx, const float* __restrict__ d0,
                    const float* __restrict__ m0, float cutoff, int predicate, const float* __restrict__ Bbar,
                    const float* __restrict__ dBdI, const float* __restrict__ rbar, int use_rbar,
                    float* __restrict__ force) {
  extern __shared__ float sm[];
  Stage st = cta_stage(N, n);
  float* spos = sm;
  float* sIbar = spos + 3 * st.n_stage_atoms;
  float* srho = sIbar + st.n_stage_atoms;
  float* ss = srho + N;
  int* sidx = reinterpret_cast<int*>(ss + N);
  float* sd0 = reinterpret_cast<float*>(sidx + N);
  float* sm0 = sd0 + U * U;
  for (int t = threadIdx.x; t < 3 * st.n_stage_atoms; t += blockDim.x) spos[t] = pos[(size_t)st.base_atom * 3 + t];
  for (int t = threadIdx.x; t < st.n_stage_atoms; t += blockDim.x)
    sIbar[t] = Bbar[st.base_atom + t] * dBdI[st.base_atom + t];
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    srho[t] = rho[t];
    ss[t] = s[t];
    sidx[t] = radidx[t];
  }
  for (int t = threadIdx.x; t < U * U; t += blockDim.x) {
    sd0[t] = d0[t];
    sm0[t] = m0[t];
  }
  __syncthreads();
  const int sub = threadIdx.x % LANES;
  int i = blockIdx.x * kPairThreads + threadIdx.x / LANES;
  const bool act = i < n && !(skip && skip[i / N] == 2);
  if (LANES == 1 && !act) return;
  if (!act) i = 0;
  int rep = i / N, a = i - rep * N;
  int lo = act ? (rep - st.first_rep) * N : 0;
  const float* rp = spos + (size_t)lo * 3;
  float tx = rp[3 * a], ty = rp[3 * a + 1], tz = rp[3 * a + 2];
  float rho_i = srho[a], s_i = ss[a], Ibar_i = sIbar[lo + a];
  int ri = sidx[a];
  const float* rbar_row = rbar + (size_t)i * N;                 // slots [i][j]: edges (j -> i)
  const float* rbar_col = rbar + (size_t)rep * N * N + a;       // slots [j][i]: edges (i -> j), stride N
  float gx = 0.0f, gy = 0.0f, gz = 0.0f;
  for (int j = sub; j < (act ? N : 0); j += LANES) {
    if (j == a) continue;
    float sx = rp[3 * j], sy = rp[3 * j + 1], sz = rp[3 * j + 2];
    float d1x, d1y, d1z, d2x, d2y, d2z;
    float r1 = eps_dist(sx, sy, sz, tx, ty, tz, d1x, d1y, d1z);   // (j -> i)
    float r2 = eps_dist(tx, ty, tz, sx, sy, sz, d2x, d2y, d2z);   // (i -> j)
    float rho_j = srho[j];
    int rj = sidx[j];
    float g1 = Ibar_i * born_integral_dr(r1, rho_i, rho_j, ss[j], U * rj + ri, sd0, sm0);
    float g2 = sIbar[lo + j] * born_integral_dr(r2, rho_j, rho_i, s_i, U * ri + rj, sd0, sm0);
    if (use_rbar) {
      // the same edge test as fill_edges_kernel (run path: eps-distance, data path: radius_graph's squared distance):
      // only slots that an edge adjoint kernel has written in this evaluation are read
      if (in_cutoff(predicate, r1, sx, sy, sz, tx, ty, tz, cutoff)) g1 += rbar_row[j];
      if (in_cutoff(predicate, r2, tx, ty, tz, sx, sy, sz, cutoff)) g2 += rbar_col[(size_t)j * N];
    }
    g1 *= rcp_fast(r1);
    g2 *= rcp_fast(r2);
    gx += -g1 * d1x + g2 * d2x;
    gy += -g1 * d1y + g2 * d2y;
    gz += -g1 * d1z + g2 * d2z;
  }
  if (LANES > 1) {
    gx = lanes_sum<LANES>(gx);
    gy = lanes_sum<LANES>(gy);
    gz = lanes_sum<LANES>(gz);
    if (!act || sub != 0) return;
  }
  force[3 * (size_t)i] -= gx;
  force[3 * (size_t)i + 1] -= gy;
  force[3 * (size_t)i + 2] -= gz;
}

// one warp per replica: E_rep = sum_i (e_i + sa_i), fixed order (lane-strided partials + shuffle tree)
__global__ void reduce_energy_kernel(const float* __restrict__ e_atom, const float* __restrict__ sa_atom, int R, int N,
                                     float* __restrict__ e_rep) {
  pdl_trigger();   // the first node adjoint kernel may set itself up while this one runs (common.cuh)
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  float acc = 0.0f;
  for (int a = lane; a < N; a += 32) {
    size_t i = (size_t)warp * N + a;
    acc += e_atom[i] + (sa_atom ? sa_atom[i] : 0.0f);
  }
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) e_rep[warp] = acc;
}

// ---------------------------------------------------------------- host launchers
static size_t stage_atoms_max(int N) { return (size_t)kPairThreads + 2 * (size_t)N; }

int launch_scan_blocks(Ctx* c, cudaStream_t st) {
  const int n = c->R * c->N;
  const int nblk = (n + kPairThreads - 1) / kPairThreads;
  scan_blocks_kernel<<<1, 1024, 0, st>>>(c->blk_sum, nblk, c->n_edges_dev, c->row_ptr + n);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_born_count(Ctx* c, const float* pos, int predicate, cudaStream_t st) {
  int n = c->R * c->N;
  int nblk = (n + kPairThreads - 1) / kPairThreads;
  size_t smem = (3 * stage_atoms_max(c->N) + 3 * (size_t)c->N + 2 * (size_t)c->U * c->U) * sizeof(float);
  if (pair_lanes(c) == 4) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(born_count_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    born_count_kernel<4><<<nblk, 4 * kPairThreads, smem, st>>>(pos, n, c->N, c->U, c->rho, c->s, c->radidx, c->alpha, c->beta,
                                                              c->gamma, c->d0, c->m0, c->cutoff, predicate, c->born,
                                                              c->dBdI, c->deg, c->blk_sum, c->skip_state);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(born_count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    born_count_kernel<1><<<nblk, kPairThreads, smem, st>>>(pos, n, c->N, c->U, c->rho, c->s, c->radidx, c->alpha, c->beta,
                                                          c->gamma, c->d0, c->m0, c->cutoff, predicate, c->born, c->dBdI,
                                                          c->deg, c->blk_sum, c->skip_state);
  }
  GNNIS_LAUNCH_CHECK();
  if (predicate < 0) return 0;   // degrees, block sums and their scan come from the cell-list back-end
  return launch_scan_blocks(c, st);
}

int launch_fill_edges(Ctx* c, const float* pos, int predicate, int* row_ptr, int* col, uint32_t* pair, float* r_e,
                      int64_t capacity, cudaStream_t st) {
  int n = c->R * c->N;
  int nblk = (n + kPairThreads - 1) / kPairThreads;
  size_t smem = 3 * stage_atoms_max(c->N) * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(fill_edges_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fill_edges_kernel<<<nblk, kPairThreads, smem, st>>>(pos, n, c->N, c->G, c->cutoff, predicate, c->deg, c->blk_sum,
                                                     capacity, row_ptr, col, pair, r_e);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_gb_energy(Ctx* c, const float* pos, const float* born_s, int delta, int want_grad, float* force,
                     cudaStream_t st) {
  int n = c->R * c->N;
  int nblk = (n + kPairThreads - 1) / kPairThreads;
  size_t smem = (5 * stage_atoms_max(c->N) + (size_t)c->N) * sizeof(float);
  if (pair_lanes(c) == 4) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(gb_energy_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gb_energy_kernel<4><<<nblk, 4 * kPairThreads, smem, st>>>(c->skip_state, pos, n, c->N, c->q, born_s, c->born, c->pref, delta,
                                                             want_grad, c->e_atom, c->dE_dBs, c->Bbar, force);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(gb_energy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gb_energy_kernel<1><<<nblk, kPairThreads, smem, st>>>(c->skip_state, pos, n, c->N, c->q, born_s, c->born, c->pref, delta,
                                                         want_grad, c->e_atom, c->dE_dBs, c->Bbar, force);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_born_adjoint(Ctx* c, const float* pos, int use_rbar, int predicate, float* force, cudaStream_t st) {
  int n = c->R * c->N;
  int nblk = (n + kPairThreads - 1) / kPairThreads;
  size_t smem = (4 * stage_atoms_max(c->N) + 3 * (size_t)c->N + 2 * (size_t)c->U * c->U) * sizeof(float);
  if (pair_lanes(c) == 4) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(born_adjoint_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    born_adjoint_kernel<4><<<nblk, 4 * kPairThreads, smem, st>>>(c->skip_state, pos, n, c->N, c->U, c->rho, c->s, c->radidx, c->d0,
                                                                c->m0, c->cutoff, predicate, c->Bbar, c->dBdI, c->rbar, use_rbar, force);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(born_adjoint_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    born_adjoint_kernel<1><<<nblk, kPairThreads, smem, st>>>(c->skip_state, pos, n, c->N, c->U, c->rho, c->s, c->radidx, c->d0,
                                                            c->m0, c->cutoff, predicate, c->Bbar, c->dBdI, c->rbar, use_rbar, force);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_reduce_energy(Ctx* c, const float* sa_atom, float* e_rep, cudaStream_t st) {
  int threads = 256;
  int blocks = (c->R * 32 + threads - 1) / threads;
  reduce_energy_kernel<<<blocks, threads, 0, st>>>(c->e_atom, sa_atom, c->R, c->N, e_rep);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

}  // namespace gnnis
