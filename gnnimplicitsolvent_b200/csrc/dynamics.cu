// dynamics.cu -- batched per-replica L-BFGS minimiser, LangevinMiddle integrator and optional harmonic bonds.
// All replicas stay resident in HBM; every round is one batched energy/force evaluation plus one small kernel.
//
// Replaces (semantics differ by design, SURVEY.md §7 hard part 5 / Appendix E):
//   simulation.minimizeEnergy(tolerance=1e-4, maxIterations=0)    Simulation/helper_functions.py:614-624
//       OpenMM runs ONE L-BFGS over all 3*R*N coordinates; here every replica has its own history, line search
//       and convergence flag, so one hard conformer cannot stall or derail the others.
//   LangevinMiddleIntegrator(300 K, 1/ps, 2 fs) + simulation.step  helper_functions.py:787-789, Simulator.py:402
#include <vector>

#include <vector>

#include "common.cuh"

namespace gnnis {

int energy_force_impl(Ctx*, const float*, float*, float*, uint32_t, cudaStream_t, struct StageTimer*);

// ---------------------------------------------------------------- harmonic bonds (stand-in vacuum term)
__global__ void bond_force_kernel(int n, int N, int n_bonds, const int* __restrict__ idx, const float* __restrict__ r0,
                                  const float* __restrict__ k, const float* __restrict__ pos, float* __restrict__ force) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int rep = i / N, a = i - rep * N;
  const float* rp = pos + (size_t)rep * N * 3;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  for (int b = 0; b < n_bonds; ++b) {
    int p = idx[2 * b], q = idx[2 * b + 1];
    if (p != a && q != a) continue;
    int o = p == a ? q : p;
    float dx = rp[3 * a] - rp[3 * o], dy = rp[3 * a + 1] - rp[3 * o + 1], dz = rp[3 * a + 2] - rp[3 * o + 2];
    float r = sqrtf(dx * dx + dy * dy + dz * dz);
    float g = -k[b] * (r - r0[b]) / r;
    fx += g * dx;
    fy += g * dy;
    fz += g * dz;
  }
  force[3 * (size_t)i] += fx;
  force[3 * (size_t)i + 1] += fy;
  force[3 * (size_t)i + 2] += fz;
}
__global__ void bond_energy_kernel(int R, int N, int n_bonds, const int* __restrict__ idx, const float* __restrict__ r0,
                                   const float* __restrict__ k, const float* __restrict__ pos, float* __restrict__ e_rep) {
  int rep = blockIdx.x * blockDim.x + threadIdx.x;
  if (rep >= R) return;
  const float* rp = pos + (size_t)rep * N * 3;
  float e = 0.f;
  for (int b = 0; b < n_bonds; ++b) {
    int p = idx[2 * b], q = idx[2 * b + 1];
    float dx = rp[3 * p] - rp[3 * q], dy = rp[3 * p + 1] - rp[3 * q + 1], dz = rp[3 * p + 2] - rp[3 * q + 2];
    float d = sqrtf(dx * dx + dy * dy + dz * dz) - r0[b];
    e += 0.5f * k[b] * d * d;
  }
  e_rep[rep] += e;
}

int launch_bond_forces(Ctx* c, const float* pos, float* e_rep, float* force, cudaStream_t st) {
  int n = c->R * c->N;
  if (force) {
    bond_force_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, c->N, c->n_bonds, c->bond_idx, c->bond_r0, c->bond_k, pos, force);
    GNNIS_LAUNCH_CHECK();
  }
  bond_energy_kernel<<<(c->R + 127) / 128, 128, 0, st>>>(c->R, c->N, c->n_bonds, c->bond_idx, c->bond_r0, c->bond_k, pos, e_rep);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- L-BFGS
constexpr int kLbThreads = 128;
enum { LB_INIT = 0, LB_SEARCH = 1, LB_DONE = 2 };

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < kLbThreads / 32; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  for (int off = 16; off; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < kLbThreads / 32; ++i) t = fmaxf(t, red[i]);
  return t;
}

struct LbArgs {
  int D, m;
  float* x;        // [R][D] accepted positions (the caller's buffer)
  float* xt;       // [R][D] trial positions
  const float* ft; // [R][D] forces at the trial positions
  const float* et; // [R]    energies at the trial positions
  float *g, *d, *S, *Y, *rho, *alpha, *e, *step, *gd, *eref;
  int *state, *hist, *status, *n_active, *stall;
  int* n_evals;     // [R] energy / force evaluations this replica took part in (gnnis_minimize_stats)
  float* grms_out;  // [R] gradient RMS at the last accepted point
  float grad_tol, max_disp;
  int stall_limit;
  int sd;   // steepest descent: no curvature pairs are kept, the first trial step follows the last accepted displacement
};

// one CTA per replica: accept / reject the trial point, update the history, build the next direction and trial.
__global__ void __launch_bounds__(kLbThreads) lbfgs_update_kernel(LbArgs a) {
  __shared__ float red[kLbThreads / 32];
  const int r = blockIdx.x, D = a.D, m = a.m;
  const size_t o = (size_t)r * D;
  int state = a.state[r];
  if (state == LB_DONE) return;
  if (threadIdx.x == 0) a.n_evals[r] += 1;
  float* x = a.x + o;
  float* xt = a.xt + o;
  const float* ft = a.ft + o;
  float* g = a.g + o;
  float* d = a.d + o;
  float* S = a.S + (size_t)r * m * D;
  float* Y = a.Y + (size_t)r * m * D;
  float* rho = a.rho + (size_t)r * m;
  float* alpha = a.alpha + (size_t)r * m;
  int hist = a.hist[r];            // low 16 bits: count, high 16 bits: head (next slot to write)
  int cnt = hist & 0xFFFF, head = hist >> 16;
  float et = a.et[r];
  float step = a.step[r];

  // finite check of the trial point
  float bad = 0.f;
  for (int t = threadIdx.x; t < D; t += kLbThreads) bad += isfinite(ft[t]) ? 0.f : 1.f;
  bad = block_sum(bad, red);
  bool finite = isfinite(et) && bad == 0.f;

  bool accept;
  if (state == LB_INIT) {
    if (!finite) {
      if (threadIdx.x == 0) {
        a.e[r] = et;               // reported as it is (NaN / inf): status 2 says the start point was not finite
        a.state[r] = LB_DONE;
        a.status[r] = 2;
        atomicSub(a.n_active, 1);
      }
      return;
    }
    accept = true;
  } else {
    // Armijo decrease, or -- when the energy difference is below the fp32 resolution of the summed energy -- the
    // approximate Wolfe conditions of Hager & Zhang: energy not higher than the noise allows and the directional
    // derivative reduced, sigma phi'(0) <= phi'(step) <= (2 delta - 1) phi'(0)
    float gtd = 0.f;
    for (int t = threadIdx.x; t < D; t += kLbThreads) gtd += -ft[t] * d[t];
    gtd = block_sum(gtd, red);
    const float e0 = a.e[r], gd0 = a.gd[r];
    const float noise = 2e-6f * fabsf(e0) + 1e-3f;
    const bool armijo = et <= e0 + 1e-4f * step * gd0;
    const bool approx_wolfe = et <= e0 + noise && gtd >= 0.9f * gd0 && gtd <= -0.9998f * gd0;
    accept = finite && (armijo || approx_wolfe);
  }

  if (accept) {
    // curvature pair
    float sy = 0.f, yy = 0.f, gg = 0.f;
    float disp = 0.f;   // steepest descent: largest coordinate displacement of the step that was just accepted
    if (a.sd && state == LB_SEARCH) {
      for (int t = threadIdx.x; t < D; t += kLbThreads) disp = fmaxf(disp, fabsf(xt[t] - x[t]));
      disp = block_max(disp, red);
    }
    if (state == LB_SEARCH && !a.sd) {
      for (int t = threadIdx.x; t < D; t += kLbThreads) {
        float s = xt[t] - x[t], y = -ft[t] - g[t];
        sy += s * y;
        yy += y * y;
      }
      sy = block_sum(sy, red);
      yy = block_sum(yy, red);
      if (sy > 1e-10f * yy && yy > 0.f) {
        float* Sk = S + (size_t)head * D;
        float* Yk = Y + (size_t)head * D;
        for (int t = threadIdx.x; t < D; t += kLbThreads) {
          Sk[t] = xt[t] - x[t];
          Yk[t] = -ft[t] - g[t];
        }
        if (threadIdx.x == 0) rho[head] = 1.0f / sy;
        head = (head + 1) % m;
        cnt = min(cnt + 1, m);
      }
    }
    for (int t = threadIdx.x; t < D; t += kLbThreads) {
      x[t] = xt[t];
      float gt = -ft[t];
      g[t] = gt;
      gg += gt * gt;
    }
    // stagnation: the energy surface is discontinuous at graph-cutoff crossings, so a replica can keep accepting
    // steps that gain nothing; 20 accepted steps without a gain above the fp32 resolution end it.  Every thread
    // reads the reference energy / stall count BEFORE the barriers of block_sum; thread 0 writes them after, so all
    // warps take the same branch below.
    const float eref = state == LB_INIT ? et : a.eref[r];
    const int st = state == LB_INIT ? 0 : a.stall[r];
    gg = block_sum(gg, red);
    __syncthreads();
    float grms = sqrtf(gg / D);
    if (threadIdx.x == 0) a.grms_out[r] = grms;
    bool stalled = false;
    if (state == LB_INIT) {
      if (threadIdx.x == 0) {
        a.eref[r] = et;
        a.stall[r] = 0;
      }
    } else if (et < eref - (1e-6f * fabsf(eref) + 1e-3f)) {
      if (threadIdx.x == 0) {
        a.eref[r] = et;
        a.stall[r] = 0;
      }
    } else {
      stalled = st + 1 >= a.stall_limit;
      if (threadIdx.x == 0) a.stall[r] = st + 1;
    }
    __syncthreads();
    if (stalled && grms >= a.grad_tol) {
      if (threadIdx.x == 0) {
        a.e[r] = et;
        a.state[r] = LB_DONE;
        a.status[r] = 3;   // stalled above the threshold (grms >= grad_tol here)
        a.hist[r] = (head << 16) | cnt;
        atomicSub(a.n_active, 1);
      }
      return;
    }
    if (grms < a.grad_tol) {
      if (threadIdx.x == 0) {
        a.e[r] = et;
        a.state[r] = LB_DONE;
        a.status[r] = 0;
        a.hist[r] = (head << 16) | cnt;
        atomicSub(a.n_active, 1);
      }
      return;
    }
    // two-loop recursion: d = -H g
    for (int t = threadIdx.x; t < D; t += kLbThreads) d[t] = g[t];
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      int i = (head - 1 - k + 2 * m) % m;
      const float* Si = S + (size_t)i * D;
      const float* Yi = Y + (size_t)i * D;
      float sq = 0.f;
      for (int t = threadIdx.x; t < D; t += kLbThreads) sq += Si[t] * d[t];
      sq = block_sum(sq, red);
      float al = rho[i] * sq;
      if (threadIdx.x == 0) alpha[i] = al;
      for (int t = threadIdx.x; t < D; t += kLbThreads) d[t] -= al * Yi[t];
      __syncthreads();
    }
    if (cnt > 0) {
      int i = (head - 1 + m) % m;
      const float* Si = S + (size_t)i * D;
      const float* Yi = Y + (size_t)i * D;
      float s_y = 0.f, y_y = 0.f;
      for (int t = threadIdx.x; t < D; t += kLbThreads) {
        s_y += Si[t] * Yi[t];
        y_y += Yi[t] * Yi[t];
      }
      s_y = block_sum(s_y, red);
      y_y = block_sum(y_y, red);
      float gam = s_y / y_y;
      for (int t = threadIdx.x; t < D; t += kLbThreads) d[t] *= gam;
      __syncthreads();
    }
    for (int k = cnt - 1; k >= 0; --k) {
      int i = (head - 1 - k + 2 * m) % m;
      const float* Si = S + (size_t)i * D;
      const float* Yi = Y + (size_t)i * D;
      float yr = 0.f;
      for (int t = threadIdx.x; t < D; t += kLbThreads) yr += Yi[t] * d[t];
      yr = block_sum(yr, red);
      float coef = alpha[i] - rho[i] * yr;
      for (int t = threadIdx.x; t < D; t += kLbThreads) d[t] += coef * Si[t];
      __syncthreads();
    }
    float gd = 0.f, dmax = 0.f;
    for (int t = threadIdx.x; t < D; t += kLbThreads) {
      float dv = -d[t];
      d[t] = dv;
      gd += g[t] * dv;
      dmax = fmaxf(dmax, fabsf(dv));
    }
    gd = block_sum(gd, red);
    dmax = block_max(dmax, red);
    if (!(gd < 0.f) || !isfinite(dmax)) {   // not a descent direction: restart with steepest descent
      cnt = 0;
      head = 0;
      gd = 0.f;
      dmax = 0.f;
      for (int t = threadIdx.x; t < D; t += kLbThreads) {
        float dv = -g[t];
        d[t] = dv;
        gd += g[t] * dv;
        dmax = fmaxf(dmax, fabsf(dv));
      }
      gd = block_sum(gd, red);
      dmax = block_max(dmax, red);
    }
    // unit step for quasi-Newton directions, displacement-capped; tiny first step for steepest descent
    step = cnt > 0 ? 1.0f : fminf(1.0f, 0.002f / dmax);
    if (a.sd && disp > 0.f) step = fminf(1.0f, 2.0f * disp / dmax);   // twice the displacement that worked last time
    if (step * dmax > a.max_disp) step = a.max_disp / dmax;
    __syncthreads();
    for (int t = threadIdx.x; t < D; t += kLbThreads) xt[t] = x[t] + step * d[t];
    if (threadIdx.x == 0) {
      a.e[r] = et;
      a.gd[r] = gd;
      a.step[r] = step;
      a.state[r] = LB_SEARCH;
      a.hist[r] = (head << 16) | cnt;
    }
  } else {
    // backtrack
    float dmax = 0.f, gg = 0.f;
    for (int t = threadIdx.x; t < D; t += kLbThreads) {
      dmax = fmaxf(dmax, fabsf(d[t]));
      gg += g[t] * g[t];
    }
    dmax = block_max(dmax, red);
    gg = block_sum(gg, red);
    step *= finite ? 0.5f : 0.1f;
    if (step * dmax < 2e-7f) {
      if (cnt > 0) {   // drop the history and retry along steepest descent
        cnt = 0;
        head = 0;
        float gd = 0.f, gmax = 0.f;
        for (int t = threadIdx.x; t < D; t += kLbThreads) {
          float dv = -g[t];
          d[t] = dv;
          gd += g[t] * dv;
          gmax = fmaxf(gmax, fabsf(dv));
        }
        gd = block_sum(gd, red);
        gmax = block_max(gmax, red);
        step = fminf(1.0f, 0.002f / gmax);
        if (threadIdx.x == 0) a.gd[r] = gd;
      } else {
        // the line search can no longer make progress in single precision: this IS the reference's stopping
        // rule (tolerance=1e-4 is below the fp32 force noise floor)
        if (threadIdx.x == 0) {
          a.state[r] = LB_DONE;
          a.status[r] = finite ? (sqrtf(gg / D) < a.grad_tol ? 0 : 3) : 2;
          a.hist[r] = 0;
          atomicSub(a.n_active, 1);
        }
        for (int t = threadIdx.x; t < D; t += kLbThreads) xt[t] = x[t];
        return;
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < D; t += kLbThreads) xt[t] = x[t] + step * d[t];
    if (threadIdx.x == 0) {
      a.step[r] = step;
      a.hist[r] = (head << 16) | cnt;
    }
  }
}

__global__ void lbfgs_init_kernel(int R, int* state, int* hist, int* status, int* n_active, int* n_evals, float* grms) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) {
    state[r] = LB_INIT;
    hist[r] = 0;
    status[r] = 1;
    n_evals[r] = 0;
    grms[r] = 0.0f;
  }
  if (r == 0) *n_active = R;
}

// ---------------------------------------------------------------- Langevin middle
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                           uint32_t out[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// v += dt F/m ; x += dt/2 v ; v = a v + sqrt(kT (1-a^2)/m) xi ; x += dt/2 v     (Appendix E)
__global__ void langevin_middle_kernel(int n, int N, const float* __restrict__ inv_mass, const float* __restrict__ force,
                                       float* __restrict__ pos, float* __restrict__ vel, float dt, float a_fric,
                                       float kT, uint64_t seed, uint64_t step) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float im = inv_mass[i % N];
  uint32_t rnd[4];
  philox4x32((uint32_t)i, (uint32_t)(step & 0xFFFFFFFFu), (uint32_t)(step >> 32), 0x4C616E67u, (uint32_t)(seed & 0xFFFFFFFFu),
             (uint32_t)(seed >> 32), rnd);
  float r1 = sqrtf(-2.0f * logf(u01(rnd[0]))), t1 = 6.283185307179586f * u01(rnd[1]);
  float r2 = sqrtf(-2.0f * logf(u01(rnd[2]))), t2 = 6.283185307179586f * u01(rnd[3]);
  float xi[3] = {r1 * cosf(t1), r1 * sinf(t1), r2 * cosf(t2)};
  float sig = sqrtf(kT * (1.0f - a_fric * a_fric) * im);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    size_t j = 3 * (size_t)i + k;
    float v = vel[j] + dt * force[j] * im;
    float x = pos[j] + 0.5f * dt * v;
    v = a_fric * v + sig * xi[k];
    x += 0.5f * dt * v;
    vel[j] = v;
    pos[j] = x;
  }
}

// The same step with holonomic distance constraints (the reference's default `constraints=HBonds`,
// helper_functions.py:937-946; OpenMM LangevinMiddleIntegrator):
//   v += dt F/m ; RATTLE(v) ; x += dt/2 v ; v = a v + noise ; x += dt/2 v ; SHAKE(x) ; v += (x_constrained - x) / dt
// One thread per (replica, cluster): a cluster is a connected set of constrained atoms (a heavy atom and its
// hydrogens, at most kMaxClusterAtoms atoms / constraints); unconstrained atoms are clusters of one.
constexpr int kMaxClusterAtoms = 8;
struct ClusterArgs {
  int n_clusters, N;
  const int* atom_ptr;   // [n_clusters + 1] -> cluster_atom
  const int* atoms;      // atom indices of a cluster
  const int* con_ptr;    // [n_clusters + 1] -> constraints
  const int2* con;       // (local a, local b) inside the cluster
  const float* con_d;    // target distance
  float tol;
};

__global__ void langevin_middle_constrained_kernel(int R, ClusterArgs c, const float* __restrict__ inv_mass,
                                                   const float* __restrict__ force, float* __restrict__ pos,
                                                   float* __restrict__ vel, float dt, float a_fric, float kT,
                                                   uint64_t seed, uint64_t step) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * c.n_clusters) return;
  const int rep = t / c.n_clusters, cl = t - rep * c.n_clusters;
  const int a0 = c.atom_ptr[cl], na = c.atom_ptr[cl + 1] - a0;
  const int k0 = c.con_ptr[cl], nk = c.con_ptr[cl + 1] - k0;
  float x0[kMaxClusterAtoms][3], x[kMaxClusterAtoms][3], v[kMaxClusterAtoms][3], im[kMaxClusterAtoms];
  size_t gi[kMaxClusterAtoms];
  for (int l = 0; l < na; ++l) {
    const int a = c.atoms[a0 + l];
    gi[l] = ((size_t)rep * c.N + a) * 3;
    im[l] = inv_mass[a];
    for (int k = 0; k < 3; ++k) {
      x0[l][k] = pos[gi[l] + k];
      v[l][k] = vel[gi[l] + k] + dt * force[gi[l] + k] * im[l];
    }
  }
  // RATTLE: remove the velocity components along the constraints
  for (int it = 0; it < 50; ++it) {
    bool done = true;
    for (int q = 0; q < nk; ++q) {
      const int2 ab = c.con[k0 + q];
      float rx = x0[ab.x][0] - x0[ab.y][0], ry = x0[ab.x][1] - x0[ab.y][1], rz = x0[ab.x][2] - x0[ab.y][2];
      float rv = rx * (v[ab.x][0] - v[ab.y][0]) + ry * (v[ab.x][1] - v[ab.y][1]) + rz * (v[ab.x][2] - v[ab.y][2]);
      float r2 = rx * rx + ry * ry + rz * rz;
      if (fabsf(rv) * dt > c.tol * r2) {
        done = false;
        float gmm = -rv / (r2 * (im[ab.x] + im[ab.y]));
        v[ab.x][0] += gmm * rx * im[ab.x];
        v[ab.x][1] += gmm * ry * im[ab.x];
        v[ab.x][2] += gmm * rz * im[ab.x];
        v[ab.y][0] -= gmm * rx * im[ab.y];
        v[ab.y][1] -= gmm * ry * im[ab.y];
        v[ab.y][2] -= gmm * rz * im[ab.y];
      }
    }
    if (done) break;
  }
  // half drift, Ornstein-Uhlenbeck refresh, half drift
  for (int l = 0; l < na; ++l) {
    const int a = c.atoms[a0 + l];
    uint32_t rnd[4];
    philox4x32((uint32_t)(rep * c.N + a), (uint32_t)(step & 0xFFFFFFFFu), (uint32_t)(step >> 32), 0x4C616E67u,
               (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32), rnd);
    float r1 = sqrtf(-2.0f * logf(u01(rnd[0]))), t1 = 6.283185307179586f * u01(rnd[1]);
    float r2 = sqrtf(-2.0f * logf(u01(rnd[2]))), t2 = 6.283185307179586f * u01(rnd[3]);
    const float xi[3] = {r1 * cosf(t1), r1 * sinf(t1), r2 * cosf(t2)};
    const float sig = sqrtf(kT * (1.0f - a_fric * a_fric) * im[l]);
    for (int k = 0; k < 3; ++k) {
      float xx = x0[l][k] + 0.5f * dt * v[l][k];
      v[l][k] = a_fric * v[l][k] + sig * xi[k];
      x[l][k] = xx + 0.5f * dt * v[l][k];
    }
  }
  // SHAKE along the old bond vectors, then the velocity correction of the middle scheme
  float xu[kMaxClusterAtoms][3];
  for (int l = 0; l < na; ++l)
    for (int k = 0; k < 3; ++k) xu[l][k] = x[l][k];
  for (int it = 0; it < 100; ++it) {
    bool done = true;
    for (int q = 0; q < nk; ++q) {
      const int2 ab = c.con[k0 + q];
      const float d = c.con_d[k0 + q];
      float rx = x[ab.x][0] - x[ab.y][0], ry = x[ab.x][1] - x[ab.y][1], rz = x[ab.x][2] - x[ab.y][2];
      float diff = d * d - (rx * rx + ry * ry + rz * rz);
      if (fabsf(diff) > 2.0f * c.tol * d * d) {
        done = false;
        float ox = x0[ab.x][0] - x0[ab.y][0], oy = x0[ab.x][1] - x0[ab.y][1], oz = x0[ab.x][2] - x0[ab.y][2];
        float gmm = diff / (2.0f * (rx * ox + ry * oy + rz * oz) * (im[ab.x] + im[ab.y]));
        x[ab.x][0] += gmm * ox * im[ab.x];
        x[ab.x][1] += gmm * oy * im[ab.x];
        x[ab.x][2] += gmm * oz * im[ab.x];
        x[ab.y][0] -= gmm * ox * im[ab.y];
        x[ab.y][1] -= gmm * oy * im[ab.y];
        x[ab.y][2] -= gmm * oz * im[ab.y];
      }
    }
    if (done) break;
  }
  const float idt = 1.0f / dt;
  for (int l = 0; l < na; ++l)
    for (int k = 0; k < 3; ++k) {
      pos[gi[l] + k] = x[l][k];
      vel[gi[l] + k] = v[l][k] + (x[l][k] - xu[l][k]) * idt;
    }
}

}  // namespace gnnis

using namespace gnnis;

template <typename T>
static int dyn_alloc(Ctx* c, T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  GNNIS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), (count ? count : 1) * sizeof(T)));
  return 0;
}

extern "C" {

int gnnis_set_bonds(gnnis_handle h, int32_t n_bonds, const int32_t* idx, const float* r0, const float* k) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (c) c->graph_gen++;
  if (!c) return -1;
  if (c->N <= 0) {
    c->err = "gnnis_set_bonds: call gnnis_set_system first";
    return -1;
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  if (n_bonds <= 0) {
    c->n_bonds = 0;
    return 0;
  }
  for (int b = 0; b < 2 * n_bonds; ++b)
    if (idx[b] < 0 || idx[b] >= c->N) {
      c->err = "gnnis_set_bonds: atom index out of range";
      return -1;
    }
  if (dyn_alloc(c, &c->bond_idx, (size_t)2 * n_bonds) || dyn_alloc(c, &c->bond_r0, (size_t)n_bonds) ||
      dyn_alloc(c, &c->bond_k, (size_t)n_bonds))
    return -1;
  GNNIS_CUDA_OK(cudaMemcpy(c->bond_idx, idx, 2 * n_bonds * sizeof(int), cudaMemcpyHostToDevice));
  GNNIS_CUDA_OK(cudaMemcpy(c->bond_r0, r0, n_bonds * sizeof(float), cudaMemcpyHostToDevice));
  GNNIS_CUDA_OK(cudaMemcpy(c->bond_k, k, n_bonds * sizeof(float), cudaMemcpyHostToDevice));
  c->n_bonds = n_bonds;
  return 0;
}

int gnnis_set_constraints(gnnis_handle h, int32_t n_con, const int32_t* idx, const float* dist, float tolerance) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (c) c->graph_gen++;
  if (!c) return -1;
  if (c->N <= 0) {
    c->err = "gnnis_set_constraints: call gnnis_set_system first";
    return -1;
  }
  if (n_con <= 0) {
    c->n_clusters = 0;
    return 0;
  }
  const int N = c->N;
  std::vector<int> parent(N);
  for (int i = 0; i < N; ++i) parent[i] = i;
  auto find = [&](int a) {
    while (parent[a] != a) a = parent[a] = parent[parent[a]];
    return a;
  };
  for (int q = 0; q < n_con; ++q) {
    int a = idx[2 * q], b = idx[2 * q + 1];
    if (a < 0 || a >= N || b < 0 || b >= N || a == b || !(dist[q] > 0.f)) {
      c->err = "gnnis_set_constraints: bad constraint";
      return -1;
    }
    parent[find(a)] = find(b);
  }
  std::vector<int> root_to_cluster(N, -1), atom_ptr(1, 0), atoms, local(N, -1);
  std::vector<std::vector<int>> members;
  for (int i = 0; i < N; ++i) {
    int r = find(i);
    if (root_to_cluster[r] < 0) {
      root_to_cluster[r] = (int)members.size();
      members.emplace_back();
    }
    members[root_to_cluster[r]].push_back(i);
  }
  for (auto& m : members) {
    if ((int)m.size() > kMaxClusterAtoms) {
      c->err = "gnnis_set_constraints: a constraint cluster has more than 8 atoms (only X-H style clusters are supported)";
      return -1;
    }
    for (size_t l = 0; l < m.size(); ++l) {
      local[m[l]] = (int)l;
      atoms.push_back(m[l]);
    }
    atom_ptr.push_back((int)atoms.size());
  }
  std::vector<std::vector<int>> con_of(members.size());
  for (int q = 0; q < n_con; ++q) con_of[root_to_cluster[find(idx[2 * q])]].push_back(q);
  std::vector<int> con_ptr(1, 0);
  std::vector<int2> con;
  std::vector<float> con_d;
  for (auto& lst : con_of) {
    for (int q : lst) {
      con.push_back(make_int2(local[idx[2 * q]], local[idx[2 * q + 1]]));
      con_d.push_back(dist[q]);
    }
    con_ptr.push_back((int)con.size());
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  if (dyn_alloc(c, &c->cl_atom_ptr, atom_ptr.size()) || dyn_alloc(c, &c->cl_atoms, atoms.size()) ||
      dyn_alloc(c, &c->cl_con_ptr, con_ptr.size()) || dyn_alloc(c, &c->cl_con, con.size()) ||
      dyn_alloc(c, &c->cl_con_d, con_d.size()))
    return -1;
  GNNIS_CUDA_OK(cudaMemcpy(c->cl_atom_ptr, atom_ptr.data(), atom_ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
  GNNIS_CUDA_OK(cudaMemcpy(c->cl_atoms, atoms.data(), atoms.size() * sizeof(int), cudaMemcpyHostToDevice));
  GNNIS_CUDA_OK(cudaMemcpy(c->cl_con_ptr, con_ptr.data(), con_ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
  GNNIS_CUDA_OK(cudaMemcpy(c->cl_con, con.data(), con.size() * sizeof(int2), cudaMemcpyHostToDevice));
  GNNIS_CUDA_OK(cudaMemcpy(c->cl_con_d, con_d.data(), con_d.size() * sizeof(float), cudaMemcpyHostToDevice));
  c->n_clusters = (int)members.size();
  c->con_tol = tolerance > 0.f ? tolerance : 1e-5f;
  return 0;
}

int gnnis_minimize(gnnis_handle h, float* pos_dev, float* e_rep_dev, int32_t* status_dev, int32_t* n_iter_host,
                   float grad_tol, int32_t max_iter, int32_t history, uint32_t flags, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if (!pos_dev || !e_rep_dev || !status_dev) {
    c->err = "gnnis_minimize: NULL argument";
    return -1;
  }
  if (c->R <= 0 || c->N <= 0) {
    c->err = "gnnis_minimize: system / replicas not set";
    return -1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  const int steepest = history < 0 ? 1 : 0;   // history < 0: steepest descent with the same line search and stopping rules
  if (history < 0) history = 1;
  if (history == 0) history = 8;
  if (history > 32) history = 32;
  if (max_iter <= 0) max_iter = 1000;
  const int R = c->R, D = 3 * c->N;
  const size_t n3 = (size_t)R * D;
  if (c->dyn_atoms != n3 || c->dyn_hist != history) {
    if (dyn_alloc(c, &c->lb_S, n3 * history) || dyn_alloc(c, &c->lb_Y, n3 * history) ||
        dyn_alloc(c, &c->lb_rho, (size_t)R * history) || dyn_alloc(c, &c->lb_alpha, (size_t)R * history) ||
        dyn_alloc(c, &c->lb_g, n3) || dyn_alloc(c, &c->lb_d, n3) || dyn_alloc(c, &c->lb_xprev, n3) ||
        dyn_alloc(c, &c->lb_ftmp, n3) || dyn_alloc(c, &c->lb_etmp, (size_t)R) || dyn_alloc(c, &c->lb_e, (size_t)R) ||
        dyn_alloc(c, &c->lb_step, (size_t)R) || dyn_alloc(c, &c->lb_eprev, (size_t)R) ||
        dyn_alloc(c, &c->lb_state, (size_t)R) || dyn_alloc(c, &c->lb_count, (size_t)R) || dyn_alloc(c, &c->lb_active, 1) ||
        dyn_alloc(c, &c->lb_eref, (size_t)R) || dyn_alloc(c, &c->lb_stall, (size_t)R) ||
        dyn_alloc(c, &c->lb_nevals, (size_t)R) || dyn_alloc(c, &c->lb_grms, (size_t)R))
      return -1;
    c->dyn_atoms = n3;
    c->dyn_hist = history;
  }
  LbArgs a{};
  a.D = D;
  a.m = history;
  a.x = pos_dev;
  a.xt = c->lb_xprev;   // trial positions
  a.ft = c->lb_ftmp;
  a.et = c->lb_etmp;
  a.g = c->lb_g;
  a.d = c->lb_d;
  a.S = c->lb_S;
  a.Y = c->lb_Y;
  a.rho = c->lb_rho;
  a.alpha = c->lb_alpha;
  a.e = c->lb_e;
  a.step = c->lb_step;
  a.gd = c->lb_eprev;
  a.eref = c->lb_eref;
  a.stall = c->lb_stall;
  a.state = c->lb_state;
  a.hist = c->lb_count;
  a.status = status_dev;
  a.n_active = c->lb_active;
  a.n_evals = c->lb_nevals;
  a.grms_out = c->lb_grms;
  a.grad_tol = grad_tol;
  a.max_disp = c->lb_max_disp;
  a.stall_limit = c->lb_stall_limit;
  a.sd = steepest;
  lbfgs_init_kernel<<<(R + 127) / 128, 128, 0, st>>>(R, a.state, a.hist, a.status, a.n_active, a.n_evals, a.grms_out);
  GNNIS_LAUNCH_CHECK();
  GNNIS_CUDA_OK(cudaMemcpyAsync(a.xt, pos_dev, n3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  uint32_t eflags = flags & ~GNNIS_FLAG_NO_FORCES;
  int it = 0, active = R;
  c->skip_state = a.state;   // finished replicas cost (almost) nothing in the following evaluations
  struct SkipReset {
    Ctx* c;
    ~SkipReset() { c->skip_state = nullptr; }
  } skip_reset{c};
  for (it = 0; it <= max_iter; ++it) {
    int rc = energy_force_impl(c, a.xt, c->lb_etmp, c->lb_ftmp, eflags, st, nullptr);
    if (rc) return rc;
    lbfgs_update_kernel<<<R, kLbThreads, 0, st>>>(a);
    GNNIS_LAUNCH_CHECK();
    if ((it % 8) == 7 || it == max_iter) {
      GNNIS_CUDA_OK(cudaMemcpyAsync(&active, a.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      GNNIS_CUDA_OK(cudaStreamSynchronize(st));
      if (active <= 0) {
        ++it;
        break;
      }
    }
  }
  GNNIS_CUDA_OK(cudaMemcpyAsync(e_rep_dev, a.e, R * sizeof(float), cudaMemcpyDeviceToDevice, st));
  GNNIS_CUDA_OK(cudaStreamSynchronize(st));
  if (n_iter_host) *n_iter_host = it;
  return 0;
}

int gnnis_minimize_stats(gnnis_handle h, int32_t* n_evals_dev, float* grad_rms_dev, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if (!c->lb_nevals || c->dyn_atoms != (size_t)c->R * 3 * c->N) {
    c->err = "gnnis_minimize_stats: no minimisation has run for the current replica set-up";
    return -1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_evals_dev) GNNIS_CUDA_OK(cudaMemcpyAsync(n_evals_dev, c->lb_nevals, c->R * sizeof(int), cudaMemcpyDeviceToDevice, st));
  if (grad_rms_dev) GNNIS_CUDA_OK(cudaMemcpyAsync(grad_rms_dev, c->lb_grms, c->R * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int gnnis_langevin(gnnis_handle h, float* pos_dev, float* vel_dev, const float* masses_host, float temperature,
                   float friction, float dt, int32_t n_steps, uint64_t seed, uint64_t first_step, uint32_t flags,
                   void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return -1;
  if (!pos_dev || !vel_dev || !masses_host) {
    c->err = "gnnis_langevin: NULL argument";
    return -1;
  }
  if (c->R <= 0 || c->N <= 0) {
    c->err = "gnnis_langevin: system / replicas not set";
    return -1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  std::vector<float> im(c->N);
  for (int i = 0; i < c->N; ++i) {
    if (!(masses_host[i] > 0.f)) {
      c->err = "gnnis_langevin: masses must be positive";
      return -1;
    }
    im[i] = 1.0f / masses_host[i];
  }
  if (dyn_alloc(c, &c->masses, (size_t)c->N)) return -1;
  GNNIS_CUDA_OK(cudaMemcpyAsync(c->masses, im.data(), c->N * sizeof(float), cudaMemcpyHostToDevice, st));
  GNNIS_CUDA_OK(cudaStreamSynchronize(st));
  const int n = c->R * c->N;
  const float a_fric = expf(-friction * dt);
  const float kT = 0.0083144626f * temperature;
  uint32_t eflags = flags & ~GNNIS_FLAG_NO_FORCES;
  for (int s = 0; s < n_steps; ++s) {
    int rc = energy_force_impl(c, pos_dev, c->h_e_dev, c->h_f_dev, eflags, st, nullptr);
    if (rc) return rc;
    if (c->n_clusters > 0) {
      ClusterArgs ca{c->n_clusters, c->N, c->cl_atom_ptr, c->cl_atoms, c->cl_con_ptr, c->cl_con, c->cl_con_d, c->con_tol};
      const int nt = c->R * c->n_clusters;
      langevin_middle_constrained_kernel<<<(nt + 127) / 128, 128, 0, st>>>(c->R, ca, c->masses, c->h_f_dev, pos_dev, vel_dev,
                                                                          dt, a_fric, kT, seed, first_step + (uint64_t)s);
    } else {
      langevin_middle_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, c->N, c->masses, c->h_f_dev, pos_dev, vel_dev, dt, a_fric, kT,
                                                             seed, first_step + (uint64_t)s);
    }
    GNNIS_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
