// forcefield.cu -- vacuum molecular-mechanics terms on the resident [R][N][3] layout.
//
// Replaces, for every replica at once, the forces OpenMM evaluates next to the GNN in the reference's replicated
// system (Simulation/Simulator.py:1597-1634, 2036-2143):
//   HarmonicBondForce       0.5 k (r - r0)^2
//   HarmonicAngleForce      0.5 k (theta - theta0)^2
//   PeriodicTorsionForce    k (1 + cos(n phi - phase))
//   Custom_electrostatic    q_i q_j 138.935456 / r                        all non-excepted intra-replica pairs, NoCutoff
//   Custom_lennard_jones    4 sqrt(e_i e_j) ((s/r)^12 - (s/r)^6), s = (s_i + s_j)/2
//   Custom_exception_force_with_scale   4 e (sr6^2 - sr6) + qq 138.935456 / r   per NonbondedForce exception
// One CTA per replica; positions and per-atom parameters live in shared memory.  Every atom is owned by one thread
// that sums, in a fixed order, the contributions of all sources (non-bonded) and of all terms it takes part in
// (per-atom term lists built on the host), so forces need no atomics and are bit-reproducible.  A term's energy is
// credited 1/arity to each of its atoms; the replica energy is a fixed-order block reduction.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace gnnis {

constexpr float kCoulombMM = 138.935456f;   // Simulator.py:2047 (the GB term uses 138.935485)
constexpr int kFfThreads = 256;   // 64 atoms per pass x 4 lanes per atom
constexpr int kFfLanes = 4;

struct FfArgs {
  int N, n_words;
  const float *charge, *sigma, *epsilon;   // [N]
  const uint32_t* excl;                    // [N][n_words] bit j of row i: pair (i, j) is an exception
  const int *term_ptr, *term_ref;          // per-atom term lists: ref = type | role << 2 | index << 4
  const int4* bond;                        // (i, j, -, -)
  const float2* bond_p;                    // (r0, k)
  const int4* angle;                       // (i, j, k, -)
  const float2* angle_p;                   // (theta0, k)
  const int4* tors;                        // (i, j, k, l)
  const float4* tors_p;                    // (n, phase, k, -)
  const int4* exc;                         // (i, j, -, -)
  const float4* exc_p;                     // (chargeprod, sigma, epsilon, -)
};

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 scale(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 ld3(const float* p, int i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

// dynamic smem: pos[3N] | q[N] | sig[N] | eps[N] | red[kFfThreads]
__global__ void __launch_bounds__(kFfThreads) mm_forces_kernel(FfArgs g, const float* __restrict__ pos,
                                                               float* __restrict__ e_rep, float* __restrict__ force,
                                                               int accumulate_energy, const int* __restrict__ skip) {
  if (skip && skip[blockIdx.x] == 2) return;
  extern __shared__ float smf[];
  float* sp = smf;
  float* sq = sp + 3 * g.N;
  float* ss = sq + g.N;
  float* se = ss + g.N;
  float* red = se + g.N;
  const int rep = blockIdx.x, tid = threadIdx.x;
  const float* rp = pos + (size_t)rep * g.N * 3;
  for (int t = tid; t < 3 * g.N; t += kFfThreads) sp[t] = rp[t];
  for (int t = tid; t < g.N; t += kFfThreads) {
    sq[t] = g.charge[t];
    ss[t] = g.sigma[t];
    se[t] = g.epsilon[t];
  }
  __syncthreads();
  // Four lanes share an atom: they split its source loop and its term list and combine their partial force / energy
  // with two shuffles in a fixed order (one thread per atom left 60 of 128 threads with a 60-iteration dependent chain).
  float e_acc = 0.0f;
  const int ln = tid % kFfLanes;
  for (int base = 0; base < g.N; base += kFfThreads / kFfLanes) {
    const int i = min(base + tid / kFfLanes, g.N - 1);
    const bool act = base + tid / kFfLanes < g.N;
    const V3 xi = ld3(sp, i);
    const float qi = sq[i] * kCoulombMM, si = ss[i], ei = se[i];
    float fx = 0.f, fy = 0.f, fz = 0.f, e = 0.f;
    // ---- non-bonded: all sources in ascending order, exceptions masked
    const uint32_t* row = g.excl + (size_t)i * g.n_words;
    for (int w = 0; w < g.n_words; ++w) {
      const uint32_t mask = row[w];
      const int j1 = act ? min(g.N, 32 * w + 32) : 0;
      for (int j = 32 * w + ln; j < j1; j += kFfLanes) {
        if (j == i || ((mask >> (j & 31)) & 1u)) continue;
        const V3 d = sub(xi, ld3(sp, j));
        const float r2 = dot(d, d);
        const float inv_r = rsqrtf(r2), inv_r2 = inv_r * inv_r;
        const float coul = qi * sq[j] * inv_r;
        const float sig = 0.5f * (si + ss[j]);
        const float eps = sqrtf(ei * se[j]);
        const float sr2 = sig * sig * inv_r2, sr6 = sr2 * sr2 * sr2;
        e += 0.5f * (coul + 4.0f * eps * (sr6 * sr6 - sr6));
        const float gmag = (coul + 24.0f * eps * (2.0f * sr6 * sr6 - sr6)) * inv_r2;   // -dE/dr / r
        fx += gmag * d.x;
        fy += gmag * d.y;
        fz += gmag * d.z;
      }
    }
    // ---- bonded terms and exceptions this atom takes part in
    for (int t = g.term_ptr[i] + ln; t < (act ? g.term_ptr[i + 1] : 0); t += kFfLanes) {
      const int ref = g.term_ref[t], type = ref & 3, role = (ref >> 2) & 3, idx = ref >> 4;
      if (type == 0) {
        const int4 a = g.bond[idx];
        const float2 p = g.bond_p[idx];
        const int o = role == 0 ? a.y : a.x;
        const V3 d = sub(xi, ld3(sp, o));
        const float r = sqrtf(dot(d, d));
        const float dr = r - p.x;
        e += 0.25f * p.y * dr * dr;
        const float gm = -p.y * dr / r;
        fx += gm * d.x;
        fy += gm * d.y;
        fz += gm * d.z;
      } else if (type == 1) {
        const int4 a = g.angle[idx];
        const float2 p = g.angle_p[idx];
        const V3 xj = ld3(sp, a.y);
        const V3 va = sub(ld3(sp, a.x), xj), vb = sub(ld3(sp, a.z), xj);
        const float la = sqrtf(dot(va, va)), lb = sqrtf(dot(vb, vb));
        const float ila = 1.0f / la, ilb = 1.0f / lb;
        float c = dot(va, vb) * ila * ilb;
        c = fminf(1.0f, fmaxf(-1.0f, c));
        const float theta = acosf(c);
        const float dth = theta - p.x;
        e += (0.5f / 3.0f) * p.y * dth * dth;
        const float sn = fmaxf(sqrtf(1.0f - c * c), 1e-6f);
        const float pre = p.y * dth / sn;   // -dE/dtheta * dtheta/dcos = k dth / sin
        // d cos / d x_i = (b_hat - c a_hat) / |a| ;  d cos / d x_k = (a_hat - c b_hat) / |b|
        const V3 ah = scale(va, ila), bh = scale(vb, ilb);
        const V3 gi = scale(sub(bh, scale(ah, c)), ila), gk = scale(sub(ah, scale(bh, c)), ilb);
        V3 gme;
        if (role == 0) gme = gi;
        else if (role == 2) gme = gk;
        else gme = {-(gi.x + gk.x), -(gi.y + gk.y), -(gi.z + gk.z)};
        fx += pre * gme.x;
        fy += pre * gme.y;
        fz += pre * gme.z;
      } else if (type == 2) {
        const int4 a = g.tors[idx];
        const float4 p = g.tors_p[idx];
        const V3 x0 = ld3(sp, a.x), x1 = ld3(sp, a.y), x2 = ld3(sp, a.z), x3 = ld3(sp, a.w);
        const V3 b1 = sub(x1, x0), b2 = sub(x2, x1), b3 = sub(x3, x2);
        const V3 n1 = cross(b1, b2), n2 = cross(b2, b3);
        const float lb2 = sqrtf(dot(b2, b2));
        const float phi = atan2f(lb2 * dot(b1, n2), dot(n1, n2));
        const float arg = p.x * phi - p.y;
        e += 0.25f * p.z * (1.0f + cosf(arg));
        const float dEdphi = -p.z * p.x * sinf(arg);
        // d phi / d x (Blondel & Karplus 1996, written for b1 = x1 - x0, b2 = x2 - x1, b3 = x3 - x2):
        // g1 = -(1 + b1.b2/|b2|^2) g0 + (b3.b2/|b2|^2) g3 ;  g2 = (b1.b2/|b2|^2) g0 - (1 + b3.b2/|b2|^2) g3
        const float in1 = 1.0f / dot(n1, n1), in2 = 1.0f / dot(n2, n2);
        const V3 g0 = scale(n1, -lb2 * in1), g3 = scale(n2, lb2 * in2);
        const float ib2 = 1.0f / (lb2 * lb2);
        const float s12 = dot(b1, b2) * ib2, s32 = dot(b3, b2) * ib2;
        V3 gme;
        if (role == 0) gme = g0;
        else if (role == 3) gme = g3;
        else if (role == 1) gme = {-g0.x - s12 * g0.x + s32 * g3.x, -g0.y - s12 * g0.y + s32 * g3.y, -g0.z - s12 * g0.z + s32 * g3.z};
        else gme = {-g3.x + s12 * g0.x - s32 * g3.x, -g3.y + s12 * g0.y - s32 * g3.y, -g3.z + s12 * g0.z - s32 * g3.z};
        fx -= dEdphi * gme.x;
        fy -= dEdphi * gme.y;
        fz -= dEdphi * gme.z;
      } else {
        const int4 a = g.exc[idx];
        const float4 p = g.exc_p[idx];
        const int o = role == 0 ? a.y : a.x;
        const V3 d = sub(xi, ld3(sp, o));
        const float r2 = dot(d, d);
        const float inv_r = rsqrtf(r2), inv_r2 = inv_r * inv_r;
        const float coul = p.x * kCoulombMM * inv_r;
        const float sr2 = p.y * p.y * inv_r2, sr6 = sr2 * sr2 * sr2;
        e += 0.5f * (coul + 4.0f * p.z * (sr6 * sr6 - sr6));
        const float gmag = (coul + 24.0f * p.z * (2.0f * sr6 * sr6 - sr6)) * inv_r2;
        fx += gmag * d.x;
        fy += gmag * d.y;
        fz += gmag * d.z;
      }
    }
#pragma unroll
    for (int o = 1; o < kFfLanes; o <<= 1) {
      e += __shfl_xor_sync(0xffffffffu, e, o);
      fx += __shfl_xor_sync(0xffffffffu, fx, o);
      fy += __shfl_xor_sync(0xffffffffu, fy, o);
      fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if (!act || ln != 0) continue;
    e_acc += e;
    if (force) {
      float* f = force + ((size_t)rep * g.N + i) * 3;
      f[0] += fx;
      f[1] += fy;
      f[2] += fz;
    }
  }
  // fixed-order block reduction of the replica energy
  red[tid] = e_acc;
  __syncthreads();
  for (int off = kFfThreads / 2; off > 0; off >>= 1) {
    if (tid < off) red[tid] += red[tid + off];
    __syncthreads();
  }
  if (tid == 0) e_rep[rep] = accumulate_energy ? e_rep[rep] + red[0] : red[0];
}

__global__ void zero_kernel(float* __restrict__ p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0f;
}

int launch_mm_forces(Ctx* c, const float* pos, float* e_rep, float* force, int accumulate, cudaStream_t st) {
  FfArgs g{};
  g.N = c->N;
  g.n_words = (c->N + 31) / 32;
  g.charge = c->ff_charge;
  g.sigma = c->ff_sigma;
  g.epsilon = c->ff_epsilon;
  g.excl = c->ff_excl;
  g.term_ptr = c->ff_term_ptr;
  g.term_ref = c->ff_term_ref;
  g.bond = c->ff_bond;
  g.bond_p = c->ff_bond_p;
  g.angle = c->ff_angle;
  g.angle_p = c->ff_angle_p;
  g.tors = c->ff_tors;
  g.tors_p = c->ff_tors_p;
  g.exc = c->ff_exc;
  g.exc_p = c->ff_exc_p;
  if (!accumulate && force) {
    size_t n3 = (size_t)c->R * c->N * 3;
    zero_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(force, n3);
    GNNIS_LAUNCH_CHECK();
  }
  size_t smem = (size_t)(6 * c->N + kFfThreads) * sizeof(float);
  mm_forces_kernel<<<c->R, kFfThreads, smem, st>>>(g, pos, e_rep, force, accumulate, c->skip_state);
  GNNIS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int ff_upload(Ctx* c, T** dst, const std::vector<T>& v) {
  if (*dst) {
    cudaFree(*dst);
    *dst = nullptr;
  }
  size_t n = v.size() ? v.size() : 1;
  GNNIS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(T)));
  if (v.size()) GNNIS_CUDA_OK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace gnnis

using namespace gnnis;

extern "C" int gnnis_set_forcefield(gnnis_handle h, const gnnis_forcefield* ff) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (c) c->graph_gen++;
  if (!c) return -1;
  if (c->N <= 0) {
    c->err = "gnnis_set_forcefield: call gnnis_set_system first";
    return -1;
  }
  if (!ff) {
    c->have_ff = false;
    return 0;
  }
  const int N = c->N;
  if (!ff->charge || !ff->sigma || !ff->epsilon) {
    c->err = "gnnis_set_forcefield: charge / sigma / epsilon are required";
    return -1;
  }
  auto bad = [&](const int32_t* idx, int count, int arity) {
    for (int t = 0; t < count * arity; ++t)
      if (idx[t] < 0 || idx[t] >= N) return true;
    return false;
  };
  if ((ff->n_bonds > 0 && bad(ff->bond_idx, ff->n_bonds, 2)) || (ff->n_angles > 0 && bad(ff->angle_idx, ff->n_angles, 3)) ||
      (ff->n_torsions > 0 && bad(ff->torsion_idx, ff->n_torsions, 4)) ||
      (ff->n_exceptions > 0 && bad(ff->exception_idx, ff->n_exceptions, 2))) {
    c->err = "gnnis_set_forcefield: atom index out of range";
    return -1;
  }
  GNNIS_CUDA_OK(cudaSetDevice(c->device));
  const int nw = (N + 31) / 32;
  std::vector<uint32_t> excl((size_t)N * nw, 0u);
  std::vector<std::vector<int>> lists(N);
  std::vector<int4> bond(ff->n_bonds), angle(ff->n_angles), tors(ff->n_torsions), exc(ff->n_exceptions);
  std::vector<float2> bond_p(ff->n_bonds), angle_p(ff->n_angles);
  std::vector<float4> tors_p(ff->n_torsions), exc_p(ff->n_exceptions);
  for (int b = 0; b < ff->n_bonds; ++b) {
    bond[b] = make_int4(ff->bond_idx[2 * b], ff->bond_idx[2 * b + 1], 0, 0);
    bond_p[b] = make_float2(ff->bond_r0[b], ff->bond_k[b]);
    for (int r = 0; r < 2; ++r) lists[ff->bond_idx[2 * b + r]].push_back(0 | (r << 2) | (b << 4));
  }
  for (int a = 0; a < ff->n_angles; ++a) {
    angle[a] = make_int4(ff->angle_idx[3 * a], ff->angle_idx[3 * a + 1], ff->angle_idx[3 * a + 2], 0);
    angle_p[a] = make_float2(ff->angle_theta0[a], ff->angle_k[a]);
    for (int r = 0; r < 3; ++r) lists[ff->angle_idx[3 * a + r]].push_back(1 | (r << 2) | (a << 4));
  }
  for (int t = 0; t < ff->n_torsions; ++t) {
    tors[t] = make_int4(ff->torsion_idx[4 * t], ff->torsion_idx[4 * t + 1], ff->torsion_idx[4 * t + 2], ff->torsion_idx[4 * t + 3]);
    tors_p[t] = make_float4((float)ff->torsion_periodicity[t], ff->torsion_phase[t], ff->torsion_k[t], 0.f);
    for (int r = 0; r < 4; ++r) lists[ff->torsion_idx[4 * t + r]].push_back(2 | (r << 2) | (t << 4));
  }
  for (int e = 0; e < ff->n_exceptions; ++e) {
    const int i = ff->exception_idx[2 * e], j = ff->exception_idx[2 * e + 1];
    if (i == j) {
      c->err = "gnnis_set_forcefield: exception between an atom and itself";
      return -1;
    }
    exc[e] = make_int4(i, j, 0, 0);
    exc_p[e] = make_float4(ff->exception_chargeprod[e], ff->exception_sigma[e], ff->exception_epsilon[e], 0.f);
    excl[(size_t)i * nw + (j >> 5)] |= 1u << (j & 31);
    excl[(size_t)j * nw + (i >> 5)] |= 1u << (i & 31);
    if (ff->exception_chargeprod[e] != 0.0f || ff->exception_epsilon[e] != 0.0f)
      for (int r = 0; r < 2; ++r) lists[ff->exception_idx[2 * e + r]].push_back(3 | (r << 2) | (e << 4));
  }
  std::vector<int> ptr(N + 1, 0), ref;
  for (int i = 0; i < N; ++i) {
    ptr[i + 1] = ptr[i] + (int)lists[i].size();
    ref.insert(ref.end(), lists[i].begin(), lists[i].end());
  }
  std::vector<float> q(ff->charge, ff->charge + N), sg(ff->sigma, ff->sigma + N), ep(ff->epsilon, ff->epsilon + N);
  if (ff_upload(c, &c->ff_charge, q) || ff_upload(c, &c->ff_sigma, sg) || ff_upload(c, &c->ff_epsilon, ep) ||
      ff_upload(c, &c->ff_excl, excl) || ff_upload(c, &c->ff_term_ptr, ptr) || ff_upload(c, &c->ff_term_ref, ref) ||
      ff_upload(c, &c->ff_bond, bond) || ff_upload(c, &c->ff_bond_p, bond_p) || ff_upload(c, &c->ff_angle, angle) ||
      ff_upload(c, &c->ff_angle_p, angle_p) || ff_upload(c, &c->ff_tors, tors) || ff_upload(c, &c->ff_tors_p, tors_p) ||
      ff_upload(c, &c->ff_exc, exc) || ff_upload(c, &c->ff_exc_p, exc_p))
    return -1;
  c->have_ff = true;
  return 0;
}
