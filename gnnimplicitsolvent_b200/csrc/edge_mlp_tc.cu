// edge_mlp_tc.cu -- the edge-message MLP and its adjoint on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same formulation and work decomposition as edge_mlp.cu (one persistent CTA per unit, 128-edge tiles, tables in
// shared memory, deterministic segmented reductions), but every dense contraction of the tile runs as
// tcgen05.mma.kind::tf32 with the accumulator in TMEM:
//     u  = rbf  . Wr^T      [128 x 24] x [24 x 64]      (A = rbf tile, written to TMEM by its 128 row-owner threads)
//     w  = v    . W2^T      [128 x 64] x [64 x 64]      (A = v = silu(u + P[tgt] + Q[src]) written back to TMEM)
//     vb = wbar . W2        [128 x 64] x [64 x 64]      adjoint
//     rb = ubar . Wr        [128 x 64] x [64 x 24]      adjoint of the radial columns
// A operands never touch shared memory: the epilogue of one product (tcgen05.ld -> registers -> bias/gather/SiLU)
// stores the next A operand straight back to TMEM (tcgen05.st), the "TS" form of tcgen05.mma.  B operands (the
// weights) are staged once per persistent CTA in shared memory in the canonical K-major 128B-swizzle layout.
//
// Precision: fp32-accurate.  Each product is evaluated as three TF32 MMAs, A_hi B_hi + A_hi B_lo + A_lo B_hi with
// x_hi = top 19 bits of x and x_lo = x - x_hi ("3xTF32"), accumulated in fp32 in TMEM; the dropped term is
// O(2^-22).  This is the mode the fp32 parity gates (energy 1e-4, forces 1e-3) are checked in.
//
// Thread mapping: 8 warps; warp w owns TMEM lanes (= tile rows) 32 (w & 3) .. +31 and columns 32 (w >> 2) .. +31,
// i.e. thread = (edge row, half of the 64 channels) -- the natural tcgen05.ld/st 32x32b shape.
#include "common.cuh"
#include "umma.cuh"

namespace gnnis {

using namespace umma;

constexpr int kTabLd = 68;   // padded row of the shared-memory node tables (bank spread for row-per-lane gathers)

struct EdgeTcArgs {
  int n_units, G, N, R;
  const int* row_ptr;
  const uint32_t* pair;
  const float* r_e;
  float inv_cutoff;
  const float *Wr, *W2, *b2, *Wr_t, *W2_t;   // torch-layout [64][20], [64][64]; transposed [20][64], [64][64]
  const float *P, *Q;
  float* a_out;
  const float* abar;
  float *Pbar, *Qbar, *rbar;
  int rbar_accumulate;
  const uint8_t* perm;
  int unit_atoms;
};

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

// issue the three TF32 products of D[128 x N] = A[128 x 8*ksteps] * B^T with A (hi, lo) in TMEM, B (hi, lo) in smem
__device__ __forceinline__ void issue_gemm_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_addr,
                                                  uint32_t b_lo_addr, int ksteps, int b_rows, uint32_t idesc) {
  uint32_t acc = 0;
#pragma unroll 1
  for (int term = 0; term < 3; ++term) {
    const uint32_t a = term == 2 ? a_lo : a_hi;
    const uint32_t b = term == 1 ? b_lo_addr : b_hi_addr;
#pragma unroll 1
    for (int s = 0; s < ksteps; ++s) {
      uint32_t baddr = b + (uint32_t)(s >> 2) * (uint32_t)b_rows * 128u + (uint32_t)(s & 3) * 32u;
      mma_tf32_ts(d_tmem, a + 8u * (uint32_t)s, smem_desc_k_sw128(baddr), idesc, acc);
      acc = 1;
    }
  }
}

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// radial basis row in registers (see rbf_row in edge_mlp.cu); o[24..31] = 0
template <bool DERIV>
__device__ __forceinline__ void rbf_regs(float r, float inv_cutoff, float (&o)[32], float (&od)[24]) {
  const float d = r * inv_cutoff;
  const float d2 = d * d, d4 = d2 * d2, d5 = d4 * d;
  const float env = 1.0f / d - 28.0f * d5 + 48.0f * (d5 * d) - 21.0f * (d5 * d2);
  float denv = 0.0f;
  if (DERIV) denv = -1.0f / d2 - 140.0f * d4 + 288.0f * d5 - 147.0f * (d5 * d);
  float s1, c1;
  sincospif(d, &s1, &c1);
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
  for (int k = kRbf; k < 32; ++k) o[k] = 0.0f;
  if (DERIV) {
#pragma unroll
    for (int k = kRbf; k < 24; ++k) od[k] = 0.0f;
  }
}

__device__ __forceinline__ void st_split(uint32_t t_hi, uint32_t t_lo, const float (&x)[32]) {
  uint32_t h[32], l[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) split_tf32(x[j], h[j], l[j]);
  tmem_st32(t_hi, h);
  tmem_st32(t_lo, l);
}

__device__ __forceinline__ void copy_rows_ld(float* __restrict__ dst, int ld_dst, const float* __restrict__ src,
                                             int ld_src, int nrows) {
  for (int t = threadIdx.x; t < nrows * 16; t += blockDim.x) {
    int r = t >> 4, c4 = t & 15;
    *reinterpret_cast<float4*>(dst + (size_t)r * ld_dst + c4 * 4) =
        *reinterpret_cast<const float4*>(src + (size_t)r * ld_src + c4 * 4);
  }
}
__device__ __forceinline__ void zero_rows_ld(float* __restrict__ dst, int ld_dst, int nrows) {
  for (int t = threadIdx.x; t < nrows * 16; t += blockDim.x) {
    int r = t >> 4, c4 = t & 15;
    *reinterpret_cast<float4*>(dst + (size_t)r * ld_dst + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// segmented add of tile rows into table rows (see seg_reduce_tile in edge_mlp.cu), table row stride `ld`
template <bool PERM>
__device__ __forceinline__ void seg_reduce_ld(const float* __restrict__ sX, const int* __restrict__ sKey,
                                              const uint8_t* __restrict__ sPerm, float* __restrict__ table, int ld,
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
        else table[cur * ld + c] += run;
      }
      cur = k;
      run = 0.0f;
      first = false;
    }
    if (k >= 0) run += v;
  }
  if (cur >= 0) {
    if (first && cont) sCarry[q * 64 + c] = run;
    else table[cur * ld + c] += run;
  }
  __syncthreads();
  if (tid < 64) {
#pragma unroll
    for (int qq = 1; qq < 4; ++qq) {
      int k = sCarryKey[qq];
      if (k >= 0) table[k * ld + tid] += sCarry[qq * 64 + tid];
    }
  }
  __syncthreads();
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
    issue_gemm_3xtf32(tm + ACC, tm + AH, tm + AL, smem_u32(bh), smem_u32(bl), K / 8, N, idesc_tf32(128, N));
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
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

// ---------------------------------------------------------------------------------------------- forward
// smem (1024-aligned): WrH 8K | WrL 8K | W2H 16K | W2L 16K | Z[128][68] | carry[256] | tgt[128] src[128] ckey[4] |
//                      b2[64] | (P, Q, acc)[3][unit_atoms][68]
constexpr uint32_t kFwdOffZ = 49152;
template <bool SMEM_TABLES>
__global__ void __launch_bounds__(kEdgeThreads, 1) edge_fwd_tc_kernel(EdgeTcArgs g) {
  extern __shared__ uint8_t smraw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t base = smem_u32(smraw);
  uint8_t* sm = smraw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t *sWrH = sm, *sWrL = sm + 8192, *sW2H = sm + 16384, *sW2L = sm + 32768;
  float* sZ = reinterpret_cast<float*>(sm + kFwdOffZ);
  float* sCarry = sZ + kTileE * kTileLd;
  int* sTgt = reinterpret_cast<int*>(sCarry + 256);
  int* sSrc = sTgt + kTileE;
  int* sCKey = sSrc + kTileE;
  float* sb2 = reinterpret_cast<float*>(sCKey + 4);
  float* sTab = sb2 + 64;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, ch = warp >> 2;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) mbar_init(&bar, 1);
  stage_b_operand(sWrH, sWrL, g.Wr, 64, 32, 64, kRbf, kRbf, 1);     // B[n = c][k] = Wr[c][k]
  stage_b_operand(sW2H, sW2L, g.W2, 64, 64, 64, 64, 64, 1);         // B[n = c][k] = W2[c][k]
  if (tid < 64) sb2[tid] = g.b2[tid];
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_slot;
  const uint32_t lane_addr = tm + ((uint32_t)(32 * (warp & 3)) << 16);
  constexpr uint32_t ACC_U = 0, ACC_W = 64, A1H = 128, A1L = 160, VH = 192, VL = 256;
  constexpr uint32_t idesc64 = idesc_tf32(128, 64);
  uint32_t phase = 0;

  for (int unit = blockIdx.x; unit < g.n_units; unit += gridDim.x) {
    const int rep0 = unit * g.G, rep1 = min(g.R, rep0 + g.G);
    const int atom0 = rep0 * g.N, natoms = (rep1 - rep0) * g.N;
    const int e_begin = g.row_ptr[atom0], e_end = g.row_ptr[atom0 + natoms];
    float *tP, *tQ, *tA;
    int ld;
    if (SMEM_TABLES) {
      ld = kTabLd;
      tP = sTab;
      tQ = tP + g.unit_atoms * kTabLd;
      tA = tQ + g.unit_atoms * kTabLd;
      copy_rows_ld(tP, ld, g.P + (size_t)atom0 * 64, 64, natoms);
      copy_rows_ld(tQ, ld, g.Q + (size_t)atom0 * 64, 64, natoms);
      zero_rows_ld(tA, ld, natoms);
    } else {
      ld = 64;
      tP = const_cast<float*>(g.P) + (size_t)atom0 * 64;
      tQ = const_cast<float*>(g.Q) + (size_t)atom0 * 64;
      tA = g.a_out + (size_t)atom0 * 64;
      zero_rows_ld(tA, ld, natoms);
    }
    __syncthreads();

    for (int t0 = e_begin; t0 < e_end; t0 += kTileE) {
      // ---- A: edge rows, radial basis -> TMEM (warps 0-3 own the 128 rows)
      if (ch == 0) {
        int e = t0 + row;
        bool valid = e < e_end;
        uint32_t pr = valid ? g.pair[e] : 0u;
        sTgt[row] = valid ? (int)(pr >> 16) : -1;
        sSrc[row] = valid ? (int)(pr & 0xFFFFu) : -1;
        float o[32], od[24];
        rbf_regs<false>(valid ? g.r_e[e] : 0.3f, g.inv_cutoff, o, od);
        st_split(lane_addr + A1H, lane_addr + A1L, o);
        tmem_st_wait();
      }
      fence_before_sync();
      __syncthreads();
      // ---- B: u = rbf . Wr^T
      if (tid == 0) {
        fence_after_sync();
        issue_gemm_3xtf32(tm + ACC_U, tm + A1H, tm + A1L, smem_u32(sWrH), smem_u32(sWrL), 3, 64, idesc64);
        mma_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      // ---- C: v = silu(u + P[tgt] + Q[src]) -> TMEM
      {
        uint32_t ur[32];
        tmem_ld32(lane_addr + ACC_U + 32 * ch, ur);
        tmem_ld_wait();
        float v[32];
        const int tg = sTgt[row], sr = sSrc[row];
        if (tg >= 0) {
          const float* pp = tP + tg * ld + 32 * ch;
          const float* qq = tQ + sr * ld + 32 * ch;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 p = *reinterpret_cast<const float4*>(pp + 4 * j4);
            float4 q = *reinterpret_cast<const float4*>(qq + 4 * j4);
            v[4 * j4 + 0] = silu_f(__uint_as_float(ur[4 * j4 + 0]) + (p.x + q.x));
            v[4 * j4 + 1] = silu_f(__uint_as_float(ur[4 * j4 + 1]) + (p.y + q.y));
            v[4 * j4 + 2] = silu_f(__uint_as_float(ur[4 * j4 + 2]) + (p.z + q.z));
            v[4 * j4 + 3] = silu_f(__uint_as_float(ur[4 * j4 + 3]) + (p.w + q.w));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.0f;
        }
        st_split(lane_addr + VH + 32 * ch, lane_addr + VL + 32 * ch, v);
        tmem_st_wait();
      }
      fence_before_sync();
      __syncthreads();
      // ---- D: w = v . W2^T
      if (tid == 0) {
        fence_after_sync();
        issue_gemm_3xtf32(tm + ACC_W, tm + VH, tm + VL, smem_u32(sW2H), smem_u32(sW2L), 8, 64, idesc64);
        mma_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      // ---- E: z = silu(w + b2) -> smem tile
      {
        uint32_t wr[32];
        tmem_ld32(lane_addr + ACC_W + 32 * ch, wr);
        tmem_ld_wait();
        float* zrow = sZ + row * kTileLd + 32 * ch;
        const float* bb = sb2 + 32 * ch;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 b = *reinterpret_cast<const float4*>(bb + 4 * j4);
          float4 z;
          z.x = silu_f(__uint_as_float(wr[4 * j4 + 0]) + b.x);
          z.y = silu_f(__uint_as_float(wr[4 * j4 + 1]) + b.y);
          z.z = silu_f(__uint_as_float(wr[4 * j4 + 2]) + b.z);
          z.w = silu_f(__uint_as_float(wr[4 * j4 + 3]) + b.w);
          *reinterpret_cast<float4*>(zrow + 4 * j4) = z;
        }
      }
      fence_before_sync();
      __syncthreads();
      // ---- F: a[tgt] += z
      seg_reduce_ld<false>(sZ, sTgt, nullptr, tA, ld, sCarry, sCKey, tid);
    }
    if (SMEM_TABLES) {
      copy_rows_ld(g.a_out + (size_t)atom0 * 64, 64, tA, ld, natoms);
      __syncthreads();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------- adjoint
// smem (1024-aligned): WrH 8K WrL 8K | W2H 16K W2L 16K | W2tH 16K W2tL 16K | WrtH 8K WrtL 8K | U[128][68] |
//   carry[256] | tgt[128] src[128] ckey[4] | b2[64] | perm[128 B] | (P, Q, abar, Pbar, Qbar)[5][unit_atoms][68]
constexpr uint32_t kBwdOffU = 98304;
template <bool SMEM_TABLES>
__global__ void __launch_bounds__(kEdgeThreads, 1) edge_bwd_tc_kernel(EdgeTcArgs g) {
  extern __shared__ uint8_t smraw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t base = smem_u32(smraw);
  uint8_t* sm = smraw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t *sWrH = sm, *sWrL = sm + 8192, *sW2H = sm + 16384, *sW2L = sm + 32768;
  uint8_t *sW2tH = sm + 49152, *sW2tL = sm + 65536, *sWrtH = sm + 81920, *sWrtL = sm + 90112;
  float* sU = reinterpret_cast<float*>(sm + kBwdOffU);
  float* sCarry = sU + kTileE * kTileLd;
  int* sTgt = reinterpret_cast<int*>(sCarry + 256);
  int* sSrc = sTgt + kTileE;
  int* sCKey = sSrc + kTileE;
  float* sb2 = reinterpret_cast<float*>(sCKey + 4);
  uint8_t* sPerm = reinterpret_cast<uint8_t*>(sb2 + 64);
  float* sTab = reinterpret_cast<float*>(sPerm + kTileE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, ch = warp >> 2;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) mbar_init(&bar, 1);
  stage_b_operand(sWrH, sWrL, g.Wr, 64, 32, 64, kRbf, kRbf, 1);      // u    : B[n = c ][k    ] = Wr[c][k]
  stage_b_operand(sW2H, sW2L, g.W2, 64, 64, 64, 64, 64, 1);          // w    : B[n = c ][k    ] = W2[c][k]
  stage_b_operand(sW2tH, sW2tL, g.W2_t, 64, 64, 64, 64, 64, 1);      // vbar : B[n = k'][k = c] = W2[c][k'] = W2_t[k'][c]
  stage_b_operand(sWrtH, sWrtL, g.Wr_t, 32, 64, kRbf, 64, 64, 1);    // rbfb : B[n = k'][k = c] = Wr[c][k'] = Wr_t[k'][c]
  if (tid < 64) sb2[tid] = g.b2[tid];
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_slot;
  const uint32_t lane_addr = tm + ((uint32_t)(32 * (warp & 3)) << 16);
  constexpr uint32_t ACC0 = 0, ACC1 = 64, ACCR = 128, A1H = 160, A1L = 192, XH = 224, XL = 288;
  constexpr uint32_t idesc64 = idesc_tf32(128, 64), idesc32 = idesc_tf32(128, 32);
  uint32_t phase = 0;

  for (int unit = blockIdx.x; unit < g.n_units; unit += gridDim.x) {
    const int rep0 = unit * g.G, rep1 = min(g.R, rep0 + g.G);
    const int atom0 = rep0 * g.N, natoms = (rep1 - rep0) * g.N;
    const int e_begin = g.row_ptr[atom0], e_end = g.row_ptr[atom0 + natoms];
    float *tP, *tQ, *tAb, *tPb, *tQb;
    int ld;
    if (SMEM_TABLES) {
      ld = kTabLd;
      tP = sTab;
      tQ = tP + g.unit_atoms * kTabLd;
      tAb = tQ + g.unit_atoms * kTabLd;
      tPb = tAb + g.unit_atoms * kTabLd;
      tQb = tPb + g.unit_atoms * kTabLd;
      copy_rows_ld(tP, ld, g.P + (size_t)atom0 * 64, 64, natoms);
      copy_rows_ld(tQ, ld, g.Q + (size_t)atom0 * 64, 64, natoms);
      copy_rows_ld(tAb, ld, g.abar + (size_t)atom0 * 64, 64, natoms);
    } else {
      ld = 64;
      tP = const_cast<float*>(g.P) + (size_t)atom0 * 64;
      tQ = const_cast<float*>(g.Q) + (size_t)atom0 * 64;
      tAb = const_cast<float*>(g.abar) + (size_t)atom0 * 64;
      tPb = g.Pbar + (size_t)atom0 * 64;
      tQb = g.Qbar + (size_t)atom0 * 64;
    }
    zero_rows_ld(tPb, ld, natoms);
    zero_rows_ld(tQb, ld, natoms);
    __syncthreads();

    for (int t0 = e_begin; t0 < e_end; t0 += kTileE) {
      float drbf[24];   // d rbf_k / dr of this thread's row (used by warps 0-3 only)
      if (ch == 0) {
        int e = t0 + row;
        bool valid = e < e_end;
        uint32_t pr = valid ? g.pair[e] : 0u;
        sTgt[row] = valid ? (int)(pr >> 16) : -1;
        sSrc[row] = valid ? (int)(pr & 0xFFFFu) : -1;
        sPerm[row] = valid ? g.perm[e] : (uint8_t)row;
        float o[32];
        rbf_regs<true>(valid ? g.r_e[e] : 0.3f, g.inv_cutoff, o, drbf);
        st_split(lane_addr + A1H, lane_addr + A1L, o);
        tmem_st_wait();
      }
      fence_before_sync();
      __syncthreads();
      // ---- u = rbf . Wr^T
      if (tid == 0) {
        fence_after_sync();
        issue_gemm_3xtf32(tm + ACC0, tm + A1H, tm + A1L, smem_u32(sWrH), smem_u32(sWrL), 3, 64, idesc64);
        mma_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      const int tg = sTgt[row], sr = sSrc[row];
      float dsu[32];
      {
        uint32_t ur[32];
        tmem_ld32(lane_addr + ACC0 + 32 * ch, ur);
        tmem_ld_wait();
        float v[32];
        if (tg >= 0) {
          const float* pp = tP + tg * ld + 32 * ch;
          const float* qq = tQ + sr * ld + 32 * ch;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 p = *reinterpret_cast<const float4*>(pp + 4 * j4);
            float4 q = *reinterpret_cast<const float4*>(qq + 4 * j4);
            silu_fd(__uint_as_float(ur[4 * j4 + 0]) + (p.x + q.x), v[4 * j4 + 0], dsu[4 * j4 + 0]);
            silu_fd(__uint_as_float(ur[4 * j4 + 1]) + (p.y + q.y), v[4 * j4 + 1], dsu[4 * j4 + 1]);
            silu_fd(__uint_as_float(ur[4 * j4 + 2]) + (p.z + q.z), v[4 * j4 + 2], dsu[4 * j4 + 2]);
            silu_fd(__uint_as_float(ur[4 * j4 + 3]) + (p.w + q.w), v[4 * j4 + 3], dsu[4 * j4 + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = 0.0f;
            dsu[j] = 0.0f;
          }
        }
        st_split(lane_addr + XH + 32 * ch, lane_addr + XL + 32 * ch, v);
        tmem_st_wait();
      }
      fence_before_sync();
      __syncthreads();
      // ---- w = v . W2^T ; wbar = abar[tgt] * silu'(w + b2)
      if (tid == 0) {
        fence_after_sync();
        issue_gemm_3xtf32(tm + ACC1, tm + XH, tm + XL, smem_u32(sW2H), smem_u32(sW2L), 8, 64, idesc64);
        mma_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      {
        uint32_t wr[32];
        tmem_ld32(lane_addr + ACC1 + 32 * ch, wr);
        tmem_ld_wait();
        float wb[32];
        if (tg >= 0) {
          const float* ab = tAb + tg * ld + 32 * ch;
          const float* bb = sb2 + 32 * ch;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 a = *reinterpret_cast<const float4*>(ab + 4 * j4);
            float4 b = *reinterpret_cast<const float4*>(bb + 4 * j4);
            wb[4 * j4 + 0] = a.x * silu_d(__uint_as_float(wr[4 * j4 + 0]) + b.x);
            wb[4 * j4 + 1] = a.y * silu_d(__uint_as_float(wr[4 * j4 + 1]) + b.y);
            wb[4 * j4 + 2] = a.z * silu_d(__uint_as_float(wr[4 * j4 + 2]) + b.z);
            wb[4 * j4 + 3] = a.w * silu_d(__uint_as_float(wr[4 * j4 + 3]) + b.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) wb[j] = 0.0f;
        }
        st_split(lane_addr + XH + 32 * ch, lane_addr + XL + 32 * ch, wb);
        tmem_st_wait();
      }
      fence_before_sync();
      __syncthreads();
      // ---- vbar = wbar . W2 ; ubar = vbar * silu'(u)
      if (tid == 0) {
        fence_after_sync();
        issue_gemm_3xtf32(tm + ACC0, tm + XH, tm + XL, smem_u32(sW2tH), smem_u32(sW2tL), 8, 64, idesc64);
        mma_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      {
        uint32_t vr[32];
        tmem_ld32(lane_addr + ACC0 + 32 * ch, vr);
        tmem_ld_wait();
        float ub[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) ub[j] = __uint_as_float(vr[j]) * dsu[j];
        float* urow = sU + row * kTileLd + 32 * ch;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<float4*>(urow + 4 * j4) = make_float4(ub[4 * j4], ub[4 * j4 + 1], ub[4 * j4 + 2], ub[4 * j4 + 3]);
        st_split(lane_addr + XH + 32 * ch, lane_addr + XL + 32 * ch, ub);
        tmem_st_wait();
      }
      fence_before_sync();
      __syncthreads();
      // ---- rbfbar = ubar . Wr  (runs while the segmented reductions below use the smem copy of ubar)
      if (tid == 0) {
        fence_after_sync();
        issue_gemm_3xtf32(tm + ACCR, tm + XH, tm + XL, smem_u32(sWrtH), smem_u32(sWrtL), 8, 32, idesc32);
        mma_commit(&bar);
      }
      seg_reduce_ld<false>(sU, sTgt, nullptr, tPb, ld, sCarry, sCKey, tid);   // Pbar[tgt] += ubar
      seg_reduce_ld<true>(sU, sSrc, sPerm, tQb, ld, sCarry, sCKey, tid);      // Qbar[src] += ubar
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      if (ch == 0) {
        uint32_t rr[32];
        tmem_ld32(lane_addr + ACCR, rr);
        tmem_ld_wait();
        if (tg >= 0) {
          float acc = 0.0f;
#pragma unroll
          for (int k = 0; k < kRbf; ++k) acc = fmaf(__uint_as_float(rr[k]), drbf[k], acc);
          int repl = tg / g.N;
          size_t slot = ((size_t)atom0 + tg) * g.N + (sr - repl * g.N);
          g.rbar[slot] = g.rbar_accumulate ? g.rbar[slot] + acc : acc;
        }
      }
      fence_before_sync();
      __syncthreads();
    }
    if (SMEM_TABLES) {
      copy_rows_ld(g.Pbar + (size_t)atom0 * 64, 64, tPb, ld, natoms);
      copy_rows_ld(g.Qbar + (size_t)atom0 * 64, 64, tQb, ld, natoms);
      __syncthreads();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------- launchers
static size_t fwd_tc_smem(const Ctx* c) {
  size_t b = kFwdOffZ + (size_t)(kTileE * kTileLd + 256 + 2 * kTileE + 4 + 64) * 4 + 1024;
  if (c->smem_tables) b += 3 * (size_t)c->G * c->N * kTabLd * 4;
  return b;
}
static size_t bwd_tc_smem(const Ctx* c) {
  size_t b = kBwdOffU + (size_t)(kTileE * kTileLd + 256 + 2 * kTileE + 4 + 64) * 4 + kTileE + 1024;
  if (c->smem_tables) b += 5 * (size_t)c->G * c->N * kTabLd * 4;
  return b;
}

static EdgeTcArgs make_tc_args(Ctx* c, int l) {
  EdgeTcArgs g{};
  g.n_units = c->n_units;
  g.G = c->G;
  g.N = c->N;
  g.R = c->R;
  g.row_ptr = c->row_ptr;
  g.pair = c->pair;
  g.r_e = c->r_e;
  g.inv_cutoff = (float)(1.0 / (double)c->cutoff);
  g.Wr = c->L[l].Wr;
  g.W2 = c->L[l].W2;
  g.b2 = c->L[l].b2;
  g.Wr_t = c->L[l].Wr_t;
  g.W2_t = c->L[l].W2_t;
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

bool tc_path_fits(const Ctx* c) { return fwd_tc_smem(c) <= 227 * 1024 && bwd_tc_smem(c) <= 227 * 1024; }

int launch_edge_fwd_tc(Ctx* c, int l, cudaStream_t st) {
  EdgeTcArgs g = make_tc_args(c, l);
  int grid = c->n_units < c->num_sms ? c->n_units : c->num_sms;
  size_t smem = fwd_tc_smem(c);
  if (c->smem_tables) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_fwd_tc_kernel<true><<<grid, kEdgeThreads, smem, st>>>(g);
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_fwd_tc_kernel<false><<<grid, kEdgeThreads, smem, st>>>(g);
  }
  GNNIS_LAUNCH_CHECK();
  return 0;
}

int launch_edge_bwd_tc(Ctx* c, int l, int rbar_accumulate, cudaStream_t st) {
  EdgeTcArgs g = make_tc_args(c, l);
  g.rbar_accumulate = rbar_accumulate;
  int grid = c->n_units < c->num_sms ? c->n_units : c->num_sms;
  size_t smem = bwd_tc_smem(c);
  if (c->smem_tables) {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_bwd_tc_kernel<true><<<grid, kEdgeThreads, smem, st>>>(g);
  } else {
    GNNIS_CUDA_OK(cudaFuncSetAttribute(edge_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_bwd_tc_kernel<false><<<grid, kEdgeThreads, smem, st>>>(g);
  }
  GNNIS_LAUNCH_CHECK();
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
