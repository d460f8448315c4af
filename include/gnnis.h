/* gnnis.h -- C ABI of the B200-native GNN implicit-solvent energy/force engine (libgnnis.so).
 *
 * Drop-in boundary for the reference's hot path.  The reference has no FFI: its boundary is a
 * torch.nn.Module (`forward(positions, solvent_model, solvent_dielectric) -> energy`) consumed by
 * openmm-torch's TorchForce (Simulation/helper_functions.py:777-785).  The Python classes in
 * gnnimplicitsolvent_b200/GNN_Models.py keep that contract and call the functions below through ctypes
 * with raw device pointers (tensor.data_ptr()) and the caller's CUDA stream.  No torch types appear here.
 *
 * Conventions: every function returns 0 on success, <0 on error (gnnis_last_error gives the text);
 * no C++ exception crosses the ABI.  Units: nm, kJ/mol, e.  All floating-point buffers are fp32.
 * `*_dev` pointers are device memory owned by the caller; `*_host` pointers are host memory.
 * `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Nothing in
 * gnnis_energy_force synchronises the host; it is CUDA-graph capturable.  Outside a capture, an evaluation whose
 * buffers, flags and handle state recur is replayed from an internal CUDA graph (bit-identical results; the
 * environment variable GNNIS_GRAPHS=0, read by gnnis_create, disables this).
 */
#ifndef GNNIS_H_
#define GNNIS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gnnis_ctx* gnnis_handle;

/* flags for gnnis_energy_force / gnnis_minimize / gnnis_langevin */
#define GNNIS_FLAG_DELTA 0x1u      /* GNN3_Multisolvent_embedding_run_multiple_Delta (GNN_Models.py:947-1065):
                                      polar energy = E_GB(B') - E_GB(B)                                    */
#define GNNIS_FLAG_NO_FORCES 0x2u  /* energy only (skips the adjoint pass)                                */
#define GNNIS_FLAG_GB_ONLY 0x4u    /* plain GB-Neck2 (GNN_GBNeck_2, GNN_Models.py:173): B' = B, no SA term */
#define GNNIS_FLAG_MLP_BF16 0x8u   /* opt-in bf16 mode of the edge MLPs (the "bf16 MLP mode" of the north star): the
                                      per-edge contractions run as tcgen05.mma.kind::f16 with bf16 operands (fp32
                                      accumulate) and the SiLU activations come from MUFU.TANH (tanh.approx.f32) instead
                                      of ex2 + rcp.  Accuracy is stated separately and is NOT inside the fp32 gates
                                      (tests/test_gpu_parity.py::test_bf16_mode); needs the tensor-core
                                      edge kernels (-3 otherwise); never used for the benchmark's headline number */
#define GNNIS_FLAG_MLP_FP16 0x100u  /* the same kernels as GNNIS_FLAG_MLP_BF16 with fp16 operands (three more mantissa bits at
                                      the same speed; the operands of these MLPs -- activations, radial basis, adjoints,
                                      weights -- fit fp16's range).  Accuracy stated in test_half_precision_modes           */
#define GNNIS_FLAG_MLP_TF32X1 0x80u /* opt-in reduced-precision mode of the edge MLPs: ONE TF32 product per contraction
                                      instead of the fp32-accurate 3xTF32 default.  Accuracy is stated separately
                                      (tests/test_gpu_parity.py::test_tf32x1_mode: energies ~1e-3, forces ~1e-2 rel);
                                      never used for the benchmark's headline number                        */
#define GNNIS_FLAG_DATA_PREDICATE 0x20u /* GNN graph by torch_cluster.radius_graph semantics, sum((x_i-x_j)^2) < r^2
                                      (data path, GNN_Models.py:345-357), instead of eps-distance r < cutoff */
#define GNNIS_FLAG_VACUUM_ONLY 0x40u  /* only the vacuum force field of gnnis_set_forcefield (no GB, no GNN): the
                                      reference's vacuum "_ref_system" simulation                          */
#define GNNIS_FLAG_CELL_LIST 0x200u   /* build the GNN graph with the cell-list back-end (csrc/cell_list.cu: atoms binned into
                                      cells >= cutoff wide, 27-cell candidate search, bit-mask rows) instead of the
                                      all-pairs sweep; the CSR arrays are bit-identical.  Default for replicas of more than
                                      GNNIS_CELL_MIN atoms (environment, default 768); GNNIS_NEIGHBORS=cell|allpairs
                                      forces one back-end for every call                                    */
#define GNNIS_FLAG_ALL_PAIRS 0x400u   /* the opposite: the all-pairs sweep of pair_kernels.cu for this call                */
#define GNNIS_FLAG_MLP_CUDA_CORES 0x10u /* force the fp32 CUDA-core edge kernels instead of the default
                                      tcgen05 tensor-core kernels (3xTF32, fp32-accurate)                 */

/* The shipped architecture's state dict (MachineLearning/trained_models/GNN.pt["model"]),
 * as host pointers to the tensors in torch layout (Linear.weight = [out][in], row-major).  Layer l = 0..2
 * is interaction{l+1} (IN_layer_all_swish_2pass_tokens, GNN_Layers.py:2211-2246). */
typedef struct gnnis_weights {
  int32_t hidden;        /* 64  */
  int32_t hidden_token;  /* 128 */
  int32_t num_solvents;  /* rows of the embedding tables (42 in GNN.pt) */
  int32_t n_rbf;         /* 20  */
  const float* solvent_embedding; /* [num_solvents][hidden]                        GNN_Models.py:361 */
  const float* gamma_embedding;   /* [num_solvents]                                GNN_Models.py:362 */
  const float* message1_w[3];     /* [hidden][cin_l + n_rbf], cin = 6, 2*hidden, 2*hidden  GNN_Layers.py:2163 */
  const float* message1_b[3];     /* [hidden] */
  const float* message2_w[3];     /* [hidden][hidden]                              GNN_Layers.py:2164 */
  const float* message2_b[3];
  const float* lin1_w[3];         /* [hidden_token][hidden + hidden]               GNN_Layers.py:2227 */
  const float* lin1_b[3];
  const float* lin2_w[3];         /* [cout_l][hidden_token], cout = hidden, hidden, 2  GNN_Layers.py:2229 */
  const float* lin2_b[3];
} gnnis_weights;

int gnnis_create(gnnis_handle* out, int device);
int gnnis_destroy(gnnis_handle h);
const char* gnnis_last_error(gnnis_handle h);   /* h may be NULL: returns the last create() error */

/* replaces load_state_dict on the reference module (helper_functions.py:772) */
int gnnis_load_weights(gnnis_handle h, const gnnis_weights* w);

/* replaces the module constructor's `parameters` table (GNN_Models.py:68-70; format GNN_Trainer.py:829-869):
 * params7_host [n_atoms][7] = [q, rho, s, alpha, beta, gamma, radindex]; d0/m0 [n_unique*n_unique] are the
 * resampled neck tables (GNN_Layers.py:913-972; also checkpoint buffers aggregate_information._d0/_m0). */
int gnnis_set_system(gnnis_handle h, int32_t n_atoms, const float* params7_host,
                     const float* d0_host, const float* m0_host, int32_t n_unique);

/* replaces set_num_reps(num_reps, solvent_models, solvent_dielectric) (GNN_Models.py:741-769) */
int gnnis_set_replicas(gnnis_handle h, int32_t n_rep, const int32_t* solvent_id_host,
                       const float* dielectric_host);

/* ctor args fraction / scaling_factor (GNN_Models.py:704,718) and the GNN cutoff (hard-coded 0.6 at :795) */
int gnnis_set_model_constants(gnnis_handle h, float fraction, float scaling_factor, float cutoff);

/* replaces torch_cluster.radius_graph / the `edge_attributes < 0.6` boolean compaction (GNN_Models.py:794-804):
 * sorted CSR by target then source.  predicate 0 = run path  ||x_src - x_tgt + 1e-6|| < cutoff,
 * predicate 1 = data path  sum((x_i-x_j)^2) < cutoff^2 (torch_cluster semantics).
 * row_ptr_dev [R*N+1] int32, col_dev [capacity] int32 (global source index), r_dev [capacity] fp32
 * eps-distance (may be NULL).  *n_edges_host receives E (this call synchronises the stream).
 * predicate | GNNIS_PREDICATE_CELL_LIST selects the cell-list back-end, predicate | GNNIS_PREDICATE_ALL_PAIRS the
 * all-pairs sweep (neither: the handle's default, see GNNIS_FLAG_CELL_LIST); both give the same bytes.
 * Returns -2 (nothing written to col/r) when capacity < E. */
#define GNNIS_PREDICATE_CELL_LIST 0x10
#define GNNIS_PREDICATE_ALL_PAIRS 0x20
int gnnis_build_neighbors(gnnis_handle h, const float* pos_dev, int32_t predicate, int32_t* row_ptr_dev,
                          int32_t* col_dev, float* r_dev, int64_t capacity, int64_t* n_edges_host, void* stream);

/* replaces forward(positions, -1, .) + energy.backward() (GNN_Models.py:873-885 + autograd):
 * pos_dev [R][N][3]; e_rep_dev [R] per-replica energies (sum = the reference's scalar);
 * force_dev [R][N][3] = -dE/dpos (may be NULL with GNNIS_FLAG_NO_FORCES). */
int gnnis_energy_force(gnnis_handle h, const float* pos_dev, float* e_rep_dev, float* force_dev,
                       uint32_t flags, void* stream);

/* same call with HOST buffers (pinned or pageable): H2D of positions, evaluation, D2H of energies and
 * forces, one stream synchronisation at the end.  This is the end-to-end plug-in path bench.py times. */
int gnnis_energy_force_host(gnnis_handle h, const float* pos_host, float* e_rep_host, float* force_host,
                            uint32_t flags);

/* pipelined form of the call above for back-to-back batches (conformer streams, trajectory re-evaluation): submit
 * enqueues H2D copy, evaluation and D2H copy and returns at once; wait blocks until the OLDEST submission has
 * delivered its results.  At most two submissions may be in flight (a third returns -4); evaluations run in
 * submission order, the copies of neighbouring calls overlap them on separate streams.  The host buffers of a
 * submission must stay valid, and its results must not be read, until its wait returns; use pinned memory for the
 * copies to be asynchronous.  Do not mix with other calls on the handle while submissions are in flight. */
int gnnis_energy_force_host_submit(gnnis_handle h, const float* pos_host, float* e_rep_host, float* force_host,
                                   uint32_t flags);
int gnnis_energy_force_host_wait(gnnis_handle h);

/* ..._run_multiple_split (GNN_Models.py:1068-1083): per-atom polar and SA energies of the LAST evaluation */
int gnnis_get_atom_energies(gnnis_handle h, float* e_atom_dev, float* sa_atom_dev, void* stream);
/* Born radii B (GBNeck_interaction output, GNN_Layers.py:982-1008) and scaled B' of the LAST evaluation */
int gnnis_get_born_radii(gnnis_handle h, float* born_dev, float* born_scaled_dev, void* stream);
/* number of GNN edges found by the LAST evaluation (synchronises the stream) */
int gnnis_last_edge_count(gnnis_handle h, int64_t* n_edges_host, void* stream);

/* Vacuum molecular-mechanics terms added to every evaluation: what OpenMM evaluates next to the
 * GNN in the reference's replicated system (Simulation/Simulator.py:1597-1634, 2036-2143).  All arrays are host
 * pointers, per MOLECULE (every replica uses the same terms); units nm, kJ/mol, rad, e.
 *   bonds      0.5 k (r - r0)^2                       HarmonicBondForce
 *   angles     0.5 k (theta - theta0)^2               HarmonicAngleForce (theta at the middle atom)
 *   torsions   k (1 + cos(n phi - phase))             PeriodicTorsionForce (phi: IUPAC dihedral of i-j-k-l)
 *   charge, sigma, epsilon [N]                        NonbondedForce particles -> Custom_electrostatic
 *                                                     (q_i q_j 138.935456 / r) + Custom_lennard_jones
 *                                                     (Lorentz-Berthelot), all non-excepted pairs, no cutoff
 *   exceptions (i, j, chargeprod, sigma, epsilon)     NonbondedForce exceptions: the pair is removed from the two
 *                                                     sums and gets 4 eps ((s/r)^12 - (s/r)^6) + chargeprod 138.935456 / r
 *                                                     (Custom_exception_force_with_scale; zero parameters = plain exclusion)
 * Pass NULL to remove the force field again. */
typedef struct gnnis_forcefield {
  int32_t n_bonds;
  const int32_t* bond_idx;        /* [n_bonds][2] */
  const float *bond_r0, *bond_k;
  int32_t n_angles;
  const int32_t* angle_idx;       /* [n_angles][3] */
  const float *angle_theta0, *angle_k;
  int32_t n_torsions;
  const int32_t* torsion_idx;     /* [n_torsions][4] */
  const int32_t* torsion_periodicity;
  const float *torsion_phase, *torsion_k;
  const float *charge, *sigma, *epsilon;   /* [n_atoms] */
  int32_t n_exceptions;
  const int32_t* exception_idx;   /* [n_exceptions][2] */
  const float *exception_chargeprod, *exception_sigma, *exception_epsilon;
} gnnis_forcefield;
int gnnis_set_forcefield(gnnis_handle h, const gnnis_forcefield* ff);

/* Batched per-replica L-BFGS (replaces simulation.minimizeEnergy(tolerance, maxIterations) driven by
 * OpenMM's LocalEnergyMinimizer, helper_functions.py:614-624; semantics differ by design: one optimiser
 * per replica, SURVEY.md §7 hard part 5).  pos_dev is updated in place.  Stops a replica when its
 * gradient RMS < grad_tol (kJ/mol/nm) or after max_iter iterations (0 = 1000).  history = number of curvature pairs
 * per replica (0 = 8, at most 32); history < 0 selects steepest descent (no pairs; the first trial step of an iteration is
 * twice the displacement accepted last) with the same acceptance and stopping rules.
 * status_dev [R] int32: 0 converged (gradient RMS below grad_tol), 1 iteration limit, 2 non-finite energy/force
 * (positions restored), 3 stopped with a gradient RMS above grad_tol because the line search can no longer lower
 * the energy at fp32 resolution (20 accepted steps without a gain, or a step below 2e-7 nm) -- the model's energy
 * surface jumps wherever a pair crosses the 0.6 nm graph cutoff, so a replica can end on such a ledge; this is
 * where OpenMM's minimizeEnergy(tolerance=1e-4) of the reference stops as well.
 * Optional harmonic bonds k/2 (r-r0)^2 given per molecule (n_bonds, bond_idx_host[2*n_bonds],
 * bond_r0_host, bond_k_host) stand in for the vacuum force field when testing. */
int gnnis_set_bonds(gnnis_handle h, int32_t n_bonds, const int32_t* bond_idx_host,
                    const float* bond_r0_host, const float* bond_k_host);
int gnnis_minimize(gnnis_handle h, float* pos_dev, float* e_rep_dev, int32_t* status_dev, int32_t* n_iter_host,
                   float grad_tol, int32_t max_iter, int32_t history, uint32_t flags, void* stream);

/* Per-replica record of the last gnnis_minimize call: number of energy / force evaluations the replica took part in
 * until it stopped, and the gradient RMS (kJ/mol/nm) at its final accepted point.  Either pointer may be NULL. */
int gnnis_minimize_stats(gnnis_handle h, int32_t* n_evals_dev, float* grad_rms_dev, void* stream);

/* Batched LangevinMiddle integrator (replaces LangevinMiddleIntegrator(T, gamma, dt) + simulation.step,
 * helper_functions.py:787-789, Simulator.py:402).  masses_host [N] (amu), temperature K, friction 1/ps,
 * dt ps.  pos_dev/vel_dev [R][N][3] updated in place; Philox-style counter RNG keyed by (seed, step, atom). */
int gnnis_langevin(gnnis_handle h, float* pos_dev, float* vel_dev, const float* masses_host,
                   float temperature, float friction, float dt, int32_t n_steps, uint64_t seed,
                   uint64_t first_step, uint32_t flags, void* stream);

/* Distance constraints of the integrator (the reference's default `constraints=HBonds`, helper_functions.py:937-946):
 * idx_host [n][2] atom pairs, dist_host [n] nm; connected constraints form clusters of at most 8 atoms (a heavy atom
 * and its hydrogens).  gnnis_langevin then runs RATTLE on the velocities and SHAKE on the positions of every cluster
 * (relative tolerance, OpenMM default 1e-5).  n = 0 removes them.  The minimiser ignores constraints, as the
 * reference's entropy path does (constraints=None). */
int gnnis_set_constraints(gnnis_handle h, int32_t n_constraints, const int32_t* idx_host, const float* dist_host,
                          float tolerance);

/* Building-block check of the tensor-core path: D[128][n] = A[128][k] * B[n][k]^T (row-major device buffers)
 * through TMEM-resident A, 128B-swizzled shared-memory B, 3xTF32 tcgen05.mma.  k, n in {32, 64}. */
int gnnis_umma_selftest(gnnis_handle h, const float* a_dev, const float* b_dev, float* d_dev, int32_t k, int32_t n,
                        void* stream);

/* Number of kernels this library launched since the handle was created (bench.py's gpu_launches). */
int64_t gnnis_launch_count(gnnis_handle h);
/* One line saying which code path the current system / replica set-up runs on: replicas per unit, sub-unit split,
 * tcgen05 or CUDA-core edge kernels (and why), adjoint from the activation stash or by recomputation (and why),
 * shared-memory table mode, stash bytes.  The string belongs to the handle and is valid until the next call. */
const char* gnnis_describe(gnnis_handle h);
/* Time of the individual kernel families of the last gnnis_profile_eval call, in ms (order: see gnnis_stage_name) */
int gnnis_profile_eval(gnnis_handle h, const float* pos_dev, float* e_rep_dev, float* force_dev, uint32_t flags,
                       void* stream, float* stage_ms_host, int32_t n_stages);
int32_t gnnis_num_stages(void);
const char* gnnis_stage_name(int32_t i);

#ifdef __cplusplus
}
#endif
#endif /* GNNIS_H_ */
