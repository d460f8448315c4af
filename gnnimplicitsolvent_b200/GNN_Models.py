"""Drop-in host-side mirror of the reference's model classes for the energy/force hot path.

Same class names, constructor arguments, `set_num_reps`, `forward` signatures, return values and state-dict
keys as MachineLearning/GNN_Models.py of the reference, but the arithmetic is done by the hand-written
sm_100a kernels of libgnnis.so through the C ABI (include/gnnis.h); PyTorch only carries tensors, streams
and autograd plumbing.  There is no CPU / eager fallback: calling `forward` without the CUDA library or
without a GPU raises.

    reference class (MachineLearning/GNN_Models.py)                 here
    GNN_GBNeck_2                                     :173           GB-only engine mode
    GNN3_Multisolvent_embedding                      :326-498       forward(data) -> (energy, forces)
    GNN3_Multisolvent_embedding_run_multiple         :700-944       forward(positions, solvent_model, dielectric)
    GNN3_Multisolvent_embedding_run_multiple_Delta   :947-1065      forward(positions)
    GNN3_Multisolvent_embedding_run_multiple_split   :1068-1083     -> (E_atom[n,1], SA_atom[n,1])

Gradient contract (openmm-torch TorchForce): the returned energy carries a
torch.autograd.Function node whose backward is `grad_out * dE/dpos`, with dE/dpos = -forces taken from the
analytic adjoint kernels of the same evaluation; `energy.backward()` therefore fills `positions.grad`
without a second pass.  `energy_force(positions[R,N,3]) -> (E[R], F[R,N,3])` is the native batched call.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, gbneck
from .engine import Engine, GnnisError

DEFAULT_UNIQUE_RADII = list(gbneck.DEFAULT_UNIQUE_RADII)   # GNN_Models.py:38 (unsorted on purpose)


class _EnergyFunction(torch.autograd.Function):
    """E[R] = engine(positions); backward: dE/dpos = -F scaled by the incoming per-replica gradient."""

    @staticmethod
    def forward(ctx, positions, module, delta):
        eng = module._engine_ready()
        e, f = eng.energy_force(positions, delta=delta, gb_only=module._gb_only, data_predicate=module._data_predicate)
        ctx.save_for_backward(f)
        ctx.in_shape = positions.shape
        ctx.in_dtype = positions.dtype
        return e

    @staticmethod
    def backward(ctx, grad_e):
        (f,) = ctx.saved_tensors
        g = -(f * grad_e.to(f.dtype).reshape(-1, 1, 1))
        return g.reshape(ctx.in_shape).to(ctx.in_dtype), None, None


class _Tables(torch.nn.Module):
    """GBNeck_interaction / GBNeck_energies_no_dielectric buffers (GNN_Layers.py:913-972, 1978-1980)."""

    def __init__(self, d0, m0, with_dielectrics):
        super().__init__()
        self.register_buffer("_d0", torch.from_numpy(np.array(d0, dtype=np.float32)))
        self.register_buffer("_m0", torch.from_numpy(np.array(m0, dtype=np.float32)))
        if with_dielectrics:
            self.register_buffer("_soluteDielectric", torch.tensor(1.0))
            self.register_buffer("_solventDielectric", torch.tensor(78.5))


class _Interaction(torch.nn.Module):
    """Parameter container of IN_layer_all_swish_2pass_tokens (GNN_Layers.py:2143-2246): message1/2, lin1/2."""

    def __init__(self, in_channels, out_channels, hidden, hidden_token, n_rbf=20):
        super().__init__()
        self.register_buffer("_FREQUENCIES", math.pi * torch.arange(1, n_rbf + 1, dtype=torch.float32))
        self.message1 = torch.nn.Linear(in_channels + n_rbf, hidden)
        self.message2 = torch.nn.Linear(hidden, hidden)
        self.lin1 = torch.nn.Linear(hidden + hidden, hidden_token)
        self.lin2 = torch.nn.Linear(hidden_token, out_channels)


class GNN3_Multisolvent_embedding(torch.nn.Module):
    """Data-path model (GNN_Models.py:326-498): forward(data) -> (energy [1,1] or [B,1], forces [n,3]).

    `data` needs `.pos [n,3]`, `.atom_features [n,7]`, `.batch [n]` (sorted), `.solvent_model_tensor [n]`,
    `.solvent_dielectric [n]`.  Ragged batches of different molecules (a13 in SURVEY.md §8) are evaluated molecule
    by molecule: graphs sharing atom count and atom_features form one batched engine call."""

    _gb_only = False
    _data_predicate = True     # torch_cluster.radius_graph semantics for the GNN graph (GNN_Models.py:345-357)

    def __init__(self, fraction=0.1, radius=0.6, max_num_neighbors=10000, parameters=None, device=None,
                 jittable=False, unique_radii=DEFAULT_UNIQUE_RADII, hidden=64, num_solvents=39, dropout_rate=0.0,
                 hidden_token=128, scaling_factor=2.0, no_batch=True):
        super().__init__()
        if dropout_rate != 0.0:
            raise NotImplementedError("dropout is a training feature; the inference engine evaluates with p = 0")
        if hidden != 64 or hidden_token != 128:
            raise NotImplementedError("libgnnis is specialised for hidden=64, hidden_token=128 (shipped GNN.pt)")
        self._fraction = float(fraction)
        self._scaling_factor = float(scaling_factor)
        self._gnn_radius = float(radius)
        self._max_num_neighbors = max_num_neighbors
        self._nobatch = bool(no_batch)
        self._unique_radii = list(unique_radii)
        self._device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self._gbparameters = None
        if parameters is not None:
            self._gbparameters = torch.tensor(np.asarray(parameters, dtype=np.float64), dtype=torch.float32)
        d0, m0, _ = gbneck.neck_tables(self._unique_radii)
        self.aggregate_information = _Tables(d0, m0, False)
        self.calculate_energies = _Tables(d0, m0, True)
        self.lin = torch.nn.Linear(1, 1)        # present in the checkpoint, unused by forward (GNN_Models.py:62)
        self.solvent_embedding = torch.nn.Embedding(num_solvents, hidden)
        self.gamma_embedding = torch.nn.Embedding(num_solvents, 1)
        self.interaction1 = _Interaction(3 + 3, hidden, hidden, hidden_token)
        self.interaction2 = _Interaction(hidden + hidden, hidden, hidden, hidden_token)
        self.interaction3 = _Interaction(hidden + hidden, 2, hidden, hidden_token)
        self.register_buffer("_num_solvents", torch.tensor(num_solvents, dtype=torch.int32))
        self.register_buffer("_no_batch", torch.tensor(bool(no_batch), dtype=torch.bool))
        self._engine: Optional[Engine] = None
        self._engine_params = None
        self._weights_dirty = True
        self._rep_cfg = None
        self._rep_dirty = True

    # ------------------------------------------------------------------ engine management
    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._weights_dirty = True
        self.__dict__["_weights_version"] = self.__dict__.get("_weights_version", 0) + 1
        return out

    def _engine_ready(self, parameters=None) -> Engine:
        if not torch.cuda.is_available():
            raise GnnisError("no CUDA device: the gnnis hot path has no CPU fallback")
        dev = self._device if self._device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        params = parameters if parameters is not None else self._gbparameters
        if params is None:
            raise ValueError("model was built without `parameters`")
        p = np.asarray(params.detach().cpu().numpy() if isinstance(params, torch.Tensor) else params, dtype=np.float32)
        if self._engine is None or self._engine_params is None or not np.array_equal(self._engine_params, p):
            if self._engine is not None:
                self._engine.close()
            sd = None if self._gb_only else {k: v.detach().cpu() for k, v in self.state_dict().items()}
            self._engine = Engine(p, sd, device=dev, unique_radii=self._unique_radii, fraction=self._fraction,
                                  scaling_factor=self._scaling_factor, cutoff=0.6 if not self._data_predicate
                                  else self._gnn_radius)
            self._engine_params = p.copy()
            self._weights_dirty = False
            self._rep_dirty = True
        elif self._weights_dirty and not self._gb_only:
            self._engine.load_state_dict({k: v.detach().cpu() for k, v in self.state_dict().items()})
            self._weights_dirty = False
        if self._rep_dirty and self._rep_cfg is not None:
            self._engine.set_replicas(*self._rep_cfg)
            self._rep_dirty = False
        return self._engine

    def _set_replica_config(self, num_reps, solvent_models, dielectrics):
        cfg = (int(num_reps), [int(s) for s in solvent_models], [float(d) for d in dielectrics])
        if cfg != self._rep_cfg:
            self._rep_cfg = cfg
            self._rep_dirty = True

    # ------------------------------------------------------------------ data path
    def forward(self, data):
        """Ragged PyG-style batches are supported by grouping the graphs by molecule (identical atom count and
        atom_features): each group is one batched evaluation on an engine cached for that molecule."""
        pos = data.pos
        n = pos.shape[0]
        batch = getattr(data, "batch", None)
        if batch is None:
            batch = torch.zeros(n, dtype=torch.int64)
        b = batch.detach().cpu().numpy().astype(np.int64)
        if n == 0 or np.any(np.diff(b) < 0):
            raise ValueError("data.batch must be non-empty and sorted (PyG batches are)")
        n_graphs = int(b.max()) + 1
        counts = np.bincount(b, minlength=n_graphs)
        starts = np.concatenate([[0], np.cumsum(counts)])
        feats = data.atom_features.detach().cpu().to(torch.float32).numpy()
        sid = data.solvent_model_tensor.detach().cpu().numpy()
        die = data.solvent_dielectric.detach().cpu().to(torch.float32).numpy()
        groups = {}
        for gi in range(n_graphs):
            lo, hi = int(starts[gi]), int(starts[gi + 1])
            if hi == lo:
                raise ValueError(f"graph {gi} of the batch is empty")
            groups.setdefault(feats[lo:hi].tobytes(), []).append(gi)
        dev = self._device if self._device.type == "cuda" else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else None
        if dev is None:
            raise GnnisError("no CUDA device: the gnnis hot path has no CPU fallback")
        e_graph = torch.empty(n_graphs, dtype=torch.float32, device=dev)
        forces = torch.empty(n, 3, dtype=torch.float32, device=dev)
        pos_dev = pos.detach().to(dev, torch.float32)
        for key, members in groups.items():
            lo0 = int(starts[members[0]])
            N = int(counts[members[0]])
            eng = self._data_engine(key, feats[lo0:lo0 + N], dev)
            idx = torch.as_tensor(np.concatenate([np.arange(starts[g], starts[g + 1]) for g in members]), device=dev)
            R = len(members)
            eng.set_replicas(R, [int(sid[starts[g]]) for g in members], [float(die[starts[g]]) for g in members])
            e_rep, f = eng.energy_force(pos_dev.index_select(0, idx).reshape(R, N, 3), gb_only=self._gb_only,
                                        data_predicate=self._data_predicate)
            e_graph[torch.as_tensor(members, device=dev)] = e_rep
            forces.index_copy_(0, idx, f.reshape(R * N, 3))
        if self._nobatch:
            energy = e_graph.sum().reshape(1, 1)
        else:
            energy = e_graph.reshape(n_graphs, 1)
        return energy, forces

    def _data_engine(self, key, params, dev) -> Engine:
        """One engine per distinct molecule of the data path (small LRU keyed by the atom_features bytes)."""
        cache = self.__dict__.setdefault("_data_engines", {})
        sd = None if self._gb_only else {k: v.detach().cpu() for k, v in self.state_dict().items()}
        ent = cache.get(key)
        if ent is None:
            if len(cache) >= 16:
                cache.pop(next(iter(cache))).close()
            ent = Engine(params, sd, device=dev, unique_radii=self._unique_radii, fraction=self._fraction,
                         scaling_factor=self._scaling_factor, cutoff=self._gnn_radius)
            cache[key] = ent
            ent._weights_version = self.__dict__.get("_weights_version", 0)
        elif ent._weights_version != self.__dict__.get("_weights_version", 0) and sd is not None:
            ent.load_state_dict(sd)
            ent._weights_version = self.__dict__.get("_weights_version", 0)
        return ent

    # ------------------------------------------------------------------ native batched call
    def energy_force(self, positions: torch.Tensor, delta: bool = False):
        """positions [R,N,3] (or [R*N,3]) -> (E[R], F[R,N,3]) on the GPU; no autograd graph."""
        eng = self._engine_ready()
        return eng.energy_force(positions, delta=delta, gb_only=self._gb_only, data_predicate=False)


class GNN_GBNeck_2(GNN3_Multisolvent_embedding):
    """Plain GB-Neck2 (GNN_Models.py:173-229): Born radii + GB energy, fixed solvent dielectric 78.5."""

    _gb_only = True

    def forward(self, data):
        n = data.pos.shape[0]
        if not hasattr(data, "solvent_dielectric"):
            data.solvent_dielectric = torch.full((n,), 78.5)
        if not hasattr(data, "solvent_model_tensor"):
            data.solvent_model_tensor = torch.zeros(n, dtype=torch.int64)
        return super().forward(data)


class GNN3_Multisolvent_embedding_run_multiple(GNN3_Multisolvent_embedding):
    """Run model consumed by TorchForce (GNN_Models.py:700-944).

    forward(positions [R*N,3] nm, solvent_model scalar, solvent_dielectric scalar) -> scalar kJ/mol.
    `solvent_model == -1` uses the per-replica solvents of `set_num_reps`; any other value evaluates every
    replica in that solvent with the given dielectric (the reference's single-replica override, :780-786)."""

    _data_predicate = False     # eps-distance `r < 0.6` compaction of the all-pairs list (GNN_Models.py:795)
    _delta = False

    def __init__(self, fraction=0.1, radius=0.6, max_num_neighbors=10000, parameters=None, device=None,
                 jittable=False, num_reps=1, gbneck_radius=10.0, unique_radii=DEFAULT_UNIQUE_RADII, hidden=64,
                 solvent_dielectric=[78.5], num_solvents=1, solvent_models=[0], hidden_token=128,
                 scaling_factor=2.0):
        super().__init__(fraction=fraction, radius=radius, max_num_neighbors=max_num_neighbors,
                         parameters=parameters, device=device, jittable=jittable, unique_radii=unique_radii,
                         hidden=hidden, num_solvents=num_solvents, hidden_token=hidden_token,
                         scaling_factor=scaling_factor)
        if parameters is None:
            raise ValueError("run models need the [N,7] `parameters` table")
        self.set_num_reps(num_reps, solvent_models, solvent_dielectric)

    def set_num_reps(self, num_reps=1, solvent_models: Sequence[int] = (0,), solvent_dielectric: Sequence[float] = (78.5,)):
        """GNN_Models.py:741-769.  O(1): no [2, R*N*(N-1)] index is materialised."""
        if len(solvent_models) < num_reps or len(solvent_dielectric) < num_reps:
            raise IndexError("need one solvent model and one dielectric per replica")
        self._num_reps = int(num_reps)
        self._solvent_models = [int(s) for s in solvent_models[:num_reps]]
        self._solvent_dielectrics = [float(d) for d in solvent_dielectric[:num_reps]]
        self._set_replica_config(self._num_reps, self._solvent_models, self._solvent_dielectrics)
        return 0

    def _select_solvents(self, solvent_model, solvent_dielectric):
        # solvent_model != -1 puts every replica into that solvent for THIS call (GNN_Models.py:780-786).  Intentional
        # divergence: the reference overwrites its per-replica buffers, so its override sticks for later calls with -1;
        # here a later call with -1 evaluates the set_num_reps configuration again.
        sm = int(solvent_model) if solvent_model is not None else -1
        if sm != -1:
            self._set_replica_config(self._num_reps, [sm] * self._num_reps, [float(solvent_dielectric)] * self._num_reps)
        else:
            self._set_replica_config(self._num_reps, self._solvent_models, self._solvent_dielectrics)

    def _energy_per_replica(self, positions, solvent_model, solvent_dielectric):
        self._select_solvents(solvent_model, solvent_dielectric)
        n = self._num_reps * self._gbparameters.shape[0]
        if positions.shape[0] * (positions.shape[1] if positions.dim() == 3 else 1) != n:
            raise ValueError(f"positions must hold {n} atoms ({self._num_reps} replicas)")
        return _EnergyFunction.apply(positions, self, self._delta)

    def forward(self, positions, solvent_model=-1, solvent_dielectric=78.5):
        return self._energy_per_replica(positions, solvent_model, solvent_dielectric).sum()

    def energy_force(self, positions: torch.Tensor, solvent_model=-1, solvent_dielectric=78.5):
        self._select_solvents(solvent_model, solvent_dielectric)
        eng = self._engine_ready()
        return eng.energy_force(positions, delta=self._delta)

    # batched drivers that keep every replica resident on the GPU
    def minimize(self, positions, grad_tol=1.0, max_iter=500, history=8):
        self._select_solvents(-1, 78.5)
        return self._engine_ready().minimize(positions, grad_tol=grad_tol, max_iter=max_iter, history=history,
                                             delta=self._delta)

    def langevin(self, positions, velocities, masses, n_steps, **kw):
        self._select_solvents(-1, 78.5)
        return self._engine_ready().langevin(positions, velocities, masses, n_steps, delta=self._delta, **kw)

    def scripted(self, outputs_forces: bool = False):
        """What the reference obtains from `torch.jit.optimize_for_inference(torch.jit.script(run_model.eval()))`
        (`helper_functions.py:777-782`): a TorchScript module for `openmmtorch.TorchForce`, here a thin scripted
        wrapper over `torch.ops.gnnis.*` (csrc/torch_op.cpp) bound to this model's resident engine.  Keep the model
        alive while the scripted module (or a file saved from it) is in use."""
        from . import torch_op
        self._select_solvents(-1, 78.5)
        return torch_op.script_engine(self._engine_ready(), outputs_forces=outputs_forces,
                                      flags=_lib.FLAG_DELTA if self._delta else 0)


class GNN3_Multisolvent_embedding_run_multiple_Delta(GNN3_Multisolvent_embedding_run_multiple):
    """GNN - GBNeck2 difference + SA (GNN_Models.py:947-1065): forward(positions) only."""

    _delta = True

    def forward(self, positions):
        return self._energy_per_replica(positions, -1, 78.5).sum()


class GNN3_Multisolvent_embedding_run_multiple_split(GNN3_Multisolvent_embedding_run_multiple):
    """Per-atom polar / apolar energies (GNN_Models.py:1068-1083): -> (energies [n,1], sa_energies [n,1])."""

    def forward(self, positions, solvent_model=-1, solvent_dielectric=78.5):
        self._select_solvents(solvent_model, solvent_dielectric)
        eng = self._engine_ready()
        eng.energy_force(positions, forces=False)
        e, sa = eng.atom_energies()
        return e.reshape(-1, 1), sa.reshape(-1, 1)


_KEEP_ALIVE: list = []


def create_gnn_model(parameters, state_dict, num_reps=1, solvent_models: List[int] = (0,),
                     solvent_dielectric: List[float] = (78.5,), cls=GNN3_Multisolvent_embedding_run_multiple,
                     device=None, fraction=0.1, scaling_factor=2.0, jit=False):
    """The model-building half of Simulation/helper_functions.py:create_gnn_model / create_gnn_sim (:750-790):
    instantiate a run model for `num_reps` replicas, load the checkpoint's state dict, freeze it.  `jit=True` returns
    the TorchScript form (`model.scripted()`), as the reference does for TorchForce; the eager model that owns the
    engine is kept alive by this module."""
    num_solvents = int(state_dict["solvent_embedding.weight"].shape[0])
    model = cls(fraction=fraction, radius=0.6, max_num_neighbors=10000, parameters=parameters, device=device,
                num_reps=num_reps, unique_radii=DEFAULT_UNIQUE_RADII, hidden=64,
                solvent_dielectric=list(solvent_dielectric), num_solvents=num_solvents,
                solvent_models=list(solvent_models), hidden_token=128, scaling_factor=scaling_factor)
    model.load_state_dict(state_dict)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    if jit:
        scripted = model.scripted()
        _KEEP_ALIVE.append(model)      # the scripted module only holds the engine's handle
        return scripted
    return model
