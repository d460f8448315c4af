"""TorchScript route to the engine: `libgnnis_torch.so` registers `torch.ops.gnnis.energy_force` and
`torch.ops.gnnis.energy` (csrc/torch_op.cpp), and the two modules below are what `torch.jit.script` / `torch.jit.save`
can serialise for `openmmtorch.TorchForce(file)` -- the reference does
`torch.jit.optimize_for_inference(torch.jit.script(run_model.eval()))` and saves it (`helper_functions.py:777-785`).

The scripted module stores the gnnis handle of a live `Engine` as an int attribute, so the file is only meaningful
inside the process that owns that engine (OpenMM runs in-process with the Python that built the model, which is the
reference's flow too); the process that loads the file must have called `torch_op.load()` first.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgnnis_torch.so")
_loaded = False


def load() -> str:
    """torch.ops.load_library of the operator shim; raises if it was not built (python -m gnnimplicitsolvent_b200.build)."""
    global _loaded
    if not _loaded:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m gnnimplicitsolvent_b200.build`")
        torch.ops.load_library(LIB_PATH)
        _loaded = True
    return LIB_PATH


class ScriptableEnergy(torch.nn.Module):
    """forward(positions[, solvent_model, solvent_dielectric]) -> scalar energy (kJ/mol) with d/dpositions = -forces:
    TorchForce's default contract (`energy.backward()`, then `positions.grad`).  The two optional scalars are the
    global parameters TorchForce appends (`helper_functions.py:784-785`).  The solvent of every replica is fixed by
    `Engine.set_replicas` when the module is scripted; the reference's run model switches ALL replicas to another
    solvent when `solvent_model != -1` (GNN_Models.py:780-786) -- the scripted module cannot, so it raises instead of
    returning energies for the wrong solvent (use the eager model's `forward(positions, solvent_model, dielectric)`
    or script a model that was set up for that solvent)."""

    def __init__(self, handle: int, n_rep: int, flags: int = 0):
        super().__init__()
        self.handle = int(handle)
        self.n_rep = int(n_rep)
        self.flags = int(flags)

    def forward(self, positions: torch.Tensor, solvent_model: Optional[torch.Tensor] = None,
                solvent_dielectric: Optional[torch.Tensor] = None) -> torch.Tensor:
        if solvent_model is not None:
            if int(solvent_model.item()) != -1:
                raise RuntimeError("gnnis scripted module: the solvent_model override (!= -1) is not supported; "
                                   "the solvents were fixed by set_num_reps when the module was scripted")
        return torch.ops.gnnis.energy(positions, self.handle, self.n_rep, self.flags).sum()


class ScriptableEnergyForces(torch.nn.Module):
    """forward(positions, ...) -> (scalar energy, forces like positions): TorchForce with `setOutputsForces(True)`."""

    def __init__(self, handle: int, n_rep: int, flags: int = 0):
        super().__init__()
        self.handle = int(handle)
        self.n_rep = int(n_rep)
        self.flags = int(flags)

    def forward(self, positions: torch.Tensor, solvent_model: Optional[torch.Tensor] = None,
                solvent_dielectric: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if solvent_model is not None:
            if int(solvent_model.item()) != -1:
                raise RuntimeError("gnnis scripted module: the solvent_model override (!= -1) is not supported; "
                                   "the solvents were fixed by set_num_reps when the module was scripted")
        e, f = torch.ops.gnnis.energy_force(positions, self.handle, self.n_rep, self.flags)
        return e.sum(), f


def script_engine(engine, outputs_forces: bool = False, flags: int = 0) -> torch.jit.ScriptModule:
    """The scripted module for `engine` (keep `engine` alive for as long as the module is used)."""
    load()
    handle = int(engine._h.value) if hasattr(engine._h, "value") else int(engine._h)
    cls = ScriptableEnergyForces if outputs_forces else ScriptableEnergy
    return torch.jit.script(cls(handle, engine.n_rep, flags).eval())
