"""Runs the reference's OWN entropy functions (Simulation/helper_functions.py:1551-1885) on duck-typed inputs
(TEST INFRASTRUCTURE ONLY; needs /root/reference, i.e. the build container).

helper_functions.py cannot be imported (OpenMM, RDKit, mdtraj at import time), but the functions on this path use
numpy only plus a handful of OpenMM unit names.  Their source is cut out of the file by AST and exec'd in a
namespace that provides those names and a stand-in `simulation` object whose forces come from a Python callable."""
import ast
import os

import numpy as np

from . import ref_harness

_NAMES = ("build_mass_weighted_hessian", "normal_modes", "calculate_entropy_from_frequencies",
          "calculate_rotational_entropy")


class _Unit:
    def __init__(self, name):
        self.name = name

    def __truediv__(self, other):
        return _Unit(f"{self.name}/{getattr(other, 'name', other)}")

    def __mul__(self, other):
        return _Unit(f"{self.name}*{getattr(other, 'name', other)}")

    __rmul__ = __mul__


class _Quantity(float):
    """Quantity(delta, unit=nanometer): behaves as the bare number (positions are kept in nm here)."""

    def __new__(cls, value, unit=None):
        return float.__new__(cls, value)


class _Array(np.ndarray):
    def value_in_unit(self, unit):
        if getattr(unit, "name", "") == "meter":
            return np.asarray(self) * 1e-9
        return np.asarray(self)


class _Mass:
    def __init__(self, m):
        self._m = m

    def value_in_unit(self, unit):
        return self._m


class _Atom:
    def __init__(self, index, mass):
        self.index = index
        self.element = type("E", (), {"mass": _Mass(mass)})()


class _State:
    def __init__(self, pos, forces):
        self._pos, self._forces = pos, forces

    def getPositions(self, asNumpy=True):
        return self._pos.copy().view(_Array)

    def getForces(self, asNumpy=True):
        return self._forces.view(_Array)


class _Context:
    def __init__(self, pos, force_fn):
        self._pos = np.array(pos, dtype=np.float64)
        self._force_fn = force_fn

    def setPositions(self, pos):
        self._pos = np.array(pos, dtype=np.float64)

    def getState(self, getPositions=False, getForces=False):
        f = self._force_fn(self._pos) if getForces else None
        return _State(self._pos, f)


class FakeSimulation:
    """context + topology with what the reference functions touch."""

    def __init__(self, positions_nm, masses, force_fn):
        self.context = _Context(positions_nm, force_fn)
        atoms = [_Atom(i, float(m)) for i, m in enumerate(masses)]
        self.topology = type("T", (), {"atoms": lambda self_: iter(atoms)})()


def load_reference_functions():
    root = ref_harness.reference_root()
    if root is None:
        raise ImportError("reference tree not available")
    path = os.path.join(root, "Simulation", "helper_functions.py")
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "Quantity": _Quantity}
    for u in ("nanometer", "dalton", "kilojoule", "mole", "meter"):
        ns[u] = _Unit(u)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in _NAMES:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return {k: ns[k] for k in _NAMES}
