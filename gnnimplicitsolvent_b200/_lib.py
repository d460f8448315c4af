"""ctypes binding of libgnnis.so (C ABI: include/gnnis.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback and no PyTorch fallback on the product path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgnnis.so")

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class GnnisWeights(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("hidden_token", C.c_int32), ("num_solvents", C.c_int32), ("n_rbf", C.c_int32),
        ("solvent_embedding", c_float_p), ("gamma_embedding", c_float_p),
        ("message1_w", c_float_p * 3), ("message1_b", c_float_p * 3),
        ("message2_w", c_float_p * 3), ("message2_b", c_float_p * 3),
        ("lin1_w", c_float_p * 3), ("lin1_b", c_float_p * 3),
        ("lin2_w", c_float_p * 3), ("lin2_b", c_float_p * 3),
    ]


class GnnisForcefield(C.Structure):
    _fields_ = [
        ("n_bonds", C.c_int32), ("bond_idx", c_int32_p), ("bond_r0", c_float_p), ("bond_k", c_float_p),
        ("n_angles", C.c_int32), ("angle_idx", c_int32_p), ("angle_theta0", c_float_p), ("angle_k", c_float_p),
        ("n_torsions", C.c_int32), ("torsion_idx", c_int32_p), ("torsion_periodicity", c_int32_p),
        ("torsion_phase", c_float_p), ("torsion_k", c_float_p),
        ("charge", c_float_p), ("sigma", c_float_p), ("epsilon", c_float_p),
        ("n_exceptions", C.c_int32), ("exception_idx", c_int32_p), ("exception_chargeprod", c_float_p),
        ("exception_sigma", c_float_p), ("exception_epsilon", c_float_p),
    ]


# every symbol include/gnnis.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gnnis_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "gnnis_destroy": (C.c_int, [C.c_void_p]),
    "gnnis_last_error": (C.c_char_p, [C.c_void_p]),
    "gnnis_load_weights": (C.c_int, [C.c_void_p, C.POINTER(GnnisWeights)]),
    "gnnis_set_system": (C.c_int, [C.c_void_p, C.c_int32, c_float_p, c_float_p, c_float_p, C.c_int32]),
    "gnnis_set_replicas": (C.c_int, [C.c_void_p, C.c_int32, c_int32_p, c_float_p]),
    "gnnis_set_model_constants": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "gnnis_build_neighbors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, c_int64_p, C.c_void_p]),
    "gnnis_energy_force": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "gnnis_energy_force_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "gnnis_energy_force_host_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "gnnis_energy_force_host_wait": (C.c_int, [C.c_void_p]),
    "gnnis_get_atom_energies": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnnis_get_born_radii": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnnis_last_edge_count": (C.c_int, [C.c_void_p, c_int64_p, C.c_void_p]),
    "gnnis_set_forcefield": (C.c_int, [C.c_void_p, C.POINTER(GnnisForcefield)]),
    "gnnis_set_constraints": (C.c_int, [C.c_void_p, C.c_int32, c_int32_p, c_float_p, C.c_float]),
    "gnnis_set_bonds": (C.c_int, [C.c_void_p, C.c_int32, c_int32_p, c_float_p, c_float_p]),
    "gnnis_minimize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_int32_p, C.c_float, C.c_int32,
                                 C.c_int32, C.c_uint32, C.c_void_p]),
    "gnnis_langevin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_float_p, C.c_float, C.c_float, C.c_float,
                                 C.c_int32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]),
    "gnnis_umma_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "gnnis_launch_count": (C.c_int64, [C.c_void_p]),
    "gnnis_describe": (C.c_char_p, [C.c_void_p]),
    "gnnis_minimize_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnnis_profile_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                     c_float_p, C.c_int32]),
    "gnnis_num_stages": (C.c_int32, []),
    "gnnis_stage_name": (C.c_char_p, [C.c_int32]),
}

FLAG_DELTA, FLAG_NO_FORCES, FLAG_GB_ONLY, FLAG_MLP_BF16, FLAG_MLP_CUDA_CORES = 0x1, 0x2, 0x4, 0x8, 0x10
FLAG_MLP_TF32X1 = 0x80
FLAG_MLP_FP16 = 0x100
FLAG_DATA_PREDICATE = 0x20
FLAG_VACUUM_ONLY = 0x40
FLAG_CELL_LIST, FLAG_ALL_PAIRS = 0x200, 0x400
PREDICATE_CELL_LIST, PREDICATE_ALL_PAIRS = 0x10, 0x20

_lib = None


def load():
    """Loads libgnnis.so and declares all prototypes. Raises RuntimeError if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m gnnimplicitsolvent_b200.build` "
            "(nvcc, sm_100a). There is no CPU / PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError => the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
