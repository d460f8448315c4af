"""gnnimplicitsolvent_b200 -- B200-native energy / force hot path of rinikerlab/GNNImplicitSolvent.

Package surface of the reference (`/root/reference/__init__.py:3-8`):
    from gnnimplicitsolvent_b200 import minimize_mol, create_gnn_model, calculate_entropy, \
        GNN3_Multisolvent_embedding_run_multiple_Delta
plus the simulator classes (`Simulator`, `Multi_simulator`) and the light `Molecule` record that stands in for an
RDKit molecule where RDKit / OpenFF are absent.  Everything numerical runs in libgnnis.so (include/gnnis.h); importing
the package needs no GPU, evaluating anything does."""
from .helper_functions import (  # noqa: F401
    MODEL_PATH,
    SOLVENT_DICT,
    calculate_entropy,
    create_gnn_model,
    get_gnn_sim,
    minimize_mol,
)
from .GNN_Models import (  # noqa: F401
    GNN3_Multisolvent_embedding,
    GNN3_Multisolvent_embedding_run_multiple,
    GNN3_Multisolvent_embedding_run_multiple_Delta,
    GNN3_Multisolvent_embedding_run_multiple_split,
)
from .Simulator import Molecule, Multi_simulator, Simulator  # noqa: F401

__all__ = ["minimize_mol", "create_gnn_model", "calculate_entropy", "GNN3_Multisolvent_embedding_run_multiple_Delta",
           "GNN3_Multisolvent_embedding", "GNN3_Multisolvent_embedding_run_multiple",
           "GNN3_Multisolvent_embedding_run_multiple_split", "Simulator", "Multi_simulator", "Molecule", "get_gnn_sim",
           "SOLVENT_DICT", "MODEL_PATH"]
