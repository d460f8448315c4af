"""Import harness for the UNMODIFIED reference model files (TEST INFRASTRUCTURE ONLY).

Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this. It puts oracle/ref_stubs (stand-ins for torch_geometric,
torch_cluster, torch_scatter, torch_sparse -- none are installed) and the reference root on
sys.path, then imports MachineLearning.GNN_Models exactly as it lies on disk
(reference: MachineLearning/GNN_Models.py:5-31 import block).

The reference root is looked up at  $GNNIS_REFERENCE_ROOT, baseline/_ref, /root/reference.
It does not exist on the GPU box: callers must handle `reference_root() is None`.
"""
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "ref_stubs")
_CANDIDATES = [
    os.environ.get("GNNIS_REFERENCE_ROOT", ""),
    os.path.join(os.path.dirname(_HERE), "baseline", "_ref"),
    "/root/reference",
]


def reference_root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "MachineLearning", "GNN_Models.py")):
            return c
    return None


def import_reference_models():
    """Returns the reference's MachineLearning.GNN_Models module (or raises ImportError)."""
    root = reference_root()
    if root is None:
        raise ImportError("reference tree not available (expected on the build box only)")
    for p in (_STUBS, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import MachineLearning.GNN_Models as M  # noqa: N812
    return M


def checkpoint_path():
    root = reference_root()
    if root is None:
        return None
    p = os.path.join(root, "MachineLearning", "trained_models", "GNN.pt")
    return p if os.path.isfile(p) else None


def build_reference_run_model(parameters, state_dict=None, num_reps=1, solvent_models=(0,),
                              solvent_dielectric=(78.5,), cls="GNN3_Multisolvent_embedding_run_multiple",
                              num_solvents=42, hidden=64, hidden_token=128, fraction=0.1,
                              scaling_factor=2.0):
    """Instantiates the reference run model on CPU the way create_gnn_sim does
    (Simulation/helper_functions.py:764-782), with weight gradients frozen."""
    import torch
    M = import_reference_models()
    model = getattr(M, cls)(
        fraction=fraction, radius=0.6, max_num_neighbors=10000, parameters=parameters,
        device="cpu", jittable=False, num_reps=1, gbneck_radius=10.0,
        unique_radii=M.DEFAULT_UNIQUE_RADII, hidden=hidden, solvent_dielectric=[78.5],
        num_solvents=num_solvents, solvent_models=[0], hidden_token=hidden_token,
        scaling_factor=scaling_factor)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    model.set_num_reps(num_reps, list(solvent_models), list(solvent_dielectric))
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model
