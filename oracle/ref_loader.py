"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (igashov/DiffLinker).

Only usable in the build container, where /root/reference exists (it does NOT exist on the GPU
box).  Used by oracle/make_golden.py to (1) pin the oracle restatement in oracle/difflinker_oracle.py
against the live reference code and (2) generate the golden vectors under tests/golden/.

The reference imports rdkit / pytorch_lightning / imageio / matplotlib / Bio at module scope; none of
them is installed here and none is touched on the sampling hot path, so they are stubbed in
sys.modules before import.  Nothing from the reference is copied: the modules are
imported from where they lie.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("DIFFLINKER_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "egnn.py"))


def _install_stubs():
    import torch.nn as nn

    for name in [
        "rdkit", "rdkit.Chem", "rdkit.Geometry", "rdkit.Chem.AllChem", "rdkit.Chem.MolStandardize",
        "rdkit.Chem.rdMolDescriptors", "rdkit.Chem.rdShapeHelpers", "rdkit.Chem.FeatMaps",
        "rdkit.Chem.FeatMaps.FeatMaps", "rdkit.RDConfig", "rdkit.six", "rdkit.six.moves", "rdkit.rdBase",
        "rdkit.RDLogger", "rdkit.Chem.Descriptors", "rdkit.Chem.QED", "rdkit.Chem.Crippen",
        "rdkit.Chem.Lipinski", "rdkit.DataStructs", "rdkit.Chem.rdMolAlign", "rdkit.Chem.Draw",
        "imageio", "matplotlib", "matplotlib.pyplot", "Bio", "Bio.PDB", "wandb", "networkx", "openbabel",
    ]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()

    if "pytorch_lightning" not in sys.modules:
        try:
            import pytorch_lightning  # noqa: F401
        except Exception:
            pl = types.ModuleType("pytorch_lightning")

            class LightningModule(nn.Module):
                def save_hyperparameters(self, *a, **k):
                    pass

                def log(self, *a, **k):
                    pass

            pl.LightningModule = LightningModule
            pl.Trainer = MagicMock()
            pl.callbacks = MagicMock()
            pl.loggers = MagicMock()
            sys.modules["pytorch_lightning"] = pl
            sys.modules["pytorch_lightning.callbacks"] = pl.callbacks
            sys.modules["pytorch_lightning.loggers"] = pl.loggers


_cache = {}


def load_reference():
    """Returns a namespace with the reference's own modules: egnn, edm, noise, utils, datasets, lightning."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT} (only present in the build container)")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    for m in ["utils", "noise", "egnn", "edm", "const", "datasets", "lightning"]:
        setattr(ns, m, importlib.import_module(f"src.{m}"))
    _cache["ns"] = ns
    return ns
