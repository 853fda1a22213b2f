"""difflinker_b200 -- B200-native (sm_100a) implementation of DiffLinker's denoising hot path.

Host-side mirror of the reference interface for the path (same names, arguments, errors):
    Dynamics, DynamicsWithPockets  (src/egnn.py)      -> egnn.py
    EDM                            (src/edm.py)       -> edm.py
    DDPM.sample_chain              (src/lightning.py) -> ddpm.py
    collate, create_templates_for_linker_generation (src/datasets.py) -> batching.py
    SizeGNN, SizeClassifier        (src/linker_size*.py) -> linker_size.py
    save_xyz_file, restore_frame   (src/visualizer.py, generate.py:163-171) -> output.py
All arithmetic of the path runs in libdifflinker_b200.so (csrc/, C-ABI in include/difflinker_b200.h).
"""
from .batching import collate, create_templates_for_linker_generation  # noqa: F401
from .ddpm import DDPM, accelerate  # noqa: F401
from .edm import EDM, InpaintingEDM  # noqa: F401
from .egnn import Dynamics, DynamicsWithPockets  # noqa: F401
from .linker_size import SizeClassifier, SizeGNN  # noqa: F401
from .noise import PredefinedNoiseSchedule  # noqa: F401
from .utils import FoundNaNException  # noqa: F401

__version__ = "0.1.0"
