"""Bond inference: `src/molecule_builder.py::build_xae_molecule` / `get_bond_order` (lines 44-102) for whole batches on the
GPU (`dl_bond_orders`) instead of an O(n^2) Python loop with one `.item()` per atom pair.
RDKit molecule construction (`build_molecule`, molecule_builder.py:29-42) stays with the caller: RDKit is not part of
the path.
"""
import torch

from . import _native
from .output import GEOM_IDX2ATOM, IDX2ATOM

# Bond lengths in pm for the atom types of the one-hot encodings (src/const.py:66-146, tables BONDS_1/2/3), keyed by the
# pair ORDERED BY TYPE INDEX -- `sorted([atom_types[i], atom_types[j]])` (molecule_builder.py:66) -- which is the only
# direction the reference ever looks up; a missing key means "no typical bond length" (get_bond_order returns 0).
SINGLE = {('C', 'C'): 154, ('C', 'O'): 143, ('C', 'N'): 147, ('C', 'F'): 135, ('C', 'S'): 182, ('C', 'Cl'): 177,
          ('C', 'Br'): 194, ('C', 'I'): 214, ('C', 'P'): 184, ('O', 'O'): 148, ('O', 'N'): 140, ('O', 'F'): 142,
          ('O', 'S'): 151, ('O', 'Cl'): 164, ('O', 'Br'): 172, ('O', 'I'): 194, ('O', 'P'): 163, ('N', 'N'): 145,
          ('N', 'F'): 136, ('N', 'S'): 168, ('N', 'Cl'): 175, ('N', 'Br'): 214, ('N', 'I'): 222, ('N', 'P'): 177,
          ('F', 'F'): 142, ('F', 'S'): 158, ('F', 'Cl'): 166, ('F', 'Br'): 178, ('F', 'I'): 187, ('F', 'P'): 156,
          ('S', 'S'): 204, ('S', 'Cl'): 207, ('S', 'Br'): 225, ('S', 'I'): 234, ('S', 'P'): 210, ('Cl', 'Cl'): 199,
          ('Cl', 'Br'): 214, ('Cl', 'P'): 203, ('Br', 'Br'): 228, ('Br', 'P'): 222, ('I', 'I'): 266, ('P', 'P'): 221}
DOUBLE = {('C', 'C'): 134, ('C', 'O'): 120, ('C', 'N'): 129, ('C', 'S'): 160, ('O', 'O'): 121, ('O', 'N'): 121,
          ('O', 'P'): 150, ('N', 'N'): 125, ('S', 'P'): 186}
TRIPLE = {('C', 'C'): 120, ('C', 'O'): 113, ('C', 'N'): 116, ('N', 'N'): 110}
MARGINS_EDM = [10, 5, 2]                                                                   # src/const.py:180


def threshold_tables(is_geom, margins=MARGINS_EDM):
    """(T,T) fp32 thresholds [min type][max type] = bond length + margin in pm; -1 where the pair is absent."""
    idx2atom = GEOM_IDX2ATOM if is_geom else IDX2ATOM
    T = len(idx2atom)
    out = []
    for table, margin in ((SINGLE, margins[0]), (DOUBLE, margins[1]), (TRIPLE, margins[2])):
        t = torch.full((T, T), -1.0)
        for a in range(T):
            for c in range(a, T):
                v = table.get((idx2atom[a], idx2atom[c]))
                if v is not None:
                    t[a, c] = float(v + margin)
        out.append(t)
    return out


@torch.no_grad()
def bond_orders(one_hot, x, node_mask, is_geom, margins=MARGINS_EDM):
    """Batched E of build_xae_molecule: (B,N,N) int8 on the inputs' device, E[b,i,j] (i > j) = bond order, else 0.
    `x` may be (B,N,3) or chain[0]-style (B,N,3+F); atom types are argmax(one_hot) (molecule_builder.py:20)."""
    dev = x.device
    if dev.type != 'cuda':
        raise RuntimeError("bond_orders runs on the GPU (no CPU fallback); move the tensors to the device")
    B, N = x.shape[:2]
    xs = x.float().contiguous()
    types = torch.argmax(one_hot, dim=2).to(torch.int32).contiguous()
    nm = (node_mask.reshape(B, N) != 0).to(torch.int8).contiguous()
    t1, t2, t3 = [t.to(dev).contiguous() for t in threshold_tables(is_geom, margins)]
    E = torch.empty((B, N, N), dtype=torch.int8, device=dev)
    lib = _native.load_library()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _native.check(lib.dl_bond_orders(B, N, t1.shape[0], xs.data_ptr(), xs.shape[2], types.data_ptr(), nm.data_ptr(),
                                         t1.data_ptr(), t2.data_ptr(), t3.data_ptr(), E.data_ptr(), st), "dl_bond_orders")
    return E


def build_xae_molecule(positions, atom_types, is_geom, margins=MARGINS_EDM):
    """Reference signature (molecule_builder.py:44): one molecule, positions (n,3) already masked, atom_types (n,).
    Returns (X, A, E) with A bool, E int32, lower-triangular ("the graph should be DIRECTED")."""
    n = positions.shape[0]
    T = len(GEOM_IDX2ATOM if is_geom else IDX2ATOM)
    one_hot = torch.nn.functional.one_hot(atom_types.long(), T).unsqueeze(0)
    E = bond_orders(one_hot.to(positions.device), positions.unsqueeze(0), torch.ones((1, n), device=positions.device),
                    is_geom, margins)[0].to(torch.int)
    return atom_types, E != 0, E
