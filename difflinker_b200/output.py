"""Output stage: what every generation script does right after `sample_chain`.

    restore_frame   generate.py:163-171 (generate_with_pocket.py:272-280, sample.py:164-171)
    save_xyz_file   src/visualizer.py:14-31 -- same signature; the text of the whole batch is produced by one native call
                    (`dl_format_xyz`) instead of a Python loop with one `.item()` per atom
"""
import ctypes as C
import os

import torch

from . import _native

# src/const.py:15,31
IDX2ATOM = {0: 'C', 1: 'O', 2: 'N', 3: 'F', 4: 'S', 5: 'Cl', 6: 'Br', 7: 'I'}
GEOM_IDX2ATOM = {0: 'C', 1: 'O', 2: 'N', 3: 'F', 4: 'S', 5: 'Cl', 6: 'Br', 7: 'I', 8: 'P'}


def restore_frame(chain0, positions, com_mask, node_mask):
    """In place on a CUDA tensor: chain0[..., :3] += mean(positions * com_mask over atoms) * node_mask
    (generate.py:165-171). `chain0` is (B,N,3) or (B,N,3+F) -- e.g. `chain[0]` straight from `sample_chain`.
    `positions` / `com_mask` may have another padded length than `chain0` / `node_mask`: generate.py passes the INPUT
    batch's positions and masks while the chain has the template's length (sampled linker sizes)."""
    if not chain0.is_cuda:
        raise RuntimeError("restore_frame runs on the GPU (no CPU fallback); move the tensors to the device")
    if chain0.dtype != torch.float32 or not chain0.is_contiguous():
        raise ValueError("chain0 must be a contiguous fp32 tensor")
    B, N, xd = chain0.shape
    dev = chain0.device
    pos = positions.to(device=dev, dtype=torch.float32).contiguous()
    if pos.dim() != 3 or pos.shape[0] != B or pos.shape[2] != 3:
        raise ValueError(f"positions must be (B, N_pos, 3), got {tuple(pos.shape)}")
    n_pos = pos.shape[1]
    cm = com_mask.to(device=dev, dtype=torch.float32).reshape(B, n_pos).contiguous()
    nm = node_mask.to(device=dev).reshape(B, N).to(torch.int8).contiguous()
    lib = _native.load_library()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _native.check(lib.dl_restore_frame2(B, N, n_pos, xd, chain0.data_ptr(), pos.data_ptr(), cm.data_ptr(),
                                            nm.data_ptr(), st), "dl_restore_frame2")
    return chain0


def format_xyz(one_hot, positions, node_mask, is_geom):
    """The text of `save_xyz_file` for every molecule of the batch: list of B `str`."""
    idx2atom = GEOM_IDX2ATOM if is_geom else IDX2ATOM
    B, N, F = one_hot.shape
    if F > len(idx2atom):
        raise KeyError(F - 1)                       # the reference fails with KeyError on idx2atom[atom]
    oh = one_hot.detach().to(device='cpu', dtype=torch.float32).contiguous()
    pos = positions.detach().to(device='cpu', dtype=torch.float32).contiguous()
    nm = (node_mask.detach().to('cpu').reshape(B, N) != 0).to(torch.int8).contiguous()
    syms = (C.c_char_p * len(idx2atom))(*[idx2atom[i].encode() for i in range(len(idx2atom))])
    offsets = (C.c_int64 * (B + 1))()
    cap = B * (N * 64 + 16)
    lib = _native.load_library()
    while True:
        buf = C.create_string_buffer(cap)
        need = lib.dl_format_xyz(B, N, F, pos.data_ptr(), pos.shape[2], oh.data_ptr(), F, nm.data_ptr(), syms,
                                 len(idx2atom), C.cast(buf, C.c_void_p), cap, C.cast(offsets, C.c_void_p))
        if need < 0:
            raise _native.NativeError(f"dl_format_xyz failed with status {need}")
        if need <= cap:
            break
        cap = int(need)
    raw = buf.raw
    return [raw[offsets[b]:offsets[b + 1]].decode() for b in range(B)]


def save_xyz_file(path, one_hot, positions, node_mask, names, is_geom, suffix=''):
    """visualizer.save_xyz_file (visualizer.py:14-31): one `<name>_<suffix>.xyz` per molecule."""
    texts = format_xyz(one_hot, positions, node_mask, is_geom)
    for name, text in zip(names, texts):
        with open(os.path.join(path, f'{name}_{suffix}.xyz'), "w") as f:
            f.write(text)
