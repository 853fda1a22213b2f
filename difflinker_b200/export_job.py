"""Serialise one sampling job for a caller that has no Python: the engine configuration, the weights under the reference's
state_dict names, the prepared inputs of `EDM.sample_chain` (reference src/edm.py:126-176), the per-step coefficient table of
the noise schedule and a Philox (seed, offset) pair. `examples/c_sampler.c` reads the file and samples through the C-ABI
(`dl_sample_chain_rng`); `read_result` parses what it writes back."""
import ctypes as C
import struct

import numpy as np
import torch

from . import _native

MAGIC = b"DLJOB1\0\0"


def write_job(path, edm, x, h, node_mask, fragment_mask, linker_mask, edge_mask, context, keep_frames, seed, offset=0,
              device_index=0):
    """Arguments as `EDM.sample_chain` takes them (`ddpm.sampler_inputs` builds them from a collated batch)."""
    dyn = edm.dynamics
    B, N = x.size(0), x.size(1)
    T = edm.T
    xd = edm.n_dims + edm.in_node_nf
    assert 1 <= keep_frames <= T and offset % 4 == 0
    xn, hn = edm.normalize(x, h)
    xh = torch.cat([xn, hn], dim=2).to(torch.float32).cpu().contiguous()
    cfg = _native.DLConfig(
        n_dims=dyn.n_dims, in_node_nf=dyn.in_node_nf, context_node_nf=dyn.context_node_nf, hidden_nf=dyn.hidden_nf,
        n_layers=dyn.n_layers, inv_sublayers=dyn.inv_sublayers, condition_time=int(dyn.condition_time),
        centering=int(dyn.centering), graph_type=_native.GRAPH_TYPES[dyn.graph_type], device=device_index,
        edge_impl=_native.EDGE_IMPLS[dyn.edge_impl], norm_constant=float(dyn.norm_constant),
        normalization_factor=float(dyn.normalization_factor))
    assert C.sizeof(cfg) == 13 * 4 and C.sizeof(_native.DLStepCoef) == 32
    coef = edm.step_coefficients(keep_frames, B)
    f32 = lambda t: t.detach().to(device='cpu', dtype=torch.float32).contiguous().numpy().tobytes()
    i8 = lambda t: t.detach().to(device='cpu', dtype=torch.int8).contiguous().numpy().tobytes()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(bytes(cfg))
        sd = dyn.dynamics.state_dict()
        f.write(struct.pack("<i", len(sd)))
        for name, p in sd.items():
            nb = f"dynamics.{name}".encode()
            f.write(struct.pack("<i", len(nb)) + nb + struct.pack("<q", p.numel()) + f32(p))
        f.write(struct.pack("<6i", B, N, T, keep_frames, xd, dyn.context_node_nf))
        f.write(struct.pack("<2Q", seed & 0xFFFFFFFFFFFFFFFF, offset))
        f.write(struct.pack("<3f", float(edm.norm_values[0]), float(edm.norm_values[1]), float(edm.norm_biases[1])))
        f.write(bytes(coef)[:(T + 1) * 32])
        f.write(xh.numpy().tobytes())
        f.write(i8(node_mask.reshape(B, N)))
        f.write(f32(fragment_mask.reshape(B, N)))
        f.write(f32(linker_mask.reshape(B, N)))
        has_em = dyn.graph_type == 'FC' and edge_mask is not None
        f.write(struct.pack("<i", int(has_em)))
        if has_em:
            em = i8(edge_mask.reshape(-1))
            assert len(em) == B * N * N
            f.write(em)
        f.write(struct.pack("<i", int(context is not None)))
        if context is not None:
            f.write(f32(context.reshape(B, N, dyn.context_node_nf)))
    return {"B": B, "N": N, "T": T, "keep_frames": keep_frames, "xd": xd}


def read_result(path, B, N, keep_frames, xd):
    """(status, philox offset consumed, chain (keep_frames,B,N,xd) float32 tensor, flags (B,) int32 tensor)."""
    raw = open(path, "rb").read()
    status, consumed = struct.unpack_from("<iQ", raw, 0)
    n = keep_frames * B * N * xd
    chain = torch.from_numpy(np.frombuffer(raw, dtype=np.float32, count=n, offset=12).copy()).view(keep_frames, B, N, xd)
    flags = torch.from_numpy(np.frombuffer(raw, dtype=np.int32, count=B, offset=12 + 4 * n).copy())
    assert len(raw) == 12 + 4 * n + 4 * B
    return status, consumed, chain, flags
