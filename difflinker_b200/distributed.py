"""Replica-level data parallelism: molecules never interact across a batch, so sampling shards
with no data-path collective. The only exchange is one broadcast of the weights at start-up (NCCL on GPUs,
gloo in the CPU tests); optionally the final results are gathered."""
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of `n_items` for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def batch_ids_for_rank(n_batches: int, rank: int, world: int):
    """Round-robin whole batches {rank, rank+world, ...}: weak scaling, seed = seed0 + batch id."""
    return list(range(rank, n_batches, world))


def broadcast_module_weights(module: torch.nn.Module, src: int = 0, device=None):
    """One flat fp32 broadcast of every parameter/buffer (about 6 MB at L=6). Returns the number of floats."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    if not tensors:
        return 0
    dev = device if device is not None else tensors[0].device
    flat = torch.cat([t.reshape(-1).to(device=dev, dtype=torch.float32) for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].reshape(t.shape).to(device=t.device, dtype=t.dtype))
            off += n
    return off


def gather_chains(chain: torch.Tensor, dst: int = 0):
    """Optional final gather of per-rank (keep,B,N,D) results on `dst` (92 KB per rank for cfg 3)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [chain]
    out = [torch.empty_like(chain) for _ in range(dist.get_world_size())] if dist.get_rank() == dst else None
    dist.gather(chain, out, dst=dst)
    return out


def slice_sampler_inputs(kw: dict, lo: int, hi: int):
    """Rows [lo, hi) of the keyword arguments `ddpm.sampler_inputs` builds for `EDM.sample_chain`. The FC edge mask is the
    flattened (B*N*N, 1) tensor of datasets.collate; the pocket variant is the per-node batch-id vector (B*N)."""
    B, N = kw['x'].shape[0], kw['x'].shape[1]
    out = {}
    for k, v in kw.items():
        if k == 'edge_mask' and v is not None:
            per = v.shape[0] // B
            out[k] = v[lo * per:hi * per]
        else:
            out[k] = None if v is None else v[lo:hi]
    return out


def sample_chain_sharded(model, data, sample_fn=None, keep_frames=None, gather=True):
    """Strong scaling of ONE batch: the template batch is built once (so every rank pads to the same
    N), each rank runs the reverse loop for its contiguous slice of the molecules with the slice's rows of the full-batch
    noise, and the chains are gathered -- the result equals `model.sample_chain(data)` on one GPU bit for bit, for any
    world size. No collective inside the loop. Returns (chain, node_mask) with the full batch on every rank when
    `gather`, else the local slice."""
    from .ddpm import sampler_inputs
    kw = sampler_inputs(model, data, sample_fn)
    B = kw['x'].shape[0]
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_range(B, rank, world)
    local = slice_sampler_inputs(kw, lo, hi)
    chain = model.edm.sample_chain(**local, keep_frames=keep_frames, batch_slice=(lo, B) if world > 1 else None)
    if world == 1:
        return chain, kw['node_mask']
    if not gather:
        return chain, local['node_mask']
    # ranks may hold different numbers of molecules: pad to the largest slice, all_gather, trim
    counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    m = max(counts)
    pad = torch.zeros((chain.shape[0], m) + tuple(chain.shape[2:]), dtype=chain.dtype, device=chain.device)
    pad[:, :hi - lo] = chain
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.cat([p[:, :c] for p, c in zip(parts, counts)], dim=1)
    return full, kw['node_mask']

