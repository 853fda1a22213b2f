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
