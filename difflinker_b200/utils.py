"""Host-side helpers of the sampling path, with the reference's names (src/utils.py)."""
import torch


class FoundNaNException(Exception):
    """Same attributes as the reference exception (src/utils.py:274-289), built from the per-molecule
    flags the device path reports (bit0: NaN in coordinates/velocity, bit1: NaN in features)."""

    def __init__(self, x=None, h=None, flags=None):
        if flags is not None:
            f = [int(v) for v in flags]
            x_idx = {i for i, v in enumerate(f) if v & 1}
            h_idx = {i for i, v in enumerate(f) if v & 2}
            steps = [v >> 8 for v in f if v >> 8]
            self.first_step = min(steps) - 1 if steps else None
        else:
            x_idx = {i for i in range(x.shape[0]) if bool(torch.isnan(x[i]).any())}
            h_idx = {i for i in range(h.shape[0]) if bool(torch.isnan(h[i]).any())}
            self.first_step = None
        self.x_h_nan_idx = x_idx & h_idx
        self.only_x_nan_idx = x_idx - h_idx
        self.only_h_nan_idx = h_idx - x_idx
        super().__init__(f"NaN in dynamics output (x&h: {sorted(self.x_h_nan_idx)}, x: {sorted(self.only_x_nan_idx)}, "
                         f"h: {sorted(self.only_h_nan_idx)})")


def sample_gaussian_with_mask(size, device, node_mask):
    """src/utils.py:189-192."""
    return torch.randn(size, device=device) * node_mask


def remove_mean_with_mask(x, node_mask):
    """src/utils.py:56-63 (the .item() assert of the reference is dropped: it is a host sync)."""
    n = node_mask.sum(1, keepdims=True)
    return x - (torch.sum(x, dim=1, keepdim=True) / n) * node_mask


def remove_partial_mean_with_mask(x, node_mask, center_of_mass_mask):
    """src/utils.py:66-74: subtract the centre of mass of the `center_of_mass_mask` atoms from all atoms."""
    n = center_of_mass_mask.sum(1, keepdims=True)
    mean = torch.sum(x * center_of_mass_mask, dim=1, keepdim=True) / n
    return x - mean * node_mask


def assert_correctly_masked(variable, node_mask):
    """src/utils.py:99-101."""
    assert (variable * (1 - node_mask)).abs().max().item() < 1e-4, "Variables not masked properly."
