import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import dl_helpers as helpers
from difflinker_b200 import synthetic, FoundNaNException, utils
from difflinker_b200.batching import collate, create_templates_for_linker_generation
from oracle import difflinker_oracle as orc
dev = torch.device('cuda', 0)
torch.set_num_threads(32)
spec = synthetic.SPECS["cfg5_sweep_N256"]
T = 6
for B in (2, 128):
    items = synthetic.make_items(spec, seed_offset=1)[:B]
    batch = collate(items)
    tpl = create_templates_for_linker_generation(batch, batch['linker_mask'].sum(1).view(-1).int())
    x = utils.remove_partial_mean_with_mask(tpl['positions'], tpl['atom_mask'], tpl['fragment_mask'])
    g = torch.Generator().manual_seed(7)
    noise = torch.randn((T + 2, B, spec.N, 3 + spec.F), generator=g)
    sel = [0, 1] if B == 2 else [0, 41, 77]
    chains = {}
    for impl in ("simt", "tcgen05"):
        ddpm, hp = helpers.build_ddpm(spec, 0, edge_impl=impl, diffusion_steps=T)
        mv = lambda v: v.to(dev)
        chains[impl] = ddpm.edm.sample_chain(x=mv(x), h=mv(tpl['one_hot']), node_mask=mv(tpl['atom_mask']), fragment_mask=mv(tpl['fragment_mask']), linker_mask=mv(tpl['linker_mask']), edge_mask=mv(tpl['edge_mask']), context=mv(tpl['fragment_mask']), keep_frames=T, noise=mv(noise)).cpu()
    # oracle on selected molecules, same noise rows
    idx = torch.tensor(sel)
    gam = orc.gamma_table('polynomial_2', T, 1e-5)
    em = tpl['edge_mask'].view(B, -1)[idx].reshape(-1, 1)
    ns = noise[:, idx]
    k = [0]
    def noise_fn(shape):
        # reference draw order: (B,N,3) then (B,N,F) per draw
        r = k[0] // 2; part = k[0] % 2; k[0] += 1
        return ns[r][:, :, :3] if part == 0 else ns[r][:, :, 3:]
    with torch.no_grad():
        want = orc.edm_sample_chain(ddpm.edm.dynamics.state_dict(), helpers.oracle_cfg(hp), gam, T, x[idx], tpl['one_hot'][idx], tpl['atom_mask'][idx], tpl['fragment_mask'][idx], tpl['linker_mask'][idx], em, tpl['fragment_mask'][idx], keep_frames=T, norm_values=(1, 4, 10), noise_fn=noise_fn)
    for impl in ("simt", "tcgen05"):
        got = chains[impl][:, idx]
        per_frame = [(float((got[f] - want[f]).abs().max()), float(want[f].abs().max())) for f in range(T)]
        print("B", B, impl, "per-frame (abs err, max|ref|) frames T-1..0:", [(round(a, 6), round(b, 1)) for a, b in per_frame[::-1]])
