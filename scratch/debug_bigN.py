import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import dl_helpers as helpers
from difflinker_b200 import synthetic, FoundNaNException
from difflinker_b200.batching import collate, create_templates_for_linker_generation
from difflinker_b200 import utils
dev = torch.device('cuda', 0)
for name in ["cfg5_sweep_N256", "cfg3_geom", "cfg4_pockets"]:
    spec = synthetic.SPECS[name]
    batch = collate(synthetic.make_items(spec, seed_offset=1))
    z, t = helpers.random_latent(batch, 5)
    ctx = helpers.context_of(batch, spec)
    outs = {}
    for impl in ("simt", "tcgen05"):
        dyn, hp = helpers.build_dynamics(spec, 0, edge_impl=impl)
        mv = lambda v: None if v is None else v.to(dev)
        try:
            outs[impl] = dyn(mv(t), mv(z), mv(batch['atom_mask']), mv(batch['linker_mask']), mv(batch['edge_mask']), mv(ctx)).cpu()
        except FoundNaNException as e:
            print(name, impl, "forward NaN", e)
    if len(outs) == 2:
        a, b = outs["simt"], outs["tcgen05"]
        err = (a - b).abs().amax(dim=(1, 2)) / a.abs().amax().clamp_min(1e-30)
        print(name, "forward simt vs tc: max rel err", float(err.max()), "worst mol", int(err.argmax()), "finite", bool(torch.isfinite(b).all()))
    # short chain with both impls, same noise
    if spec.pocket: continue
    chains = {}
    for impl in ("simt", "tcgen05"):
        ddpm, hp = helpers.build_ddpm(spec, 0, edge_impl=impl, diffusion_steps=10)
        tpl = create_templates_for_linker_generation(batch, batch['linker_mask'].sum(1).view(-1).int())
        x = utils.remove_partial_mean_with_mask(tpl['positions'], tpl['atom_mask'], tpl['fragment_mask'])
        g = torch.Generator().manual_seed(7)
        noise = torch.randn((12, spec.B, spec.N, 3 + spec.F), generator=g)
        mv = lambda v: v.to(dev)
        try:
            chains[impl] = ddpm.edm.sample_chain(x=mv(x), h=mv(tpl['one_hot']), node_mask=mv(tpl['atom_mask']), fragment_mask=mv(tpl['fragment_mask']), linker_mask=mv(tpl['linker_mask']), edge_mask=mv(tpl['edge_mask']), context=mv(tpl['fragment_mask']), keep_frames=1, noise=mv(noise)).cpu()
        except FoundNaNException as e:
            print(name, impl, "chain NaN first_step", e.first_step, str(e)[:200])
    if len(chains) == 2:
        a, b = chains["simt"], chains["tcgen05"]
        print(name, "chain simt vs tc max abs diff", float((a - b).abs().max()), "max|x|", float(a[0][..., :3].abs().max()))
