"""Error of the GPU sampler against the reference's golden chains at the benchmarked shapes (diagnostic print-out of what
tests/test_gpu_parity.py::test_public_ddpm_sample_chain_matches_reference_golden asserts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import dl_helpers as helpers
from difflinker_b200 import synthetic
from difflinker_b200.batching import collate, create_templates_for_linker_generation

d = torch.device("cuda", 0)
for name in sys.argv[1:] or ["chain_cfg2_zinc_T500", "chain_cfg2_zinc_L8_T500", "chain_cfg3_geom_T500"]:
    meta, a = helpers.load_golden(name)
    spec = helpers.spec_by_name(meta["spec"])
    for impl in ("auto", "simt"):
        ddpm, hp = helpers.build_ddpm(spec, meta["seed"], edge_impl=impl, diffusion_steps=meta["table_timesteps"])
        ddpm.edm.T = meta["T"]
        ddpm = ddpm.to(d)
        data = collate(synthetic.make_items(spec, batch=meta["batch"]))
        tpl = create_templates_for_linker_generation(data, data['linker_mask'].sum(1).view(-1).int())
        B, N = tpl['positions'].shape[:2]
        noise = helpers.noise_tensor(meta["noise_seed"], meta["T"], B, N, spec.F)
        ddpm.edm.draw_noise = lambda *args, **kw: noise.to(d)
        chain, nm = ddpm.sample_chain({k: (v.to(d) if torch.is_tensor(v) else v) for k, v in data.items()}, keep_frames=meta["keep_frames"])
        chain = chain.cpu(); want = a["chain"]
        lm = tpl['linker_mask']
        dx = ((chain[0][..., :3] - want[0][..., :3]) * lm).abs()
        per_mol = dx.flatten(1).max(1).values
        scale = want[0][..., :3].abs().max().item()
        print(f"{name} [{impl}]: max|dx| {dx.max().item():.3e} / max|x| {scale:.3e} = {dx.max().item() / scale:.3e}; per molecule "
              + " ".join(f"{v:.1e}" for v in per_mol.tolist()) + f"; types equal {torch.equal(chain[0][..., 3:], want[0][..., 3:])}", flush=True)
