"""Debug: neighbour-list statistics in forward mode vs sampler mode for the pockets workload."""
import ctypes, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ["DL_DEBUG_CUT"] = "1"
import torch
import dl_helpers as helpers
from difflinker_b200 import synthetic, _native, utils
from difflinker_b200.batching import collate

spec = synthetic.SPECS["cfg4_pockets"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ddpm, hp = helpers.build_ddpm(spec, 0)
synthetic.init_reference_like_weights(ddpm, 1.0)
ddpm.edm.T = T
d = torch.device('cuda:0')
ddpm = ddpm.to(d)
data = collate(synthetic.make_items(spec, batch=B))
print("cpu edge estimate", sum(e for e, _ in synthetic.cutoff_edge_counts(data, spec.graph_type)))
lib = _native.load_library()
st = (ctypes.c_int64 * 4)()
dd = {k: (v.to(d) if torch.is_tensor(v) else v) for k, v in data.items()}
chain, nm = ddpm.sample_chain(dd, keep_frames=1)
eng = ddpm.edm.dynamics.engine(0)
lib.dl_cut_graph_stats(eng, st)
print("after chain", list(st))
x = chain[0][..., :3].cpu()
print("final |x| max", x.abs().max().item(), "linker |x| max", (x * data['linker_mask']).abs().max().item())
# forward mode on the template input
z = torch.cat([utils.remove_partial_mean_with_mask(data['positions'], data['atom_mask'], data['fragment_mask']), data['one_hot'] / 4], dim=2)
t = torch.full((B, 1), 0.5)
ctx = helpers.context_of(data, spec)
out = ddpm.edm.dynamics(t.to(d), z.to(d), dd['atom_mask'], dd['linker_mask'], dd['edge_mask'], ctx.to(d))
lib.dl_cut_graph_stats(eng, st)
print("after forward", list(st))
