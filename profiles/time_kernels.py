"""Live per-kernel times of Dynamics.forward (DL_TIME_KERNELS=1: CUDA events around every launch, warm caches, back to back).
Usage: DL_TIME_KERNELS=1 python profiles/time_kernels.py [workload] [reps]"""
import os, sys
os.environ.setdefault("DL_TIME_KERNELS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difflinker_b200 import DDPM, synthetic
from difflinker_b200.batching import collate
from difflinker_b200.ddpm import sampler_inputs

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_zinc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
spec = synthetic.SPECS[wl]
hp = synthetic.model_hparams(spec)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
ddpm = DDPM(**hp)
synthetic.init_reference_like_weights(ddpm, coord_gain=100.0 if spec.N <= 64 else 1.0)
ddpm = ddpm.to(dev)
data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in collate(synthetic.make_items(spec)).items()}
kw = sampler_inputs(ddpm, data)
z = torch.cat([kw['x'], kw['h'] / 4], dim=2)
z = z * kw['fragment_mask'] + torch.randn_like(z) * kw['linker_mask']
t = torch.full((spec.B, 1), 0.5, device=dev)
dyn = ddpm.edm.dynamics
for _ in range(reps + 3):
    dyn(t, z, kw['node_mask'], kw['linker_mask'], kw['edge_mask'], kw['context'])
torch.cuda.synchronize()
del ddpm, dyn
import gc; gc.collect()
