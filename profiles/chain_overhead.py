"""Where the time of one DDPM.sample_chain call goes outside the device loop (DL_TIME_CHAIN=1 prints the native side's breakdown)."""
import os, sys, time
os.environ.setdefault("DL_TIME_CHAIN", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difflinker_b200 import DDPM, synthetic
from difflinker_b200.batching import collate
from difflinker_b200.ddpm import sampler_inputs

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_zinc"
spec = synthetic.SPECS[wl]
hp = synthetic.model_hparams(spec)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
ddpm = DDPM(**hp)
synthetic.init_reference_like_weights(ddpm, coord_gain=100.0 if spec.N <= 64 else 1.0)
ddpm = ddpm.to(dev)
data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in collate(synthetic.make_items(spec)).items()}
kw = sampler_inputs(ddpm, data)
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ddpm.edm.sample_chain(**kw, keep_frames=1)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ddpm.sample_chain(data, keep_frames=1)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"call {i}: edm.sample_chain wall {1e3 * (t1 - t0):.2f} ms (device loop {ddpm.edm.last_loop_ms:.2f}); DDPM.sample_chain wall {1e3 * (t2 - t1):.2f} ms", file=sys.stderr)
