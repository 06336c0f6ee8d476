"""Host-thread sweep of the CPU reference step (to pick the thread count the reference arm uses)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import restate_net
sd = restate_net.init_state(restate_net.param_shapes(50, 16, True, 64), 0, scale_final=0.001)
p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
x = torch.randn(8, 3, 256, 256)
print("cpu_count", os.cpu_count())
for th in (8, 16, 32, 64, os.cpu_count()):
    torch.set_num_threads(th)
    t0 = time.perf_counter()
    o = restate_net.forward(p, x, num_layers=50, training=True)
    o.square().mean().backward()
    print("threads", th, "fwd+bwd 8 imgs: %.2f s" % (time.perf_counter() - t0), flush=True)
