"""Runs the same network step through the CUDA ops and through the torch-CPU
emulation of the C ABI, recording every op call's tensor arguments after the
call, and reports the first call whose outputs diverge while its inputs agree."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from oracle import refshim, restate_net
from tests import golden_inputs as gi, emul_ops
from epipolarpose_b200 import ops as real_ops, net

_zeros, _zeros_like = torch.zeros, torch.zeros_like
torch.empty = lambda *a, **k: _zeros(*a, **k)          # make partial phase writes comparable
torch.empty_like = lambda *a, **k: _zeros_like(*a, **k)
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 18
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
J, D, N = 3, HW // 4, 4

def recorder(mod, log):
    ns = types.SimpleNamespace()
    for name in dir(mod):
        f = getattr(mod, name)
        if callable(f) and not name.startswith("_") and name not in ("make_geom", "device_check"):
            def mk(name, f):
                def w(*a, **k):
                    f(*a, **k)
                    log.append((name, [(x.detach().float().cpu().clone() if isinstance(x, torch.Tensor) else None)
                                       for x in list(a) + list(k.values())]))
                return w
            setattr(ns, name, mk(name, f))
        else:
            setattr(ns, name, f)
    return ns

sd = restate_net.init_state(restate_net.param_shapes(layers, J, True, D), 5)
x = torch.from_numpy(gi.images(N, HW, 5))
plan = net.PoseNetPlan(layers, J, True, D, (HW, HW))
gout = torch.from_numpy(gi.grad_like((N, HW // 4, HW // 4, J * D), 6))
logs = {}
for tag, mod, dev in (("gpu", real_ops, "cuda:0"), ("cpu", emul_ops, "cpu")):
    log = []
    eng = net.Engine(plan, precision=prec, ops=recorder(mod, log))
    params = {k: v.clone().to(dev) for k, v in sd.items()}
    logits, _, S = eng.forward(x.to(dev), params, training=True)
    grads = {k: torch.zeros_like(v) for k, v in params.items() if v.is_floating_point() and "running" not in k}
    eng.backward(S, gout.to(dev).contiguous(), None, params, grads)
    logs[tag] = log
a, b = logs["gpu"], logs["cpu"]
print("calls", len(a), len(b))
bad = 0
for i, ((na, ta), (nb, tb)) in enumerate(zip(a, b)):
    assert na == nb
    errs = []
    for u, v in zip(ta, tb):
        if u is None or v is None or u.numel() == 0:
            errs.append(None); continue
        errs.append(float((u - v).abs().max() / v.abs().max().clamp_min(1e-30)))
    worst = max([e for e in errs if e is not None] or [0])
    if worst > 1e-3:
        print(i, na, ["%.1e" % e if e is not None else "-" for e in errs], [tuple(u.shape) if u is not None else None for u in ta])
        bad += 1
        if bad > 40: break
print("done")
