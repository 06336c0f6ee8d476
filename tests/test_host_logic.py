"""CPU: host-side logic (network plan, geometry / tap tables, weight packing,
autograd wiring, optimiser, config, dataset, loops) through the torch-CPU
emulation of the C ABI (tests/emul_ops.py), and the C-ABI library itself
(loads; exports every symbol include/epb.h declares; refuses to compute
without a GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import refshim, restate, restate_net
from tests import emul_ops, golden_inputs as gi
from tests.conftest import ROOT, relerr


def test_c_abi_exports_match_header():
    from epipolarpose_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "epb.h")).read()
    declared = set(re.findall(r"\b(epb_[a-z0-9_]+)\s*\(", hdr)) - {"epb_conv_geom"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().epb_version() >= 100


def test_product_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from epipolarpose_b200 import _lib, ops
    with pytest.raises(_lib.EpbError):
        ops.device_check()
    import lib.models as models
    cfg = refshim.make_cfg(num_layers=18, num_joints=2, depth_res=4, image_size=(32, 32))
    m = models.pose3d_resnet.get_pose_net(cfg, False)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 32, 32))            # no CPU fallback
    import lib.core.integral_loss as il
    with pytest.raises(_lib.EpbError):
        il.softmax_integral_tensor(torch.zeros(1, 8, 4, 4), 2, True, 4, 4, 4)


def test_config_parses_reference_yamls(tmp_path):
    from lib.core.config import config, update_config, reset_config, get_model_name, gen_config
    y = tmp_path / "exp.yaml"
    y.write_text("GPUS: '0'\nMODEL:\n  NUM_JOINTS: 16\n  VOLUME: true\n  EXTRA:\n    NUM_LAYERS: 50\n"
                 "LOSS:\n  FN: SmoothL1JointLocationLoss\nTRAIN:\n  LR_STEP:\n  - 90\n  - 120\n")
    reset_config()
    update_config(str(y))
    assert config.MODEL.NUM_JOINTS == 16 and config.TRAIN.LR_STEP == [90, 120]
    assert get_model_name(config)[0] == "pose3d_resnet_50"
    bad = tmp_path / "bad.yaml"
    bad.write_text("MODEL:\n  NOT_A_KEY: 1\n")
    with pytest.raises(ValueError):
        update_config(str(bad))
    bad.write_text("NOPE: 1\n")
    with pytest.raises(ValueError):
        update_config(str(bad))
    gen_config(str(tmp_path / "dump.yaml"))
    reset_config()


@pytest.mark.parametrize("layers,volume", [(18, True), (34, True), (50, False)])
def test_engine_matches_oracle_through_emulated_abi(layers, volume):
    """Whole-network forward + backward wiring (geometry tables, phase
    decomposition of stride-2 dgrad / deconv, packing, BN, residual merges)."""
    import lib.models as models
    J, D, HW, N = 3, 8, 64, 2
    cfg = refshim.make_cfg(num_layers=layers, num_joints=J, volume=volume, depth_res=D,
                           image_size=(HW, HW))
    shapes = restate_net.param_shapes(layers, J, volume, D)
    sd = restate_net.init_state(shapes, 9)
    m = models.pose3d_resnet.get_pose_net(cfg, False, ops=emul_ops)
    assert list(m.state_dict().keys()) == list(shapes.keys())
    m.load_state_dict(sd)
    m.train()
    x = torch.from_numpy(gi.images(N, HW, 9))
    p = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
             else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
    ref = restate_net.forward(p, x.double(), num_layers=layers, volume=volume, image_size=(HW, HW))
    out = m(x)
    refs = ref if isinstance(ref, tuple) else (ref,)
    outs = out if isinstance(out, tuple) else (out,)
    gs = [torch.from_numpy(gi.grad_like(o.shape, 10 + i)) for i, o in enumerate(outs)]
    for o, r in zip(outs, refs):
        assert relerr(o.detach().numpy(), r.detach().numpy()) <= 2e-3
    sum((o * g).sum() for o, g in zip(outs, gs)).backward()
    sum((r * g.double()).sum() for r, g in zip(refs, gs)).backward()
    if layers != 50:      # R50 at 64x64 / batch 2 is chaotic in fp32 (2x2x2 BN statistics)
        for k, q in m.named_parameters():
            assert relerr(q.grad.numpy(), p[k].grad.numpy()) <= 5e-3, k
    assert int(m.state_dict()["bn1.num_batches_tracked"]) == 1


def test_losses_and_decode_surface_emulated():
    import lib.core.integral_loss as il
    import lib.core.inference as inf
    il._backend[0] = emul_ops
    inf._backend[0] = emul_ops
    try:
        N, J, D = 2, 3, 8
        x = torch.from_numpy(gi.logits(N, J, D, D, D, 4, 2.0)).requires_grad_(True)
        gt, wt = gi.labels(N, J, 4)
        for cls, kind in ((il.L1JointLocationLoss, "l1"), (il.SmoothL1JointLocationLoss, "smoothl1"),
                          (il.L2JointLocationLoss, "mse")):
            for norm in (False, True):
                x.grad = None
                loss = cls(J, norm=norm)(x, torch.from_numpy(gt), torch.from_numpy(wt))
                loss.backward()
                c = restate.softmax_integral(x.detach().numpy(), J, D, D, D)
                rl, dc = restate.weighted_loss(kind, c, gt, wt, True, norm)
                assert abs(loss.item() - rl) <= 1e-5
                rg = restate.softmax_integral_grad(x.detach().numpy(), dc, J, D, D, D)
                assert relerr(x.grad.numpy(), rg) <= 1e-3
        with pytest.raises(AssertionError):
            il.L1JointLocationLoss(J)(x, torch.from_numpy(gt).requires_grad_(True), torch.from_numpy(wt))
        hm = gi.argmax_heatmaps()
        p, mv = inf.get_max_preds(hm)
        rp, rm, _ = restate.get_max_preds(hm)
        assert np.array_equal(p, rp) and np.array_equal(mv, rm)
        with pytest.raises(AssertionError):
            inf.get_max_preds(hm[0])
    finally:
        il._backend[0] = __import__("epipolarpose_b200.ops", fromlist=["ops"])
        inf._backend[0] = il._backend[0]


def test_fused_optimizers_match_torch():
    import lib.utils.utils as U
    U._backend[0] = emul_ops
    try:
        torch.manual_seed(0)
        for kind in ("adam", "sgd"):
            ps = [torch.nn.Parameter(torch.randn(s)) for s in ((7, 3), (64,), (5, 5, 3), (17,))]
            qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
            if kind == "adam":
                a, b = U.FusedAdam(ps, lr=1e-2), torch.optim.Adam(qs, lr=1e-2)
            else:
                a = U.FusedSGD(ps, lr=1e-2, momentum=0.9, weight_decay=1e-4, nesterov=True)
                b = torch.optim.SGD(qs, lr=1e-2, momentum=0.9, weight_decay=1e-4, nesterov=True)
            sched = torch.optim.lr_scheduler.MultiStepLR(a, [2], 0.1)
            schedb = torch.optim.lr_scheduler.MultiStepLR(b, [2], 0.1)
            for it in range(4):
                for p, q in zip(ps, qs):
                    g = torch.randn_like(p)
                    p.grad, q.grad = g.clone(), g.clone()
                a.step(); b.step(); sched.step(); schedb.step()
            for p, q in zip(ps, qs):
                assert relerr(p.detach().numpy(), q.detach().numpy()) <= 1e-5
    finally:
        U._backend[0] = __import__("epipolarpose_b200.ops", fromlist=["ops"])


def test_synthetic_dataset_contract_and_pairing():
    from lib.core.config import config, reset_config
    import lib.dataset as dataset
    reset_config()
    config.MODEL.NUM_JOINTS = 16
    ds = eval("dataset.synthetic_h36m")(cfg=config, root="", image_set="train", is_train=True)
    img, label, weight, meta = ds[5]
    assert img.shape == (3, 256, 256) and img.dtype == torch.float32
    assert label.shape == (48,) and weight.shape == (48,)
    for k in ("center_x", "center_y", "width", "height", "scale", "rot", "R", "T", "f", "c",
              "projection_matrix", "image"):
        assert k in meta
    P = restate.projection_matrix(meta["R"], meta["T"], meta["f"], meta["c"])
    assert np.allclose(P, meta["projection_matrix"])
    b = next(iter(ds.pair_batch_sampler(3)))
    half = len(b) // 2
    for i in range(half):                # halves pair neighbouring views of the same tuple
        a, c = ds.db[b[i]], ds.db[b[half + i]]
        assert a["tuple"] == c["tuple"] and (a["view"], c["view"]) in ((0, 1), (3, 2))
    assert len(ds.db) == len(ds)


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lib.models as models
    cfg = refshim.make_cfg(num_layers=18, num_joints=2, volume=True, depth_res=4, image_size=(32, 32))
    m = models.pose3d_resnet.get_pose_net(cfg, False, ops=emul_ops)
    m.load_state_dict(restate_net.init_state(restate_net.param_shapes(18, 2, True, 4), 1))
    m.train()
    x = torch.from_numpy(gi.images(4, 32, 100))[2 * rank:2 * rank + 2]
    g = torch.from_numpy(gi.grad_like((4, 8, 8, 8), 101))[2 * rank:2 * rank + 2]
    import torch.distributed as d2
    orig = d2.all_reduce

    def avg(t, op=None):          # gloo has no AVG: emulate with SUM / world
        orig(t)
        t /= world
    d2.all_reduce = avg
    (m(x) * g).sum().backward()
    q.put((rank, {k: p.grad.numpy().copy() for k, p in m.named_parameters()}))
    dist.destroy_process_group()


def test_data_parallel_allreduce_world2_gloo():
    """Two ranks, each with its own half batch: after the single all-reduce in
    the model's backward both ranks hold identical (averaged) gradients."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    ps = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(60)
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k
    assert any(np.abs(v).max() > 0 for v in res[0].values())


def test_bench_reference_arm_cli():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0", "--layers", "18", "--cpu-tuples", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0
