"""Throughput of the small kernels around the CNN (HBM / latency bound; no tensor cores):
fused heat-map + joint loss, H36M evaluation, the four triangulators.  One JSON line each:
algorithmic bytes per unit / CUDA-event time, against the measured HBM peak.

    python tools/bench_aux.py
"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from epipolarpose_b200 import ops

dev = torch.device("cuda:0")
peak = 6479.6
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=9):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e-3


def line(name, units, unit_name, bytes_per_unit, t):
    gbs = units * bytes_per_unit / t / 1e9
    print(json.dumps({"kernel": name, "units": units, "unit": unit_name, "ms": round(t * 1e3, 4),
                      "units_per_s": round(units / t, 1), "bytes_per_unit": bytes_per_unit,
                      "achieved_gbs": round(gbs, 1), "hbm_peak_gbs": peak, "frac": round(gbs / peak, 4)}))


# fused heat-map MSE + L1 joint loss: 128 / 1024 images x 17 joints x 64 x 64
for N in (128, 1024):
    J, HW = 17, 64 * 64
    hm = torch.randn(N * J, HW, device=dev); tg = torch.rand(N * J, HW, device=dev)
    wh = torch.ones(N * J, device=dev)
    x = torch.rand(N, J * 3, device=dev) - 0.5; t = torch.rand(N, J * 3, device=dev) - 0.5
    w = torch.ones(N, J * 3, device=dev)
    loss = torch.empty(3, device=dev); dhm = torch.empty_like(hm); dx = torch.empty_like(x)
    tt = timeit(lambda: ops.heatmap_joint_loss(hm, tg, wh, N * J, HW, 1.0, x, t, w, x.numel(), 1, float(N),
                                               1.0, loss, dhm, dx))
    line("heatmap_joint_loss N=%d" % N, N * J * HW, "heat-map element", 12, tt)

# H36M evaluation: 2^18 samples x 17 joints (pred + gt 816 B, cam 40 B, metrics 72 B, per-joint 136 B, pck 68 B)
S, J = 1 << 18, 17
pred = torch.randn(S, J, 3, device=dev, dtype=torch.float64) * 50 + 500
gt = pred + torch.randn(S, J, 3, device=dev, dtype=torch.float64) * 5
pred[:, :, 2] = torch.randn(S, J, device=dev, dtype=torch.float64) * 300; gt[:, :, 2] = pred[:, :, 2] + 20
cam = torch.tensor([1145.0, 1144.0, 512.0, 515.0, 4500.0], device=dev, dtype=torch.float64).repeat(S, 1)
met = torch.empty(S, 9, device=dev, dtype=torch.float64); pj = torch.empty(S, J, device=dev, dtype=torch.float64)
pck = torch.empty(S, J, device=dev, dtype=torch.int32)
mask = sum(1 << j for j in [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15])
tt = timeit(lambda: ops.h36m_eval(pred, gt, cam, S, J, 0, mask, 150.0, met, pj, pck, None))
line("h36m_eval", S, "sample", 816 + 40 + 72 + 136 + 68, tt)

# triangulators: 17-joint pairs, 1144 B per pair: the real size of a step (64 pairs) and a saturating one (2^20)
J = 17
rng = np.random.default_rng(0)
sys.path.insert(0, ROOT)
from oracle import restate          # camera synthesis only (bench input), not a compute path
R, T, f, c, P = restate.synthetic_cameras(rng, 64, 4)
X = rng.normal(0, 400, (64, J, 3))
u1 = np.stack([restate.project(P[i, 0], X[i]) for i in range(64)]) + rng.normal(0, 3, (64, J, 2))
u2 = np.stack([restate.project(P[i, 1], X[i]) for i in range(64)]) + rng.normal(0, 3, (64, J, 2))
for NP in (64, 1 << 20):
    rep = NP // 64
    tu1 = torch.from_numpy(np.tile(u1, (rep, 1, 1))).to(dev); tu2 = torch.from_numpy(np.tile(u2, (rep, 1, 1))).to(dev)
    tP1 = torch.from_numpy(np.tile(P[:, 0], (rep, 1, 1))).to(dev).contiguous()
    tP2 = torch.from_numpy(np.tile(P[:, 1], (rep, 1, 1))).to(dev).contiguous()
    Xo = torch.empty(NP, J, 3, device=dev, dtype=torch.float64); st = torch.empty(NP, J, device=dev, dtype=torch.int32)
    for m, name in ((0, "linear_eigen"), (1, "linear_LS"), (2, "iterative_LS"), (3, "polynomial")):
        tt = timeit(lambda: ops.triangulate(tu1, tu2, 2, tP1, tP2, NP, J, m, 3e-5, Xo, st), reps=5)
        line("triangulate %s, %d pairs" % (name, NP), NP, "17-joint pair", 1144, tt)
    del tu1, tu2, tP1, tP2, Xo, st

# Adam over the R50 parameter vector: 4 reads + 3 writes of 4 B per parameter
n = 34272704
pp, gg, m1, m2 = (torch.randn(n, device=dev) * 1e-2 for _ in range(4))
m2.abs_()
tt = timeit(lambda: ops.adam_step(pp, gg, m1, m2, n, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1))
line("adam_step (R50, 34.27 M parameters)", n, "parameter", 28, tt)
del pp, gg, m1, m2

# input pipeline after decode: 512 frames of 1000 x 1002 x 3 (H36M camera frames) -> 256 x 256 patches.
# Algorithmic bytes per output pixel: 12 written (3 float planes) + 12 read (4 bilinear taps x 3 channels, the
# source footprint of a 0.8-1.2x crop is about one source pixel per output pixel) = 24 B.
B, HI, WI, PW = 512, 1002, 1000, 256
img = torch.randint(0, 256, (HI * WI * 3,), dtype=torch.uint8, device=dev)
per = (img.numel() + 15) // 16 * 16
base = torch.zeros(per * 8, dtype=torch.uint8, device=dev)
for i in range(8):
    base[i * per:i * per + img.numel()] = torch.roll(img, i * 977)
offs = torch.tensor([(i % 8) * per for i in range(B)], dtype=torch.int64, device=dev)
hwp = torch.tensor([[HI, WI, WI * 3]] * B, dtype=torch.int32, device=dev)
g = torch.Generator().manual_seed(3)
box = torch.stack([500 + 40 * torch.randn(B, generator=g, dtype=torch.float64),
                   500 + 40 * torch.randn(B, generator=g, dtype=torch.float64),
                   torch.full((B,), 300.0, dtype=torch.float64), torch.full((B,), 300.0, dtype=torch.float64),
                   1 + 0.25 * (torch.rand(B, generator=g, dtype=torch.float64) - 0.5),
                   30 * (torch.rand(B, generator=g, dtype=torch.float64) - 0.5)], dim=1).contiguous().to(dev)
flip = (torch.rand(B, generator=g) < 0.5).to(torch.int32).to(dev)
col = (0.8 + 0.4 * torch.rand(B, 3, generator=g)).to(dev)
ms = [0.485, 0.456, 0.406, 0.229, 0.224, 0.225]
out = torch.empty(B, 3, PW, PW, device=dev); trans = torch.empty(B, 6, device=dev, dtype=torch.float64)
tt = timeit(lambda: ops.patch_sample(base, offs, hwp, box, flip, col, ms, B, PW, PW, out, trans))
line("patch_sample 256x256 (warpAffine+colour+normalise)", B * PW * PW, "output pixel", 24, tt)

from epipolarpose_b200.lib.utils.augmentation import pack_occluders
rs = np.random.RandomState(5)
occ_lists = []
for b in range(B):
    lst = []
    for k in range(rs.randint(1, 8)):                       # augmentation.py:65 count = randint(1, 8)
        h, w = rs.randint(30, 120), rs.randint(30, 120)
        rgba = rs.randint(0, 256, (h, w, 4)).astype(np.uint8)
        lst.append((rgba, (int(rs.randint(0, PW)), int(rs.randint(0, PW)))))
    occ_lists.append(lst)
ob, od, oc = pack_occluders(occ_lists, dev)
occ_px = float(sum(r.shape[0] * r.shape[1] for l in occ_lists for r, _ in l)) / (B * PW * PW)
tt = timeit(lambda: ops.patch_sample_occ(base, offs, hwp, box, flip, col, ms, B, PW, PW, ob, od, oc, out, trans))
line("patch_sample_occ (+%.2f RGBA occluder px per output px)" % occ_px, B * PW * PW, "output pixel",
     round(24 + 4 * occ_px, 2), tt)

# decode: hard argmax and get_final_preds over 1024 x 17 heat-maps of 64 x 64 (4 B per element read once)
N, J, H, W = 1024, 17, 64, 64
hm = torch.rand(N, J, H, W, device=dev)
ctr = torch.full((N, 2), 500.0, dtype=torch.float64, device=dev)
scl = torch.full((N, 2), 1.5, dtype=torch.float64, device=dev)
preds = torch.empty(N, J, 2, device=dev); mv = torch.empty(N, J, 1, device=dev)
tt = timeit(lambda: ops.final_preds(hm, N, J, H, W, ctr, scl, True, preds, mv))
line("final_preds (argmax + refine + transform_preds)", N * J * H * W, "heat-map element", 4, tt)
