"""BatchNorm-backward passes of the split path on one tensor (target for ncu / A-B timing):
    python tools/one_bn16.py M C [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarpose_b200 import ops

M, C = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
z = torch.randn(M, C, device=dev)
dy = torch.randn(M, C, device=dev) * 1e-4
mean, var = z.mean(0), z.var(0, unbiased=False)
invstd = 1.0 / torch.sqrt(var + 1e-5)
gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
scale, shift = gamma * invstd, beta - mean * gamma * invstd
sums = torch.zeros(2 * C, device=dev, dtype=torch.float64)
mx = torch.zeros(2 * C, device=dev)
dz = torch.empty(2, M, C, device=dev, dtype=torch.float16)
a = torch.empty(2, M, C, device=dev, dtype=torch.float16)
sc = torch.empty(2, device=dev)
asc = torch.tensor([16.0, 1 / 16.0, 4000.0, 0.0], device=dev)
dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, name, nbytes):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print("%-24s M=%d C=%d: %.3f ms, %.0f GB/s" % (name, M, C, t, nbytes / t / 1e6))


def reduce():
    sums.zero_(); mx.zero_()
    ops.bn_bwd_reduce_mx(dy, z, None, scale, shift, mean, invstd, 1, M, C, sums, mx)


def apply():
    ops.bn_bwd_apply_split(dy, z, None, scale, shift, mean, invstd, gamma, 1, sums, mx, M, C, dz, sc, None, dg, db)


def act():
    ops.bn_act_split(z, scale, shift, None, None, None, None, None, 1, M, C, a, asc)


def fused():
    ops.bn_bwd_split(dy, z, None, scale, shift, mean, invstd, gamma, 1, M, C, dz, sc, None, dg, db)


mask = (torch.rand(M, C, device=dev) > 0.5).to(torch.float16)
bits = torch.randint(0, 256, (M * C // 8,), device=dev, dtype=torch.uint8)


def fused_mask():
    ops.bn_bwd_split(dy, z, mask, scale, shift, mean, invstd, gamma, 0, M, C, dz, sc, dy, dg, db)


def fused_bits():
    ops.bn_bwd_split(dy, z, None, scale, shift, mean, invstd, gamma, 0, M, C, dz, sc, dy, dg, db, mask_bits=bits)


timeit(fused, "bwd_split", 20.0 * M * C)
timeit(fused_mask, "bwd_split+mask+inplace", 28.0 * M * C)
timeit(fused_bits, "bwd_split+bits+inplace", 24.25 * M * C)
timeit(reduce, "reduce_mx", 8.0 * M * C)
reduce()
timeit(apply, "apply", 12.0 * M * C)
timeit(act, "bn_act", 8.0 * M * C)
