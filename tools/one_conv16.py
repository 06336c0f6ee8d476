"""One conv layer on the split-fp16 path, run a few times (target for ncu and A/B timing):

    python tools/one_conv16.py N H W Cin Cout k stride [what=fprop|dgrad|wgrad] [reps] [kind=conv|deconv]
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarpose_b200 import net, ops

N, H, W, cin, cout, k, s = [int(a) for a in sys.argv[1:8]]
what = sys.argv[8] if len(sys.argv) > 8 else "fprop"
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 5
kind = sys.argv[10] if len(sys.argv) > 10 else "conv"
dev = torch.device("cuda:0")
conv = net.Conv("t", kind, cin, cout, k, s, (k // 2) if kind == "conv" else 1, 0)
torch.manual_seed(0)
Ho, Wo = conv.out_hw(H, W)
T = k * k


def split_of(t):
    h = torch.empty(2 * t.numel(), device=dev, dtype=torch.float16)
    sc = torch.ones(2, device=dev)
    ops.split16_batch(ops.SplitBatch([(t.reshape(-1), h, sc)]))
    return h, sc


x, x_sc = split_of(torch.relu(torch.randn(N, H, W, conv.cin_p, device=dev)))
x = x.view(2, N, H, W, conv.cin_p)
dz, dz_sc = split_of(torch.randn(N, Ho, Wo, conv.cout_p, device=dev) * 1e-4)
dz = dz.view(2, N, Ho, Wo, conv.cout_p)
wf, wf_sc = split_of(torch.randn(conv.cout_p * T * conv.cin_p, device=dev) * 0.05)
wd, wd_sc = split_of(torch.randn(conv.cin_p * T * conv.cout_p, device=dev) * 0.05)
out = torch.empty(N, Ho, Wo, conv.cout_p, device=dev)
din = torch.empty(N, H, W, conv.cin_p, device=dev)
dw = torch.zeros(conv.cout_p * T * conv.cin_p, device=dev)
ws = torch.empty(48 << 20, device=dev)
stats = torch.zeros(2 * conv.cout_p, device=dev, dtype=torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    if what == "fprop":
        for g in conv.fprop_geoms(ops, N, H, W, 3):
            if g is not None:
                g.in_relu, g.accumulate = 0, 0
                ops.conv16_fprop(g, x, x_sc, wf, wf_sc, out, None, stats)
    elif what == "dgrad":
        for g in conv.dgrad_geoms(ops, N, H, W, 3):
            if g is not None:
                g.in_relu, g.accumulate = 0, 0
                ops.conv16_fprop(g, dz, dz_sc, wd, wd_sc, din, None, None)
    else:
        for g in conv.fprop_geoms(ops, N, H, W, 3):
            if g is not None:
                g.in_relu, g.accumulate = 0, 0
                ops.conv16_wgrad(g, x, x_sc, dz, dz_sc, dw, ws)


run(); run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
M = N * (Ho * Wo if kind == "conv" else H * W)
fl = 2.0 * M * cin * cout * k * k
t = sorted(ts)[len(ts) // 2]
print("%s16 %s N%d %dx%d %d->%d k%d s%d: %.3f ms (median of %d), %.1f TFLOP/s algorithmic"
      % (what, kind, N, H, W, cin, cout, k, s, t, reps, fl / t / 1e9))

if int(os.environ.get("EPB_C16_PROBE", "0")) & 32 and what != "wgrad":
    # pipeline trace of cluster 0 (csrc/conv16.cu C16_TR): clock64() ticks, SM clock ~1.9 GHz
    import ctypes
    import numpy as np
    from epipolarpose_b200 import _lib
    buf = np.zeros(4 * 2 * 256, dtype=np.int64)
    _lib.call("epb_debug_conv16_trace", ctypes.c_void_p(buf.ctypes.data), buf.size)
    tr = buf.reshape(4, 2, 256)
    mma = tr[1, 0]
    n_t = int((mma > 0).sum()) // 4
    mma = mma[:4 * n_t].reshape(n_t, 4)                     # (a) ready (b) buffer free (c) operands (d) issued
    epi = tr[2, 0]
    n_e = int((epi > 0).sum()) // 3
    epi = epi[:3 * n_e].reshape(n_e, 3)                     # (a) waiting (b) accumulator complete (c) written
    prod = tr[0, 0]
    prod = prod[prod > 0]
    us = lambda c: c / 1.93e3
    print("tiles traced: MMA %d, epilogue %d, producer k-blocks %d" % (n_t, n_e, prod.size))
    if n_t > 6:
        sl = slice(3, n_t - 1)
        print("MMA issuer per tile (us, median): period %.2f | wait buffer %.2f | wait operands %.2f | issue %.2f" %
              (np.median(us(np.diff(mma[:, 0])[sl])), np.median(us(mma[sl, 1] - mma[sl, 0])),
               np.median(us(mma[sl, 2] - mma[sl, 1])), np.median(us(mma[sl, 3] - mma[sl, 2]))))
    if n_e > 6:
        sl = slice(3, n_e - 1)
        print("epilogue warp per tile (us, median): period %.2f | wait accumulator %.2f | drain + store %.2f" %
              (np.median(us(np.diff(epi[:, 0])[sl])), np.median(us(epi[sl, 1] - epi[sl, 0])),
               np.median(us(epi[sl, 2] - epi[sl, 1]))))
        if n_t == n_e:
            print("accumulator complete after the last MMA was issued (us, median): %.2f" %
                  np.median(us(epi[sl, 1] - mma[sl, 3])))
            print("buffer seen free after the epilogue finished the tile two back (us, median): %.2f" %
                  np.median(us(mma[5:n_t - 1, 1] - epi[3:n_t - 3, 2])))
    if prod.size > 8:
        print("producer: k-block issue period %.2f us (median)" % np.median(us(np.diff(prod)[3:-1])))
    ch = tr[3, 0]
    n_c = int((ch > 0).sum()) // 6
    if n_c > 10:
        ch = ch[:6 * n_c].reshape(n_c, 6)     # (a) start (b) box free (c) columns in registers (d) staged (e) fenced (f) store issued
        sl = slice(8, n_c - 1)
        d = lambda i, j: np.median(us(ch[sl, j] - ch[sl, i]))
        print("epilogue chunk (us, median): wait box %.3f | wait TMEM load %.3f | scale + stage %.3f | fence %.3f | "
              "issue store %.3f | statistics + loop %.3f | total %.3f" %
              (d(0, 1), d(1, 2), d(2, 3), d(3, 4), d(4, 5), np.median(us(ch[9:n_c, 0] - ch[8:n_c - 1, 5])),
               np.median(us(np.diff(ch[:, 0])[sl]))))
