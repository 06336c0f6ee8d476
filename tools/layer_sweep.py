"""Median time of a fixed set of bench layers (R50, 128 images) in ONE process: A/B target for
library variants (EPB_LIB_PATH=build/variants/libepb_<name>.so python tools/layer_sweep.py).

    python tools/layer_sweep.py [reps]
"""
import os, sys
os.environ.setdefault("EPB_OVERLAP_WGRAD", "0")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarpose_b200 import net, ops

LAYERS = [  # N H W Cin Cout k stride what
    (128, 64, 64, 64, 256, 1, 1, "fprop"), (128, 64, 64, 256, 64, 1, 1, "fprop"),
    (128, 64, 64, 64, 64, 3, 1, "fprop"), (128, 32, 32, 128, 512, 1, 1, "fprop"),
    (128, 32, 32, 512, 128, 1, 1, "fprop"), (128, 16, 16, 256, 1024, 1, 1, "fprop"),
    (128, 16, 16, 1024, 256, 1, 1, "fprop"), (128, 16, 16, 256, 256, 3, 1, "fprop"),
    (128, 64, 64, 256, 1024, 1, 1, "dgrad"), (128, 64, 64, 256, 1024, 1, 1, "fprop"),
    (128, 32, 32, 128, 128, 3, 1, "fprop"), (128, 8, 8, 512, 2048, 1, 1, "fprop"),
    (128, 8, 8, 512, 512, 3, 1, "fprop"), (128, 64, 64, 256, 128, 1, 2, "fprop"),
    (128, 64, 64, 64, 64, 3, 1, "wgrad"), (128, 16, 16, 256, 256, 3, 1, "wgrad"),
    (128, 64, 64, 256, 1024, 1, 1, "wgrad"), (128, 64, 64, 64, 256, 1, 1, "wgrad"),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
dev = torch.device("cuda:0")
eng = net.Engine(None, precision=3)
eng.dev = dev
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("lib", os.environ.get("EPB_LIB_PATH", "in-tree"))
tot = 0.0
for (N, H, W, cin, cout, k, s, what) in LAYERS:
    conv = net.Conv("t", "conv", cin, cout, k, s, k // 2, 0)
    torch.manual_seed(0)
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    x = torch.randn(N, H, W, conv.cin_p, device=dev)
    sc, sh = torch.rand(conv.cin_p, device=dev) + 0.5, torch.randn(conv.cin_p, device=dev) * 0.1
    wf, wd = conv.pack(ops, w)
    Ho, Wo = conv.out_hw(H, W)
    gout = torch.randn(N, Ho, Wo, conv.cout_p, device=dev)
    gw = torch.zeros_like(w)

    def run():
        if what == "fprop":
            eng._conv_fwd(conv, x, N, H, W, wf, affine=(sc, sh))
        elif what == "dgrad":
            eng._conv_dgrad(conv, gout, N, H, W, wd)
        else:
            eng._conv_wgrad(conv, x, gout, N, H, W, gw, affine=(sc, sh))

    run(); run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    tot += t
    fl = 2.0 * N * Ho * Wo * cin * cout * k * k
    print("%-5s %3dx%-3d %4d->%-4d k%d s%d  %.3f ms  %.1f TF/s" % (what, H, W, cin, cout, k, s, t, fl / t / 1e9))
    del x, gout, w, wf, wd
print("total %.3f ms" % tot)
