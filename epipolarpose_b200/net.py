"""PoseResNet forward / backward engine over the libepb.so kernels.

Data layout in HBM: activations are NHWC float32 pixel rows; each conv keeps
its RAW output z (pre-BatchNorm) plus per-channel (scale, shift) so that the
BatchNorm+ReLU of layer L is applied on the fly by the operand loader of layer
L+1 (and by wgrad of L+1) instead of a separate elementwise pass.  Only block
outputs (after the residual add) and the stem pool output are materialised.

Architecture restated from the reference lib/models/pose3d_resnet.py:91-212
(constructor :93-126, forward :185-212).  All compute goes through
epipolarpose_b200.ops (hand-written CUDA); torch only owns memory.
"""
import torch

from . import ops as _default_ops

BN_MOMENTUM = 0.1   # pose3d_resnet.py:8
BN_EPS = 1e-5

RESNET_SPEC = {18: ("basic", [2, 2, 2, 2]), 34: ("basic", [3, 4, 6, 3]),
               50: ("bottleneck", [3, 4, 6, 3]), 101: ("bottleneck", [3, 4, 23, 3]),
               152: ("bottleneck", [3, 8, 36, 3])}   # pose3d_resnet.py:288-292


def _pad(c, m):
    return (c + m - 1) // m * m


class Conv:
    """One Conv2d / ConvTranspose2d layer: geometry tables + packed weights."""

    def __init__(self, name, kind, cin, cout, k, stride, pad, opad=0, bias=False):
        self.name, self.kind = name, kind
        self.cin, self.cout, self.k, self.stride, self.pad, self.opad = cin, cout, k, stride, pad, opad
        self.bias = bias
        self.cin_p = _pad(cin, 4)
        self.cout_p = _pad(cout, 4)
        self._gcache = {}

    def out_hw(self, h, w):
        if self.kind == "conv":
            return ((h + 2 * self.pad - self.k) // self.stride + 1,
                    (w + 2 * self.pad - self.k) // self.stride + 1)
        return ((h - 1) * self.stride - 2 * self.pad + self.k + self.opad,
                (w - 1) * self.stride - 2 * self.pad + self.k + self.opad)

    # ---- geometry (see include/epb.h epb_conv_geom) -------------------------
    def _gather_geoms(self, ops, N, Hs, Ws, Cs, Hd, Wd, Cd, precision):
        """Strided gather: dst[i,j] = sum_t src[i*s + r - p, j*s + c - p] (conv
        fprop; deconv dgrad).  src=[N,Hs,Ws,Cs], dst=[N,Hd,Wd,Cd]."""
        k, s, p = self.k, self.stride, self.pad
        taps = [(r - p, c - p, r * k + c) for r in range(k) for c in range(k)]
        return [ops.make_geom(N, Hs, Ws, Cs, Hd, Wd, Cd, Hd, Wd, 1, 0, 0, s, taps, k * k,
                              precision=precision)]

    def _scatter_geoms(self, ops, N, Hs, Ws, Cs, Hd, Wd, Cd, precision):
        """Transposed: dst[i*s - p + r] += src[i] (deconv fprop; conv dgrad),
        one gather per output phase: dst[y] with y = i*s + a reads
        src[i + (a + p - r)/s] for the taps r = (a + p) mod s."""
        k, s, p = self.k, self.stride, self.pad
        geoms = []
        for a in range(s):
            for b in range(s):
                Hp = (Hd - a + s - 1) // s
                Wp = (Wd - b + s - 1) // s
                if Hp <= 0 or Wp <= 0:
                    continue
                taps = [((a + p - r) // s, (b + p - c) // s, r * k + c)
                        for r in range(k) for c in range(k)
                        if (a + p - r) % s == 0 and (b + p - c) % s == 0]
                if not taps:
                    geoms.append(None)   # phase receives nothing: stays zero
                    continue
                geoms.append(ops.make_geom(N, Hs, Ws, Cs, Hd, Wd, Cd, Hp, Wp, s, a, b, 1, taps,
                                           k * k, precision=precision))
        return geoms

    def fprop_geoms(self, ops, N, H, W, precision):
        key = ("f", id(ops), N, H, W, precision)
        if key not in self._gcache:           # geometry tables are static per shape: build once
            Ho, Wo = self.out_hw(H, W)
            f = self._gather_geoms if self.kind == "conv" else self._scatter_geoms
            self._gcache[key] = f(ops, N, H, W, self.cin_p, Ho, Wo, self.cout_p, precision)
        return self._gcache[key]

    def dgrad_geoms(self, ops, N, H, W, precision):
        """src = dOut [N,Ho,Wo,cout], dst = dIn [N,H,W,cin]."""
        key = ("d", id(ops), N, H, W, precision)
        if key not in self._gcache:
            Ho, Wo = self.out_hw(H, W)
            f = self._scatter_geoms if self.kind == "conv" else self._gather_geoms
            self._gcache[key] = f(ops, N, Ho, Wo, self.cout_p, H, W, self.cin_p, precision)
        return self._gcache[key]

    # ---- weights --------------------------------------------------------------
    def pack(self, ops, w):
        """state_dict weight -> (fprop operand [cout_p][T][cin_p], dgrad operand
        [cin_p][T][cout_p])."""
        T = self.k * self.k
        wf = torch.zeros(self.cout_p * T * self.cin_p, device=w.device, dtype=torch.float32)
        wd = torch.zeros(self.cin_p * T * self.cout_p, device=w.device, dtype=torch.float32)
        A, B = w.shape[0], w.shape[1]
        if self.kind == "conv":      # [O][I][k][k]
            ops.pack_weight(w, wf, A, B, self.k, self.k, 0, self.cin_p)
            ops.pack_weight(w, wd, A, B, self.k, self.k, 1, self.cout_p)
        else:                        # [I][O][k][k]
            ops.pack_weight(w, wf, A, B, self.k, self.k, 1, self.cin_p)
            ops.pack_weight(w, wd, A, B, self.k, self.k, 0, self.cout_p)
        return wf, wd

    def unpack_grad(self, ops, dwp, grad_out):
        """packed fprop-operand gradient -> state_dict layout."""
        A, B = grad_out.shape[0], grad_out.shape[1]
        swap = 0 if self.kind == "conv" else 1
        ops.pack_weight(dwp, grad_out, A, B, self.k, self.k, swap, self.cin_p, 1)


class PoseNetPlan:
    """Static description of the network (layer list in state_dict order)."""

    def __init__(self, num_layers=50, num_joints=17, volume=True, depth_res=64,
                 image_size=(256, 256), deconv_filters=(256, 256, 256), deconv_kernels=(4, 4, 4),
                 deconv_with_bias=False, final_kernel=1):
        self.kind, self.layers = RESNET_SPEC[num_layers]
        self.num_joints, self.volume, self.depth_res = num_joints, volume, depth_res
        self.image_size = tuple(image_size)
        self.exp = 4 if self.kind == "bottleneck" else 1
        self.stem = Conv("conv1", "conv", 3, 64, 7, 2, 3)
        # the stem runs as a 1x1 conv over its patch matrix (K = 7*7*3 = 147 -> 160)
        self.stem_kpad = _pad(7 * 7 * 3, 32)
        self.stem_col = Conv("conv1", "conv", self.stem_kpad, 64, 1, 1, 0)
        self.blocks = []
        inpl = 64
        for li, (planes, nb) in enumerate(zip((64, 128, 256, 512), self.layers)):
            stride = 1 if li == 0 else 2
            for b in range(nb):
                p = "layer%d.%d" % (li + 1, b)
                s = stride if b == 0 else 1
                blk = {"name": p, "convs": [], "bns": [], "down": None}
                if self.kind == "bottleneck":    # pose3d_resnet.py:50-66
                    blk["convs"] = [Conv(p + ".conv1", "conv", inpl, planes, 1, 1, 0),
                                    Conv(p + ".conv2", "conv", planes, planes, 3, s, 1),
                                    Conv(p + ".conv3", "conv", planes, planes * 4, 1, 1, 0)]
                    blk["bns"] = [(p + ".bn1", planes), (p + ".bn2", planes), (p + ".bn3", planes * 4)]
                else:                             # :19-29
                    blk["convs"] = [Conv(p + ".conv1", "conv", inpl, planes, 3, s, 1),
                                    Conv(p + ".conv2", "conv", planes, planes, 3, 1, 1)]
                    blk["bns"] = [(p + ".bn1", planes), (p + ".bn2", planes)]
                if b == 0 and (s != 1 or inpl != planes * self.exp):   # :130-135
                    blk["down"] = (Conv(p + ".downsample.0", "conv", inpl, planes * self.exp, 1, s, 0),
                                   (p + ".downsample.1", planes * self.exp))
                inpl = planes * self.exp
                self.blocks.append(blk)
        self.trunk_channels = inpl
        self.deconvs = []
        for i, (nf, k) in enumerate(zip(deconv_filters, deconv_kernels)):   # :158-183
            kk, pad, opad = {4: (4, 1, 0), 3: (3, 1, 1), 2: (2, 0, 0)}[k]
            self.deconvs.append((Conv("deconv_layers.%d" % (3 * i), "deconv", inpl, nf, kk, 2, pad,
                                      opad, bias=deconv_with_bias),
                                 ("deconv_layers.%d" % (3 * i + 1), nf)))
            inpl = nf
        out_ch = num_joints * depth_res if volume else num_joints      # :116-122
        self.final = Conv("final_layer", "conv", inpl, out_ch, final_kernel, 1,
                          1 if final_kernel == 3 else 0, bias=True)
        self.fc = None
        if not volume:                                                  # :124-126
            self.fc = Conv("depth_fc", "conv", 2048, num_joints * depth_res, 1, 1, 0, bias=True)
            self.pool_k = int(image_size[0] / 2 ** 5)

    def all_convs(self):
        out = [self.stem]
        for blk in self.blocks:
            out += blk["convs"]
            if blk["down"]:
                out.append(blk["down"][0])
        out += [d[0] for d in self.deconvs] + [self.final]
        if self.fc:
            out.append(self.fc)
        return out

    def all_bns(self):
        out = [("bn1", 64)]
        for blk in self.blocks:
            out += blk["bns"]
            if blk["down"]:
                out.append(blk["down"][1])
        out += [d[1] for d in self.deconvs]
        return out


# Gradient stages, in the order the backward pass completes them (parameters of a stage are
# contiguous in the flat gradient buffer): the data-parallel all-reduce of a stage is issued as
# soon as its weight gradients are unpacked and overlaps the rest of the backward pass.
STAGES = (("deconv_layers", "final_layer", "depth_fc"), ("layer4",), ("layer3",), ("layer2",),
          ("layer1", "conv1", "bn1"))


def stage_of(name):
    for k, prefixes in enumerate(STAGES):
        if any(name == p or name.startswith(p + ".") for p in prefixes):
            return k
    raise KeyError(name)


class _BNState:
    __slots__ = ("name", "C", "scale", "shift", "mean", "invstd", "M")


class Engine:
    """Executes a PoseNetPlan.  `params` maps state_dict names to tensors."""

    def __init__(self, plan, precision=0, ops=None, wgrad_precision=None):
        self.plan = plan
        self.precision = precision
        # Weight gradients are leaf outputs: their rounding error is not propagated through
        # further layers (no ReLU-mask flips downstream), and it averages over the N*H*W
        # reduction.  A separate precision can therefore be chosen for wgrad
        # (EPB_WGRAD_PRECISION=tf32|tf32x3|fp32; default: same as `precision`).
        import os
        env = os.environ.get("EPB_WGRAD_PRECISION")
        if wgrad_precision is None and env:
            wgrad_precision = {"fp32": 0, "tf32": 1, "tf32x3": 3}[env]
        self.wgrad_precision = precision if wgrad_precision is None else wgrad_precision
        if precision == 0:
            self.wgrad_precision = 0
        self.ops = ops or _default_ops
        # K of the stem's patch matrix (Engine16: 192); plan is None for bare conv helpers
        self.stem_kpad = plan.stem_kpad if plan is not None else None
        self.stem_col = plan.stem_col if plan is not None else None

    # ------------------------------------------------------------------ helpers
    def _pack_weights(self, params):
        """state_dict weights -> persistent GEMM operands {name: (fprop, dgrad)} with ONE
        batched conversion launch (epb_pack_weight_batch).  The buffers and the job table are
        built once per parameter storage; padding rows / columns stay zero."""
        plan, ops = self.plan, self.ops
        convs = plan.all_convs()
        key = tuple(params[c.name + ".weight"].data_ptr() for c in convs) + (str(self.dev),)
        st = getattr(self, "_wstate", None)
        if st is None or st["key"] != key:
            packed, jobs = {}, []
            for conv in convs:
                w = params[conv.name + ".weight"]
                if conv is plan.stem:
                    kpad = self.stem_kpad
                    wcol = torch.zeros(64 * kpad, device=self.dev, dtype=torch.float32)
                    jobs.append((w, wcol, 64, 3, 49, 0, 3, 0, kpad))
                    packed[conv.name] = (wcol, None)
                    continue
                T = conv.k * conv.k
                A, B = w.shape[0], w.shape[1]
                wd = torch.zeros(conv.cin_p * T * conv.cout_p, device=self.dev, dtype=torch.float32)
                sf, sd = (0, 1) if conv.kind == "conv" else (1, 0)
                if self._aliases_param(conv, w):
                    wf = w.detach().reshape(-1)          # [Cout][Cin] IS the packed fprop operand of a 1x1 conv
                else:
                    wf = torch.zeros(conv.cout_p * T * conv.cin_p, device=self.dev, dtype=torch.float32)
                    jobs.append((w, wf, A, B, T, sf, conv.cin_p, 0, T * conv.cin_p))
                jobs.append((w, wd, A, B, T, sd, conv.cout_p, 0, T * conv.cout_p))
                packed[conv.name] = (wf, wd)
            st = {"key": key, "packed": packed, "batch": ops.PackBatch(jobs)}
            self._wstate = st
        ops.pack_weight_batch(st["batch"])
        return st["packed"]

    def _aliases_param(self, conv, t):
        """A pad-free 1x1 convolution's [Cout][Cin][1][1] tensor is already the packed fprop operand
        [Cout][T = 1][Cin] (and its gradient the packed weight gradient): engines that read the fp32
        operand with plain loads (alias_1x1) skip the pack / unpack of those layers."""
        plan = self.plan
        if conv is plan.fc or conv is plan.final:        # may run on the fp32-operand kernels
            return False
        return (getattr(self, "alias_1x1", False) and conv.kind == "conv" and conv.k == 1
                and conv.cin_p == t.shape[1] and conv.cout_p == t.shape[0] and t.is_contiguous()
                and t.data_ptr() % 16 == 0)

    def _grad_state(self, grads):
        """Persistent backward scratch: packed weight-gradient accumulators (one flat buffer,
        one memset per step), the BatchNorm-backward sums, and the batched unpack into the
        state_dict-shaped `grads`."""
        plan, ops = self.plan, self.ops
        convs = plan.all_convs()
        key = tuple(grads[c.name + ".weight"].data_ptr() for c in convs) + (str(self.dev),)
        st = getattr(self, "_gstate", None)
        if st is None or st["key"] != key:
            sizes = []
            for conv in convs:
                if conv is plan.stem:
                    sizes.append(64 * self.stem_kpad)
                else:
                    sizes.append(conv.cout_p * conv.k * conv.k * conv.cin_p)
            flat = torch.zeros(sum(sizes), device=self.dev, dtype=torch.float32)
            dwp, jobs, off = {}, [[] for _ in STAGES], 0
            for conv, n in zip(convs, sizes):
                view = flat[off:off + n]
                off += n
                g = grads[conv.name + ".weight"]
                if conv is not plan.stem and self._aliases_param(conv, g):
                    dwp[conv.name] = g.reshape(-1)       # the weight gradient lands in the state_dict layout
                    continue
                dwp[conv.name] = view
                js = jobs[stage_of(conv.name)]
                if conv is plan.stem:
                    js.append((view, g, 64, 3, 49, 0, 3, 1, self.stem_kpad))
                else:
                    T = conv.k * conv.k
                    js.append((view, g, g.shape[0], g.shape[1], T, 0 if conv.kind == "conv" else 1,
                               conv.cin_p, 1, T * conv.cin_p))
            bns = plan.all_bns()
            sums = torch.zeros(sum(2 * C for _, C in bns), device=self.dev, dtype=torch.float64)
            bsum, off = {}, 0
            for name, C in bns:
                bsum[name] = sums[off:off + 2 * C]
                off += 2 * C
            st = {"key": key, "flat": flat, "dwp": dwp, "sums": sums, "bsum": bsum,
                  "batches": [ops.PackBatch(j) if j else None for j in jobs]}
            self._gstate = st
        return st

    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(shape, device=self.dev, dtype=dtype)

    def _conv_fwd(self, conv, x, N, H, W, wf, affine=None, bias=None, stats=None, relu=1):
        ops = self.ops
        Ho, Wo = conv.out_hw(H, W)
        geoms = conv.fprop_geoms(ops, N, H, W, self.precision)
        need_zero = any(g is None for g in geoms)
        out = (torch.zeros if need_zero else torch.empty)(
            (N, Ho, Wo, conv.cout_p), device=self.dev, dtype=torch.float32)
        sc, sh = affine if affine is not None else (None, None)
        for g in geoms:
            if g is None:
                continue
            g.in_relu = relu if affine is not None else 0
            g.accumulate = 0
            ops.conv_fprop(g, x, wf, out, sc, sh, bias, stats)
        return out, Ho, Wo

    def _conv_dgrad(self, conv, dout, N, H, W, wd, accumulate_into=None):
        """dIn [N,H,W,cin_p] from dOut; H,W are the conv INPUT dims."""
        ops = self.ops
        geoms = conv.dgrad_geoms(ops, N, H, W, self.precision)
        if accumulate_into is not None:
            din = accumulate_into
        else:
            need_zero = any(g is None for g in geoms)
            din = (torch.zeros if need_zero else torch.empty)(
                (N, H, W, conv.cin_p), device=self.dev, dtype=torch.float32)
        for g in geoms:
            if g is None:
                continue
            g.accumulate = 1 if accumulate_into is not None else 0
            g.in_relu = 0
            ops.conv_fprop(g, dout, wd, din, None, None, None, None)
        return din

    def _conv_wgrad(self, conv, x, dout, N, H, W, grad_out, affine=None, relu=1, post=None):
        """Weight gradient.  Off the critical path of the backward pass (nothing downstream
        consumes it before the optimiser), so it is issued on a side stream and overlaps the
        next layers' dgrad / BatchNorm-backward kernels; `backward` joins the streams.
        `post` (optional) runs right after it ON THE SAME STREAM (layout fix-ups of the
        gradient)."""
        side = getattr(self, "_side", None)
        if side is None:
            self._conv_wgrad_now(conv, x, dout, N, H, W, grad_out, affine, relu)
            if post is not None:
                post()
            return
        main = torch.cuda.current_stream()
        side.wait_stream(main)                    # dout / x are ready on the main stream
        self._keep.append((x, dout, affine, grad_out))   # keep operands alive until the join
        with torch.cuda.stream(side):
            self._conv_wgrad_now(conv, x, dout, N, H, W, grad_out, affine, relu)
            if post is not None:
                post()

    def _conv_wgrad_now(self, conv, x, dout, N, H, W, grad_out, affine=None, relu=1):
        ops = self.ops
        T = conv.k * conv.k
        gs = getattr(self, "_gs", None)
        batched = gs is not None and gs["dwp"].get(conv.name) is not None and \
            gs["dwp"][conv.name].numel() == conv.cout_p * T * conv.cin_p
        # inside backward(): accumulate into the step's flat buffer (zeroed once, unpacked in
        # one launch at the end); stand-alone calls convert immediately
        dwp = gs["dwp"][conv.name] if batched else \
            torch.zeros(conv.cout_p * T * conv.cin_p, device=self.dev, dtype=torch.float32)
        sc, sh = affine if affine is not None else (None, None)
        for g in conv.fprop_geoms(ops, N, H, W, self.precision):
            if g is None:
                continue
            g.in_relu = relu if affine is not None else 0
            g.accumulate = 0
            g.precision = self.wgrad_precision
            ops.conv_wgrad(g, x, dout, dwp, sc, sh)
            g.precision = self.precision
        if not batched:
            conv.unpack_grad(ops, dwp, grad_out)

    def _bn_train(self, name, C, stats, M, params, new_buffers, act_scale=None):
        """act_scale = (stats2, scale2, shift2, res_sc, sc): also publish the scale of the layer's
        post-activation split tensor in the same launch (split path)."""
        ops = self.ops
        st = _BNState()
        st.name, st.C, st.M = name, C, M
        st.scale, st.shift, st.mean, st.invstd = (self._new(C), self._new(C), self._new(C),
                                                  self._new(C))
        rm = params[name + ".running_mean"]
        rv = params[name + ".running_var"]
        if act_scale is None:
            ops.bn_finalize(stats, M, C, params[name + ".weight"], params[name + ".bias"], BN_EPS,
                            BN_MOMENTUM, rm, rv, st.scale, st.shift, st.mean, st.invstd)
        else:
            ops.bn_finalize_scale(stats, M, C, params[name + ".weight"], params[name + ".bias"], BN_EPS,
                                  BN_MOMENTUM, rm, rv, st.scale, st.shift, st.mean, st.invstd, *act_scale)
        nbt = params.get(name + ".num_batches_tracked")
        if nbt is not None:
            tick = getattr(self, "_nbt_tick", None)
            if tick is None:
                nbt += 1
            else:
                tick.append(nbt)            # one multi-tensor increment at the end of the forward
        return st

    def _bn_eval(self, name, C, params):
        st = _BNState()
        st.name, st.C, st.M = name, C, 0
        st.scale, st.shift = self._new(C), self._new(C)
        st.mean = st.invstd = None
        self.ops.bn_eval_affine(C, params[name + ".weight"], params[name + ".bias"],
                                params[name + ".running_mean"], params[name + ".running_var"],
                                BN_EPS, st.scale, st.shift)
        return st

    # ------------------------------------------------------------------ forward
    def forward(self, x_nchw, params, training=True, save=True):
        """x_nchw [N,3,H,W] float32 contiguous.  Returns (logits_nhwc, depth or
        None, saved) where logits_nhwc is [N,H/4,W/4,cout_p]."""
        ops, plan = self.ops, self.plan
        self.dev = x_nchw.device
        N, _, H, W = x_nchw.shape
        S = {"N": N, "H": H, "W": W, "packed": {}, "bn": {}, "blocks": []}
        # per-forward BN statistics accumulators (one memset)
        bns = plan.all_bns()
        offs, tot = {}, 0
        for name, C in bns:
            offs[name] = tot
            tot += 2 * C
        stats_all = torch.zeros(tot, device=self.dev, dtype=torch.float64) if training else None

        def stats_of(name, C):
            return stats_all[offs[name]:offs[name] + 2 * C] if training else None

        def bn(name, C, M):
            st = self._bn_train(name, C, stats_of(name, C), M, params, None) if training \
                else self._bn_eval(name, C, params)
            S["bn"][name] = st
            return st

        S["packed"] = self._pack_weights(params)     # every layer's operands, one launch

        def packed(conv):
            return S["packed"][conv.name][0]

        # ---- stem (pose3d_resnet.py:186-189)
        stem, scol, kpad = plan.stem, self.stem_col, self.stem_kpad
        x = self._new(N, H, W, stem.cin_p)
        ops.nchw_to_nhwc(x_nchw, x, N, 3, H, W, stem.cin_p)
        H1, W1 = stem.out_hw(H, W)
        col = self._new(N, H1, W1, kpad)
        ops.im2col(x, col, N, H, W, stem.cin_p, 3, 7, 7, 2, 3, H1, W1, kpad)
        # conv1.weight [64,3,7,7] -> [64][(r,s,c)] zero-padded to [64][kpad]: the fprop operand
        # of the 1x1 conv over the patch matrix (packed with the other layers)
        z0, _, _ = self._conv_fwd(scol, col, N, H1, W1, packed(stem), stats=stats_of("bn1", 64))
        b0 = bn("bn1", 64, N * H1 * W1)
        H2, W2 = (H1 + 2 - 3) // 2 + 1, (W1 + 2 - 3) // 2 + 1
        cur = self._new(N, H2, W2, 64)
        argidx = self._new(N, H2, W2, 64, dtype=torch.uint8)
        ops.bn_relu_maxpool(z0, b0.scale, b0.shift, cur, argidx, N, H1, W1, 64)
        S["stem"] = (col, z0, argidx, H1, W1, H2, W2)
        h, w = H2, W2

        # ---- residual stages (:191-194)
        for blk in plan.blocks:
            rec = {"in": cur, "h": h, "w": w, "z": [], "hw": []}
            src, aff = cur, None
            hh, ww = h, w
            for ci, conv in enumerate(blk["convs"]):
                bname, C = blk["bns"][ci]
                z, ho, wo = self._conv_fwd(conv, src, N, hh, ww, packed(conv), affine=aff,
                                           stats=stats_of(bname, C))
                st = bn(bname, C, N * ho * wo)
                rec["z"].append(z)
                rec["hw"].append((hh, ww))
                src, aff = z, (st.scale, st.shift)
                hh, ww = ho, wo
            last = S["bn"][blk["bns"][-1][0]]
            out = self._new(N, hh, ww, blk["convs"][-1].cout_p)
            M = N * hh * ww
            if blk["down"]:
                dconv, (dname, dC) = blk["down"]
                zd, _, _ = self._conv_fwd(dconv, cur, N, h, w, packed(dconv),
                                          stats=stats_of(dname, dC))
                dst = bn(dname, dC, M)
                rec["zd"] = zd
                ops.bn_act(src, last.scale, last.shift, zd, dst.scale, dst.shift, 1, out, M,
                           out.shape[-1])
            else:
                ops.bn_act(src, last.scale, last.shift, cur, None, None, 1, out, M, out.shape[-1])
            rec["out"] = out
            S["blocks"].append(rec)
            cur, h, w = out, hh, ww

        S["trunk"] = (cur, h, w)
        # ---- deconv head (:198) : BN+ReLU of each deconv fused into the next loader
        src, aff = cur, None
        S["deconv"] = []
        for conv, (bname, C) in plan.deconvs:
            bias = params.get(conv.name + ".bias")
            z, ho, wo = self._conv_fwd(conv, src, N, h, w, packed(conv), affine=aff, bias=bias,
                                       stats=stats_of(bname, C))
            st = bn(bname, C, N * ho * wo)
            S["deconv"].append((src, aff, z, h, w))
            src, aff, h, w = z, (st.scale, st.shift), ho, wo
        # ---- final 1x1 / 3x3 conv with bias (:199)
        fin = plan.final
        fbias = params[fin.name + ".bias"]
        if fin.cout_p != fin.cout:
            fb = torch.zeros(fin.cout_p, device=self.dev)
            fb[:fin.cout] = fbias
            fbias = fb
        logits, ho, wo = self._conv_fwd(fin, src, N, h, w, packed(fin), affine=aff, bias=fbias)
        S["final"] = (src, aff, h, w)
        depth = None
        if plan.fc is not None:                 # :202-210
            tr, th, tw = S["trunk"]
            assert th == plan.pool_k and tw == plan.pool_k, "AvgPool(k) -> 1x1 expected"
            pooled = self._new(N, 1, 1, 2048)
            ops.avgpool(tr, pooled, N, th * tw, 2048)
            depth, _, _ = self._conv_fwd(plan.fc, pooled, N, 1, 1, packed(plan.fc),
                                         bias=params["depth_fc.bias"])
            S["fc"] = pooled
        return logits, depth, (S if save else None)

    # ------------------------------------------------------------------ backward
    def _bn_bwd(self, st, dy, z, y_out, relu, params, grads):
        """Returns dz.  Fills grads[name.weight/.bias]."""
        ops = self.ops
        C = st.C
        M = z.numel() // C
        gs = getattr(self, "_gs", None)
        sums = gs["bsum"][st.name] if gs is not None else \
            torch.zeros(2 * C, device=self.dev, dtype=torch.float64)
        ops.bn_bwd_reduce(dy, z, y_out, st.scale, st.shift, st.mean, st.invstd, relu, M, C, sums)
        dz = torch.empty_like(z)
        ops.bn_bwd_apply(dy, z, y_out, st.scale, st.shift, st.mean, st.invstd,
                         params[st.name + ".weight"], relu, sums, M, C, dz,
                         grads[st.name + ".weight"], grads[st.name + ".bias"])
        return dz

    def _stage_done(self, k):
        """Every gradient of stage k has been issued: unpack its packed weight gradients into the
        state_dict-shaped buffers (on the weight-gradient stream, behind them) and hand the stage
        to the caller (the data-parallel all-reduce of its slice of the flat buffer)."""
        if k in self._stages_done:
            return
        self._stages_done.add(k)
        batch = self._gs["batches"][k]
        side = self._side
        if side is None:
            if batch is not None:
                self.ops.pack_weight_batch(batch)
            if self._on_stage is not None:
                self._on_stage(k)
            return
        side.wait_stream(torch.cuda.current_stream())       # BatchNorm gradients of the stage (main stream)
        with torch.cuda.stream(side):
            if batch is not None:
                self.ops.pack_weight_batch(batch)
            if self._on_stage is not None:
                self._on_stage(k)

    def backward(self, S, dlogits, ddepth, params, grads, on_stage=None):
        """dlogits [N,Ho,Wo,cout_p] NHWC contiguous.  grads: dict name -> tensor
        (state_dict shape) to be filled for every trainable parameter.  on_stage(k) is called
        (with the weight-gradient stream current) as soon as stage k of net.STAGES is complete."""
        ops, plan = self.ops, self.plan
        N = S["N"]
        anchor = dlogits if dlogits is not None else self._head.planes    # split path: gradient sink
        self.dev = anchor.device
        import os
        self._side, self._keep = None, []
        if anchor.is_cuda and os.environ.get("EPB_OVERLAP_WGRAD", "1") != "0":
            if getattr(self, "_side_stream", None) is None:
                self._side_stream = torch.cuda.Stream()
            self._side = self._side_stream
        self._gs = self._grad_state(grads)
        self._gs["flat"].zero_()
        self._gs["sums"].zero_()
        self._on_stage, self._stages_done = on_stage, set()
        try:
            self._backward(S, dlogits, ddepth, params, grads)
            for k in range(len(STAGES)):
                self._stage_done(k)                                   # whatever is still open
        finally:
            if self._side is not None:
                torch.cuda.current_stream().wait_stream(self._side)   # join before grads are used
            self._side, self._keep, self._gs, self._on_stage = None, [], None, None

    def _backward(self, S, dlogits, ddepth, params, grads):
        ops, plan = self.ops, self.plan
        N = S["N"]

        def wd_of(conv):
            return S["packed"][conv.name][1]

        # ---- final layer
        fin = plan.final
        src, aff, h, w = S["final"]
        Ho, Wo = fin.out_hw(h, w)
        gb = grads[fin.name + ".bias"]
        if fin.cout_p != fin.cout:
            tmp = torch.empty(fin.cout_p, device=self.dev)
            ops.colsum(dlogits, N * Ho * Wo, fin.cout_p, tmp)
            gb.copy_(tmp[:fin.cout])
        else:
            ops.colsum(dlogits, N * Ho * Wo, fin.cout_p, gb)
        self._conv_wgrad(fin, src, dlogits, N, h, w, grads[fin.name + ".weight"], affine=aff)
        dcur = self._conv_dgrad(fin, dlogits, N, h, w, wd_of(fin))
        # ---- deconv head, reversed
        for (conv, (bname, C)), (dsrc, daff, z, dh, dw) in zip(reversed(plan.deconvs),
                                                               reversed(S["deconv"])):
            st = S["bn"][bname]
            dz = self._bn_bwd(st, dcur, z, None, 1, params, grads)
            if conv.bias:
                oh, ow = conv.out_hw(dh, dw)
                ops.colsum(dz, N * oh * ow, conv.cout_p, grads[conv.name + ".bias"])
            self._conv_wgrad(conv, dsrc, dz, N, dh, dw, grads[conv.name + ".weight"], affine=daff)
            dcur = self._conv_dgrad(conv, dz, N, dh, dw, wd_of(conv))
        # ---- VOLUME=False depth head
        if plan.fc is not None and ddepth is not None:
            tr, th, tw = S["trunk"]
            dd = ddepth.reshape(N, 1, 1, -1).contiguous()
            ops.colsum(dd, N, plan.fc.cout_p, grads["depth_fc.bias"])
            self._conv_wgrad(plan.fc, S["fc"], dd, N, 1, 1, grads["depth_fc.weight"])
            dpool = self._conv_dgrad(plan.fc, dd, N, 1, 1, wd_of(plan.fc))
            ops.avgpool_bwd(dpool, dcur, N, th * tw, 2048, 1)
        self._stage_done(0)                     # head (deconvs, final layer, depth_fc) complete
        # ---- residual stages, reversed
        prev_stage = None
        for blk, rec in zip(reversed(plan.blocks), reversed(S["blocks"])):
            sk = stage_of(blk["name"])
            if prev_stage is not None and sk != prev_stage:
                self._stage_done(prev_stage)
            prev_stage = sk
            out, xin, h, w = rec["out"], rec["in"], rec["h"], rec["w"]
            nconv = len(blk["convs"])
            dres = None
            if blk["down"]:
                dconv, (dname, dC) = blk["down"]
                dzd = self._bn_bwd(S["bn"][dname], dcur, rec["zd"], out, 0, params, grads)
            g = dcur
            y_out = out
            for ci in range(nconv - 1, -1, -1):
                conv = blk["convs"][ci]
                st = S["bn"][blk["bns"][ci][0]]
                z = rec["z"][ci]
                # last BN of the block: mask = (block output > 0); inner BNs: own ReLU
                dz = self._bn_bwd(st, g, z, y_out, 0 if y_out is not None else 1, params, grads)
                y_out = None
                hh, ww = rec["hw"][ci]
                if ci == 0:
                    self._conv_wgrad(conv, xin, dz, N, hh, ww, grads[conv.name + ".weight"])
                else:
                    pst = S["bn"][blk["bns"][ci - 1][0]]
                    self._conv_wgrad(conv, rec["z"][ci - 1], dz, N, hh, ww,
                                     grads[conv.name + ".weight"], affine=(pst.scale, pst.shift))
                g = self._conv_dgrad(conv, dz, N, hh, ww, wd_of(conv))
            if blk["down"]:
                self._conv_wgrad(dconv, xin, dzd, N, h, w, grads[dconv.name + ".weight"])
                self._conv_dgrad(dconv, dzd, N, h, w, wd_of(dconv), accumulate_into=g)
                dcur = g
            else:
                nd = torch.empty_like(g)
                ops.add_masked(g, dcur, out, nd, g.numel())
                dcur = nd
        # ---- stem
        col, z0, argidx, H1, W1, H2, W2 = S["stem"]
        gpool = self._new(N, H1, W1, 64)
        ops.maxpool_bwd(dcur, argidx, gpool, N, H1, W1, 64)
        dz0 = self._bn_bwd(S["bn"]["bn1"], gpool, z0, None, 1, params, grads)
        # the stem's gradient w.r.t. the [64][kpad] patch-matrix operand lands in the flat
        # buffer; the batched unpack maps its first 147 columns back to [64,3,7,7]
        self._conv_wgrad(self.stem_col, col, dz0, N, H1, W1, None)
