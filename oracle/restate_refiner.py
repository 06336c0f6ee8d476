"""TEST INFRASTRUCTURE ONLY -- functional torch-CPU restatement of the refiner MLP (reference
refiner/model.py:15-143) as a pure function of a state_dict, used as the oracle of
epipolarpose_b200/refiner (pinned against the unmodified module through
tests/golden/refiner.npz).  Dropout is an explicit list of keep masks (None = no dropout)."""
import collections

import torch
import torch.nn.functional as F


def param_shapes(linear_size=1024, input_size=45, output_size=45):
    """Ordered {state_dict key: shape} of LinearModelPG (refiner/model.py:71-115)."""
    d = collections.OrderedDict()

    def lin(p, o, i):
        d[p + ".weight"] = (o, i)
        d[p + ".bias"] = (o,)

    def bn(p, c):
        d[p + ".weight"] = (c,)
        d[p + ".bias"] = (c,)
        d[p + ".running_mean"] = (c,)
        d[p + ".running_var"] = (c,)
        d[p + ".num_batches_tracked"] = ()
    for s in range(2):
        for i in (1, 2, 3, 4):
            lin("linear_stages.%d.w%d" % (s, i), linear_size, linear_size)
        for i in (1, 2, 3, 4):
            bn("linear_stages.%d.batch_norm%d" % (s, i), linear_size)
    lin("w1", linear_size, input_size)
    lin("w2", output_size, linear_size)
    lin("w3", linear_size, output_size)
    lin("w4", output_size, linear_size)
    bn("batch_norm1", linear_size)
    bn("batch_norm3", linear_size)
    return d


def init_state(shapes, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = collections.OrderedDict()
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shp, generator=g)
        elif "batch_norm" in k and k.endswith(".weight"):
            sd[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        else:
            sd[k] = torch.randn(shp, generator=g) * (2.0 / shp[1]) ** 0.5     # kaiming_normal_ scale
    return sd


def forward(sd, x, training=True, masks=None, p_dropout=0.0, new_stats=None):
    """-> (p1, p2).  masks: iterator of keep masks consumed in call order (10 dropout sites)."""
    it = iter(masks) if masks is not None else None

    def block(p, i, h):
        h = F.linear(h, sd["%sw%d.weight" % (p, i)], sd["%sw%d.bias" % (p, i)])
        n = "%sbatch_norm%d" % (p, i)
        rm, rv = sd[n + ".running_mean"].clone(), sd[n + ".running_var"].clone()
        h = F.batch_norm(h, rm, rv, sd[n + ".weight"], sd[n + ".bias"], training, 0.1, 1e-5)
        if new_stats is not None:
            new_stats[n + ".running_mean"], new_stats[n + ".running_var"] = rm, rv
        h = torch.relu(h)
        if it is not None and training:
            h = h * next(it).to(h.dtype) / (1.0 - p_dropout)
        return h

    def stage(p, h):
        y = block(p, 2, block(p, 1, h))
        out = h + y
        y = block(p, 4, block(p, 3, out))
        return out + y
    inp = block("", 1, x)
    s1 = stage("linear_stages.0.", inp)
    p1 = F.linear(s1, sd["w2.weight"], sd["w2.bias"])
    y = block("", 3, p1)
    y = s1 + y + inp
    y = stage("linear_stages.1.", y)
    y = inp + y
    p2 = F.linear(y, sd["w4.weight"], sd["w4.bias"])
    return p1, p2
