"""Refiner network -- host-side mirror of the reference refiner/model.py surface
(weight_init :8-12, LinearPG :15-68, LinearModelPG :71-143, get_model :146-150): same class names,
constructor arguments, state_dict keys / shapes / default initialisation (the sub-modules are
torch.nn.Linear / BatchNorm1d objects used as PARAMETER CONTAINERS only), same forward signature
`model(x [N, input_size]) -> (p1, p2)`.  The arithmetic runs through epipolarpose_b200.mlp.MLPEngine
on the libepb.so kernels.  Not built: leaky=True, bn=False."""
import torch
import torch.nn as nn

from epipolarpose_b200 import mlp as _mlp


def weight_init(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight)


class LinearPG(nn.Module):
    """Parameter container of one residual stage (reference :15-37); evaluated by LinearModelPG."""

    def __init__(self, linear_size, p_dropout=0.5, bias=True, bn=True, leaky=False):
        super().__init__()
        if leaky or not bn or not bias:
            raise NotImplementedError("refiner stages are built for bias=True, bn=True, leaky=False")
        self.l_size, self.bn, self.leaky = linear_size, bn, leaky
        for i in (1, 2, 3, 4):
            setattr(self, "w%d" % i, nn.Linear(linear_size, linear_size, bias=bias))
        for i in (1, 2, 3, 4):
            setattr(self, "batch_norm%d" % i, nn.BatchNorm1d(linear_size))

    def forward(self, x):
        raise RuntimeError("LinearPG is evaluated inside LinearModelPG (one fused graph)")


class _RefinerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        names = module._param_names
        p = dict(zip(names, params))
        p.update(module._buffers_flat())
        eng = module._engine()
        need = any(ctx.needs_input_grad)
        with torch.no_grad():
            p1, p2, record = eng.forward(x.contiguous().float(), p, module.training, module.p_dropout,
                                         want_grad=need)
        ctx.module, ctx.record = module, record
        return p1.clone(), p2.clone()

    @staticmethod
    def backward(ctx, dp1, dp2):
        module = ctx.module
        c = lambda d: None if d is None else d.contiguous()
        with torch.no_grad():
            dx, pg = module._engine().backward(ctx.record, c(dp1), c(dp2))
        return (None, dx) + tuple(pg.get(n) for n in module._param_names)


class LinearModelPG(nn.Module):
    def __init__(self, linear_size=1024, num_stage=2, p_dropout=0.5, input_size=15 * 3,
                 output_size=15 * 3, bias=True, bn=True, leaky=False, precision="tf32x3"):
        super().__init__()
        if leaky or not bn or not bias or num_stage != 2:
            raise NotImplementedError("refiner is built for num_stage=2, bias=True, bn=True, leaky=False")
        if linear_size % 4:
            raise ValueError("linear_size must be a multiple of 4")
        self.linear_size, self.bn, self.leaky = linear_size, bn, leaky
        self.p_dropout, self.num_stage = p_dropout, num_stage
        self.input_size, self.output_size = input_size, output_size
        self.linear_stages = nn.ModuleList([LinearPG(linear_size, p_dropout, bias=bias, bn=bn, leaky=leaky)
                                            for _ in range(num_stage)])
        self.w1 = nn.Linear(input_size, linear_size, bias=bias)
        self.w2 = nn.Linear(linear_size, output_size, bias=bias)
        self.w3 = nn.Linear(output_size, linear_size, bias=bias)
        self.w4 = nn.Linear(linear_size, output_size, bias=bias)
        self.batch_norm1 = nn.BatchNorm1d(linear_size)
        self.batch_norm3 = nn.BatchNorm1d(linear_size)
        self._precision = {"fp32": 0, "tf32": 1, "tf32x3": 3}[precision]
        self._param_names = [n for n, _ in self.named_parameters()]
        self._eng = None

    _backend = [None]      # test hook: an ops module emulating the C ABI on the CPU

    def _engine(self):
        if self._eng is None:
            self._eng = _mlp.MLPEngine(self._precision, ops=self._backend[0])
        return self._eng

    def _buffers_flat(self):
        return dict(self.named_buffers())

    def forward(self, x):
        params = [p for _, p in self.named_parameters()]
        return _RefinerFn.apply(self, x, *params)


def get_model(weights, **kwargs):
    model = LinearModelPG(**kwargs)
    if weights:
        model.load_state_dict(torch.load(weights)['state_dict'])
    return model
