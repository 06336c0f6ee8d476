"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference (mkocabas/EpipolarPose) read-only from
/root/reference under the minimal shim set SURVEY.md section 7 lists, so that
its own functions can be executed on CPU to (a) pin the numpy/C restatements
in oracle/restate.py and (b) generate the golden vectors committed under
tests/golden/.  /root/reference exists only in the build container, never on
the GPU box: nothing under tests -m gpu, smoke() or bench.py may import this
module at run time.

Shims (each one only makes a 2019 code base importable on py3.12/numpy2/torch2.11;
none changes arithmetic):
  np.int/np.float aliases          lib/utils/prep_h36m.py:68-69, lib/dataset/h36m.py:23
  stub matplotlib/mpl_toolkits/h5py lib/utils/img_utils.py:4-5, lib/utils/cameras.py:1
  easydict.EasyDict stand-in        lib/core/config.py:5
  torch.cuda.comm.broadcast -> [t]  lib/core/integral_loss.py:61-63
  yaml.load -> safe_load            lib/core/config.py:173
"""
import os
import sys
import types

REF_ROOT = os.environ.get("EPB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib"))


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


_installed = False


def install():
    """Install shims and put the reference root on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "h5py"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name == "mpl_toolkits.mplot3d":
                    m.Axes3D = object
                    m.axes3d = object
                m.__path__ = []
                sys.modules[name] = m
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    import torch
    import torch.cuda
    try:
        import torch.cuda.comm  # noqa: F401
    except Exception:
        torch.cuda.comm = types.ModuleType("torch.cuda.comm")
    torch.cuda.comm.broadcast = lambda t, devices=None: [t if devices is None or devices[0] is None else t.to("cuda:%d" % devices[0])]
    import yaml
    _orig = yaml.load
    yaml.load = lambda f, Loader=None: _orig(f, Loader=Loader or yaml.SafeLoader)
    # reference `lib` must win over any other `lib` on the path for oracle use
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    _installed = True


def ref():
    """Return a namespace with the reference modules on the hot path."""
    install()
    import importlib
    ns = types.SimpleNamespace()
    ns.pose3d_resnet = importlib.import_module("lib.models.pose3d_resnet")
    ns.integral_loss = importlib.import_module("lib.core.integral_loss")
    ns.triangulation = importlib.import_module("lib.utils.triangulation")
    ns.img_utils = importlib.import_module("lib.utils.img_utils")
    ns.inference = importlib.import_module("lib.core.inference")
    ns.cameras = importlib.import_module("lib.utils.cameras")
    ns.prep_h36m = importlib.import_module("lib.utils.prep_h36m")
    ns.config = importlib.import_module("lib.core.config")
    ns.function = importlib.import_module("lib.core.function")
    ns.utils = importlib.import_module("lib.utils.utils")
    return ns


def make_cfg(num_layers=50, num_joints=17, volume=True, depth_res=64,
             image_size=(256, 256), deconv_with_bias=False, final_kernel=1):
    """SimpleNamespace cfg carrying exactly the fields PoseResNet reads
    (lib/models/pose3d_resnet.py:95-97,118,125-126,296,302-303)."""
    S = types.SimpleNamespace
    extra = S(NUM_LAYERS=num_layers, DECONV_WITH_BIAS=deconv_with_bias,
              NUM_DECONV_LAYERS=3, NUM_DECONV_FILTERS=[256, 256, 256],
              NUM_DECONV_KERNELS=[4, 4, 4], FINAL_CONV_KERNEL=final_kernel)
    model = S(EXTRA=extra, VOLUME=volume, NUM_JOINTS=num_joints,
              DEPTH_RES=depth_res, IMAGE_SIZE=list(image_size),
              INIT_WEIGHTS=False, PRETRAINED="")
    return S(MODEL=model)
