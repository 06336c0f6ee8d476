"""The drop-in claim as a test: the UNMODIFIED reference scripts/train.py
(:66-188) runs over the mirror.  A scratch tree is laid out as INTEGRATION.md section 1 says --
<tmp>/scripts -> /root/reference/scripts (read-only link), <tmp>/lib -> epipolarpose_b200/lib --
and `python scripts/train.py --cfg <yaml>` is executed for one epoch with
DATASET.DATASET: synthetic_h36m.  No GPU in the build container: the C ABI is replaced by its
CPU emulation (tests/emul_ops.py) from a sitecustomize hook that touches neither the script nor
the mirror; `.cuda()` is a no-op there.  Skipped where /root/reference does not exist (the GPU
box): tests/test_gpu_parity.py::test_reference_script_flow replays the same call sequence on
the device."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from tests.conftest import ROOT

REF = "/root/reference"

YAML = """\
GPUS: '0'
DATA_DIR: ''
OUTPUT_DIR: '{out}'
LOG_DIR: '{out}/log'
WORKERS: 0
PRINT_FREQ: 1
EXP_NAME: dropin
DATASET:
  DATASET: synthetic_h36m
  ROOT: ''
  TRAIN_SET: train
  TEST_SET: valid
  SYNTHETIC_LEN: 8
MODEL:
  NAME: pose3d_resnet
  INIT_WEIGHTS: false
  PRETRAINED: ''
  IMAGE_SIZE: [64, 64]
  NUM_JOINTS: 4
  DEPTH_RES: 16
  VOLUME: true
  EXTRA:
    NUM_LAYERS: 18
    NUM_DECONV_LAYERS: 3
    NUM_DECONV_FILTERS: [256, 256, 256]
    NUM_DECONV_KERNELS: [4, 4, 4]
    FINAL_CONV_KERNEL: 1
    DECONV_WITH_BIAS: false
LOSS:
  FN: L1JointLocationLoss
  NORM: false
TRAIN:
  BATCH_SIZE: 4
  SHUFFLE: false
  BEGIN_EPOCH: 0
  END_EPOCH: 1
  OPTIMIZER: adam
  LR: 0.001
  LR_FACTOR: 0.1
  LR_STEP: [1]
  ONLINE_TRIANGULATION: false
TEST:
  BATCH_SIZE: 4
"""

SITECUSTOMIZE = """\
# test-only: route the C ABI to its CPU emulation and make .cuda() a no-op (no GPU here)
import sys
sys.path.insert(0, {root!r})
import torch
torch.nn.Module.cuda = lambda self, *a, **k: self
torch.Tensor.cuda = lambda self, *a, **k: self
import epipolarpose_b200.ops as _o
from tests import emul_ops as _e
for _n in dir(_e):
    if not _n.startswith("_") and callable(getattr(_e, _n)):
        setattr(_o, _n, getattr(_e, _n))
sys.path.insert(0, {tree!r})
import lib.core.integral_loss, lib.core.inference, lib.utils.img_utils, lib.utils.triangulation
import lib.dataset.h36m_eval, lib.utils.utils
for _mod in (lib.core.integral_loss, lib.core.inference, lib.utils.img_utils, lib.utils.triangulation,
             lib.dataset.h36m_eval, lib.utils.utils):
    _mod._backend[0] = _e                      # the mirrors' own test hook
import lib.models.pose3d_resnet as _m
_init = _m.PoseResNet.__init__
def _patched(self, *a, **k):
    k.setdefault("ops", _e)
    _init(self, *a, **k)
_m.PoseResNet.__init__ = _patched
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")),
                    reason="the unmodified reference tree is only present in the build container")
@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-emulation run (build container)")
def test_unmodified_train_script_runs_over_the_mirror(tmp_path):
    tree = tmp_path / "tree"
    tree.mkdir()
    os.symlink(os.path.join(REF, "scripts"), tree / "scripts")                  # UNMODIFIED scripts
    os.symlink(os.path.join(ROOT, "epipolarpose_b200", "lib"), tree / "lib")    # the mirror as `lib`
    out = tmp_path / "output"
    cfg = tmp_path / "dropin.yaml"
    cfg.write_text(YAML.format(out=str(out)))
    inject = tmp_path / "inject"
    inject.mkdir()
    (inject / "sitecustomize.py").write_text(SITECUSTOMIZE.format(root=ROOT, tree=str(tree)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(inject), ROOT]), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, str(tree / "scripts" / "train.py"), "--cfg", str(cfg), "--workers", "0"],
                       cwd=str(tree), env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    log = r.stdout + r.stderr
    assert "Epoch: [0][0/" in log and "Loss" in log            # train_integral's log line (function.py:55-63)
    assert "saving final model state" in log                   # scripts/train.py:183-186
    finals = [os.path.join(d, f) for d, _, fs in os.walk(str(out)) for f in fs if f == "final_state.pth.tar"]
    assert len(finals) == 1
    sd = torch.load(finals[0], map_location="cpu")
    assert "conv1.weight" in sd and "final_layer.bias" in sd and all(torch.isfinite(v).all() for v in sd.values()
                                                                      if v.is_floating_point())
    ck = [os.path.join(d, f) for d, _, fs in os.walk(str(out)) for f in fs if f == "checkpoint.pth.tar"]
    assert ck and {"epoch", "state_dict", "perf", "optimizer"} <= set(torch.load(ck[0], map_location="cpu",
                                                                                 weights_only=False))
