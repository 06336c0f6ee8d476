from . import pose3d_resnet
