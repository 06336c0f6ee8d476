"""Model cfg for synthetic runs (bench.py, tools/): a namespace carrying exactly the fields
PoseResNet reads (reference lib/models/pose3d_resnet.py:95-97,118,125-126,296,302-303),
i.e. the MODEL section of experiments/h36m/train-ss.yaml with the fields overridden."""
import types


def make_cfg(num_layers=50, num_joints=17, volume=True, depth_res=64, image_size=(256, 256),
             deconv_with_bias=False, final_kernel=1):
    S = types.SimpleNamespace
    extra = S(NUM_LAYERS=num_layers, DECONV_WITH_BIAS=deconv_with_bias, NUM_DECONV_LAYERS=3,
              NUM_DECONV_FILTERS=[256, 256, 256], NUM_DECONV_KERNELS=[4, 4, 4],
              FINAL_CONV_KERNEL=final_kernel)
    model = S(EXTRA=extra, VOLUME=volume, NUM_JOINTS=num_joints, DEPTH_RES=depth_res,
              IMAGE_SIZE=list(image_size), INIT_WEIGHTS=False, PRETRAINED="")
    return S(MODEL=model)
