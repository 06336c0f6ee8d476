"""Synthetic-occlusion augmentation -- host-side mirror of the reference
lib/utils/augmentation.py (`load_occluders` :8-58, `occlude_with_objects` :61-78, `paste_over`
:81-114, `resize_by_factor` :117-123, `list_filepaths` :126-129).

The blend itself (paste_over: float32 alpha*src + (1-alpha)*dst truncated to uint8, occluders
applied in order) runs inside the patch kernel epb_patch_sample_occ, fused with the crop, the
colour scale and the normalisation.  What stays on the host is the augmentation's parameter
side: the random draws (same np.random / random calls in the same order as the reference) and
the cv2.resize of each chosen occluder (a <= 256x256 RGBA image) -- `draw_occluders` returns
them as the list the kernel consumes.  `occlude_with_objects` keeps the reference's
numpy-in / numpy-out form on top of the same kernel path."""
import os.path
import random
import xml.etree.ElementTree

import numpy as np

MAX_OCCLUDERS = 7          # count = np.random.randint(1, 8)


def load_occluders(pascal_voc_root_path):
    """reference :8-58 (dataset side: needs the Pascal VOC tree, PIL and OpenCV)."""
    import cv2
    import PIL.Image
    occluders = []
    structuring_element = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (8, 8))
    for annotation_path in list_filepaths(os.path.join(pascal_voc_root_path, 'Annotations')):
        xml_root = xml.etree.ElementTree.parse(annotation_path).getroot()
        if xml_root.find('segmented').text == '0':
            continue
        boxes = []
        for i_obj, obj in enumerate(xml_root.findall('object')):
            is_person = (obj.find('name').text == 'person')
            is_difficult = (obj.find('difficult').text != '0')
            is_truncated = (obj.find('truncated').text != '0')
            if not is_person and not is_difficult and not is_truncated:
                bndbox = obj.find('bndbox')
                boxes.append((i_obj, [int(bndbox.find(s).text) for s in ['xmin', 'ymin', 'xmax', 'ymax']]))
        if not boxes:
            continue
        im_filename = xml_root.find('filename').text
        im = np.asarray(PIL.Image.open(os.path.join(pascal_voc_root_path, 'JPEGImages', im_filename)))
        labels = np.asarray(PIL.Image.open(os.path.join(pascal_voc_root_path, 'SegmentationObject',
                                                        im_filename.replace('jpg', 'png'))))
        for i_obj, (xmin, ymin, xmax, ymax) in boxes:
            object_mask = (labels[ymin:ymax, xmin:xmax] == i_obj + 1).astype(np.uint8) * 255
            object_image = im[ymin:ymax, xmin:xmax]
            if cv2.countNonZero(object_mask) < 500:
                continue
            eroded = cv2.erode(object_mask, structuring_element)
            object_mask[eroded < object_mask] = 192
            object_with_mask = np.concatenate([object_image, object_mask[..., np.newaxis]], axis=-1)
            occluders.append(resize_by_factor(object_with_mask, 0.5))
    return occluders


def resize_by_factor(im, factor):
    """reference :117-123 (bilinear for up-, area interpolation for down-scaling)."""
    import cv2
    new_size = tuple(np.round(np.array([im.shape[1], im.shape[0]]) * factor).astype(int))
    interp = cv2.INTER_LINEAR if factor > 1.0 else cv2.INTER_AREA
    return cv2.resize(im, new_size, fx=factor, fy=factor, interpolation=interp)


def list_filepaths(dirpath):
    names = os.listdir(dirpath)
    return sorted(filter(os.path.isfile, [os.path.join(dirpath, name) for name in names]))


def draw_occluders(width, height, occluders):
    """The parameter half of occlude_with_objects (:61-78) for a width x height image: the same
    draws in the same order -> [(rgba uint8 [h, w, 4], (cx, cy) int), ...] (1..7 entries)."""
    width_height = np.asarray([width, height])
    im_scale_factor = min(width_height) / 256
    count = np.random.randint(1, 8)
    out = []
    for _ in range(count):
        occluder = random.choice(occluders)
        random_scale_factor = np.random.uniform(0.2, 1.0)
        scale_factor = random_scale_factor * im_scale_factor
        occluder = resize_by_factor(occluder, scale_factor)
        center = np.random.uniform([0, 0], width_height)
        c = np.round(center).astype(np.int32)                  # paste_over :98
        out.append((np.ascontiguousarray(occluder[..., :4], dtype=np.uint8), (int(c[0]), int(c[1]))))
    return out


def pack_occluders(per_sample, dev):
    """[[(rgba, (cx, cy)), ...] per sample] -> (occ_base uint8, occ_desc int64 [B,7,5],
    occ_count int32 [B]) device tensors for epb_patch_sample_occ."""
    import torch
    B = len(per_sample)
    desc = np.zeros((B, MAX_OCCLUDERS, 5), dtype=np.int64)
    count = np.zeros(B, dtype=np.int32)
    chunks, pos = [], 0
    for b, lst in enumerate(per_sample):
        if len(lst) > MAX_OCCLUDERS:
            raise ValueError("at most %d occluders per sample" % MAX_OCCLUDERS)
        count[b] = len(lst)
        for k, (rgba, (cx, cy)) in enumerate(lst):
            if rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4:
                raise ValueError("occluders must be uint8 [h, w, 4] (RGB + alpha)")
            desc[b, k] = (pos, rgba.shape[1], rgba.shape[0], cx, cy)
            chunks.append(rgba.reshape(-1))
            pos += (rgba.size + 15) // 16 * 16
    base = np.zeros(max(pos, 16), dtype=np.uint8)
    p = 0
    for c in chunks:
        base[p:p + c.size] = c
        p += (c.size + 15) // 16 * 16
    return (torch.from_numpy(base).to(dev), torch.from_numpy(desc).to(dev), torch.from_numpy(count).to(dev))


def paste_over(im_src, im_dst, center):
    """reference :81-114: alpha-blend the RGBA `im_src` onto the uint8 RGB `im_dst` IN PLACE,
    centred at `center` (through the patch kernel with an identity crop)."""
    out = _blend(im_dst, [(np.ascontiguousarray(im_src, dtype=np.uint8),
                           tuple(int(v) for v in np.round(center).astype(np.int32)))])
    im_dst[...] = out


def occlude_with_objects(im, occluders):
    """reference :61-78: numpy uint8 [H, W, 3] in, occluded copy out."""
    return _blend(im, draw_occluders(im.shape[1], im.shape[0], occluders))


def _blend(im, lst):
    """uint8 RGB image + occluder list -> blended uint8 image via epb_patch_sample_occ (identity
    crop: box = whole image, patch = image size; BGR<->RGB swap undone on both sides)."""
    from . import img_utils as iu
    import torch
    H, W = im.shape[0], im.shape[1]
    bgr = np.ascontiguousarray(im[:, :, ::-1])
    out, _, _ = iu.generate_patch_batch_device([bgr], [W * 0.5], [H * 0.5], [W], [H], W, H,
                                               occluders=[lst])
    return np.ascontiguousarray(out[0].permute(1, 2, 0).cpu().numpy()).astype(np.uint8)
