"""openglottal_b200: B200-native drop-in for OpenGlottal's unet-only hot path.

Public names mirror ``openglottal`` (/root/reference/openglottal/__init__.py) for the path
this package accelerates: ``UNet``, ``extract_features_unet``, ``unet_segment_frame``,
``_kinematic_features``; plus the batched / sharded entry points.
"""
__version__ = "0.1.0"

from .unet import UNet
from .utils import (
    unet_segment_frame,
    unet_segment_frames,
    bgr_to_gray,
    load_frames_bgr,
    load_frames_bgr_parallel,
    video_info,
    dice,
    iou,
    dice_iou_batch,
    gated_area,
    letterbox_geometry,
    letterbox_crops,
    unletterbox_area,
    segment_crops,
    resize_u8_linear,
    prob_resize_mask,
    segment_frames_reference_resize,
)
from .features import (
    _kinematic_features,
    kinematic_features_device,
    segment_clip,
    masks_for_clip,
    decode_gray_clip,
    extract_features_unet,
    extract_features_unet_frames,
    extract_features_yolo_crop_unet,
)

__all__ = [
    "UNet",
    "unet_segment_frame",
    "unet_segment_frames",
    "bgr_to_gray",
    "load_frames_bgr",
    "load_frames_bgr_parallel",
    "video_info",
    "decode_gray_clip",
    "dice",
    "iou",
    "dice_iou_batch",
    "gated_area",
    "letterbox_geometry",
    "letterbox_crops",
    "unletterbox_area",
    "segment_crops",
    "resize_u8_linear",
    "prob_resize_mask",
    "segment_frames_reference_resize",
    "_kinematic_features",
    "kinematic_features_device",
    "segment_clip",
    "masks_for_clip",
    "extract_features_unet",
    "extract_features_unet_frames",
    "extract_features_yolo_crop_unet",
]
