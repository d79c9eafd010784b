"""The ~20 configuration keys the region pipeline reads.

The reference reads them from a global mutable ``cfg`` AttrDict
(lib/core/config.py:22); here they are an explicit object.  ``RegionConfig.from_cfg``
adapts a reference-style ``cfg`` (anything with attribute access) so the shims
can be dropped into the unmodified model builders, and ``set_cfg`` / ``get_cfg``
hold the process-wide default the reference-signature shims consult.
"""
import copy
import math
from dataclasses import dataclass, field


@dataclass
class RpnMode:
    pre_nms_topN: int
    post_nms_topN: int
    nms_thresh: float = 0.7
    min_size: float = 0.0


@dataclass
class RegionConfig:
    # {TRAIN,TEST}.RPN_* (lib/core/config.py:132-149, 202-215; yaml e2e_mask_rcnn_R-50-FPN_1x.yaml:37-43)
    train: RpnMode = field(default_factory=lambda: RpnMode(2000, 2000))
    test: RpnMode = field(default_factory=lambda: RpnMode(1000, 1000))
    # FPN.* (config.py:679-722)
    rpn_min_level: int = 2
    rpn_max_level: int = 6
    roi_min_level: int = 2
    roi_max_level: int = 5
    roi_canonical_scale: float = 224.0
    roi_canonical_level: int = 4
    rpn_collect_scale: float = 1.0
    rpn_anchor_start_size: int = 32
    rpn_aspect_ratios: tuple = (0.5, 1, 2)
    # BBOX_XFORM_CLIP (config.py:1009)
    bbox_xform_clip: float = math.log(1000.0 / 16.0)
    # MODEL / MRCNN (config.py:740-775)
    num_classes: int = 81
    mrcnn_resolution: int = 28
    mrcnn_thresh_binarize: float = 0.5
    mrcnn_cls_specific_mask: bool = True
    # TEST.* / MODEL.* read by box_results_with_nms_and_limit and im_detect_bbox (config.py:190,219,224,359,371,417)
    test_score_thresh: float = 0.05
    test_nms: float = 0.3
    test_detections_per_im: int = 100
    test_num_det_per_class: int = 0
    test_soft_nms: bool = False
    test_bbox_vote: bool = False
    bbox_reg_weights: tuple = (10.0, 10.0, 5.0, 5.0)
    # TRAIN.* / MODEL.* read by the label assignment (config.py:99-127; roi_data/fast_rcnn.py:132-213)
    train_batch_size_per_im: int = 512
    train_fg_fraction: float = 0.25
    train_fg_thresh: float = 0.5
    train_bg_thresh_hi: float = 0.5
    train_bg_thresh_lo: float = 0.0
    cls_agnostic_bbox_reg: bool = False
    mask_on: bool = False
    # TRAIN.RPN_* / RPN.* / FPN.* read by the RPN label assignment (config.py:48,118-145,666-709; roi_data/rpn.py)
    fpn_on: bool = True
    multilevel_rpn: bool = True
    fpn_coarsest_stride: int = 32
    train_max_size: int = 1333
    train_rpn_positive_overlap: float = 0.7
    train_rpn_negative_overlap: float = 0.3
    train_rpn_fg_fraction: float = 0.5
    train_rpn_batch_size_per_im: int = 256
    train_rpn_straddle_thresh: float = 0
    rpn_stride: int = 16
    rpn_sizes: tuple = (64, 128, 256, 512)
    rpn_single_aspect_ratios: tuple = (0.5, 1, 2)
    identity_training: bool = False
    # lib_vos extras of box_results_with_nms_and_limit / nms_with_mask_iou (config.py:948-953)
    test_num_det_per_class_pre: int = 0
    test_num_det_per_class_post: int = 0
    test_nms_cross_class: float = 0.0
    test_nms_with_mask_iou: float = 0.0
    test_nms_small_box_iou: float = 0.0
    test_nms_small_box_score_threshold: float = 0.0

    def mode(self, training):
        return self.train if training else self.test

    def collect_post_topN(self, training):
        # int(POST_NMS_TOP_N * RPN_COLLECT_SCALE + 0.5), collect_and_distribute...py:93
        return int(self.mode(training).post_nms_topN * self.rpn_collect_scale + 0.5)

    @classmethod
    def from_cfg(cls, cfg):
        def mode(m):
            return RpnMode(int(m.RPN_PRE_NMS_TOP_N), int(m.RPN_POST_NMS_TOP_N), float(m.RPN_NMS_THRESH),
                           float(m.RPN_MIN_SIZE))
        return cls(
            train=mode(cfg.TRAIN), test=mode(cfg.TEST),
            rpn_min_level=int(cfg.FPN.RPN_MIN_LEVEL), rpn_max_level=int(cfg.FPN.RPN_MAX_LEVEL),
            roi_min_level=int(cfg.FPN.ROI_MIN_LEVEL), roi_max_level=int(cfg.FPN.ROI_MAX_LEVEL),
            roi_canonical_scale=float(cfg.FPN.ROI_CANONICAL_SCALE),
            roi_canonical_level=int(cfg.FPN.ROI_CANONICAL_LEVEL),
            rpn_collect_scale=float(cfg.FPN.RPN_COLLECT_SCALE),
            rpn_anchor_start_size=int(cfg.FPN.RPN_ANCHOR_START_SIZE),
            rpn_aspect_ratios=tuple(cfg.FPN.RPN_ASPECT_RATIOS),
            bbox_xform_clip=float(cfg.BBOX_XFORM_CLIP),
            num_classes=int(cfg.MODEL.NUM_CLASSES),
            mrcnn_resolution=int(cfg.MRCNN.RESOLUTION),
            mrcnn_thresh_binarize=float(cfg.MRCNN.THRESH_BINARIZE),
            mrcnn_cls_specific_mask=bool(cfg.MRCNN.CLS_SPECIFIC_MASK),
            test_score_thresh=float(cfg.TEST.SCORE_THRESH), test_nms=float(cfg.TEST.NMS),
            test_detections_per_im=int(cfg.TEST.DETECTIONS_PER_IM),
            # lib/core/test.py:785 reads TEST.NUM_DET_PER_CLASS, which lib/core/config.py never defines
            # (it has NUM_DET_PER_CLASS_PRE / _POST, config.py:948-949): absent means 0 here
            test_num_det_per_class=int(getattr(cfg.TEST, "NUM_DET_PER_CLASS", 0) or 0),
            test_soft_nms=bool(cfg.TEST.SOFT_NMS.ENABLED), test_bbox_vote=bool(cfg.TEST.BBOX_VOTE.ENABLED),
            bbox_reg_weights=tuple(float(x) for x in cfg.MODEL.BBOX_REG_WEIGHTS),
            # label-assignment knobs: a partial cfg (inference only) keeps the reference's defaults (config.py:99-127)
            train_batch_size_per_im=int(getattr(cfg.TRAIN, "BATCH_SIZE_PER_IM", 512)),
            train_fg_fraction=float(getattr(cfg.TRAIN, "FG_FRACTION", 0.25)),
            train_fg_thresh=float(getattr(cfg.TRAIN, "FG_THRESH", 0.5)),
            train_bg_thresh_hi=float(getattr(cfg.TRAIN, "BG_THRESH_HI", 0.5)),
            train_bg_thresh_lo=float(getattr(cfg.TRAIN, "BG_THRESH_LO", 0.0)),
            cls_agnostic_bbox_reg=bool(getattr(cfg.MODEL, "CLS_AGNOSTIC_BBOX_REG", False)),
            mask_on=bool(getattr(cfg.MODEL, "MASK_ON", False)),
            fpn_on=bool(getattr(cfg.FPN, "FPN_ON", True)), multilevel_rpn=bool(getattr(cfg.FPN, "MULTILEVEL_RPN", True)),
            fpn_coarsest_stride=int(getattr(cfg.FPN, "COARSEST_STRIDE", 32)),
            train_max_size=int(getattr(cfg.TRAIN, "MAX_SIZE", 1333)),
            train_rpn_positive_overlap=float(getattr(cfg.TRAIN, "RPN_POSITIVE_OVERLAP", 0.7)),
            train_rpn_negative_overlap=float(getattr(cfg.TRAIN, "RPN_NEGATIVE_OVERLAP", 0.3)),
            train_rpn_fg_fraction=float(getattr(cfg.TRAIN, "RPN_FG_FRACTION", 0.5)),
            train_rpn_batch_size_per_im=int(getattr(cfg.TRAIN, "RPN_BATCH_SIZE_PER_IM", 256)),
            train_rpn_straddle_thresh=float(getattr(cfg.TRAIN, "RPN_STRADDLE_THRESH", 0)),
            rpn_stride=int(getattr(getattr(cfg, "RPN", None), "STRIDE", 16)),
            rpn_sizes=tuple(getattr(getattr(cfg, "RPN", None), "SIZES", (64, 128, 256, 512))),
            rpn_single_aspect_ratios=tuple(getattr(getattr(cfg, "RPN", None), "ASPECT_RATIOS", (0.5, 1, 2))),
            identity_training=bool(getattr(cfg.MODEL, "IDENTITY_TRAINING", False)),
            test_num_det_per_class_pre=int(getattr(cfg.TEST, "NUM_DET_PER_CLASS_PRE", 0) or 0),
            test_num_det_per_class_post=int(getattr(cfg.TEST, "NUM_DET_PER_CLASS_POST", 0) or 0),
            test_nms_cross_class=float(getattr(cfg.TEST, "NMS_CROSS_CLASS", 0.0) or 0.0),
            test_nms_with_mask_iou=float(getattr(cfg.TEST, "NMS_WITH_MASK_IOU", 0.0) or 0.0),
            test_nms_small_box_iou=float(getattr(cfg.TEST, "NMS_SMALL_BOX_IOU", 0.0) or 0.0),
            test_nms_small_box_score_threshold=float(getattr(cfg.TEST, "NMS_SMALL_BOX_SCORE_THRESHOLD", 0.0) or 0.0),
        )


_default = RegionConfig()


def get_cfg():
    return _default


def set_cfg(cfg):
    """Install the process-wide default (a RegionConfig, or a reference-style cfg)."""
    global _default
    _default = cfg if isinstance(cfg, RegionConfig) else RegionConfig.from_cfg(cfg)
    return _default


def clone_cfg():
    return copy.deepcopy(_default)
