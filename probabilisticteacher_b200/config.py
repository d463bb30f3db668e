"""Configuration for the hot path: the keys of `pt/config.py:20-92` (add_config) on top of the
detectron2 v0.5 defaults the path reads, with YAML `_BASE_` inheritance and `KEY value` overrides
as used by `train.sh:5-12`. yacs is not required."""
import ast
import copy
import os

import yaml


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_dict(self, d):
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge_from_dict(v)
            else:
                if isinstance(v, str) and v.startswith("(") and v.endswith(")"):
                    v = ast.literal_eval(v)
                self[k] = _wrap(v)

    def merge_from_file(self, path):
        with open(path) as f:
            d = yaml.safe_load(f) or {}
        base = d.pop("_BASE_", None)
        if base:
            self.merge_from_file(os.path.join(os.path.dirname(path), base))
        self.merge_from_dict(d)

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for k, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            node[parts[-1]] = _wrap(v)

    def freeze(self):
        return self


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, CfgNode):
        n = CfgNode()
        for k, x in v.items():
            n[k] = _wrap(x)
        return n
    return v


CN = CfgNode


def get_cfg():
    """detectron2 v0.5 defaults (subset read by the hot path) + pt/config.py add_config."""
    c = _wrap({
        "VERSION": 2,
        "MODEL": {
            "META_ARCHITECTURE": "GuassianGeneralizedRCNN", "DEVICE": "cuda", "WEIGHTS": "", "MASK_ON": False,
            "PIXEL_MEAN": [103.530, 116.280, 123.675], "PIXEL_STD": [1.0, 1.0, 1.0],
            "BACKBONE": {"NAME": "build_vgg_backbone", "FREEZE_AT": 2},
            "VGG": {"DEPTH": 16, "OUT_FEATURES": ["vgg_block5"], "NORM": "None", "CONV5_OUT_CHANNELS": 512,
                    "PRETRAIN": "./vgg16_caffe.pth"},
            "ANCHOR_GENERATOR": {
                "NAME": "DefaultAnchorGenerator", "SIZES": [[128, 256, 512]], "ASPECT_RATIOS": [[0.5, 1.0, 2.0]],
                "OFFSET": 0.0,
                "ANCHOR": [[[181.0193, 90.5097], [128.0, 128.0], [90.5097, 181.0193], [362.0387, 181.0193],
                            [256.0, 256.0], [181.0193, 362.0387], [724.0773, 362.0387], [512.0, 512.0],
                            [362.0387, 724.0773]]]},
            "PROPOSAL_GENERATOR": {"NAME": "GuassianRPN", "MIN_SIZE": 0},
            "RPN": {"HEAD_NAME": "GuassianRPNHead", "IN_FEATURES": ["vgg_block5"], "BOUNDARY_THRESH": -1,
                    "IOU_THRESHOLDS": [0.3, 0.7], "IOU_LABELS": [0, -1, 1], "BATCH_SIZE_PER_IMAGE": 256,
                    "POSITIVE_FRACTION": 0.25, "BBOX_REG_LOSS_TYPE": "smooth_l1", "BBOX_REG_LOSS_WEIGHT": 1.0,
                    "BBOX_REG_WEIGHTS": (1.0, 1.0, 1.0, 1.0), "SMOOTH_L1_BETA": 0.0, "LOSS_WEIGHT": 1.0,
                    "PRE_NMS_TOPK_TRAIN": 12000, "PRE_NMS_TOPK_TEST": 6000, "POST_NMS_TOPK_TRAIN": 2000,
                    "POST_NMS_TOPK_TEST": 1000, "NMS_THRESH": 0.7},
            "ROI_HEADS": {"NAME": "GuassianROIHead", "NUM_CLASSES": 8, "IN_FEATURES": ["vgg_block5"],
                          "IOU_THRESHOLDS": [0.5], "IOU_LABELS": [0, 1], "BATCH_SIZE_PER_IMAGE": 512,
                          "POSITIVE_FRACTION": 0.25, "SCORE_THRESH_TEST": 0.05, "NMS_THRESH_TEST": 0.5,
                          "PROPOSAL_APPEND_GT": True},
            "ROI_BOX_HEAD": {"NAME": "FastRCNNConvFCHead", "BBOX_REG_LOSS_TYPE": "smooth_l1",
                             "BBOX_REG_LOSS_WEIGHT": 1.0, "BBOX_REG_WEIGHTS": (10.0, 10.0, 5.0, 5.0),
                             "SMOOTH_L1_BETA": 0.0, "POOLER_RESOLUTION": 7, "POOLER_SAMPLING_RATIO": 0,
                             "POOLER_TYPE": "ROIAlignV2", "NUM_FC": 2, "FC_DIM": 1024, "NUM_CONV": 0,
                             "CLS_AGNOSTIC_BBOX_REG": False, "TRAIN_ON_PRED_BOXES": False},
        },
        "INPUT": {"MIN_SIZE_TRAIN": (600,), "MAX_SIZE_TRAIN": 1333, "MIN_SIZE_TEST": 600, "MAX_SIZE_TEST": 1333,
                  "FORMAT": "BGR", "RANDOM_FLIP": "horizontal"},
        "DATASETS": {"TRAIN": (), "TEST": ()},
        "DATALOADER": {"NUM_WORKERS": 4},
        "SOLVER": {"LR_SCHEDULER_NAME": "WarmupMultiStepLR", "MAX_ITER": 40000, "BASE_LR": 0.001, "MOMENTUM": 0.9,
                   "WEIGHT_DECAY": 0.0001, "GAMMA": 0.1, "STEPS": (30000,), "WARMUP_FACTOR": 0.001,
                   "WARMUP_ITERS": 1000, "WARMUP_METHOD": "linear", "CHECKPOINT_PERIOD": 5000, "IMS_PER_BATCH": 16,
                   "AMP": {"ENABLED": False}},
        "TEST": {"DETECTIONS_PER_IMAGE": 100, "EVAL_PERIOD": 0},
        "OUTPUT_DIR": "./output",
        "SEED": -1,
    })
    add_config(c)
    return c


def add_config(cfg):
    """`pt/config.py:20-92`."""
    _C = cfg
    _C.SOLVER.IMG_PER_BATCH_LABEL = 16
    _C.SOLVER.IMG_PER_BATCH_UNLABEL = 16
    _C.SOLVER.FACTOR_LIST = (1,)
    _C.SOLVER.REFERENCE_WORLD_SIZE = 1
    _C.SOLVER.REFERENCE_BATCH_SIZE = 0
    _C.DATASETS.TRAIN_LABEL = ("coco_2017_train",)
    _C.DATASETS.TRAIN_UNLABEL = ("coco_2017_train",)
    _C.DATASETS.CROSS_DATASET = True
    _C.TEST.EVALUATOR = "COCOeval"
    _C.UNSUPNET = CN()
    _C.UNSUPNET.Trainer = "pt"
    _C.UNSUPNET.PSEUDO_BBOX_SAMPLE = "all"
    _C.UNSUPNET.TEACHER_UPDATE_ITER = 1
    _C.UNSUPNET.BURN_UP_STEP = 4000
    _C.UNSUPNET.EMA_KEEP_RATE = 0.0
    _C.UNSUPNET.LOSS_WEIGHT_TYPE = "standard"
    _C.UNSUPNET.SOURCE_LOSS_WEIGHT = 1.0
    _C.UNSUPNET.TARGET_UNSUP_LOSS_WEIGHT = 1.0
    _C.UNSUPNET.GUASSIAN = True
    _C.UNSUPNET.TAU = [0.5, 0.5]
    _C.UNSUPNET.EFL = True
    _C.UNSUPNET.EFL_LAMBDA = [0.5, 0.5]
    _C.UNSUPNET.MODEL_TYPE = "GUASSIAN"
    return cfg


def c2f_config():
    """configs/pt/final_c2f.yaml on top of configs/Guassian-RCNN-VGG.yaml with the train.sh overrides
    (DifferentiableAnchorGenerator, EFL, lambda 0.5/0.5, tau 0.5/0.5), restated as values."""
    c = get_cfg()
    c.MODEL.RPN.PRE_NMS_TOPK_TEST = 6000
    c.MODEL.RPN.POST_NMS_TOPK_TEST = 1000
    c.MODEL.ROI_HEADS.NUM_CLASSES = 8
    c.SOLVER.BASE_LR = 0.016
    c.SOLVER.STEPS = (30000,)
    c.SOLVER.MAX_ITER = 30000
    c.SOLVER.WARMUP_ITERS = 400
    c.SOLVER.CHECKPOINT_PERIOD = 4000
    c.SOLVER.REFERENCE_BATCH_SIZE = 16
    c.UNSUPNET.EMA_KEEP_RATE = 0.9996
    c.UNSUPNET.BURN_UP_STEP = 4000
    c.UNSUPNET.TAU = [0.5, 0.5]
    c.MODEL.ANCHOR_GENERATOR.NAME = "DifferentiableAnchorGenerator"
    c.TEST.EVAL_PERIOD = 400
    return c


def k2c_config():
    """configs/pt/final_k2c.yaml (KITTI -> Cityscapes, car only): final_c2f with NUM_CLASSES = 1; the
    train.sh overrides are kept."""
    c = c2f_config()
    c.MODEL.ROI_HEADS.NUM_CLASSES = 1
    return c
