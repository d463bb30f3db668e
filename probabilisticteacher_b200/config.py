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


def validate_cfg(cfg):
    """Refuses configurations whose value the reference honours but this path would silently ignore (the kernels are
    specialised to the structure every reference config uses: configs/Guassian-RCNN-VGG.yaml + configs/pt/*.yaml on
    detectron2 v0.5 defaults). Raises ValueError naming the key. Values that ARE honoured (thresholds, batch sizes,
    top-k sizes, loss weights, UNSUPNET.*, NUM_CLASSES, anchor shapes, pixel statistics ...) are not restricted."""
    m = cfg.MODEL

    def need(key, ok, why):
        if not ok:
            raise ValueError(f"{key}: unsupported value on the B200 path ({why})")

    need("MODEL.MASK_ON", not m.MASK_ON, "box heads only")
    need("MODEL.VGG.DEPTH", m.VGG.DEPTH == 16, "VGG16 backbone")
    need("MODEL.VGG.OUT_FEATURES", list(m.VGG.OUT_FEATURES) == ["vgg_block5"], "single stride-16 feature map")
    need("MODEL.VGG.NORM", str(m.VGG.NORM) in ("None", ""), "no normalisation layers (pt/config.py:74)")
    need("MODEL.BACKBONE.FREEZE_AT", m.BACKBONE.FREEZE_AT >= 1, "the fused first conv is forward-only")
    n_cell = len(m.ANCHOR_GENERATOR.SIZES[0]) * len(m.ANCHOR_GENERATOR.ASPECT_RATIOS[0])
    if m.ANCHOR_GENERATOR.NAME == "DefaultAnchorGenerator":
        need("MODEL.ANCHOR_GENERATOR.SIZES x ASPECT_RATIOS", len(m.ANCHOR_GENERATOR.SIZES) == 1 and n_cell == 9,
             "9 cell anchors on one level")
    else:
        need("MODEL.ANCHOR_GENERATOR.ANCHOR", len(m.ANCHOR_GENERATOR.ANCHOR) == 1 and len(m.ANCHOR_GENERATOR.ANCHOR[0]) == 9,
             "9 learnable (w, h) pairs on one level (pt/config.py:84-92)")
    r = m.RPN
    need("MODEL.RPN.IN_FEATURES", list(r.IN_FEATURES) == ["vgg_block5"], "single level")
    need("MODEL.RPN.IOU_LABELS", list(r.IOU_LABELS) == [0, -1, 1] and len(r.IOU_THRESHOLDS) == 2, "bg / ignore / fg matcher")
    need("MODEL.RPN.BOUNDARY_THRESH", r.BOUNDARY_THRESH == -1, "no anchor boundary filter")
    need("MODEL.RPN.BBOX_REG_WEIGHTS", tuple(float(x) for x in r.BBOX_REG_WEIGHTS) == (1.0, 1.0, 1.0, 1.0), "unit weights")
    need("MODEL.RPN.BBOX_REG_LOSS_WEIGHT", float(r.BBOX_REG_LOSS_WEIGHT) == 1.0, "LOSS_WEIGHT scales both RPN losses")
    h = m.ROI_HEADS
    need("MODEL.ROI_HEADS.IN_FEATURES", list(h.IN_FEATURES) == ["vgg_block5"], "single level")
    need("MODEL.ROI_HEADS.IOU_LABELS", list(h.IOU_LABELS) == [0, 1] and len(h.IOU_THRESHOLDS) == 1, "bg / fg matcher")
    need("MODEL.ROI_HEADS.PROPOSAL_APPEND_GT", bool(h.PROPOSAL_APPEND_GT), "ground truth is always appended")
    need("MODEL.ROI_HEADS.NUM_CLASSES", 1 <= h.NUM_CLASSES <= 14, "cls_score + bbox_pred share one 128-row block: K + 1 + 8K <= 128")
    b = m.ROI_BOX_HEAD
    need("MODEL.ROI_BOX_HEAD.NAME", b.NAME == "FastRCNNConvFCHead" and b.NUM_FC == 2 and b.NUM_CONV == 0, "two FC layers")
    need("MODEL.ROI_BOX_HEAD.FC_DIM", b.FC_DIM % 256 == 0, "GEMM tile width")
    need("MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION", b.POOLER_RESOLUTION == 7, "ROIAlign kernel geometry")
    need("MODEL.ROI_BOX_HEAD.POOLER_TYPE", b.POOLER_TYPE == "ROIAlignV2" and b.POOLER_SAMPLING_RATIO == 0,
         "aligned ROIAlign with the adaptive sampling grid")
    need("MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG", not b.CLS_AGNOSTIC_BBOX_REG, "class-specific Gaussian box head")
    need("MODEL.ROI_BOX_HEAD.TRAIN_ON_PRED_BOXES", not b.TRAIN_ON_PRED_BOXES, "")
    need("MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT", float(b.BBOX_REG_LOSS_WEIGHT) == 1.0, "")
    need("UNSUPNET.MODEL_TYPE", cfg.UNSUPNET.MODEL_TYPE == "GUASSIAN", "Gaussian heads only")
    return cfg
