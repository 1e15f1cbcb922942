"""Config keys the hot path reads: detectron2-0.3 defaults, defrcn/config/defaults.py:7-35
and main.py:36-44 (MODEL.ADDITION.*).  A yacs `CfgNode` built by the reference's own `get_cfg()` works
unchanged; `get_cfg()` here is the stand-alone equivalent for benches/tests."""
import copy


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

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError("Non-existent config key: %s" % key)
            node[parts[-1]] = val
        return self


def get_cfg():
    C = CfgNode
    return C(
        MODEL=C(
            DEVICE="cuda", KEYPOINT_ON=False, MASK_ON=False,
            ROI_HEADS=C(NAME="Res5ROIHeads", NUM_CLASSES=20, BATCH_SIZE_PER_IMAGE=512, POSITIVE_FRACTION=0.25,
                        SCORE_THRESH_TEST=0.05, NMS_THRESH_TEST=0.5, IN_FEATURES=["res4"], PROPOSAL_APPEND_GT=True,
                        IOU_THRESHOLDS=[0.5], IOU_LABELS=[0, 1], OUTPUT_LAYER="FastRCNNOutputLayers",
                        CLS_DROPOUT=False, DROPOUT_RATIO=0.8, ENABLE_DECOUPLE=False, BACKWARD_SCALE=1.0,
                        FREEZE_FEAT=False),
            ROI_BOX_HEAD=C(NAME="", POOLER_RESOLUTION=7, POOLER_TYPE="ROIAlignV2", POOLER_SAMPLING_RATIO=0,
                           CLS_AGNOSTIC_BBOX_REG=False, SMOOTH_L1_BETA=0.0, BBOX_REG_WEIGHTS=(10.0, 10.0, 5.0, 5.0)),
            RESNETS=C(DEPTH=101, NUM_GROUPS=1, WIDTH_PER_GROUP=64, RES2_OUT_CHANNELS=256, STRIDE_IN_1X1=True,
                      NORM="FrozenBN", DEFORM_ON_PER_STAGE=[False, False, False, False]),
            RPN=C(ENABLE_DECOUPLE=False, BACKWARD_SCALE=1.0),
            ADDITION=C(NAME=None, INFERENCE_WITH_GT=False, TEACHER_TRAINING=False, STUDENT_TRAINING=False,
                       DISTIL_MODE=False, FREEZEATTENTION=False),
            # b200roi extensions (absent keys fall back to these defaults when a reference cfg is passed)
            B200=C(CHANNELS_LAST=True, RES5_DTYPE="bfloat16", RES5_IMPL="tcgen05", STATIC_SAMPLING=False, EMBED_DIR="datasets", SKIP_DEAD_BINS=True, FUSED_TRAINING=True, COSINE_LOGITS=False,
                   COSINE_TAU=20.0),
        ),
        TEST=C(DETECTIONS_PER_IMAGE=100, PCB_ENABLE=False, PCB_MODELTYPE="resnet", PCB_MODELPATH="", PCB_ALPHA=0.50,
               PCB_UPPER=1.0, PCB_LOWER=0.05),
        DATASETS=C(TRAIN=("voc_2007_trainval_all1_1shot_seed0",), TEST=("voc_2007_test_all1",)),
    )


def b200_opt(cfg, key, default):
    """Read MODEL.B200.<key> when present (reference cfgs do not carry the extension block)."""
    try:
        return cfg.MODEL.B200[key]
    except (AttributeError, KeyError, TypeError):
        return default
