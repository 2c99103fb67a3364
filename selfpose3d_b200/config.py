"""Minimal attribute-dict configuration for standalone use (bench, smoke, tests).

The reference's config system (``lib/core/config.py``) is out of scope and is reused
unchanged when the reference's tools drive this backend; the model classes only read
attributes, so any object with the same attribute names works.  ``default_config()``
carries the keys the hot path reads, at the values of the shipped SSL config
``configs/panoptic_ssl/resnet50/cam5_posenet.yaml`` but with the BASELINE geometry
(network input 288x384 ``[w,h]``, heat-maps 72x96).
"""
from __future__ import annotations


class AttrDict(dict):
    """dict with attribute access (what the model classes need from EasyDict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def default_config():
    c = AttrDict()
    c.BACKBONE_MODEL = "pose_resnet"
    c.MODEL = "multi_person_posenet_ssv"
    c.WITH_ATTN = False
    c.ATTN_WEIGHT = 0.1
    c.ATTN_NUM_LAYERS = 18
    c.USE_L1 = False
    c.L1_WEIGHT = 0.01
    c.L1_ATTN = False
    c.EVAL_ROOTNET_ONLY = False
    c.COCO_TO_PANOPTIC_MAPPING = [5, 0, 11, 5, 7, 9, 11, 13, 15, 6, 8, 10, 12, 14, 16]

    c.NETWORK = AttrDict(
        PRETRAINED="", IMAGE_SIZE=[288, 384], HEATMAP_SIZE=[72, 96], NUM_JOINTS=15, SIGMA=3, BETA=100.0,
        USE_GT=False, TRAIN_ONLY_2D=False, TRAIN_ONLY_ROOTNET=False, ROOTNET_ROOTHM=True,
        ROOTNET_TRAIN_SYNTH=False, FREEZE_ROOTNET=True, SINGLE_AUG_TRAINING_POSENET=False,
        ROOT_CONSISTENCY_LOSS=True, WEIGHT_ROOT_SYN=100.0, WEIGHT_ROOT_REG=1.0, INIT_TRAIN_EPOCHS_ROOTNET=0,
        ROOTNET_SYN_RANGE=[[2500.0, -2000.0], [1500.0, -1500.0], [250.0, -300.0]])
    c.POSE_RESNET = AttrDict(
        NUM_LAYERS=50, DECONV_WITH_BIAS=False, NUM_DECONV_LAYERS=3, NUM_DECONV_FILTERS=[256, 256, 256],
        NUM_DECONV_KERNELS=[4, 4, 4], FINAL_CONV_KERNEL=1)
    c.DATASET = AttrDict(ROOTIDX=2, ROOTIDX_PSEUDO=2, TEST_DATASET="panoptic", CAMERA_NUM=5)
    c.TRAIN = AttrDict(BATCH_SIZE=1, L1_EPOCH=5)
    c.TEST = AttrDict(BATCH_SIZE=4)
    c.MULTI_PERSON = AttrDict(
        SPACE_SIZE=[8000.0, 8000.0, 2000.0], SPACE_CENTER=[0.0, -500.0, 800.0],
        INITIAL_CUBE_SIZE=[80, 80, 20], MAX_PEOPLE_NUM=10, THRESHOLD=0.3)
    c.PICT_STRUCT = AttrDict(GRID_SIZE=[2000.0, 2000.0, 2000.0], CUBE_SIZE=[64, 64, 64])
    return c
