"""Helpers of the "tools run unchanged" tests: a YAML for the unmodified reference tools (small geometry, synthetic
``panoptic_synth*`` datasets, one GPU), a seeded checkpoint, and the launcher command line."""
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def have_reference():
    return os.path.isfile(os.path.join(REF, "tools", "evaluate.py"))


def write_yaml(path, out_dir, ssl, with_attn=False):
    cfg = {
        "CUDNN": {"BENCHMARK": False, "DETERMINISTIC": False, "ENABLED": True},
        "BACKBONE_MODEL": "pose_resnet", "MODEL": "multi_person_posenet_ssv" if ssl else "multi_person_posenet",
        "DATA_DIR": "", "GPUS": "0", "OUTPUT_DIR": out_dir, "LOG_DIR": out_dir, "WORKERS": 0, "PRINT_FREQ": 1,
        "WITH_SSV": bool(ssl), "WITH_ATTN": bool(with_attn), "ATTN_WEIGHT": 0.1, "ATTN_NUM_LAYERS": 18,
        "USE_L1": True, "L1_WEIGHT": 0.01, "L1_ATTN": bool(with_attn),
        "DATASET": {"COLOR_RGB": True, "TRAIN_DATASET": "panoptic_synth_ssv" if ssl else "panoptic_synth",
                    "TEST_DATASET": "panoptic_synth", "DATA_FORMAT": "jpg", "DATA_AUGMENTATION": False, "FLIP": False,
                    "ROOT": "", "TEST_SUBSET": "validation", "TRAIN_SUBSET": "train", "ROOTIDX": 2, "CAMERA_NUM": 5},
        "NETWORK": {"PRETRAINED_BACKBONE": "", "PRETRAINED": "", "TARGET_TYPE": "gaussian", "INIT_ROOTNET": "",
                    "INIT_ALL": "", "TRAIN_BACKBONE": True, "TRAIN_ONLY_ROOTNET": False, "ROOTNET_TRAIN_SYNTH": False,
                    "FREEZE_ROOTNET": True, "IMAGE_SIZE_ORIG": [1920, 1080], "IMAGE_SIZE": [96, 128],
                    "HEATMAP_SIZE": [24, 32], "SIGMA": 3, "NUM_JOINTS": 15, "USE_GT": False, "ROOTNET_ROOTHM": True},
        "POSE_RESNET": {"FINAL_CONV_KERNEL": 1, "DECONV_WITH_BIAS": False, "NUM_DECONV_LAYERS": 3,
                        "NUM_DECONV_FILTERS": [256, 256, 256], "NUM_DECONV_KERNELS": [4, 4, 4], "NUM_LAYERS": 50},
        "LOSS": {"USE_TARGET_WEIGHT": True},
        "TRAIN": {"BATCH_SIZE": 1, "SHUFFLE": False, "BEGIN_EPOCH": 0, "END_EPOCH": 1, "RESUME": False,
                  "OPTIMIZER": "adam", "LR": 0.0001, "LR_FACTOR": 0.1, "LR_STEP": [5, 7], "L1_EPOCH": 0},
        "TEST": {"MODEL_FILE": "model_best.pth.tar", "BATCH_SIZE": 1},
        "DEBUG": {"DEBUG": False, "SAVE_HEATMAPS_GT": False, "SAVE_HEATMAPS_PRED": False, "SAVE_3D_POSES": False,
                  "SAVE_3D_ROOTS": False},
        "MULTI_PERSON": {"SPACE_SIZE": [8000.0, 8000.0, 2000.0], "SPACE_CENTER": [0.0, -500.0, 800.0],
                         "INITIAL_CUBE_SIZE": [16, 16, 8], "MAX_PEOPLE_NUM": 2, "THRESHOLD": -1000000.0},
        "PICT_STRUCT": {"GRID_SIZE": [2000.0, 2000.0, 2000.0], "CUBE_SIZE": [16, 16, 16]},
    }
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return cfg


def write_checkpoint(path, yaml_cfg):
    """A seeded "trained-like" state dict of the model the YAML describes (strict=True loadable by the tool)."""
    sys.path.insert(0, ROOT)
    from selfpose3d_b200 import synthetic
    from selfpose3d_b200.config import default_config
    from selfpose3d_b200.models import multi_person_posenet, multi_person_posenet_ssv
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = yaml_cfg["NETWORK"]["IMAGE_SIZE"], yaml_cfg["NETWORK"]["HEATMAP_SIZE"]
    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = yaml_cfg["MULTI_PERSON"]["INITIAL_CUBE_SIZE"]
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = yaml_cfg["MULTI_PERSON"]["MAX_PEOPLE_NUM"]
    cfg.PICT_STRUCT.CUBE_SIZE = yaml_cfg["PICT_STRUCT"]["CUBE_SIZE"]
    cfg.WITH_ATTN = yaml_cfg["WITH_ATTN"]
    mod = multi_person_posenet_ssv if yaml_cfg["WITH_SSV"] else multi_person_posenet
    cfg.MODEL = yaml_cfg["MODEL"]
    model = mod.get_multi_person_pose_net(cfg, is_train=False)
    torch.save(synthetic.trained_like_state_dict(model, seed=11), path)


def command(tool, *tool_args, emulate=False):
    runner = os.path.join(ROOT, "tests", "run_tool_emulated.py") if emulate else os.path.join(ROOT, "integration", "run_tool.py")
    return [sys.executable, runner, "--sp3d-shims", "--sp3d-synthetic-data", os.path.join(REF, "tools", tool)] + list(tool_args)
