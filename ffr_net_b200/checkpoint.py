"""Checkpoint container of the reference: gzip-compressed torch.save files `<name>.pth.gzip`
holding {'RecNet': state_dict, 'optimizer': state_dict, 'epoch': int, 'iter': int} (models/trainer.py:201-224,
utils/utils.py:110-123), plus the plain `se50.pth` backbone file ir_se_50_512 loads (model_ir_se50.py:143-154).
Files written here load in the reference and vice versa; the state_dict key layouts are those of the reference
(tests/test_host_cpu.py::test_state_dict_layout_matches_reference). Host-side only."""
import gzip
import os

import torch


def save(obj, save_path):
    """utils.save, utils/utils.py:110-115."""
    with gzip.GzipFile(save_path, "wb") as f:
        torch.save(obj, f)


def load(read_path, map_location=None):
    """utils.load, utils/utils.py:117-123 ('.gzip' -> gzip stream, anything else -> plain torch.load)."""
    if read_path.endswith(".gzip"):
        with gzip.open(read_path, "rb") as f:
            return torch.load(f, map_location=map_location, weights_only=False)
    return torch.load(read_path, map_location=map_location, weights_only=False)


def resolve(ckpt_dir, file_name):
    """Trainer.load_model's file resolution, models/trainer.py:202-210: 'latest' = last `*pth.gzip` of the sorted
    directory listing; a name containing '/' is taken as a path."""
    if file_name == "latest":
        weights = sorted(x for x in os.listdir(ckpt_dir) if x.endswith("pth.gzip"))
        if not weights:
            raise FileNotFoundError("no *.pth.gzip checkpoint in %s" % ckpt_dir)
        file_name = weights[-1]
    else:
        file_name = file_name + ".pth.gzip"
    return file_name if "/" in file_name else os.path.join(ckpt_dir, file_name)


def save_model(recnet, optimizer, ckpt_dir, file_name, extra_info=None):
    """Trainer.save_model, models/trainer.py:216-224."""
    weight_dict = {"RecNet": recnet.state_dict(), "optimizer": optimizer.state_dict()}
    if extra_info is not None:
        weight_dict.update(extra_info)
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, file_name + ".pth.gzip")
    save(weight_dict, path)
    return path


def load_model(recnet, ckpt_dir, file_name, map_location=None):
    """Trainer.load_model, models/trainer.py:201-214: non-strict load of weights['RecNet'] (the optimizer state is NOT
    restored, as in the reference); returns the start point {'epoch', 'iter'}."""
    weights = load(resolve(ckpt_dir, file_name), map_location=map_location)
    recnet.load_state_dict(weights["RecNet"], strict=False)
    return {"epoch": weights["epoch"], "iter": weights["iter"]}
