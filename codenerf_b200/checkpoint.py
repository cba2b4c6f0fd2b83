"""Checkpoint wire formats of the reference, so its files flow through this path unchanged.

* ``models.pth`` (reference src/trainer.py:165-174): a ``torch.save`` dict with ``model_params`` (the module's
  ``state_dict``: 28 fp32 tensors, names and shapes as in src/model.py:11-34), ``shape_code_params`` /
  ``texture_code_params`` (``nn.Embedding.state_dict()``: one ``weight`` [n_objects, latent]), ``niter``, ``nepoch``.
  The bf16 stage images this library packs from the weights are derived state and are never written.
* ``codes.pth`` (reference src/optimizer.py:137-146): ``ids``, ``num_obj``, ``optimized_shapecodes``,
  ``optimized_texturecodes`` ([n_objects, latent]), ``psnr_eval``, ``ssim_eval`` (dicts keyed by object index).

Everything here is host-side dictionary plumbing; no arithmetic of the hot path.
"""
import os

import torch

MODEL_KEYS = ("model_params", "shape_code_params", "texture_code_params", "niter", "nepoch")
CODES_KEYS = ("ids", "num_obj", "optimized_shapecodes", "optimized_texturecodes", "psnr_eval", "ssim_eval")


def models_dict(model, shape_codes, texture_codes, niter, nepoch):
    """The dictionary trainer.py:166-171 builds.  `shape_codes` / `texture_codes`: nn.Embedding (or a [n, latent] tensor)."""
    def emb_state(e):
        return e.state_dict() if hasattr(e, "state_dict") else {"weight": torch.as_tensor(e)}
    return {"model_params": model.state_dict(), "shape_code_params": emb_state(shape_codes),
            "texture_code_params": emb_state(texture_codes), "niter": int(niter), "nepoch": int(nepoch)}


def save_models(save_dir, model, shape_codes, texture_codes, niter, nepoch, iteration=None):
    """trainer.py:165-174: always `models.pth`, and `<iteration>.pth` when a checkpoint iteration is given."""
    d = models_dict(model, shape_codes, texture_codes, niter, nepoch)
    os.makedirs(save_dir, exist_ok=True)
    if iteration is not None:
        torch.save(d, os.path.join(save_dir, f"{iteration}.pth"))
    torch.save(d, os.path.join(save_dir, "models.pth"))
    return d


def load_models(path, model=None, map_location="cpu"):
    """optimizer.py:208-216: load on the CPU, `load_state_dict` into `model` (strict: the key set is part of the
    boundary), and return (dict, mean shape code [1, latent], mean texture code [1, latent])."""
    saved = torch.load(path, map_location=torch.device(map_location), weights_only=False)
    missing = [k for k in MODEL_KEYS if k not in saved]
    if missing:
        raise KeyError(f"{path}: not a CodeNeRF models.pth (missing {missing})")
    if model is not None:
        model.load_state_dict(saved["model_params"])
    mean_shape = torch.mean(saved["shape_code_params"]["weight"], dim=0).reshape(1, -1)
    mean_texture = torch.mean(saved["texture_code_params"]["weight"], dim=0).reshape(1, -1)
    return saved, mean_shape, mean_texture


def codes_dict(ids, num_obj, shapecodes, texturecodes, psnr_eval=None, ssim_eval=None):
    return {"ids": ids, "num_obj": int(num_obj), "optimized_shapecodes": shapecodes,
            "optimized_texturecodes": texturecodes, "psnr_eval": psnr_eval if psnr_eval is not None else {},
            "ssim_eval": ssim_eval if ssim_eval is not None else {}}


def save_codes(save_dir, ids, num_obj, shapecodes, texturecodes, psnr_eval=None, ssim_eval=None):
    """optimizer.py:137-146 (`codes.pth` is rewritten after every test object)."""
    d = codes_dict(ids, num_obj, shapecodes, texturecodes, psnr_eval, ssim_eval)
    os.makedirs(save_dir, exist_ok=True)
    torch.save(d, os.path.join(save_dir, "codes.pth"))
    return d


def load_codes(path, map_location="cpu"):
    saved = torch.load(path, map_location=torch.device(map_location), weights_only=False)
    missing = [k for k in CODES_KEYS if k not in saved]
    if missing:
        raise KeyError(f"{path}: not a CodeNeRF codes.pth (missing {missing})")
    return saved
