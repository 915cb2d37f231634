"""Generates tests/golden/preprocess.pt by running the REFERENCE's own input pipeline on seeded synthetic images:
``ResizeLongestSide.apply_image`` (model/segment_anything/utils/transforms.py — PIL through torchvision),
``LazySupervisedDataset.preprocess`` / ``pad_tensor_channelwise`` (datasets/LazySupervisedDataset.py:446-502, called as
unbound methods on an uninitialised instance: they only read class attributes), the installed CLIPImageProcessor and
cv2's nearest resize (:519).  Inputs are NOT stored: tests/golden/inputs.py:preprocess_image regenerates them.

Run:  python tests/golden/make_golden_preprocess.py     (needs /root/reference, PIL, torchvision, cv2)
"""
import os
import sys
import types

import cv2
import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

ds, ds_moe, ds_layer = types.ModuleType("deepspeed"), types.ModuleType("deepspeed.moe"), types.ModuleType("deepspeed.moe.layer")
ds_layer.MoE = type("MoE", (nn.Module,), {})
ds.moe, ds_moe.layer = ds_moe, ds_layer
sys.modules.update({"deepspeed": ds, "deepspeed.moe": ds_moe, "deepspeed.moe.layer": ds_layer})

import PIL  # noqa: E402
import transformers  # noqa: E402
from transformers import CLIPImageProcessor  # noqa: E402

from datasets.LazySupervisedDataset import LazySupervisedDataset  # noqa: E402
from model.segment_anything.utils.transforms import ResizeLongestSide  # noqa: E402

import inputs as gi  # noqa: E402

import hashlib  # noqa: E402


def digest(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


dsobj = object.__new__(LazySupervisedDataset)  # preprocess() only reads the class-level pixel_mean / std / clip mean


def make_processor(size):
    """The reference pins transformers 4.31, whose CLIPImageProcessor is the numpy / PIL implementation (rescale in
    float64 -> fp32, then normalize).  transformers 5 keeps that arithmetic as CLIPImageProcessorPil; its default
    torchvision backend fuses rescale into normalize and differs in the last fp32 bit (<= 2.4e-7), so it is NOT used."""
    kw = dict(size={"shortest_edge": size}, crop_size={"height": size, "width": size})
    try:
        from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
        return CLIPImageProcessorPil(**kw)
    except ImportError:
        return CLIPImageProcessor(**kw)  # transformers < 5: the only implementation


out = {"versions": {"PIL": PIL.__version__, "transformers": transformers.__version__, "cv2": cv2.__version__},
       "cases": []}
for i, (h, w, ls, lc) in enumerate(gi.PREPROCESS_SIZES):
    img = gi.preprocess_image(i, h, w)
    r = ResizeLongestSide(ls).apply_image(img)
    image_sam = dsobj.preprocess(torch.from_numpy(r).permute(2, 0, 1).contiguous(), ls)
    rc = ResizeLongestSide(lc).apply_image(img)
    clip_u8 = dsobj.preprocess(torch.from_numpy(rc).permute(2, 0, 1).contiguous(), lc, normalize=False)
    image_clip = make_processor(lc).preprocess(clip_u8, return_tensors="pt")["pixel_values"][0]
    m = gi.preprocess_mask(i, h, w)
    rm = ResizeLongestSide(lc).apply_image(m)
    rm = dsobj.preprocess(torch.from_numpy(rm).contiguous(), lc, normalize=False, is_mask=True)
    rm_grid = cv2.resize(np.array(rm), None, fx=1 / 14, fy=1 / 14, interpolation=cv2.INTER_NEAREST)
    assert image_sam.dtype == torch.float32 and clip_u8.dtype == torch.uint8 and image_clip.dtype == torch.float32
    tensors = {"sam_u8": torch.from_numpy(r.copy()), "image_sam": image_sam, "clip_u8": clip_u8.clone(),
               "image_clip": image_clip, "region_u8": rm.clone(), "region_grid": torch.from_numpy(rm_grid.copy())}
    case = {"hw": (h, w), "targets": (ls, lc), "resize": tuple(r.shape[:2]),
            "sha256": {k: digest(v) for k, v in tensors.items()}, "region_grid": tensors["region_grid"]}
    if lc <= 72:
        case["tensors"] = tensors  # small enough to keep whole, for debugging a digest mismatch
    out["cases"].append(case)
    print(i, (h, w), (ls, lc), tuple(r.shape), tuple(rm_grid.shape))
# ---- ICL exemplar masks for the MaskTokenEncoder: ICLLazySupervisedDataset._preprocess_encoder_mask (:77-85)
from datasets.ICLLazySupervisedDataset import ICLLazySupervisedDataset  # noqa: E402

out["encoder_masks"] = []
for i, (h, w, ls, lc) in enumerate(gi.PREPROCESS_SIZES):
    icl = object.__new__(ICLLazySupervisedDataset)
    icl.transform_clip, icl.clip_img_size = ResizeLongestSide(lc), lc
    em = icl._preprocess_encoder_mask(gi.preprocess_mask(i, h, w))
    assert em.dtype == torch.float32 and tuple(em.shape) == (1, lc, lc)
    out["encoder_masks"].append({"sha256": digest(em), "ones": int(em.sum())})

# ---- sentinel tokenisation: the reference's tokenizer_image_token on a deterministic stub tokenizer
from datasets.LazySupervisedDataset import tokenizer_image_token  # noqa: E402

out["tokenize"] = [{"bos": bos, "prompt": pr, "ids": tokenizer_image_token(pr, gi.StubTokenizer(bos))}
                   for bos in (True, False) for pr in gi.TOKENIZE_PROMPTS]
torch.save(out, os.path.join(HERE, "preprocess.pt"))
print("preprocess.pt %.1f KiB" % (os.path.getsize(os.path.join(HERE, "preprocess.pt")) / 1024))
