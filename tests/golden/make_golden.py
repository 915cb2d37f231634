"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own code (imported from /root/reference
in the authoring container; it does not exist on the GPU box, so only the committed .pt files travel).

Each fixture stores inputs + the reference's fp32 outputs; weights are NOT stored — they are regenerated
deterministically by oracle/weights.py (torch CPU generator, fixed seeds), loaded into the reference's own modules
here with load_state_dict(strict=True) and into the oracle / the CUDA path in the tests.

Run:  python tests/golden/make_golden.py        (needs /root/reference; deepspeed is stubbed)
"""
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

# ---- stub the absent deepspeed import (model/MedPLIB.py:21, medplib_moe_llama.py:30)
ds, ds_moe, ds_layer = types.ModuleType("deepspeed"), types.ModuleType("deepspeed.moe"), types.ModuleType("deepspeed.moe.layer")


class _MoE(nn.Module):
    pass


ds_layer.MoE = _MoE
ds.moe, ds_moe.layer = ds_moe, ds_layer
sys.modules.update({"deepspeed": ds, "deepspeed.moe": ds_moe, "deepspeed.moe.layer": ds_layer})
torch.Tensor.cuda = lambda self, *a, **k: self  # MedPLIB.py:302 / LISA.py:316 call .cuda() unconditionally

from oracle import weights  # noqa: E402

torch.manual_seed(0)
f32 = torch.float32


def save(name, **kw):
    path = os.path.join(HERE, name + ".pt")
    torch.save(kw, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def strip(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


# ------------------------------------------------------------------------------------------------ SAM-Med2D encoder
from functools import partial  # noqa: E402

from model.segment_anything_med2d.modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, TwoWayTransformer  # noqa: E402

import inputs as gi  # noqa: E402  (tests/golden/inputs.py)


def make_sam_encoder():
    c = gi.SAM_ENC_CFG
    enc = ImageEncoderViT(depth=c["depth"], embed_dim=c["embed_dim"], img_size=c["image_size"], mlp_ratio=4,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=c["num_heads"],
                          patch_size=c["patch_size"], qkv_bias=True, use_rel_pos=True,
                          global_attn_indexes=[2, 5, 8, 11], window_size=14, out_chans=c["out_chans"],
                          adapter_train=True)
    enc.load_state_dict(weights.sam_encoder(c, seed=gi.SAM_ENC_SEED, dtype=f32), strict=True)
    with torch.no_grad():
        out = enc(gi.sam_encoder_images())
    save("sam_encoder", out=out)


def make_sam_head():
    dim = 256
    pe = PromptEncoder(embed_dim=dim, image_embedding_size=(16, 16), input_image_size=(256, 256), mask_in_chans=16)
    md = MaskDecoder(num_multimask_outputs=3,
                     transformer=TwoWayTransformer(depth=2, embedding_dim=dim, mlp_dim=2048, num_heads=8),
                     transformer_dim=dim, iou_head_depth=3, iou_head_hidden_dim=256)
    sd = weights.sam_head(seed=gi.SAM_HEAD_SEED, dtype=f32)
    pe.load_state_dict(strip(sd, "prompt_encoder."), strict=False)  # point/mask-downscaling params are unused here
    md.load_state_dict(strip(sd, "mask_decoder."), strict=True)
    emb, text = gi.sam_head_inputs()
    with torch.no_grad():
        sparse, dense = pe(points=None, boxes=None, masks=None, text_embeds=text)
        dpe = pe.get_dense_pe()
        masks, iou = md(image_embeddings=emb, image_pe=dpe, sparse_prompt_embeddings=sparse,
                        dense_prompt_embeddings=dense, multimask_output=False)
        masks3, iou3 = md(image_embeddings=emb, image_pe=dpe, sparse_prompt_embeddings=sparse,
                          dense_prompt_embeddings=dense, multimask_output=True)
    save("sam_head", dense_pe_sample=dpe[0, :, ::5, ::5].clone(), masks=masks, iou=iou,
         masks3_sample=masks3[:, :, ::8, ::8].clone(), iou3=iou3)


# ------------------------------------------------------------------------------------------------ multimodal glue
from model.medplib.model import medplib_arch as ref_arch  # noqa: E402
from model.medplib.model.multimodal_projector.builder import build_vision_projector  # noqa: E402


def make_arch():
    D, Dv = 64, 32
    params, feats, masks, fmap, rmasks = gi.arch_inputs(D, Dv)
    cfg = types.SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=Dv, hidden_size=D)
    proj = build_vision_projector(cfg)
    comp = ref_arch.TokenCompressor(D, 256)
    menc = ref_arch.MaskTokenEncoder(D, 64)
    proj.load_state_dict(params["projector"], strict=True)
    comp.load_state_dict(params["compressor"], strict=True)
    menc.load_state_dict(params["mask_encoder"], strict=True)
    with torch.no_grad():
        pj = proj(feats)
        cp = comp(pj)
        me = menc(masks)
    fake = types.SimpleNamespace(get_model=lambda: types.SimpleNamespace(max_sample_point=512))
    with torch.no_grad():
        rf = ref_arch.LlavaMetaForCausalLM.extract_region_feature(fake, fmap, rmasks, original_dtype=f32,
                                                                  return_dtype=f32)
    # outputs are subsampled along tokens to keep the fixture small (every 7th token, all channels)
    save("arch", proj_out=pj[:, ::7].clone(), comp_out=cp[:, ::5].clone(), menc_out=me.clone(), region_out=rf)


def make_splice():
    """prepare_inputs_labels_for_multimodal driven through a minimal stand-in for the model object."""

    class Fake(ref_arch.LlavaMetaForCausalLM):
        def __init__(self, use_se, embed):
            V, D = embed.shape
            self.config = types.SimpleNamespace(mm_use_im_start_end=use_se, tune_mm_mlp_adapter=False,
                                                mm_token_compress=False, region_geo_sampler=False)
            self.embed = nn.Embedding(V, D)
            self.embed.weight.data = embed.clone()
            self.device = torch.device("cpu")
            self._tower = types.SimpleNamespace(dummy_feature=torch.zeros(1, D))
            self._model = types.SimpleNamespace(embed_tokens=self.embed, max_sample_point=512,
                                                get_vision_tower=lambda: self._tower)

        def get_model(self):
            return self._model

        def get_vision_tower(self):
            return self._tower

        def encode_images(self, images, region_flag=False, region_geo_sampler=False):
            # `images` here ARE the per-image features [n, n_img, D]; region map = 2 * features
            return images, images, (2 * images if region_flag else None)

        def encode_masks(self, masks):
            return masks  # likewise: the exemplar-mask "images" are already their token features

    cases = []
    for use_se in (True, False):
        d = gi.splice_inputs(use_se)
        fake = Fake(use_se, d["embed"])
        with torch.no_grad():
            _, am1, _, emb1, lab1 = fake.prepare_inputs_labels_for_multimodal(
                d["ids"], d["am"], None, d["labels"], d["feats_r"], d["region_masks"], d["valid"])
            _, am2, _, emb2, lab2 = fake.prepare_inputs_labels_for_multimodal(
                d["ids2"], d["am"], None, d["labels"], d["feats"], None, None)
        # ICL separate mode: stacks of images / masks per sample, ragged lengths
        di = gi.icl_splice_inputs(use_se)
        fake = Fake(use_se, di["embed"])
        with torch.no_grad():
            _, am3, _, emb3, lab3 = fake.prepare_inputs_labels_for_multimodal(
                di["ids"], di["am"], None, di["labels"], di["img"], None, None, mask_images=di["msk"],
                image_token_types=di["types"])
        cases.append(dict(use_se=use_se, emb1=emb1, lab1=lab1, am1=am1, emb2=emb2, lab2=lab2, am2=am2,
                          emb3=emb3, lab3=lab3, am3=am3))
    save("splice", cases=cases)


# ------------------------------------------------------------------------------------------------ heads + losses
def make_heads():
    from model import MedPLIB as ref  # noqa
    d = gi.heads_inputs()
    cls = ref.MedPLIBForCausalLM
    pp = {}
    for name, (inp, orig) in gi.POSTPROCESS_CASES.items():
        pp[name] = cls.postprocess_masks(None, d["low"], inp, orig)[..., ::3, ::3].clone()
    fake = types.SimpleNamespace(seg_token_idx=42, config=types.SimpleNamespace(mm_token_compress=False),
                                 get_model=lambda: types.SimpleNamespace(
                                     get_vision_tower=lambda: types.SimpleNamespace(num_patches=5)))
    m1 = cls.build_seg_token_mask(fake, d["ids"])
    m2 = cls.build_seg_token_mask(fake, d["ids"], image_token_lengths=[[3], [2, 4]])
    pred, gt, piou = d["pred"], d["gt"], d["piou"]
    losses = dict(bce=ref.sigmoid_ce_loss(pred, gt, num_masks=1), dice=ref.dice_loss(pred, gt, num_masks=1),
                  iou=ref.MaskIoULoss()(pred, gt, piou), focal=ref.FocalLoss()(pred, gt))
    save("heads", postprocess=pp, seg_mask=m1, seg_mask_lengths=m2, losses=losses)


if __name__ == "__main__":
    make_sam_encoder()
    make_sam_head()
    make_arch()
    make_splice()
    make_heads()
