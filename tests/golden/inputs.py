"""Deterministic inputs of the golden fixtures (torch CPU generator, fixed seeds). Shared by make_golden.py (which runs
the reference on them) and by the tests (which run the oracle / the CUDA path on them), so only the reference's OUTPUTS
need to be stored in the .pt files."""
import torch

SAM_ENC_CFG = dict(embed_dim=64, depth=4, num_heads=1, image_size=256, patch_size=16, out_chans=32)
SAM_ENC_SEED, SAM_HEAD_SEED = 11, 12


def sam_encoder_images():
    return torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16).float()


def sam_head_inputs(dim=256):
    g = torch.Generator().manual_seed(6)
    emb = torch.randn(1, dim, 16, 16, generator=g).to(torch.bfloat16).float()
    text = torch.randn(1, 1, dim, generator=g).to(torch.bfloat16).float()
    return emb, text


def arch_inputs(D=64, Dv=32):
    g = torch.Generator().manual_seed(7)
    params = {}
    shapes = dict(projector={"0.weight": (D, Dv), "0.bias": (D,), "2.weight": (D, D), "2.bias": (D,)},
                  compressor={"norm.weight": (D,), "norm.bias": (D,), "proj.weight": (D, D), "proj.bias": (D,)},
                  mask_encoder={"encoder.0.weight": (64, 1, 3, 3), "encoder.0.bias": (64,),
                                "encoder.2.weight": (128, 64, 3, 3), "encoder.2.bias": (128,),
                                "encoder.4.weight": (256, 128, 3, 3), "encoder.4.bias": (256,),
                                "encoder.6.weight": (256, 256, 3, 3), "encoder.6.bias": (256,),
                                "proj.weight": (D, 256), "proj.bias": (D,), "norm.weight": (D,), "norm.bias": (D,)})
    for mod, d in shapes.items():
        params[mod] = {k: (torch.randn(s, generator=g) * (0.2 if len(s) < 4 else (s[1] * 9) ** -0.5)) for k, s in d.items()}
    feats = torch.randn(2, 576, Dv, generator=g)
    masks = (torch.rand(2, 1, 336, 336, generator=g) > 0.8).float()
    fmap = torch.randn(2, 576, D, generator=g)
    rmasks = [[(torch.rand(24, 24, generator=g) > 0.7).float(), (torch.rand(24, 24, generator=g) > 0.9).float()],
              [(torch.rand(24, 24, generator=g) > 0.5).float()]]
    return params, feats, masks, fmap, rmasks


def splice_inputs(use_se, D=16, V=50, n_img=6):
    g = torch.Generator().manual_seed(8 + int(use_se))
    embed = torch.randn(V, D, generator=torch.Generator().manual_seed(9))
    ids = torch.randint(3, V, (2, 12), generator=g)
    ids[0, 2], ids[0, 7] = -200, -300
    ids[1, 1] = -200
    labels = ids.clone()
    labels[labels < 0] = -100
    am = torch.ones(2, 12, dtype=torch.bool)
    am[1, 9:] = False
    feats = torch.randn(2, n_img, D, generator=g)
    feats_r = torch.randn(2, 4, D, generator=g)  # 2x2 feature map for the region path
    rm = [[(torch.rand(24, 24, generator=g) > 0.6).float()]]  # one entry per VALID sample
    valid = [[True], []]
    ids2 = ids.clone()
    ids2[0, 7] = 5
    return dict(embed=embed, ids=ids, ids2=ids2, labels=labels, am=am, feats=feats, feats_r=feats_r, region_masks=rm,
                valid=valid)


def icl_splice_inputs(use_se, D=16, V=50, n_img_tok=4, n_mask_tok=2):
    """MedPLIB-ICL separate mode (medplib_arch.py:246-266): per sample a stack of image features and a stack of
    exemplar-mask features, consumed by the IMAGE sentinels in the order image_token_types names."""
    g = torch.Generator().manual_seed(18 + int(use_se))
    embed = torch.randn(V, D, generator=torch.Generator().manual_seed(9))
    ids = torch.randint(3, V, (2, 16), generator=g)
    types = [["image", "mask", "image"], ["image", "mask", "image", "mask", "image"]]
    for b, pos in enumerate(((1, 5, 9), (0, 3, 6, 9, 12))):
        for p in pos:
            ids[b, p] = -200
    labels = ids.clone()
    labels[labels < 0] = -100
    am = torch.ones(2, 16, dtype=torch.bool)
    img = [torch.randn(2, n_img_tok, D, generator=g), torch.randn(3, n_img_tok, D, generator=g)]
    msk = [torch.randn(1, n_mask_tok, D, generator=g), torch.randn(2, n_mask_tok, D, generator=g)]
    return dict(embed=embed, ids=ids, labels=labels, am=am, img=img, msk=msk, types=types)


def heads_inputs():
    g = torch.Generator().manual_seed(10)
    low = torch.randn(1, 1, 64, 64, generator=g)
    ids = torch.randint(3, 40, (2, 10), generator=g)
    ids[0, 1], ids[0, 6], ids[1, 0], ids[1, 4], ids[1, 9] = -200, 42, -200, -200, 42
    pred = torch.randn(1, 30, 40, generator=g) * 3
    gt = (torch.rand(1, 30, 40, generator=g) > 0.6).float()
    piou = torch.rand(1, generator=g)
    return dict(low=low, ids=ids, pred=pred, gt=gt, piou=piou)


POSTPROCESS_CASES = dict(square=((256, 256), (336, 336)), wide=((256, 192), (300, 225)), tiny=((48, 256), (60, 320)))


# ---- image input pipeline: (H, W, SAM target, CLIP target); targets other than 256 / 336 keep the
# fixture small — the reference's ResizeLongestSide / preprocess are size-parametric
PREPROCESS_SIZES = [(512, 512, 256, 336), (300, 451, 256, 336), (1024, 768, 256, 336), (97, 130, 256, 336),
                    (256, 256, 256, 336), (336, 200, 256, 336), (1300, 1777, 256, 336),
                    (50, 40, 64, 84), (211, 97, 64, 84), (64, 84, 64, 84), (500, 333, 32, 42), (31, 200, 48, 70)]


def preprocess_image(i, h, w):
    """u8 RGB [h, w, 3] numpy: low-frequency structure + noise, so that both antialiasing and rounding matter."""
    g = torch.Generator().manual_seed(100 + i)
    base = torch.nn.functional.interpolate(torch.rand(1, 3, 8, 8, generator=g), size=(h, w), mode="bilinear")[0]
    img = (base * 200 + torch.rand(3, h, w, generator=g) * 80 - 12).clamp(0, 255)
    return img.permute(1, 2, 0).contiguous().to(torch.uint8).numpy()


def preprocess_mask(i, h, w):
    """u8 {0,1} [h, w] numpy region mask: a filled ellipse."""
    g = torch.Generator().manual_seed(200 + i)
    cy, cx, ry, rx = (torch.rand(4, generator=g) * torch.tensor([h, w, h / 2, w / 2]) + torch.tensor([0, 0, 2, 2])).tolist()
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1).to(torch.uint8).numpy()


class StubTokenizer:
    """Deterministic stand-in for the LLaMA tokenizer (no tokenizer model offline): whitespace words -> ids by a fixed
    hash, `<region>` / `</region>` / `<im_start>` as single added tokens, optional BOS — enough to exercise the sentinel
    logic of tokenizer_image_token (datasets/LazySupervisedDataset.py:353-388)."""
    ADDED = {"<region>": 32005, "</region>": 32006, "<im_start>": 32001, "<im_end>": 32002, "<SEG>": 32003}

    def __init__(self, bos=True):
        self.bos_token_id = 1 if bos else None
        self.pad_token_id, self.eos_token_id = 0, 2

    def __call__(self, text, add_special_tokens=True):
        import re
        from types import SimpleNamespace
        ids = [1] if (self.bos_token_id is not None and add_special_tokens) else []
        for piece in re.findall(r"</?[a-z_A-Z]+>|\S+", text):
            ids.append(self.ADDED.get(piece, 3 + sum(ord(c) * (i + 7) for i, c in enumerate(piece)) % 31000))
        return SimpleNamespace(input_ids=ids)


TOKENIZE_PROMPTS = [
    "USER: <image>\nWhat is shown in <region></region> ? ASSISTANT:",
    "<image>",
    "no image here, only <region></region> and <region></region>",
    "<im_start><image><im_end> two images <image> and text <region> x </region>",
    "",
    "trailing image <image>",
]
