"""Streaming serving loop: the step immediately after the hot path.

``generate_stream`` is the token-by-token loop of the reference's worker (model/serve/model_worker.py:296-541) over
this package's model: one prefill ``model(**kw)`` with the spliced image, then one ``model(input_ids=[[tok]],
past_key_values=cache)`` per new token (the persistent one-kernel decode step), the choice made on the device
(argmax, or temperature sampling), text streamed every ``stream_interval`` tokens, and — when the answer stops and
contains ``<SEG>`` — the mask tail (text_hidden_fcs on the row in front of the first ``<SEG>`` -> SAM-Med2D encoder ->
prompt encoder -> mask decoder -> ``sigmoid > 0.1`` -> sparse (row, col) list).  The HTTP / controller / heartbeat
plumbing around it (model_worker.py:60-160, 545-620) is out of scope: a FastAPI handler would wrap the
generator exactly as the reference's ``worker_generate_stream`` does.

Differences from the reference, all deliberate: the SAM-Med2D image encoder runs on a side stream as soon as the request
arrives instead of after the last token (same arithmetic, ready when the mask tail needs it); only the final-norm hidden
state is kept per step (``output_hidden_states=False`` — the reference keeps all 33 and reads ``[-1]``); the per-step
host read is the one token id (the reference also reads it, to test for EOS).
"""
import json

import torch

IMAGE_TOKEN_INDEX = -200   # utils/utils.py
REGION_TOKEN_INDEX = -300


def insert_region_placeholders(input_ids, id_open, id_close):
    """model_worker.py:310-318: put the REGION sentinel between every adjacent ``<region></region>`` pair."""
    ids = list(input_ids)
    i = 0
    while i < len(ids) - 1:
        if ids[i] == id_open and ids[i + 1] == id_close:
            ids.insert(i + 1, REGION_TOKEN_INDEX)
            i += 1
        i += 1
    return ids


def tokenize_with_sentinels(prompt, tokenizer, return_tensors=None):
    """``tokenizer_image_token`` (datasets/LazySupervisedDataset.py:353-388, model_worker.py:307-318): tokenise the text
    around every ``<image>`` separately, keep ONE leading BOS, put the IMAGE sentinel (-200) at each ``<image>`` and the
    REGION sentinel (-300) inside each ``<region></region>`` pair."""
    chunks = [tokenizer(c).input_ids for c in prompt.split("<image>")]
    bos = getattr(tokenizer, "bos_token_id", None)
    has_bos = bool(chunks and chunks[0] and chunks[0][0] == bos)
    ids = [chunks[0][0]] if has_bos else []
    for n, chunk in enumerate(chunks):
        if n:
            ids.append(IMAGE_TOKEN_INDEX)
        ids.extend(chunk[1:] if has_bos else chunk)
    ids = insert_region_placeholders(ids, tokenizer("<region>", add_special_tokens=False).input_ids[0],
                                     tokenizer("</region>", add_special_tokens=False).input_ids[0])
    if return_tensors is None:
        return ids
    if return_tensors != "pt":
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return torch.tensor(ids, dtype=torch.long)


def encode_sparse(mask):
    """model_worker.py:519-523: list of [row, col] of the non-zero pixels."""
    return torch.nonzero(torch.as_tensor(mask)).tolist()


def _spliced_extra(model, input_ids):
    """Rows the image sentinels add to the sequence (the reference hard-codes 575 = one 576-token image in front)."""
    n_img = int((input_ids == IMAGE_TOKEN_INDEX).sum())
    if n_img == 0:
        return 0
    if getattr(model.config, "mm_token_compress", False):
        per = getattr(model.config, "mm_compressed_token_count", 256)
    else:
        per = model.get_model().get_vision_tower().num_patches
    return n_img * (per - 1)


@torch.no_grad()
def generate_stream(model, tokenizer, input_ids, images_clip=None, images_sam=None, resize=None, original_size=None,
                    region_masks=None, valid_region_masks_bool=None, temperature=1.0, max_new_tokens=256,
                    stop_str=None, stream_interval=1, prompt_text="", mask_threshold=0.1, as_bytes=False,
                    forced_tokens=None):
    """Yields the reference's records ``{"text", "mask", "height", "width", "error_code"}`` (``as_bytes``: the exact
    ``json + b"\\0"`` frames of model_worker.py:537) while decoding ONE request.

    input_ids: [1, n] with the IMAGE / REGION sentinels already inserted (``tokenizer_image_token`` +
    ``insert_region_placeholders``); images_clip [1,3,336,336] / images_sam [1,3,256,256] in the model's dtype on its
    device (``medplib_b200.preprocess.ImagePreprocessor(out_dtype=torch.bfloat16)`` produces both from the raw image);
    resize = (h, w) at SAM scale, original_size = (H, W) of the raw image.  forced_tokens {step: id} overrides the
    choice at that step (tests / benchmarks with random weights, which never emit ``<SEG>``), as in ``generate``.
    """
    max_new_tokens = min(int(max_new_tokens), 1024)
    stop_idx = None
    if stop_str is not None:
        ids = tokenizer(stop_str).input_ids
        stop_idx = ids[0] if len(ids) == 1 else None
    eos = getattr(tokenizer, "eos_token_id", None)
    dev = input_ids.device
    attention_mask = input_ids.ne(tokenizer.pad_token_id) if getattr(tokenizer, "pad_token_id", None) is not None \
        else torch.ones_like(input_ids, dtype=torch.bool)
    extra = _spliced_extra(model, input_ids)

    # SAM-Med2D encoder: independent of the language model -> side stream, joined before the mask tail
    image_embeddings, side = None, None
    if images_sam is not None and images_sam.is_cuda:
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=images_sam.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            image_embeddings = model.get_visual_embs(images_sam)

    output_ids = input_ids[0].tolist()
    pred_ids, hidden_rows = [], []
    cache, cur = None, input_ids
    encoded_mask, height, width = [], 0, 0
    for i in range(max_new_tokens):
        out = model(input_ids=cur, use_cache=True, attention_mask=attention_mask, past_key_values=cache,
                    images=images_clip, region_masks=region_masks, valid_region_masks_bool=valid_region_masks_bool,
                    output_hidden_states=False, return_dict=True, logits_rows="last")
        cache = out.past_key_values
        last_logits = out.logits[0, -1]
        if temperature < 1e-4:
            token = int(torch.argmax(last_logits))
        else:
            token = int(torch.multinomial(torch.softmax(last_logits / temperature, dim=-1), num_samples=1))
        if forced_tokens is not None and i in forced_tokens:
            token = int(forced_tokens[i])
        output_ids.append(token)
        pred_ids.append(token)
        hidden_rows.append(out.hidden_states[-1][0])  # [T or 1, D], after the final RMSNorm
        stopped = (stop_idx is not None and token == stop_idx) or (eos is not None and token == eos)
        cur = torch.tensor([[token]], dtype=input_ids.dtype, device=dev)

        if i % stream_interval == 0 or i == max_new_tokens - 1 or stopped:
            cur_out = tokenizer.decode(pred_ids, skip_special_tokens=True)
            if stop_str:
                pos = cur_out.rfind(stop_str)
                if pos != -1:
                    cur_out, stopped = cur_out[:pos], True
            if stopped:
                # position t is the row in front of token t+1 (model_worker.py:449-461)
                is_seg = [t == model.seg_token_idx for t in output_ids[1:]]
                if any(is_seg) and images_sam is not None:
                    hidden = torch.cat(hidden_rows, dim=0)  # [T_spliced + n_new - 1, D]
                    row = extra + is_seg.index(True)  # the first <SEG> when there are several
                    pred_embeddings = model._seg_embeddings(hidden[row:row + 1])
                    if side is not None:
                        torch.cuda.current_stream().wait_stream(side)
                        image_embeddings.record_stream(torch.cuda.current_stream())
                    elif image_embeddings is None:
                        image_embeddings = model.get_visual_embs(images_sam)
                    masks, _ = model._decode_masks(pred_embeddings, image_embeddings, [tuple(resize)],
                                                   [tuple(original_size)])
                    pred = (torch.sigmoid(masks[0].float()) > mask_threshold).int().squeeze(0)
                    height, width = int(pred.shape[0]), int(pred.shape[1])
                    encoded_mask = encode_sparse(pred.cpu())
            ret = {"text": prompt_text + cur_out, "mask": encoded_mask, "height": str(height), "width": str(width),
                   "error_code": 0}
            yield (json.dumps(ret).encode() + b"\0") if as_bytes else ret
        if stopped:
            break
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)
