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


# ----------------------------------------------------------------------------------------------- continuous batching
class _EngineBackend:
    """Device side of the batcher: one shared KV cache [L, max_batch, H, max_len, 128], the B = 1 prefill of a new
    request (the tcgen05 path of ``model(...)``) copied into its slot, and ONE persistent decode kernel launch per step
    for every active slot (llama_decode.cu: per-sequence RoPE positions via mpl_llama_io.rope_pos, ragged lengths via
    the key mask)."""

    def __init__(self, model, max_batch, max_len):
        from . import ops
        self.ops, self.model = ops, model
        self.eng = model._llama()
        self.B, self.Tmax = max_batch, max_len
        self.cache = self.eng.new_cache(max_batch, max_len)
        dev = model.lm_head.weight.device
        self.mask = torch.zeros((max_batch, max_len), dtype=torch.uint8, device=dev)
        self.dev = dev

    def prefill(self, req):
        am = torch.ones_like(req["input_ids"], dtype=torch.bool)
        out = self.model(input_ids=req["input_ids"], use_cache=True, attention_mask=am, past_key_values=None,
                         images=req.get("images_clip"), region_masks=req.get("region_masks"),
                         valid_region_masks_bool=req.get("valid_region_masks_bool"), output_hidden_states=False,
                         return_dict=True, logits_rows="last")
        kv = out.past_key_values
        return out.logits[0, -1], out.hidden_states[-1][0], kv, int(kv.len)

    def install(self, slot, kv, start, n):
        self.cache.k[:, slot, :, start:start + n] = kv.k[:, 0, :, :n]
        self.cache.v[:, slot, :, start:start + n] = kv.v[:, 0, :, :n]
        self.mask[slot].zero_()
        self.mask[slot, start:start + n] = 1

    def release(self, slot):
        self.mask[slot].zero_()

    def decode(self, tokens, col, rope_pos):
        """tokens / rope_pos: python lists of max_batch ints (idle slots: any valid id, position 0). Returns
        (logits f32 [B, V], hidden bf16 [B, D])."""
        ops = self.ops
        self.mask[:, col] = 1  # every row attends to its own new key (idle rows: only that one -> finite garbage)
        idx = torch.tensor(tokens, dtype=torch.int32, device=self.dev)
        x = ops.gather_rows(idx, self.model.model.embed_tokens.weight).view(self.B, 1, -1)
        rp = torch.tensor(rope_pos, dtype=torch.int32, device=self.dev)
        self.cache.len = col
        out = self.eng.forward(x, self.cache, kv_mask=self.mask, rope_pos=rp)
        hidden = out["last_hidden_state"][:, -1]
        return ops.linear(hidden, self.model.lm_head.weight, out_dtype=torch.float32), hidden


class ContinuousBatcher:
    """Continuous batching over the serving loop: up to ``max_batch`` (<= 8) requests decode together,
    one persistent-kernel launch per token step for all of them; a request joins as soon as a slot is free (its prefill
    runs alone, then its keys are copied into the slot) and leaves when it stops, without waiting for the others.

    Layout: all sequences share the cache write column ``col``; sequence b owns columns [start_b, col) of row b and the
    key mask hides everything else, and its RoPE angle is its OWN token count (mpl_llama_io.rope_pos), so a sequence
    admitted late -- or after the column jumped to make room for a longer prompt -- computes exactly what it would
    compute alone. Records have the reference worker's shape (model_worker.py:537); the mask tail runs per request when
    it stops (text_hidden_fcs on the row in front of its first ``<SEG>``).
    """

    def __init__(self, model, tokenizer, max_batch=8, max_len=2048, temperature=0.0, stop_str=None, mask_threshold=0.1,
                 backend=None):
        assert 1 <= max_batch <= 8
        self.model, self.tok = model, tokenizer
        self.B, self.Tmax = max_batch, max_len
        self.temperature, self.stop_str, self.thr = temperature, stop_str, mask_threshold
        self.stop_idx = None
        if stop_str is not None:
            ids = tokenizer(stop_str).input_ids
            self.stop_idx = ids[0] if len(ids) == 1 else None
        self.backend = backend if backend is not None else _EngineBackend(model, max_batch, max_len)
        self.slots = [None] * max_batch
        self.queue = []
        self.col = 0
        self.next_id = 0
        self.steps = 0

    # ------------------------------------------------------------------ requests
    def submit(self, input_ids, images_clip=None, images_sam=None, resize=None, original_size=None, max_new_tokens=256,
               region_masks=None, valid_region_masks_bool=None, prompt_text="", forced_tokens=None, request_id=None):
        rid = self.next_id if request_id is None else request_id
        self.next_id += 1
        self.queue.append(dict(id=rid, input_ids=input_ids, images_clip=images_clip, images_sam=images_sam, resize=resize,
                               original_size=original_size, max_new=min(int(max_new_tokens), 1024),
                               region_masks=region_masks, valid_region_masks_bool=valid_region_masks_bool,
                               prompt_text=prompt_text, forced=forced_tokens or {}))
        return rid

    def idle(self):
        return not self.queue and all(s is None for s in self.slots)

    def _choose(self, logits_row, req, i):
        if self.temperature < 1e-4:
            tok = int(torch.argmax(logits_row))
        else:
            tok = int(torch.multinomial(torch.softmax(logits_row / self.temperature, dim=-1), num_samples=1))
        return int(req["forced"].get(i, tok))

    def _admit(self, records):
        """Queued requests take free slots: prefill alone, keys copied behind the shared column."""
        while self.queue and any(s is None for s in self.slots):
            if all(s is None for s in self.slots):
                self.col = 0  # nothing in flight: start the column over
            req = self.queue[0]
            n_est = req["input_ids"].shape[1] + _spliced_extra(self.model, req["input_ids"])
            if max(self.col, n_est) + req["max_new"] + 1 > self.Tmax:
                if all(s is None for s in self.slots):
                    raise ValueError("request does not fit max_len")
                return  # drain first; the column restarts at 0 once the batch is empty
            logits, hidden, kv, n = self.backend.prefill(req)
            start = max(self.col, n) - n
            self.queue.pop(0)
            slot = self.slots.index(None)
            self.col = max(self.col, n)
            self.backend.install(slot, kv, start, n)
            st = dict(req=req, n_prompt=n, extra=n - req["input_ids"].shape[1], pred=[], seg_row=None, done=False)
            # a <SEG> already in the prompt: the row in front of the first one (model_worker.py:449-461)
            ids0 = req["input_ids"][0].tolist()
            in_prompt = [t == self.model.seg_token_idx for t in ids0[1:]]
            if any(in_prompt):
                r = st["extra"] + in_prompt.index(True)
                st["seg_row"] = hidden[r:r + 1].clone()
            tok = self._choose(logits, req, 0)
            self._accept(st, tok, hidden[-1:], records)
            self.slots[slot] = None if st["done"] else st
            if st["done"]:
                self.backend.release(slot)

    def _accept(self, st, tok, hidden_row, records):
        """Token `tok` was produced from `hidden_row` (the final-norm row in front of it)."""
        req = st["req"]
        if tok == self.model.seg_token_idx and st["seg_row"] is None:
            st["seg_row"] = hidden_row.clone()
        st["pred"].append(tok)
        eos = getattr(self.tok, "eos_token_id", None)
        stopped = (self.stop_idx is not None and tok == self.stop_idx) or (eos is not None and tok == eos) or \
            len(st["pred"]) >= req["max_new"]
        text = self.tok.decode(st["pred"], skip_special_tokens=True)
        if self.stop_str:
            pos = text.rfind(self.stop_str)
            if pos != -1:
                text, stopped = text[:pos], True
        mask, h, w = [], 0, 0
        if stopped:
            st["done"] = True
            if st["seg_row"] is not None and req["images_sam"] is not None:
                emb = self.model._seg_embeddings(st["seg_row"])
                img = self.model.get_visual_embs(req["images_sam"])
                masks, _ = self.model._decode_masks(emb, img, [tuple(req["resize"])], [tuple(req["original_size"])])
                pred = (torch.sigmoid(masks[0].float()) > self.thr).int().squeeze(0)
                h, w = int(pred.shape[0]), int(pred.shape[1])
                mask = encode_sparse(pred.cpu())
        records.append((req["id"], {"text": req["prompt_text"] + text, "mask": mask, "height": str(h), "width": str(w),
                                    "error_code": 0}, stopped))

    # ------------------------------------------------------------------ one token step for every active slot
    @torch.no_grad()
    def step(self):
        """Admit what fits, then decode ONE token for every active request. Returns [(request id, record, done)]."""
        records = []
        self._admit(records)
        active = [b for b, s in enumerate(self.slots) if s is not None]
        if not active:
            return records
        tokens = [s["pred"][-1] if s is not None else 0 for s in self.slots]
        rope = [s["n_prompt"] + len(s["pred"]) - 1 if s is not None else 0 for s in self.slots]
        logits, hidden = self.backend.decode(tokens, self.col, rope)
        self.col += 1
        self.steps += 1
        for b in active:
            st = self.slots[b]
            tok = self._choose(logits[b], st["req"], len(st["pred"]))
            self._accept(st, tok, hidden[b:b + 1], records)
            if st["done"]:
                self.slots[b] = None
                self.backend.release(b)
        return records

    def run(self):
        """Drive every submitted request to completion; returns {request id: final record}."""
        final = {}
        while not self.idle():
            for rid, rec, done in self.step():
                if done:
                    final[rid] = rec
        return final
