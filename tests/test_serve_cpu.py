"""CPU: host logic of the streaming serving loop (medplib_b200/serve.py) against the reference worker's
semantics (model/serve/model_worker.py:296-541) with a scripted stand-in for the model: which tokens are fed back, when
records are emitted, stop-token / stop-string / EOS handling, which hidden row feeds the mask tail, sparse mask encoding."""
import json
from types import SimpleNamespace

import torch

from medplib_b200 import serve

SEG, EOS, PAD, V, D = 50, 2, 0, 64, 8
N_PATCH = 4


class Tok:
    pad_token_id, eos_token_id = PAD, EOS

    def __call__(self, s, add_special_tokens=True):
        return SimpleNamespace(input_ids=[ord(c) - 32 for c in s])

    def decode(self, ids, skip_special_tokens=True):
        return "".join(chr(i + 32) for i in ids if not (skip_special_tokens and i in (EOS, PAD, SEG)))


class Model:
    """Emits `script` token by token; hidden row p carries the value p (spliced position) in every channel."""
    seg_token_idx = SEG
    config = SimpleNamespace(mm_token_compress=False)

    def __init__(self, script):
        self.script, self.calls, self.tail_rows = script, [], []

    def get_model(self):
        return SimpleNamespace(get_vision_tower=lambda: SimpleNamespace(num_patches=N_PATCH))

    def __call__(self, input_ids, past_key_values, images, **kw):
        assert kw["use_cache"] and kw["return_dict"] and kw["logits_rows"] == "last"
        self.calls.append(input_ids.tolist())
        if past_key_values is None:
            n_img = int((input_ids == serve.IMAGE_TOKEN_INDEX).sum())
            T = input_ids.shape[1] + n_img * (N_PATCH - 1)
            past, step = SimpleNamespace(len=T, step=0), 0
            hidden = torch.arange(T, dtype=torch.float32)[None, :, None].expand(1, T, D)
        else:
            assert input_ids.shape == (1, 1) and int(input_ids) == self.script[past_key_values.step - 1]
            past = SimpleNamespace(len=past_key_values.len + 1, step=past_key_values.step)
            hidden = torch.full((1, 1, D), float(past.len - 1))
        logits = torch.zeros(1, 1, V)
        logits[0, 0, self.script[past.step]] = 5.0
        past.step += 1
        return SimpleNamespace(logits=logits, past_key_values=past, hidden_states=(hidden,))

    def _seg_embeddings(self, rows):
        self.tail_rows.append(rows[:, 0].tolist())
        return rows

    def get_visual_embs(self, x):
        return torch.zeros(1, 4, 2, 2)

    def _decode_masks(self, pe, ie, resize_list, size_list):
        assert resize_list == [(3, 4)] and size_list == [(5, 6)]
        m = torch.full((1, 5, 6), -9.0)
        m[0, 1, 2] = m[0, 4, 5] = 3.0
        m[0, 0, 0] = -2.1  # sigmoid = 0.109 > 0.1: kept, like the reference's 0.1 threshold
        return [m], None


def prompt():
    return torch.tensor([[1, serve.IMAGE_TOKEN_INDEX, 40, 41, 42]])


def run(script, **kw):
    m = Model(script)
    recs = list(serve.generate_stream(m, Tok(), prompt(), images_clip=torch.zeros(1, 3, 4, 4),
                                      images_sam=torch.zeros(1, 3, 4, 4), resize=(3, 4), original_size=(5, 6),
                                      temperature=0.0, **kw))
    return m, recs


def test_tokens_are_fed_back_and_text_streams_until_eos():
    m, recs = run([33, 34, 35, EOS, 36], max_new_tokens=10)
    assert m.calls[0] == prompt().tolist() and m.calls[1:] == [[[33]], [[34]], [[35]]]  # nothing after EOS
    assert [r["text"] for r in recs] == ["A", "AB", "ABC", "ABC"]
    assert all(r["mask"] == [] and r["height"] == "0" and r["error_code"] == 0 for r in recs)
    assert m.tail_rows == []


def test_stop_string_and_single_token_stop():
    _, recs = run([33, 3, 3, 34, 35], max_new_tokens=10, stop_str="##", prompt_text="Q: ")
    assert recs[-1]["text"] == "Q: A" and len(recs) == 3  # "A", "A#", then "A##" cut at the stop string
    _, recs = run([33, 3, 34], max_new_tokens=10, stop_str="#")  # '#' is one token: stop on the id itself
    assert [r["text"] for r in recs] == ["A", "A"]


def test_stream_interval_and_max_new_tokens():
    _, recs = run([33, 34, 35, 36, 37, 38, 39, 40], max_new_tokens=6, stream_interval=4)
    assert [r["text"] for r in recs] == ["A", "ABCDE", "ABCDEF"]  # i = 0, 4 and the last step


def test_mask_tail_uses_the_row_in_front_of_the_first_seg():
    m, recs = run([33, SEG, 34, SEG, EOS], max_new_tokens=10, as_bytes=False)
    # spliced prompt = 5 ids + 3 extra image rows = 8 rows (0..7); generated token j sits at spliced position 8 + j, so
    # the row in front of the first <SEG> (generated index 1) is position 8 — the reference's 575-offset rule with 3
    assert m.tail_rows == [[8.0]]
    assert all(r["mask"] == [] for r in recs[:-1])
    last = recs[-1]
    assert last["mask"] == [[0, 0], [1, 2], [4, 5]] and (last["height"], last["width"]) == ("5", "6")
    assert last["text"] == "AB"


def test_byte_frames_match_the_worker_protocol():
    m = Model([33, EOS])
    frames = list(serve.generate_stream(m, Tok(), prompt(), temperature=0.0, max_new_tokens=4, as_bytes=True))
    assert all(f.endswith(b"\0") for f in frames)
    assert json.loads(frames[-1][:-1]) == {"text": "A", "mask": [], "height": "0", "width": "0", "error_code": 0}


def test_region_placeholders():
    o, c = 60, 61
    assert serve.insert_region_placeholders([5, o, c, 6, o, c], o, c) == [5, o, -300, c, 6, o, -300, c]
    assert serve.insert_region_placeholders([o, 7, c, o], o, c) == [o, 7, c, o]
    assert serve.encode_sparse(torch.tensor([[0, 1], [1, 0]])) == [[0, 1], [1, 0]]


def test_sentinel_tokenisation_matches_the_reference():
    """tests/golden/preprocess.pt['tokenize'] = the reference's own tokenizer_image_token on the stub tokenizer."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import inputs as gi
    gold = torch.load(os.path.join(here, "golden", "preprocess.pt"), weights_only=False)["tokenize"]
    assert len(gold) == 2 * len(gi.TOKENIZE_PROMPTS)
    for case in gold:
        got = serve.tokenize_with_sentinels(case["prompt"], gi.StubTokenizer(case["bos"]))
        assert got == case["ids"], case["prompt"]
    t = serve.tokenize_with_sentinels(gi.TOKENIZE_PROMPTS[0], gi.StubTokenizer(), return_tensors="pt")
    assert t.dtype == torch.long and t.tolist() == gold[0]["ids"]
