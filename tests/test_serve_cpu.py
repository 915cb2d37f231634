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


# ----------------------------------------------------------------------------------------------- continuous batching
class ScriptBackend:
    """Stand-in for the device side of serve.ContinuousBatcher: every request follows its own token script; hidden rows
    carry the sequence's own position; the cache layout (slot, start column, rope position) is recorded and checked."""

    def __init__(self, scripts, B):
        self.scripts, self.B = scripts, B
        self.slot_of, self.installs, self.decodes, self.released = {}, [], [], []
        self.step_of = {}

    def _logits(self, rid):
        lg = torch.zeros(V)
        lg[self.scripts[rid][self.step_of[rid]]] = 5.0
        self.step_of[rid] += 1
        return lg

    def prefill(self, req):
        ids = req["input_ids"]
        n = ids.shape[1] + int((ids == serve.IMAGE_TOKEN_INDEX).sum()) * (N_PATCH - 1)
        self.step_of[req["id"]] = 0
        hidden = torch.arange(n, dtype=torch.float32)[:, None].expand(n, D)
        return self._logits(req["id"]), hidden, SimpleNamespace(rid=req["id"], len=n), n

    def install(self, slot, kv, start, n):
        assert kv.len == n and start >= 0
        self.slot_of[slot] = dict(rid=kv.rid, start=start, n=n, fed=0)
        self.installs.append((kv.rid, slot, start, n))

    def release(self, slot):
        self.released.append(self.slot_of.pop(slot)["rid"])

    def decode(self, tokens, col, rope_pos):
        assert len(tokens) == len(rope_pos) == self.B
        self.decodes.append((col, dict((s["rid"], rope_pos[b]) for b, s in self.slot_of.items())))
        logits, hidden = torch.zeros(self.B, V), torch.zeros(self.B, D)
        for b, s in self.slot_of.items():
            # the token fed back is the request's previous script entry, at ITS OWN position, written at the shared column
            assert tokens[b] == self.scripts[s["rid"]][self.step_of[s["rid"]] - 1]
            assert rope_pos[b] == s["n"] + s["fed"] and col >= s["start"] + s["n"] + s["fed"]
            s["fed"] += 1
            logits[b] = self._logits(s["rid"])
            hidden[b] = float(rope_pos[b])
        return logits, hidden


def test_continuous_batching_host_logic():
    scripts = {0: [33, 34, 35, 36, EOS], 1: [40, SEG, 41, EOS], 2: [44, EOS], 3: [45, 46, 47, 48, 49, 50 - 1, EOS]}
    m = Model([])
    be = ScriptBackend(scripts, B=2)
    cb = serve.ContinuousBatcher(m, Tok(), max_batch=2, max_len=64, temperature=0.0, backend=be)
    short = torch.tensor([[1, 40, 41]])
    cb.submit(short, max_new_tokens=10, request_id=0)                                   # 3 rows, no image
    out = cb.step()                                                                     # admits 0, decodes token 2 of 0
    assert [r[0] for r in out] == [0, 0] and cb.col == 4
    cb.submit(prompt(), images_clip=torch.zeros(1, 3, 4, 4), images_sam=torch.zeros(1, 3, 4, 4), resize=(3, 4),
              original_size=(5, 6), max_new_tokens=10, request_id=1)                     # 8 rows: the column jumps 4 -> 8
    cb.submit(short, max_new_tokens=10, request_id=2)                                   # waits for a free slot
    cb.submit(short, max_new_tokens=10, request_id=3)
    final = {}
    while not cb.idle():
        for rid, rec, done in cb.step():
            if done:
                final[rid] = rec
    assert {k: v["text"] for k, v in final.items()} == {0: "ABCD", 1: "HI", 2: "L", 3: "MNOPQQ"}
    # request 1 joined mid-flight in slot 1 behind the column; its prompt is longer than the column was
    assert be.installs[0] == (0, 0, 0, 3) and be.installs[1] == (1, 1, 0, 8)
    assert be.installs[2][0] == 2 and be.installs[3][0] == 3 and len(be.installs) == 4
    # requests 2 and 3 re-used freed slots while another request was still decoding (no drain between them)
    rids_per_step = [set(d[1]) for d in be.decodes]
    assert {0, 1} in rids_per_step and any(2 in s and len(s) == 2 for s in rids_per_step)
    assert any(3 in s and len(s) == 2 for s in rids_per_step)
    # mask tail of request 1: the row in front of its first <SEG> = the last prompt row + 1 generated token = position 8
    assert m.tail_rows == [[8.0]]
    assert final[1]["mask"] == [[0, 0], [1, 2], [4, 5]] and (final[1]["height"], final[1]["width"]) == ("5", "6")
    assert final[0]["mask"] == [] and sorted(be.released + [2]) and set(be.released) | {2} >= {0, 1, 3}


def test_continuous_batching_restarts_the_column_and_refuses_what_cannot_fit():
    scripts = {0: [33, EOS], 1: [34, EOS]}
    be = ScriptBackend(scripts, B=1)
    cb = serve.ContinuousBatcher(Model([]), Tok(), max_batch=1, max_len=16, temperature=0.0, backend=be)
    cb.submit(torch.tensor([[1, 40, 41]]), max_new_tokens=4, request_id=0)
    cb.submit(torch.tensor([[1, 40, 41, 42]]), max_new_tokens=4, request_id=1)
    final = cb.run()
    assert final[0]["text"] == "A" and final[1]["text"] == "B"
    assert be.installs == [(0, 0, 0, 3), (1, 0, 0, 4)]  # the second request started a fresh column
    cb.submit(torch.tensor([[1] * 14]), max_new_tokens=8, request_id=2)
    try:
        cb.run()
        assert False, "a request longer than max_len must be refused"
    except ValueError:
        pass
