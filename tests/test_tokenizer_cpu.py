"""CLIP byte-level BPE (lang-seg_b200/tokenizer.py) against an independent implementation of the same algorithm —
transformers' CLIPTokenizer (Rust `tokenizers` BPE backend) — on a SYNTHETIC merges table: the released
bpe_simple_vocab_16e6.txt.gz is not available offline, but the algorithm is data-independent, so a merges table
learned here from a small corpus exercises the same code paths (rank-ordered merging, </w> marker, byte fallback for
non-ASCII, the contraction / letter / digit / punctuation pre-tokeniser, lower-casing, SOT / EOT / padding / overflow)."""
import collections
import gzip
import os

import pytest
import torch

import lseg_b200  # noqa: F401
from lseg_b200 import tokenizer as T
from parity_util import synth


def _learn_merges(words, n_merges):
    """Plain BPE training (most frequent adjacent pair first, ties broken lexicographically) over byte-encoded words."""
    be = T.bytes_to_unicode()
    vocab = collections.Counter()
    for w in words:
        sym = [be[b] for b in w.encode("utf-8")]
        sym[-1] += "</w>"
        vocab[tuple(sym)] += 1
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for sym, c in vocab.items():
            for a, b in zip(sym[:-1], sym[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])
        merges.append(best)
        new_vocab = collections.Counter()
        for sym, c in vocab.items():
            out, i = [], 0
            while i < len(sym):
                if i < len(sym) - 1 and (sym[i], sym[i + 1]) == best:
                    out.append(sym[i] + sym[i + 1])
                    i += 2
                else:
                    out.append(sym[i])
                    i += 1
            new_vocab[tuple(out)] += c
        vocab = new_vocab
    return merges


CORPUS = ("wall building sky floor tree ceiling road bed windowpane grass cabinet sidewalk person earth door table "
          "mountain plant curtain chair car water painting sofa shelf house sea mirror rug field armchair seat fence desk "
          "rock wardrobe lamp bathtub railing cushion base box column signboard chest of drawers counter sand sink "
          "skyscraper fireplace refrigerator grandstand path stairs runway case pool table pillow screen door stairway "
          "river bridge bookcase blind coffee table toilet flower book hill bench countertop stove palm kitchen island "
          "computer swivel chair boat bar arcade machine hovel bus towel light truck tower chandelier awning streetlight "
          "booth television receiver airplane dirt track apparel pole land bannister escalator ottoman bottle buffet "
          "poster stage van ship fountain conveyer belt canopy washer plaything swimming pool stool barrel basket "
          "waterfall tent bag minibike cradle oven ball food step tank trade name microwave pot animal bicycle lake "
          "dishwasher screen blanket sculpture hood sconce vase traffic light tray ashcan fan pier crt screen plate "
          "monitor bulletin board shower radiator glass clock flag it's don't they're we've i'm you'll he'd 42 2024").split()

SAMPLES = ["wall", "a photo of a cat", "Traffic Light", "the dog's bowl, isn't it?", "crt screen;  monitor", "others",
           "café au lait", "naïve 3d-printed Ünïcödé!!", "x" * 5, "it's 2024: they've 12 cats & 3 dogs", "",
           "swimming pool / kitchen island", "<|startoftext|> inner marker"]


@pytest.fixture(scope="module")
def tok():
    return T.BPETokenizer(_learn_merges(CORPUS, 400))


def test_vocabulary_layout(tok):
    assert len(tok.encoder) == 512 + len(tok.bpe_ranks) + 2
    assert tok.encoder["!"] == 0 and tok.encoder["!</w>"] == 256
    assert tok.sot == len(tok.encoder) - 2 and tok.eot == len(tok.encoder) - 1
    # with the released table (48 894 merges) these are CLIP's 49406 / 49407
    assert 512 + 48894 == T.SOT and 512 + 48894 + 1 == T.EOT


def test_bpe_matches_transformers(tok):
    from transformers import CLIPTokenizer
    ref = CLIPTokenizer(vocab=dict(tok.encoder), merges=[tuple(m) for m in tok.bpe_ranks])
    for text in SAMPLES + synth.ade20k_labels():
        want = ref(text)["input_ids"]  # <sot> ids <eot>
        got = [tok.sot] + tok.encode(text) + [tok.eot]
        assert got == want, (text, got, want)


def test_tokenize_contract(tok):
    out = tok.tokenize(["wall", "traffic light", ""])
    assert out.shape == (3, 77) and out.dtype == torch.int64
    assert (out[:, 0] == tok.sot).all()
    assert out[2, 1] == tok.eot and out[2, 2:].sum() == 0
    eot_pos = out.argmax(dim=-1)  # EOT is the largest id: CLIP pools at text.argmax(-1)
    assert (out[torch.arange(3), eot_pos] == tok.eot).all()
    with pytest.raises(RuntimeError, match="too long for context length"):
        tok.tokenize(["z q " * 60])


def test_vocab_file_loader_and_resolution(tmp_path, monkeypatch):
    """from_file reads CLIP's file format (header line, one merge per line, gzip); tokenize() prefers it; without any
    vocabulary and without the explicit stand-in switch tokenize() refuses instead of hashing silently."""
    merges = _learn_merges(CORPUS, 120)
    path = tmp_path / T.VOCAB_FILE
    with gzip.open(path, "wb") as f:
        f.write(("#version: synthetic\n" + "\n".join(" ".join(m) for m in merges) + "\n").encode("utf-8"))
    t = T.BPETokenizer.from_file(str(path))
    assert list(t.bpe_ranks) == merges
    monkeypatch.setenv("LSEG_CLIP_BPE", str(path))
    T._default_tokenizer.cache_clear()
    try:
        assert T.find_vocab_file() == str(path)
        assert torch.equal(T.tokenize(["wall", "sky"]), t.tokenize(["wall", "sky"]))
    finally:
        monkeypatch.delenv("LSEG_CLIP_BPE")
        T._default_tokenizer.cache_clear()
    if T.find_vocab_file() is None:
        was = T._stand_in_enabled
        T.enable_stand_in(False)
        try:
            with pytest.raises(FileNotFoundError, match="CLIP BPE vocabulary"):
                T.tokenize(["wall"])
            T.enable_stand_in(True)
            with pytest.warns(RuntimeWarning) if not T._warned else _nullcontext():
                a = T.tokenize(["wall", "traffic light"])
            assert torch.equal(a, synth.tokenize(["wall", "traffic light"]))
        finally:
            T.enable_stand_in(was)


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
