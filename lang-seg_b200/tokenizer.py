"""Label tokenisation for the drop-in LSegNet — CLIP's byte-level BPE (reference: clip.tokenize, called at
modules/models/lseg_net.py:158,164 and lseg_net_zs.py:171-175; algorithm of CLIP@04f4dc2 `simple_tokenizer.py`, restated).

    tokenize(labels) -> int64 [K, 77]: <|startoftext|> (49406), BPE ids, <|endoftext|> (49407), zero padding;
                        RuntimeError when a prompt does not fit the context (as clip.tokenize does).

The algorithm lives here; the DATA it needs — the merges table `bpe_simple_vocab_16e6.txt.gz` that ships inside the
`clip` package — cannot be redistributed or downloaded in the build environment, so it is looked up at run time:
    $LSEG_CLIP_BPE, <this package>/bpe_simple_vocab_16e6.txt.gz, the installed `clip` package's directory.
Resolution order of `tokenize`:
    1. the vocabulary file is found         -> the native BPE below (ids identical to clip.tokenize);
    2. the real `clip` package is importable -> clip.tokenize (it carries the file);
    3. otherwise                             -> ERROR, unless the deterministic stand-in has been enabled explicitly
       (`enable_stand_in()` / LSEG_ALLOW_HASH_TOKENIZER=1): word -> 1000 + crc32(word) % 40000, same shape / SOT / EOT /
       padding contract. Tests, bench.py and smoke() enable it — there are no real weights offline either — but a
       deployment with a real CLIP checkpoint must never get hashed ids silently (they index unrelated embedding rows).
Callers may also pass pre-tokenised int64 [K,77] tensors straight to LSegNet.forward.
"""
import gzip
import html
import os
import warnings
import zlib
from functools import lru_cache

import torch

CONTEXT = 77
SOT, EOT = 49406, 49407
VOCAB_FILE = "bpe_simple_vocab_16e6.txt.gz"

_stand_in_enabled = bool(os.environ.get("LSEG_ALLOW_HASH_TOKENIZER"))
_warned = False


def enable_stand_in(on=True):
    """Allow the crc32 stand-in when no BPE vocabulary is available (synthetic-weight tests / benchmarks only)."""
    global _stand_in_enabled
    _stand_in_enabled = bool(on)


# ---------------------------------------------------------------------------------------------------------------------
# byte-level BPE
# ---------------------------------------------------------------------------------------------------------------------
@lru_cache()
def bytes_to_unicode():
    """Reversible byte -> printable unicode character table (the 188 printable latin-1 bytes map to themselves, the rest
    to code points from 256 upwards), so that BPE operates on strings without whitespace / control characters."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def _get_pairs(word):
    return {(a, b) for a, b in zip(word[:-1], word[1:])}


def _whitespace_clean(text):
    return " ".join(text.split())


def _basic_clean(text):
    try:  # CLIP runs ftfy.fix_text first; it only matters for mojibake input and is optional here
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    return html.unescape(html.unescape(text)).strip()


class BPETokenizer:
    """CLIP's SimpleTokenizer: `merges` is the ordered list of (left, right) merge rules (48 894 for the released
    vocabulary). Vocabulary ids: 256 byte symbols, the same 256 with the end-of-word marker, one id per merge, SOT, EOT."""

    def __init__(self, merges):
        import regex
        self.byte_encoder = bytes_to_unicode()
        vocab = list(self.byte_encoder.values())
        vocab = vocab + [v + "</w>" for v in vocab]
        merges = [tuple(m) for m in merges]
        vocab.extend("".join(m) for m in merges)
        vocab.extend(["<|startoftext|>", "<|endoftext|>"])
        self.encoder = dict(zip(vocab, range(len(vocab))))
        self.bpe_ranks = dict(zip(merges, range(len(merges))))
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}
        self.pat = regex.compile(
            r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
            regex.IGNORECASE)
        self.sot = self.encoder["<|startoftext|>"]
        self.eot = self.encoder["<|endoftext|>"]

    @classmethod
    def from_file(cls, path):
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rb") as f:
            lines = f.read().decode("utf-8").split("\n")
        lines = lines[1:49152 - 256 - 2 + 1]  # header line, then exactly the merges the released vocabulary uses
        return cls([tuple(line.split()) for line in lines if line.strip()])

    def bpe(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        pairs = _get_pairs(word)
        if not pairs:
            return token + "</w>"
        while True:
            bigram = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if bigram not in self.bpe_ranks:
                break
            first, second = bigram
            new_word = []
            i = 0
            while i < len(word):
                try:
                    j = word.index(first, i)
                except ValueError:
                    new_word.extend(word[i:])
                    break
                new_word.extend(word[i:j])
                i = j
                if word[i] == first and i < len(word) - 1 and word[i + 1] == second:
                    new_word.append(first + second)
                    i += 2
                else:
                    new_word.append(word[i])
                    i += 1
            word = tuple(new_word)
            if len(word) == 1:
                break
            pairs = _get_pairs(word)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text):
        ids = []
        text = _whitespace_clean(_basic_clean(text)).lower()
        for token in self.pat.findall(text):
            token = "".join(self.byte_encoder[b] for b in token.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(token).split(" "))
        return ids

    def tokenize(self, texts, context_length=CONTEXT):
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros((len(texts), context_length), dtype=torch.int64)
        for i, text in enumerate(texts):
            ids = [self.sot] + self.encode(text) + [self.eot]
            if len(ids) > context_length:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            out[i, : len(ids)] = torch.tensor(ids, dtype=torch.int64)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# vocabulary lookup + public entry point
# ---------------------------------------------------------------------------------------------------------------------
def find_vocab_file():
    cands = [os.environ.get("LSEG_CLIP_BPE"), os.path.join(os.path.dirname(os.path.abspath(__file__)), VOCAB_FILE)]
    try:
        import importlib.util
        spec = importlib.util.find_spec("clip")
        if spec is not None and spec.origin:
            cands.append(os.path.join(os.path.dirname(spec.origin), VOCAB_FILE))
    except (ImportError, ValueError):
        pass
    for c in cands:
        if c and os.path.isfile(c):
            return c
    return None


@lru_cache()
def _default_tokenizer():
    path = find_vocab_file()
    return BPETokenizer.from_file(path) if path else None


def _hash_tokenize(labels, context_length=CONTEXT):
    if isinstance(labels, str):
        labels = [labels]
    out = torch.zeros((len(labels), context_length), dtype=torch.int64)
    for i, text in enumerate(labels):
        ids = [SOT] + [1000 + (zlib.crc32(w.encode("utf-8")) % 40000) for w in text.lower().strip().split()] + [EOT]
        if len(ids) > context_length:
            raise RuntimeError(f"Input {text} is too long for context length {context_length}")
        out[i, : len(ids)] = torch.tensor(ids, dtype=torch.int64)
    return out


def tokenize(labels, context_length=CONTEXT):
    global _warned
    tok = _default_tokenizer()
    if tok is not None:
        return tok.tokenize(labels, context_length)
    try:
        import clip  # the reference's tokenizer, when the real package is installed (its own errors propagate)
    except ImportError:
        clip = None
    if clip is not None and hasattr(clip, "tokenize") and getattr(clip, "__file__", None):
        return clip.tokenize(labels).to(torch.int64)
    if not _stand_in_enabled:
        raise FileNotFoundError(
            f"CLIP BPE vocabulary '{VOCAB_FILE}' not found (looked at $LSEG_CLIP_BPE, the lseg_b200 package directory and "
            f"an installed `clip` package). Put the file there, or pass pre-tokenised int64 [K,77] tensors to forward(); "
            f"for synthetic-weight tests call lseg_b200.tokenizer.enable_stand_in() / set LSEG_ALLOW_HASH_TOKENIZER=1.")
    if not _warned:
        warnings.warn("lseg_b200: no CLIP BPE vocabulary available — using the crc32 stand-in tokenizer (ids are NOT "
                      "CLIP's; fine for synthetic weights only)", RuntimeWarning, stacklevel=2)
        _warned = True
    return _hash_tokenize(labels, context_length)
