"""Label tokenisation for the drop-in LSegNet (reference: clip.tokenize, lseg_net.py:158,164).

The CLIP BPE vocabulary (bpe_simple_vocab_16e6.txt.gz) is not available offline, so:
  * if the real `clip` package is importable, its tokenizer is used (identical ids to the reference);
  * otherwise a deterministic stand-in with the same contract (int64 [K,77], SOT 49406, EOT 49407,
    zero padding, error when a prompt exceeds the context) maps each lower-cased word to
    1000 + crc32(word) % 40000. The real BPE is a "next" row of SURVEY.md section 8(f).
Callers may also pass pre-tokenised int64 [K,77] tensors straight to LSegNet.forward.
"""
import zlib

import torch

CONTEXT = 77
SOT, EOT = 49406, 49407


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
    try:
        import clip  # the reference's tokenizer, when installed
        if hasattr(clip, "tokenize") and getattr(clip, "__file__", None):
            return clip.tokenize(labels).to(torch.int64)
    except Exception:
        pass
    return _hash_tokenize(labels, context_length)
