"""BERT-uncased WordPiece tokenisation without the `transformers` package.

run_ldm_sampler.py:28-46 builds token ids with HF's `BertTokenizer.from_pretrained(vocab_dir)`
(`truncation=True, max_length=77, padding="max_length"`).  This module restates that published
algorithm (BERT `BasicTokenizer` + `WordpieceTokenizer`, uncased) over the same `vocab.txt`, so the
drop-in CLI needs neither HF nor a network:

  clean (drop control characters, normalise whitespace) -> space out CJK ideographs -> split on
  whitespace -> lower-case, NFD, strip combining marks -> split off punctuation -> greedy
  longest-match-first word pieces with the `##` continuation prefix ([UNK] for a word with an
  unmatched remainder or more than 100 characters) -> [CLS] ids [SEP], truncated to max_length
  and padded with [PAD].

tests/golden/wordpiece_small.json (made by tests/golden/make_wordpiece_golden.py with HF's
BertTokenizerFast on the reference's bert_model/vocab.txt) pins it, including the two known-answer
vectors of convert_ckpt_pytorch_to_tf2.py:384-392.
"""
import os
import unicodedata

import numpy as np


def load_vocab(vocab_dir_or_file):
    path = vocab_dir_or_file
    if os.path.isdir(path):
        path = os.path.join(path, "vocab.txt")
    vocab = {}
    with open(path, encoding="utf-8") as f:
        for i, line in enumerate(f):
            vocab[line.rstrip("\n")] = i
    return vocab


def _is_whitespace(ch):
    return ch in " \t\n\r" or unicodedata.category(ch) == "Zs"


def _is_control(ch):
    if ch in "\t\n\r":
        return False
    return unicodedata.category(ch).startswith("C")


def _is_punctuation(ch):
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp):
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or
            0x2A700 <= cp <= 0x2B73F or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or
            0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


def basic_tokenize(text):
    out = []
    for ch in text:
        cp = ord(ch)
        if cp == 0 or cp == 0xFFFD or _is_control(ch):
            continue
        if _is_whitespace(ch):
            out.append(" ")
        elif _is_cjk(cp):
            out.append(f" {ch} ")
        else:
            out.append(ch)
    words = []
    for tok in "".join(out).split():
        tok = unicodedata.normalize("NFD", tok.lower())
        tok = "".join(c for c in tok if unicodedata.category(c) != "Mn")
        cur = []
        for c in tok:
            if _is_punctuation(c):
                if cur:
                    words.append("".join(cur))
                    cur = []
                words.append(c)
            else:
                cur.append(c)
        if cur:
            words.append("".join(cur))
    return words


def wordpiece(word, vocab, unk="[UNK]", max_chars=100):
    if len(word) > max_chars:
        return [unk]
    pieces, start = [], 0
    while start < len(word):
        end, cur = len(word), None
        while start < end:
            sub = word[start:end]
            if start > 0:
                sub = "##" + sub
            if sub in vocab:
                cur = sub
                break
            end -= 1
        if cur is None:
            return [unk]
        pieces.append(cur)
        start = end
    return pieces


def encode(text, vocab, max_length=77):
    """ids of `[CLS] pieces [SEP]`, truncated to max_length and padded with [PAD] to max_length."""
    ids = []
    for w in basic_tokenize(text):
        ids.extend(vocab[p] for p in wordpiece(w, vocab))
    ids = ids[:max_length - 2]
    ids = [vocab["[CLS]"]] + ids + [vocab["[SEP]"]]
    return ids + [vocab["[PAD]"]] * (max_length - len(ids))


def get_token_ids(prompt, vocab_dir, batch_size, max_length=77):
    """run_ldm_sampler.py:28-46: B unconditional rows (empty prompt) then B conditional rows, int64."""
    vocab = load_vocab(vocab_dir)
    cond, uncond = encode(prompt, vocab, max_length), encode("", vocab, max_length)
    return np.array([uncond] * batch_size + [cond] * batch_size, dtype=np.int64)
