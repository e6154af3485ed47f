#!/usr/bin/env python
"""Generates tests/golden/wordpiece_small.json with HF's BertTokenizerFast on the reference's
bert_model/vocab.txt (build container only).  The fixture holds prompts, their padded id vectors and
the (token -> id) entries of every token that occurs -- enough for a greedy longest-match tokenizer to
reproduce the ids (a longer match absent from the reduced vocabulary is absent from the full one too)."""
import json
import os

from transformers import BertTokenizerFast

PROMPTS = [
    "a virus monster is playing guitar, oil on canvas", "",
    "A painting of a squirrel eating a burger", "Café déjà-vu!! naïve façade", "北京 and 東京 skyline at night",
    "unbelievablenesses xqzw antidisestablishmentarianism", "hello 🙂 world", "semi-colon; colon: dash—emdash … ellipsis",
    "  multiple   spaces\tand\nnewlines  ", "UPPER lower MiXeD 1234 5,678.90 $%&", "it's a dog's life, isn't it?",
    "an astronaut riding a horse in the style of picasso " * 12,
    "x" * 120 + " tail", "straße ÅNGSTRÖM ﬁnal", "photo-realistic 8k render, trending on artstation #art @user",
]


def main():
    tok = BertTokenizerFast.from_pretrained("/root/reference/bert_model")
    ids = [tok(p, truncation=True, max_length=77, padding="max_length")["input_ids"] for p in PROMPTS]
    used = sorted({i for row in ids for i in row})
    vocab = {tok.convert_ids_to_tokens(i): i for i in used}
    for t in ("[PAD]", "[UNK]", "[CLS]", "[SEP]"):
        vocab[t] = tok.convert_tokens_to_ids(t)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wordpiece_small.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"prompts": PROMPTS, "ids": ids, "vocab": vocab}, f, ensure_ascii=False)
    print(len(PROMPTS), "prompts,", len(vocab), "vocabulary entries ->", path)


if __name__ == "__main__":
    main()
