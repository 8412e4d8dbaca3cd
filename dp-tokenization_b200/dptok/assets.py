"""Offline-trained stand-in tokenizers with the shapes BASELINE.json names.

Real Llama-2 / GPT-2 / Llama-3 vocab files are not on disk and there is no
network, so vocabularies of the same size, alphabet and
pre-tokenizer rule are trained with the installed ``tokenizers`` on the
synthetic corpora of ``synth.py``.  The trained ``tokenizer.json`` files are
committed (gzip) under ``assets/`` so tests and the bench are reproducible and
need no training at run time; ``python -m dptok.assets`` regenerates them.

Families:
  llama2  - BPE, byte_fallback, normaliser Prepend('▁')+Replace(' ','▁'), no pre-tokenizer,
            specials <unk>,<s>,</s> = 0,1,2 and <0x00>..<0xFF> = 3..258 (Llama-2 layout).
  gpt2    - byte-level BPE, ByteLevel(add_prefix_space=False, use_regex=True).
  llama3  - byte-level BPE, Split(llama-3 regex, isolated) + ByteLevel(use_regex=False).
  bloom   - byte-level BPE, Split(bloom regex, isolated) + ByteLevel(use_regex=False).
"""
from __future__ import annotations

import gzip
import json
import os
import sys

ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "assets")

LLAMA3_REGEX = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*"
                r"|\s*[\r\n]+|\s+(?!\S)|\s+")
BLOOM_REGEX = r" ?[^(\s|[.,!?…。，、।۔،])]+"

SPECS = {
    # name: (family, vocab size)
    "llama2_32k": ("llama2", 32000),
    "gpt2_50k": ("gpt2", 50257),
    "llama3_128k": ("llama3", 128256),
    "bloom_8k": ("bloom", 8192),
    "llama2_2k": ("llama2", 2048),
    "gpt2_3k": ("gpt2", 3000),
}


def _training_docs(family: str, n_bytes: int):
    from . import synth
    docs = synth.sample_text(n_bytes, seed=1234, flavour="en")
    if family != "llama2":
        docs += synth.sample_text(n_bytes // 3, seed=1235, flavour="de")
        docs += synth.sample_text(n_bytes // 3, seed=1236, flavour="ar")
    # whitespace runs, newlines and digits so those tokens exist as in real vocabs
    extra = ["".join(chr(c) for c in range(33, 127)) + " " + " ".join(chr(c) for c in range(33, 127))]
    for k in range(400):
        if family == "llama2":
            # Llama-2 has no '\n' / '\t' piece (they byte-fall-back to <0x0A>/<0x09>), but has '▁▁'-run pieces
            extra.append(" " * (k % 15 + 1) + "x %d %d,%d" % (k, k * 37, k * 1001))
        else:
            extra.append("  " * (k % 9 + 1) + "x\n\n" + "\t" * (k % 3) + "%d %d,%d" % (k, k * 37, k * 1001))
    return docs + extra * 5


def build(name: str, train_bytes: int | None = None):
    from tokenizers import Regex, Tokenizer, decoders, models, normalizers, pre_tokenizers, trainers
    family, vocab_size = SPECS[name]
    if train_bytes is None:
        train_bytes = 6_000_000 if vocab_size < 10000 else 40_000_000
    docs = _training_docs(family, train_bytes)
    if family == "llama2":
        tok = Tokenizer(models.BPE(unk_token="<unk>", byte_fallback=True, fuse_unk=True))
        tok.normalizer = normalizers.Sequence([normalizers.Prepend("▁"), normalizers.Replace(" ", "▁")])
        # training only: never merge across a '▁' (SentencePiece split_by_whitespace)
        tok.pre_tokenizer = pre_tokenizers.Split(Regex("▁[^▁]*|[^▁]+"), behavior="isolated")
        specials = ["<unk>", "<s>", "</s>"] + ["<0x%02X>" % b for b in range(256)]
        trainer = trainers.BpeTrainer(vocab_size=vocab_size, special_tokens=specials, show_progress=False,
                                      max_token_length=16)
        tok.train_from_iterator(docs, trainer)
        spec = json.loads(tok.to_str())
        spec["pre_tokenizer"] = None
        spec["added_tokens"] = [t for t in spec["added_tokens"] if t["content"] in ("<unk>", "<s>", "</s>")]
        spec["decoder"] = {"type": "Sequence", "decoders": [
            {"type": "Replace", "pattern": {"String": "▁"}, "content": " "},
            {"type": "ByteFallback"}, {"type": "Fuse"}, {"type": "Strip", "content": " ", "start": 1, "stop": 0}]}
        spec["post_processor"] = {
            "type": "TemplateProcessing",
            "single": [{"SpecialToken": {"id": "<s>", "type_id": 0}}, {"Sequence": {"id": "A", "type_id": 0}}],
            "pair": [{"SpecialToken": {"id": "<s>", "type_id": 0}}, {"Sequence": {"id": "A", "type_id": 0}},
                     {"SpecialToken": {"id": "<s>", "type_id": 1}}, {"Sequence": {"id": "B", "type_id": 1}}],
            "special_tokens": {"<s>": {"id": "<s>", "ids": [1], "tokens": ["<s>"]}}}
        return spec
    tok = Tokenizer(models.BPE())
    if family == "gpt2":
        tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True)
    else:
        rx = LLAMA3_REGEX if family == "llama3" else BLOOM_REGEX
        tok.pre_tokenizer = pre_tokenizers.Sequence([
            pre_tokenizers.Split(Regex(rx), behavior="isolated"),
            pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=vocab_size, special_tokens=[], show_progress=False,
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    tok.train_from_iterator(docs, trainer)
    return json.loads(tok.to_str())


def path(name: str) -> str:
    return os.path.join(ASSET_DIR, name + ".tokenizer.json.gz")


def load_spec(name: str) -> dict:
    with gzip.open(path(name), "rt", encoding="utf-8") as f:
        return json.load(f)


def load_tokenizer(name: str):
    """``tokenizers.Tokenizer`` for a committed asset."""
    from tokenizers import Tokenizer
    return Tokenizer.from_str(json.dumps(load_spec(name)))


def load_hf(name: str):
    """``transformers.PreTrainedTokenizerFast`` wrapper: the object type the
    reference's adapters receive (tokenizer_utils.py:52,98)."""
    from transformers import PreTrainedTokenizerFast
    family = SPECS[name][0]
    kw = dict(bos_token="<s>", eos_token="</s>", unk_token="<unk>") if family == "llama2" else {}
    return PreTrainedTokenizerFast(tokenizer_object=load_tokenizer(name), **kw)


def main(argv):
    os.makedirs(ASSET_DIR, exist_ok=True)
    names = argv or list(SPECS)
    for name in names:
        spec = build(name)
        with gzip.GzipFile(path(name), "wb", mtime=0) as f:
            f.write(json.dumps(spec, ensure_ascii=False, separators=(",", ":")).encode("utf-8"))
        print(name, len(spec["model"]["vocab"]), os.path.getsize(path(name)))


if __name__ == "__main__":
    main(sys.argv[1:])
