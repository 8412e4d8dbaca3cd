"""B200 implementation behind the call surface of the reference's ``packages/tokenizer_utils.py``.

* ``dp_tokenize_llama(llama_tokenizer, pretokenize_option='llama') -> (dp_tokenize, decode_dp_tokenization)``
  (tokenizer_utils.py:52-96)
* ``dp_tokenize_bloom(bloom_tokenizer, HF_CACHE_DIR) -> (dp_tokenize, decode_dp_tokenization)``
  (tokenizer_utils.py:98-181)
* ``pretokenize_with_llama``, ``pretokenize_raw``, ``merge_tokens`` (tokenizer_utils.py:7-50)

Callers still pass a HuggingFace tokenizer plus strings and get ``List[int]`` back.  The returned
``dp_tokenize`` additionally carries ``.batch(list_of_str) -> List[List[int]]`` and ``.engine`` for
the throughput path.  Every id comes from the CUDA kernels; there is no CPU DP.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List

import numpy as np
import torch

from dptok import _cabi
from dptok.engine import Engine, pack_documents
from dptok.vocab import CompiledVocab, bytelevel_to_bytes

from .dp_tokenize import compute_shortest_tokenizations, obtain_longest_token

SPACE_TOKEN = "▁"


class _BiMap(dict):
    """Minimal stand-in for ``bidict`` (not installed here): dict with ``.inverse``."""

    @property
    def inverse(self):
        inv = self.__dict__.get("_inv")
        if inv is None or len(inv) != len(self):
            inv = {v: k for k, v in self.items()}
            self.__dict__["_inv"] = inv
        return inv


def merge_tokens(tokens, sep="Ġ"):
    """Glue token strings into words: a new word starts at every token that starts with ``sep``
    (tokenizer_utils.py:7-22)."""
    words: List[str] = []
    for k, t in enumerate(tokens):
        if k == 0 or t.startswith(sep):
            words.append(t)
        else:
            words[-1] += t
    return words


def pretokenize_with_llama(tokenizer, vocab_bidict):
    """Default-tokenizer-driven word split (tokenizer_utils.py:24-31).  Host side: this IS the
    tokenizer's own encode; the device rule DPT_RULE_SPM_LLAMA reproduces it for unambiguous text."""
    inverse = vocab_bidict.inverse

    def pretokenize(input_str):
        return merge_tokens([inverse[t] for t in tokenizer.encode(input_str)], sep=SPACE_TOKEN)

    return pretokenize


def pretokenize_raw(manual_mapping):
    """Whitespace split into unit lists (tokenizer_utils.py:33-50): unit 0 is U+2581+c0, every ' '
    becomes the unit U+2581 and starts a new word, chars in ``manual_mapping.inverse`` are renamed."""
    rename = manual_mapping.inverse

    def pretokenize(input_str):
        units = list(input_str)
        words, start = [], 0
        for i, ch in enumerate(units):
            if i == 0:
                units[i] = SPACE_TOKEN + ch
            elif ch == " ":
                units[i] = SPACE_TOKEN
                words.append(units[start:i])
                start = i
            elif ch in rename:
                units[i] = rename[ch]
        words.append(units[start:])
        return words

    return pretokenize


def _ids_per_doc(res, n_docs) -> List[List[int]]:
    ids = res.ids.cpu().numpy()
    offs = res.doc_tok_offs.cpu().numpy()
    return [ids[offs[d]:offs[d + 1]].tolist() for d in range(n_docs)]


def _raise_untokenizable():
    # the reference drops into ipdb (tokenizer_utils.py:72-73) and then fails in
    # obtain_longest_token([]) (dp_tokenize.py:82-84)
    raise ValueError("max() arg is an empty sequence")


def _spm_backend_matches_device_rule(tokenizer) -> bool:
    """True when the tokenizer's backend is what DPT_RULE_SPM_LLAMA implements: normaliser = Prepend(U+2581) then
    Replace(' ', U+2581) (or the equivalent Metaspace pre-tokenizer with prepend_scheme 'first' and no split), no other
    pre-tokenizer, post-processor = '<s>' in front of the single sequence and nothing else.  Tokenizers without a
    ``tokenizers`` backend (the slow SentencePiece classes) are judged by their flags and by the construction probe."""
    backend = getattr(tokenizer, "backend_tokenizer", None) or getattr(tokenizer, "_tokenizer", None)
    if getattr(tokenizer, "add_eos_token", False):
        return False
    if backend is None or not hasattr(backend, "to_str"):
        return bool(getattr(tokenizer, "add_bos_token", True)) and getattr(tokenizer, "bos_token", "<s>") == "<s>"
    try:
        spec = json.loads(backend.to_str())
    except Exception:
        return False
    norm, pre, post = spec.get("normalizer"), spec.get("pre_tokenizer"), spec.get("post_processor")

    def is_prepend(d):
        return d.get("type") == "Prepend" and d.get("prepend") == SPACE_TOKEN

    def is_replace(d):
        return d.get("type") == "Replace" and d.get("pattern") == {"String": " "} and d.get("content") == SPACE_TOKEN

    if norm is not None:
        steps = norm.get("normalizers") if norm.get("type") == "Sequence" else [norm]
        if not (len(steps) == 2 and is_prepend(steps[0]) and is_replace(steps[1])) or pre is not None:
            return False
    else:
        if not (pre and pre.get("type") == "Metaspace" and pre.get("replacement") == SPACE_TOKEN and
                pre.get("prepend_scheme") == "first" and pre.get("split") is False):
            return False
    if not post or post.get("type") != "TemplateProcessing":
        return False
    single = post.get("single") or []
    if len(single) != 2 or single[0].get("SpecialToken", {}).get("id") != "<s>" or "Sequence" not in single[1]:
        return False
    return True


def dp_tokenize_llama(llama_tokenizer, pretokenize_option="llama", cache_dir=None, device=None):
    t2i_dict = llama_tokenizer.get_vocab()
    vocab_bidict = _BiMap(t2i_dict)
    compiled = CompiledVocab.cached(t2i_dict, "spm", cache_dir)
    merges = CompiledVocab.merges_of(llama_tokenizer, t2i_dict)
    if merges:
        # the tokenizer's merge table travels with the vocabulary: runs of spaces / U+2581 are then cut on the device where
        # the default tokenizer cuts them (tokenizer_utils.py:7-31) instead of sending their documents to the host split
        compiled.set_merges(merges)
    engine = Engine(compiled, device)
    info = engine.vocab.info
    manual_mapping = _BiMap({"<0x0A>": "\n"})
    host_split = pretokenize_with_llama(llama_tokenizer, vocab_bidict)
    specials = set(getattr(llama_tokenizer, "all_special_tokens", []) or [])
    try:
        specials |= set(llama_tokenizer.get_added_vocab().keys())
    except Exception:
        pass
    specials = [s for s in specials if s]
    # The device rule DPT_RULE_SPM_LLAMA hard-codes ONE pipeline: Prepend(U+2581) + Replace(' ', U+2581), no
    # pre-tokenizer, the literal '<s>' in front of every document and nothing behind it.  Vocabulary facts alone do not say
    # that the tokenizer is built that way (add_bos_token=False, add_eos_token=True, another BOS string, no dummy prefix,
    # an extra normaliser...): the backend's own description must match, and a probe at construction compares the device
    # rule with the tokenizer-driven split on a few strings.  Otherwise every text takes the host split (DP on the GPU).
    device_rule_ok = bool(info.marker_leading_only and info.byte_fallback) and _spm_backend_matches_device_rule(llama_tokenizer)
    dev = engine.device

    def _encode_presplit(words: List[str]) -> List[int]:
        text, offs = pack_documents([w.encode("utf-8") for w in words])
        res = engine.encode_words(torch.from_numpy(text.copy()).to(dev), torch.from_numpy(offs).to(dev))
        if int(res.counters[_cabi.CTR_UNTOKENIZABLE]):
            _raise_untokenizable()
        return res.ids.cpu().tolist()

    def _needs_host_split(s: str) -> bool:
        return (not device_rule_ok) or s == "" or any(sp in s for sp in specials)

    def _encode_many(texts: List[str]) -> List[List[int]]:
        out: List = [None] * len(texts)
        if len(texts) == 1 and not _needs_host_split(texts[0]):
            # one string per call (main_analyze_s2orc.py:78): the single-document path - no allocation, one copy each way
            one = engine.encode_one(texts[0].encode("utf-8"), _cabi.RULE_SPM_LLAMA)
            if one is not None and not one[1]:
                if one[2][_cabi.CTR_UNTOKENIZABLE]:
                    _raise_untokenizable()
                return [one[0].tolist()]
        easy = [k for k, s in enumerate(texts) if not _needs_host_split(s)]
        if easy:
            text, offs = pack_documents([texts[k].encode("utf-8") for k in easy])
            res = engine.encode_corpus(torch.from_numpy(text.copy()).to(dev), torch.from_numpy(offs).to(dev),
                                       _cabi.RULE_SPM_LLAMA)
            per_doc = _ids_per_doc(res, len(easy))
            ambiguous = res.doc_flags.cpu().numpy() & _cabi.DF_AMBIGUOUS
            untok = int(res.counters[_cabi.CTR_UNTOKENIZABLE])
            for pos, k in enumerate(easy):
                if not ambiguous[pos]:
                    out[k] = per_doc[pos]
            if untok:
                # the counter is global: find out whether one of the documents the device rule decided holds the word
                # (the reference raises for that document; ambiguous ones are re-done below and raise there if need be)
                for pos, k in enumerate(easy):
                    if not ambiguous[pos]:
                        t1, o1 = pack_documents([texts[k].encode("utf-8")])
                        r1 = engine.encode_corpus(torch.from_numpy(t1.copy()).to(dev), torch.from_numpy(o1).to(dev),
                                                  _cabi.RULE_SPM_LLAMA)
                        if int(r1.counters[_cabi.CTR_UNTOKENIZABLE]):
                            _raise_untokenizable()
        hard = [k for k in range(len(texts)) if out[k] is None]
        if hard:
            # word split depends on the BPE merge order (or the text holds special tokens): ask the tokenizer for the
            # split exactly like pretokenize_with_llama, then ONE device pass over the words of all these texts
            per_doc = [host_split(texts[k]) for k in hard]
            flat = [w.encode("utf-8") for doc in per_doc for w in doc]
            if flat:
                text, offs = pack_documents(flat)
                res = engine.encode_words(torch.from_numpy(text.copy()).to(dev), torch.from_numpy(offs).to(dev),
                                          want_tok_offs=True)
                if int(res.counters[_cabi.CTR_UNTOKENIZABLE]):
                    _raise_untokenizable()
                ids = res.ids.cpu().numpy()
                to = res.word_tok_offs.cpu().numpy()
            w = 0
            for k, doc in zip(hard, per_doc):
                out[k] = ids[to[w]:to[w + len(doc)]].tolist() if doc else []
                w += len(doc)
        return out

    if device_rule_ok:
        probes = ["the weather", "Hello world, this is a test.", "a\nb c", "x", "caf\u00e9 \u4e2d\u6587 ok", "tab\there (1.5%) end."]
        try:
            device_rule_ok = False
            want = []
            for s_ in probes:
                words = host_split(s_)
                want.append(_encode_presplit(words) if words else [])
            device_rule_ok = True
            got = _encode_many(probes)
            device_rule_ok = got == want
        except Exception:
            device_rule_ok = False

    if pretokenize_option == "llama":
        def dp_tokenize(input_str) -> List[int]:
            return _encode_many([input_str])[0]
        dp_tokenize.batch = _encode_many
    elif pretokenize_option == "raw":
        raw_split = pretokenize_raw(manual_mapping)
        vocab = frozenset(t2i_dict)  # (immutable: dp_tokenize._engine_for recognises the object and compiles it once)

        def dp_tokenize(input_str) -> List[int]:
            ids: List[int] = []
            for word_units in raw_split(input_str):
                options, _ = compute_shortest_tokenizations(word_units, vocab, False, None, 1)
                for token in obtain_longest_token(options):
                    ids.append(t2i_dict[token])
            return ids
        dp_tokenize.batch = lambda texts: [dp_tokenize(s) for s in texts]
    else:
        # the reference leaves pretokenize_func unbound for any other value (tokenizer_utils.py:62-65)
        def dp_tokenize(input_str):
            raise UnboundLocalError("cannot access local variable 'pretokenize_func'")
    dp_tokenize.engine = engine

    def decode_dp_tokenization(encoding: List[int]):
        return llama_tokenizer.decode(encoding)[4:]  # strips "<s> " (tokenizer_utils.py:84)

    return dp_tokenize, decode_dp_tokenization


_BLOOM_SNAPSHOT = "models--bigscience--bloom-3b/snapshots/52bc5b43010b4844513826b8be3f78c7344c37d7/tokenizer.json"


def _bytelevel_vocab(tokenizer, HF_CACHE_DIR) -> Dict[str, int]:
    """token -> ENUMERATION index of ``model.vocab`` (tokenizer_utils.py:105-113)."""
    path = f"{HF_CACHE_DIR}/{_BLOOM_SNAPSHOT}" if HF_CACHE_DIR else None
    if path and os.path.isfile(path):
        with open(path, "r") as f:
            spec = json.load(f)
    else:
        backend = getattr(tokenizer, "backend_tokenizer", None) or tokenizer._tokenizer
        spec = json.loads(backend.to_str())
    return {token: index for index, token in enumerate(spec["model"]["vocab"])}


_LLAMA3_SPLIT = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*"
                 r"|\s*[\r\n]+|\s+(?!\S)|\s+")


_BLOOM_SPLIT = r" ?[^(\s|[.,!?…。，、।۔،])]+"


def _device_split_rule(tokenizer):
    """DPT_RULE_* of the tokenizer's pre-tokenizer if the device implements exactly that split (GPT-2 ByteLevel regex,
    Llama-3 or BLOOM Split regex + ByteLevel) and nothing rewrites the text before it; None -> pre_tokenize_str on the
    host."""
    try:
        backend = getattr(tokenizer, "backend_tokenizer", None) or tokenizer._tokenizer
        spec = json.loads(backend.to_str())
    except Exception:
        return None
    if spec.get("normalizer") is not None:
        return None
    pt = spec.get("pre_tokenizer") or {}

    def is_bytelevel(d, use_regex):
        return (d.get("type") == "ByteLevel" and not d.get("add_prefix_space", True) and
                bool(d.get("use_regex", True)) == use_regex)

    if is_bytelevel(pt, True):
        return _cabi.RULE_GPT2
    if pt.get("type") == "Sequence":
        steps = pt.get("pretokenizers", [])
        if (len(steps) == 2 and steps[0].get("type") == "Split" and steps[0].get("behavior") == "Isolated" and
                not steps[0].get("invert", False) and is_bytelevel(steps[1], False)):
            rx = steps[0].get("pattern", {}).get("Regex")
            if rx == _LLAMA3_SPLIT:
                return _cabi.RULE_LLAMA3
            if rx == _BLOOM_SPLIT:
                return _cabi.RULE_BLOOM
    return None


def dp_tokenize_bloom(bloom_tokenizer, HF_CACHE_DIR, cache_dir=None, device=None):
    vocab_to_index = _BiMap(_bytelevel_vocab(bloom_tokenizer, HF_CACHE_DIR))
    engine = Engine(CompiledVocab.cached(vocab_to_index, "bytelevel", cache_dir), device)
    dev = engine.device
    single = np.zeros(256, dtype=bool)
    for b in range(256):
        single[b] = engine.vocab.lookup(bytes([b])) >= 0

    def pretokenize(input_str):
        pieces = bloom_tokenizer._tokenizer.pre_tokenizer.pre_tokenize_str(input_str)
        return [p[0] for p in pieces]

    def _encode_pieces(pieces: List[str]) -> List[int]:
        if not pieces:
            return []
        raw = []
        for piece in pieces:
            try:
                data = bytelevel_to_bytes(piece)
            except KeyError as e:  # char outside the byte alphabet: vocab_to_index[c] fails (tokenizer_utils.py:149)
                raise KeyError(e.args[0])
            arr = np.frombuffer(data, dtype=np.uint8)
            if not single[arr].all():
                missing = int(arr[~single[arr]][0])
                raise KeyError(piece[[int(x) for x in arr].index(missing)])
            raw.append(data)
        text, offs = pack_documents(raw)
        res = engine.encode_words(torch.from_numpy(text.copy()).to(dev), torch.from_numpy(offs).to(dev))
        return res.ids.cpu().tolist()

    device_rule = _device_split_rule(bloom_tokenizer) if bool(single.all()) else None

    def dp_tokenize(input_str) -> List[int]:
        if device_rule is not None and input_str:
            # the tokenizer's split regex on the device, single-document path (every byte is a token here: nothing the
            # reference would raise for)
            one = engine.encode_one(input_str.encode("utf-8"), device_rule)
            if one is not None:
                return one[0].tolist()
        return _encode_pieces(pretokenize(input_str))

    def _batch(texts: List[str]) -> List[List[int]]:
        if device_rule is not None and all(texts):
            # the tokenizer's split regex runs on the GPU (kernel A of the corpus pipeline): no host pre-tokenization
            text, offs = pack_documents([s.encode("utf-8") for s in texts])
            res = engine.encode_corpus(torch.from_numpy(text.copy()).to(dev), torch.from_numpy(offs).to(dev), device_rule)
            return _ids_per_doc(res, len(texts))
        per_doc = [pretokenize(s) for s in texts]
        flat = [p for doc in per_doc for p in doc]
        if not flat:
            return [[] for _ in texts]
        if not bool(single.all()):  # some byte is no token: per document, with the reference's KeyError (tokenizer_utils.py:149)
            return [_encode_pieces(doc) for doc in per_doc]
        raw = [bytelevel_to_bytes(p) for p in flat]
        text, offs = pack_documents(raw)
        res = engine.encode_words(torch.from_numpy(text.copy()).to(dev), torch.from_numpy(offs).to(dev),
                                  want_tok_offs=True)
        ids = res.ids.cpu().numpy()
        to = res.word_tok_offs.cpu().numpy()
        out, w = [], 0
        for doc in per_doc:
            out.append(ids[to[w]:to[w + len(doc)]].tolist())
            w += len(doc)
        return out

    dp_tokenize.batch = _batch
    dp_tokenize.engine = engine
    dp_tokenize.device_rule = device_rule

    def decode_dp_tokenization(encoding: List[int]):
        return bloom_tokenizer.decode(encoding)

    return dp_tokenize, decode_dp_tokenization
