"""Per-kernel CUDA-event times of dpt_encode_words (pre-split words, the host-split fallback of both adapters) on the
words of a sentence-pair corpus (a word = a space + what follows it, the shape a byte-level pre-tokenizer gives).

    python tools/bench_words.py [n_bytes]                 (GPU box; DPT_LIB_PATH selects another build of the library)
"""
import json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dp-tokenization_b200")]

import numpy as np
import torch

from dptok import _cabi, assets, synth
from dptok import engine as eng_mod
from dptok.vocab import CompiledVocab


def main():
    n_bytes = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    v2i = {t: k for k, t in enumerate(assets.load_spec("gpt2_50k")["model"]["vocab"])}
    eng = eng_mod.Engine(CompiledVocab.from_token_map(v2i, "bytelevel"), 0)
    cache = os.path.join(ROOT, "tools", "_words_text.npy")  # development: a corpus generated beforehand (host generator, slow)
    text = np.load(cache) if os.path.isfile(cache) else synth.gen_sentence_pairs(n_bytes, seed=0)[0]
    cut = np.flatnonzero(text == 0x20)
    cut = cut[cut > 0]
    offs = np.concatenate([[0], cut, [len(text)]]).astype(np.int64)
    d_text, d_offs = torch.from_numpy(text).cuda(), torch.from_numpy(offs).cuda()
    for _ in range(3):
        res = eng.encode_words(d_text, d_offs)
    torch.cuda.synchronize()
    reps = 10
    eng_mod.profile_enable(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        res = eng.encode_words(d_text, d_offs)
    b.record()
    torch.cuda.synchronize()
    eng_mod.profile_enable(False)
    rows = eng_mod.profile_report()
    print(json.dumps({"lib": os.path.basename(_cabi.LIB_PATH), "bytes": len(text), "words": len(offs) - 1, "tokens": res.n_ids,
                      "ms_per_call_incl_host": a.elapsed_time(b) / reps,
                      "kernels_ms_per_call": {n: ms / reps for n, _, ms in rows},
                      "kernel_sum_ms": sum(ms for _, _, ms in rows) / reps}))


if __name__ == "__main__":
    main()
