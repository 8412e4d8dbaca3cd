"""Development aid: run the corpus pipeline once on a bench workload and print the device-side work-list counters
(PipeCtl at the start of the range workspace): words per length class, odd / deferred / long words."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dp-tokenization_b200"))
import numpy as np, torch
from dptok import _cabi, assets, synth
from dptok._cabi import lib
from dptok.engine import Engine
from dptok.vocab import CompiledVocab

wl = sys.argv[1] if len(sys.argv) > 1 else "s2orc_llama2"
mb = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
n = int(mb * 1e6)
if wl == "s2orc_llama2":
    text, offs = synth.gen_documents(n, seed=0, lexicon=synth.make_lexicon(200_000, seed=0))
    t2i = assets.load_hf("llama2_32k").get_vocab(); fam = "spm"; rule = _cabi.RULE_SPM_LLAMA
elif wl == "pairs_gpt2":
    text, offs = synth.gen_sentence_pairs(n, seed=0)
    t2i = {t: k for k, t in enumerate(assets.load_spec("gpt2_50k")["model"]["vocab"])}; fam = "bytelevel"; rule = _cabi.RULE_GPT2
else:
    t_en, o_en = synth.gen_documents(int(0.6 * n), seed=0, words_per_doc=(1200, 9000))
    t_ar, o_ar = synth.gen_documents(int(0.4 * n), seed=1000, flavour="ar", words_per_doc=(800, 6000))
    text = np.concatenate([t_en, t_ar]); offs = np.concatenate([o_en, o_ar[1:] + o_en[-1]])
    t2i = {t: k for k, t in enumerate(assets.load_spec("llama3_128k")["model"]["vocab"])}; fam = "bytelevel"; rule = _cabi.RULE_LLAMA3
eng = Engine(CompiledVocab.from_token_map(t2i, fam), 0)
d_text = torch.from_numpy(text.copy()).cuda(); d_offs = torch.from_numpy(offs).cuda()
res = eng.encode_corpus(d_text, d_offs, rule)
res = eng.encode_corpus(d_text, d_offs, rule, ids_cap=res.n_ids + 1024, word_cap=res.n_words + 1024)
torch.cuda.synchronize()
word_cap = res.n_words + 1024
tb = lib.dpt_corpus_table_workspace(len(text), word_cap, 0)
tb_al = (tb + 255) // 256 * 256
ctl = eng._ws[tb_al:tb_al + 96].cpu().numpy().view(np.uint32)
print(wl, "bytes", len(text), "words", res.n_words, "tokens", res.n_ids)
print("ticket_a,c", ctl[0:2], "n_pending[0..3]", ctl[2:6], "n_odd", ctl[6], "n_long", ctl[7], "n_defer", ctl[8])
