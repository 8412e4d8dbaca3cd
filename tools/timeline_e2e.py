"""Print the per-chunk timeline of Engine.encode_corpus_host (development aid): CUDA events on the three streams."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dp-tokenization_b200"))
import numpy as np, torch
from dptok import _cabi, assets, synth
from dptok.engine import Engine
from dptok.vocab import CompiledVocab

torch.cuda.set_device(0)
text, doc_offs = synth.gen_documents(100_000_000, seed=0, lexicon=synth.make_lexicon(200_000, seed=0))
eng = Engine(CompiledVocab.from_token_map(assets.load_hf("llama2_32k").get_vocab(), "spm"), 0)
h_text = torch.from_numpy(text.copy()).pin_memory()
h_ids = torch.empty(len(text) // 2 + 200000, dtype=torch.int32).pin_memory()
chunk = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 25 << 20
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ov = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
for _ in range(3):
    eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_SPM_LLAMA, chunk_bytes=chunk, n_streams=ns, out_ids=h_ids, overlap=ov)
torch.cuda.synchronize()
eng._trace = []
t0 = time.perf_counter()
r = eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_SPM_LLAMA, chunk_bytes=chunk, n_streams=ns, out_ids=h_ids, overlap=ov)
torch.cuda.synchronize()
print("wall", (time.perf_counter() - t0) * 1e3, "ms")
start = eng._trace[0][2]
for name, k, ev, host_t in sorted(eng._trace, key=lambda t: start.elapsed_time(t[2])):
    print(f"{name:10s} chunk {k}: gpu {start.elapsed_time(ev):7.3f} ms   host-enqueue {1e3*(host_t - eng._trace[0][3]):7.3f} ms")

# per-kernel device time summed over the ranges of one call (CUDA events around every launch)
from dptok import engine as eng_mod
eng._trace = None
eng_mod.profile_enable(True)
r = eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_SPM_LLAMA, chunk_bytes=chunk, n_streams=ns, out_ids=h_ids, overlap=ov)
torch.cuda.synchronize()
eng_mod.profile_enable(False)
tot = 0.0
for name, cnt, ms in eng_mod.profile_report():
    print(f"{name:22s} launches {cnt:3d}  total {ms:7.3f} ms  per launch {ms/cnt:7.3f}")
    tot += ms
print("sum of kernel times", tot)
