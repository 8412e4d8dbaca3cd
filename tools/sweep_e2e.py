"""Sweep chunk size / stream count of Engine.encode_corpus_host on the bench workload (development aid)."""
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
d_text = torch.empty(len(text), dtype=torch.uint8, device=0)
# raw PCIe numbers for context
for name, fn in (("H2D 100MB", lambda: d_text.copy_(h_text, non_blocking=True)),
                 ("D2H 95MB", lambda: h_ids[:23_700_000].copy_(torch.empty(23_700_000, dtype=torch.int32, device=0), non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, (time.perf_counter() - t0) / 5 * 1e3, "ms")
ovs = [bool(int(x)) for x in os.environ.get("SWEEP_OVERLAP", "0,1").split(",")]
for chunk_mb in [int(x) for x in os.environ.get("SWEEP_CHUNKS", "8,12,16,25,34,50").split(",")]:
    for ns in [int(x) for x in os.environ.get("SWEEP_SLOTS", "3,8").split(",")]:
        for ov in ovs:
            kw = dict(chunk_bytes=chunk_mb << 20, n_streams=ns, out_ids=h_ids, overlap=ov)
            r = eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_SPM_LLAMA, **kw)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5):
                r = eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_SPM_LLAMA, **kw)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
            print(f"chunk {chunk_mb:4d} MB slots {ns} overlap {int(ov)}: {dt*1e3:7.3f} ms  {len(text)/dt/1e9:6.2f} GB/s  chunks {r.n_chunks}", flush=True)
