#!/usr/bin/env python
"""Benchmark of the shortest-tokenization hot path (BASELINE.json metric: corpus bytes/s & tokens/s per GPU,
% of HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A step = one pass of the whole path (boundary/normalise -> DP -> compaction -> counters [+ NCCL count
reduction]) over one batch of synthetic S2ORC-shaped text.  Workload at N=1 = BASELINE.json configs[1]:
Llama-2-shaped 32k SentencePiece-BPE vocab, 100 MB of abstracts.  Multi-GPU = document-sharded, each rank its own
100 MB shard (weak scaling), counters all-reduced once per step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dp-tokenization_b200"))

WORKLOADS = {
    # name: (asset, vocabulary family, rule, description)
    "s2orc_llama2": ("llama2_32k", "spm", "RULE_SPM_LLAMA",
                     "configs[1]: Llama-2-shaped 32k SentencePiece-BPE vocab, 100 MB synthetic S2ORC-shaped abstracts per GPU"),
    "pairs_gpt2": ("gpt2_50k", "bytelevel", "RULE_GPT2",
                   "configs[2]: GPT-2-shaped 50k byte-level BPE vocab, en/de biomedical-translation-shaped sentence pairs"),
    "longdocs_llama3": ("llama3_128k", "bytelevel", "RULE_LLAMA3",
                        "configs[3]-shaped: Llama-3-shaped 128k byte-level vocab, 8-64 KB documents (60 % English, 40 % "
                        "Arabic script with diacritics)"),
}
METRIC = "corpus_bytes_per_sec"
UNIT = "bytes/s"
ASSET = "llama2_32k"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (every ~1 ms; the timed
    region is a few milliseconds, too short for `nvidia-smi -lms`), falling back to one nvidia-smi query."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = False
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _poll(self):
        nv = self._nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                if get_reasons:
                    r = int(get_reasons(self._h))
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self._nvml is not None:
            import threading
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._thread is not None:
            self._stop = True
            self._thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=statistics.median(self.samples), reasons=sorted(self.reasons), samples=len(self.samples),
                           source="nvml, 1 ms polling during the timed region")
                return out
        try:  # fallback: one query right after the region
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            f = subprocess.check_output(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                         str(self.gpu)], timeout=10).decode().strip().split(",")
            f = [x.strip() for x in f]
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6])
                       if v.lower().startswith("active")]
            out.update(sm_mhz=float(f[0]), sm_max_mhz=float(f[1]), reasons=reasons, samples=1,
                       source="nvidia-smi, one sample right after the timed region")
        except Exception:
            pass
        return out


def cpu_reference_leg(n_sample_bytes: int, budget_s: float):
    """The reference's CPU path (oracle port, kind 'port') on all host cores, bounded sample."""
    from dptok import synth
    from oracle import cpu_baseline
    docs = synth.sample_text(n_sample_bytes, seed=0)
    r = cpu_baseline.run(ASSET, docs, budget_s=budget_s)
    return r


def run_reference_arm(args):
    """bench.py --impl reference: the reference's CPU path (oracle port) on all host cores; every step is a bounded
    sample of the configs[1] workload, sized so that the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from dptok import synth
    from oracle import cpu_baseline
    n_steps = args.warmup + args.steps
    budget = max(0.5, min(8.0, 150.0 / max(n_steps, 1)))
    docs = synth.sample_text(int(1_500_000 * budget) + 200_000, seed=0)
    runner = cpu_baseline.Runner(ASSET)
    runner.warm(docs)
    total = {"bytes": 0, "tokens": 0, "seconds": 0.0}
    per = max(1, len(docs) // max(n_steps, 1))
    for k in range(n_steps):
        # a different rotation of the sample every step; Runner.run hands out bounded batches and waits for each
        r = runner.run(docs[(k * per) % len(docs):] + docs[:(k * per) % len(docs)], budget_s=budget)
        if k >= args.warmup:
            for key in total:
                total[key] += r[key]
    runner.close()
    cores = runner.procs
    v = total["bytes"] / total["seconds"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total["seconds"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "tokens_per_sec": total["tokens"] / total["seconds"],
        "config": {"workload": "configs[1]: Llama-2-shaped 32k SentencePiece-BPE vocab, synthetic S2ORC-shaped abstracts",
                   "vocab": ASSET, "note": "reference CPU path = oracle port of packages/dp_tokenize.py + "
                   "tokenizer_utils.dp_tokenize_llama (pure-Python reference cannot travel to the GPU box)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{total['bytes']} bytes of the same synthetic corpus (seed 0) over {args.steps} steps of "
                                   f"{budget:.1f} s each, multiprocessing over documents"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mb", type=float, default=100.0, help="corpus bytes per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="s2orc_llama2", choices=sorted(WORKLOADS),
                    help="default = BASELINE.json configs[1]; the others are extra measurements, not the contract line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from dptok import _cabi, assets, engine as eng_mod, synth
    from dptok.engine import Engine
    from dptok.sharded import reduce_counters
    from dptok.vocab import CompiledVocab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: dptok has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.cuda.current_device()
    if world > 1:
        # NCCL prints its version banner to STDOUT when the communicator is created; the contract is ONE JSON line on
        # stdout, so create the communicator (init + one collective) with fd 1 pointing at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    # ---- workload: this rank's shard (documents are independent; each rank generates its own) ----
    n_bytes_target = int(args.size_mb * 1e6)
    asset, family, rule_name, workload_desc = WORKLOADS[args.workload]
    RULE = getattr(_cabi, rule_name)
    if args.workload == "s2orc_llama2":
        lexicon = synth.make_lexicon(200_000, seed=0)
        text, doc_offs = synth.gen_documents(n_bytes_target, seed=rank, lexicon=lexicon)
    elif args.workload == "pairs_gpt2":
        text, doc_offs = synth.gen_sentence_pairs(n_bytes_target, seed=rank)
    else:
        t_en, o_en = synth.gen_documents(int(0.6 * n_bytes_target), seed=rank, words_per_doc=(1200, 9000))
        t_ar, o_ar = synth.gen_documents(int(0.4 * n_bytes_target), seed=rank + 1000, flavour="ar", words_per_doc=(800, 6000))
        text = np.concatenate([t_en, t_ar])
        doc_offs = np.concatenate([o_en, o_ar[1:] + o_en[-1]])
    n_bytes, n_docs = len(text), len(doc_offs) - 1
    if family == "spm":
        t2i = assets.load_hf(asset).get_vocab()
    else:
        t2i = {t: k for k, t in enumerate(assets.load_spec(asset)["model"]["vocab"])}
    engine = Engine(CompiledVocab.from_token_map(t2i, family), dev)
    h_text = torch.from_numpy(text).pin_memory()
    h_offs = torch.from_numpy(doc_offs).pin_memory()
    d_text = h_text.to(dev, non_blocking=True)
    d_offs = h_offs.to(dev, non_blocking=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        res = engine.encode_corpus(d_text, d_offs, RULE)
        stats = reduce_counters(res.counters, None) if world > 1 else None
        return res, stats

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res, _ = step()
    n_tokens, n_words = res.n_ids, res.n_words
    ids_cap, word_cap = res.n_ids + 1024, res.n_words + 1024

    # ---- timed region: K steps, device time by CUDA events, L2 flushed between steps --------------
    launches0 = eng_mod.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        res = engine.encode_corpus(d_text, d_offs, RULE, ids_cap=ids_cap, word_cap=word_cap)
        if world > 1:
            reduce_counters(res.counters, None)
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = eng_mod.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    # ---- the same K steps once more with a CUDA event pair around every kernel launch (on the launching stream):
    #      per-kernel times for the roofline object.  Kept out of the timed region above: 2 x launches event records
    #      per step are not part of the path.
    eng_mod.profile_enable(True)
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        engine.encode_corpus(d_text, d_offs, RULE, ids_cap=ids_cap, word_cap=word_cap)
    barrier()
    eng_mod.profile_enable(False)
    prof = eng_mod.profile_report()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([n_bytes, n_tokens, n_words], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    g_bytes, g_tokens, g_words = [int(x) for x in tot.tolist()]
    value = g_bytes * args.steps / (total_ms / 1e3)

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    peak, peak_src = measured_peaks()
    # Algorithmic bytes: the whole pass moves AB = text read once + ids + per-word lengths
    # written; each kernel of the pipeline is charged the part of it (plus the 4-byte word refs between the kernels)
    # that it actually reads or writes, so the dominant kernel's `achieved` is its own bytes over its own time.
    ab = n_bytes + 4 * n_tokens + 4 * n_words
    kernel_ab = {
        "k_scan_dedup": n_bytes + 4 * n_words,                    # text read, refs written
        "k_scan_dedup_bl": n_bytes + 4 * n_words,
        "k_emit": 4 * n_words + 4 * n_tokens + 5 * n_words,       # refs read; ids, lengths, flags written
    }
    roofline = None
    if prof:
        name, cnt, ms = prof[0]
        per_launch_ms = ms / cnt
        launches_per_step = cnt / args.steps
        k_ab = kernel_ab.get(name, ab)
        achieved = k_ab / launches_per_step / (per_launch_ms / 1e3) / 1e9
        whole = ab * args.steps / (total_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                    "kernel_algorithmic_bytes_per_launch": k_ab / launches_per_step,
                    "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": ms / sum(r[2] for r in prof),
                    "algorithmic_bytes_per_step": ab,
                    "whole_path_achieved_gbs": whole, "whole_path_frac": whole / peak,
                    "kernels_ms_per_step": {r[0]: r[2] / args.steps for r in prof}}
        tr = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.isfile(tr):
            try:
                roofline["traffic"] = json.load(open(tr)).get(name)
            except Exception:
                pass

    # ---- end to end through the public API with HOST buffers -------------------------------------
    # Engine.encode_corpus_host: pinned host text -> chunked H2D / kernels / D2H on 4 streams -> pinned host ids,
    # document offsets and counters.  Both copies are inside the timed region.
    e2e = None
    if not args.no_e2e:
        h_ids = torch.empty(ids_cap, dtype=torch.int32).pin_memory()

        def e2e_step():
            r = engine.encode_corpus_host(h_text, doc_offs, RULE, out_ids=h_ids)
            if world > 1:
                c = torch.from_numpy(r.counters).to(dev)
                reduce_counters(c, None)
            return r
        r = e2e_step()
        assert r.n_ids == n_tokens and int(r.counters[2]) == n_tokens, "end-to-end path disagrees with the resident path"
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": g_bytes * args.steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(n_bytes + 8 * (n_docs + r.n_chunks)),
               "d2h_bytes_per_step": int(4 * r.n_ids + 9 * n_docs + 104 * r.n_chunks),
               "tokens_per_sec": g_tokens * args.steps / float(tt.item()),
               "ms_per_step": 1e3 * float(tt.item()) / args.steps, "chunks": r.n_chunks,
               "how": "Engine.encode_corpus_host: 16 MB ranges at document boundaries, copy-in / compute / copy-out streams, one word table for the whole corpus, wall clock"}

        # the same call with compact ids (uint16; this vocabulary has at most 65,536 entries): the host path is bound by
        # PCIe bytes, so halving the bytes that leave the GPU is what moves it.  Extra information, not the `e2e` line.
        if max(t2i.values()) <= 0xFFFF:
            h_ids16 = torch.empty(ids_cap, dtype=torch.uint16).pin_memory()
            r16 = engine.encode_corpus_host(h_text, doc_offs, RULE, out_ids=h_ids16, ids_dtype=torch.uint16)
            assert r16.n_ids == n_tokens
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                r16 = engine.encode_corpus_host(h_text, doc_offs, RULE, out_ids=h_ids16, ids_dtype=torch.uint16)
            barrier()
            dt16 = time.perf_counter() - t0
            tt16 = torch.tensor([dt16], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt16, op=dist.ReduceOp.MAX)
            e2e["compact_u16_ids"] = {"value": g_bytes * args.steps / float(tt16.item()), "unit": UNIT,
                                      "ms_per_step": 1e3 * float(tt16.item()) / args.steps,
                                      "d2h_bytes_per_step": int(2 * r16.n_ids + 9 * n_docs + 104 * r16.n_chunks),
                                      "how": "same call with ids_dtype=torch.uint16 (dpt_narrow_ids_u16 on the device)"}

    # ---- CPU baseline beside it (rank 0, N=1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "s2orc_llama2":
        r = cpu_reference_leg(16_000_000, budget_s=15.0)
        cpu = {"value": r["bytes_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "tokens_per_sec": r["tokens_per_s"],
               "sample": f"{r['bytes']} bytes / {r['docs']} documents of the same synthetic corpus generator (seed 0), "
                         f"{r['seconds']:.1f} s, reference algorithm (enumerating DP) under multiprocessing"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "tokens_per_sec": g_tokens * args.steps / (total_ms / 1e3),
            "bytes_per_token": g_bytes / max(g_tokens, 1),
            "config": {"workload": workload_desc, "vocab": asset, "bytes_per_gpu": n_bytes, "docs_per_gpu": n_docs,
                       "words_per_gpu": n_words, "tokens_per_gpu": n_tokens, "parallelism": f"doc-sharded x{world}",
                       "l2": "256 MiB buffer written between timed steps (L2 flush)", "timing": "CUDA events per step"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "wall_s_timed_region": t_wall, "step_ms": step_ms,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
