#!/usr/bin/env python
"""Benchmark of the shortest-tokenization hot path (BASELINE.json metric: corpus bytes/s & tokens/s per GPU,
% of HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A step = one pass of the whole path (boundary rule -> dedup -> DP -> compaction -> counters) over one batch of
synthetic text resident in HBM.

  N = 1   `value` = BASELINE.json configs[1]: Llama-2-shaped 32k SentencePiece-BPE vocabulary, 100 MB of S2ORC-shaped
          abstracts.  `other_workloads` carries the other configurations, each with its own ms_per_step, roofline and
          cpu_baseline: configs[2] (GPT-2-shaped 50k byte-level vocabulary, 100 MB of en/de sentence pairs), configs[3]
          (Llama-3-shaped 128k vocabulary, 1 GB of 8-64 KB documents generated on the device), one GPU's share of
          configs[4] (1.25 GB of Arabic-script text generated on the device), and configs[1] with the share of DISTINCT
          words raised (the pipeline solves each distinct word once: its speed depends on the redundancy of the text).
          `latency_us`: the reference's per-document call shape (main_analyze_s2orc.py:78).
  N > 1   `value` = configs[3] STRONG-scaled: every rank generates the same seeded 1 GB corpus on its device, takes its
          byte-balanced document range (dptok.sharded.shard_bounds) and tokenizes it; no data-path collective; the
          counters are all-reduced ONCE after the timed region and checked against rank 0's single-GPU pass over the
          whole corpus (also timed: `strong_scaling.n1_ms_per_step`).  `other_workloads`: configs[4] (1.25 GB of Arabic
          text per rank = 10 GB at 8 GPUs, NCCL sum of {bytes, words, tokens, untokenizable}) and the round-1 weak
          scaling of configs[1] (100 MB per rank).
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

METRIC = "corpus_bytes_per_sec"
UNIT = "bytes/s"
WORKLOADS = {
    # name: (asset, vocabulary family, rule, description)
    "s2orc_llama2": ("llama2_32k", "spm", "RULE_SPM_LLAMA",
                     "configs[1]: Llama-2-shaped 32k SentencePiece-BPE vocab, 100 MB synthetic S2ORC-shaped abstracts"),
    "pairs_gpt2": ("gpt2_50k", "bytelevel", "RULE_GPT2",
                   "configs[2]: GPT-2-shaped 50k byte-level BPE vocab, 100 MB of en/de biomedical-translation-shaped sentence pairs"),
    "longdocs_llama3": ("llama3_128k", "bytelevel", "RULE_LLAMA3",
                        "configs[3]: Llama-3-shaped 128k byte-level vocab, 1 GB of 8-64 KB documents (60 % English, 40 % "
                        "Arabic script with diacritics), generated on the device"),
    "arabic_llama3": ("llama3_128k", "bytelevel", "RULE_LLAMA3",
                      "configs[4]: Llama-3-shaped 128k vocab, Arabic-script UTF-8 text (2-byte letters, 10 % of the words "
                      "with diacritics), 1.25 GB per GPU generated on the device (10 GB at 8 GPUs)"),
}
SEED_GLOBAL = 20250101  # device-generated corpora: the same on every rank


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


# =====================================================================================================================
# reference arm: the reference's CPU implementation on the host cores (no GPU, none of our kernels)
# =====================================================================================================================
def host_prefix_docs(workload: str, n_bytes: int):
    """The first documents (about n_bytes) of the corpus the GPU arm tokenizes for `workload`, built on the HOST."""
    from dptok import synth
    if workload == "s2orc_llama2":
        text, offs = synth.gen_documents(100_000_000, seed=0, lexicon=synth.make_lexicon(200_000, seed=0))
        raw = text.tobytes()
        docs, k = [], 0
        while k < len(offs) - 1 and offs[k] < n_bytes:
            docs.append(raw[offs[k]:offs[k + 1]].decode("utf-8"))
            k += 1
        return docs, "the first %d documents of the GPU arm's 100 MB corpus (host generator, seed 0)" % len(docs)
    from dptok import synth_device
    if workload == "longdocs_llama3":
        en = synth_device.HostLexicon(synth.make_lexicon(200_000, seed=0))
        ar = synth_device.HostLexicon(synth.make_arabic_lexicon(100_000, seed=0))
        n_docs = max(4, n_bytes // 36_000)
        docs = synth_device.generate_host(en, n_docs, SEED_GLOBAL, (1200, 9000), lex_b=ar, frac_b=0.4)
    else:
        ar = synth_device.HostLexicon(synth.make_arabic_lexicon(100_000, seed=0))
        n_docs = max(4, n_bytes // 8_000)
        docs = synth_device.generate_host(ar, n_docs, SEED_GLOBAL + 4, (200, 1200))
    return [d.decode("utf-8") for d in docs], ("the first %d documents of the GPU arm's device-generated corpus (host port "
                                               "of the generator, same seed)" % len(docs))


def run_reference_arm(args):
    """bench.py --impl reference: the reference's own CPU path (oracle/_ref = the unmodified reference modules when staged,
    else the oracle port) on all host cores, on the first documents of the corpus the GPU arm tokenizes at this N; every
    step is a bounded sample sized so that the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_baseline
    workload = args.workload or ("s2orc_llama2" if args.gpus <= 1 else "longdocs_llama3")
    asset, family, _rule, desc = WORKLOADS[workload]
    n_steps = args.warmup + args.steps
    budget = max(0.5, min(8.0, 150.0 / max(n_steps, 1)))
    docs, sample_desc = host_prefix_docs(workload, int(1_500_000 * budget) + 200_000)
    runner = cpu_baseline.Runner(asset, family=family)
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
    v = total["bytes"] / total["seconds"]
    how = ("the UNMODIFIED reference adapter (oracle/_ref, staged by oracle/make_ref.py) called per document"
           if runner.kind == "reference" else "oracle port of packages/dp_tokenize.py + the reference adapter")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total["seconds"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak" if args.gpus <= 1 else "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "tokens_per_sec": total["tokens"] / total["seconds"],
        "config": {"workload": desc, "vocab": asset, "note": "reference CPU path: " + how},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": runner.procs, "kind": runner.kind,
                         "sample": f"{total['bytes']} bytes over {args.steps} steps of {budget:.1f} s each, {sample_desc}, "
                                   "multiprocessing over documents"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# =====================================================================================================================
# our arm
# =====================================================================================================================
class Ctx:
    """torch / dptok handles of this rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.cuda.current_device()
        if self.world > 1:
            # NCCL prints its version banner to STDOUT when the communicator is created; the contract is ONE JSON line on
            # stdout, so create the communicator (init + one collective) with fd 1 pointing at stderr
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                import datetime
                # (a rank that fails must cost the others minutes, not the default ten per collective)
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank),
                                        timeout=datetime.timedelta(seconds=240))
                dist.all_reduce(torch.zeros(1, device=self.dev))
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.engines = {}
        self.lex = {}
        self.peak, self.peak_src = measured_peaks()

    def barrier(self, collective=True):
        # collective=False: a measurement only THIS rank makes (rank 0's single-GPU pass at N > 1) must not enter a
        # collective the other ranks never call
        if self.world > 1 and collective:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def engine(self, asset, family):
        from dptok import assets
        from dptok.engine import Engine
        from dptok.vocab import CompiledVocab
        if asset not in self.engines:
            if family == "spm":
                t2i = assets.load_hf(asset).get_vocab()
            else:
                t2i = {t: k for k, t in enumerate(assets.load_spec(asset)["model"]["vocab"])}
            self.engines[asset] = (Engine(CompiledVocab.from_token_map(t2i, family), self.dev), t2i)
        return self.engines[asset]

    def lexicon(self, flavour):
        from dptok import synth, synth_device
        if flavour not in self.lex:
            words = synth.make_lexicon(200_000, seed=0) if flavour == "en" else synth.make_arabic_lexicon(100_000, seed=0)
            self.lex[flavour] = (words, synth_device.DeviceLexicon(words, self.dev))
        return self.lex[flavour]


def build_workload(cx: Ctx, name: str, size_mb: float, seed_rank: int = 0, suffix_prob: float = 0.0,
                   doc_range=None):
    """-> dict(d_text, d_offs, h_text (pinned, host-generated workloads only), h_offs, desc ...)."""
    import numpy as np
    from dptok import _cabi, synth, synth_device
    torch = cx.torch
    asset, family, rule_name, desc = WORKLOADS[name]
    n_target = int(size_mb * 1e6)
    wl = dict(name=name, asset=asset, family=family, rule=getattr(_cabi, rule_name), desc=desc, h_text=None, h_offs=None,
              generated="host")
    if name == "s2orc_llama2" and suffix_prob == 0.0:
        words, _ = cx.lexicon("en")
        text, offs = synth.gen_documents(n_target, seed=seed_rank, lexicon=words)
    elif name == "pairs_gpt2":
        text, offs = synth.gen_sentence_pairs(n_target, seed=seed_rank)
    else:
        wl["generated"] = "device"
        if name == "s2orc_llama2":      # redundancy sweep: the device generator with hash suffixes on the words
            _, en = cx.lexicon("en")
            wpd = (150, 250)
            n_docs = synth_device.docs_for_bytes(n_target, wpd, en.mean_len + 4.0 * suffix_prob)
            gen = dict(lex_a=en, n_docs=n_docs, seed=SEED_GLOBAL + 1, words_per_doc=wpd, suffix_prob=suffix_prob)
        elif name == "longdocs_llama3":
            _, en = cx.lexicon("en")
            _, ar = cx.lexicon("ar")
            wpd = (1200, 9000)
            n_docs = synth_device.docs_for_bytes(n_target, wpd, 0.6 * en.mean_len + 0.4 * ar.mean_len)
            gen = dict(lex_a=en, n_docs=n_docs, seed=SEED_GLOBAL, words_per_doc=wpd, lex_b=ar, frac_b=0.4)
        else:
            _, ar = cx.lexicon("ar")
            wpd = (200, 1200)
            n_docs = synth_device.docs_for_bytes(n_target, wpd, ar.mean_len)
            gen = dict(lex_a=ar, n_docs=n_docs, seed=SEED_GLOBAL + 4, words_per_doc=wpd)
        if doc_range is not None:      # this rank's documents of a global corpus: (rank, world)
            r, w = doc_range
            per = gen["n_docs"]
            gen["doc_base"] = r * per  # every rank `per` documents: the global corpus has w * per
        d_text, d_offs = synth_device.generate(device=cx.dev, **gen)
        torch.cuda.synchronize()
        wl.update(d_text=d_text, d_offs=d_offs, n_bytes=int(d_text.numel()), n_docs=int(d_offs.numel() - 1))
        return wl
    wl["h_text"] = torch.from_numpy(np.ascontiguousarray(text).copy()).pin_memory()
    wl["h_offs"] = offs
    wl.update(d_text=wl["h_text"].to(cx.dev), d_offs=torch.from_numpy(offs).to(cx.dev), n_bytes=len(text), n_docs=len(offs) - 1)
    return wl


def measure_resident(cx: Ctx, engine, d_text, d_offs, rule, steps, warmup, with_profile=True, sampler=None, collective=True):
    """K timed steps with the text resident in HBM: CUDA events per step, L2 flushed between steps.  Then (optionally) the
    same K steps once more with a CUDA event pair around every kernel launch (per-kernel times for the roofline)."""
    from dptok import engine as eng_mod
    torch = cx.torch
    for _ in range(max(warmup, 3)):
        res = engine.encode_corpus(d_text, d_offs, rule)
    n_tokens, n_words = res.n_ids, res.n_words
    ids_cap, word_cap = res.n_ids + 1024, res.n_words + 1024
    # a corpus of mostly distinct words overflows the default word table (n_bytes / 48 slots): the engine then sizes it for
    # the worst case and runs the pass again.  The timed steps ask for that size at once, so a step is ONE pass.
    engine.encode_corpus(d_text, d_offs, rule, ids_cap=ids_cap, word_cap=word_cap)
    worst = int(engine.last_worst)  # 0 typical, 2 roomy word table, 1 worst-case sizes (chosen by the engine during the warm-up)
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    # the last warm-up pass runs with everything the timed region needs already set up (sampler thread, events), so that only
    # the barrier lies between it and the first timed step: a pass after a longer idle gap ran 4 % slower than the others
    engine.encode_corpus(d_text, d_offs, rule, ids_cap=ids_cap, word_cap=word_cap, worst_case=worst)
    launches0 = eng_mod.launch_count()
    cx.barrier(collective)
    t_wall0 = time.perf_counter()
    for k in range(steps):
        cx.flush.fill_(k & 0xFF)
        ev[k][0].record()
        res = engine.encode_corpus(d_text, d_offs, rule, ids_cap=ids_cap, word_cap=word_cap, worst_case=worst)
        ev[k][1].record()
    cx.barrier(collective)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = eng_mod.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    prof = None
    if with_profile:
        eng_mod.profile_enable(True)
        for k in range(steps):
            cx.flush.fill_(k & 0xFF)
            engine.encode_corpus(d_text, d_offs, rule, ids_cap=ids_cap, word_cap=word_cap, worst_case=worst)
        cx.torch.cuda.synchronize()
        eng_mod.profile_enable(False)
        prof = eng_mod.profile_report()
    return dict(step_ms=step_ms, total_ms=sum(step_ms), n_tokens=n_tokens, n_words=n_words, launches=launches, prof=prof,
                counters=[int(x) for x in res.counters.cpu().tolist()], t_wall=t_wall, clocks=clocks, res=res,
                ids_cap=ids_cap, word_cap=word_cap, worst=worst)


def roofline_of(cx: Ctx, m, n_bytes, steps, total_ms=None):
    """Algorithmic bytes: the whole pass moves AB = text read once + ids + per-word lengths
    written; each kernel of the pipeline is charged the part of it (plus the 4-byte word refs between the kernels) that it
    actually reads or writes, so the dominant kernel's `achieved` is its own bytes over its own time."""
    prof = m["prof"]
    if not prof:
        return None
    n_tokens, n_words = m["n_tokens"], m["n_words"]
    ab = n_bytes + 4 * n_tokens + 4 * n_words
    kernel_ab = {
        "k_scan_dedup": n_bytes + 4 * n_words,                    # text read, refs written
        "k_scan_dedup_bl": n_bytes + 4 * n_words,
        "k_emit": 4 * n_words + 4 * n_tokens + 5 * n_words,       # refs read; ids, lengths, flags written
    }
    name, cnt, ms = prof[0]
    per_launch_ms = ms / cnt
    launches_per_step = cnt / steps
    k_ab = kernel_ab.get(name, ab)
    achieved = k_ab / launches_per_step / (per_launch_ms / 1e3) / 1e9
    total_ms = m["total_ms"] if total_ms is None else total_ms
    whole = ab * steps / (total_ms / 1e3) / 1e9
    out = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": cx.peak, "unit": "GB/s",
           "frac": achieved / cx.peak, "traffic": None, "peak_source": cx.peak_src,
           "kernel_algorithmic_bytes_per_launch": k_ab / launches_per_step,
           "kernel_ms_per_launch": per_launch_ms,
           # share of the STEP (kernels on the library's side streams overlap the lock-step DP kernel: the per-kernel times
           # do not add up to the step)
           "kernel_share_of_step": ms / total_ms if total_ms > 0 else None,
           "algorithmic_bytes_per_step": ab,
           "whole_path_achieved_gbs": whole, "whole_path_frac": whole / cx.peak,
           "kernels_ms_per_step": {r[0]: r[2] / steps for r in prof},
           "note": "kernels on the library's side stream (k_dp_warp_*, k_dp_distinct) run beside k_dp_lock_*: their times overlap"}
    tr = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.isfile(tr):
        try:
            t = json.load(open(tr))
            # the capture was taken on a 100 MB corpus: only a launch over (about) as many bytes can be compared with it
            if abs(n_bytes - t.get("_capture_corpus_bytes", 100_000_000)) <= 0.05 * t.get("_capture_corpus_bytes", 100_000_000):
                out["traffic"] = t.get(name)
        except Exception:
            pass
    return out


def cpu_baseline_of(cx: Ctx, wl, budget_s: float, sample_bytes: int):
    """The reference's CPU path on the FIRST documents of this workload's corpus (copied from the device)."""
    from oracle import cpu_baseline
    torch = cx.torch
    k = int(torch.searchsorted(wl["d_offs"], torch.tensor([sample_bytes], device=cx.dev, dtype=torch.int64)).item())
    k = max(1, min(k, wl["n_docs"]))
    offs = wl["d_offs"][:k + 1].cpu().numpy()
    raw = wl["d_text"][:int(offs[k])].cpu().numpy().tobytes()
    docs = [raw[offs[j]:offs[j + 1]].decode("utf-8") for j in range(k)]
    r = cpu_baseline.run(wl["asset"], docs, budget_s=budget_s, family=wl["family"])
    return {"value": r["bytes_per_s"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "tokens_per_sec": r["tokens_per_s"],
            "sample": f"{r['bytes']} bytes / {r['docs']} documents = the first documents of this workload's corpus, "
                      f"{r['seconds']:.1f} s, " + ("the unmodified reference adapter per document (oracle/_ref)"
                                                   if r["kind"] == "reference" else "oracle port of the reference") +
                      " under multiprocessing"}


def workload_record(cx: Ctx, name, size_mb, steps, warmup, cpu_budget, suffix_prob=0.0):
    wl = build_workload(cx, name, size_mb, suffix_prob=suffix_prob)
    engine, _t2i = cx.engine(wl["asset"], wl["family"])
    m = measure_resident(cx, engine, wl["d_text"], wl["d_offs"], wl["rule"], steps, warmup)
    ms = m["total_ms"] / steps
    rec = {"workload": wl["desc"], "vocab": wl["asset"], "bytes": wl["n_bytes"], "docs": wl["n_docs"], "words": m["n_words"],
           "tokens": m["n_tokens"], "generated_on": wl["generated"], "ms_per_step": ms,
           "value": wl["n_bytes"] / (ms / 1e3), "unit": UNIT, "tokens_per_sec": m["n_tokens"] / (ms / 1e3),
           "counters": dict(zip(("bytes", "words", "tokens", "untokenizable"), m["counters"])),
           "roofline": roofline_of(cx, m, wl["n_bytes"], steps)}
    if cpu_budget > 0:
        rec["cpu_baseline"] = cpu_baseline_of(cx, wl, cpu_budget, 3_000_000)
    return rec, wl, m


def distinct_share(cx: Ctx, engine, wl, word_cap, worst=False):
    """Share of distinct words of a corpus = words the DP kernels solved / word occurrences (device counters)."""
    import numpy as np
    from dptok._cabi import lib
    tb = lib.dpt_corpus_table_workspace(wl["n_bytes"], word_cap, int(worst))
    tb_al = (tb + 255) // 256 * 256
    ctl = engine._ws[tb_al:tb_al + 64].cpu().numpy().view(np.uint32)
    distinct_share.last = {"queued_by_length_class": [int(x) for x in ctl[2:7]], "odd_words": int(ctl[7]), "long_words": int(ctl[8]),
                           "deferred_words": int(ctl[9])}
    return int(ctl[2:7].sum() + ctl[7])  # n_pending[0..4] + n_odd


def latency_record(cx: Ctx, wl_s2orc):
    """Per-call latency of the reference's call shape: dp_tokenize(one_document) (main_analyze_s2orc.py:78,
    tokenizer_utils.py:66-80) on ~1.4 KB abstracts, next to the CPU path's time for the same calls."""
    from dptok import assets
    from packages.tokenizer_utils import dp_tokenize_llama
    from oracle import adapters
    tok = assets.load_hf("llama2_32k")
    enc, _dec = dp_tokenize_llama(tok, device=cx.dev)
    offs = wl_s2orc["h_offs"]
    raw = wl_s2orc["h_text"].numpy()[:int(offs[260])].tobytes()
    docs = [raw[offs[j]:offs[j + 1]].decode("utf-8") for j in range(260)]
    for d in docs[:20]:
        enc(d)
    cx.torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = [enc(d) for d in docs[20:220]]
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = [adapters.llama_encode(tok, d) for d in docs[20:40]]
    dt_ref = time.perf_counter() - t0
    assert ref == out[:20], "per-call path disagrees with the oracle"
    rec = {"call": "dp_tokenize_llama(tok)(abstract)", "mean_doc_bytes": sum(len(d.encode()) for d in docs[20:220]) / 200,
           "latency_us": 1e6 * dt / 200, "calls": 200,
           "cpu_oracle_latency_us": 1e6 * dt_ref / 20, "cpu_oracle": "oracle.adapters.llama_encode (closed-form DP, one core), 20 calls"}
    # the byte-level adapter on the same abstracts (GPT-2-shaped 50k vocabulary, split regex on the device)
    try:
        from packages.tokenizer_utils import dp_tokenize_bloom
        btok = assets.load_hf("gpt2_50k")
        benc, _bdec = dp_tokenize_bloom(btok, None, device=cx.dev)
        for d in docs[:20]:
            benc(d)
        cx.torch.cuda.synchronize()
        t0 = time.perf_counter()
        bout = [benc(d) for d in docs[20:220]]
        bdt = time.perf_counter() - t0
        bvocab = {t: k for k, t in enumerate(assets.load_spec("gpt2_50k")["model"]["vocab"])}
        t0 = time.perf_counter()
        bref = [adapters.bytelevel_encode(btok, bvocab, d) for d in docs[20:30]]
        bdt_ref = time.perf_counter() - t0
        assert bref == bout[:10], "byte-level per-call path disagrees with the oracle"
        rec["bytelevel"] = {"call": "dp_tokenize_bloom(tok, None)(abstract)", "latency_us": 1e6 * bdt / 200, "calls": 200,
                            "device_rule": getattr(benc, "device_rule", None),
                            "cpu_oracle_latency_us": 1e6 * bdt_ref / 10}
    except Exception as e:  # noqa: BLE001
        rec["bytelevel"] = {"error": repr(e)}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mb", type=float, default=None, help="corpus bytes of the `value` workload (default: its configuration's size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip other_workloads / redundancy / latency")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="development: measure this workload as `value` (N = 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: dptok has no CPU fallback"}))
        return 2
    cx = Ctx()
    if cx.world > 1:
        return main_multi(cx, args)
    import numpy as np
    from dptok import engine as eng_mod
    steps = args.steps
    name = args.workload or "s2orc_llama2"
    size_mb = args.size_mb or {"s2orc_llama2": 100.0, "pairs_gpt2": 100.0, "longdocs_llama3": 1000.0, "arabic_llama3": 1250.0}[name]
    wl = build_workload(cx, name, size_mb)
    engine, t2i = cx.engine(wl["asset"], wl["family"])
    sampler = ClockSampler(cx.local_rank)
    m = measure_resident(cx, engine, wl["d_text"], wl["d_offs"], wl["rule"], steps, args.warmup, sampler=sampler)
    n_bytes, n_docs, n_tokens, n_words = wl["n_bytes"], wl["n_docs"], m["n_tokens"], m["n_words"]
    total_ms = m["total_ms"]
    value = n_bytes * steps / (total_ms / 1e3)
    roofline = roofline_of(cx, m, n_bytes, steps)
    n_distinct = distinct_share(cx, engine, wl, m["word_cap"], m["worst"])

    # ---- end to end through the public API with HOST buffers -------------------------------------
    e2e = None
    if not args.no_e2e and wl["h_text"] is not None:
        h_text, doc_offs, RULE = wl["h_text"], wl["h_offs"], wl["rule"]
        h_ids = torch.empty(m["ids_cap"], dtype=torch.int32).pin_memory()
        r = engine.encode_corpus_host(h_text, doc_offs, RULE, out_ids=h_ids)
        assert r.n_ids == n_tokens and int(r.counters[2]) == n_tokens, "end-to-end path disagrees with the resident path"
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = engine.encode_corpus_host(h_text, doc_offs, RULE, out_ids=h_ids)
        cx.barrier()
        dt = time.perf_counter() - t0
        e2e = {"value": n_bytes * steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(n_bytes + 8 * (n_docs + r.n_chunks)),
               "d2h_bytes_per_step": int(4 * r.n_ids + 9 * n_docs + 104 * r.n_chunks),
               "tokens_per_sec": n_tokens * steps / dt, "ms_per_step": 1e3 * dt / steps, "chunks": r.n_chunks,
               "how": "Engine.encode_corpus_host: pinned host text -> ranges at document boundaries -> H2D / kernels / D2H on "
                      "copy-in / compute / copy-out streams -> pinned host int32 ids + document offsets + counters, wall clock"}
        if max(t2i.values()) <= 0xFFFF:
            h_ids16 = torch.empty(m["ids_cap"], dtype=torch.uint16).pin_memory()
            r16 = engine.encode_corpus_host(h_text, doc_offs, RULE, out_ids=h_ids16, ids_dtype=torch.uint16)
            assert r16.n_ids == n_tokens
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                r16 = engine.encode_corpus_host(h_text, doc_offs, RULE, out_ids=h_ids16, ids_dtype=torch.uint16)
            cx.barrier()
            dt16 = time.perf_counter() - t0
            e2e["compact_u16_ids"] = {"value": n_bytes * steps / dt16, "unit": UNIT, "ms_per_step": 1e3 * dt16 / steps,
                                      "d2h_bytes_per_step": int(2 * r16.n_ids + 9 * n_docs + 104 * r16.n_chunks),
                                      "how": "same call with ids_dtype=torch.uint16 (dpt_narrow_ids_u16 on the device)"}
        if hasattr(engine, "corpus_lengths_host"):
            rl = engine.corpus_lengths_host(h_text, doc_offs, RULE)
            assert int(rl.counters[2]) == n_tokens
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                rl = engine.corpus_lengths_host(h_text, doc_offs, RULE)
            cx.barrier()
            dtl = time.perf_counter() - t0
            e2e["lengths_only"] = {"value": n_bytes * steps / dtl, "unit": UNIT, "ms_per_step": 1e3 * dtl / steps,
                                   "d2h_bytes_per_step": int(8 * (n_docs + 1) + 32),
                                   "how": "Engine.corpus_lengths_host: per-document token counts only (what the statistics loops "
                                          "main_analyze_s2orc.py:271 / main_biomed_translation.py:75-76 consume)"}

    # ---- CPU baseline beside it -------------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline_of(cx, wl, 15.0, 16_000_000)

    # ---- the other configurations, the redundancy sweep, the per-call latency ----------------------------------------
    others, redundancy, latency = {}, None, None
    if not args.no_others and name == "s2orc_llama2":
        k = max(5, min(steps, 10))
        cb = 0.0 if args.no_cpu_baseline else 5.0
        for other, mb in (("pairs_gpt2", 100.0), ("longdocs_llama3", 1000.0), ("arabic_llama3", 1250.0)):
            try:
                rec, owl, om = workload_record(cx, other, mb, k, 3, cb)
                others[other] = rec
                del owl, om
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001  (an extra measurement must not take the contract line down)
                others[other] = {"error": repr(e)}
        redundancy = {"how": "configs[1] vocabulary and rule, 100 MB from the device generator; a word gets 4 hash-derived letters "
                             "appended with probability p, which makes that occurrence a distinct word", "rows": []}
        for p in (0.0, 0.2, 1.0):
            try:
                rec, owl, om = workload_record(cx, "s2orc_llama2", 100.0, k, 3, 0.0, suffix_prob=p)
                oeng, _ = cx.engine(owl["asset"], owl["family"])
                nd = distinct_share(cx, oeng, owl, om["word_cap"], om["worst"])
                redundancy["rows"].append({"suffix_prob": p, "distinct_words": nd, "words": rec["words"],
                                           "distinct_share": nd / max(rec["words"], 1), "worst_case_table": om["worst"],
                                           "dp_work_lists": getattr(distinct_share, "last", None),
                                           "ms_per_step": rec["ms_per_step"],
                                           "value": rec["value"], "kernels_ms_per_step": rec["roofline"]["kernels_ms_per_step"]})
                del owl, om
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                redundancy["rows"].append({"suffix_prob": p, "error": repr(e)})
        try:
            latency = latency_record(cx, wl)
        except Exception as e:  # noqa: BLE001
            latency = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "tokens_per_sec": n_tokens * steps / (total_ms / 1e3),
        "bytes_per_token": n_bytes / max(n_tokens, 1),
        "config": {"workload": wl["desc"], "vocab": wl["asset"], "bytes_per_gpu": n_bytes, "docs_per_gpu": n_docs,
                   "words_per_gpu": n_words, "tokens_per_gpu": n_tokens, "distinct_words": n_distinct,
                   "parallelism": "1 GPU", "l2": "256 MiB buffer written between timed steps (L2 flush)",
                   "timing": "CUDA events per step"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(m["launches"]),
        "clocks": m["clocks"], "wall_s_timed_region": m["t_wall"], "step_ms": m["step_ms"],
        "other_workloads": others, "redundancy": redundancy, "latency": latency,
    }
    print(json.dumps(line))
    return 0


def main_multi(cx: Ctx, args):
    """N > 1: configs[3] strong-scaled over document shards; configs[4] and configs[1] (weak) in other_workloads."""
    import numpy as np
    torch, dist = cx.torch, cx.dist
    from dptok import engine as eng_mod
    from dptok.sharded import shard_bounds
    steps, world, rank = args.steps, cx.world, cx.rank
    size_mb = args.size_mb or 1000.0
    wl = build_workload(cx, "longdocs_llama3", size_mb)       # the same global corpus on every rank
    engine, _t2i = cx.engine(wl["asset"], wl["family"])
    h_offs = wl["d_offs"].cpu().numpy()
    b = shard_bounds(h_offs, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    s_text = wl["d_text"][int(h_offs[lo]):int(h_offs[hi])].clone()  # this rank's shard in a buffer of its own (16-byte aligned)
    s_offs = (wl["d_offs"][lo:hi + 1] - wl["d_offs"][lo]).contiguous()
    sampler = ClockSampler(cx.local_rank)
    m = measure_resident(cx, engine, s_text, s_offs, wl["rule"], steps, args.warmup, sampler=sampler)
    # ONE reduction of the counters, after the timed region (north_star: "only for the final token-count reduction")
    counters = torch.tensor(m["counters"], dtype=torch.int64, device=cx.dev)
    t = torch.tensor([m["total_ms"]], dtype=torch.float64, device=cx.dev)
    dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    g = [int(x) for x in counters.tolist()]
    shard_bytes = torch.tensor([int(s_text.numel())], dtype=torch.int64, device=cx.dev)
    gathered = [torch.zeros_like(shard_bytes) for _ in range(world)]
    dist.all_gather(gathered, shard_bytes)
    roofline = roofline_of(cx, m, int(s_text.numel()), steps)
    # rank 0 alone: the whole corpus on one GPU - the reference point of the strong scaling and of the counters
    strong = None
    if rank == 0:
        m1 = measure_resident(cx, engine, wl["d_text"], wl["d_offs"], wl["rule"], max(3, min(steps, 5)), 3, with_profile=False,
                              collective=False)
        n1_ms = m1["total_ms"] / len(m1["step_ms"])
        strong = {"n1_ms_per_step": n1_ms, "nN_ms_per_step": total_ms / steps, "speedup": n1_ms / (total_ms / steps),
                  "counters_single_gpu": m1["counters"], "counters_reduced": g, "counters_match": m1["counters"] == g,
                  "how": "rank 0 tokenizes the whole corpus alone (other ranks idle) right after the timed region: same box, "
                         "same corpus"}
        del m1
    cx.barrier()
    value = g[0] * steps / (total_ms / 1e3)
    # ---- end to end at N GPUs: each rank's shard from PINNED HOST memory through Engine.encode_corpus_host ------------------
    e2e = None
    if not args.no_e2e:
        k = max(3, min(steps, 5))
        h_text = torch.empty(s_text.numel(), dtype=torch.uint8).pin_memory()
        h_text.copy_(s_text)
        h_so = s_offs.cpu().numpy()
        r = engine.encode_corpus_host(h_text, h_so, wl["rule"])
        assert r.n_ids == m["n_tokens"], "end-to-end path disagrees with the resident path"
        h_ids = torch.empty(r.n_ids + 1024, dtype=torch.int32).pin_memory()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            r = engine.encode_corpus_host(h_text, h_so, wl["rule"], out_ids=h_ids)
        cx.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=cx.dev)
        io = torch.tensor([int(s_text.numel() + 8 * (hi - lo + r.n_chunks)), int(4 * r.n_ids + 9 * (hi - lo) + 104 * r.n_chunks)],
                          dtype=torch.int64, device=cx.dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(io, op=dist.ReduceOp.SUM)
        e2e = {"value": g[0] * k / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(io[0].item()),
               "d2h_bytes_per_step": int(io[1].item()), "ms_per_step": 1e3 * float(dt.item()) / k, "steps": k,
               "how": "every rank: its shard in pinned host memory -> Engine.encode_corpus_host (H2D / kernels / D2H on three "
                      "streams) -> pinned host int32 ids; wall clock, max over ranks; bytes summed over ranks"}
        del h_text, h_ids

    # ---- configs[4]: Arabic-script text generated on the device, 1.25 GB per rank ---------------------------------------------
    others = {}
    if not args.no_others:
        k = max(3, min(steps, 5))
        try:
            del s_text, s_offs
            wl_keep_desc = wl["desc"]
            wl = None
            torch.cuda.empty_cache()
            awl = build_workload(cx, "arabic_llama3", 1250.0, doc_range=(rank, world))
            aeng, _ = cx.engine(awl["asset"], awl["family"])
            am = measure_resident(cx, aeng, awl["d_text"], awl["d_offs"], awl["rule"], k, 3)
            ac = torch.tensor(am["counters"], dtype=torch.int64, device=cx.dev)
            at = torch.tensor([am["total_ms"]], dtype=torch.float64, device=cx.dev)
            dist.all_reduce(ac, op=dist.ReduceOp.SUM)       # the NCCL count reduction of configs[4]
            dist.all_reduce(at, op=dist.ReduceOp.MAX)
            ag = [int(x) for x in ac.tolist()]
            a_ms = float(at.item()) / k
            others["arabic_llama3"] = {
                "workload": awl["desc"], "scaling": "weak", "ms_per_step": a_ms, "value": ag[0] / (a_ms / 1e3), "unit": UNIT,
                "tokens_per_sec": ag[2] / (a_ms / 1e3),
                "counters_all_ranks": dict(zip(("bytes", "words", "tokens", "untokenizable"), ag)),
                "bytes_per_token": ag[0] / max(ag[2], 1), "roofline_rank0": roofline_of(cx, am, awl["n_bytes"], k)}
            del awl, am
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            others["arabic_llama3"] = {"error": repr(e)}
            wl_keep_desc = WORKLOADS["longdocs_llama3"][3]
        try:
            swl = build_workload(cx, "s2orc_llama2", 100.0, seed_rank=rank)
            seng, _ = cx.engine(swl["asset"], swl["family"])
            sm = measure_resident(cx, seng, swl["d_text"], swl["d_offs"], swl["rule"], k, 3, with_profile=False)
            sc = torch.tensor([swl["n_bytes"]], dtype=torch.int64, device=cx.dev)
            stt = torch.tensor([sm["total_ms"]], dtype=torch.float64, device=cx.dev)
            dist.all_reduce(sc, op=dist.ReduceOp.SUM)
            dist.all_reduce(stt, op=dist.ReduceOp.MAX)
            s_ms = float(stt.item()) / k
            others["s2orc_llama2_weak"] = {"workload": swl["desc"] + ", per GPU (weak scaling, as in round 1)", "scaling": "weak",
                                           "ms_per_step": s_ms, "value": int(sc.item()) / (s_ms / 1e3), "unit": UNIT}
        except Exception as e:  # noqa: BLE001
            others["s2orc_llama2_weak"] = {"error": repr(e)}
    else:
        wl_keep_desc = wl["desc"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic (generated on the device)",
            "tokens_per_sec": g[2] * steps / (total_ms / 1e3), "bytes_per_token": g[0] / max(g[2], 1),
            "config": {"workload": wl_keep_desc, "vocab": "llama3_128k", "global_bytes": g[0], "global_words": g[1],
                       "global_tokens": g[2], "untokenizable": g[3], "shard_bytes": [int(x.item()) for x in gathered],
                       "parallelism": f"doc-sharded x{world} (dptok.sharded.shard_bounds: contiguous document ranges, equal "
                                      "bytes), no data-path collective, counters all-reduced once after the timed region",
                       "l2": "256 MiB buffer written between timed steps (L2 flush)", "timing": "CUDA events per step, max over ranks"},
            "strong_scaling": strong, "roofline": roofline, "cpu_baseline": None,
            "e2e": e2e, "gpu_launches": int(m["launches"]), "clocks": m["clocks"], "wall_s_timed_region": m["t_wall"],
            "step_ms": m["step_ms"], "other_workloads": others,
        }
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
