"""Host driver above the C ABI: owns device buffers (torch tensors), streams and capacity retries.

PyTorch is plumbing here (device memory, streams, torch.distributed); every result comes from the
kernels in ``csrc/`` through ``include/dptok.h``.  No CPU execution path exists.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import lib, check
from .vocab import CompiledVocab


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("dptok needs a CUDA device: the shortest-tokenization path has no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@dataclass
class EncodeResult:
    ids: torch.Tensor            # int32[n_ids] device
    word_lens: torch.Tensor      # int32[n_words] device  (len_dp[n], dp_tokenize.py:70)
    word_flags: torch.Tensor     # uint8[n_words] device  (WF_* bits)
    word_tok_offs: Optional[torch.Tensor]  # int64[n_words+1] device
    counters: torch.Tensor       # int64[4] device {bytes, words, tokens, untokenizable}
    n_ids: int
    n_words: int
    doc_tok_offs: Optional[torch.Tensor] = None  # int64[n_docs+1] device
    doc_flags: Optional[torch.Tensor] = None     # uint8[n_docs] device


@dataclass
class HostResult:
    ids: torch.Tensor            # int32[n_ids] pinned host
    doc_tok_offs: np.ndarray     # int64[n_docs+1]
    doc_flags: np.ndarray        # uint8[n_docs]
    counters: np.ndarray         # int64[4] {bytes, words, tokens, untokenizable}
    n_ids: int
    n_chunks: int


class Engine:
    """One compiled vocabulary on one GPU."""

    def __init__(self, vocab: CompiledVocab, device: Optional[int] = None):
        _require_cuda()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.vocab = vocab.upload(self.device)
        self._ws: Optional[torch.Tensor] = None

    # ---- buffers -------------------------------------------------------------------------
    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- PRESPLIT: words already split (and normalised) by the caller ---------------------------
    def encode_words(self, text: torch.Tensor, word_offs: torch.Tensor, want_tok_offs: bool = False,
                     ids_cap: Optional[int] = None) -> EncodeResult:
        """``text`` uint8 device tensor, ``word_offs`` int64[n_words+1] device tensor."""
        assert text.dtype == torch.uint8 and word_offs.dtype == torch.int64
        assert text.is_cuda and word_offs.is_cuda
        n_words = word_offs.numel() - 1
        n_bytes = text.numel()
        dev = self.device
        if ids_cap is None:
            ids_cap = n_bytes // 2 + n_words + 64
        worst = 0
        with torch.cuda.device(dev):
            for attempt in range(4):
                ids = torch.empty(max(ids_cap, 1), dtype=torch.int32, device=dev)
                lens = torch.empty(max(n_words, 1), dtype=torch.int32, device=dev)
                flags = torch.empty(max(n_words, 1), dtype=torch.uint8, device=dev)
                tok_offs = torch.empty(n_words + 1, dtype=torch.int64, device=dev) if want_tok_offs else None
                counters = torch.empty(4, dtype=torch.int64, device=dev)
                n_out = torch.empty(8, dtype=torch.int64, device=dev)
                ws_bytes = lib.dpt_encode_words_workspace(n_bytes, n_words, worst)
                ws = self._workspace(ws_bytes)
                check(lib.dpt_encode_words(self.vocab.handle, _ptr(text), _ptr(word_offs), n_words, n_bytes, _ptr(ids),
                                           ids_cap, _ptr(lens), _ptr(flags), _ptr(tok_offs), _ptr(counters), _ptr(n_out),
                                           _ptr(ws), ws.numel(), self._stream()))
                h = n_out.cpu().tolist()  # synchronises the stream
                retry = False
                if h[_cabi.NOUT_IDS] > ids_cap:
                    ids_cap = h[_cabi.NOUT_IDS]
                    retry = True
                if h[_cabi.NOUT_POOL_REQ] > h[_cabi.NOUT_POOL_CAP]:
                    if worst:
                        raise _cabi.DptError(_cabi.ECAPACITY, f"worst-case workspace still too small: {h}")
                    worst = 1
                    retry = True
                if not retry:
                    break
            else:
                raise _cabi.DptError(_cabi.ECAPACITY, f"capacity retries exhausted: {h}")
        return EncodeResult(ids[:h[0]], lens[:n_words], flags[:n_words], tok_offs, counters, h[0], n_words)

    # ---- corpus path: raw documents -> ids ------------------------------------------------------
    def encode_corpus(self, text: torch.Tensor, doc_offs: torch.Tensor, rule: int,
                      ids_cap: Optional[int] = None, word_cap: Optional[int] = None,
                      force_general: bool = False, worst_case=0) -> EncodeResult:
        """``text`` uint8 device tensor of concatenated non-empty documents; ``doc_offs`` int64[n_docs+1]
        (``doc_offs[0] == 0``, ``doc_offs[-1] == len(text)``).

        The corpus pipeline (``dpt_encode_corpus``): scan + dedup -> one DP per distinct word -> scan + emit,
        enqueued without host synchronisation; this wrapper then reads the 64-byte status vector and retries with
        larger buffers only if a capacity was exceeded.  ``force_general`` runs the non-deduplicating multi-kernel
        path instead (``dpt_encode_corpus_general``; the cross-check).  ``worst_case``: 1 / True sizes the word table and the
        odd-word and long-word scratch for ANY text of this size at once (a corpus of mostly distinct words overflows the
        default table of n_bytes / 48 slots: the first pass then reports it and the call runs a second time); 2 = the
        typical sizes with a roomy table (n_bytes / 10 slots); ``self.last_worst`` says which one the last call used."""
        assert text.dtype == torch.uint8 and doc_offs.dtype == torch.int64 and text.is_cuda and doc_offs.is_cuda
        n_bytes = text.numel()
        n_docs = doc_offs.numel() - 1
        dev = self.device
        if ids_cap is None:
            ids_cap = n_bytes // 2 + 2 * n_docs + 64
        if word_cap is None:
            word_cap = n_bytes // 3 + 2 * n_docs + 64
        if force_general:
            return self._encode_corpus_general(text, doc_offs, rule, ids_cap, word_cap)
        # 0 typical / 2 roomy word table / 1 worst case.  The engine remembers that a corpus of this size needed more than
        # the typical table (many words overflowed into the odd-word path, or a capacity was exceeded) and asks for it at
        # once the next time.
        hint = getattr(self, "_table_hint", {}).get(rule, 0)
        worst = 1 if (worst_case is True or worst_case == 1) else (2 if worst_case == 2 else hint)
        with torch.cuda.device(dev):
            for attempt in range(5):
                ws = self._workspace(lib.dpt_encode_corpus_workspace(rule, n_bytes, n_docs, word_cap, worst))
                ids = torch.empty(ids_cap, dtype=torch.int32, device=dev)
                lens = torch.empty(word_cap, dtype=torch.int32, device=dev)
                flags = torch.empty(word_cap, dtype=torch.uint8, device=dev)
                doc_tok = torch.empty(n_docs + 1, dtype=torch.int64, device=dev)
                doc_flags = torch.empty(n_docs, dtype=torch.uint8, device=dev)
                counters = torch.empty(4, dtype=torch.int64, device=dev)
                n_out = torch.empty(8, dtype=torch.int64, device=dev)
                check(lib.dpt_encode_corpus(self.vocab.handle, rule, _ptr(text), n_bytes, _ptr(doc_offs), n_docs, _ptr(ids),
                                            ids_cap, _ptr(lens), _ptr(flags), word_cap, _ptr(doc_tok), _ptr(doc_flags),
                                            _ptr(counters), _ptr(n_out), _ptr(ws), ws.numel(), worst, self._stream()))
                h = n_out.cpu().tolist()  # synchronises the stream
                retry = False
                if h[_cabi.NOUT_WORDS] > word_cap:
                    word_cap = h[_cabi.NOUT_WORDS] + 64
                    retry = True
                if h[2] > h[3] or h[4] > h[5] or h[6] > h[7]:
                    if worst == 1:
                        raise _cabi.DptError(_cabi.ECAPACITY, f"worst-case workspace still too small: {h}")
                    worst = 1
                    retry = True
                if h[_cabi.NOUT_IDS] > ids_cap and not retry:
                    ids_cap = h[_cabi.NOUT_IDS] + 64
                    retry = True
                if not retry:
                    break
            else:
                raise _cabi.DptError(_cabi.ECAPACITY, f"capacity retries exhausted: {h}")
        nw = h[_cabi.NOUT_WORDS]
        if worst == 0 and h[6] * 50 > nw:   # > 2 % of the words were solved per occurrence: the table was too small
            worst_next = 2
        else:
            worst_next = worst
        if worst_next:
            if not hasattr(self, "_table_hint"):
                self._table_hint = {}
            self._table_hint[rule] = worst_next
        self.last_worst = worst
        self.last_n_out = h
        return EncodeResult(ids[:h[0]], lens[:nw], flags[:nw], None, counters, h[0], nw, doc_tok, doc_flags)

    # ---- ONE document per call: the reference's call shape (main_analyze_s2orc.py:78, main_biomed_translation.py:142) --
    ONE_MAX_BYTES = 48 << 10

    def encode_one(self, data: bytes, rule: int):
        """Token ids of ONE document (``bytes``, at most ``ONE_MAX_BYTES``) through the corpus pipeline, for callers that
        tokenize a string at a time.  Everything such a call pays besides the kernels is per-call overhead, so nothing is
        allocated and only one wait is made: the text and its two document offsets leave in ONE pinned-host -> device copy,
        the pipeline is enqueued behind it, and ids + status + counters + flags come back in ONE device -> pinned-host
        copy.  -> (ids: np.ndarray int32, doc_flag: int, counters: list[int]) or None when a capacity was exceeded (the
        caller falls back to ``encode_corpus``, which retries)."""
        n = len(data)
        if n == 0 or n > self.ONE_MAX_BYTES:
            return None
        dev = self.device
        if not hasattr(self, "_one"):
            self._one = {}
        st = self._one.get(rule)
        if st is None:
            cap = self.ONE_MAX_BYTES
            ids_cap, word_cap = cap // 2 + 66, cap // 3 + 66
            with torch.cuda.device(dev):
                h_in = torch.empty(16 + cap, dtype=torch.uint8).pin_memory()
                d_in = torch.empty(16 + cap, dtype=torch.uint8, device=dev)
                # one output block: [n_out 8 x i64][counters 4 x i64][doc_tok 2 x i64][doc_flags 16 B][ids]
                d_out = torch.empty(128 + 4 * ids_cap, dtype=torch.uint8, device=dev)
                h_out = torch.empty(128 + 4 * ids_cap, dtype=torch.uint8).pin_memory()
                lens = torch.empty(word_cap, dtype=torch.int32, device=dev)
                flags = torch.empty(word_cap, dtype=torch.uint8, device=dev)
                ws = torch.empty(int(lib.dpt_encode_corpus_workspace(rule, cap, 1, word_cap, 0)), dtype=torch.uint8, device=dev)
            st = self._one[rule] = dict(h_in=h_in, d_in=d_in, d_out=d_out, h_out=h_out, lens=lens, flags=flags, ws=ws,
                                  ids_cap=ids_cap, word_cap=word_cap, h_in_np=h_in.numpy(), h_out_np=h_out.numpy(),
                                  h_offs=h_in.numpy()[:16].view(np.int64), h_hdr=h_out.numpy()[:112].view(np.int64),
                                  ev=torch.cuda.Event())
        h_in_np = st["h_in_np"]
        st["h_offs"][0] = 0
        st["h_offs"][1] = n
        h_in_np[16:16 + n] = np.frombuffer(data, dtype=np.uint8)
        d_in, d_out = st["d_in"], st["d_out"]
        base_in, base_out = d_in.data_ptr(), d_out.data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            d_in[:16 + n].copy_(st["h_in"][:16 + n], non_blocking=True)
            check(lib.dpt_encode_corpus(self.vocab.handle, rule, C.c_void_p(base_in + 16), n, C.c_void_p(base_in), 1,
                                        C.c_void_p(base_out + 128), st["ids_cap"], _ptr(st["lens"]), _ptr(st["flags"]),
                                        st["word_cap"], C.c_void_p(base_out + 96), C.c_void_p(base_out + 112),
                                        C.c_void_p(base_out + 64), C.c_void_p(base_out), _ptr(st["ws"]), st["ws"].numel(), 0,
                                        C.c_void_p(stream.cuda_stream)))
            out_bytes = 128 + 4 * min(st["ids_cap"], n // 2 + 66)
            st["h_out"][:out_bytes].copy_(d_out[:out_bytes], non_blocking=True)
            st["ev"].record(stream)
            st["ev"].synchronize()
        h = st["h_hdr"]
        n_ids = int(h[0])
        if (n_ids > n // 2 + 66 or int(h[1]) > st["word_cap"] or int(h[2]) > int(h[3]) or int(h[4]) > int(h[5]) or
                int(h[6]) > int(h[7])):
            return None
        ids = st["h_out_np"][128:128 + 4 * n_ids].view(np.int32)
        return ids, int(st["h_out_np"][112]), [int(x) for x in h[8:12]]

    # ---- corpus path from HOST buffers: chunked, H2D / kernels / D2H overlapped ---------------------------------------
    def encode_corpus_host(self, h_text: torch.Tensor, doc_offs: np.ndarray, rule: int, chunk_bytes: int = 16 << 20,
                           n_streams: int = 3, out_ids: Optional[torch.Tensor] = None,
                           overlap: bool = False, ids_dtype: torch.dtype = torch.int32,
                           want_ids: bool = True) -> "HostResult":
        """``h_text``: uint8 HOST tensor (pinned for full speed) of concatenated non-empty documents; ``doc_offs``:
        int64[n_docs+1] numpy array.  The corpus is cut at document boundaries into ranges of about ``chunk_bytes``.
        Range k is copied into its place in ONE device text buffer on the copy-in stream, tokenized by
        ``dpt_encode_corpus_range`` on the compute stream (ranges in order; the word table is shared by all ranges
        of the call, so a word is still solved once per corpus) and its ids are copied out on the
        copy-out stream over a ring of ``n_streams`` output slots: PCIe in both directions and the SMs work at the same
        time.  Returns host tensors.

        Measured on the B200 box (100 MB corpus, int32 ids): 3.3 ms, about the same for every chunk size from 16 to 34 MB,
        3 or 8 slots, with or without ``overlap`` (scan / DP / emit of neighbouring ranges on three streams).  The host
        link moves 75 GB/s with both directions busy (H2D alone 1.8 ms, D2H alone 1.7 ms: profiles/r2_pcie_duplex.txt), so
        2.6 ms is the floor for 100.5 + 95.2 MB; the rest is the per-range latency of the DP launches
        (profiles/r2_e2e_timeline_n1.txt).  ``ids_dtype=torch.uint16`` (vocabularies whose ids all fit
        16 bits: Llama-2 32k, GPT-2 50k) narrows the ids on the device (``dpt_narrow_ids_u16``) so that half as many
        bytes cross PCIe on the way out; an id that does not fit raises.  ``want_ids=False``: the ids stay on the device
        and only the per-document token offsets (8 bytes per document), flags and counters come back - what the statistics
        loops of the reference consume (``len(dp_tokenize(a))``, main_analyze_s2orc.py:271,
        main_biomed_translation.py:75-76); see ``corpus_lengths_host``."""
        if ids_dtype not in (torch.int32, torch.uint16):
            raise ValueError("ids_dtype must be torch.int32 or torch.uint16")
        narrow = ids_dtype == torch.uint16
        assert h_text.dtype == torch.uint8 and not h_text.is_cuda
        doc_offs = np.ascontiguousarray(doc_offs, dtype=np.int64)
        n_docs = len(doc_offs) - 1
        n_bytes = int(doc_offs[-1])
        dev = self.device
        cuts = [0]
        while cuts[-1] < n_docs:
            target = doc_offs[cuts[-1]] + chunk_bytes
            nxt = int(np.searchsorted(doc_offs, target, side="right")) - 1
            cuts.append(min(n_docs, max(nxt, cuts[-1] + 1)))
        n_chunks = len(cuts) - 1
        max_b = max(int(doc_offs[cuts[k + 1]] - doc_offs[cuts[k]]) for k in range(n_chunks))
        max_d = max(cuts[k + 1] - cuts[k] for k in range(n_chunks))
        ids_cap = max_b // 2 + 2 * max_d + 64
        word_cap = max_b // 3 + 2 * max_d + 64
        word_cap_total = n_bytes // 3 + 2 * n_docs + 64
        if not want_ids:
            out_ids = torch.empty(0, dtype=ids_dtype)
        elif out_ids is None:
            out_ids = torch.empty(n_bytes // 2 + 2 * n_docs + 64, dtype=ids_dtype).pin_memory()
        assert out_ids.dtype == ids_dtype
        out_doc_tok = np.zeros(n_docs + 1, dtype=np.int64)
        out_doc_flags = np.zeros(n_docs, dtype=np.uint8)
        totals = np.zeros(4, dtype=np.int64)
        with torch.cuda.device(dev):
            key = (rule, n_bytes, n_docs, max_b, max_d, n_streams, narrow)
            if getattr(self, "_host_key", None) != key:
                self._host = None
                ws_bytes = lib.dpt_encode_corpus_range_workspace(rule, max_b, max_d, word_cap, 0)
                tws_bytes = lib.dpt_corpus_table_workspace(n_bytes, word_cap_total, 0)
                self._host = dict(
                    streams=[torch.cuda.Stream(device=dev) for _ in range(5)],  # copy-in, scan/compute, copy-out, DP, emit
                    ev_reset=torch.cuda.Event(),
                    d_text=torch.empty(n_bytes, dtype=torch.uint8, device=dev),
                    d_offs=torch.empty(n_docs + 1, dtype=torch.int64, device=dev),
                    h_offs=torch.empty(n_docs + 1, dtype=torch.int64).pin_memory(),
                    table_ws=torch.empty(int(tws_bytes), dtype=torch.uint8, device=dev),
                    slots=[dict(ev_in=torch.cuda.Event(), ev_a=torch.cuda.Event(), ev_ab=torch.cuda.Event(),
                                ev_b=torch.cuda.Event(),
                                ev_comp=torch.cuda.Event(),
                                ev_out=torch.cuda.Event(),
                                ids=torch.empty(ids_cap, dtype=torch.int32, device=dev),
                                ids16=torch.empty(ids_cap if narrow else 1, dtype=torch.uint16, device=dev),
                                ovf=torch.zeros(1, dtype=torch.int64, device=dev),
                                lens=torch.empty(word_cap, dtype=torch.int32, device=dev),
                                flags=torch.empty(word_cap, dtype=torch.uint8, device=dev),
                                doc_tok=torch.empty(max_d + 1, dtype=torch.int64, device=dev),
                                doc_flags=torch.empty(max_d, dtype=torch.uint8, device=dev),
                                counters=torch.empty(4, dtype=torch.int64, device=dev),
                                n_out=torch.empty(8, dtype=torch.int64, device=dev),
                                h_small=torch.zeros(13, dtype=torch.int64).pin_memory(),
                                h_doc_tok=torch.empty(max_d + 1, dtype=torch.int64).pin_memory(),
                                h_doc_flags=torch.empty(max_d, dtype=torch.uint8).pin_memory(),
                                ws=torch.empty(int(ws_bytes), dtype=torch.uint8, device=dev)) for _ in range(n_streams)])
                self._host_key = key
            H = self._host
            slots = H["slots"]
            s_in, s_comp, s_out, s_dp, s_emit = H["streams"]
            cur = torch.cuda.current_stream(dev)
            for st in H["streams"]:
                st.wait_stream(cur)
            H["h_offs"].copy_(torch.from_numpy(doc_offs))
            with torch.cuda.stream(s_in):
                H["d_offs"].copy_(H["h_offs"], non_blocking=True)
            ids_base = 0
            overflow = False
            pending = []  # (chunk index, slot), in order
            trace = getattr(self, "_trace", None)

            def mark(name, k, stream):
                if trace is not None:
                    import time as _t
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record(stream)
                    trace.append((name, k, ev, _t.perf_counter()))

            def finalize(k, sl):
                nonlocal ids_base, overflow
                lo, hi = cuts[k], cuts[k + 1]
                nd = hi - lo
                sl["ev_comp"].synchronize()                 # status vector + document offsets of range k are on the host
                h = sl["h_small"].tolist()
                if h[1] > word_cap or h[0] > ids_cap or h[2] > h[3] or h[4] > h[5] or h[6] > h[7]:
                    overflow = True                         # a capacity was exceeded: the whole corpus is redone below
                    sl["ev_out"].record(s_out)
                    return
                n_ids = h[0]
                if narrow and h[12]:
                    sl["ovf"].zero_()
                    raise _cabi.DptError(_cabi.EINVAL, f"ids_dtype=uint16: {h[12]} token ids of range {k} do not fit 16 bits")
                with torch.cuda.stream(s_out):              # the host has seen ev_comp: the ids are complete
                    mark("d2h-begin", k, s_out)
                    if want_ids:
                        out_ids[ids_base:ids_base + n_ids].copy_((sl["ids16"] if narrow else sl["ids"])[:n_ids], non_blocking=True)
                    sl["ev_out"].record(s_out)
                    mark("d2h-end", k, s_out)
                out_doc_tok[lo:hi] = sl["h_doc_tok"][:nd].numpy() + ids_base
                out_doc_flags[lo:hi] = sl["h_doc_flags"][:nd].numpy()
                totals[:] += np.asarray(h[8:12], dtype=np.int64)
                ids_base += n_ids

            def range_call(sl, k, phases, reset, stream):
                lo, hi = cuts[k], cuts[k + 1]
                check(lib.dpt_encode_corpus_range(
                    self.vocab.handle, rule, _ptr(H["d_text"]), n_bytes, _ptr(H["d_offs"]), n_docs, int(doc_offs[lo]),
                    int(doc_offs[hi]), lo, hi, reset, word_cap_total, _ptr(sl["ids"]), ids_cap, _ptr(sl["lens"]),
                    _ptr(sl["flags"]), word_cap, _ptr(sl["doc_tok"]), _ptr(sl["doc_flags"]), _ptr(sl["counters"]),
                    _ptr(sl["n_out"]), _ptr(H["table_ws"]), H["table_ws"].numel(), _ptr(sl["ws"]), sl["ws"].numel(), 0,
                    phases, C.c_void_p(stream.cuda_stream)))

            for k in range(n_chunks):
                sl = slots[k % n_streams]
                if len(pending) >= n_streams:               # the slot's previous range must have left the GPU
                    finalize(*pending.pop(0))
                lo, hi = cuts[k], cuts[k + 1]
                b0, b1 = int(doc_offs[lo]), int(doc_offs[hi])
                with torch.cuda.stream(s_in):               # copy-in engine: range after range
                    mark("h2d-begin", k, s_in)
                    H["d_text"][b0:b1].copy_(h_text[b0:b1], non_blocking=True)
                    sl["ev_in"].record(s_in)
                    mark("h2d-end", k, s_in)
                # overlap=False: the three phases of a range run back to back on ONE compute stream
                if overlap:
                    # Three in-order streams, one per phase: scan(k+1) and emit(k-1) run beside the DP kernel of range
                    # k, whose small launches are latency-bound (a batch of long words takes ~0.1-0.2 ms however few
                    # words there are) and leave most of the GPU idle.  Scans stay in range order (a range may only
                    # reference table slots claimed by itself or an earlier range), DP kernels too (emit(k) needs every
                    # DP result up to k: the DP stream is in order, so ev_b of k covers them).
                    with torch.cuda.stream(s_comp):
                        s_comp.wait_event(sl["ev_in"])
                        s_comp.wait_event(sl["ev_comp"])    # the slot's previous range has been emitted (workspace reuse)
                        mark("comp-begin", k, s_comp)
                        range_call(sl, k, 1, 1 if k == 0 else 0, s_comp)
                        sl["ev_a"].record(s_comp)
                    with torch.cuda.stream(s_dp):
                        s_dp.wait_event(sl["ev_a"])
                        range_call(sl, k, 2, 0, s_dp)
                        sl["ev_b"].record(s_dp)
                with torch.cuda.stream(s_emit if overlap else s_comp):
                    cs = s_emit if overlap else s_comp
                    if overlap:
                        cs.wait_event(sl["ev_b"])
                    else:
                        cs.wait_event(sl["ev_in"])
                    cs.wait_event(sl["ev_out"])             # the slot's previous ids have been copied out
                    if not overlap:
                        mark("comp-begin", k, cs)
                    range_call(sl, k, 4 if overlap else 7, 0 if overlap else (1 if k == 0 else 0), cs)
                    if narrow:
                        check(lib.dpt_narrow_ids_u16(_ptr(sl["ids"]), _ptr(sl["n_out"]), ids_cap, _ptr(sl["ids16"]),
                                                     _ptr(sl["ovf"]), C.c_void_p(cs.cuda_stream)))
                        sl["h_small"][12:13].copy_(sl["ovf"], non_blocking=True)
                    sl["h_small"][:8].copy_(sl["n_out"], non_blocking=True)
                    sl["h_small"][8:12].copy_(sl["counters"], non_blocking=True)
                    sl["h_doc_tok"][:hi - lo + 1].copy_(sl["doc_tok"][:hi - lo + 1], non_blocking=True)
                    sl["h_doc_flags"][:hi - lo].copy_(sl["doc_flags"][:hi - lo], non_blocking=True)
                    sl["ev_comp"].record(cs)
                    mark("comp-end", k, cs)
                pending.append((k, sl))
            while pending:
                finalize(*pending.pop(0))
            s_out.synchronize()
            if overflow:
                # rare: a range exceeded a capacity.  The text is resident: redo the corpus through the retrying
                # device-resident call and copy the result out.
                for st in H["streams"]:
                    st.synchronize()
                res = self.encode_corpus(H["d_text"], H["d_offs"], rule)
                ids_base = res.n_ids
                if want_ids and out_ids.numel() < ids_base:
                    out_ids = torch.empty(ids_base, dtype=torch.int32).pin_memory()
                if narrow and ids_base and int(res.ids.max().item()) > 0xFFFF:
                    raise _cabi.DptError(_cabi.EINVAL, "ids_dtype=uint16: token ids do not fit 16 bits")
                if want_ids:
                    out_ids[:ids_base].copy_(res.ids)
                out_doc_tok[:] = res.doc_tok_offs.cpu().numpy()
                out_doc_flags[:] = res.doc_flags.cpu().numpy()
                totals[:] = np.asarray(res.counters.cpu().tolist(), dtype=np.int64)
        out_doc_tok[n_docs] = ids_base
        return HostResult(out_ids[:ids_base] if want_ids else out_ids, out_doc_tok, out_doc_flags, totals, ids_base, n_chunks)

    def corpus_lengths_host(self, h_text: torch.Tensor, doc_offs: np.ndarray, rule: int, **kw) -> "HostResult":
        """Lengths-only result of ``encode_corpus_host``: ``np.diff(result.doc_tok_offs)`` = tokens per document (the
        reference's ``len(dp_tokenize(document))``), ``result.counters`` = {bytes, words, tokens, untokenizable}; no token
        id crosses PCIe (8 bytes per document instead of 4 bytes per token on the way out)."""
        return self.encode_corpus_host(h_text, doc_offs, rule, want_ids=False, **kw)

    def _encode_corpus_general(self, text, doc_offs, rule, ids_cap, word_cap) -> EncodeResult:
        """General multi-kernel CUDA path (normalise -> DP count -> scan -> DP emit): any word length."""
        n_bytes = text.numel()
        n_docs = doc_offs.numel() - 1
        dev = self.device
        worst = 0
        with torch.cuda.device(dev):
            for attempt in range(5):
                ids = torch.empty(ids_cap, dtype=torch.int32, device=dev)
                lens = torch.empty(word_cap, dtype=torch.int32, device=dev)
                flags = torch.empty(word_cap, dtype=torch.uint8, device=dev)
                doc_tok = torch.empty(n_docs + 1, dtype=torch.int64, device=dev)
                doc_flags = torch.empty(n_docs, dtype=torch.uint8, device=dev)
                counters = torch.empty(4, dtype=torch.int64, device=dev)
                n_out = torch.zeros(8, dtype=torch.int64, device=dev)
                ws_bytes = lib.dpt_encode_corpus_general_workspace(rule, n_bytes, n_docs, word_cap, worst)
                ws = self._workspace(ws_bytes)
                rc = lib.dpt_encode_corpus_general(self.vocab.handle, rule, _ptr(text), n_bytes, _ptr(doc_offs), n_docs,
                                                   _ptr(ids), ids_cap, _ptr(lens), _ptr(flags), word_cap, _ptr(doc_tok),
                                                   _ptr(doc_flags), _ptr(counters), _ptr(n_out), _ptr(ws), ws.numel(), worst,
                                                   self._stream())
                if rc not in (_cabi.OK, _cabi.ECAPACITY):
                    check(rc)
                h = n_out.cpu().tolist()
                retry = False
                if h[_cabi.NOUT_WORDS] > word_cap:
                    word_cap = h[_cabi.NOUT_WORDS] + 64
                    retry = True
                if h[_cabi.NOUT_NORM_REQ] > h[_cabi.NOUT_NORM_CAP] or h[_cabi.NOUT_POOL_REQ] > h[_cabi.NOUT_POOL_CAP]:
                    if worst:
                        raise _cabi.DptError(_cabi.ECAPACITY, f"worst-case workspace still too small: {h}")
                    worst = 1
                    retry = True
                if h[_cabi.NOUT_IDS] > ids_cap:
                    ids_cap = h[_cabi.NOUT_IDS] + 64
                    retry = True
                if rc == _cabi.ECAPACITY and not retry:
                    check(rc)
                if not retry:
                    break
            else:
                raise _cabi.DptError(_cabi.ECAPACITY, f"capacity retries exhausted: {h}")
        nw = h[_cabi.NOUT_WORDS]
        return EncodeResult(ids[:h[0]], lens[:nw], flags[:nw], None, counters, h[0], nw, doc_tok, doc_flags)

    # ---- lattice of one word (enumerate-all API) --------------------------------------------------
    def _upload_word(self, data: bytes, unit_starts: Optional[Sequence[int]]):
        """ONE host-to-device copy for a word's bytes and (optionally) its unit-start flags -> (text ptr, flags ptr, keep)."""
        n = len(data)
        host = np.zeros(2 * n if unit_starts is not None else n, dtype=np.uint8)
        host[:n] = np.frombuffer(data, dtype=np.uint8)
        if unit_starts is not None:
            for p in unit_starts:
                if p < n:
                    host[n + p] = 1
        d = torch.from_numpy(host).to(self.device)
        return C.c_void_p(d.data_ptr()), (C.c_void_p(d.data_ptr() + n) if unit_starts is not None else None), d

    def lattice(self, data: bytes, unit_starts: Optional[Sequence[int]] = None):
        """Returns (len_dp: list[int], preds: list[list[int]]) over unit positions 0..n_units.  One copy in, one launch, one
        copy out per call: the outputs are carved out of one int32 buffer and read back together."""
        n = len(data)
        dev = self.device
        with torch.cuda.device(dev):
            p_text, p_us, _keep = self._upload_word(data, unit_starts)
            pred_cap = max(64, 8 * (n + 1))
            while True:
                # [n_out 2 | len_dp n+2 | pred_offs n+3 | pred pred_cap | scratch n+2]
                o_len, o_po, o_pr = 2, 2 + (n + 2), 2 + (n + 2) + (n + 3)
                o_sc = o_pr + pred_cap
                buf = torch.zeros(o_sc + n + 2, dtype=torch.int32, device=dev)
                base = buf.data_ptr()
                at = lambda k: C.c_void_p(base + 4 * k)  # noqa: E731
                check(lib.dpt_lattice_word(self.vocab.handle, p_text, n, p_us, at(o_len), at(o_po), at(o_pr), pred_cap,
                                           at(0), at(o_sc), self._stream()))
                h = buf[:o_sc].cpu().numpy()  # synchronises
                n_units, n_pred = int(h[0]), int(h[1])
                if n_pred <= pred_cap:
                    break
                pred_cap = n_pred
        ld = h[o_len:o_len + n_units + 1].tolist()
        po = h[o_po:o_po + n_units + 2].tolist()
        pr = h[o_pr:o_pr + n_pred].tolist()
        preds = [pr[po[u]:po[u + 1]] for u in range(n_units + 1)]
        return ld, preds

    def min_tokens(self, data: bytes, unit_starts: Optional[Sequence[int]] = None) -> float:
        """dp[n] of the infinity-initialised length-only DP (inspect_tokenizer.py:77-86); float('inf') when unreachable."""
        n = len(data)
        if n == 0:
            return 0
        dev = self.device
        with torch.cuda.device(dev):
            p_text, p_us, _keep = self._upload_word(data, unit_starts)
            buf = torch.empty(1 + n + 2, dtype=torch.int32, device=dev)  # [out 1 | scratch n+2]
            check(lib.dpt_min_tokens_word(self.vocab.handle, p_text, n, p_us, C.c_void_p(buf.data_ptr()),
                                          C.c_void_p(buf.data_ptr() + 4), self._stream()))
            v = int(buf[0].item())
        return float("inf") if v < 0 else v

    # ---- training-data feed on device ------------------------------------------------------------
    def pad_batch(self, res: EncodeResult, pad_id: int, doc_begin: int = 0, n_rows: Optional[int] = None,
                  row_len: Optional[int] = None, pad_left: bool = False, labels: Optional[EncodeResult] = None):
        """``input_ids`` / ``attention_mask`` (int64 [n_rows, row_len], on the device) for documents
        ``doc_begin .. doc_begin + n_rows`` of an encode result: what ``tokenized_dict['input_ids']`` +
        ``[1] * len`` (main_analyze_s2orc.py:87-89) become under ``DataCollatorWithPadding``.  With ``labels`` (a second
        encode result over the same documents, e.g. the target sentences) every row is ids + label ids like the custom
        collator of main_biomed_translation.py:104-124.  ``row_len`` defaults to the longest row (one device max)."""
        offs = res.doc_tok_offs
        n_docs = offs.numel() - 1
        if n_rows is None:
            n_rows = n_docs - doc_begin
        dev = self.device
        with torch.cuda.device(dev):
            lens = offs[doc_begin + 1:doc_begin + n_rows + 1] - offs[doc_begin:doc_begin + n_rows]
            if labels is not None:
                lo = labels.doc_tok_offs
                lens = lens + (lo[doc_begin + 1:doc_begin + n_rows + 1] - lo[doc_begin:doc_begin + n_rows])
            if row_len is None:
                row_len = max(int(lens.max().item()), 1)
            input_ids = torch.empty((n_rows, row_len), dtype=torch.int64, device=dev)
            mask = torch.empty((n_rows, row_len), dtype=torch.int64, device=dev)
            row_lens = torch.empty(n_rows, dtype=torch.int64, device=dev)
            check(lib.dpt_pad_batch(_ptr(res.ids), _ptr(offs), _ptr(labels.ids) if labels is not None else None,
                                    _ptr(labels.doc_tok_offs) if labels is not None else None, doc_begin, n_rows, row_len,
                                    int(pad_id), 1 if pad_left else 0, _ptr(input_ids), _ptr(mask), _ptr(row_lens),
                                    self._stream()))
        return input_ids, mask, row_lens

    # ---- decode / round trip on device ---------------------------------------
    def roundtrip_ok(self, res: EncodeResult, text: torch.Tensor, doc_offs: torch.Tensor, skip_bos: bool) -> torch.Tensor:
        n_docs = doc_offs.numel() - 1
        ok = torch.empty(n_docs, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.dpt_roundtrip_check(self.vocab.handle, _ptr(res.ids), _ptr(res.doc_tok_offs), _ptr(text),
                                          _ptr(doc_offs), n_docs, 1 if skip_bos else 0, _ptr(ok), self._stream()))
        return ok


def launch_count() -> int:
    return int(lib.dpt_launch_count())


def profile_enable(on: bool):
    lib.dpt_profile_enable(1 if on else 0)


def profile_report():
    """-> list of (kernel name, launches, total_ms) sorted by time; synchronises and clears."""
    need = C.c_int64()
    check(lib.dpt_profile_report(None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    check(lib.dpt_profile_report(buf, need.value, C.byref(need)))
    rows = []
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        rows.append((name, int(cnt), float(ms)))
    return rows


def pack_documents(docs: Sequence[bytes]):
    """list of byte strings -> (uint8 array, int64 offsets) on the host."""
    offs = np.zeros(len(docs) + 1, dtype=np.int64)
    np.cumsum([len(d) for d in docs], out=offs[1:])
    text = np.frombuffer(b"".join(docs), dtype=np.uint8)
    return text, offs
