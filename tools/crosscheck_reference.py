"""Time-boxed cross-check of every restatement of the DP against the UNMODIFIED reference (needs /root/reference; CPU only).

    python tools/crosscheck_reference.py [minutes] [seed]

Per random case (word of 1..16 units over a small alphabet incl. multi-byte characters, vocabulary of 1..20 random pieces,
sometimes a list of multi-character units): the reference's compute_shortest_tokenizations + obtain_longest_token
(packages/dp_tokenize.py, imported through oracle/ref_harness.py) against
  * oracle.dp_oracle.enumerate_shortest (literal restatement) and select_shortest (closed form of SURVEY 8.1),
  * oracle/dp_oracle.c through oracle.c_oracle.COracle (the bulk checker of the GPU tests),
  * the device code of csrc/dpt_dp_core.h compiled for the host (tests/host_sim: dpt_forward + dpt_backward_emit).
Exit status 1 on the first mismatch (printed with its seed)."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dp-tokenization_b200"), os.path.join(ROOT, "tests")]

import ctypes  # noqa: E402

import conftest  # noqa: E402
from helpers import make_sim_vocab, pack, sim_word  # noqa: E402
from oracle import dp_oracle, ref_harness  # noqa: E402
from oracle.c_oracle import COracle  # noqa: E402


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 7_000_000
    assert ref_harness.available(), "the reference tree is not here"
    dp, _ = ref_harness.load()
    sim = conftest.build_host_sim()
    t_end = time.time() + 60 * minutes
    it = 0
    while time.time() < t_end:
        seed = seed0 + it
        rng = random.Random(seed)
        alpha = rng.choice(["ab", "abc", "abé▁日", "xyzé", "aб日\U0001f600"])
        s = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 16)))
        vocab = {"".join(rng.choice(alpha) for _ in range(rng.randint(1, 5))) for _ in range(rng.randint(1, 20))}
        if rng.random() < 0.6:
            vocab |= set(alpha)
        units = s
        if it % 5 == 0:
            units, k = [], 0
            while k < len(s):
                step = rng.randint(1, 2)
                units.append(s[k:k + step])
                k += step
        ref_all, ref_len = dp.compute_shortest_tokenizations(units, vocab, False, "")
        ref_sel = dp.obtain_longest_token(ref_all) if ref_all else None
        ok = dp_oracle.enumerate_shortest(units, vocab) == (ref_all, ref_len)
        sel, l2, cnt = dp_oracle.select_shortest(units, vocab)
        ok = ok and l2 == ref_len and cnt == len(ref_all) and sel == ref_sel
        if isinstance(units, str):  # the byte-level checkers take code-point units
            bv = {t.encode(): i for i, t in enumerate(sorted(vocab))}
            inv = {i: t for t, i in bv.items()}
            text, offs = pack([s.encode()])
            ids, lens, untok = COracle(bv, 1).encode_words(text, offs)
            ok = ok and int(lens[0]) == ref_len and bool(untok[0]) == (ref_sel is None)
            if ref_sel is not None:
                ok = ok and [inv[i].decode() for i in ids.tolist()] == ref_sel
            h = make_sim_vocab(sim, bv, 1)
            r, wl, sids = sim_word(sim, h, s.encode())
            ok = ok and wl == ref_len and (r == -1) == (ref_sel is None)
            if ref_sel is not None:
                ok = ok and [inv[i].decode() for i in sids] == ref_sel
            sim.sim_vocab_destroy(ctypes.c_void_p(h))
        if not ok:
            print(f"MISMATCH seed={seed} units={units!r} vocab={sorted(vocab)!r} ref={(ref_sel, ref_len)}")
            return 1
        it += 1
    print(f"done: {it} cases, 0 mismatches, seeds {seed0}..{seed0 + it - 1}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
