"""What the host link gives: H2D alone, D2H alone, and both at once on two streams (development aid; evidence for the
end-to-end numbers in profiles/README.md).  usage: python tools/pcie_duplex.py [MB]"""
import sys
import time

import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = mb * 1000 * 1000
dev = torch.device("cuda", 0)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(do_in, do_out, reps=20, chunks=1):
    step = n // chunks
    for _ in range(3):
        if do_in:
            d_in.copy_(h_in, non_blocking=True)
        if do_out:
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for c in range(chunks):
            sl = slice(c * step, (c + 1) * step)
            if do_in:
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt


for chunks in (1, 6):
    a = run(True, False, chunks=chunks)
    b = run(False, True, chunks=chunks)
    c = run(True, True, chunks=chunks)
    print(f"{mb} MB each way, {chunks} chunk(s): H2D alone {n / a / 1e9:6.1f} GB/s ({a * 1e3:.2f} ms)   D2H alone {n / b / 1e9:6.1f} GB/s "
          f"({b * 1e3:.2f} ms)   both at once {2 * n / c / 1e9:6.1f} GB/s in total ({c * 1e3:.2f} ms)")
