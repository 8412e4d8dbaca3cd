"""Build / run compile-time tuning variants of libdptok.so (development aid).

    python tools/variants.py build name=-DDPT_PB_CTAS=16 other="-DDPT_PB_REFILL=24 -DDPT_PC_CTAS=3" ...   (CPU box)
    python tools/variants.py run [bench args]                                                        (GPU box)

`run` benches every lib/variants/libdptok_*.so (and the default library first) with bench.py and prints the
per-kernel CUDA-event times."""
import glob, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dp-tokenization_b200", "csrc")
VDIR = os.path.join(ROOT, "dp-tokenization_b200", "lib", "variants")


def build(specs):
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, extra = spec.split("=", 1)
        out = os.path.join(VDIR, f"libdptok_{name}.so")
        subprocess.check_call(["make", "-C", CSRC, "-B", f"OUT={out}", f"EXTRA={extra}"], stdout=subprocess.DEVNULL)
        with open(out + ".flags", "w") as f:
            f.write(extra + "\n")
        print("built", name, extra)


def run(args):
    libs = [("default", None)] + [(os.path.basename(p)[9:-3], p) for p in sorted(glob.glob(os.path.join(VDIR, "libdptok_*.so")))]
    for name, path in libs:
        env = dict(os.environ)
        if path:
            env["DPT_LIB_PATH"] = path
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", os.environ.get("DPT_VARIANT_STEPS", "10"), "--warmup", "3", "--no-cpu-baseline", "--no-e2e"] + args
        p = subprocess.run(cmd, env=env, capture_output=True, text=True)
        try:
            line = json.loads(p.stdout.strip().splitlines()[-1])
            k = line["roofline"]["kernels_ms_per_step"]
            flags = open(path + ".flags").read().strip() if path and os.path.isfile(path + ".flags") else ""
            print(f"{name:14s} step {line['ms_per_step']:.4f} ms | " + " ".join(f"{n[2:]}={v:.4f}" for n, v in k.items()) + f" | {flags}", flush=True)
        except Exception as e:
            print(name, "FAILED", e, p.stderr[-500:], flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
