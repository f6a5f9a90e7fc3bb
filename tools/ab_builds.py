"""Build-variant A/B for the solve kernel (development aid).

HERE (no GPU needed, nvcc cross-compiles; the .so files are git-ignored but travel with the gpurun snapshot):
    python tools/ab_builds.py build rows64:-DTTMPC_STAGE_ROWS=64 spec:-DTTMPC_SPECULATION=1 ...
      -> tools/libttmpc_ab_<name>.so for every  name:flags  (flags separated by commas)
ON THE GPU BOX:
    python tools/ab_builds.py run [B] [N] [steps]
      -> quick_bench of the shipped library and of every tools/libttmpc_ab_*.so, one process each (TTMPC_LIB), results
         in gpurun_out/ab_builds.txt.  Extra environment (e.g. TTMPC_SPECULATE=3) is passed through.
    python tools/ab_builds.py clean
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tools")
sys.path.insert(0, ROOT)


def build(specs):
    from car_trailer_mpc_b200.build import build_library
    for spec in specs:
        name, _, flags = spec.partition(":")
        out = os.path.join(TOOLS, f"libttmpc_ab_{name}.so")
        os.environ["TTMPC_NVCC_FLAGS"] = " ".join(f for f in flags.split(",") if f)
        os.environ["TTMPC_BUILD_OUT"] = out
        print(name, "->", build_library(force=True), flush=True)
    os.environ.pop("TTMPC_NVCC_FLAGS", None)
    os.environ.pop("TTMPC_BUILD_OUT", None)


def run(args):
    B, N, steps = (args + ["65536", "40", "5"][len(args):])[:3]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    libs = [("shipped", None)] + [(os.path.basename(p)[len("libttmpc_ab_"):-3], p)
                                  for p in sorted(glob.glob(os.path.join(TOOLS, "libttmpc_ab_*.so")))]
    with open(os.path.join(ROOT, "gpurun_out", "ab_builds.txt"), "w") as f:
        for name, path in libs + [libs[0]]:   # the shipped library again at the end: drift of the box over the run
            env = dict(os.environ)
            if path:
                env["TTMPC_LIB"] = path
            r = subprocess.run([sys.executable, os.path.join(TOOLS, "quick_bench.py"), B, N, steps], env=env,
                               capture_output=True, text=True, timeout=300)
            line = (r.stdout.strip().splitlines() or [r.stderr.strip()[-300:]])[-1]
            print(f"{name:16s} {line}", flush=True)
            f.write(f"{name:16s} {line}\n")


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else ""
    if cmd == "build":
        build(sys.argv[2:])
    elif cmd == "run":
        run(sys.argv[2:])
    elif cmd == "clean":
        for p in glob.glob(os.path.join(TOOLS, "libttmpc_ab_*.so")):
            os.remove(p)
    else:
        print(__doc__)
