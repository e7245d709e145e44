"""Compile the standalone hardware probes (sm_100a) into tools/probes/bin/ (git-ignored, travels with gpurun)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "probes")
BIN = os.path.join(SRC, "bin")


def main():
    os.makedirs(BIN, exist_ok=True)
    for f in sorted(os.listdir(SRC)):
        if not f.endswith(".cu"):
            continue
        src, out = os.path.join(SRC, f), os.path.join(BIN, f[:-3])
        deps = [src, os.path.join(HERE, "..", "cat-seg_b200", "csrc", "umma.cuh")]
        if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
            continue
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", src, "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise SystemExit(f"nvcc failed on {f}")


if __name__ == "__main__":
    main()
