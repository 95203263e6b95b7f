"""Builds the oracle's C restatements (oracle/*.c -> oracle/_build/liboracle.so) with gcc. Test infrastructure only.
Strict IEEE: -O2 -ffp-contract=off (no FMA contraction) so the distances are bit-identical to the numpy restatement
and to the CUDA kernel's __fmul_rn/__fadd_rn arithmetic."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_build" / "liboracle.so"


def build(force: bool = False) -> Path:
    srcs = sorted(HERE.glob("*.c"))
    if not srcs:
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    if force or not OUT.exists() or OUT.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(OUT),
               *map(str, srcs), "-lm"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"gcc failed:\n{res.stdout}\n{res.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
