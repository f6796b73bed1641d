"""Stage the reference's Python packages for the GPU box: /root/reference/{pyramid_dit,video_vae,diffusion_schedulers,
trainer_misc} (+ the top-level utils.py they import) -> baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels with the snapshot like the built .so files).

TEST / BASELINE INFRASTRUCTURE.  The GPU box has no /root/reference; the drop-in test (tests/test_dropin_gpu.py) and
bench.py's `gpu_eager_baseline` leg import the UNMODIFIED reference from this copy through oracle/pin/ref_shim.py.  Nothing is
edited: files are copied byte for byte (checked below), never committed, and no product module imports them.

    python oracle/pin/stage_reference.py        (also run by __graft_entry__.build() when /root/reference exists)
"""
from __future__ import annotations

import filecmp
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
SRC = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"
PACKAGES = ("pyramid_dit", "video_vae", "diffusion_schedulers", "trainer_misc")
TOP_FILES = ("utils.py",)          # video_vae/modeling_causal_conv.py:11 imports the context-parallel helpers from it


def stage(verbose: bool = True) -> bool:
    if not SRC.exists():
        if verbose:
            print(f"[stage_reference] {SRC} not present (GPU box): using the staged copy at {DST}" if DST.exists()
                  else f"[stage_reference] neither {SRC} nor {DST} exists")
        return DST.exists()
    DST.mkdir(parents=True, exist_ok=True)
    n = 0
    for pkg in PACKAGES:
        for f in (SRC / pkg).rglob("*.py"):
            out = DST / f.relative_to(SRC)
            out.parent.mkdir(parents=True, exist_ok=True)
            if not out.exists() or not filecmp.cmp(f, out, shallow=False):
                shutil.copyfile(f, out)
            n += 1
    for name in TOP_FILES:
        if not (DST / name).exists() or not filecmp.cmp(SRC / name, DST / name, shallow=False):
            shutil.copyfile(SRC / name, DST / name)
        n += 1
    (DST / "STAGED_FROM").write_text(f"{SRC} (unmodified copy of {', '.join(PACKAGES)}; {n} files)\n")
    if verbose:
        print(f"[stage_reference] {n} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
