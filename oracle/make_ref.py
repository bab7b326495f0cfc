"""Stage the UNMODIFIED reference for the CPU / eager-GPU comparison arms (TEST / BENCH INFRASTRUCTURE ONLY).

    python oracle/make_ref.py            # dev container only: needs /root/reference

The reference is a script tree (no setup.py / pyproject), so the base contract's `pip install --target baseline/_ref`
does not apply; this script is its equivalent: it copies the reference's Python files byte-for-byte into
`baseline/_ref/` (git-ignored - no reference source enters the history - but NOT gpurun-ignored, so it travels to
the GPU box, where /root/reference does not exist).  `oracle/ref_loader.py` imports it from there.
`__graft_entry__.build()` runs this when /root/reference is present.
"""
import os
import shutil
import sys

REF = os.environ.get("SSV_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(ref=REF, dst=DST):
    if not os.path.isdir(ref):
        return False
    n = 0
    for sub in ("utils", "models", "networks"):
        for dirpath, _, files in os.walk(os.path.join(ref, sub)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(dirpath, f)
                out = os.path.join(dst, os.path.relpath(src, ref))
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(src, out)
                n += 1
    with open(os.path.join(dst, "STAGED_FROM"), "w") as fh:
        fh.write(f"{ref} ({n} files, copied unmodified by oracle/make_ref.py)\n")
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else f"{REF} not found", DST)
    sys.exit(0 if ok else 1)
