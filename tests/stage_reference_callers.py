#!/usr/bin/env python
"""Stage the reference's own caller scripts for tests/test_reference_callers.py (BASELINE configs 1-2: "crbe.py and
experiments/crbe_experiments.py run unchanged").

The GPU box has no /root/reference, and the product has no CPU fallback, so the reference's files can only be executed
against the drop-in module where a GPU is -- they have to travel with the snapshot.  This script copies them, byte for
byte, into tests/_reference_callers/ (git-ignored: reference sources never enter this repository's history) and
checks / records their sha256 in tests/golden/reference_callers.sha256, so the test can assert it ran the UNMODIFIED
files.  Run it in the build container right before `gpurun`, and delete the directory afterwards:

    python tests/stage_reference_callers.py [--record]      # --record rewrites the .sha256 file
    python tests/stage_reference_callers.py --clean
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CRBE_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_reference_callers")
SHA = os.path.join(HERE, "golden", "reference_callers.sha256")


def main_block(path):
    """The `if __name__ == '__main__':` block of the reference's crbe.py (crbe.py:665-704), verbatim."""
    lines = open(path, encoding="utf-8").read().splitlines(keepends=True)
    start = next(i for i, ln in enumerate(lines) if ln.startswith("if __name__ =="))
    return "".join(lines[start:])


def main():
    if "--clean" in sys.argv:
        shutil.rmtree(OUT, ignore_errors=True)
        return
    os.makedirs(OUT, exist_ok=True)
    files = {"crbe_experiments.py": open(os.path.join(REF, "experiments", "crbe_experiments.py"), encoding="utf-8").read(),
             "crbe_main_block.py": main_block(os.path.join(REF, "crbe.py"))}
    sums = {}
    for name, text in files.items():
        with open(os.path.join(OUT, name), "w", encoding="utf-8") as f:
            f.write(text)
        sums[name] = hashlib.sha256(text.encode("utf-8")).hexdigest()
    if "--record" in sys.argv or not os.path.exists(SHA):
        with open(SHA, "w") as f:
            for name in sorted(sums):
                f.write(f"{sums[name]}  {name}\n")
    recorded = dict(reversed(ln.split()) for ln in open(SHA).read().splitlines() if ln.strip())
    assert recorded == sums, "staged files differ from the recorded checksums"
    print("staged", sorted(sums), "->", OUT)


if __name__ == "__main__":
    main()
