"""TEST / BASELINE INFRASTRUCTURE ONLY — recipe for oracle/_ref: an UNMODIFIED, git-ignored copy of the three reference
packages the CPU arm executes (clip/, trainers/, Dassl.pytorch/dassl/ of Zehong-Ma/OVMR).

    python oracle/build_ref.py            # run by __graft_entry__.build() whenever /root/reference exists

The reference is pure Python with no installable package metadata (no setup.py / pyproject at its root), so
`pip install --target` has nothing to build; copying the package directories is the whole "install".  oracle/_ref/ is
listed in .gitignore (it never enters history) but not in .gpurunignore, so it travels to the GPU box with the snapshot,
where `bench.py --impl reference` and the `cpu_baseline` leg run it on the host cores through the import shims of
oracle/ref_loader.py (`cpu_baseline.kind = "reference"`).  Nothing under ovmr_b200/ reads it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("OVMR_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = [("clip", "clip"), ("trainers", "trainers"), (os.path.join("Dassl.pytorch", "dassl"), os.path.join("Dassl.pytorch", "dassl"))]


def build(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, "clip", "model.py")):
        if verbose:
            print(f"oracle/build_ref.py: {SRC} not present — keeping whatever oracle/_ref already holds")
        return os.path.isfile(os.path.join(DST, "clip", "model.py"))
    for src_rel, dst_rel in PACKAGES:
        src, dst = os.path.join(SRC, src_rel), os.path.join(DST, dst_rel)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name in ("LICENSE",):
        if os.path.isfile(os.path.join(SRC, name)):
            shutil.copy(os.path.join(SRC, name), os.path.join(DST, name))
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("Unmodified copy of clip/, trainers/ and Dassl.pytorch/dassl/ from Zehong-Ma/OVMR (see LICENSE), made by\n"
                "oracle/build_ref.py for the CPU reference arm of bench.py.  Git-ignored; not product source.\n")
    if verbose:
        print(f"oracle/build_ref.py: copied {[p for p, _ in PACKAGES]} -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
