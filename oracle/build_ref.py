"""TEST INFRASTRUCTURE — recipe for oracle/_ref: the UNMODIFIED reference made available to the GPU box.

The reference is pure Python (no native code to compile), so "building" it is a verbatim copy of its package
sources from where they lie (/root/reference/seesaw/**/*.py) into oracle/_ref/seesaw/.  oracle/_ref/ is listed in
.gitignore (reference sources never enter this repository's history) but not in .gpurunignore, so it travels to
the GPU box like the built .so; there oracle/refstubs.py imports it under the same stub modules as in the build
container and `bench.py --impl reference` / the `cpu_baseline` leg time the reference's OWN functions
(MultiscaleIndex._query_prelim, compute_exact_knn) on the box's host cores — `kind: "reference"`.

    python oracle/build_ref.py        # needs /root/reference; __graft_entry__.build() runs it when present
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SEESAW_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(verbose=True) -> bool:
    src_pkg = os.path.join(SRC, "seesaw")
    if not os.path.isdir(src_pkg):
        if verbose:
            print(f"{src_pkg} not present: oracle/_ref left as it is", file=sys.stderr)
        return os.path.isdir(os.path.join(DST, "seesaw"))
    dst_pkg = os.path.join(DST, "seesaw")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    n = 0
    for root, dirs, files in os.walk(src_pkg):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "attic")]
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(root, f), SRC)
            out = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), out)
            n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write(f"verbatim copy of {n} .py files of {src_pkg} (orm011/seesaw 1.3.0), made by oracle/build_ref.py; "
                 "git-ignored, never committed\n")
    if verbose:
        print(f"oracle/_ref: {n} reference source files copied from {src_pkg}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
