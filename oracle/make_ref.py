#!/usr/bin/env python
"""Stage the reference's OWN files for the scoring path into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it
travels to the GPU box like a built .so) — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python (no native code to compile), so "building" it is copying the handful of modules its
ranker imports, byte for byte, from where they lie under /root/reference.  Nothing is edited and nothing is committed:
`git ls-files oracle/_ref` stays empty.  oracle/ref_loader.py imports them with the four import-time shims of
SURVEY.md §8c (stub `faiss`, `ujson` → json, CPU device, no pinned/cuda zeros), and `bench.py --impl reference` then
times the reference's own ColbertRanker.rank_forward + BaseModel.score on the box's host cores (kind "reference").

    python oracle/make_ref.py        (run by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import os
import shutil
import sys

SRC = os.environ.get("COLBERT_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = [
    "colbert/__init__.py",
    "colbert/ranking/__init__.py",
    "colbert/ranking/colbert_ranker.py",        # ColbertRanker: __init__/_load_parts/init_ranker/rank_forward
    "colbert/modeling/__init__.py",
    "colbert/modeling/BaseModel.py",            # BaseModel.score / get_representation
    "colbert/indexing/__init__.py",
    "colbert/indexing/loaders.py",              # get_parts / load_doclens
    "colbert/indexing/index_manager.py",        # load_index_part
    "colbert/utils/__init__.py",
    "colbert/utils/utils.py",                   # print_message / flatten
]


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "colbert")):
        print(f"make_ref: {SRC} not present, nothing staged", file=sys.stderr)
        return 0
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append(f"{hashlib.sha256(open(src, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    print(f"make_ref: staged {len(FILES)} reference files into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
