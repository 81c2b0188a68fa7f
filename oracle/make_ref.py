"""TEST / BASELINE INFRASTRUCTURE ONLY -- stage the reference's hot-path files for boxes without /root/reference.

    python oracle/make_ref.py            (called by __graft_entry__.build() where /root/reference exists)

The reference is a script tree (no setup.py / pyproject.toml, so there is nothing to pip-install into baseline/_ref).
Its MPS path is four Python files.  This recipe copies them, byte for byte, into the git-ignored directory
oracle/_ref/ -- out of the repository history, but inside the snapshot gpurun sends to the GPU box -- so that
`bench.py --impl reference` and the `cpu_baseline` leg can time the UNMODIFIED reference on the box's host cores
(through the import shim of oracle/ref_runner.py, complex128) instead of the numpy port.  Nothing under
perceptron_dqa_b200/ reads oracle/_ref; without it the CPU legs fall back to the port and say so (`kind: "port"`).
A MANIFEST with the sha256 of every staged file is written next to them.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DQA_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("dQA_mps.py", "lib/tenn.py", "lib/dQA_utils.py", "lib/dQA_loss.py", "LICENSE")


def stage(quiet=False):
    if not os.path.isfile(os.path.join(SRC, "dQA_mps.py")):
        if not quiet:
            print(f"make_ref: {SRC} not present, nothing staged")
        return False
    lines = []
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            lines.append(f"{hashlib.sha256(f.read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "MANIFEST"), "w") as f:
        f.write("# staged verbatim from baronefr/perceptron-dqa (Apache-2.0) by oracle/make_ref.py\n" + "\n".join(lines) + "\n")
    if not quiet:
        print(f"make_ref: staged {len(lines)} files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
