"""Recipe for oracle/_ref: a byte-for-byte copy of the reference's Python modules on the hot path.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python (no native code to compile), so "building" it means
making its UNMODIFIED modules importable where /root/reference is not mounted (the GPU box): this script copies the
files below from /root/reference into oracle/_ref/ -- git-ignored (never committed), not gpurun-ignored (it travels to
the GPU box like the built libsqlx.so).  `oracle/ref_shim.py` imports the reference from /root/reference when it is
mounted and from oracle/_ref otherwise; `bench.py --impl reference` then times the reference's own
Trainer.generate_images_pred + compute_losses + Depth_Decoder_QueryTr.forward on the box's host cores.

  python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
import filecmp
import os
import shutil
import sys

SRC = os.environ.get("SQLX_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# modules the hot path imports (trainer.py:14-20 pulls in utils, kitti_utils, layers, datasets, networks)
FILES = ["trainer.py", "layers.py", "options.py", "utils.py", "kitti_utils.py", "SQLdepth.py", "finetune/loss.py",
         "evaluate_depth_config.py"]
DIRS = ["networks", "datasets"]


def build(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "trainer.py")):
        if verbose:
            print("oracle/build_ref.py: %s not mounted, nothing to do" % SRC)
        return False
    n = 0
    todo = list(FILES)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(SRC, d))):
            if name.endswith(".py"):
                todo.append(os.path.join(d, name))
    for rel in todo:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            n += 1
    with open(os.path.join(DST, "SOURCE.txt"), "w") as f:
        f.write("unmodified copies of %d files from %s (hisfog/SfMNeXt-Impl); made by oracle/build_ref.py\n" % (len(todo), SRC))
    if verbose:
        print("oracle/build_ref.py: %d files in %s (%d refreshed)" % (len(todo), DST, n))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() or True else 1)
