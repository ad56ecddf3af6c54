"""Stage the handful of UNMODIFIED reference files the on-GPU acceptance run needs under
baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun; /root/reference does not exist
there).  Nothing is edited and nothing is committed.  No-op when /root/reference is absent.

Files: ACM-Pytorch driver + models + splits for cora/squirrel, BaseLogger.py, the Cora
Planetoid pickles and the Squirrel Geom-GCN text files."""
import glob
import os
import shutil
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def stage():
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)

    def cp(rel):
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copy2(src, dst)

    for rel in ("BaseLogger.py", "ACM-Pytorch/__init__.py", "ACM-Pytorch/train.py", "ACM-Pytorch/utils.py",
                "ACM-Pytorch/arg_parser.py", "ACM-Pytorch/logger.py", "ACM-Pytorch/models/__init__.py",
                "ACM-Pytorch/models/models.py", "ACM-Pytorch/models/layers.py",
                "ACM-Geometric/layers.py", "ACM-Geometric/models.py", "ACM-Geometric/utils.py"):
        cp(rel)
    for pat in ("data/ind.cora.*", "new_data/squirrel/*", "ACM-Pytorch/splits/cora_split_0.6_0.2_*.npz",
                "ACM-Pytorch/splits/squirrel_split_0.6_0.2_*.npz"):
        for src in glob.glob(os.path.join(REF, pat)):
            cp(os.path.relpath(src, REF))
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference tree absent: nothing staged", DST)
