"""TEST / BENCH INFRASTRUCTURE -- not product code.

Installs the UNMODIFIED reference (stefanopini/simple-HRNet) into oracle/_ref/ so that it can travel to the GPU box with
the gpurun snapshot (`/root/reference` does not exist there).  This is the Python equivalent of compiling a C reference
into oracle/_ref/: the reference has no setup.py (it is not pip-installable), its hot path is four pure-Python files
plus their package markers, and `pip install --target` would do nothing but copy them.  oracle/_ref/ is git-ignored
(reference sources never enter the repo history) but not gpurun-ignored.

Installed: SimpleHRNet.py, models_/{__init__,hrnet,modules,poseresnet}.py  (the files SURVEY.md section 8a cites).
Consumers (only these): bench.py's CPU / `--impl reference` / `--impl reference-cuda` arms and
tests/test_gpu_reference_seam.py (the INTEGRATION.md stub against the real class).  Run by __graft_entry__.build()
whenever /root/reference is present."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["SimpleHRNet.py", "models_/__init__.py", "models_/hrnet.py", "models_/modules.py", "models_/poseresnet.py", "LICENSE"]


def install(src="/root/reference"):
    if not os.path.isdir(src):
        return None
    for f in FILES:
        s, d = os.path.join(src, f), os.path.join(DST, f)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    return DST


def ref_path():
    """Directory to put on sys.path to import the real reference, or None."""
    for cand in (os.environ.get("SIMPLE_HRNET_REF"), DST, "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "SimpleHRNet.py")):
            return cand
    return None


def import_reference():
    """(SimpleHRNet class, HRNet class, PoseResNet class) of the unmodified reference, or None if it is not installed."""
    p = ref_path()
    if p is None:
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    try:
        from SimpleHRNet import SimpleHRNet       # noqa: E402
        from models_.hrnet import HRNet           # noqa: E402
        from models_.poseresnet import PoseResNet  # noqa: E402
    except Exception:
        return None
    return SimpleHRNet, HRNet, PoseResNet


if __name__ == "__main__":
    print(install())
