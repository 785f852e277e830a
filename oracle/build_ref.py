"""Recipe for oracle/_ref/reference.zip — TEST / MEASUREMENT INFRASTRUCTURE ONLY.

The reference is pure Python: nothing of it compiles.  So that `bench.py --impl reference`, the `cpu_baseline` leg
and the `gpu_reference` leg time the REAL reference modules on the GPU box (where /root/reference does not exist),
this script packs the UNMODIFIED files the path needs, read where they lie under /root/reference, into ONE binary
artefact, oracle/_ref/reference.zip:

    models/SMOW_Net.py  models/SMOW_Net_LW.py  utils/func.py  utils/loss_f.py  utils/metric_tool.py

oracle/_ref/ is git-ignored (no reference source ever enters the history) but not gpurun-ignored, so the archive
travels to the GPU box exactly like the built .so files.  Python imports straight from the archive (zipimport);
oracle/ref_runtime.py is the only loader.  __graft_entry__.build() runs this when /root/reference is present.

    python -m oracle.build_ref
"""
import os
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SMOW_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "reference.zip")
FILES = ["models/SMOW_Net.py", "models/SMOW_Net_LW.py", "utils/func.py", "utils/loss_f.py", "utils/metric_tool.py"]


def build(force=False):
    """Returns the archive path, or None when neither the reference tree nor a previously built archive exists."""
    if not os.path.isdir(REF):
        return OUT if os.path.exists(OUT) else None
    srcs = [os.path.join(REF, f) for f in FILES]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = OUT + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for pkg in ("models", "utils"):                       # the reference relies on namespace packages
            z.writestr("smow_reference/%s/__init__.py" % pkg, "")
        z.writestr("smow_reference/__init__.py", "")
        for f, s in zip(FILES, srcs):
            z.write(s, "smow_reference/" + f)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
