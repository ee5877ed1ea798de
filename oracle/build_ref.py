"""Recipe for `oracle/_ref/`: stage the UNMODIFIED reference sources of the hot path so that the reference's own
`SceneRF.render_rays_batch` (scenerf/models/scenerf.py:392-471, scenerf_bf.py:420-494) can be timed as the CPU arm
on the GPU box, where /root/reference does not exist.

    python oracle/build_ref.py            # copies from /root/reference (or $SCENERF_REFERENCE)

TEST INFRASTRUCTURE ONLY -- like everything under oracle/: executed by bench.py's CPU legs (`--impl reference`,
`cpu_baseline`) and by tests; never imported by the product path (scenerf_b200/).  The reference is pure Python: there
is nothing to compile, the "build" is a verbatim file copy.  Outputs go ONLY into oracle/_ref/, which is git-ignored
(the reference's sources never enter this repository's history) but not gpurun-ignored, so it travels to the GPU box
with the snapshot like the built .so files.  oracle/_ref/MANIFEST.json records the sha256 of every staged file.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
# the eight model files of the path + the two loss modules scenerf.py imports at module level
FILES = ["scenerf/models/scenerf.py", "scenerf/models/scenerf_bf.py", "scenerf/models/utils.py", "scenerf/models/pe.py",
         "scenerf/models/resnetfc.py", "scenerf/models/spherical_mapping.py", "scenerf/models/ray_som_kl.py",
         "scenerf/models/unet2d_sphere.py", "scenerf/loss/depth_metrics.py", "scenerf/loss/ss_loss.py"]


def build(reference: str | None = None, quiet: bool = False) -> bool:
    """Returns True if oracle/_ref is (now) populated; False when the reference tree is not present here."""
    ref = reference or os.environ.get("SCENERF_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "scenerf", "models")):
        return os.path.exists(os.path.join(OUT, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(ref, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "files": manifest}, f, indent=1)
    if not quiet:
        print("staged %d reference files into %s" % (len(FILES), OUT))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
