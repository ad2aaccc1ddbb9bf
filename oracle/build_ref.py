"""Recipe for ``oracle/_ref/``: the reference's own certify-path modules, taken from where they lie under
``/root/reference`` so that ``bench.py --impl reference`` can time the UNMODIFIED reference (``Smooth.certify`` ->
``WrappedModel.forward`` -> ``lat2embs`` ...) on the GPU box's host cores, where ``/root/reference`` does not exist.

    python -m oracle.build_ref

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is collecting the files its path
imports: nothing is edited, and ``oracle/_ref/`` is git-ignored -- reference sources never enter
this repository's history; like the built ``.so`` the directory travels to the GPU box with the snapshot.  The
out-of-scope vendored trees (``models/*_tf_official``, autoattack examples / TF variants, figures) are not collected.
"""
from __future__ import annotations

import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
# what `import main_attack`, `attack_utils.gen_utils`, `models.smoothing_model`, `smoothing.smooth` pull in
FILES = ["main_attack.py", "certify.py"]
GLOBS = ["smoothing/*.py", "models/*.py", "attack_utils/*.py", "autoattack/*.py", "utils/*.py"]
BOUNDARIES = [f"boundaries/stylegan_ffhq_{a}_w_boundary.npy" for a in ("age", "eyeglasses", "gender", "pose", "smile")]


def build(verbose: bool = False) -> str:
    """Populate oracle/_ref/ (idempotent).  Returns the directory; raises when /root/reference is absent and nothing was
    collected before."""
    if not os.path.isdir(os.path.join(SRC, "smoothing")):
        if os.path.isfile(os.path.join(DST, "smoothing", "smooth.py")):
            return DST
        raise RuntimeError(f"{SRC} is not available and {DST} has not been collected")
    rel = list(FILES) + BOUNDARIES
    for g in GLOBS:
        rel += [os.path.relpath(p, SRC) for p in sorted(glob.glob(os.path.join(SRC, g)))]
    for r in rel:
        dst = os.path.join(DST, r)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, r), dst)
        if verbose:
            print("collected", r)
    return DST


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
