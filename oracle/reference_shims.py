"""Import the UNMODIFIED reference from /root/reference on a CPU-only host.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden*.py`` (golden vectors, build container) and by ``bench.py``'s
CPU legs, which on the GPU box import the unmodified files collected under ``oracle/_ref`` by ``oracle/build_ref.py``
(``/root/reference`` does not exist there).

Shims (none of them touches the arithmetic on the path):
  1. stub modules ``matplotlib``, ``matplotlib.pyplot``, ``matplotlib.image``, ``facenet_pytorch``;
  2. ``scipy.stats.binom_test`` (removed in SciPy 1.12) -> ``binomtest(...).pvalue``;
  3. stub ``statsmodels.stats.proportion.proportion_confint`` (Clopper-Pearson via scipy beta);
  4. CPU device patches: ``torch.cuda.is_available`` forced True only while ``main_attack`` is
     imported (it asserts it at import), then ``model_settings.USE_CUDA=False`` and the module-level
     ``DEVICE`` constants set to cpu;
  5. a scratch cwd with ``boundaries`` symlinked and the fixture files the path loads by relative
     name (``data/stylegan_ffhq_1M/w.npy``, ``embeddings/embs_insightface.pth``,
     ``weights/ms1mv3_arcface_r50/backbone.pth``).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

def _reference_root() -> str:
    """/root/reference in the build container; on the GPU box the collected copy ``oracle/_ref`` (oracle/build_ref.py,
    git-ignored, unmodified files).  CFR_REFERENCE_ROOT overrides (testing the collected copy here)."""
    env = os.environ.get("CFR_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/smoothing"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _reference_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "smoothing", "smooth.py"))


def install_shims() -> None:
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        mpl.image = types.ModuleType("matplotlib.image")
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot
        sys.modules["matplotlib.image"] = mpl.image
    if "facenet_pytorch" not in sys.modules:
        fn = types.ModuleType("facenet_pytorch")

        class InceptionResnetV1(torch.nn.Module):       # never constructed on the insightface path
            def __init__(self, *a, **k):
                raise RuntimeError("facenet_pytorch is not installed (stub)")
        fn.InceptionResnetV1 = InceptionResnetV1
        sys.modules["facenet_pytorch"] = fn
    import scipy.stats as st
    if not hasattr(st, "binom_test"):
        st.binom_test = lambda k, n=None, p=0.5: st.binomtest(int(k), int(n), p).pvalue
    if "statsmodels" not in sys.modules:
        sm = types.ModuleType("statsmodels")
        sms = types.ModuleType("statsmodels.stats")
        smp = types.ModuleType("statsmodels.stats.proportion")

        def proportion_confint(count, nobs, alpha=0.05, method="normal"):
            assert method == "beta"
            lo = 0.0 if count == 0 else st.beta.ppf(alpha / 2, count, nobs - count + 1)
            hi = 1.0 if count == nobs else st.beta.isf(alpha / 2, count + 1, nobs - count)
            return lo, hi
        smp.proportion_confint = proportion_confint
        sm.stats = sms
        sms.proportion = smp
        sys.modules["statsmodels"] = sm
        sys.modules["statsmodels.stats"] = sms
        sys.modules["statsmodels.stats.proportion"] = smp


def import_reference(device: str = "cpu"):
    """Returns a namespace with the reference modules on the path (Smooth, L2Certificate,
    WrappedModel, gen_utils, model_settings, ...).  Must be called with cwd = scratch dir."""
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.model_settings as model_settings
    if device == "cpu":
        model_settings.USE_CUDA = False
    real_avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: True
    try:
        import main_attack
    finally:
        torch.cuda.is_available = real_avail
    import attack_utils.gen_utils as gen_utils
    dev = torch.device(device)
    gen_utils.DEVICE = dev
    main_attack.DEVICE = dev
    from smoothing.smooth import Smooth
    from smoothing.certificate import L2Certificate
    from models.smoothing_model import WrappedModel
    import models.iresnet as iresnet
    import models.stylegan_generator_model as sgm
    ns = types.SimpleNamespace(Smooth=Smooth, L2Certificate=L2Certificate, WrappedModel=WrappedModel,
                               gen_utils=gen_utils, main_attack=main_attack, model_settings=model_settings,
                               iresnet=iresnet, sgm=sgm, device=dev)
    return ns


def make_scratch(root: str, w_npy: np.ndarray, gallery: torch.Tensor, backbone_sd: dict) -> str:
    """Lay out the files certify's path reads by relative name (see module docstring, item 5)."""
    os.makedirs(root, exist_ok=True)
    link = os.path.join(root, "boundaries")
    if not os.path.exists(link):
        os.symlink(os.path.join(REFERENCE_ROOT, "boundaries"), link)
    os.makedirs(os.path.join(root, "data", "stylegan_ffhq_1M"), exist_ok=True)
    np.save(os.path.join(root, "data", "stylegan_ffhq_1M", "w.npy"), w_npy)
    os.makedirs(os.path.join(root, "embeddings"), exist_ok=True)
    torch.save(gallery, os.path.join(root, "embeddings", "embs_insightface.pth"))
    os.makedirs(os.path.join(root, "weights", "ms1mv3_arcface_r50"), exist_ok=True)
    torch.save(backbone_sd, os.path.join(root, "weights", "ms1mv3_arcface_r50", "backbone.pth"))
    return root


def load_stylegan_into(model, g_sd: dict) -> None:
    """Load the fixture synthesis/truncation tensors into a reference ModStyleGANGenerator.model
    (``mapping.*`` keeps its constructor init: not on the path)."""
    sd = model.state_dict()
    for k, v in g_sd.items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), (k, tuple(v.shape))
        sd[k] = v
    model.load_state_dict(sd)
