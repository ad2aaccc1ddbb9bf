"""Writes certifyingfacerecognition_b200/data/iresnet50_bn_calib_v3.npz: the BatchNorm running statistics of the
seeded synthetic iresnet50 after the calibration pass of ``oracle/fixtures.build_models`` (16 fixture latents through
the CPU restatement).  TEST / BENCH INFRASTRUCTURE: the product package only *loads* this file to give the random-init
benchmark network sane activations (certifyingfacerecognition_b200/synthetic.py); it never runs oracle code.

  python -m oracle.make_bn_calibration
"""
import os

import numpy as np

from . import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "certifyingfacerecognition_b200", "data", "iresnet50_bn_calib_v3.npz")


def main() -> None:
    _, f_sd = fixtures.build_models(cache_dir=os.path.join(ROOT, ".fixture_cache"))
    stats = {k: v.numpy().astype(np.float32) for k, v in f_sd.items()
             if k.endswith("running_mean") or k.endswith("running_var")}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **stats)
    print(OUT, len(stats), "arrays")


if __name__ == "__main__":
    main()
