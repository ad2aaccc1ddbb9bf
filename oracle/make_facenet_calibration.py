"""Writes certifyingfacerecognition_b200/data/facenet_bn_calib_v1.npz: BatchNorm running statistics of the seeded
synthetic InceptionResnetV1 after one calibration pass (oracle/facenet.py, calibrate=True) over the 160^2 images of the
first 16 fixture latents.  TEST / BENCH INFRASTRUCTURE (the product package only loads the file).

  python -m oracle.make_facenet_calibration
"""
import os

import numpy as np
import torch

from certifyingfacerecognition_b200 import synthetic

from . import facenet, fixtures
from . import mc_path as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "certifyingfacerecognition_b200", "data", "facenet_bn_calib_v1.npz")


def main() -> None:
    torch.set_num_threads(os.cpu_count())
    g_sd = fixtures.stylegan_weights()
    w = torch.from_numpy(fixtures.latents(16))
    ims = []
    with torch.no_grad():
        for i in range(0, 16, M.GAN_CHUNK):
            ims.append(M.transform(M.easy_synthesize(w[i:i + M.GAN_CHUNK], g_sd, literal=False), size=facenet.INPUT_RES))
        sd = synthetic.facenet_weights(calibrated=False)
        emb = facenet.forward(torch.cat(ims), sd, calibrate=True)
        emb2 = facenet.forward(torch.cat(ims), sd)                    # eval mode on the calibrated statistics
    stats = {k: v.numpy().astype(np.float32) for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")}
    np.savez_compressed(OUT, **stats)
    d = torch.cdist(emb2, emb2)
    print(OUT, len(stats), "arrays; eval-mode embeddings: pairwise distance min/mean",
          float(d[d > 0].min()), float(d[d > 0].mean()), "calib-vs-eval cos", float(torch.nn.functional.cosine_similarity(emb, emb2).min()))


if __name__ == "__main__":
    main()
