"""tests/golden/mapping_vectors.npz: the UNMODIFIED reference MappingModule (models/stylegan_generator_model.py, imported
from /root/reference) on seeded weights / latents.  TEST INFRASTRUCTURE ONLY.   python -m oracle.make_golden_mapping"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    from certifyingfacerecognition_b200 import synthetic
    sys.path.insert(0, "/root/reference")
    from models.stylegan_generator_model import MappingModule          # noqa: E402  (reference, unmodified)
    sd = synthetic.mapping_weights()
    m = MappingModule()
    missing = m.load_state_dict({k[len("mapping."):]: v for k, v in sd.items()}, strict=True)
    print("load_state_dict:", missing)
    m.eval()
    z_raw = np.random.RandomState(7).randn(6, 512)
    norm = np.linalg.norm(z_raw, axis=1, keepdims=True)                 # ModStyleGANGenerator.preprocess 'Z' (:177-180)
    z = (z_raw / norm * np.sqrt(512)).astype(np.float32)
    with torch.no_grad():
        w = m(torch.from_numpy(z)).numpy()
    out = os.path.join(ROOT, "tests", "golden", "mapping_vectors.npz")
    np.savez_compressed(out, z_raw=z_raw.astype(np.float32), z=z, w=w)
    print(out, "w mean/std/absmax", w.mean(), w.std(), np.abs(w).max())


if __name__ == "__main__":
    main()
