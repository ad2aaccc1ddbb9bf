"""Throughput of the generate_data.py path: Z -> mapping -> synthesis -> 1024^2 RGB, images/s
(device time, no PNG encoding -- that part is host-side cv2 in the reference as well)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from certifyingfacerecognition_b200 import synthetic
    from certifyingfacerecognition_b200.models.stylegan_generator import StyleGANGenerator
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    reps = 6
    g = StyleGANGenerator({**synthetic.stylegan_weights(), **synthetic.mapping_weights()}, batch_size=batch)
    z = g.easy_sample(batch, "Z")
    for _ in range(2):
        g.easy_synthesize(z, latent_space_type="Z")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        zz = torch.from_numpy(z).cuda()
        w = g.map_latents(zz)
        g.lib.cfr_truncate(w.data_ptr(), g.synth.w_avg.data_ptr(), 0.7, batch, g.synth.wp2.data_ptr(), g._stream())
        g.synth.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"metric": "1024^2 images generated / s (Z -> W -> synthesis -> RGB)", "value": reps * batch / (ms * 1e-3),
                      "unit": "images/s", "batch": batch, "ms_per_batch": ms / reps}))


if __name__ == "__main__":
    main()
