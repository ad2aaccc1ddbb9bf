"""generate_data.py of the reference (generate_data.py:27-123) on the CUDA engine: sample / load latent codes, map Z -> W,
synthesise StyleGAN-FFHQ 1024^2 images, write `ims/%06d.png` and `{z,w,wp,styleNN}.npy` under the output directory.

Same flags and file layout as the reference.  Only the `stylegan_ffhq` model is built (PGGAN is outside SURVEY.md
section 8); weights come from `models/pretrain/stylegan_ffhq.pth` (models/model_settings.py) or `--weights`; with
`--synthetic` the seeded random-init fixture is used instead (no checkpoint is available offline)."""
from __future__ import annotations

import argparse
import os
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

from certifyingfacerecognition_b200.attack_utils.gen_utils import STYLEGAN_PATH
from certifyingfacerecognition_b200.attack_utils.proj_utils import set_seed


def parse_args():
    p = argparse.ArgumentParser(description="Generate images with given model.")
    p.add_argument("-m", "--model_name", type=str, required=True, choices=["stylegan_ffhq"],
                   help="Name of the model for generation. (required)")
    p.add_argument("-o", "--output_dir", type=str, required=True, help="Directory to save the output results. (required)")
    p.add_argument("-i", "--latent_codes_path", type=str, default="",
                   help="If specified, will load latent codes from given path instead of randomly sampling. (optional)")
    p.add_argument("-n", "--num", type=int, default=1,
                   help="Number of images to generate. Ignored if `latent_codes_path` is specified. (default: 1)")
    p.add_argument("-s", "--latent_space_type", type=str, default="z",
                   choices=["z", "Z", "w", "W", "wp", "wP", "Wp", "WP"], help="Latent space used in Style GAN. (default: `Z`)")
    p.add_argument("-S", "--generate_style", action="store_true",
                   help="If specified, will generate layer-wise style codes in Style GAN.")
    p.add_argument("-I", "--generate_image", action="store_false",
                   help="If specified, will skip generating images in Style GAN. (default: generate images)")
    # additions
    p.add_argument("--weights", type=str, default=None, help="generator state dict (default: models/pretrain/stylegan_ffhq.pth)")
    p.add_argument("--synthetic", action="store_true", help="use the seeded random-init generator (no checkpoint needed)")
    p.add_argument("--batch", type=int, default=4, help="images per synthesis batch (reference: MAX_IMAGES_ON_DEVICE = 4)")
    return p.parse_args()


def _write_png(path: str, image: torch.Tensor) -> None:
    """generate_data.py:107-109: 255 * CHW float -> HWC, written as PNG (cv2 wants BGR; PIL takes RGB directly)."""
    arr = (255.0 * image.cpu().numpy().transpose(1, 2, 0))
    try:
        import cv2
        cv2.imwrite(path, arr[:, :, ::-1])
    except ImportError:
        from PIL import Image
        Image.fromarray(np.clip(arr, 0, 255).astype(np.uint8)).save(path)


def main() -> None:
    args = parse_args()
    device = torch.device("cuda")
    set_seed(device, seed=2)                                   # generate_data.py:27
    from certifyingfacerecognition_b200.models.stylegan_generator import StyleGANGenerator
    ims_out_dir = os.path.join(args.output_dir, "ims")
    Path(ims_out_dir).mkdir(parents=True, exist_ok=True)
    if args.synthetic:
        from certifyingfacerecognition_b200 import synthetic
        g_sd = {**synthetic.stylegan_weights(), **synthetic.mapping_weights()}
    else:
        path = args.weights or STYLEGAN_PATH
        if not os.path.isfile(path):
            raise SystemExit(f"generator weights `{path}` not found (pass --weights, or --synthetic for the seeded fixture)")
        g_sd = torch.load(path, map_location="cpu")
    model = StyleGANGenerator(g_sd, batch_size=args.batch, device=device)
    kwargs = {"latent_space_type": args.latent_space_type}
    if os.path.isfile(args.latent_codes_path):
        latent_codes = model.preprocess(np.load(args.latent_codes_path), **kwargs)
    else:
        latent_codes = model.easy_sample(args.num, **kwargs)
    total_num = latent_codes.shape[0]
    results = defaultdict(list)
    n_done = 0
    for batch in model.get_batch_inputs(latent_codes):
        outputs = model.easy_synthesize(batch, **kwargs, generate_style=args.generate_style,
                                        generate_image=args.generate_image)
        for key, val in outputs.items():
            if key == "image":
                for image in val:
                    _write_png(os.path.join(ims_out_dir, f"{n_done:06d}.png"), image)
                    n_done += 1
            else:
                results[key].append(val)
        if "image" not in outputs:
            n_done += batch.shape[0]
    for key, val in results.items():
        np.save(os.path.join(args.output_dir, f"{key}.npy"), np.concatenate(val, axis=0))
    print(f"generated {total_num} samples under {args.output_dir}")


if __name__ == "__main__":
    main()
