"""SURVEY.md section 8e, partition A on real GPUs: the MC samples of one identity split over 2 / 4 / 8 NCCL ranks (one
process per GPU, Philox counter = global sample index, one int64 all-reduce per pass) give exactly the single-GPU vote
counts and the same Smooth.certify result.  Skips on a box with one GPU (the gloo world-2 tests in test_distributed_cpu.py
cover the host logic there)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["CFR_ROOT"])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from certifyingfacerecognition_b200 import synthetic
from certifyingfacerecognition_b200.models.smoothing_model import WrappedModel
from certifyingfacerecognition_b200.smoothing import L2Certificate, Smooth
dev = torch.device("cuda", local)
g_sd, f_sd = synthetic.build_models()
dirs = torch.from_numpy(np.load(os.path.join(os.environ["CFR_ROOT"], "tests", "golden", "dirs.npy")))
lat = torch.from_numpy(synthetic.latents(4))
model = WrappedModel(dirs.to(dev), "insightface", generator_state=g_sd, frm_state=f_sd, latents=lat,
                     orig_embs=torch.zeros(1, 512), chunk=16)
rows = model.engine.embed_latents(lat).cpu()
decoys = model.engine.embed_latents(lat[0:1] + 0.4 * dirs).cpu()          # mixed votes
model.engine.set_gallery(synthetic.synthetic_gallery(torch.cat([rows, decoys]), 300))
sigma = 0.3 * torch.tensor([0.25, 0.25, 0.04, 0.25, 0.64], device=dev)     # anisotropic (certify.py:85-95)
sm = Smooth(model, 300, sigma, L2Certificate(1, device=dev), seed=77, process_group=dist.group.WORLD if world > 1 else None)
z, x = lat[0:1].to(dev), torch.zeros(1, 5, device=dev)
c1 = sm._sample_noise(z, x, 37, 16, device=dev)
c2 = sm._sample_noise(z, x, 101, 16, device=dev)
res = sm.certify(z, x, torch.tensor([int(c2.argmax())], device=dev), 24, 88, 0.001, 16, device=dev)
# several identities certified together: selection / estimation passes share program runs and all-reduces
many = sm.certify_many(lat.to(dev), x, torch.tensor([0, 1, 3, 3], device=dev), 21, 45, 0.001)
if rank == 0:
    np.savez(os.environ["CFR_OUT"], c1=c1, c2=c2, res=np.array(res, dtype=np.float64), many=np.array(many, dtype=np.float64))
if world > 1:
    dist.destroy_process_group()
'''


def _run(world, out, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, CFR_ROOT=ROOT, CFR_OUT=out)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    subprocess.run(cmd, env=env, check=True, timeout=900)
    return np.load(out)


def test_sample_sharding_over_nccl_ranks_is_exact(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA devices")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    ref = _run(1, str(tmp_path / "w1.npz"), tmp_path)
    assert ref["c1"].sum() == 37 and ref["c2"].sum() == 101 and len(np.nonzero(ref["c2"])[0]) >= 2
    assert ref["many"].shape == (4, 2) and ref["many"][2, 1] == 0.0          # identity 2 labelled 3: early exit
    for world in (2, 4, 8):
        if world > n:
            break
        got = _run(world, str(tmp_path / f"w{world}.npz"), tmp_path)
        assert np.array_equal(got["c1"], ref["c1"]), world
        assert np.array_equal(got["c2"], ref["c2"]), world
        assert np.array_equal(got["res"], ref["res"]), world
        assert np.array_equal(got["many"], ref["many"]), world
