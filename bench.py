"""Benchmark of the MC-certification hot path (BASELINE.json metric: MC samples/s, StyleGAN-1024 -> ArcFace vote).

    python bench.py --gpus N --steps K --warmup W          # our arm   (torchrun launches it for N > 1)
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (oracle port) on host cores

A "step" is one certification batch of ``--batch`` MC samples (BASELINE config 2: batch 250) of one identity:
noise -> latent -> StyleGAN-FFHQ-1024 synthesis -> bilinear 112 -> ArcFace iresnet50 -> argmin over a
5000-row gallery -> int64 votes.  Weights are random-init of the named architectures (seeded fixtures), data
synthetic.  N > 1: identities are sharded over the ranks (no data-path collective, weak scaling); with
``--shard samples`` every step's batch is split over the ranks and the counts are summed by one NCCL all-reduce.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GALLERY = 5000
SIGMA = 0.1
GFLOP_PER_SAMPLE_REFERENCE_FORM = 76.49   # BASELINE.md section 3 (9-tap up-convs)
GFLOP_PER_SAMPLE_SUBPIXEL_FORM = 67.59    # ... with the four non-fused up-convs in 4-tap sub-pixel form (ours)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_fixture(n_ids: int):
    """Seeded random-init weights of the named architectures + synthetic latents (package module, no oracle code)."""
    from certifyingfacerecognition_b200 import synthetic as fixtures
    g_sd, f_sd = fixtures.build_models()
    dirs = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "dirs.npy")))
    lat = torch.from_numpy(fixtures.latents(n_ids))
    return g_sd, f_sd, dirs, lat, fixtures


def cpu_arm(steps: int, warmup: int, sample: int):
    """The reference's algorithm (oracle/mc_path.py restatement; /root/reference cannot travel to the GPU box) on
    the host cores: each step classifies ``sample`` MC samples of one identity end to end."""
    from oracle import mc_path as M
    torch.set_num_threads(os.cpu_count())
    g_sd, f_sd, dirs, lat, fixtures = build_fixture(8)
    gallery = fixtures.synthetic_gallery(torch.randn(8, 512, generator=torch.Generator().manual_seed(0)) * 1.5, N_GALLERY)
    x, sigma = torch.zeros(1, 5), torch.tensor([SIGMA])
    gen = torch.Generator().manual_seed(1234)

    def step(i):
        z = lat[i % lat.shape[0]:i % lat.shape[0] + 1]
        classify = lambda p: M.wrapped_forward(z, p, dirs, gallery, g_sd, f_sd, literal=True)
        return M.sample_noise_counts(classify, x, sigma, sample, sample, N_GALLERY, generator=gen)
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = time.perf_counter() - t0
    return sample * steps / dt, dt, os.cpu_count()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=250, help="MC samples per step (BASELINE config 2: 250)")
    ap.add_argument("--chunk", type=int, default=125, help="samples per GAN+FRM program run")
    ap.add_argument("--frm-group", type=int, default=2, help="synthesis chunks per ArcFace program run")
    ap.add_argument("--shard", default="identities", choices=["identities", "samples"])
    ap.add_argument("--cpu-sample", type=int, default=4, help="MC samples per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frm", default="insightface", choices=["insightface", "facenet"],
                    help="face recognition model: ArcFace iresnet50 (headline, BASELINE config 2) or FaceNet "
                         "InceptionResnetV1 at 160^2 (config 4; GPU arm only)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"isotropic certify batch: {args.batch} MC samples/step, sigma={SIGMA}, StyleGAN-FFHQ-1024 + ArcFace "
                f"iresnet50 random-init, {N_GALLERY}-row synthetic gallery (BASELINE config 2)")
    gflop_per_sample = GFLOP_PER_SAMPLE_SUBPIXEL_FORM
    if args.frm == "facenet":
        workload = (f"isotropic certify batch: {args.batch} MC samples/step, sigma={SIGMA}, StyleGAN-FFHQ-1024 + FaceNet "
                    f"InceptionResnetV1 @160 random-init, {N_GALLERY}-row synthetic gallery (BASELINE config 4)")
        gflop_per_sample = GFLOP_PER_SAMPLE_SUBPIXEL_FORM - 12.62 + 2.835      # iresnet50 -> InceptionResnetV1 (counted
        args.no_cpu_baseline = True                                           # from the layer table, models/facenet.py)

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, args.steps)
        val, dt, cores = cpu_arm(steps, args.warmup, args.cpu_sample)
        line = {"impl": "reference", "metric": "MC samples/sec (StyleGAN1024->ArcFace vote)", "value": val,
                "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "step": f"{args.cpu_sample} MC samples of one identity (bounded sample)"},
                "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                                 "sample": f"{args.cpu_sample} samples/step x {steps} steps, oracle/mc_path.py (torch fp32, "
                                           f"{cores} threads)"},
                "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from certifyingfacerecognition_b200 import _lib as L
    from certifyingfacerecognition_b200.engine import Engine
    lib = L.load()
    n_ids = 64
    g_sd, f_sd, dirs, lat, fixtures = build_fixture(n_ids)
    if args.frm == "facenet":
        f_sd = fixtures.facenet_weights()
    eng = Engine(g_sd, f_sd, dirs, torch.zeros(1, 512), chunk=args.chunk, frm_group=args.frm_group,
                 frm="insightface" if args.frm == "insightface" else "facenet-vggface2")
    true_rows = eng.embed_latents(lat).cpu()
    eng.set_gallery(fixtures.synthetic_gallery(true_rows, N_GALLERY))
    dev = eng.device
    lat_d = lat.to(dev)
    x_d, sigma_d = torch.zeros(5, device=dev), torch.tensor([SIGMA], device=dev)
    counts = torch.zeros(N_GALLERY, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    per_rank = args.batch if args.shard == "identities" else args.batch // world
    state = {"draws": 0}

    def step(i):
        ident = (i * world + rank) % n_ids if args.shard == "identities" else i % n_ids
        off = state["draws"] + (0 if args.shard == "identities" else rank * per_rank)
        counts.zero_()
        eng.sample_votes(lat_d[ident], x_d, sigma_d, per_rank, seed=1234, sample_offset=off, counts=counts)
        if args.shard == "samples" and world > 1:
            dist.all_reduce(counts)
        state["draws"] += args.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = lib.cfr_launch_count()
    lib.cfr_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        step(1000 + i)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    prof = {}
    for kind, name in ((0, "igemm"), (1, "halo")):
        t_ms, work, n_l = C.c_double(), C.c_double(), C.c_int64()
        L.check(lib.cfr_profile_read(kind, C.byref(t_ms), C.byref(work), C.byref(n_l)))
        prof[name] = (t_ms.value, work.value, n_l.value)
    lib.cfr_profile_enable(0)
    launches = lib.cfr_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = tmax[0].item(), int(tsum[1].item())
    total_samples = per_rank * world * args.steps
    value = total_samples / (ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (H2D of z/x/sigma, D2H of the counts, sync per step)
    z_h = [lat[(i * world + rank) % n_ids].numpy().copy() for i in range(args.steps)]
    x_h, s_h = np.zeros(5, dtype=np.float32), np.array([SIGMA], dtype=np.float32)
    c_h = np.zeros(N_GALLERY, dtype=np.int64)
    sptr = C.c_void_p(stream.cuda_stream)
    for i in range(2):
        L.check(lib.cfr_sample_votes_host(eng.sampler, z_h[i % len(z_h)].ctypes.data, x_h.ctypes.data, s_h.ctypes.data, 1,
                                          per_rank, 99, 0, c_h.ctypes.data, sptr))
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        L.check(lib.cfr_sample_votes_host(eng.sampler, z_h[i].ctypes.data, x_h.ctypes.data, s_h.ctypes.data, 1, per_rank,
                                          99, i * args.batch, c_h.ctypes.data, sptr))
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_samples / te.item()
    assert int(c_h.sum()) == per_rank

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    # DRAM bytes per launch of the halo kernel from the committed `ncu --set full` capture of this very command
    # (profiles/traffic_r01_halo.json; only meaningful for the chunk size it was captured at)
    halo_traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r01_halo.json")) as fh:
            tj = json.load(fh)
        if args.chunk == 125:
            halo_traffic = tj["avg_dram_bytes_per_launch"] / 1e6
    except (OSError, KeyError, ValueError):
        pass
    ig_ms, ig_flops, ig_n = prof["igemm"]
    ha_ms, ha_bytes, ha_n = prof["halo"]
    ig_tf = ig_flops / (ig_ms * 1e-3) / 1e12 if ig_ms > 0 else 0.0
    ha_gbs = ha_bytes / (ha_ms * 1e-3) / 1e9 if ha_ms > 0 else 0.0
    # the kernel with the larger share of the step is "the dominant kernel"; the other one is reported beside it
    halo_roof = {"kernel": "conv_halo_kernel (tcgen05 halo-resident conv: StyleGAN layers 13-17, Cin<=64; HBM-bound "
                           "by arithmetic intensity, SURVEY.md section 8d)",
                 "bound": "hbm", "achieved": ha_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                 "frac": ha_gbs / pk["hbm_gbs"], "traffic": halo_traffic, "traffic_unit": "MB per launch (dram read+write, ncu)",
                 "peak_source": pk["source"],
                 "launches_timed": int(ha_n), "avg_launch_ms": ha_ms / max(1, ha_n),
                 "alg_mbytes_per_launch": ha_bytes / max(1, ha_n) / 1e6, "share_of_step": ha_ms / ms if ms > 0 else None}
    igemm_roof = {"kernel": "conv_igemm_kernel (tcgen05 implicit GEMM: StyleGAN layers 1-12 + every iresnet50 conv / FC)",
                  "bound": "tensor", "achieved": ig_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                  "frac": ig_tf / pk["tf_sustained"], "traffic": None, "peak_source": pk["source"] + " (sustained)",
                  "launches_timed": int(ig_n), "avg_launch_ms": ig_ms / max(1, ig_n),
                  "alg_gflop_per_launch": ig_flops / max(1, ig_n) / 1e9, "share_of_step": ig_ms / ms if ms > 0 else None}
    dominant, other = (halo_roof, igemm_roof) if ha_ms >= ig_ms else (igemm_roof, halo_roof)
    line = {
        "metric": "MC samples/sec (StyleGAN1024->ArcFace vote)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.shard == "identities" else "strong", "vs_baseline": None,
        "dtype": "fp16 operands / fp32 accumulate", "data": "synthetic",
        "config": {"workload": workload, "chunk": args.chunk, "frm_group": args.frm_group, "shard": args.shard, "gallery": N_GALLERY,
                   "l2": "working set per step (activations, GBs) far exceeds the 126 MB L2; no flush needed",
                   "gflop_per_sample_algorithmic": gflop_per_sample,
                   "pipeline_tflops": value * gflop_per_sample / 1e3,
                   "pipeline_frac_of_bf16_sustained": value * gflop_per_sample / 1e3 / (pk["tf_sustained"] * world)},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": 528 * 4, "d2h_bytes_per_step": N_GALLERY * 8,
                "api": "cfr_sample_votes_host (C ABI, host buffers, one call per step)"},
        "gpu_launches": int(launches),
        "roofline": dominant,
        "roofline_second_kernel": other,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        val, dt, cores = cpu_arm(2, 1, args.cpu_sample)
        line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_sample} samples/step x 2 steps (+1 warm-up) of the same workload, "
                                          f"oracle/mc_path.py torch fp32 on {cores} threads"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
