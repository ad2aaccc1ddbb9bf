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
            self.rows.append((time.perf_counter(), [f.strip() for f in line.split(",")]))

    def mark(self):
        """Start of the timed region: only samples taken from here on count (nvidia-smi itself needs a few hundred ms to
        deliver its first line, so it is started before the warm-up steps)."""
        self.t_mark = time.perf_counter()

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
        t_mark = getattr(self, "t_mark", 0.0)
        inside = [r for t, r in self.rows if t >= t_mark]
        if not inside:                      # region shorter than one sampling period: the nearest sample
            inside = [r for _, r in self.rows[-1:]]
        for r in inside:
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


def cpu_arm(steps: int, warmup: int, sample: int, one_thread_sample: int = 0):
    # the reference prints progress lines on stdout; the bench's stdout carries exactly ONE JSON line
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        return _cpu_arm(steps, warmup, sample, one_thread_sample)


def _cpu_arm(steps: int, warmup: int, sample: int, one_thread_sample: int = 0):
    """The reference's CPU path on the host cores; each step classifies ``sample`` MC samples of one identity end to end
    (one batch of BASELINE config 1, whose batch size is 10).

    kind "reference": the UNMODIFIED reference modules -- ``Smooth._sample_noise`` -> ``WrappedModel.forward`` ->
    ``lat2embs`` -> ``compute_probs`` (smooth.py:109-138) -- imported from ``oracle/_ref`` (collected by
    oracle/build_ref.py; /root/reference itself when present) through the shims of SURVEY.md section 8c, fixture weights
    loaded into the reference's own StyleGAN / iresnet50 modules.  kind "port": oracle/mc_path.py, only when the
    collected reference is missing.  Returns (samples/s, seconds, threads, kind, samples/s on ONE thread or None)."""
    torch.set_num_threads(os.cpu_count())
    g_sd, f_sd, dirs, lat, fixtures = build_fixture(8)
    gallery = fixtures.synthetic_gallery(torch.randn(8, 512, generator=torch.Generator().manual_seed(0)) * 1.5, N_GALLERY)
    x, sigma = torch.zeros(1, 5), torch.tensor([SIGMA])
    from oracle import reference_shims as RS
    if RS.available():
        import tempfile
        kind = "reference"
        cwd = os.getcwd()
        scratch = tempfile.mkdtemp(prefix="cfr_ref_bench_")
        RS.make_scratch(scratch, lat.numpy(), gallery, f_sd)
        os.chdir(scratch)               # the reference reads its files by cwd-relative names
        try:
            ref = RS.import_reference("cpu")
            model = ref.WrappedModel(dirs, "insightface", n_embs=N_GALLERY, load_embs=True)
        finally:
            os.chdir(cwd)
        RS.load_stylegan_into(model.generator.model, g_sd)
        model.generator.model.eval()
        model.eval()
        smooth = ref.Smooth(model, N_GALLERY, sigma, ref.L2Certificate(1, device=ref.device))
        torch.manual_seed(1234)

        def step(i, n=sample):
            z = lat[i % lat.shape[0]:i % lat.shape[0] + 1]
            return smooth._sample_noise(z, x, n, n, device=ref.device)
    else:
        from oracle import mc_path as M
        kind = "port"
        gen = torch.Generator().manual_seed(1234)

        def step(i, n=sample):
            z = lat[i % lat.shape[0]:i % lat.shape[0] + 1]
            classify = lambda p: M.wrapped_forward(z, p, dirs, gallery, g_sd, f_sd, literal=True)
            return M.sample_noise_counts(classify, x, sigma, n, n, N_GALLERY, generator=gen)
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        c = step(warmup + i)
    dt = time.perf_counter() - t0
    assert int(c.sum()) == sample
    one = None
    if one_thread_sample > 0:
        torch.set_num_threads(1)
        t1 = time.perf_counter()
        step(0, one_thread_sample)
        one = one_thread_sample / (time.perf_counter() - t1)
        torch.set_num_threads(os.cpu_count())
    return sample * steps / dt, dt, os.cpu_count(), kind, one


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=250, help="MC samples per step (BASELINE config 2: 250)")
    ap.add_argument("--chunk", type=int, default=250, help="samples per GAN+FRM program run")
    ap.add_argument("--frm-group", type=int, default=1, help="synthesis chunks per ArcFace program run")
    ap.add_argument("--shard", default="identities", choices=["identities", "samples"],
                    help="how the certification-batch loop uses N > 1 ranks (the certify loop always shards samples)")
    ap.add_argument("--group", type=int, default=20,
                    help="identities certified together per step of the certify loop (Smooth.certify_many); 1 = one at a time")
    ap.add_argument("--headline-batches", action="store_true",
                    help="N > 1: keep the certification-batch loop as the headline instead of BASELINE config 3")
    ap.add_argument("--cpu-sample", type=int, default=10,
                    help="MC samples per CPU-baseline step (BASELINE config 1 classifies batches of 10)")
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
        val, dt, cores, kind, one = cpu_arm(steps, args.warmup, args.cpu_sample, one_thread_sample=2)
        what = ("the unmodified reference (oracle/_ref: Smooth._sample_noise -> WrappedModel.forward -> lat2embs, torch fp32)"
                if kind == "reference" else "oracle/mc_path.py (torch fp32 port)")
        line = {"impl": "reference", "metric": "MC samples/sec (StyleGAN1024->ArcFace vote)", "value": val,
                "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "step": f"{args.cpu_sample} MC samples of one identity = one batch of "
                                                         "BASELINE config 1 (n0=10, n=100, batch 10): a bounded sample"},
                "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind,
                                 "value_1_thread": one,
                                 "sample": f"{args.cpu_sample} samples/step x {steps} steps, {what}, {cores} threads; "
                                           "1-thread figure from one batch of 2 samples"},
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
    from certifyingfacerecognition_b200.models.smoothing_model import WrappedModel
    from certifyingfacerecognition_b200.smoothing import L2Certificate, Smooth
    lib = L.load()
    n_ids = 64
    g_sd, f_sd, dirs, lat, fixtures = build_fixture(n_ids)
    if args.frm == "facenet":
        f_sd = fixtures.facenet_weights()
    dev = torch.device("cuda", local_rank)
    # the drop-in classes own the engine (one set of buffers); the gallery is computed by the engine, as
    # WrappedModel(load_embs=False) does (smoothing_model.py:48-53), then padded to 5000 rows
    model = WrappedModel(dirs.to(dev), "insightface" if args.frm == "insightface" else "facenet-vggface2",
                         generator_state=g_sd, frm_state=f_sd, latents=lat, orig_embs=torch.zeros(1, 512),
                         chunk=args.chunk, frm_group=args.frm_group)
    eng = model.engine
    true_rows = eng.embed_latents(lat).cpu()
    eng.set_gallery(fixtures.synthetic_gallery(true_rows, N_GALLERY))
    model.orig_embs = eng.gallery
    lat_d = lat.to(dev)
    x_d, sigma_d = torch.zeros(5, device=dev), torch.tensor([SIGMA], device=dev)
    counts = torch.zeros(N_GALLERY, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    ident_mode = args.shard == "identities"
    per_rank = args.batch if ident_mode else args.batch // world
    state = {"draws": 0}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm, on_start=None):
        """K steps on the launch stream between a barrier + synchronize on both sides; device time, max over ranks."""
        for i in range(warm):
            fn(i)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if on_start:
            on_start()
        barrier()
        ev0.record(stream)
        for i in range(steps):
            fn(1000 + i)
        ev1.record(stream)
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def wall(fn, steps, warm, on_start=None):
        """The same through host buffers, wall clock (max over ranks)."""
        for i in range(warm):
            fn(i)
        if on_start:
            on_start()
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(1000 + i)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- (A) certification batches (BASELINE config 2): identities sharded over the ranks, no data-path collective
    #      (--shard samples: every batch split over the ranks + one all-reduce)
    def batch_step(i):
        ident = (i * world + rank) % n_ids if ident_mode else i % n_ids
        off = state["draws"] + (0 if ident_mode else rank * per_rank)
        counts.zero_()
        eng.sample_votes(lat_d[ident], x_d, sigma_d, per_rank, seed=1234, sample_offset=off, counts=counts)
        if not ident_mode and world > 1:
            dist.all_reduce(counts)
        state["draws"] += args.batch

    z_h = [lat[(i * world + rank) % n_ids].numpy().copy() for i in range(args.steps)]
    x_h, s_h = np.zeros(5, dtype=np.float32), np.array([SIGMA], dtype=np.float32)
    c_h = np.zeros(N_GALLERY, dtype=np.int64)
    sptr = C.c_void_p(stream.cuda_stream)

    def batch_step_host(i):
        j = i % len(z_h)
        L.check(lib.cfr_sample_votes_host(eng.sampler, z_h[j].ctypes.data, x_h.ctypes.data, s_h.ctypes.data, 1, per_rank,
                                          99, j * args.batch, c_h.ctypes.data, sptr))

    # ---- (B) whole certifications (BASELINE config 3): Smooth.certify through the drop-in API, anisotropic sigma
    #      (certify.py:85-95: sigma * eps^2), N0 = 100 selection + n = 1000 estimation samples of ONE identity split over the
    #      ranks by global sample index; both passes end in an NCCL int64 all-reduce of the [5000] vote counts (the first
    #      one feeds the early-exit decision every rank must agree on, smooth.py:66-68)
    eps2 = torch.tensor([0.25, 0.25, 0.04, 0.25, 0.64], device=dev)          # red_ellipse_mat_inv (proj_utils.py:16-21)
    smooth = Smooth(model, N_GALLERY, SIGMA * eps2, L2Certificate(1, device=dev), seed=4321,
                    process_group=dist.group.WORLD if world > 1 else None)
    N0, NEST, ALPHA = 100, 1000, 0.001
    zero5 = torch.zeros(1, 5, device=dev)
    cert_stat = {"certified": 0, "calls": 0}

    G = max(1, args.group)
    label_t = torch.arange(n_ids, device=dev)
    stage = torch.empty(G, 512).pin_memory()

    def certify_step(i, host=False):
        """One step = the certification of G identities (Smooth.certify_many: their selection passes share program runs and
        one all-reduce, then the estimation passes; G = 1 is the reference's one-identity-at-a-time loop)."""
        ids = [(i * G + k) % n_ids for k in range(G)]
        if host:                       # this step's latents: gathered into pinned staging, copied H2D inside the timed region
            torch.index_select(lat, 0, torch.tensor(ids), out=stage)
            z = stage.to(dev, non_blocking=True)
        else:
            z = lat_d[ids]
        if G == 1:
            res = [smooth.certify(z, zero5, label_t[ids], N0, NEST, ALPHA, args.batch, device=dev)]
        else:
            res = smooth.certify_many(z, zero5, label_t[ids], N0, NEST, ALPHA, args.batch, device=dev)
        cert_stat["calls"] += G
        cert_stat["certified"] += sum(int(p == j and gap > 0) for (p, gap), j in zip(res, ids))

    headline_certify = world > 1 and ident_mode and not args.headline_batches
    # at least 10 untimed steps (~0.35 s): the SM clock is still settling under the power cap during the first few, and the
    # timed region would otherwise read 3-4 % low ("warmup" echoes the W asked for, "warmup_steps_run" the number run)
    warm = max(10, args.warmup)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    results, snap = {}, {}

    def start_headline():
        if sampler:
            sampler.mark()
        snap["launches"] = lib.cfr_launch_count()

    def mark_draws():
        snap["draws"] = smooth._draws

    for name in (("batches", "certify") if headline_certify else ("certify", "batches")):   # headline loop last
        is_headline = name == ("certify" if headline_certify else "batches")
        if is_headline and sampler:
            sampler.start()                 # before the warm-up steps; samples count from start_headline() on
        hook = start_headline if is_headline else None
        if name == "batches":
            ms = timed(batch_step, args.steps, warm, hook)
            total = per_rank * world * args.steps
        else:
            ms = timed(certify_step, args.steps, warm if is_headline else 1,
                       (lambda: (mark_draws(), start_headline())) if is_headline else mark_draws)
            total = smooth._draws - snap["draws"]            # global MC samples classified (N0 + n per certified identity)
        if is_headline:
            launches = lib.cfr_launch_count() - snap["launches"]
            clocks = sampler.stop() if sampler else None
            # per-kernel roofline pass: the same K steps once more with the two-stream overlap OFF and a CUDA-event pair
            # around every conv launch (with the FRM side running concurrently on its own stream an event pair would
            # also time the wait for free SMs); the kernel's share is taken against THIS pass's step time
            eng.set_overlap(False)
            lib.cfr_profile_enable(1)
            ms_serial = timed(batch_step if name == "batches" else certify_step, args.steps, 1)
            prof = {}
            for kind, pname in ((0, "igemm"), (1, "halo")):
                t_ms, work, n_l = C.c_double(), C.c_double(), C.c_int64()
                L.check(lib.cfr_profile_read(kind, C.byref(t_ms), C.byref(work), C.byref(n_l)))
                prof[pname] = (t_ms.value, work.value, n_l.value)
            lib.cfr_profile_enable(0)
            eng.set_overlap(True)
        results[name] = {"ms": ms, "samples": total, "value": total / (ms * 1e-3)}
    # end to end through the public entry with HOST buffers
    if headline_certify:
        e2e_s = wall(lambda i: certify_step(i, host=True), args.steps, 1, mark_draws)
        e2e_total = smooth._draws - snap["draws"]
        e2e_api = ("Smooth.certify (drop-in Python API): latent from pinned host memory per identity, vote counts read back "
                   "to the host after each of the two all-reduced passes")
        h2d, d2h = G * 512 * 4, 2 * G * N_GALLERY * 8
    else:
        e2e_s = wall(batch_step_host, args.steps, 2)
        e2e_total = per_rank * world * args.steps
        e2e_api = "cfr_sample_votes_host (C ABI, host buffers, one call per step)"
        h2d, d2h = 528 * 4, N_GALLERY * 8
        assert int(c_h.sum()) == per_rank
    e2e_value = e2e_total / e2e_s
    head = results["certify" if headline_certify else "batches"]
    ms, value = head["ms"], head["value"]
    lt = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(lt)
    launches = int(lt.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    # DRAM bytes per launch of the halo kernel from the committed `ncu --set full` capture of this command (it cannot be
    # measured in-process; the file names the command, chunk size and build it was taken at, and is only used for that
    # chunk size -- a kernel change makes it stale, which `traffic_source` lets a reader check)
    halo_traffic, traffic_source = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r02_halo.json")) as fh:
            tj = json.load(fh)
        if tj.get("chunk") == args.chunk:
            halo_traffic = tj["avg_dram_bytes_per_launch"] / 1e6
            traffic_source = "profiles/traffic_r02_halo.json: " + tj.get("source", "")
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
                 "traffic_source": traffic_source,
                 "peak_source": pk["source"],
                 "launches_timed": int(ha_n), "avg_launch_ms": ha_ms / max(1, ha_n),
                 "alg_mbytes_per_launch": ha_bytes / max(1, ha_n) / 1e6,
                 "share_of_step": ha_ms / ms_serial if ms_serial > 0 else None}
    igemm_roof = {"kernel": "conv_igemm_kernel (tcgen05 implicit GEMM: StyleGAN layers 1-12 + every iresnet50 conv / FC)",
                  "bound": "tensor", "achieved": ig_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                  "frac": ig_tf / pk["tf_sustained"], "traffic": None, "peak_source": pk["source"] + " (sustained)",
                  "launches_timed": int(ig_n), "avg_launch_ms": ig_ms / max(1, ig_n),
                  "alg_gflop_per_launch": ig_flops / max(1, ig_n) / 1e9,
                  "share_of_step": ig_ms / ms_serial if ms_serial > 0 else None}
    dominant, other = (halo_roof, igemm_roof) if ha_ms >= ig_ms else (igemm_roof, halo_roof)
    if headline_certify:
        workload = (f"anisotropic certify (BASELINE config 3): {G} identities per step (Smooth.certify_many), each N0={N0} + "
                    f"n={NEST} MC samples split over {world} ranks by global sample index, sigma = {SIGMA} * eps^2 along the five "
                    f"W-boundaries, NCCL int64 all-reduce of the [{N_GALLERY}] vote counts after each pass; StyleGAN-FFHQ-1024 + "
                    f"ArcFace iresnet50 random-init, {N_GALLERY}-row synthetic gallery")
    second = results["batches" if headline_certify else "certify"]
    line = {
        "metric": "MC samples/sec (StyleGAN1024->ArcFace vote)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": warm, "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "strong" if (headline_certify or not ident_mode) else "weak", "vs_baseline": None,
        "dtype": "fp16 operands / fp32 accumulate (StyleGAN layers 1-%d: fp16 hi/lo split operands, fp32 activations)"
                 % eng.hp_layers,
        "data": "synthetic",
        "config": {"workload": workload, "chunk": args.chunk, "frm_group": args.frm_group,
                   "shard": "samples" if headline_certify else args.shard, "gallery": N_GALLERY,
                   "hp_layers": eng.hp_layers,
                   "overlap": "FRM + match + vote of group i run on a second stream under the synthesis of group i+1",
                   "l2": "working set per step (activations, GBs) far exceeds the 126 MB L2; no flush needed",
                   "gflop_per_sample_algorithmic": gflop_per_sample,
                   "pipeline_tflops": value * gflop_per_sample / 1e3,
                   "pipeline_frac_of_bf16_sustained": value * gflop_per_sample / 1e3 / (pk["tf_sustained"] * world)},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": e2e_api},
        "gpu_launches": int(launches),
        "roofline": dominant,
        "roofline_second_kernel": other,
        "roofline_pass": {"what": "per-launch CUDA-event times of the two conv kernels, taken in a second pass of the same K "
                                  "steps with the two-stream overlap off (serial launches); share_of_step is against this pass",
                          "ms_per_step_serial": ms_serial / args.steps},
        "clocks": clocks,
        # both partitions of SURVEY.md section 8e are timed in every run; the headline is (B) for N > 1, (A) for N = 1
        "samples_sharded": {"what": "BASELINE config 3: whole certifications (N0=100 + n=1000), samples of one identity "
                                    "split over the ranks, two int64 all-reduces per identity; device-timed",
                            "value": results["certify"]["value"], "unit": "samples/s",
                            "ms_per_identity": results["certify"]["ms"] / args.steps / G, "identities_per_step": G,
                            "identities_certified": cert_stat["certified"], "certify_calls": cert_stat["calls"]},
        "identities_sharded": {"what": f"BASELINE config 2: batches of {args.batch} samples, one identity per rank and step, "
                                       "no data-path collective; device-timed",
                               "value": results["batches"]["value"], "unit": "samples/s",
                               "ms_per_step": results["batches"]["ms"] / args.steps},
    }
    if not args.no_cpu_baseline and world == 1:
        val, dt, cores, kind, _ = cpu_arm(2, 1, args.cpu_sample)
        line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": cores, "kind": kind,
                                "sample": f"{args.cpu_sample} samples/step x 2 steps (+1 warm-up) of the same workload, "
                                          + ("the unmodified reference (oracle/_ref) " if kind == "reference" else
                                             "oracle/mc_path.py ") + f"torch fp32 on {cores} threads"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
