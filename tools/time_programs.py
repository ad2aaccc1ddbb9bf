import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from certifyingfacerecognition_b200 import synthetic as fixtures
from certifyingfacerecognition_b200.engine import Engine
g_sd, f_sd = fixtures.build_models()
dirs = torch.from_numpy(np.load("tests/golden/dirs.npy"))
for chunk in (250, 125):
    eng = Engine(g_sd, f_sd, dirs, torch.zeros(8, 512), chunk=chunk)
    eng.embed_latents(torch.from_numpy(fixtures.latents(chunk)))
    torch.cuda.synchronize()
    for prog, name in ((eng.frm, "frm"), (eng.synth, "synth")):
        for _ in range(2): prog.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): prog.run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"chunk {chunk} {name}: {ms:.3f} ms per run = {ms*1e3/chunk:.2f} us/sample, {prog.num_launches} launches")
    del eng; torch.cuda.empty_cache()
