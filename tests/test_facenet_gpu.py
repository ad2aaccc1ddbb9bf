"""BASELINE config 4: FaceNet (InceptionResnetV1 at 160^2) as the FRM.  PARITY UNPINNED -- facenet_pytorch is not
available, so the CUDA program is checked against oracle/facenet.py, an independent torch fp32
restatement of the published architecture, on seeded synthetic weights."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(golden, models):
    from certifyingfacerecognition_b200 import synthetic
    from certifyingfacerecognition_b200.engine import Engine
    g_sd, _ = models
    f_sd = synthetic.facenet_weights()
    dirs = torch.from_numpy(golden["dirs"])
    eng = Engine(g_sd, f_sd, dirs, torch.zeros(8, 512), chunk=4, frm="facenet-vggface2", keep_planar=True)
    return eng, g_sd, f_sd, dirs


def test_facenet_program_matches_oracle_on_identical_images(setup):
    """The FRM alone: the oracle's own 160^2 images (rounded to fp16) through the CUDA program -> cosine >= 0.9999."""
    from certifyingfacerecognition_b200.models.facenet import FaceNetProgram
    from oracle import facenet as O
    from oracle import fixtures, mc_path as M
    _, g_sd, f_sd, _ = setup
    n = 4
    w = torch.from_numpy(fixtures.latents(n))
    with torch.no_grad():
        img = M.transform(M.easy_synthesize(w, g_sd, literal=False), size=O.INPUT_RES)
        ref = O.forward(img, f_sd)
    buf = torch.zeros(n, 160, 160, 16, dtype=torch.float16, device="cuda")
    buf[..., :3] = img.permute(0, 2, 3, 1).cuda().half()
    prog = FaceNetProgram(f_sd, n, buf)
    prog.run()
    torch.cuda.synchronize()
    emb = prog.emb.cpu()
    assert torch.allclose(emb.norm(dim=1), torch.ones(n), atol=1e-4)          # F.normalize
    cos = F.cosine_similarity(emb, ref)
    assert cos.min().item() >= 0.9999, cos


def test_facenet_embeddings_match_oracle_end_to_end(setup):
    """StyleGAN (fp16 tensor-core pipeline) -> 160^2 -> FaceNet vs the all-fp32 oracle.  The random-init InceptionResnetV1
    is far more sensitive to pixel noise than the iresnet50 fixture (its program alone reproduces the oracle to 1.0000
    on identical images, see above), so the end-to-end bar here is cosine >= 0.99 and -- what matters for the votes --
    an error that is small against the distance to the nearest other identity."""
    from oracle import facenet as O
    from oracle import fixtures, mc_path as M
    eng, g_sd, f_sd, _ = setup
    w = torch.from_numpy(fixtures.latents(6))
    emb = eng.embed_latents(w).cpu()
    with torch.no_grad():
        img = M.transform(M.easy_synthesize(w, g_sd, literal=False), size=O.INPUT_RES)
        ref = O.forward(img, f_sd)
    cos = F.cosine_similarity(emb, ref)
    assert cos.min().item() >= 0.99, cos
    d_self = (emb - ref).norm(dim=1)
    d_other = torch.cdist(ref, ref) + 10 * torch.eye(6)
    assert (d_self < 0.25 * d_other.min(dim=1).values).all(), (d_self, d_other.min(dim=1).values)


def test_facenet_votes_on_its_own_gallery(setup):
    """End to end with the FaceNet FRM: gallery = embeddings of the identities; small perturbations vote for the identity."""
    from oracle import fixtures
    eng, _, _, _ = setup
    w = torch.from_numpy(fixtures.latents(8))
    eng.set_gallery(eng.embed_latents(w))
    counts, ex = eng.sample_votes(w[3:4], torch.zeros(1, 5), torch.tensor([0.02]), 12, seed=5, want_pred=True)
    torch.cuda.synchronize()
    assert counts.sum().item() == 12 and counts[3].item() == 12


def test_wrapped_model_accepts_facenet_names(setup, golden):
    from certifyingfacerecognition_b200 import synthetic
    from certifyingfacerecognition_b200.models.smoothing_model import WrappedModel
    from oracle import fixtures
    _, g_sd, f_sd, dirs = setup
    lat = torch.from_numpy(fixtures.latents(4))
    m = WrappedModel(dirs.cuda(), "facenet", n_embs=-1, load_embs=False, generator_state=g_sd, frm_state=f_sd, latents=lat,
                     chunk=4)
    assert m.orig_embs.shape == (4, 512)
    probs = m(m.latents[1:2], torch.zeros(1, 1, 1, 5, device="cuda"))
    assert probs.shape == (1, 4) and int(probs.argmax(1)) == 1
