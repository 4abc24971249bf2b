"""Run-to-run determinism of the pipelined kernels.  Every kernel on the path is written to
be deterministic (fixed tile -> CTA assignment, fixed accumulation order, top-k keys totally
ordered), so repeated runs on the same input must agree bit for bit; a missing barrier or a
stage reused too early in the TMA / mbarrier / TMEM pipelines shows up here as a sporadic
mismatch even when the error is too small for a tolerance test to notice.  Sizes are chosen
so that every CTA (pair) walks several tiles and the rings wrap many times."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

REPEATS = 8


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyvisim_b200 import encoders, features, retrieval
    from pyvisim_b200.encoders._base_encoder import kmeans_from_centers
    import types
    return types.SimpleNamespace(enc=encoders, feat=features, ret=retrieval, km=kmeans_from_centers)


def ragged_offsets(rng, n_images, t_lo, t_hi):
    t = rng.integers(t_lo, t_hi + 1, n_images)
    offs = np.zeros(n_images + 1, np.int64)
    np.cumsum(t, out=offs[1:])
    return torch.from_numpy(offs)


def test_fv_tensor_path_is_deterministic(mods):
    rng = np.random.default_rng(1)
    offs = ragged_offsets(rng, 700, 1, 3000)                      # more images than one chunk of 592, ragged
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((int(offs[-1]), 128), device="cuda", generator=g).abs_().mul_(40).clamp_(0, 255).floor_()
    enc = mods.enc.FisherVectorEncoder(feature_extractor=mods.feat.Descriptors(128),
                                       weights=mods.enc.GMMWeights.OXFORD102_K256_SIFT_PCA)
    first = enc.encode_descriptors(x, offs).clone()
    assert torch.isfinite(first).all()
    assert torch.allclose(first.norm(dim=1), torch.ones(700, device="cuda"), atol=1e-4)
    for _ in range(REPEATS):
        assert torch.equal(enc.encode_descriptors(x, offs), first)
    # one stream or two, one big call or many small ones: same bits
    assert torch.equal(enc.encode_descriptors(x, offs, n_streams=1), first)
    assert torch.equal(enc.encode_descriptors(x, offs, images_per_call=97), first)


@pytest.mark.parametrize("d,t_hi,n_images", [(514, 300, 1500), (128, 2500, 300), (64, 800, 600)])
def test_vlad_tensor_path_is_deterministic(mods, d, t_hi, n_images):
    rng = np.random.default_rng(d)
    offs = ragged_offsets(rng, n_images, 0, t_hi)                # includes images without descriptors
    g = torch.Generator(device="cuda").manual_seed(d)
    x = torch.randn((int(offs[-1]), d), device="cuda", generator=g)
    centers = x[torch.randperm(x.shape[0], device="cuda", generator=g)[:256]].cpu().numpy() + 0.05
    enc = mods.enc.VLADEncoder(feature_extractor=mods.feat.Descriptors(d), kmeans_model=mods.km(centers))
    first, labels = enc.encode_descriptors(x, offs, return_labels=True)
    first, labels = first.clone(), labels.clone()
    assert torch.isfinite(first).all() and int(labels.min()) >= 0 and int(labels.max()) < 256
    for _ in range(REPEATS):
        out, lab = enc.encode_descriptors(x, offs, return_labels=True)
        assert torch.equal(lab, labels) and torch.equal(out, first)
    assert torch.equal(enc.encode_descriptors(x, offs, images_per_call=211, n_streams=1), first)


@pytest.mark.parametrize("nq,ndb,d,k", [(4096, 8192, 1024, 100), (1000, 20000, 512, 10), (300, 70000, 256, 300)])
def test_similarity_topk_is_deterministic(mods, nq, ndb, d, k):
    g = torch.Generator(device="cuda").manual_seed(nq)
    db = mods.ret.l2_normalize(torch.randn((ndb, d), device="cuda", generator=g), "bf16")
    q = db[:nq]
    s0, i0 = mods.ret.cosine_topk(q, db, k)
    s0, i0 = s0.clone(), i0.clone()
    assert torch.all(i0[:, 0] == torch.arange(nq, device="cuda"))          # every row finds itself first
    assert torch.all(s0[:, :-1] >= s0[:, 1:])
    for _ in range(REPEATS):
        s, i = mods.ret.cosine_topk(q, db, k)
        assert torch.equal(i, i0) and torch.equal(s, s0)
