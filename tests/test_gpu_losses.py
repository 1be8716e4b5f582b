"""GPU suite (-m gpu): the loss-side callers (tcam_wsol_video_b200/losses.py) against the oracle."""
import numpy as np
import pytest

from conftest import rel_err
from tcam_wsol_video_b200 import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def test_con_ran_field_tcams(torch_cuda, oracle_mod):
    torch = torch_cuda
    from tcam_wsol_video_b200.losses import ConRanFieldTcams
    n, k, h, w = 4, 2, 48, 56
    raw = torch.from_numpy(synth.make_images(n, h, w, "natural", seed=3))            # CPU, 0..255
    logits = torch.randn((n, k, h, w), generator=torch.Generator().manual_seed(3)).cuda().requires_grad_(True)
    mod = ConRanFieldTcams(cuda_id=0, lambda_=2e-9, sigma_rgb=15., sigma_xy=100., scale_factor=1.0,
                           start_epoch=2, end_epoch=-1)
    assert mod.__name__ == "con_ran_field_tcams"
    off = mod(epoch=1, fcams=logits, raw_img=raw)
    assert off.item() == 0.0                                                           # before start_epoch
    loss = mod(epoch=2, fcams=logits, raw_img=raw)
    loss.backward()
    probs = torch.softmax(logits.detach().cpu(), dim=1)
    want, want_grad, _ = oracle_mod.densecrf_loss_fwd_bwd(raw.numpy(), probs.numpy(), 15., 100., 2e-9,
                                                          oracle_mod.port_bilateralfilter_batch)
    assert abs(loss.item() - 2e-9 * float(want)) < REL_TOL * abs(2e-9 * float(want))
    # chain rule through the softmax, done by torch on both sides
    p = torch.softmax(logits.detach().cpu().requires_grad_(True), dim=1)
    ref_logits = logits.detach().cpu().requires_grad_(True)
    torch.softmax(ref_logits, dim=1).backward(torch.from_numpy(want_grad))
    assert rel_err(logits.grad.cpu().numpy(), ref_logits.grad.numpy()) < REL_TOL
    # single-channel map -> (1-sigmoid, sigmoid)
    one = torch.randn((n, 1, h, w), generator=torch.Generator().manual_seed(4)).cuda()
    l1 = mod(epoch=5, fcams=one, raw_img=raw)
    s = torch.sigmoid(one.cpu())
    w1, _, _ = oracle_mod.densecrf_loss_fwd_bwd(raw.numpy(), torch.cat((1 - s, s), 1).numpy(), 15., 100., 1.0,
                                                oracle_mod.port_bilateralfilter_batch)
    assert abs(l1.item() - 2e-9 * float(w1)) < REL_TOL * abs(2e-9 * float(w1))


def test_rgb_joint_con_ran_field_tcams(torch_cuda, oracle_mod):
    """Clips of 3 and 2 frames plus a singleton (skipped): batched per-clip lattices == the reference's
    per-clip loop over width-concatenated frames, averaged over clips (tcam.py:191-205)."""
    torch = torch_cuda
    from tcam_wsol_video_b200.losses import RgbJointConRanFieldTcams, group_ordered_frames
    h, w, k = 24, 32, 2
    seq = torch.tensor([7, 3, 7, 3, 7, 9, 3, 4, 4])
    frm = torch.tensor([2, 0, 0, 2, 1, 5, 1, 1, 0])
    n = len(seq)
    raw = torch.from_numpy(synth.make_images(n, h, w, "natural", seed=8))
    logits = torch.randn((n, k, h, w), generator=torch.Generator().manual_seed(8)).cuda().requires_grad_(True)
    groups = group_ordered_frames(seq, frm)
    assert groups == [[1, 6, 3], [8, 7], [2, 4, 0], [5]]
    mod = RgbJointConRanFieldTcams(cuda_id=0, lambda_=1e-6, sigma_rgb=15., scale_factor=1.0)
    loss = mod(epoch=0, fcams=logits, raw_img=raw, seq_iter=seq, frm_iter=frm)
    loss.backward()
    probs = torch.softmax(logits.detach().cpu(), dim=1).numpy()
    total, grads = 0.0, np.zeros_like(probs)
    clips = [g for g in groups if len(g) >= 2]
    for g in clips:
        img = np.concatenate([raw.numpy()[i] for i in g], axis=2)[None]      # width concat
        seg = np.concatenate([probs[i] for i in g], axis=2)[None]
        l, gr, _ = oracle_mod.color_densecrf_loss_fwd_bwd(img, seg, 15., 1e-6 / len(clips),
                                                          oracle_mod.port_colorbilateralfilter_batch)
        total += 1e-6 * float(l) / len(clips)
        for j, i in enumerate(g):
            grads[i] = gr[0][:, :, j * w:(j + 1) * w]
    assert abs(loss.item() - total) < REL_TOL * abs(total)
    ref_logits = logits.detach().cpu().requires_grad_(True)
    torch.softmax(ref_logits, dim=1).backward(torch.from_numpy(grads))
    assert rel_err(logits.grad.cpu().numpy(), ref_logits.grad.numpy()) < REL_TOL
    # pair_samples keeps the reference signature
    pi, pc = mod.pair_samples([2, 4, 0], raw, torch.from_numpy(probs))
    assert pi.shape == (1, 3, h, 3 * w) and pc.shape == (1, k, h, 3 * w)
    assert torch.equal(pi[0, :, :, w:2 * w], raw[4])


def test_self_learning_tcams_consumes_seeds(torch_cuda):
    torch = torch_cuda
    from tcam_wsol_video_b200.losses import SelfLearningTcams
    from tcam_wsol_video_b200.tcam_seeding import TCAMSeeder
    b, h, w = 4, 64, 64
    cam = torch.rand((b, 1, h, w), generator=torch.Generator().manual_seed(2)).cuda()
    roi = (cam > 0.5).long()
    seeder = TCAMSeeder(seed_tech="seed_weighted", min_=1, max_=1, max_p=0.6, min_p=0.1, fg_erode_k=11,
                        fg_erode_iter=0, ksz=3, support_background=True, multi_label_flag=False,
                        seg_ignore_idx=-255, cuda_id=0, roi_method="roi_all", p_min_area_roi=0.05, use_roi=True)
    seeds = seeder(x=cam, roi=roi)
    fcams = torch.randn((b, 2, h, w), device="cuda", requires_grad=True)
    loss = SelfLearningTcams(cuda_id=0, lambda_=1.0, seg_ignore_idx=-255)(epoch=0, fcams=fcams, seeds=seeds)
    want = torch.nn.functional.cross_entropy(fcams, seeds, ignore_index=-255)
    assert torch.allclose(loss, want)
    loss.backward()
    assert int((seeds >= 0).sum()) <= b * 2 * 9 and fcams.grad.abs().sum() > 0


@pytest.mark.parametrize("k", [2, 10, 3])
def test_fused_softmax_matches_unfused(torch_cuda, k):
    """DenseCRFLossFromLogits == DenseCRFLoss o softmax, value and gradient w.r.t. the logits (rel 1e-4; the
    only difference is exp() rounding), with float and uint8 images."""
    torch = torch_cuda
    from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLoss, DenseCRFLossFromLogits
    from tcam_wsol_video_b200.losses import ConRanFieldTcams
    n, h, w = 3, 56, 64
    raw = torch.from_numpy(synth.make_images(n, h, w, "natural", seed=13))
    logits = (3 * torch.randn((n, k, h, w), generator=torch.Generator().manual_seed(13))).cuda()
    a = logits.clone().requires_grad_(True)
    b = logits.clone().requires_grad_(True)
    la = DenseCRFLoss(1e-6, 15., 100., 1.0)(images=raw, segmentations=torch.softmax(a, dim=1))
    lb = DenseCRFLossFromLogits(1e-6, 15., 100., 1.0)(images=raw.to(torch.uint8).cuda(), logits=b)
    la.backward()
    lb.backward()
    assert abs(la.item() - lb.item()) < REL_TOL * abs(la.item())
    assert rel_err(b.grad.cpu().numpy(), a.grad.cpu().numpy()) < REL_TOL
    c = logits.clone().requires_grad_(True)
    mod = ConRanFieldTcams(fuse_softmax=True, cuda_id=0, lambda_=1e-6, sigma_rgb=15., sigma_xy=100., scale_factor=1.0)
    lc = mod(epoch=0, fcams=c, raw_img=raw)
    lc.backward()
    assert abs(la.item() - lc.item()) < REL_TOL * abs(la.item())
    assert rel_err(c.grad.cpu().numpy(), a.grad.cpu().numpy()) < REL_TOL


def test_exact_gradient_option(torch_cuda, oracle_mod):
    """Opt-in exact gradient: -(A + A^T) S g / N.  Checked three ways: (1) the transposed filter really is the
    transpose, <V, A S> == <A^T V, S>; (2) the directional derivative of the (quadratic) loss along a random V
    equals <grad_exact, V>; (3) the default stays the reference's -2 g A S / N."""
    torch = torch_cuda
    from tcam_wsol_video_b200 import _lib, ops
    from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLoss
    n, k, h, w = 2, 3, 40, 48
    raw = torch.from_numpy(synth.make_images(n, h, w, "noise", seed=21))
    gen = torch.Generator().manual_seed(21)
    s = torch.softmax(torch.randn((n, k, h, w), generator=gen), dim=1).cuda()
    v = torch.randn((n, k, h, w), generator=gen).cuda()
    cfg = _lib.make_config(_lib.FEAT_XY_RGB, 3, 15.0, 100.0)
    a_s, _, _ = ops.crf_forward(raw, s, cfg, want_loss=False)
    at_v = ops.crf_filter_transposed(raw, v, cfg)
    lhs = (v.double() * a_s.double()).sum().item()
    rhs = (at_v.double() * s.double()).sum().item()
    assert abs(lhs - rhs) < 1e-5 * abs(lhs)
    # the filter is NOT symmetric: A^T V differs from A V
    a_v, _, _ = ops.crf_forward(raw, v, cfg, want_loss=False)
    assert rel_err(at_v.cpu().numpy(), a_v.cpu().numpy()) > 1e-4

    weight = 1e-3
    s1 = s.clone().requires_grad_(True)
    DenseCRFLoss(weight, 15.0, 100.0, 1.0, exact_gradient=True)(images=raw, segmentations=s1).backward()
    # L(S + eV) = L(S) + e * <grad, V> + e^2 * L_2: read the linear term off two evaluations
    def loss_at(t):
        return DenseCRFLoss(weight, 15.0, 100.0, 1.0)(images=raw, segmentations=(s + t * v)).double().item()
    eps = 1.0
    directional = (loss_at(eps) - loss_at(-eps)) / (2 * eps)
    got = (s1.grad.double() * v.double()).sum().item()
    assert abs(got - directional) < 2e-4 * abs(directional)
    s2 = s.clone().requires_grad_(True)
    DenseCRFLoss(weight, 15.0, 100.0, 1.0)(images=raw, segmentations=s2).backward()
    want = (-2.0 * weight * a_s / n)
    assert rel_err(s2.grad.cpu().numpy(), want.cpu().numpy()) < 1e-6


def test_lattice_reuse(torch_cuda, oracle_mod):
    """One lattice, many filters (tcamcrf_lattice_build / _apply): equals the one-shot filter and the oracle, can
    be applied repeatedly (the first splat turns entry indices into tagged vertex ids in place), A^T included."""
    torch = torch_cuda
    from tcam_wsol_video_b200 import _lib, ops
    n, k, h, w = 3, 5, 37, 52
    raw_np = synth.make_images(n, h, w, "noise", seed=5)
    raw = torch.from_numpy(raw_np)
    gen = torch.Generator().manual_seed(5)
    a = torch.rand((n, k, h, w), generator=gen).cuda()
    b = torch.randn((n, k, h, w), generator=gen).cuda()
    cfg = _lib.make_config(_lib.FEAT_XY_RGB, 3, 15.0, 100.0)
    lat = ops.Lattice(raw, cfg, k, device=a.device)
    st, m = lat.status()
    assert st == 0 and m > 0
    want_a = oracle_mod.port_bilateralfilter_batch(raw_np, a.cpu().numpy(), n, k, h, w, 15.0, 100.0).reshape(n, k, h, w)
    want_b = oracle_mod.port_bilateralfilter_batch(raw_np, b.cpu().numpy(), n, k, h, w, 15.0, 100.0).reshape(n, k, h, w)
    got_a1 = lat.apply(a)
    got_b = lat.apply(b)
    got_a2, loss = lat.apply(a, want_loss=True, n_norm=float(n))
    assert rel_err(got_a1.cpu().numpy(), want_a) < REL_TOL
    assert rel_err(got_b.cpu().numpy(), want_b) < REL_TOL
    assert rel_err(got_a2.cpu().numpy(), want_a) < REL_TOL
    want_loss = -(a.double().cpu().numpy() * want_a.astype(np.float64)).sum() / n
    assert abs(loss.item() - want_loss) < REL_TOL * abs(want_loss)
    # one-shot path on the same inputs
    one_shot, _, _ = ops.crf_forward(raw, a, cfg, want_loss=False)
    assert rel_err(got_a1.cpu().numpy(), one_shot.cpu().numpy()) < 1e-5
    # transposed on the kept lattice == transposed one-shot, and <b, A a> == <A^T b, a>
    at_b = lat.apply(b, transposed=True)
    assert rel_err(at_b.cpu().numpy(), ops.crf_filter_transposed(raw, b, cfg).cpu().numpy()) < 1e-5
    lhs = (b.double() * got_a1.double()).sum().item()
    rhs = (at_b.double() * a.double()).sum().item()
    assert abs(lhs - rhs) < 1e-5 * abs(lhs)
    with pytest.raises(_lib.TcamCrfError):
        lat.apply(a[:, :2])


@pytest.mark.parametrize("shape,quirk", [((2, 3, 32, 32), True), ((2, 2, 24, 40), True), ((1, 4, 30, 30), False)])
def test_dense_crf_filter_mean_field(torch_cuda, oracle_mod, shape, quirk):
    """DenseCRFFilter (mean-field refinement, crf_post_processing.py:33-135) against the numpy restatement of the
    pydensecrf calls built on the oracle's filter.  Tolerance: 1e-3 absolute on probabilities -- every iteration
    multiplies the filter's rel 1e-6 by compat=10 inside a softmax."""
    torch = torch_cuda
    from oracle import crf_post_processing as ocp
    from tcam_wsol_video_b200.crf_post_processing import DenseCRFFilter
    n, k, h, w = shape
    raw_np = synth.make_images(n, h, w, "natural", seed=9)
    gen = torch.Generator().manual_seed(9)
    seg = torch.softmax(2.0 * torch.randn((n, k, h, w), generator=gen), dim=1)
    itera = 3
    flt = DenseCRFFilter(sigma_rgb=15.7, sigma_xy=100.2, scale_factor=1.0, itera=itera, quirk_transposed_image=quirk)
    assert (flt.sigma_rgb, flt.sigma_xy) == (15, 100)
    got = flt(torch.from_numpy(raw_np), seg)                       # CPU in, CPU out like the reference
    assert got.device.type == "cpu" and got.shape == seg.shape
    for i in range(n):
        want = ocp.mean_field(raw_np[i], seg[i].numpy(), 15, 100, itera, oracle_mod.port_bilateralfilter_batch,
                              quirk=quirk)
        assert np.abs(got[i].numpy() - want).max() < 1e-3
        assert np.abs(got[i].numpy().sum(0) - 1.0).max() < 1e-5
    # CUDA in -> CUDA out, same numbers
    got_cuda = flt(torch.from_numpy(raw_np).cuda(), seg.cuda())
    assert got_cuda.is_cuda and np.abs(got_cuda.cpu().numpy() - got.numpy()).max() < 1e-4
    # refinement does something, and itera=0 returns the (rescaled) input
    assert np.abs(got.numpy() - seg.numpy()).max() > 1e-2
    same = DenseCRFFilter(15, 100, 1.0, 0)(torch.from_numpy(raw_np), seg)
    assert torch.equal(same, seg)


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable-frames", "pinned-frames"])
def test_module_matches_the_reference_autograd_function(torch_cuda, pinned):
    """DenseCRFLoss (CUDA) against loss and gradient produced by the reference's own DenseCRFLossFunction executed on
    CPU tensors with the reference's C++ behind it (tests/golden/py/py_dense_crf_loss.npz).  The frames stay on the
    CPU as in the reference's trainer; pinned ones take tcamcrf_loss_forward_host_frames."""
    import os
    torch = torch_cuda
    from conftest import GOLDEN
    from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLoss
    g = np.load(os.path.join(GOLDEN, "py", "py_dense_crf_loss.npz"))
    seg = torch.from_numpy(g["seg"]).cuda().requires_grad_(True)
    image = torch.from_numpy(g["image"])
    loss = DenseCRFLoss(float(g["weight"]), 15.0, 100.0, 1.0)(images=image.pin_memory() if pinned else image,
                                                              segmentations=seg)
    loss.backward()
    assert loss.shape == (1,)
    assert abs(loss.item() - float(g["loss"][0])) < REL_TOL * abs(float(g["loss"][0]))
    assert rel_err(seg.grad.cpu().numpy(), g["grad"]) < REL_TOL


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable-frames", "pinned-frames"])
def test_colour_module_matches_the_reference_autograd_function(torch_cuda, pinned):
    """ColorDenseCRFLoss (CUDA) against the reference's own ColorDenseCRFLossFunction executed on CPU tensors."""
    import os
    torch = torch_cuda
    from conftest import GOLDEN
    from tcam_wsol_video_b200.color_dense_crf_loss import ColorDenseCRFLoss
    g = np.load(os.path.join(GOLDEN, "py", "py_color_dense_crf_loss.npz"))
    seg = torch.from_numpy(g["seg"]).cuda().requires_grad_(True)
    image = torch.from_numpy(g["image"])
    loss = ColorDenseCRFLoss(float(g["weight"]), 15.0, 1.0)(images=image.pin_memory() if pinned else image,
                                                            segmentations=seg)
    loss.backward()
    assert abs(loss.item() - float(g["loss"][0])) < REL_TOL * abs(float(g["loss"][0]))
    assert rel_err(seg.grad.cpu().numpy(), g["grad"]) < REL_TOL


def test_tcam_step_in_a_cuda_graph(torch_cuda):
    """The TCAM loss step -- temporal max + seeding (rng_parity=False: counts and draws in the kernel), CRF from logits,
    cross-entropy on the seeds, backward -- captured with torch.cuda.graph on a stream the library has never seen (new
    workspace, no density hint, no host round trip anywhere) and replayed on new inputs: same CRF gradient as eager."""
    torch = torch_cuda
    from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLossFromLogits
    from tcam_wsol_video_b200.tcam_seeding import TCAMSeeder
    n, t, h, w = 4, 3, 64, 72
    dev = torch.device("cuda", 0)
    seeder = TCAMSeeder(seed_tech="seed_weighted", min_=1, max_=1, max_p=0.6, min_p=0.1, fg_erode_k=11, fg_erode_iter=0,
                        ksz=3, support_background=True, multi_label_flag=False, seg_ignore_idx=-255, cuda_id=0,
                        roi_method="roi_all", p_min_area_roi=0.05, use_roi=True, rng_parity=False)
    crf = DenseCRFLossFromLogits(1.0, 15.0, 100.0, 1.0)
    cams = torch.zeros((n, t, h, w), device=dev)
    roi = torch.ones((n, 1, h, w), dtype=torch.long, device=dev)
    img8 = torch.zeros((n, 3, h, w), dtype=torch.uint8, device=dev)
    logits = torch.zeros((n, 2, h, w), device=dev, requires_grad=True)
    seeds_out = torch.zeros((n, h, w), dtype=torch.long, device=dev)

    def load(seed):
        g = torch.Generator().manual_seed(seed)
        cams.copy_(torch.rand((n, t, h, w), generator=g))
        img8.copy_(torch.from_numpy(synth.make_images(n, h, w, "natural", seed=seed).astype(np.uint8)))
        with torch.no_grad():
            logits.copy_(torch.randn((n, 2, h, w), generator=g))

    def step():
        seeds, _ = seeder.forward_stack(cams, roi)
        seeds_out.copy_(seeds)
        l_crf = crf(images=img8, logits=logits)
        loss = l_crf + torch.nn.functional.cross_entropy(logits, seeds, ignore_index=-255)
        loss.backward()
        return l_crf

    load(1)
    graph = torch.cuda.CUDAGraph()
    logits.grad = None
    with torch.cuda.graph(graph):
        l_graph = step()
    for seed in (2, 3):
        load(seed)
        logits.grad.zero_()
        graph.replay()
        torch.cuda.synchronize()
        got_l = l_graph.item()
        s = seeds_out.clone()
        assert set(torch.unique(s).tolist()) <= {-255, 0, 1} and (s == 1).any() and (s == 0).any()
        got = logits.grad.clone()
        # eager: the same CRF term, and the cross-entropy on the seeds the replay picked
        z = logits.detach().clone().requires_grad_(True)
        l_e = crf(images=img8, logits=z)
        (l_e + torch.nn.functional.cross_entropy(z, s, ignore_index=-255)).backward()
        assert abs(got_l - l_e.item()) < 1e-5 * abs(l_e.item())
        assert rel_err(got.cpu().numpy(), z.grad.cpu().numpy()) < 1e-5


def _seeder_kw(**kw):
    base = dict(seed_tech="seed_weighted", min_=1, max_=1, max_p=0.6, min_p=0.1, fg_erode_k=11, fg_erode_iter=0,
                ksz=3, support_background=True, multi_label_flag=False, seg_ignore_idx=-255, cuda_id=0,
                roi_method="roi_all", p_min_area_roi=0.05, use_roi=True)
    base.update(kw)
    return base


@pytest.mark.parametrize("cfg,shape,k_classes", [
    (dict(), (4, 64, 72), 2),
    (dict(min_=3, max_=5, ksz=5), (3, 40, 44), 2),            # overlapping windows, fg/bg conflicts
    (dict(min_=6, max_=6, ksz=4, min_p=0.3, max_p=0.3), (2, 12, 14), 2),   # even window, seeds at the borders
    (dict(min_=2, max_=0, ksz=3), (3, 33, 35), 4),            # background seeds only, four channels
])
def test_sparse_seed_cross_entropy(torch_cuda, cfg, shape, k_classes):
    """SelfLearningTcams on SparseSeeds (the cross-entropy computed from the labelled pixels alone) against torch's
    CrossEntropyLoss on the label map the same seeds paint (dlib/losses/tcam.py:48-77): value and gradient."""
    torch = torch_cuda
    from tcam_wsol_video_b200.losses import SelfLearningTcams
    from tcam_wsol_video_b200.tcam_seeding import SparseSeeds, TCAMSeeder
    b, h, w = shape
    g = torch.Generator().manual_seed(b * 100 + h)
    low = torch.rand((b, 1, 7, 7), generator=g)
    cam = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False).cuda()
    cam[b - 1] = 0.5                                           # a flat CAM: no seeds for that sample
    roi = (cam >= cam.flatten(1).median(dim=1).values.view(b, 1, 1, 1)).long()
    seeder = TCAMSeeder(**_seeder_kw(**cfg))
    torch.manual_seed(7)
    dense = seeder(cam, roi)
    torch.manual_seed(7)
    sparse = seeder(cam, roi, sparse=True)
    assert isinstance(sparse, SparseSeeds) and torch.equal(sparse.dense(), dense)
    sl = SelfLearningTcams(cuda_id=0, lambda_=0.7)
    z1 = torch.randn((b, k_classes, h, w), generator=g).cuda().requires_grad_(True)
    z2 = z1.detach().clone().requires_grad_(True)
    l_sparse = sl(fcams=z1, seeds=sparse)
    l_dense = sl(fcams=z2, seeds=dense)
    l_sparse.backward()
    l_dense.backward()
    assert abs(l_sparse.item() - l_dense.item()) < 1e-6 * abs(l_dense.item())
    assert rel_err(z1.grad.cpu().numpy(), z2.grad.cpu().numpy()) < 1e-6
    # nothing labelled at all: NaN, like torch's mean over no element
    none = SparseSeeds(torch.full_like(sparse.sel, -1), sparse.ksz, -255, h, w)
    assert torch.isnan(sl(fcams=z1.detach(), seeds=none)).all()


def test_fused_tcam_losses(torch_cuda):
    """FusedTcamLosses: SelfLearningTcams + ConRanFieldTcams as one autograd node (sparse cross-entropy added in place
    to the CRF gradient).  Same total, same terms, same gradient as the two modules run separately on the label map;
    epoch gating of either term falls back to the separate modules."""
    torch = torch_cuda
    from tcam_wsol_video_b200.losses import ConRanFieldTcams, FusedTcamLosses, SelfLearningTcams
    from tcam_wsol_video_b200.tcam_seeding import TCAMSeeder
    b, h, w = 5, 48, 56
    g = torch.Generator().manual_seed(3)
    low = torch.rand((b, 3, 7, 7), generator=g)
    cams = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False).cuda()
    roi = torch.ones((b, 1, h, w), dtype=torch.long, device="cuda")
    raw = torch.from_numpy(synth.make_images(b, h, w, "natural", seed=8).astype(np.uint8)).cuda()
    seeder = TCAMSeeder(**_seeder_kw(min_=2, max_=2, ksz=3, rng_parity=False))
    seeds, _ = seeder.forward_stack(cams, roi, sparse=True)
    sl = SelfLearningTcams(cuda_id=0, lambda_=1.0)
    crf = ConRanFieldTcams(cuda_id=0, lambda_=2e-9, sigma_rgb=15., sigma_xy=100., scale_factor=1.0, fuse_softmax=True)
    fused = FusedTcamLosses(sl, crf)
    z1 = torch.randn((b, 2, h, w), generator=g).cuda().requires_grad_(True)
    z2 = z1.detach().clone().requires_grad_(True)
    total = fused(epoch=0, fcams=z1, raw_img=raw, seeds=seeds)
    total.backward()
    want_sl = sl(epoch=0, fcams=z2, seeds=seeds.dense())
    want_crf = crf(epoch=0, fcams=z2, raw_img=raw)
    (want_sl + want_crf).backward()
    assert abs(total.item() - (want_sl + want_crf).item()) < 1e-6 * abs((want_sl + want_crf).item())
    assert abs(fused.last_ce.item() * sl.lambda_ - want_sl.item()) < 1e-6 * abs(want_sl.item())
    assert abs(fused.last_crf.item() - want_crf.item()) < 1e-5 * abs(want_crf.item())
    assert rel_err(z1.grad.cpu().numpy(), z2.grad.cpu().numpy()) < 1e-5
    # the CRF term switched off by its epoch window: the fused module returns the cross-entropy alone
    crf_late = ConRanFieldTcams(cuda_id=0, lambda_=2e-9, sigma_rgb=15., sigma_xy=100., scale_factor=1.0,
                                fuse_softmax=True, start_epoch=5, end_epoch=9)
    only_sl = FusedTcamLosses(sl, crf_late)(epoch=0, fcams=z1.detach(), raw_img=raw, seeds=seeds)
    assert abs(only_sl.item() - want_sl.item()) < 1e-6 * abs(want_sl.item())
