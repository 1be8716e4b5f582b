"""Generates tests/golden/py/*.npz by EXECUTING the reference's own Python for the pure-torch / pure-python pieces
around the hot path (no GPU, no third-party packages needed).  Run in the build container only:

    python tests/golden/make_golden_py.py

The reference package cannot be imported here (dlib/__init__ chains need pydensecrf, kornia, skimage, munch ...), so
the function bodies are cut out of the reference files with `ast` -- from where they lie under /root/reference,
nothing is copied into this repository -- compiled and run as they are:

  * _WSOLDataset.re_normalize_cam, _get_lef_knn, _get_right_knn   dlib/datasets/wsol_loader.py:448-459, 630-635
  * the torch.maximum chain of __getitem__                         dlib/datasets/wsol_loader.py:591-600 (restated:
                                                                   three lines inside a 150-line method)
  * Trainer.prepare_std_cams_disq                                  dlib/learning/train_wsol.py:417-432
  * _SFG / _SBG (fg / bg seed sampling modules, whole classes)      dlib/cams/tcam_seeding.py:490-592
  * TCAMSeeder + _OneSample (whole classes; kornia's dilation and  dlib/cams/tcam_seeding.py:44-260, 433-488
    torch.device(cuda_id) stubbed, see below)
  * DenseCRFLossFunction (whole class, forward + backward)          dlib/crf/dense_crf_loss.py:30-74, with the
                                                                   reference's own C++ (oracle/_ref) behind the
                                                                   bilateralfilter_batch name instead of the SWIG module

Stored: inputs and the reference functions' outputs.  tests/test_oracle.py checks the oracle restatements against
them (CPU), tests/test_gpu_seeding.py the kernels (GPU box, where /root/reference does not exist).
"""
import ast
import os
import sys
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tcam_wsol_video_b200 import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "py")


def cut(path, cls, name):
    """Source of method `name` of class `cls` in the reference file `path`, dedented, decorators dropped."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == name:
                    lines = src.splitlines()[item.lineno - 1:item.end_lineno]
                    return textwrap.dedent("\n".join(lines))
    raise KeyError((path, cls, name))


def cut_class(path, cls):
    """Source of a whole top-level class of the reference file `path`."""
    src = open(os.path.join(REF, path)).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            return "\n".join(src.splitlines()[node.lineno - 1:node.end_lineno])
    raise KeyError((path, cls))


def load_class(path, cls, env):
    scope = dict(env)
    exec(compile(cut_class(path, cls), f"{path}:{cls}", "exec"), scope)
    return scope[cls]


def load(path, cls, name, env):
    code = cut(path, cls, name)
    scope = dict(env)
    exec(compile(code, f"{path}:{cls}.{name}", "exec"), scope)
    return scope[name]


def main():
    from typing import Tuple
    os.makedirs(OUT, exist_ok=True)
    env = {"torch": torch, "F": F, "Tuple": Tuple}
    loader_cls = "WSOLImageLabelDataset"
    re_norm = load("dlib/datasets/wsol_loader.py", loader_cls, "re_normalize_cam", env)
    left = load("dlib/datasets/wsol_loader.py", loader_cls, "_get_lef_knn", env)
    right = load("dlib/datasets/wsol_loader.py", loader_cls, "_get_right_knn", env)
    prep = load("dlib/learning/train_wsol.py", "Trainer", "prepare_std_cams_disq", env)

    # --- temporal aggregation: B samples x T frames of low-resolution CAMs, with the odd values the loader guards
    low = torch.from_numpy(synth.make_low_res_cams(6, 5, 28, 28, seed=11))          # [B,T,1,h,w]
    low[1, 2, 0, 3, 4] = float("nan")
    low[2, 0, 0, 0, 0] = float("inf")
    low[3, 1, 0, 5, 5] = float("-inf")
    out = {"cams": low.numpy()}
    for h_t in (0.0, 10.0, 50.0):
        agg = []
        for b in range(low.shape[0]):
            std = None
            for t in range(low.shape[1]):
                c = low[b, t]
                if h_t > 0:
                    c = re_norm(c, h=h_t)                                       # wsol_loader.py:594-595
                std = c if std is None else torch.maximum(std, c)               # wsol_loader.py:597-600
            agg.append(std)
        out[f"agg_h{int(h_t)}"] = torch.stack(agg).numpy()
    out["renorm_single_h10"] = re_norm(low[0, 0], h=10.0).numpy()
    np.savez_compressed(os.path.join(OUT, "py_temporal_agg.npz"), **out)

    # --- prepare_std_cams_disq (self is unused by the method body)
    std = torch.from_numpy(synth.make_low_res_cams(3, 1, 28, 28, seed=4))[:, 0]       # [B,1,h,w]
    std[0, 0, 2, 3] = float("nan")
    std[1, 0, 0, 0] = float("inf")
    std[2, 0, 27, 27] = float("-inf")
    res = {"std_cams": std.numpy()}
    for size in ((224, 224), (96, 160), (28, 28)):
        res[f"out_{size[0]}x{size[1]}"] = prep(None, std, size).numpy()
    np.savez_compressed(os.path.join(OUT, "py_prepare_std_cams.npz"), **res)

    # --- frame pickers
    frames = [f"shot0_{i:03d}.jpg" for i in range(7)]
    pick = {}
    for k in (1, 2, 4):
        for f in (0, 1, 3, 5, 6):
            pick[f"left_k{k}_f{f}"] = np.array(left(frames, frames[f], k), dtype=object).astype(str)
            pick[f"right_k{k}_f{f}"] = np.array(right(frames, frames[f], k), dtype=object).astype(str)
    np.savez_compressed(os.path.join(OUT, "py_frame_pickers.npz"), frames=np.array(frames), **pick)
    # --- fg / bg seed sampling: the reference's own _SFG / _SBG modules (dlib/cams/tcam_seeding.py:490-592) on the
    #     CPU generator; oracle/seeding.py must reproduce them draw for draw
    import types
    import torch.nn as nn
    consts = types.SimpleNamespace(SEED_UNIFORM='seed_uniform', SEED_WEIGHTED='seed_weighted')
    senv = {"torch": torch, "nn": nn, "constants": consts}
    SFG = load_class("dlib/cams/tcam_seeding.py", "_SFG", senv)
    SBG = load_class("dlib/cams/tcam_seeding.py", "_SBG", senv)
    g = torch.Generator().manual_seed(77)
    lowc = torch.rand((4, 1, 7, 9), generator=g)
    cam = F.interpolate(lowc, size=(48, 56), mode="bilinear", align_corners=False)[:, 0]
    cam[3] = torch.round(cam[3] * 8) / 8                                      # many ties: stable-sort order matters
    roi = (cam >= cam.flatten(1).median(dim=1).values.view(4, 1, 1)).long()
    seeds = {"cam": cam.numpy(), "roi": roi.numpy()}
    for ci, (tech, max_, min_, max_p, min_p, use_roi) in enumerate([
            ('seed_weighted', 1, 1, 0.6, 0.1, True), ('seed_uniform', 1, 1, 0.6, 0.1, True),
            ('seed_weighted', 5, 3, 0.2, 0.2, False), ('seed_uniform', 10, 10, 0.2, 0.2, True)]):
        fgm, bgm = SFG(max_p=max_p, max_=max_, seed_tech=tech), SBG(min_p=min_p, min_=min_, seed_tech='seed_uniform')
        torch.manual_seed(1000 + ci)
        fgs, bgs = [], []
        for i in range(cam.shape[0]):                                          # TCAMSeeder's per-sample order
            fg0 = torch.zeros((48, 56), dtype=torch.long)
            bg0 = torch.zeros((48, 56), dtype=torch.long)
            fgs.append(fgm(cam=cam[i], roi=roi[i] if use_roi else None, fg=fg0))
            bgs.append(bgm(cam=cam[i], bg=bg0))
        seeds[f"case{ci}_fg"] = torch.stack(fgs).numpy()
        seeds[f"case{ci}_bg"] = torch.stack(bgs).numpy()
        seeds[f"case{ci}_cfg"] = np.array([tech, str(max_), str(min_), str(max_p), str(min_p), str(int(use_roi))])
    np.savez_compressed(os.path.join(OUT, "py_seed_sampling.npz"), **seeds)
    # --- the whole TCAMSeeder.forward (dlib/cams/tcam_seeding.py:44-260 + _OneSample :433-488) on the CPU: per-sample
    #     loop, dilation of the seed maps, fg/bg conflict cancellation, ignore index.  Two things are stubbed and said
    #     so: kornia's dilation / erosion (not installed) by the flat max-pool of oracle/seeding.py, and
    #     torch.device(cuda_id) by the CPU device (no GPU in the build container).  roi is always passed, so the
    #     skimage-based GetRoiSingleCam / STOtsu members are constructed as dummies and never called.
    from oracle import seeding as oseed

    class _TorchOnCpu:
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def device(*_a, **_k):
            return torch.device("cpu")

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    from typing import Callable
    consts2 = types.SimpleNamespace(SEED_UNIFORM='seed_uniform', SEED_WEIGHTED='seed_weighted',
                                    SEED_TECHS=['seed_uniform', 'seed_weighted'],
                                    ROI_SELECT=['roi_all', 'roi_high_density', 'largest'])
    tenv = {"torch": _TorchOnCpu(), "nn": nn, "constants": consts2, "Tuple": Tuple, "Callable": Callable,
            "STOtsu": _Dummy, "GetRoiSingleCam": _Dummy,
            "dilation": lambda x, kernel: oseed.flat_dilation(x, kernel.shape[0]),
            "erosion": lambda x, kernel: -oseed.flat_dilation(-x, kernel.shape[0])}
    for cname in ("_SFG", "_SBG", "_OneSample", "TCAMSeeder"):
        exec(compile(cut_class("dlib/cams/tcam_seeding.py", cname), f"tcam_seeding.py:{cname}", "exec"), tenv)
    full = {"cam": cam.numpy(), "roi": roi.numpy()}
    for ci, kw in enumerate([
            dict(seed_tech='seed_weighted', min_=1, max_=1, min_p=0.1, max_p=0.6, ksz=3, use_roi=True),
            dict(seed_tech='seed_uniform', min_=10, max_=10, min_p=0.2, max_p=0.2, ksz=1, use_roi=False),
            dict(seed_tech='seed_weighted', min_=4, max_=4, min_p=0.3, max_p=0.3, ksz=5, use_roi=True)]):
        mod = tenv["TCAMSeeder"](fg_erode_k=11, fg_erode_iter=0, support_background=True, multi_label_flag=False,
                                 seg_ignore_idx=-255, cuda_id=0, roi_method='roi_all', p_min_area_roi=0.05, **kw)
        torch.manual_seed(2000 + ci)
        full[f"case{ci}_out"] = mod(cam.unsqueeze(1), roi.unsqueeze(1)).numpy()
        full[f"case{ci}_cfg"] = np.array([f"{k}={v}" for k, v in kw.items()])
    np.savez_compressed(os.path.join(OUT, "py_tcam_seeder.npz"), **full)

    # --- DenseCRFLossFunction itself (dlib/crf/dense_crf_loss.py:30-74), run on CPU tensors: the SWIG module is
    #     replaced by the reference's own C++ built in place (oracle/_ref), torch.cuda.synchronize() by a no-op (there
    #     is no GPU in the build container), the AMP decorators by the identity
    import oracle
    if oracle.have_ref():
        def bilateralfilter_batch(images, ins, outs, N, K, H, W, sigma_rgb, sigma_xy):
            outs[:] = oracle.ref_bilateralfilter_batch(images, ins, N, K, H, W, sigma_rgb, sigma_xy)
        ident = lambda f: f
        fenv = {"torch": torch, "np": np, "Function": torch.autograd.Function, "custom_fwd": ident,
                "custom_bwd": ident, "bilateralfilter_batch": bilateralfilter_batch}
        real_sync = torch.cuda.synchronize
        torch.cuda.synchronize = lambda *a, **k: None
        try:
            Fn = load_class("dlib/crf/dense_crf_loss.py", "DenseCRFLossFunction", fenv)
            n, k, h, w = 2, 3, 31, 37
            img = torch.from_numpy(synth.make_images(n, h, w, "natural", seed=5))
            seg = torch.from_numpy(synth.make_segs(n, k, h, w, seed=5)).requires_grad_(True)
            loss = 1e-3 * Fn.apply(img, seg, 15.0, 100.0)
            loss.backward()
        finally:
            torch.cuda.synchronize = real_sync
        np.savez_compressed(os.path.join(OUT, "py_dense_crf_loss.npz"), image=img.numpy(), seg=seg.detach().numpy(),
                            loss=loss.detach().numpy(), grad=seg.grad.numpy(), weight=np.float32(1e-3))
        # ... and ColorDenseCRFLossFunction (dlib/crf/color_dense_crf_loss.py:33-76) on a width-concatenated clip
        def colorbilateralfilter_batch(images, ins, outs, N, K, H, W, sigma_rgb, DIM):
            outs[:] = oracle.ref_colorbilateralfilter_batch(images, ins, N, K, H, W, sigma_rgb, DIM)
        cenv = dict(fenv, colorbilateralfilter_batch=colorbilateralfilter_batch)
        torch.cuda.synchronize = lambda *a, **k: None
        try:
            CFn = load_class("dlib/crf/color_dense_crf_loss.py", "ColorDenseCRFLossFunction", cenv)
            n, k, h, w = 2, 2, 24, 3 * 20
            cimg = torch.from_numpy(synth.make_images(n, h, w, "natural", seed=6))
            cseg = torch.from_numpy(synth.make_segs(n, k, h, w, seed=6)).requires_grad_(True)
            closs = 1e-3 * CFn.apply(cimg, cseg, 15.0)
            closs.backward()
        finally:
            torch.cuda.synchronize = real_sync
        np.savez_compressed(os.path.join(OUT, "py_color_dense_crf_loss.npz"), image=cimg.numpy(),
                            seg=cseg.detach().numpy(), loss=closs.detach().numpy(), grad=cseg.grad.numpy(),
                            weight=np.float32(1e-3))
    # --- GetRoiSingleCam (dlib/cams/tcam_seeding.py:316-430), the CPU ROI of a CAM, executed with REAL OpenCV (cv2 is in
    #     this image): thresholding, component statistics and selection, cv2.findContours / boundingRect through the
    #     reference's own compute_bboxes_from_scoremaps_ext_contours (dlib/utils/wsol.py:81-150) and get_largest_bbox
    #     (tcam_seeding.py:34-41).  scikit-image is not installed: skimage.measure.label(connectivity=1) is stood in
    #     for by scipy.ndimage.label (4-neighbour structure, raster-order numbering like skimage's) and
    #     threshold_otsu by oracle/seeding.py's restatement of skimage 0.17.2 -- those two stay unpinned and are named
    #     so in DESIGN.md.  np.float (removed from numpy 2) is mapped to float.
    try:
        import cv2
        from scipy import ndimage

        class _NP:
            float = float

            def __getattr__(self, name):
                return getattr(np, name)

        class _Measure:
            @staticmethod
            def label(blobs, background=0, connectivity=1, return_num=False):
                assert connectivity == 1 and background == 0 and not return_num
                return ndimage.label(blobs, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])[0]

        wsrc = open(os.path.join(REF, "dlib/utils/wsol.py")).read()
        wenv = {"np": np, "cv2": cv2, "Union": __import__("typing").Union,
                "_CONTOUR_INDEX": 1 if cv2.__version__.split('.')[0] == '3' else 0}
        for node in ast.parse(wsrc).body:
            if isinstance(node, ast.FunctionDef) and node.name in ("check_scoremap_validity", "check_box_convention",
                                                                   "compute_bboxes_from_scoremaps_ext_contours"):
                exec(compile("\n".join(wsrc.splitlines()[node.lineno - 1:node.end_lineno]), "wsol.py", "exec"), wenv)
        tsrc = open(os.path.join(REF, "dlib/cams/tcam_seeding.py")).read()
        renv = {"np": _NP(), "torch": torch, "Tuple": Tuple, "constants": types.SimpleNamespace(
                    ROI_SELECT=['roi_all', 'roi_high_density', 'largest'], ROI_ALL='roi_all',
                    ROI_H_DENSITY='roi_high_density', ROI_LARGEST='largest'),
                "measure": _Measure, "threshold_otsu": lambda a: oseed.threshold_otsu_skimage(a),
                "compute_bboxes_from_scoremaps_ext_contours": wenv["compute_bboxes_from_scoremaps_ext_contours"],
                "check_box_convention": wenv["check_box_convention"]}
        for node in ast.parse(tsrc).body:
            if isinstance(node, ast.FunctionDef) and node.name == "get_largest_bbox":
                exec(compile("\n".join(tsrc.splitlines()[node.lineno - 1:node.end_lineno]), "tcam_seeding.py", "exec"), renv)
        exec(compile(cut_class("dlib/cams/tcam_seeding.py", "GetRoiSingleCam"), "tcam_seeding.py:GetRoiSingleCam", "exec"),
             renv)
        gen = torch.Generator().manual_seed(91)
        rois = {}
        ci = 0
        cj = 0
        for (h, w, low_r) in ((48, 64, 6), (37, 53, 5), (96, 96, 9), (160, 144, 7)):
            lowc = torch.rand((3, 1, low_r, low_r + 1), generator=gen)
            cams = F.interpolate(lowc, size=(h, w), mode="bilinear", align_corners=False)[:, 0]
            cams[1] = cams[1] * (torch.rand((h, w), generator=gen) > 0.35)      # speckle: many small components
            for i in range(cams.shape[0]):
                rois[f"cam{cj}"] = cams[i].numpy()
                cj += 1
                for method in ('roi_high_density', 'largest'):
                    for thresh in (None, 0.55):
                        getter = renv["GetRoiSingleCam"](roi_method=method, p_min_area_roi=0.05)
                        roi_o, mask_o, bbox_o = getter(cams[i], thresh=thresh)
                        rois[f"c{ci}_cfg"] = np.array([method, "" if thresh is None else str(thresh), str(cj - 1)])
                        rois[f"c{ci}_roi"] = roi_o.numpy().astype(np.int8)
                        rois[f"c{ci}_mask"] = mask_o.numpy().astype(np.int8)
                        rois[f"c{ci}_bbox"] = bbox_o.numpy()
                        ci += 1
        rois["n_cases"] = np.int64(ci)
        rois["cv2_version"] = np.array(cv2.__version__)
        np.savez_compressed(os.path.join(OUT, "py_get_roi_single_cam.npz"), **rois)
        print(f"wrote py_get_roi_single_cam.npz ({ci} cases, OpenCV {cv2.__version__})")
    except ImportError as exc:
        print("skipped py_get_roi_single_cam.npz:", exc)

    # --- clip grouping of the joint colour CRF (dlib/losses/tcam.py:32-45 group_ordered_frames, :207-232 pair_samples)
    src = open(os.path.join(REF, "dlib/losses/tcam.py")).read()
    gscope = {"torch": torch, "Tuple": Tuple}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "group_ordered_frames":
            exec(compile("\n".join(src.splitlines()[node.lineno - 1:node.end_lineno]), "tcam.py", "exec"), gscope)
    pair = load("dlib/losses/tcam.py", "RgbJointConRanFieldTcams", "pair_samples", gscope)
    seq_iter = torch.tensor([3., 1., 3., 1., 7., 3., 1., 3.])
    frm_iter = torch.tensor([2., 0., 0., 1., 0., 1., 2., 3.])
    groups = gscope["group_ordered_frames"](seq_iter, frm_iter)
    groups = [[int(i) for i in grp] for grp in groups]
    gen = torch.Generator().manual_seed(3)
    imgs = torch.randint(0, 256, (8, 3, 6, 5), generator=gen).float()
    probs = torch.rand((8, 2, 6, 5), generator=gen)
    clip = {"seq_iter": seq_iter.numpy(), "frm_iter": frm_iter.numpy(), "imgs": imgs.numpy(), "probs": probs.numpy(),
            "n_groups": np.int64(len(groups))}
    for gi, grp in enumerate(groups):
        clip[f"group{gi}"] = np.array(grp, dtype=np.int64)
        if len(grp) > 1:
            pi, pc = pair(o_idx=grp, imgs=imgs, prob_cams=probs)
            clip[f"pair{gi}_img"] = pi.numpy()
            clip[f"pair{gi}_cam"] = pc.numpy()
    np.savez_compressed(os.path.join(OUT, "py_clip_grouping.npz"), **clip)
    print("wrote py_temporal_agg.npz, py_prepare_std_cams.npz, py_frame_pickers.npz, py_seed_sampling.npz, "
          "py_dense_crf_loss.npz, py_clip_grouping.npz")


if __name__ == "__main__":
    main()
