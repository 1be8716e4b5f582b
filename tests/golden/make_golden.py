"""Generates tests/golden/*.npz from the REFERENCE's own C++ (oracle/_ref, built in place from
/root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture stores the inputs (uint8 image, float32 segmentations), and what the reference
produced for them: the filtered tensor AS, the loss and gradient of DenseCRFLossFunction
(dlib/crf/dense_crf_loss.py:56-74 restated in oracle.densecrf_loss_fwd_bwd, with the native
call going to the reference build), and the lattice size M.  The large 224x224 fixture keeps
only a strided sample of AS plus float64 checksums to stay small.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from tcam_wsol_video_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, kind, N, K, H, W, color(DIM or 0), sigma_rgb, sigma_xy
    ("bf5_noise_n2_k2_32x40", "noise", 2, 2, 32, 40, 0, 15.0, 100.0),
    ("bf5_natural_n2_k3_31x37", "natural", 2, 3, 31, 37, 0, 15.0, 100.0),
    ("bf5_noise_n1_k10_24x24_s", "noise", 1, 10, 24, 24, 0, 7.5, 20.0),
    ("cbf3_noise_n2_k2_32x40", "noise", 2, 2, 32, 40, 3, 15.0, 0.0),
    ("cbf3_natural_n1_k2_28x84", "natural", 1, 2, 28, 84, 3, 15.0, 0.0),
    ("cbf1_noise_n1_k2_32x32", "noise", 1, 2, 32, 32, 1, 15.0, 0.0),
]
BIG = ("bf5_noise_n1_k2_224x224", "noise", 1, 2, 224, 224, 0, 15.0, 100.0)


def run_case(kind, N, K, H, W, dim, srgb, sxy, seed):
    channels = dim if dim else 3
    img = synth.make_images(N, H, W, kind, seed=seed, channels=channels)
    seg = synth.make_segs(N, K, H, W, seed=seed)
    assert oracle.have_ref(), "oracle/_ref missing: run `make -C oracle` where /root/reference is mounted"
    if dim:
        loss, grad, AS = oracle.color_densecrf_loss_fwd_bwd(img, seg, srgb, 1.0, oracle.ref_colorbilateralfilter_batch)
        M = oracle.ref_lattice_color(img[0], H, W, srgb, dim).m
    else:
        loss, grad, AS = oracle.densecrf_loss_fwd_bwd(img, seg, srgb, sxy, 1.0, oracle.ref_bilateralfilter_batch)
        M = oracle.ref_lattice_bilateral(img[0], H, W, srgb, sxy).m
    return img, seg, loss, grad, AS, M


def main():
    for i, (name, kind, N, K, H, W, dim, srgb, sxy) in enumerate(CASES):
        img, seg, loss, grad, AS, M = run_case(kind, N, K, H, W, dim, srgb, sxy, seed=100 + i)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), image_u8=img.astype(np.uint8), seg=seg, AS=AS,
                            loss=np.float32(loss), grad=grad, M0=np.int32(M), dim=np.int32(dim),
                            sigma_rgb=np.float32(srgb), sigma_xy=np.float32(sxy), kind=kind, seed=np.int32(100 + i))
        print(name, "loss", loss, "M0", M)
    name, kind, N, K, H, W, dim, srgb, sxy = BIG
    img, seg, loss, grad, AS, M = run_case(kind, N, K, H, W, dim, srgb, sxy, seed=0)
    flat = AS.ravel()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), AS_sample=flat[::97].copy(), stride=np.int32(97),
                        AS_sum=np.float64(flat.astype(np.float64).sum()),
                        AS_sqsum=np.float64((flat.astype(np.float64) ** 2).sum()), loss=np.float32(loss),
                        M0=np.int32(M), image_sum=np.float64(img.sum(dtype=np.float64)),
                        seg_sum=np.float64(seg.sum(dtype=np.float64)), dim=np.int32(dim),
                        sigma_rgb=np.float32(srgb), sigma_xy=np.float32(sxy), kind=kind, seed=np.int32(0),
                        shape=np.array([N, K, H, W], dtype=np.int32))
    print(name, "loss", loss, "M0", M)


if __name__ == "__main__":
    main()
