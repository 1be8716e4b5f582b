"""tcam_wsol_video_b200 -- B200-native (sm_100a) DenseCRF-loss hot path of TCAM.

Public surface mirrors the reference's own modules for this path:

* ``DenseCRFLoss`` / ``ColorDenseCRFLoss``      (dlib/crf/dense_crf_loss.py, color_dense_crf_loss.py)
* ``bilateralfilter[_batch]`` / ``colorbilateralfilter[_batch]``  (the SWIG modules' call contract)
* ``temporal_cam_max`` / ``TCAMSeeder``          (dlib/datasets/wsol_loader.py:585-600, dlib/cams/tcam_seeding.py)

Everything runs through ``csrc/libtcamcrf.so`` (hand-written CUDA behind a C ABI,
``include/tcamcrf.h``).  There is no CPU fallback: importing the package is cheap,
but the first call raises if the library is missing or no B200 is visible.
"""
__version__ = "0.1.0"
