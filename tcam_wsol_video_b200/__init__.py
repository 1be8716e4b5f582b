"""tcam_wsol_video_b200 -- B200-native (sm_100a) DenseCRF-loss hot path of TCAM.

Public surface mirrors the reference's own modules for this path:

* ``DenseCRFLoss`` / ``ColorDenseCRFLoss``      (dlib/crf/dense_crf_loss.py, color_dense_crf_loss.py)
* ``bilateralfilter[_batch]`` / ``colorbilateralfilter[_batch]``  (the SWIG modules' call contract)
* ``temporal_cam_max`` / ``TCAMSeeder`` / ``GetRoiSingleCam``   (dlib/datasets/wsol_loader.py:585-600, dlib/cams/tcam_seeding.py)
* ``temporal``: frame pickers, ``re_normalize_cam``, ``prepare_std_cams_disq``   (wsol_loader.py:448-459,630-635, train_wsol.py:417-432)
* ``losses``: ``ConRanFieldTcams``, ``RgbJointConRanFieldTcams``, ``SelfLearningTcams``   (dlib/losses/tcam.py)
* ``crf_post_processing.DenseCRFFilter``        (dlib/crf/crf_post_processing.py; mean field on one reusable lattice)
* ``ops.Lattice``: build once, apply many times (A and A^T); ``dist``: batch sharding + scalar loss all-reduce

Everything runs through ``csrc/libtcamcrf.so`` (hand-written CUDA behind a C ABI,
``include/tcamcrf.h``).  There is no CPU fallback: importing the package is cheap,
but the first call raises if the library is missing or no B200 is visible.
"""
__version__ = "0.1.0"
