"""Timeline of one tcamcrf_loss_fwd_bwd_host call (TCAMCRF_HOST_TRACE=1): python tools/host_trace.py [K]"""
import ctypes, sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tcam_wsol_video_b200 import _lib, synth
lib=_lib.load()
N,K,H,W=32,int(sys.argv[1]) if len(sys.argv)>1 else 10,224,224
img=torch.from_numpy(synth.make_images(N,H,W,"noise",seed=0)).pin_memory()
seg=torch.from_numpy(synth.make_segs(N,K,H,W,seed=0)).pin_memory()
loss=torch.zeros(1).pin_memory(); grad=torch.empty(N,K,H,W).pin_memory()
cfg=_lib.make_config(_lib.FEAT_XY_RGB,3,15.0,100.0)
def step():
    _lib.check(lib.tcamcrf_loss_fwd_bwd_host(ctypes.byref(cfg), img.data_ptr(), seg.data_ptr(), loss.data_ptr(), grad.data_ptr(), N,K,H,W, 2e-9),"x")
for _ in range(5): step()
_lib.set_tuning("HOST_TRACE", 1)
t0=time.perf_counter(); step(); print("wall ms", (time.perf_counter()-t0)*1e3)
_lib.set_tuning("HOST_TRACE", 0)
t0=time.perf_counter()
for _ in range(20): step()
print("mean wall ms over 20", (time.perf_counter()-t0)*1e3/20)
