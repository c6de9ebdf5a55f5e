import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, lensed_b200 as L
print("LD_LIBRARY_PATH", os.environ.get("LD_LIBRARY_PATH"))
print("LCU_NVRTC_FLAGS", os.environ.get("LCU_NVRTC_FLAGS"))
print([l.split()[-1] for l in open("/proc/self/maps") if "nvrtc" in l][:3])
img = np.zeros((256, 256), np.float32)
for dev in (-1, 0):
    ctx = L.Context(device=dev)
    for fl in (0, 4, 8, 1, 5, 13):
        m = L.Model(ctx, ["sie_plus_shear", "sersic", "sersic", "sky"], img, img, rule="g7k15", psf=np.ones((25, 25), np.float32)/625, flags=fl)
        print(dev, fl, len(m.cubin), hashlib.md5(m.cubin).hexdigest(), flush=True)
