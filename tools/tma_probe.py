import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import homography_js_b200 as hg
from oracle.homography_ref import RefHomography, RefImageData
rng = np.random.default_rng(0)
W, H = 640, 480
img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
for kind, src, dst in [("affine", [[0,0],[0,H],[W,0]], [[10,20],[30,H+40],[W+5,8]]),
                       ("projective", [[0,0],[0,H],[W,0],[W,H]], [[W/10,0],[W/10,H],[W,H/4],[W,3*H/4]])]:
    a = RefHomography(kind); b = hg.Homography(kind)
    a.setReferencePoints(src, dst); b.setReferencePoints(src, dst)
    ra = a.warp(RefImageData(img.reshape(-1).copy(), W, H), None, True) if False else a.warp(RefImageData(img.reshape(-1).copy(), W, H))
    rb = b.warp(hg.ImageData(img.reshape(-1).copy(), W, H))
    print(kind, ra.width, ra.height, rb.width, rb.height, "equal:", np.array_equal(np.asarray(ra.data), np.asarray(rb.data)), flush=True)
