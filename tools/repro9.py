import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import homography_js_b200 as hg
from oracle import oracle as O
ctx = hg.Context(0)
for seed in range(5):
    rng = np.random.default_rng(600 + seed)
    W, H = 120, 90
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ctx.image_set(img, W, H)
    h6 = [1e-3, -2e-3, 1 / 64, -1 / 50, 3e-4][seed]
    h = np.array([rng.uniform(0.5, 1.5), rng.uniform(-0.2, 0.2), rng.uniform(-10, 10),
                  rng.uniform(-0.2, 0.2), rng.uniform(0.5, 1.5), rng.uniform(-10, 10), h6, -0.0 if seed % 2 else 0.0])
    print("seed", seed, flush=True)
    got = ctx.warp_inverse_matrix(h, -30, -20, 260, 170)
    want = O.warp_inverse_geometric(img, W, H, h, -30, -20, 260, 170)
    print("  equal", np.array_equal(got, want), flush=True)
