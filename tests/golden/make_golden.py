"""Regenerates tests/golden/*.npz from the reference's own fixtures (run in the build container only).

    python tests/golden/make_golden.py

Reads /root/reference/test/*.png (the reference's test inputs and its single pixel-exact known-answer
output, test/transformedImage.png == Documentation/exampleImages/nodeExampleOutput.png, produced by
test/nodeTest.js) and stores the decoded RGBA8 arrays, so that nothing at test time needs
/root/reference (it does not exist on the GPU box).  No reference source code is copied.
"""
import hashlib
import os

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def rgba(path):
    im = Image.open(path)
    assert im.mode == "RGBA", (path, im.mode)
    return np.asarray(im, dtype=np.uint8)


def main():
    src = rgba(f"{REF}/test/testImgLogoBlack.png")
    out = rgba(f"{REF}/test/transformedImage.png")
    doc = rgba(f"{REF}/Documentation/exampleImages/nodeExampleOutput.png")
    assert np.array_equal(out, doc), "the two copies of the node golden differ"
    np.savez_compressed(
        os.path.join(HERE, "node_test_golden.npz"),
        src=src, out=out,
        # test/nodeTest.js:5-6
        src_points=np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.float64),
        dst_points=np.array([[1 / 10, 1 / 2], [0, 1], [9 / 10, 1 / 2], [1, 1]], np.float64),
        sha256_out_png=np.frombuffer(hashlib.sha256(open(f"{REF}/test/transformedImage.png", "rb").read()).digest(), np.uint8),
    )
    # the other two 400x400 inputs used by test/test.js (inputs only: the reference has no goldens for them)
    np.savez_compressed(os.path.join(HERE, "test_inputs.npz"),
                        testImg=rgba(f"{REF}/test/testImg.png"),
                        testImgLogoWhite=rgba(f"{REF}/test/testImgLogoWhite.png"))
    print("src", src.shape, "out", out.shape)


if __name__ == "__main__":
    main()
