"""homography.js_b200 — a Blackwell-native (sm_100a) image-warp engine behind the Homography.js class surface.

    from homography_js_b200 import Homography          # repo-root shim (the directory name has a dot)
    h = Homography("projective"); h.setReferencePoints(src, dst); h.setImage(img); out = h.warp()

Layout: csrc/ (hand-written CUDA kernels + the C ABI of include/hgwarp.h), _abi.py (ctypes binding of
that ABI), homography.py (host mirror of the reference class), js/ (the Node.js shim + N-API addon
source a maintainer would ship), build.py (nvcc recipe).
"""
from . import _abi, workloads
from ._abi import Context, HgError, HgFrame, HgStreamInfo, Pipe, device_count
from .homography import Homography, HomographyError, ImageData

__all__ = ["Homography", "HomographyError", "ImageData", "Context", "HgError", "HgFrame", "HgStreamInfo", "Pipe", "device_count", "_abi", "workloads"]
