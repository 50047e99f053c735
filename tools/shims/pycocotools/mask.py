"""Import-only: the synthetic drop-in runs never encode RLE masks."""


def _unavailable(*a, **k):
    raise RuntimeError("pycocotools is not installed in this image (import shim)")


encode = decode = area = toBbox = iou = merge = frPyObjects = _unavailable
