import numpy as np


def ResizeNearestNeighbor(images, size, align_corners=False, half_pixel_centers=False):
    """out[y, x] = in[floor(y * in_h / out_h), floor(x * in_w / out_w)] (align_corners=False)."""
    x = np.asarray(images)
    assert not align_corners and not half_pixel_centers
    ih, iw = x.shape[1], x.shape[2]
    oh, ow = int(size[0]), int(size[1])
    ys = np.minimum((np.arange(oh) * (ih / oh)).astype(np.int64), ih - 1)
    xs = np.minimum((np.arange(ow) * (iw / ow)).astype(np.int64), iw - 1)
    return x[:, ys][:, :, xs]
