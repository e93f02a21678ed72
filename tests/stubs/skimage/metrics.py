"""skimage.metrics as test.py:13-14 imports it, backed by this package's evaluation kernels:
structural_similarity -> sesr_pytorch_quantize_b200.evalpost.ssim (CUDA), peak_signal_noise_ratio -> its one-line formula."""
import numpy as np
import torch


def peak_signal_noise_ratio(image_true, image_test, *, data_range=None):
    err = np.mean((np.asarray(image_true, dtype=np.float64) - np.asarray(image_test, dtype=np.float64)) ** 2)
    return 10 * np.log10((data_range ** 2) / err)


def structural_similarity(im1, im2, *, data_range=None, channel_axis=None, multichannel=False, **_):
    from sesr_pytorch_quantize_b200.evalpost import ssim
    a, b = np.asarray(im1, dtype=np.float32), np.asarray(im2, dtype=np.float32)
    if channel_axis is not None:
        a, b = np.moveaxis(a, channel_axis, 0), np.moveaxis(b, channel_axis, 0)
    return ssim(torch.from_numpy(np.ascontiguousarray(a)).cuda(), torch.from_numpy(np.ascontiguousarray(b)).cuda(),
                data_range=float(data_range))
