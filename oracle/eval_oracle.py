"""CPU ORACLE (numpy) for the evaluation post-step of the reference's test.py:128-170.  TEST INFRASTRUCTURE ONLY.
Restates, for one frame as the script does: the nearest-neighbour x2 upsample-and-add (:149-155), clip (:157),
compute_psnr (:128-131), rgb_to_yuv (:133-136) and the per-model metric choice (:165-170)."""
import numpy as np


def compute_psnr(img_pred, img_true, data_range=255., eps=1e-8):             # test.py:128-131
    err = np.mean((img_pred - img_true) ** 2)
    return 10. * np.log10((data_range ** 2) / (err + eps))


def rgb_to_yuv(img):                                                         # test.py:133-136 (HWC, range 0-1 -> 0-255)
    return np.clip(np.matmul(img, np.array([65.481, 128.553, 24.966])) + 16., 0, 255.)


def post(y, x, r):
    """y [C,Hr,Wr], x [C,H,W] or None -> clipped HWC image as test.py:148-157."""
    g = y.astype(np.float32)
    if x is not None:
        g = (g + np.repeat(np.repeat(x.astype(np.float32), r, axis=1), r, axis=2)).astype(np.float32)
    return np.clip(g.transpose(1, 2, 0), 0, 1)


def psnr(img_hwc, gt_hwc, mflag):
    if mflag == 5:
        return compute_psnr(gt_hwc[:, :, 0] * 255., img_hwc[:, :, 0] * 255.)
    if mflag == 6:
        return compute_psnr(rgb_to_yuv(gt_hwc), rgb_to_yuv(img_hwc))
    mse = np.mean((gt_hwc.astype(np.float64) - img_hwc.astype(np.float64)) ** 2)        # skimage compare_psnr, data_range 1
    return 10. * np.log10(1.0 / mse)


def ssim(a, b, data_range=1.0, win_size=7):
    """skimage.metrics.structural_similarity restated (skimage 0.19+ `_structural_similarity.py`: uniform filter of
    win_size, K1 = 0.01, K2 = 0.03, use_sample_covariance=True, mean over the map cropped by (win_size - 1) // 2).
    a, b: [planes, H, W] or [H, W]; planes are averaged like channel_axis.  skimage itself is not installed in this
    image (parity unpinned against the third-party code; the formula is its published one), scipy's uniform_filter is."""
    from scipy.ndimage import uniform_filter
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.ndim == 2:
        a, b = a[None], b[None]
    k1, k2 = 0.01, 0.03
    npix = win_size ** 2
    cov_norm = npix / (npix - 1)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    pad = (win_size - 1) // 2
    vals = []
    for x, y in zip(a, b):
        ux, uy = uniform_filter(x, size=win_size), uniform_filter(y, size=win_size)
        uxx, uyy, uxy = uniform_filter(x * x, size=win_size), uniform_filter(y * y, size=win_size), uniform_filter(x * y, size=win_size)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        vals.append(s[pad:-pad, pad:-pad].mean())
    return float(np.mean(vals))
