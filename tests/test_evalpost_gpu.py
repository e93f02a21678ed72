"""Evaluation post-step (test.py:148-170) on device vs its numpy restatement: the clipped image bit for bit; the PSNR of
one frame to 1e-9 relative where the reference computes in float64 (fp64 sums in a different order), and to 5e-6 for the
x4 metric, which the reference evaluates in float32 (`gts*255.` on float32 arrays, np.mean in float32)."""
import numpy as np
import pytest
import torch

import eval_oracle as eo
from sesr_pytorch_quantize_b200.evalpost import sr_post, METRIC_OF_MFLAG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mflag,C,r,add", [(6, 3, 2, True), (5, 1, 4, False), (3, 3, 1, False)])
def test_post_and_psnr_vs_oracle(mflag, C, r, add):
    rng = np.random.default_rng(mflag)
    H, W = 37, 51
    x = rng.random((1, C, H, W), dtype=np.float32)
    y = (rng.random((1, C, H * r, W * r), dtype=np.float32) * 1.4 - 0.2).astype(np.float32)     # exercises both clip sides
    gt = rng.random((1, C, H * r, W * r), dtype=np.float32)
    img, psnr = sr_post(torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda() if add else None, torch.from_numpy(gt).cuda(),
                        METRIC_OF_MFLAG[mflag])
    want = eo.post(y[0], x[0] if add else None, r)
    assert np.array_equal(img.cpu().numpy()[0].transpose(1, 2, 0), want)
    want_psnr = float(eo.psnr(want, gt[0].transpose(1, 2, 0), mflag))
    assert abs(psnr - want_psnr) <= (5e-6 if mflag == 5 else 1e-9) * abs(want_psnr)


def test_post_without_ground_truth_and_errors():
    from sesr_pytorch_quantize_b200 import SesrqError
    y = torch.rand(2, 3, 8, 12).cuda() * 2 - 0.5
    img, psnr = sr_post(y)
    assert psnr is None and torch.equal(img, y.clamp(0, 1))
    with pytest.raises(SesrqError):
        sr_post(y.cpu())
    with pytest.raises(SesrqError):
        sr_post(y, torch.rand(2, 3, 5, 6).cuda())


@pytest.mark.parametrize("shape", [(3, 45, 70), (1, 7, 7), (2, 3, 33, 64), (1, 130, 97)])
def test_ssim_vs_oracle(shape):
    """SSIM kernel (test.py:172-175, skimage defaults) against the scipy/numpy restatement, incl. the clipped-image
    regime of the script (values in [0, 1], data_range 1) and a window that exactly fills the image."""
    from sesr_pytorch_quantize_b200.evalpost import ssim
    rng = np.random.default_rng(sum(shape))
    a = rng.random(shape, dtype=np.float32)
    b = np.clip(a + 0.1 * rng.standard_normal(shape).astype(np.float32), 0, 1)
    got = ssim(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), data_range=1.0)
    want = eo.ssim(a.reshape(-1, shape[-2], shape[-1]), b.reshape(-1, shape[-2], shape[-1]), data_range=1.0)
    assert abs(got - want) <= 1e-9, (got, want)
    assert abs(ssim(torch.from_numpy(a).cuda(), torch.from_numpy(a).cuda()) - 1.0) <= 1e-12


def test_ssim_errors():
    from sesr_pytorch_quantize_b200 import SesrqError
    from sesr_pytorch_quantize_b200.evalpost import ssim
    with pytest.raises(SesrqError):
        ssim(torch.rand(3, 6, 20).cuda(), torch.rand(3, 6, 20).cuda())
    with pytest.raises(SesrqError):
        ssim(torch.rand(3, 16, 20), torch.rand(3, 16, 20))
