"""CPU-only: the six-tap form of the in-model mel upsampler used by csrc/sashimi_kernels.cu::mel_upsample_kernel
equals ConvTranspose2d(1, 1, (3, 2s), stride (1, s), padding (1, s/2)) + leaky-ReLU(0.4)
(models/sashimi.py:133-141, models/wavenet.py:62-70).  The CUDA kernel itself is compared with the reference
goldens in tests/test_gpu_models.py (tiny_*_cond, full_unet_d32_cond)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def six_tap(x, w, bias, s):
    cb, Fb, Win = x.shape
    out = np.empty((cb, Fb, Win * s), dtype=np.float64)
    for xo in range(Win * s):
        k0 = (xo + s // 2) % s
        acc = np.full((cb, Fb), bias, dtype=np.float64)
        for kf in range(3):
            for q in range(2):
                kx = k0 + q * s
                num = xo + s // 2 - kx
                if num < 0 or num // s >= Win:
                    continue
                xi = num // s
                for f in range(Fb):
                    fi = f + 1 - kf
                    if 0 <= fi < Fb:
                        acc[:, f] += x[:, fi, xi] * w[kf, kx]
        out[:, :, xo] = acc
    return np.where(out > 0, out, 0.4 * out)


@pytest.mark.parametrize("s,Fb,Win", [(16, 5, 4), (4, 3, 7), (2, 4, 5), (16, 80, 3)])
def test_six_tap_transposed_conv(s, Fb, Win):
    g = torch.Generator().manual_seed(s * 100 + Win)
    x = torch.randn(2, Fb, Win, generator=g, dtype=torch.float64)
    w = torch.randn(1, 1, 3, 2 * s, generator=g, dtype=torch.float64)
    b = torch.randn(1, generator=g, dtype=torch.float64)
    ref = F.leaky_relu(F.conv_transpose2d(x.unsqueeze(1), w, b, stride=(1, s), padding=(1, s // 2)), 0.4).squeeze(1)
    got = six_tap(x.numpy(), w[0, 0].numpy(), float(b), s)
    assert ref.shape == (2, Fb, Win * s)
    np.testing.assert_allclose(got, ref.numpy(), atol=1e-12)
