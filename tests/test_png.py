"""Device-side PNG writer (`dl_png_stored`, B200_PNG=gpu) against its oracle and against real PNG readers.
CPU: the oracle itself is a valid PNG that PIL and zlib decode to the input pixels.  GPU: the kernel's bytes
equal the oracle's byte for byte, for sizes whose scanline stream spans one / several / exactly-full stored blocks."""
import io
import zlib

import numpy as np
import pytest

SIZES = [(8, 8), (64, 64), (128, 96), (512, 512), (85, 257), (768, 512), (5, 4369)]   # (h, w); 5*(1+3*4369) = 65540


def _img(h, w, seed=0):
    return np.random.RandomState(seed).randint(0, 256, size=(h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w", SIZES)
def test_oracle_png_is_a_valid_png(h, w):
    from PIL import Image
    from oracle.png import png_stored
    img = _img(h, w)
    data = png_stored(img)
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    back = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert np.array_equal(back, img)
    # the IDAT payload is a plain zlib stream of the filtered scanlines
    n = int.from_bytes(data[33:37], "big")
    raw = zlib.decompress(data[41:41 + n])
    assert len(raw) == h * (1 + 3 * w)


def test_size_formula_matches_oracle():
    from oracle.png import png_stored
    import dreamlab_b200.lib as lib
    for h, w in SIZES:
        assert lib.png_stored_size(h, w) == len(png_stored(_img(h, w)))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", SIZES)
def test_cuda_png_bytes_equal_oracle(h, w):
    import torch
    from PIL import Image
    from oracle.png import png_stored
    import dreamlab_b200.lib as lib
    imgs = np.stack([_img(h, w, seed=s) for s in range(3)])
    out, size = lib.png_stored(torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    for i in range(3):
        got = out[i, :size].tobytes()
        ref = png_stored(imgs[i])
        assert len(got) == len(ref)
        assert got == ref, next(j for j in range(len(ref)) if got[j] != ref[j])
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(got)).convert("RGB")), imgs[i])
