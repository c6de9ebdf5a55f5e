import numpy as np

from lensed_b200 import fits


def test_roundtrip_layers_and_section(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.random((20, 30)).astype(np.float32)
    b = rng.random((20, 30)).astype(np.float32)
    path = str(tmp_path / "x.fits")
    fits.write_layers(path, [a, b], ["IMG", "RES"])
    hdus = fits.read_hdus(path)
    # the reference's results layout (src/data.c:129-165): empty primary HDU, then one IMAGE extension per layer
    assert len(hdus) == 3 and hdus[0][1] is None and hdus[0][0]["NAXIS"] == 0 and hdus[0][0]["BITPIX"] == 16
    assert "ORIGIN" in hdus[0][0] and "DATE" in hdus[0][0]
    assert hdus[1][0]["EXTNAME"] == "IMG" and hdus[2][0]["EXTNAME"] == "RES" and hdus[1][0]["XTENSION"] == "IMAGE"
    img, pcs = fits.read_image(path)
    assert np.array_equal(img, a) and pcs == (1.0, 1.0, 1.0, 1.0)
    res, _ = fits.read_image(path + "[RES]")
    assert np.array_equal(res, b)
    # CFITSIO image section: 1-based inclusive, x range first (examples/test_sersic_bulge.ini:2)
    cut, pcs = fits.read_image(path + "[3:12,5:9]")
    assert cut.shape == (5, 10) and np.array_equal(cut, a[4:9, 2:12]) and pcs == (3.0, 5.0, 1.0, 1.0)
    # increments and reversed ranges: origin = first pixel of the section, scale = +-increment (src/data.c:262-270)
    cut, pcs = fits.read_image(path + "[3:12:3,9:5]")
    assert np.array_equal(cut, a[8:3:-1, 2:12:3]) and pcs == (3.0, 9.0, 3.0, -1.0)
    cut, pcs = fits.read_image(path + "[30:1,20:1:2]")
    assert np.array_equal(cut, a[::-2, ::-1]) and pcs == (30.0, 20.0, -1.0, -2.0)
    import pytest
    with pytest.raises(ValueError, match="not supported"):
        fits.read_image(path + "[*,5:9]")
    # a plain image file: first layer in the primary HDU
    fits.write_layers(path, [a], ["IMG"], empty_primary=False)
    hdus = fits.read_hdus(path)
    assert len(hdus) == 1 and np.array_equal(hdus[0][1], a)
