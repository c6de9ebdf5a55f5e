import numpy as np

from lensed_b200 import fits


def test_roundtrip_layers_and_section(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.random((20, 30)).astype(np.float32)
    b = rng.random((20, 30)).astype(np.float32)
    path = str(tmp_path / "x.fits")
    fits.write_layers(path, [a, b], ["IMG", "RES"])
    hdus = fits.read_hdus(path)
    assert len(hdus) == 2 and hdus[1][0]["EXTNAME"] == "RES"
    img, pcs = fits.read_image(path)
    assert np.array_equal(img, a) and pcs == (1.0, 1.0, 1.0, 1.0)
    res, _ = fits.read_image(path + "[RES]")
    assert np.array_equal(res, b)
    # CFITSIO image section: 1-based inclusive, x range first (examples/test_sersic_bulge.ini:2)
    cut, pcs = fits.read_image(path + "[3:12,5:9]")
    assert cut.shape == (5, 10) and np.array_equal(cut, a[4:9, 2:12]) and pcs == (3.0, 5.0, 1.0, 1.0)
