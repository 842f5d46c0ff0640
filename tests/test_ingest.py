"""CPU: data-ingest helpers (sort / rescale / magnitude conversion conventions of the notebooks)."""
import os

import numpy as np
import pytest

from gaussn_b200 import ingest

REF_DATA = "/root/reference/data"


def test_get_flux_and_rescale():
    flux, err = ingest.get_flux([27.5, 25.0, 30.0], [0.1, 0.1, 0.2])
    np.testing.assert_allclose(flux, [1.0, 10.0, 0.1])
    np.testing.assert_allclose(err, flux * np.array([0.1, 0.1, 0.2]) * np.log(10) / 2.5)
    y, e, f = ingest.rescale([1.0, 3.0, 5.0], [0.4, 0.4, 0.8])
    assert f == 4.0 and np.allclose(y, [0.25, 0.75, 1.25]) and np.allclose(e, [0.1, 0.1, 0.2])


def test_sort_rows_gives_the_order_prepare_indices_assumes():
    t = np.array([3.0, 1.0, 2.0, 5.0, 4.0, 0.5])
    band = np.array(["r", "g", "r", "g", "g", "r"])
    image = np.array(["image_2", "image_1", "image_1", "image_2", "image_1", "image_2"])
    ts, bs, ims, extra = ingest.sort_rows(t, band, image, np.arange(6))
    assert list(bs) == ["g", "g", "g", "r", "r", "r"]
    assert list(ims) == ["image_1", "image_1", "image_2", "image_1", "image_2", "image_2"]
    assert list(ts) == [1.0, 4.0, 5.0, 2.0, 0.5, 3.0] and list(extra) == [1, 4, 3, 2, 5, 0]


def test_load_glsn_csv_roundtrip(tmp_path):
    p = tmp_path / "sim.csv"
    p.write_text("time,band,flux,fluxerr,zp,zpsys,image\n"
                 "5.0,lssti,4.0,0.5,27,AB,image_2\n0.0,lssti,2.0,0.5,27,AB,image_1\n"
                 "1.0,F110W,6.0,1.0,27,AB,image_1\n3.0,F110W,10.0,1.0,27,AB,image_2\n2.0,F110W,7.0,1.0,27,AB,image_1\n")
    d = ingest.load_glsn_csv(str(p))
    assert list(d["band"]) == ["F110W", "F110W", "F110W", "lssti", "lssti"]
    assert list(d["x"]) == [1.0, 2.0, 3.0, 0.0, 5.0] and d["rescale"] == 8.0
    np.testing.assert_allclose(d["y"], np.array([6.0, 7.0, 10.0, 2.0, 4.0]) / 8.0)
    raw = ingest.load_glsn_csv(str(p), do_rescale=False)
    assert raw["rescale"] == 1.0 and list(raw["y"]) == [6.0, 7.0, 10.0, 2.0, 4.0]


def test_load_cosmograil_roundtrip(tmp_path):
    p = tmp_path / "q.txt"
    p.write_text("mhjd\tmag_A\tmagerr_A\tmag_B\tmagerr_B\ttelescope\n"
                 "10.0\t27.5\t0.1\t25.0\t0.1\tWFI\n5.0\t25.0\t0.2\t27.5\t0.1\tEulerCAM\n")
    d = ingest.load_cosmograil(str(p), images=("A", "B"))
    assert list(d["band"]) == ["EulerCAM", "EulerCAM", "WFI", "WFI"]
    assert list(d["image"]) == ["image_1", "image_2", "image_1", "image_2"]
    np.testing.assert_allclose(d["y"] * d["rescale"], [10.0, 1.0, 1.0, 10.0])
    e = ingest.load_cosmograil(str(p), images=("A", "B"), telescope="EulerCAM")
    assert len(e["x"]) == 2 and e["rescale"] == d["rescale"]     # factor from the whole table, as the notebook does


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference data only exist in the build container")
def test_loaders_reproduce_the_golden_datasets(golden):
    d = ingest.load_glsn_csv(f"{REF_DATA}/glSN_forGP_sims/simSN_107.csv")
    g = golden["datasets"]["simSN_107"]
    # the golden arrays were parsed by pandas' fast float parser, which may be one ulp off the correctly
    # rounded value Python's float() gives: flux, its range and their quotient each add one: compare to a few ulp
    for k in ("x", "y", "yerr"):
        np.testing.assert_allclose(d[k], g[k], rtol=2e-15, atol=0)
    assert list(d["band"]) == g["band"] and list(d["image"]) == g["image"]
    assert d["rescale"] == pytest.approx(g["rescale"], rel=2e-15)
    q = ingest.load_cosmograil(f"{REF_DATA}/COSMOGRAIL/DES_J0408-5354.txt", telescope="EulerCAM")
    gq = golden["datasets"]["J0408_euler"]
    np.testing.assert_allclose(q["x"], gq["x"], rtol=4.5e-16); np.testing.assert_allclose(q["y"], gq["y"], rtol=1e-15)
    np.testing.assert_allclose(q["yerr"], gq["yerr"], rtol=1e-15); assert q["rescale"] == pytest.approx(gq["rescale"], rel=1e-15)
