"""Loaders for the two data formats the reference ships, with the notebooks' preparation steps
(SURVEY.md section 8f, rank 4).

``GP._prepare_indices`` only COUNTS rows per (band, image) (gaussn/gausSN.py:84-91); rows that are not
already sorted by band, then image, silently get the wrong delays and magnifications.  The notebooks sort
by (band, image, time) and divide flux and error by ``max(flux) - min(flux)`` before fitting
(ipynb/fitting_SLSN_example.ipynb cell 3, ipynb/fitting_glQSO_example.ipynb cell 5); these helpers do exactly
that and return plain numpy arrays ready for ``GP.optimize_parameters``.
"""
from __future__ import annotations

import csv

import numpy as np


def get_flux(mag, magerr, zp=27.5):
    """Magnitudes to fluxes: ``10**((mag - zp) / -2.5)``, ``flux * magerr * ln(10) / 2.5``
    (ipynb/fitting_glQSO_example.ipynb cell 5)."""
    mag, magerr = np.asarray(mag, dtype=np.float64), np.asarray(magerr, dtype=np.float64)
    flux = 10 ** ((mag - zp) / -2.5)
    return flux, flux * magerr * np.log(10) / 2.5


def sort_rows(time, band, image, *columns):
    """Stable sort by (band, image, time), the order ``_prepare_indices`` assumes.  Returns the sorted
    (time, band, image, *columns)."""
    time, band, image = np.asarray(time, dtype=np.float64), np.asarray(band), np.asarray(image)
    order = np.lexsort((time, image, band))
    return (time[order], band[order], image[order], *[np.asarray(c)[order] for c in columns])


def rescale(flux, fluxerr):
    """Divide by ``max(flux) - min(flux)`` so the light curve spans one unit (gaussn/gausSN.py:120-127);
    returns (flux, fluxerr, factor)."""
    flux, fluxerr = np.asarray(flux, dtype=np.float64), np.asarray(fluxerr, dtype=np.float64)
    factor = float(np.max(flux) - np.min(flux))
    return flux / factor, fluxerr / factor, factor


def _finish(time, band, image, flux, fluxerr, do_rescale, extra=None):
    cols = [] if extra is None else list(extra.values())
    out = sort_rows(time, band, image, flux, fluxerr, *cols)
    time, band, image, flux, fluxerr = out[:5]
    factor = 1.0
    if do_rescale:
        flux, fluxerr, factor = rescale(flux, fluxerr)
    data = dict(x=time, y=flux, yerr=fluxerr, band=band, image=image, rescale=factor)
    if extra is not None:
        for k, v in zip(extra.keys(), out[5:]):
            data[k] = v
    return data


def load_glsn_csv(path, do_rescale=True):
    """A ``data/glSN_forGP_sims/simSN_*.csv`` file: columns ``time,band,flux,fluxerr,zp,zpsys,image``."""
    with open(path, newline="") as f:
        rows = list(csv.DictReader(f))
    col = lambda k: np.array([r[k] for r in rows])
    return _finish(col("time").astype(float), col("band"), col("image"), col("flux").astype(float), col("fluxerr").astype(float),
                   do_rescale, dict(zp=col("zp").astype(float), zpsys=col("zpsys")))


def load_cosmograil(path, images=("A", "B", "D"), telescope=None, zp=27.5, do_rescale=True):
    """A COSMOGRAIL light-curve table (``data/COSMOGRAIL/DES_J0408-5354.txt``): tab separated
    ``mhjd, mag_X, magerr_X, ..., telescope``.  Image k of `images` becomes ``image_{k+1}``, the telescope is the
    band.  The rescale factor is taken over ALL telescopes before `telescope` is selected, as the notebook does."""
    with open(path, newline="") as f:
        rows = list(csv.DictReader(f, delimiter="\t"))
    t = np.array([float(r["mhjd"]) for r in rows])
    tel = np.array([r["telescope"] for r in rows])
    time, band, image, flux, err = [], [], [], [], []
    for k, tag in enumerate(images):
        fl, fe = get_flux([float(r[f"mag_{tag}"]) for r in rows], [float(r[f"magerr_{tag}"]) for r in rows], zp)
        time.append(t); band.append(tel); image.append(np.repeat(f"image_{k + 1}", len(rows))); flux.append(fl); err.append(fe)
    data = _finish(np.concatenate(time), np.concatenate(band), np.concatenate(image), np.concatenate(flux), np.concatenate(err), do_rescale)
    if telescope is not None:
        keep = data["band"] == telescope
        for k in ("x", "y", "yerr", "band", "image"):
            data[k] = data[k][keep]
    return data
