"""Host-side beam / detection-PSF formation.

This is the once-per-microscope set-up that ``pysted.base.Microscope.cache`` performs
(called from the reference at ``gym_sted/utils.py:164,178,616``): the excitation focus, the
STED donut and the detection PSF for *unit power* on the datamap pixel grid.  It runs once
per process (the reference pickles it to ``.microscope_cache.pkl``), so it stays on the host
in float64 numpy; everything that depends on the action (powers, dwell time) or on the
fluorophore is formed on the device by ``sted_make_effective`` (csrc/sted_kernels.cu).

pySTED is not vendored in the reference (``setup.py:13``) and is absent from the build box, so
the formulas are restated from the published theory pySTED cites:

* excitation: Richards & Wolf vectorial focus of a linearly polarised plane wave
  (integrals I0, I1, I2 over the aperture half angle);
* depletion: the same with a 0-2pi vortex phase and circular polarisation (five integrals,
  Hao et al. 2010), plus the ``zero_residual`` floor;
* detection: Gaussian emission PSF of FWHM lambda/(2 NA) integrated over each pixel,
  convolved with the circular pinhole (Willig 2006 eq. 3), max-normalised, times the
  objective transmission.

The quadrature here is a fixed 96-point Gauss-Legendre rule evaluated for every pixel at
once; the oracle (``oracle/pysted_oracle.py``) re-does it pixel by pixel with
``scipy.integrate.quad`` the way pySTED does, and the tests compare the two.
"""
from __future__ import annotations

import numpy as np
import scipy.constants
import scipy.special

H_PLANCK = scipy.constants.h
C_LIGHT = scipy.constants.c

_GL_ORDER = 96


def psf_side(lambda_: float, na: float, pixelsize: float) -> int:
    """Odd pixel count of a beam grid: ``int(2.233*lambda/(NA*px)/2)*2+1``."""
    diameter = 2.233 * lambda_ / (na * pixelsize)
    return int(diameter / 2) * 2 + 1


def _polar_grid(n_pixels: int):
    center = n_pixels // 2
    yy, xx = np.mgrid[0:n_pixels, 0:n_pixels]
    w_rel = (xx - center).astype(np.float64)
    h_rel = (center - yy).astype(np.float64)
    return np.arctan2(h_rel, w_rel), np.hypot(w_rel, h_rel)


def _aperture_nodes(alpha: float):
    x, w = np.polynomial.legendre.leggauss(_GL_ORDER)
    theta = 0.5 * alpha * (x + 1.0)
    return theta, 0.5 * alpha * w


def fwhm_pixels(profile: np.ndarray) -> float:
    """Full width at half maximum of a single-peaked 1-D profile, in (integer) pixels:
    distance between the samples nearest to half maximum on each side of the peak."""
    hm = profile.max() / 2.0
    i_max = int(np.argmax(profile))
    left = int(np.abs(profile[:i_max] - hm).argmin())
    right = int(np.abs(profile[i_max:] - hm).argmin()) + i_max
    return float(right - left)


def fwhm_donut_pixels(profile: np.ndarray):
    """Outer and inner full widths at half maximum of a symmetric two-crest profile."""
    hm = profile.max() / 2.0
    center = len(profile) // 2
    i_max = int(np.argmax(profile[: center + 1]))
    outer_l = int(np.abs(profile[:i_max] - hm).argmin())
    inner_l = int(np.abs(profile[i_max: center + 1] - hm).argmin()) + i_max
    return float(2 * (center - outer_l)), float(2 * (center - inner_l))


def gaussian_beam_intensity(lambda_, power, n, na, transmission, pixelsize,
                            polarization=np.pi / 2, beta=np.pi / 4) -> np.ndarray:
    """Excitation focus (W/m^2) for ``power`` watts, time averaged (``GaussianBeam.get_intensity``)."""
    n_pixels = psf_side(lambda_, na, pixelsize)
    angle, radius = _polar_grid(n_pixels)
    alpha = np.arcsin(na / n)
    k = 2.0 * np.pi * n / lambda_
    theta, w = _aperture_nodes(alpha)
    kr = (k * radius * pixelsize)[..., None]
    sq, s, c = np.sqrt(np.cos(theta)), np.sin(theta), np.cos(theta)
    i1 = (w * sq * s * (1 + c) * scipy.special.jv(0, kr * s)).sum(-1)
    i2 = (w * sq * s * s * scipy.special.jv(1, kr * s)).sum(-1)
    i3 = (w * sq * s * (1 - c) * scipy.special.jv(2, kr * s)).sum(-1)
    ax = np.sin(polarization)
    ay = np.cos(polarization) * np.exp(1j * beta)
    ex = -1j * (ax * (i1 + i3 * np.cos(2 * angle)) + ay * i3 * np.sin(2 * angle))
    ey = -1j * (ax * i3 * np.sin(2 * angle) + ay * (i1 - i3 * np.cos(2 * angle)))
    ez = -2 * i2 * (ax * np.cos(angle) + ay * np.sin(angle))
    intensity = np.abs(ex) ** 2 + np.abs(ey) ** 2 + np.abs(ez) ** 2
    intensity /= intensity.max()
    r = fwhm_pixels(intensity[n_pixels // 2])
    area_fwhm = np.pi * (r * pixelsize) ** 2 / 2.0
    return intensity * transmission * power / area_fwhm


def donut_beam_intensity(lambda_, power, n, na, transmission, pixelsize, tau, rate,
                         zero_residual=0.0, polarization=np.pi / 4, beta=np.pi / 2) -> np.ndarray:
    """STED donut (W/m^2) for ``power`` watts of *average* power.

    The returned map is the *instantaneous* (in-pulse) intensity: average / (tau*rate), which is
    what the reference's own restatement assumes (``gym_sted/utils.py:619``: "the instant sted
    intensity (instant p_sted = p_sted/(self.sted.tau * self.sted.rate))").
    """
    n_pixels = psf_side(lambda_, na, pixelsize)
    angle, radius = _polar_grid(n_pixels)
    alpha = np.arcsin(na / n)
    k = 2.0 * np.pi * n / lambda_
    theta, w = _aperture_nodes(alpha)
    kr = (k * radius * pixelsize)[..., None]
    sq, s, c = np.sqrt(np.cos(theta)), np.sin(theta), np.cos(theta)
    i1 = (w * sq * s * (1 + c) * scipy.special.jv(1, kr * s)).sum(-1)
    i2 = (w * sq * s * (1 - c) * scipy.special.jv(1, kr * s)).sum(-1)
    i3 = (w * sq * s * (1 - c) * scipy.special.jv(3, kr * s)).sum(-1)
    i4 = (w * sq * s * s * scipy.special.jv(0, kr * s)).sum(-1)
    i5 = (w * sq * s * s * scipy.special.jv(2, kr * s)).sum(-1)
    ax = np.sin(polarization)
    ay = np.cos(polarization) * np.exp(1j * beta)
    e1, em, e2, e3 = (np.exp(1j * angle), np.exp(-1j * angle), np.exp(2j * angle), np.exp(3j * angle))
    ex = ax * (i1 * e1 - i2 / 2 * em + i3 / 2 * e3) - ay * 0.5j * (i2 * em + i3 * e3)
    ey = -ax * 0.5j * (i2 * em + i3 * e3) + ay * (i1 * e1 + i2 / 2 * em - i3 / 2 * e3)
    ez = ax * 1j * (i4 - i5 * e2) - ay * (i4 + i5 * e2)
    intensity = np.abs(ex) ** 2 + np.abs(ey) ** 2 + np.abs(ez) ** 2
    intensity /= intensity.max()
    r_out, r_in = fwhm_donut_pixels(intensity[n_pixels // 2])
    area_fwhm = np.pi * (r_out * pixelsize) ** 2 / 2.0 - np.pi * (r_in * pixelsize) ** 2 / 2.0
    intensity = intensity * transmission * power / area_fwhm
    if power > 0 and zero_residual > 0:
        old_max = intensity.max()
        intensity = intensity + zero_residual * old_max
        intensity = intensity / intensity.max() * old_max
    return intensity / (tau * rate)


def emission_psf(lambda_, na, pixelsize) -> np.ndarray:
    """Pixel-integrated Gaussian emission PSF, FWHM = lambda/(2 NA) (``Fluorescence.get_psf``)."""
    n_pixels = psf_side(lambda_, na, pixelsize)
    center = n_pixels // 2
    sigma = lambda_ / (2.0 * na) / (2.0 * np.sqrt(2.0 * np.log(2.0)))
    edges = (np.arange(n_pixels + 1) - center - 0.5) * pixelsize
    cdf = 0.5 * (1.0 + scipy.special.erf(edges / (sigma * np.sqrt(2.0))))
    g = np.diff(cdf)
    return np.outer(g, g)


def pinhole_mask(radius, pixelsize, n_pixels) -> np.ndarray:
    center = n_pixels // 2
    yy, xx = np.mgrid[0:n_pixels, 0:n_pixels]
    return (((xx - center) ** 2 + (yy - center) ** 2) <= (radius / pixelsize) ** 2).astype(np.float64)


def convolve2d_same(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Direct 'same' 2-D convolution of two equal odd-sized squares (zero boundary)."""
    n = a.shape[0]
    p = n // 2
    ap = np.pad(a, p)
    out = np.zeros_like(a)
    bf = b[::-1, ::-1]
    for dy in range(n):
        for dx in range(n):
            wgt = bf[dy, dx]
            if wgt != 0.0:
                out += wgt * ap[dy:dy + n, dx:dx + n]
    return out


def detection_psf(lambda_, na, transmission, pixelsize, n_airy) -> np.ndarray:
    """``Detector.get_detection_psf``: emission PSF (*) pinhole, max-normalised, times transmission."""
    psf = emission_psf(lambda_, na, pixelsize)
    radius = n_airy * 0.61 * lambda_ / na
    pin = pinhole_mask(radius, pixelsize, psf.shape[0])
    det = convolve2d_same(psf, pin)
    return det / det.max() * transmission


def resize_common(*images: np.ndarray):
    """``pysted.utils.resize``: centre-pad every map to the largest (odd) side."""
    side = max(im.shape[0] for im in images)
    out = []
    for im in images:
        p = (side - im.shape[0]) // 2
        out.append(np.pad(im, p) if p else im.copy())
    return tuple(out)
