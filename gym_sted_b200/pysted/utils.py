"""``pysted.utils`` helpers visible from the reference: ``resize`` (inside ``Microscope.cache``), ``fwhm`` /
``fwhm_donut`` (beam normalisation) and ``mse_calculator`` (``gym_sted/envs/sted_env.py:423``)."""
import numpy as np

from ..physics import fwhm_donut_pixels as fwhm_donut  # noqa: F401
from ..physics import fwhm_pixels as fwhm  # noqa: F401
from ..physics import pinhole_mask as pinhole  # noqa: F401
from ..physics import resize_common as resize  # noqa: F401


def mse_calculator(array1, array2):
    """Mean squared error of two equally shaped arrays."""
    a, b = np.asarray(array1, dtype=np.float64), np.asarray(array2, dtype=np.float64)
    return float(np.mean((a - b) ** 2))
