"""``pysted.base`` names the reference touches (``gym_sted/utils.py:157-180,616-657``)."""
from ..base import (Datamap, Detector, DonutBeam, Fluorescence, GaussianBeam, Microscope,  # noqa: F401
                    Objective)
