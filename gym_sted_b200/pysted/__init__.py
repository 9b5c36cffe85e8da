"""``pysted``-shaped module tree over the CUDA hot path.

gym-sted imports ``from pysted import base, utils`` and ``from pysted import exp_data_gen as dg``
(``gym_sted/utils.py:11-12``) and ``import pysted.base`` (``gym_sted/envs/timed_sted_env.py:8``).
``gym_sted_b200.install_as_pysted()`` registers this package under the name ``pysted`` so the
unmodified reference runs on ``libsted_b200.so`` (exercised by
``tests/test_reference_dropin.py``).
"""
from . import base, exp_data_gen, utils  # noqa: F401
