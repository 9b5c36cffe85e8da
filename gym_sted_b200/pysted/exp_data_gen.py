"""``pysted.exp_data_gen`` surface used by ``gym_sted/utils.py:57-76``: ``Synapse``."""
from ..synapse import Synapse  # noqa: F401
