"""Drop-in for the reference's combination_layer.py."""
from fira_icse_b200.modules import CombinationLayer  # noqa: F401
