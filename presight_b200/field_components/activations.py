"""trunc_exp (reference: nerfstudio/field_components/activations.py:28-54) on the b200 kernel."""
from ..ops import trunc_exp  # noqa: F401
