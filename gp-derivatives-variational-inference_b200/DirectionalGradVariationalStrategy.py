"""Same module name as the reference's directionalvi/DirectionalGradVariationalStrategy.py."""
from dsvgp_b200.gp import DirectionalGradVariationalStrategy  # noqa: F401
