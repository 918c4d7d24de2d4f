"""Same module name as the reference's directionalvi/GradVariationalStrategy.py."""
from dsvgp_b200.gp import GradVariationalStrategy  # noqa: F401
