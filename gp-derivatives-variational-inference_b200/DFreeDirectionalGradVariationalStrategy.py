"""Same module name as the reference's directionalvi/DFreeDirectionalGradVariationalStrategy.py, which exports its
class under the name DirectionalGradVariationalStrategy (dfree_directional_vi.py:14 imports it that way)."""
from dsvgp_b200.gp import DFreeDirectionalGradVariationalStrategy
from dsvgp_b200.gp import DFreeDirectionalGradVariationalStrategy as DirectionalGradVariationalStrategy  # noqa: F401
