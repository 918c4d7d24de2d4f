"""Same module name as the reference's directionalvi/SharedDirectionalGradVariationalStrategy.py, which exports its
class under the name DirectionalGradVariationalStrategy (shared_directional_vi.py:13 imports it that way)."""
from dsvgp_b200.gp import SharedDirectionalGradVariationalStrategy
from dsvgp_b200.gp import SharedDirectionalGradVariationalStrategy as DirectionalGradVariationalStrategy  # noqa: F401
