"""Same module name as the reference's directionalvi/RBFKernelDirectionalGrad.py, so
`from RBFKernelDirectionalGrad import RBFKernelDirectionalGrad` keeps working; the class is CUDA-backed."""
from dsvgp_b200.gp import RBFKernelDirectionalGrad  # noqa: F401
