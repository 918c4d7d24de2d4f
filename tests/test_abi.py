"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/dsvgp_b200.h declares; the Python mirror keeps the reference's names, signatures and state-dict keys;
the product never imports the oracle and refuses to run without CUDA."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gp-derivatives-variational-inference_b200")


def test_library_exports_every_declared_symbol():
    from dsvgp_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    protos = _lib.parse_header()
    assert len(protos) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/dsvgp_b200.h but not exported"
    assert lib.dsvgp_version() >= 100 and lib.dsvgp_built_for_sm() == 100
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dsvgp_\w+)", out))
    assert exported == set(protos), exported ^ set(protos)


def test_library_is_built_for_sm100a_only():
    from dsvgp_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_chol_plan_host_side():
    from dsvgp_b200 import _lib
    for Mq in (1, 60, 96, 113, 1024, 3072, 3200, 5000):
        Mp, nb0, nlev = _lib.chol_plan(Mq)
        assert Mp == nb0 << nlev and Mp >= Mq and nb0 <= 112 and nb0 % 4 == 0
        assert Mp < 1.15 * Mq + 8


def test_cpu_tensors_are_refused():
    from dsvgp_b200 import _lib, ops
    with pytest.raises(_lib.DsvgpError):
        ops.hyp_from_raw(torch.zeros(1))


def test_state_dict_keys_match_reference():
    import directional_vi
    from dsvgp_b200 import gp
    m = directional_vi.GPModel(torch.rand(5, 3), torch.eye(3)[:2].repeat(5, 1), 3)
    lik = gp.GaussianLikelihood()
    want = {"variational_strategy.inducing_points", "variational_strategy.inducing_directions",
            "variational_strategy.updated_strategy", "variational_strategy.variational_params_initialized",
            "variational_strategy._variational_distribution.variational_mean",
            "variational_strategy._variational_distribution.chol_variational_covar", "mean_module.constant",
            "covar_module.raw_outputscale", "covar_module.base_kernel.raw_lengthscale"}
    assert set(m.state_dict()) == want
    assert set(lik.state_dict()) == {"noise_covar.raw_noise"}
    assert m.num_inducing == 5 and m.num_directions == 2
    assert [n.split(".")[-1] for n, _ in m.named_variational_parameters()] == ["variational_mean", "chol_variational_covar"]
    assert m.covar_module.base_kernel.raw_lengthscale.shape == (1, 1) and m.covar_module.raw_outputscale.dim() == 0
    # old checkpoints without the flag load with a warning and get updated_strategy = False (DGVS.py:17-29)
    sd = {k: v for k, v in m.state_dict().items() if not k.endswith("updated_strategy")}
    with pytest.warns(gp.OldVersionWarning):
        m.load_state_dict(sd)
    assert not bool(m.variational_strategy.updated_strategy)


def test_signatures_match_reference():
    import dfree_directional_vi
    import directional_vi
    import grad_svgp
    from dsvgp_b200 import gp
    assert list(inspect.signature(directional_vi.train_gp).parameters)[:8] == [
        "train_dataset", "num_inducing", "num_directions", "minibatch_size", "minibatch_dim", "num_epochs",
        "learning_rate_hypers", "learning_rate_ngd"]
    assert list(inspect.signature(directional_vi.eval_gp).parameters) == [
        "test_dataset", "model", "likelihood", "mll_type", "num_directions", "minibatch_size", "minibatch_dim"]
    assert list(inspect.signature(grad_svgp.train_gp).parameters)[:2] == ["train_dataset", "dim"]
    assert list(inspect.signature(gp.DirectionalGradVariationalStrategy.__init__).parameters)[1:] == [
        "model", "inducing_points", "inducing_directions", "variational_distribution", "learn_inducing_locations"]
    assert list(inspect.signature(gp.GradVariationalStrategy.__init__).parameters)[1:] == [
        "model", "inducing_points", "variational_distribution", "learn_inducing_locations"]
    assert list(inspect.signature(gp.DirectionalGradVariationalStrategy.forward).parameters)[1:5] == [
        "x", "inducing_points", "inducing_values", "variational_inducing_covar"]
    import DFreeDirectionalGradVariationalStrategy as dfree_mod
    assert dfree_mod.DirectionalGradVariationalStrategy is gp.DFreeDirectionalGradVariationalStrategy
    assert dfree_directional_vi.GPModel.strategy_class is gp.DFreeDirectionalGradVariationalStrategy
    y, dirs = directional_vi.select_cols_of_y(torch.arange(12.).reshape(3, 4), 2, 3)
    assert y.shape == (3, 3) and dirs.shape == (2, 3) and torch.equal(dirs.sum(1), torch.ones(2))


def test_asserts_of_the_reference_are_kept():
    import directional_vi
    with pytest.raises(AssertionError):
        directional_vi.train_gp(None, num_directions=2, minibatch_dim=1)
    m = directional_vi.GPModel(torch.rand(4, 2), torch.eye(2).repeat(4, 1), 2)
    with pytest.raises(AssertionError):      # data directions per point != inducing directions per point (DGVS.py:106)
        m.variational_strategy.forward(torch.rand(3, 2), None, None, derivative_directions=torch.eye(2)[:1].repeat(3, 1))
    with pytest.raises(NotImplementedError):
        directional_vi.GPModel(torch.rand(4, 2), torch.eye(2).repeat(4, 1), 2, variational_strategy="CIQ")


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "dsvgp_oracle" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_missing_library_fails_loudly(tmp_path):
    code = ("import sys, importlib.util, os\n"
            f"sys.path.insert(0, {PKG!r})\n"
            "import dsvgp_b200._lib as L\n")
    env = dict(os.environ)
    # simulate a missing extension by pointing the loader at a copy of the package without the .so
    import shutil
    dst = tmp_path / "pkg"
    shutil.copytree(os.path.join(PKG, "dsvgp_b200"), dst / "dsvgp_b200", ignore=shutil.ignore_patterns("*.so", "__pycache__"))
    os.makedirs(tmp_path / "include")
    shutil.copy(os.path.join(ROOT, "include", "dsvgp_b200.h"), tmp_path / "include" / "dsvgp_b200.h")
    code = f"import sys\nsys.path.insert(0, {str(dst)!r})\nimport dsvgp_b200\n"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "no CPU or PyTorch fallback" in r.stderr
