"""The C-ABI library loads and exports every symbol include/ihmr_b200.h declares (no compute
calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from ihmr_b200 import _lib
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ihmr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ihmr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from ihmr_b200 import _lib
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert set(_lib.EXPORTS) == set(names)


def test_version_and_error_conventions(lib):
    assert lib.ihmr_abi_version() == 1
    assert lib.ihmr_mano_workspace_bytes(0) == 0
    assert lib.ihmr_mano_workspace_bytes(16) > 16 * 2334 * 4
    assert lib.ihmr_opt_workspace_bytes(4) > lib.ihmr_mano_workspace_bytes(8)
    # invalid arguments are reported by status + thread-local message, nothing is dereferenced
    rc = lib.ihmr_mano_forward(None, 1, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"invalid argument" in lib.ihmr_last_error()
    handle = ctypes.c_void_p()
    rc = lib.ihmr_model_create(*([None] * 9), 0, ctypes.byref(handle))
    assert rc == -1 and not handle.value


def test_stage_struct_matches_header_and_strategy_dicts():
    from ihmr_b200 import _lib
    from ihmr_b200.strategies import opt_default, strategies
    assert ctypes.sizeof(_lib.Stage) == 4 * (3 + 6 + 1 + 4 + 4 + 1 + 1)
    assert ctypes.sizeof(_lib.Targets) == 5 * ctypes.sizeof(ctypes.c_void_p)
    st = _lib.make_stage(opt_default[2])
    assert st.update_mask == (_lib.P_L_POSE | _lib.P_R_POSE) and st.epoch == 300
    assert abs(st.w_finger_reg - 1e5) < 1e-3 and st.n_filters == 2
    assert list(st.filter_loss)[:2] == [0, 1] and list(st.filter_percent)[:2] == [0.0, -10.0]
    assert [s["epoch"] for s in strategies["opt_fixed100"]] == [24] * 4
    with pytest.raises(_lib.IhmrError):
        _lib.make_stage(dict(opt_default[0], select_loss="joints_3d_loss"))     # opt_utils.py:57-67


def test_no_cpu_fallback(model_root):
    """The product path must fail loudly without a GPU instead of computing on the CPU."""
    import torch
    from ihmr_b200 import _lib, mano_layer, sdf_loss
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    layer = mano_layer.create(os.path.join(model_root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True)
    with pytest.raises(_lib.IhmrError):
        layer(global_orient=torch.zeros(1, 3), hand_pose=torch.zeros(1, 45), betas=torch.zeros(1, 10))
    with pytest.raises(_lib.IhmrError):
        sdf_loss.SDFLoss(layer.faces, layer.faces)(torch.zeros(1, 2, 778, 3))
    from ihmr_b200.optimize_model import OptimizeModel
    from tests import helpers as H
    with pytest.raises(_lib.IhmrError):
        OptimizeModel(H.make_opt(model_root, 1))
