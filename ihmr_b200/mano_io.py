"""Read a MANO_{LEFT,RIGHT}.pkl (real or synthetic) into plain float32 numpy arrays.

Mirrors what smplx 0.1.28 ``MANO.__init__`` [UPSTREAM] extracts from the pickle the
reference passes at ``src/models/optimize_model.py:103-106``.  Real MANO pickles hold
``chumpy`` arrays and a scipy sparse ``J_regressor``; chumpy is not installed here, so a
stand-in class that only keeps the underlying ndarray is registered for unpickling.
"""
from __future__ import annotations

import pickle
from typing import Dict

import numpy as np


class _ChStub:
    """Unpickling stand-in for chumpy.ch.Ch: keeps the raw array in ``x``."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})

    def __array__(self, dtype=None, copy=None):
        arr = np.asarray(self.__dict__.get("x", self.__dict__.get("_x")))
        return arr.astype(dtype) if dtype is not None else arr


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("chumpy"):
            return _ChStub
        return super().find_class(module, name)


def _dense(a) -> np.ndarray:
    if hasattr(a, "toarray"):
        a = a.toarray()
    return np.asarray(a)


def load_mano_pkl(path: str) -> Dict[str, np.ndarray]:
    """Returns v_template (778,3), shapedirs (778,3,10), posedirs (135,2334) [M4 layout],
    J_regressor (16,778), lbs_weights (778,16), parents (16,) with parents[0] = -1,
    faces (1538,3) int64, hands_mean (45,)."""
    with open(path, "rb") as fh:
        raw = _Unpickler(fh, encoding="latin1").load()
    f32 = np.float32
    v_template = _dense(raw["v_template"]).astype(f32)
    nv = v_template.shape[0]
    shapedirs = _dense(raw["shapedirs"]).astype(f32)[:, :, :10]
    posedirs = _dense(raw["posedirs"]).astype(f32)          # (778, 3, 135)
    npf = posedirs.shape[-1]
    posedirs = np.ascontiguousarray(posedirs.reshape(nv * 3, npf).T)   # (135, 2334), M4
    parents = np.asarray(raw["kintree_table"])[0].astype(np.int64).copy()
    parents[0] = -1
    return dict(
        v_template=v_template,
        shapedirs=np.ascontiguousarray(shapedirs),
        posedirs=posedirs,
        J_regressor=_dense(raw["J_regressor"]).astype(f32),
        lbs_weights=_dense(raw["weights"]).astype(f32),
        parents=parents,
        faces=_dense(raw["f"]).astype(np.int64),
        hands_mean=_dense(raw["hands_mean"]).astype(f32).reshape(-1),
    )
