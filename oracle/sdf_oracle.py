"""ORACLE (test infrastructure, not product code) — CPU restatement of ``sdf.SDFLoss``.

PARITY UNPINNED: the reference imports ``SDFLoss`` from the un-vendored, un-pinned package
``sdf`` (github.com/penincillin/SDF_ihmr @ git HEAD, /root/reference/docs/install.md:37;
import at src/models/loss_utils.py:13, construction :34-38, call :181-182).  This restates
``sdf/sdf_loss.py::SDFLoss.forward`` of that package's lineage under assumptions A1-A8 of
SURVEY.md §8(c) / Appendix B.  What the in-tree reference code does pin, and what this module
honours: constructor ``SDFLoss(faces_right, faces_left, robustifier=None)``; call
``(hand_verts (B,2,778,3), return_per_vert_loss=True, return_origin_scale_loss=True)`` ->
``(losses (B,), per_vert, origin_scale (B,1556))`` with origin_scale in metres, right-hand
vertices first (src/utils/evaluator.py:119-120,169).

The voxel field itself is computed by the brute-force C restatement ``sdf_oracle.c``
(built by oracle/Makefile); a pure-numpy twin is kept for cross-checking on tiny grids.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libsdf_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(path)
        for suf in ("f32", "f64"):
            getattr(lib, f"sdf_grid_{suf}").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                       ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                       ctypes.c_void_p]
            getattr(lib, f"sdf_grid_{suf}").restype = None
        _LIB = lib
    return _LIB


def sdf_grid(verts: torch.Tensor, faces: torch.Tensor, grid_size: int = 32) -> torch.Tensor:
    """phi (n, G, G, G) [z][y][x] for n meshes with normalised vertices (n, nv, 3).
    Restates the `sdf_cuda.sdf(phi, faces, vertices)` extension entry point (A1, A3, A4)."""
    assert verts.dim() == 3 and verts.shape[2] == 3
    v = verts.detach().contiguous().cpu()
    f = faces.detach().to(torch.int32).contiguous().cpu()
    n, nv = v.shape[:2]
    phi = torch.empty(n, grid_size, grid_size, grid_size, dtype=v.dtype)
    fn = _lib().sdf_grid_f32 if v.dtype == torch.float32 else _lib().sdf_grid_f64
    fn(v.data_ptr(), f.data_ptr(), n, nv, f.shape[0], grid_size, phi.data_ptr())
    return phi


# ------------------------------------------------------------------ numpy twin (tiny cases)
def _ray_cross_np(P, face, q):
    def side(i0, i1):
        fwd = i0 < i1
        lo, hi = (P[i0], P[i1]) if fwd else (P[i1], P[i0])
        e = (hi[1] - lo[1]) * (q[2] - lo[2]) - (hi[2] - lo[2]) * (q[1] - lo[1])
        return ((e >= 0) if fwd else (e < 0)), (e if fwd else -e)
    ia, ib, ic = face
    p0, wc = side(ia, ib)
    p1, wa = side(ib, ic)
    p2, wb = side(ic, ia)
    if not ((p0 and p1 and p2) or (not p0 and not p1 and not p2)):
        return 0
    s = (wa + wb) + wc
    if s == 0:
        return 0
    x = ((wa * P[ia][0] + wb * P[ib][0]) + wc * P[ic][0]) / s
    return int(x > q[0])


def _pt_tri_dist2_np(p, a, b, c):
    """Brute-force independent formulation: min over the face interior projection and the 3 edges."""
    def seg(p, u, v):
        d = v - u
        t = np.clip(np.dot(p - u, d) / max(np.dot(d, d), 1e-300), 0, 1)
        r = p - (u + t * d)
        return np.dot(r, r)
    best = min(seg(p, a, b), seg(p, b, c), seg(p, c, a))
    n = np.cross(b - a, c - a)
    nn_ = np.dot(n, n)
    if nn_ > 0:
        t = np.dot(p - a, n) / nn_
        proj = p - t * n
        # barycentric inside test
        c0 = np.dot(np.cross(b - a, proj - a), n)
        c1 = np.dot(np.cross(c - b, proj - b), n)
        c2 = np.dot(np.cross(a - c, proj - c), n)
        if c0 >= 0 and c1 >= 0 and c2 >= 0:
            best = min(best, t * t * nn_)
    return best


def sdf_grid_numpy(verts: np.ndarray, faces: np.ndarray, grid_size: int) -> np.ndarray:
    """Slow reference of the reference: pure-Python loops, float64, for tiny grids/meshes only."""
    G = grid_size
    P = np.asarray(verts, dtype=np.float64)
    phi = np.zeros((G, G, G))
    for zi in range(G):
        for yi in range(G):
            for xi in range(G):
                q = np.array([(2 * xi + 1 - G) / G, (2 * yi + 1 - G) / G, (2 * zi + 1 - G) / G])
                cross = sum(_ray_cross_np(P, f, q) for f in faces)
                if cross % 2 == 1:
                    phi[zi, yi, xi] = np.sqrt(min(_pt_tri_dist2_np(q, P[f[0]], P[f[1]], P[f[2]]) for f in faces))
    return phi


# --------------------------------------------------------------------------- the loss module
class SDFLoss(nn.Module):
    """Restatement of sdf.SDFLoss for two hands (Appendix B of SURVEY.md)."""

    def __init__(self, faces_right, faces_left, grid_size=32, robustifier=None, debugging=False, ray_axis=0):
        super().__init__()
        # A4 audit knob (not a reference argument): world axis of the parity ray.  The field, the boxes and the
        # trilinear sampling are symmetric under a cyclic permutation of the coordinates, so a ray along axis a is
        # the +x statement below applied to (w[a], w[a+1], w[a+2]).
        self.ray_axis = int(ray_axis)
        self.register_buffer("faces_right", torch.tensor(np.asarray(faces_right).astype(np.int32)))
        self.register_buffer("faces_left", torch.tensor(np.asarray(faces_left).astype(np.int32)))
        self.grid_size = grid_size
        self.robustifier = robustifier

    @torch.no_grad()
    def boxes(self, hand_verts, scale_factor=0.2):
        lo = hand_verts.min(dim=2)[0]                       # (B,2,3)
        hi = hand_verts.max(dim=2)[0]
        center = (lo + hi) * 0.5                           # A2: bbox midpoint
        scale = (1 + scale_factor) * 0.5 * (hi - lo).max(dim=-1)[0]   # (B,2)
        return center, scale

    @torch.no_grad()
    def grids(self, hand_verts, center, scale):
        B = hand_verts.shape[0]
        U = (hand_verts - center[:, :, None, :]) / scale[:, :, None, None]
        phi_r = sdf_grid(U[:, 0], self.faces_right, self.grid_size)
        phi_l = sdf_grid(U[:, 1], self.faces_left, self.grid_size)
        return torch.stack([phi_r, phi_l], dim=1)          # (B,2,G,G,G)

    def forward(self, hand_verts, return_per_vert_loss=False, return_origin_scale_loss=False,
                scale_factor=0.2):
        B = hand_verts.shape[0]
        if self.ray_axis:
            hand_verts = hand_verts[..., [(self.ray_axis + c) % 3 for c in range(3)]]
        center, scale = self.boxes(hand_verts, scale_factor)
        phi = self.grids(hand_verts, center, scale)
        psi = [None, None]                                 # psi[o] = hand o sampled in the other grid
        for h in (0, 1):
            o = 1 - h
            p = (hand_verts[:, o] - center[:, h, None, :]) / scale[:, h, None, None]   # A5
            val = F.grid_sample(phi[:, h][:, None], p.view(B, -1, 1, 1, 3), mode="bilinear",
                                padding_mode="zeros", align_corners=False).view(B, -1)  # A3
            psi[o] = val
        if self.robustifier:
            rob = []
            for v in psi:
                frac = (v / self.robustifier) ** 2
                rob.append(frac / (frac + 1))
            cur = rob
        else:
            cur = psi
        per_vert = torch.cat([cur[0], cur[1]], dim=1)                     # right verts first
        losses = (cur[0].sum(dim=1) + cur[1].sum(dim=1)) / 4.0            # A6: / n_hands**2
        origin = torch.cat([psi[0] * scale[:, 1:2], psi[1] * scale[:, 0:1]], dim=1).detach()  # A7
        if return_per_vert_loss and return_origin_scale_loss:
            return losses, per_vert, origin
        if return_per_vert_loss:
            return losses, per_vert
        if return_origin_scale_loss:
            return losses, origin
        return losses


class SDFLoss_Single(SDFLoss):
    """Name exported by the package and imported at src/models/loss_utils.py:13; unused on the
    IHMR-OPT path."""
