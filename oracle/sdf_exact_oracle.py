"""ORACLE (test infrastructure, not product code) — brute-force CPU statement of the EXACT (grid-free) penetration mode.

This mode is NOT part of the reference: SURVEY.md §8(f) rank 3 defines it as a separately named alternative to the
32^3 voxel field of ``sdf.SDFLoss`` (loss_utils.py:181-182).  With the conventions of Appendix B (bounding-box centre
c_h and scale s_h of the grid hand h, query hand o = 1 - h):

    p_v   = (V_o[v] - c_h) / s_h
    psi_v = dist(p_v, mesh(U_h, F_h))   if p_v is inside the mesh (odd number of +x ray crossings)   else 0
    loss  = ( sum_v psi_L[v] + sum_v psi_R[v] ) / 4 ,   origin_scale = psi * s_h (metres)
    d loss / d V_o[v] = (p_v - c*_v) / (4 psi_v s_h)    with c*_v the closest point of the mesh (c_h, s_h, mesh: no grad)

i.e. the limit of the reference's field for an infinitely fine grid.  The inside test uses the same arithmetic contract as
oracle/sdf_oracle.c and the CUDA kernels (fp32, no FMA contraction, edge functions on (low id, high id) ordering), evaluated
here with numpy float32 operations; distances and closest points are computed in float64.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _boxes(V):
    lo, hi = V.min(0), V.max(0)
    cen = ((lo + hi) * F32(0.5)).astype(F32)
    scale = F32(0.6) * (hi - lo).max().astype(F32)
    return cen, F32(scale)


def _inside(P, faces, q):
    """+x ray crossing parity of the fp32 point q against all faces (vectorised over faces, fp32 unfused)."""
    ia, ib, ic = faces[:, 0], faces[:, 1], faces[:, 2]

    def side(i0, i1):
        fwd = i0 < i1
        lo = np.where(fwd[:, None], P[i0], P[i1])
        hi = np.where(fwd[:, None], P[i1], P[i0])
        e = ((hi[:, 1] - lo[:, 1]) * (q[2] - lo[:, 2])).astype(F32) - ((hi[:, 2] - lo[:, 2]) * (q[1] - lo[:, 1])).astype(F32)
        e = e.astype(F32)
        return np.where(fwd, e >= 0, e < 0), np.where(fwd, e, -e).astype(F32)
    p0, wc = side(ia, ib)
    p1, wa = side(ib, ic)
    p2, wb = side(ic, ia)
    hit = (p0 & p1 & p2) | (~p0 & ~p1 & ~p2)
    s = ((wa + wb).astype(F32) + wc).astype(F32)
    hit &= s != 0
    with np.errstate(divide="ignore", invalid="ignore"):
        num = (((wa * P[ia, 0]).astype(F32) + (wb * P[ib, 0]).astype(F32)).astype(F32) + (wc * P[ic, 0]).astype(F32)).astype(F32)
        x = (num / s).astype(F32)
    return int(np.count_nonzero(hit & (x > q[0]))) % 2 == 1


def _closest_points(p, A, B, C):
    """closest point of every triangle (A,B,C rows) to p, float64 (Ericson, Real-Time Collision Detection 5.1.5)."""
    ab, ac, ap = B - A, C - A, p - A
    d1, d2 = (ab * ap).sum(1), (ac * ap).sum(1)
    bp = p - B
    d3, d4 = (ab * bp).sum(1), (ac * bp).sum(1)
    cp = p - C
    d5, d6 = (ab * cp).sum(1), (ac * cp).sum(1)
    vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
    out = np.empty_like(A)
    done = np.zeros(len(A), bool)

    def put(mask, val):
        m = mask & ~done
        out[m] = val[m]
        done[m] = True
    with np.errstate(divide="ignore", invalid="ignore"):
        put((d1 <= 0) & (d2 <= 0), A)
        put((d3 >= 0) & (d4 <= d3), B)
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), A + (d1 / (d1 - d3))[:, None] * ab)
        put((d6 >= 0) & (d5 <= d6), C)
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), A + (d2 / (d2 - d6))[:, None] * ac)
        put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), B + ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[:, None] * (C - B))
        den = 1.0 / (va + vb + vc)
        put(np.ones(len(A), bool), A + (vb * den)[:, None] * ab + (vc * den)[:, None] * ac)
    return out


def exact_penetration(hand_verts: np.ndarray, faces_right: np.ndarray, faces_left: np.ndarray):
    """hand_verts (B,2,778,3) fp32 -> losses (B), per_vert (B,1556) [right verts first], origin_scale (B,1556) metres,
    grad (B,2,778,3) = d losses[b] / d hand_verts[b]."""
    hv = np.asarray(hand_verts, F32)
    Bn, nv = hv.shape[0], hv.shape[2]
    faces = (np.asarray(faces_right), np.asarray(faces_left))
    losses = np.zeros(Bn)
    per_vert, origin = np.zeros((Bn, 2 * nv)), np.zeros((Bn, 2 * nv))
    grad = np.zeros(hv.shape)
    for b in range(Bn):
        for h in (0, 1):
            o = 1 - h
            cen, scale = _boxes(hv[b, h])
            U = ((hv[b, h] - cen) / scale).astype(F32)
            P = ((hv[b, o] - cen) / scale).astype(F32)
            lo, hi = U.min(0), U.max(0)
            F = faces[h]
            A, Bv, C = U[F[:, 0]].astype(np.float64), U[F[:, 1]].astype(np.float64), U[F[:, 2]].astype(np.float64)
            for v in range(nv):
                q = P[v]
                if q[0] > hi[0] or q[1] < lo[1] or q[1] > hi[1] or q[2] < lo[2] or q[2] > hi[2]:
                    continue                   # a +x ray from here cannot cross the mesh an odd number of times
                if not _inside(U, F, q):
                    continue
                cp = _closest_points(q.astype(np.float64), A, Bv, C)
                d2 = ((cp - q.astype(np.float64)) ** 2).sum(1)
                k = int(np.argmin(d2))
                psi = float(np.sqrt(d2[k]))
                per_vert[b, o * nv + v] = psi
                origin[b, o * nv + v] = psi * float(scale)
                losses[b] += psi / 4.0
                if psi > 0:
                    grad[b, o, v] = (q.astype(np.float64) - cp[k]) / (4.0 * psi * float(scale))
    return losses, per_vert, origin, grad
