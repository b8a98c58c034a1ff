"""Seeded synthetic MANO-shaped hand model and synthetic two-hand frames.

No MANO_{LEFT,RIGHT}.pkl exists offline, so every test and
benchmark runs on a seeded stand-in with the exact tensor shapes, key names and
topology counts of the real files the reference loads at
``src/models/optimize_model.py:103-106`` (778 vertices, 1538 faces, 16 joints,
MANO ``parents``, a 16-edge open wrist loop).  A real ``MANO_RIGHT.pkl`` can be
dropped into the same directory and is read through the same loader
(``ihmr_b200.mano_layer.load_mano_pkl``).

The frame generator follows SURVEY.md §8(d): frames are generated in fixed
blocks indexed by frame id, so sharding a range over ranks never changes a frame.
Everything here is numpy only; nothing here is on the timed path.
"""
from __future__ import annotations

import os
import pickle
from typing import Callable, Dict, Tuple

import numpy as np

NUM_VERTS = 778
NUM_FACES = 1538
NUM_JOINTS = 16
NUM_BETAS = 10
NUM_POSE_FEAT = 135
PARENTS = np.array([-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14], dtype=np.int64)
# fingertip vertex ids the reference appends to the 16 joints
# (src/models/optimize_model.py:99): thumb, index, middle, ring, pinky
TIP_IDS = np.array([744, 320, 443, 554, 671], dtype=np.int64)

FRAME_BLOCK = 256  # frames are generated in blocks of this many ids


# --------------------------------------------------------------------------- mesh
def _profile(x: np.ndarray, length: float) -> Tuple[np.ndarray, np.ndarray]:
    """Half-width (z) and half-thickness (y) of the mitten at abscissa x."""
    x0 = 0.58 * length
    a_w, a0 = 0.030, 0.046
    b_w, b0 = 0.013, 0.016
    t = np.clip(x / x0, 0.0, 1.0)
    s = t * t * (3 - 2 * t)
    a = a_w + (a0 - a_w) * s
    b = b_w + (b0 - b_w) * s
    cap = np.sqrt(np.clip(1.0 - ((x - x0) / (length - x0)) ** 2, 0.0, 1.0))
    a = np.where(x > x0, a0 * cap, a)
    b = np.where(x > x0, b0 * cap, b)
    return a, b


def _ring_counts(perims: np.ndarray, total: int, last: int) -> np.ndarray:
    """Integer vertex count per ring, proportional to perimeter, summing to total."""
    lo, hi = 1.0, 1e5
    for _ in range(200):
        c = 0.5 * (lo + hi)
        n = np.maximum(6, np.round(c * perims)).astype(int)
        n[-1] = last
        if n.sum() > total:
            hi = c
        else:
            lo = c
    n = np.maximum(6, np.round(lo * perims)).astype(int)
    n[-1] = last
    k = 0
    order = np.argsort(-perims[:-1])
    while n.sum() != total:
        i = order[k % len(order)]
        n[i] += 1 if n.sum() < total else -1
        k += 1
    return n


def _stitch(a_idx, a_ang, b_idx, b_ang):
    """Zipper-triangulate the band between two rings (angles increasing, one turn)."""
    n, m = len(a_idx), len(b_idx)
    # rotate ring b so it starts at the first angle >= a_ang[0]
    two_pi = 2 * np.pi
    rel = np.mod(b_ang - a_ang[0], two_pi)
    j0 = int(np.argmin(rel))
    b_idx = np.roll(b_idx, -j0)
    b_unw = np.sort(np.mod(np.roll(b_ang, -j0) - a_ang[0], two_pi))
    a_unw = np.mod(a_ang - a_ang[0], two_pi)
    a_unw = np.concatenate([a_unw, [two_pi + a_unw[0]]])
    b_unw = np.concatenate([b_unw, [two_pi + b_unw[0]]])
    tris = []
    i = j = 0
    while i < n or j < m:
        if j == m or (i < n and a_unw[i + 1] <= b_unw[j + 1]):
            tris.append((a_idx[i % n], a_idx[(i + 1) % n], b_idx[j % m]))
            i += 1
        else:
            tris.append((a_idx[i % n], b_idx[(j + 1) % m], b_idx[j % m]))
            j += 1
    return tris


def _build_mesh(rng: np.random.Generator):
    length = 0.185
    n_rings = 36
    # ring abscissae from near the tip (first) down to the wrist (last, x=0)
    u = (np.arange(n_rings) + 0.6) / (n_rings - 0.4)
    xs = length * np.cos(0.5 * np.pi * u) ** 0.85
    xs[-1] = 0.0
    a, b = _profile(xs, length)
    perims = np.pi * (3 * (a + b) - np.sqrt((3 * a + b) * (a + 3 * b)))
    counts = _ring_counts(perims, NUM_VERTS - 1, 16)

    verts = [np.array([length, 0.0, 0.0])]
    rings, angs = [], []
    for r in range(n_rings):
        n = counts[r]
        phase = rng.uniform(0, 1)
        ang = 2 * np.pi * (np.arange(n) + phase) / n
        # superellipse cross-section (a little boxy, like a palm)
        ca, sa = np.cos(ang), np.sin(ang)
        p = 2.6
        rad = (np.abs(ca) ** p + np.abs(sa) ** p) ** (-1.0 / p)
        ring = np.stack([np.full(n, xs[r]), b[r] * rad * sa, a[r] * rad * ca], axis=1)
        start = len(verts)
        verts.extend(list(ring))
        rings.append(np.arange(start, start + n))
        angs.append(ang)
    verts = np.array(verts)
    # deterministic jitter so no edge is axis-aligned
    verts[1:] += rng.normal(0, 2.5e-4, size=verts[1:].shape)

    faces = []
    n0 = len(rings[0])
    for i in range(n0):
        faces.append((0, rings[0][(i + 1) % n0], rings[0][i]))
    for r in range(n_rings - 1):
        faces.extend(_stitch(rings[r], angs[r], rings[r + 1], angs[r + 1]))
    faces = np.array(faces, dtype=np.int64)
    assert verts.shape == (NUM_VERTS, 3) and faces.shape == (NUM_FACES, 3), (verts.shape, faces.shape)

    # consistent outward orientation: signed volume (hole is small) must be > 0
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    vol = np.sum(np.einsum("ij,ij->i", v0, np.cross(v1, v2))) / 6.0
    if vol < 0:
        faces = faces[:, ::-1].copy()
    return verts, faces


def _joint_targets() -> Tuple[np.ndarray, np.ndarray]:
    """16 rest joints and 5 fingertip targets (thumb, index, middle, ring, pinky)."""
    J = np.zeros((16, 3))
    J[0] = [0.012, 0.0, 0.0]
    fingers = {  # joint base id -> (z offset, x of 3 joints, tip x)
        1: (0.030, [0.095, 0.125, 0.147], 0.166),   # index
        4: (0.010, [0.100, 0.133, 0.157], 0.178),   # middle
        7: (-0.032, [0.088, 0.112, 0.130], 0.146),  # little
        10: (-0.011, [0.097, 0.128, 0.150], 0.170),  # ring
    }
    tips = np.zeros((5, 3))
    tip_slot = {1: 1, 4: 2, 10: 3, 7: 4}
    for base, (z, xs, tx) in fingers.items():
        for k in range(3):
            J[base + k] = [xs[k], 0.002, z]
        tips[tip_slot[base]] = [tx, 0.0, z]
    # thumb sticks out on the +z side
    J[13] = [0.035, 0.004, 0.034]
    J[14] = [0.060, 0.004, 0.041]
    J[15] = [0.082, 0.004, 0.043]
    tips[0] = [0.102, 0.0, 0.044]
    return J, tips


def _seg_dist(p: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(float(ab @ ab), 1e-12), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[:, None] * ab), axis=1)


def make_mano_model(seed: int = 0) -> Dict[str, np.ndarray]:
    """Synthetic right-hand model with the key set of a MANO pickle.

    Keys/shapes as consumed by smplx 0.1.28 ``MANO.__init__`` [UPSTREAM]:
    v_template (778,3), shapedirs (778,3,10), posedirs (778,3,135),
    J_regressor (16,778), weights (778,16), kintree_table (2,16), f (1538,3),
    hands_mean (45,), hands_components (45,45).
    """
    rng = np.random.default_rng([seed, 7781538])
    verts, faces = _build_mesh(rng)
    J_t, tips = _joint_targets()

    # relabel so that the reference's hard-coded tip ids sit on the fingertips
    perm = np.arange(NUM_VERTS)  # perm[new_id] = old_id
    taken = set()
    for tid, tgt in zip(TIP_IDS, tips):
        d = np.linalg.norm(verts - tgt, axis=1)
        for old in np.argsort(d):
            if int(old) not in taken:
                break
        taken.add(int(old))
        cur = int(np.where(perm == old)[0][0])
        perm[[tid, cur]] = perm[[cur, tid]]
    inv = np.empty_like(perm)
    inv[perm] = np.arange(NUM_VERTS)
    verts = verts[perm]
    faces = inv[faces]

    # J_regressor: dense, rows sum to 1
    d2 = ((verts[None, :, :] - J_t[:, None, :]) ** 2).sum(-1)
    Jr = np.exp(-d2 / (2 * 0.018 ** 2))
    Jr /= Jr.sum(1, keepdims=True)

    # skinning weights: dense softmax of squared bone distance, rows sum to 1
    J_reg = Jr @ verts
    child = {0: None}
    for j in range(1, 16):
        child[PARENTS[j]] = child.get(PARENTS[j])
    ends = {}
    tip_of = {3: 1, 6: 2, 9: 4, 12: 3, 15: 0}
    for j in range(16):
        kids = [k for k in range(16) if PARENTS[k] == j]
        if j == 0:
            ends[j] = np.array([0.06, 0.0, 0.0])
        elif kids:
            ends[j] = J_reg[kids[0]]
        else:
            ends[j] = tips[tip_of[j]]
    dist = np.stack([_seg_dist(verts, J_reg[j], ends[j]) for j in range(16)], axis=1)
    logit = -(dist ** 2) / (2 * 0.011 ** 2)
    logit -= logit.max(1, keepdims=True)
    W = np.exp(logit)
    W /= W.sum(1, keepdims=True)

    # smooth basis over the surface for the blend shapes
    c = verts - verts.mean(0)
    sc = c / np.abs(c).max(0)
    basis = np.stack([np.ones(NUM_VERTS), sc[:, 0], sc[:, 1], sc[:, 2],
                      sc[:, 0] * sc[:, 2], np.sin(3 * sc[:, 0]), np.cos(3 * sc[:, 2]),
                      sc[:, 0] ** 2], axis=1)                       # (778, 8)
    coef = rng.normal(0, 1.0, size=(8, 3, NUM_BETAS))
    shapedirs = 0.002 * np.einsum("vb,bck->vck", basis, coef) / np.sqrt(8)
    coef_p = rng.normal(0, 1.0, size=(8, 3, NUM_POSE_FEAT))
    posedirs = 0.0005 * (0.7 * np.einsum("vb,bck->vck", basis, coef_p) / np.sqrt(8)
                         + 0.3 * rng.normal(0, 1.0, size=(NUM_VERTS, 3, NUM_POSE_FEAT)))
    hands_mean = rng.normal(0, 0.1, size=45)

    kintree = np.stack([np.where(PARENTS < 0, 2 ** 32 - 1, PARENTS), np.arange(16)]).astype(np.int64)
    return dict(
        v_template=verts.astype(np.float32),
        shapedirs=shapedirs.astype(np.float32),
        posedirs=posedirs.astype(np.float32),
        J_regressor=Jr.astype(np.float32),
        weights=W.astype(np.float32),
        kintree_table=kintree,
        f=faces.astype(np.int64),
        hands_mean=hands_mean.astype(np.float32),
        hands_components=np.eye(45, dtype=np.float32),
    )


def mirror_model(right: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Left-hand twin: x-mirrored geometry, reversed face winding."""
    left = {k: np.array(v, copy=True) for k, v in right.items()}
    left["v_template"][:, 0] *= -1
    left["shapedirs"][:, 0, :] *= -1
    left["posedirs"][:, 0, :] *= -1
    left["f"] = right["f"][:, ::-1].copy()
    return left


def write_mano_pkls(model_root: str, seed: int = 0) -> None:
    """Write MANO_RIGHT.pkl / MANO_LEFT.pkl where optimize_model.py:103 looks for them."""
    os.makedirs(model_root, exist_ok=True)
    right = make_mano_model(seed)
    for name, m in (("MANO_RIGHT.pkl", right), ("MANO_LEFT.pkl", mirror_model(right))):
        with open(os.path.join(model_root, name), "wb") as fh:
            pickle.dump(m, fh, protocol=2)


# ------------------------------------------------------------------------- frames
def _block(seed: int, block_id: int, mode: str) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng([seed, 1000, block_id])
    n = FRAME_BLOCK
    f32 = np.float32
    true_pose = np.zeros((n, 96))
    true_pose[:, 0:3] = rng.uniform(-np.pi, np.pi, (n, 3)) * 0.5
    true_pose[:, 48:51] = rng.uniform(-np.pi, np.pi, (n, 3)) * 0.5
    true_pose[:, 3:48] = np.clip(rng.normal(0, 0.3, (n, 45)), -1, 1)
    true_pose[:, 51:96] = np.clip(rng.normal(0, 0.3, (n, 45)), -1, 1)
    true_shape = np.clip(rng.normal(0, 1.0, (n, 20)), -2, 2)
    true_trans = rng.normal(0, 0.08, (n, 3))
    cam = np.concatenate([rng.uniform(4, 6, (n, 1)), rng.uniform(-0.1, 0.1, (n, 2))], axis=1)
    if mode == "collision":
        # mirrored identical poses, wrists ~1 cm apart: near-coincident hands (cfg5)
        true_pose[:, 48:96] = true_pose[:, 0:48]
        true_pose[:, 49::3] *= -1
        true_pose[:, 50::3] *= -1
        true_shape[:, 10:] = true_shape[:, :10]
        true_trans = rng.normal(0, 0.01, (n, 3))
    init_pose = true_pose + rng.normal(0, 0.1, (n, 96))
    init_shape = true_shape + rng.normal(0, 0.3, (n, 20))
    init_trans = true_trans + rng.normal(0, 0.02, (n, 3))
    noise_j3d = rng.normal(0, 0.005, (n, 42, 3))
    noise_j2d = rng.normal(0, 0.02, (n, 42, 2))
    return dict(true_pose=true_pose.astype(f32), true_shape=true_shape.astype(f32),
                true_trans=true_trans.astype(f32), cam=cam.astype(f32),
                init_pose=init_pose.astype(f32), init_shape=init_shape.astype(f32),
                init_trans=init_trans.astype(f32), noise_j3d=noise_j3d.astype(f32),
                noise_j2d=noise_j2d.astype(f32))


def make_raw_frames(start: int, count: int, seed: int = 0, mode: str = "typical") -> Dict[str, np.ndarray]:
    """Seeded per-frame parameters for frame ids [start, start+count)."""
    assert mode in ("typical", "collision")
    b0, b1 = start // FRAME_BLOCK, (start + count - 1) // FRAME_BLOCK
    blocks = [_block(seed, b, mode) for b in range(b0, b1 + 1)]
    out = {}
    lo = start - b0 * FRAME_BLOCK
    for k in blocks[0]:
        out[k] = np.concatenate([blk[k] for blk in blocks], axis=0)[lo:lo + count]
    out["index"] = np.arange(start, start + count, dtype=np.int64)
    return out


def finish_frames(raw: Dict[str, np.ndarray],
                  two_hand_forward: Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray]
                  ) -> Dict[str, np.ndarray]:
    """Turn raw parameters into the 17-key batch dict of src/data/opt_dataset.py:176-196.

    ``two_hand_forward(pose (B,96), shape (B,20), trans (B,3)) -> joints (B,42,3)`` is the
    IHMR two-hand MANO forward (oracle on CPU in tests, the CUDA path in the benchmark).
    """
    f32 = np.float32
    n = raw["true_pose"].shape[0]
    joints = np.asarray(two_hand_forward(raw["true_pose"], raw["true_shape"], raw["true_trans"]), dtype=f32)
    cam = raw["cam"]
    j2d = cam[:, None, 0:1] * (joints[:, :, :2] + cam[:, None, 1:3])
    ones = np.ones((n, 42, 1), f32)
    init_j3d = joints + raw["noise_j3d"]
    init_j2d = j2d + raw["noise_j2d"]
    w1 = np.ones((n, 1, 1), f32)
    hand_trans = (joints[:, 21] - joints[:, 0])[:, None, :]
    init_trans_j = (init_j3d[:, 21] - init_j3d[:, 0])[:, None, :]
    return dict(
        joints_2d=np.concatenate([j2d, ones], 2).astype(f32),
        joints_3d=np.concatenate([joints, ones], 2).astype(f32),
        mano_pose=raw["true_pose"].astype(f32),
        mano_betas=raw["true_shape"].astype(f32),
        mano_params_weight=np.ones((n, 2), f32),
        hand_trans=np.concatenate([hand_trans, w1], 2).astype(f32),
        hand_type_array=np.ones((n, 2), f32),
        hand_type_valid=np.ones((n, 1), f32),
        scale_ratio=np.ones((n,), f32),
        index=raw["index"].astype(np.int64),
        init_cam=cam.astype(f32),
        init_shape_params=raw["init_shape"].astype(f32),
        init_pose_params=raw["init_pose"].astype(f32),
        init_hand_trans=np.concatenate([raw["init_trans"][:, None, :], w1], 2).astype(f32),
        init_joints_2d=np.concatenate([init_j2d, ones], 2).astype(f32),
        init_joints_3d=np.concatenate([init_j3d, ones], 2).astype(f32),
        init_hand_trans_j=np.concatenate([init_trans_j, w1], 2).astype(f32),
    )
