"""Property tests (hypothesis): size-independent invariants of the path, on the oracle (CPU) and on the CUDA kernels.

The penetration field is built in the box-normalised frame of each hand, so the op is invariant
under a similarity transform of the whole frame (loss unchanged, origin-scale values times s, gradient divided by s) and
under a cyclic relabelling of the world axes when the parity ray is relabelled with them (bitwise: every component goes
through the same arithmetic).  MANO is equivariant under a rotation of the global orientation about the wrist.
"""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from ihmr_b200 import dist as idist
from ihmr_b200 import synthetic
from oracle import mano_oracle, sdf_oracle

# derandomize: the same examples in every run (the driver's runs are then the runs these were developed against)
COMMON = dict(deadline=None, derandomize=True, database=None,
              suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


# ------------------------------------------------------------------------------------------ host logic (CPU)
@settings(max_examples=200, **COMMON)
@given(total=st.integers(0, 100000), world=st.integers(1, 64))
def test_shards_partition_the_frames(total, world):
    """dist.shard_range: contiguous, ordered, complete, sizes within one frame of each other (optimize.py:40-47's
    DistributedSampler split, without its padding)."""
    pos, sizes = 0, []
    for r in range(world):
        start, count = idist.shard_range(total, r, world)
        assert start == pos and count >= 0
        pos += count
        sizes.append(count)
    assert pos == total and max(sizes) - min(sizes) <= 1


@settings(max_examples=25, **COMMON)
@given(start=st.integers(0, 5000), count=st.integers(1, 9), split=st.integers(1, 8))
def test_synthetic_frames_are_a_function_of_the_frame_id(start, count, split):
    """Frame k is the same whichever shard generates it (what lets every rank build its own inputs)."""
    whole = synthetic.make_raw_frames(start, count, seed=0)
    cut = min(split, count)
    a, b = synthetic.make_raw_frames(start, cut, seed=0), synthetic.make_raw_frames(start + cut, count - cut, seed=0)
    for k in ("true_pose", "true_shape", "true_trans"):
        assert np.array_equal(whole[k], np.concatenate([a[k], b[k]]))


def _frames(oracle_layers, mode, start, count):
    raw = synthetic.make_raw_frames(start, count, seed=0, mode=mode)
    with torch.no_grad():
        rv, lv, _ = mano_oracle.two_hand_forward(oracle_layers[0], torch.tensor(raw["true_pose"]), torch.tensor(raw["true_shape"]),
                                                 torch.tensor(raw["true_trans"]))
    return torch.stack([rv, lv], 1)


@settings(max_examples=6, **COMMON)
@given(start=st.integers(0, 2000), log2s=st.integers(-2, 2), t=st.tuples(*[st.floats(-0.5, 0.5, width=32)] * 3))
def test_oracle_penetration_is_similarity_invariant(oracle_layers, start, log2s, t):
    hv = _frames(oracle_layers, "collision", start, 1)
    ref = sdf_oracle.SDFLoss(oracle_layers[0].faces, oracle_layers[1].faces)
    l0, _, o0 = ref(hv, True, True)
    s = 2.0 ** log2s
    l1, _, o1 = ref(hv * s + torch.tensor(t, dtype=torch.float32), True, True)
    # a translation re-rounds the box-normalised coordinates: a voxel within an ulp of the surface may change side
    assert float((l1 - l0).abs().max()) <= 2e-3 * max(float(l0.abs().max()), 1e-3)
    assert float((o1 - o0 * s).abs().max()) <= 5e-4 * s * max(float(o0.abs().max()), 1e-3)


@settings(max_examples=8, **COMMON)
@given(seed=st.integers(0, 10000), axis_angle=st.tuples(*[st.floats(-2.0, 2.0, width=32)] * 3))
def test_oracle_mano_global_rotation_equivariance(oracle_layers, seed, axis_angle):
    """Composing the global orientation with a rotation Q turns the mesh about the wrist: v' = Q (v - J0) + J0."""
    from scipy.spatial.transform import Rotation as R
    g = torch.Generator().manual_seed(seed)
    orient, pose, betas = torch.randn(1, 3, generator=g) * 0.5, torch.randn(1, 45, generator=g) * 0.3, torch.randn(1, 10, generator=g) * 0.5
    import copy
    layer = copy.deepcopy(oracle_layers[0]).double()
    out = layer(global_orient=orient.double(), hand_pose=pose.double(), betas=betas.double())
    Q = R.from_rotvec(np.asarray(axis_angle, np.float64))
    composed = (Q * R.from_rotvec(orient[0].double().numpy())).as_rotvec()
    out2 = layer(global_orient=torch.tensor(composed).view(1, 3), hand_pose=pose.double(), betas=betas.double())
    j0 = out.joints[0, 0].numpy()
    want = (out.vertices[0].numpy() - j0) @ Q.as_matrix().T + j0
    assert np.abs(out2.vertices[0].numpy() - want).max() <= 1e-7          # (smplx adds 1e-8 to the rotation vector: M2)


# ------------------------------------------------------------------------------------------------ CUDA kernels
def _cuda_sdf(layers, **kw):
    from ihmr_b200 import sdf_loss
    return sdf_loss.SDFLoss(layers[0].faces, layers[1].faces, **kw).cuda()


@pytest.fixture(scope="module")
def cuda_layers(model_root):
    import os
    from ihmr_b200.mano_layer import create
    return (create(os.path.join(model_root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True).cuda(),
            create(os.path.join(model_root, "MANO_LEFT.pkl"), "mano", use_pca=False, is_rhand=False).cuda())


@pytest.mark.gpu
@settings(max_examples=12, **COMMON)
@given(start=st.integers(0, 60000), axis=st.integers(1, 2), mode=st.sampled_from(["typical", "collision"]))
def test_cuda_penetration_axis_relabelling_is_bitwise(oracle_layers, cuda_layers, start, axis, mode):
    """World axes relabelled cyclically and the parity ray relabelled with them: bitwise the same losses and origin-scale
    values, the gradient relabelled (every component goes through the same instructions)."""
    hv = _frames(oracle_layers, mode, start, 24).cuda()
    perm = [(axis + c) % 3 for c in range(3)]            # internal component c of the kernels = world component perm[c]
    inv = [perm.index(c) for c in range(3)]
    hx = hv.clone().requires_grad_(True)
    lx, _, ox = _cuda_sdf(cuda_layers)(hx, True, True)
    lx.sum().backward()
    # a world whose component perm[c] holds what x-ray world component c held, with the ray along world axis `axis`
    hp = hv[..., inv].contiguous().requires_grad_(True)
    lp, _, op = _cuda_sdf(cuda_layers, ray_axis=axis)(hp, True, True)
    lp.sum().backward()
    assert torch.equal(lx, lp) and torch.equal(ox, op)
    assert torch.equal(hx.grad, hp.grad[..., perm])


@pytest.mark.gpu
@settings(max_examples=10, **COMMON)
@given(start=st.integers(0, 60000), log2s=st.integers(-3, 3), t=st.tuples(*[st.floats(-1.0, 1.0, width=32)] * 3),
       mode=st.sampled_from(["typical", "collision"]))
def test_cuda_penetration_similarity_invariance(oracle_layers, cuda_layers, start, log2s, t, mode):
    """loss(s V + t) = loss(V), origin(s V + t) = s origin(V), grad(s V + t) = grad(V) / s.  A power-of-two scale alone is
    bitwise (exponent shift); a translation re-rounds the normalised coordinates, so a tolerance applies."""
    hv = _frames(oracle_layers, mode, start, 32).cuda()
    s = 2.0 ** log2s
    mod = _cuda_sdf(cuda_layers)
    h0 = hv.clone().requires_grad_(True)
    l0, _, o0 = mod(h0, True, True)
    l0.sum().backward()
    h1 = (hv * s).requires_grad_(True)
    l1, _, o1 = mod(h1, True, True)
    l1.sum().backward()
    assert torch.equal(l0, l1) and torch.equal(o0 * s, o1) and torch.equal(h0.grad, h1.grad * s)
    h2 = (hv * s + torch.tensor(t, device="cuda")).requires_grad_(True)
    l2, _, o2 = mod(h2, True, True)
    l2.sum().backward()
    tot = max(float(l0.sum()), 1e-3)
    assert abs(float(l2.sum()) - float(l0.sum())) <= 2e-3 * tot
    assert float((o2 - o0 * s).abs().max()) <= 1e-3 * s * max(float(o0.abs().max()), 1e-3)


@pytest.mark.gpu
@settings(max_examples=10, **COMMON)
@given(seed=st.integers(0, 100000), axis_angle=st.tuples(*[st.floats(-2.0, 2.0, width=32)] * 3))
def test_cuda_mano_global_rotation_equivariance(cuda_layers, seed, axis_angle):
    from scipy.spatial.transform import Rotation as R
    g = torch.Generator().manual_seed(seed)
    n = 64
    orient, pose, betas = torch.randn(n, 3, generator=g) * 0.5, torch.randn(n, 45, generator=g) * 0.3, torch.randn(n, 10, generator=g) * 0.5
    layer = cuda_layers[0]
    out = layer(global_orient=orient.cuda(), hand_pose=pose.cuda(), betas=betas.cuda())
    Q = R.from_rotvec(np.asarray(axis_angle, np.float64))
    composed = (Q * R.from_rotvec(orient.double().numpy())).as_rotvec().astype(np.float32)
    out2 = layer(global_orient=torch.tensor(composed).cuda(), hand_pose=pose.cuda(), betas=betas.cuda())
    j0 = out.joints[:, 0:1].double().cpu().numpy()
    want = (out.vertices.double().cpu().numpy() - j0) @ Q.as_matrix().T + j0
    assert np.abs(out2.vertices.cpu().numpy() - want).max() <= 1e-5          # north_star: 1e-5 m on vertices
