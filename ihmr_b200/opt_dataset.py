"""Input side of the refinement path: the annotation + prior-prediction records the
reference's ``OPTDataset`` serves (/root/reference/src/data/opt_dataset.py:17-198, data_utils.py:42-70), turned
into the 17 tensors ``OptimizeModel.set_input`` takes (keys at opt_dataset.py:176-196).

Same record arithmetic as the reference, sample by sample (``OPTDataset.__getitem__``), plus what a B200 needs
and a per-sample DataLoader cannot give: ``batches()`` assembles whole pinned batches with a handful of
vectorised numpy operations (a 65536-frame batch is 220 MB) and hands every rank a contiguous block of frames
(SURVEY.md §8(e)).  The image itself is never decoded — the refinement path only needs its height and width for
the 2-D joint normalisation (data_preprocess.py:45-60,162-169), read from the file header.
"""
from __future__ import annotations

import os.path as osp
import pickle
from typing import Dict, Iterator, List, Sequence, Tuple

import numpy as np
import torch

BATCH_KEYS = ("joints_2d", "joints_3d", "mano_pose", "mano_betas", "mano_params_weight", "hand_trans", "hand_type_array",
              "hand_type_valid", "scale_ratio", "index", "init_cam", "init_shape_params", "init_pose_params",
              "init_hand_trans", "init_joints_2d", "init_joints_3d", "init_hand_trans_j")    # opt_dataset.py:176-196


def load_pkl(path):
    with open(path, "rb") as fh:
        return pickle.load(fh, encoding="latin1")


def save_pkl(path, obj):
    with open(path, "wb") as fh:
        pickle.dump(obj, fh, protocol=2)


def load_anno_pred_data(data_root: str, anno_path: str, pred_res_path: str) -> List[dict]:
    """data_utils.py:42-70: annotation records keyed by ``img_path`` merged with the prior network's predictions."""
    anno = {d["img_path"]: d for d in load_pkl(osp.join(data_root, anno_path))}
    pred = load_pkl(osp.join(data_root, pred_res_path))
    out = []
    for key, rec in anno.items():
        p = pred[key]
        for k in ("pred_cam_params", "pred_shape_params", "pred_pose_params", "pred_hand_trans"):
            rec[k] = p[k]
        for k in ("joints_2d", "joints_3d"):
            rec[f"pred_{k}"] = p[k]
        rec["img_feat"] = p["img_feat"]
        out.append(rec)
    assert len(out) > 0, "Data List must have data."
    return out


def image_size(path: str) -> Tuple[int, int]:
    """(height, width) from the file header; the pixels are not decoded."""
    try:
        from PIL import Image
        with Image.open(path) as im:
            w, h = im.size
        return h, w
    except ImportError:
        import cv2
        img = cv2.imread(path)
        return img.shape[0], img.shape[1]


def hand_type_str2array(hand_type: str) -> np.ndarray:
    """data_preprocess.py:35-42"""
    if hand_type == "right":
        return np.array([1, 0], dtype=np.float32)
    if hand_type == "left":
        return np.array([0, 1], dtype=np.float32)
    assert hand_type == "interacting", f"{hand_type} not supported."
    return np.array([1, 1], dtype=np.float32)


class OPTDataset:
    """``OPTDataset(opt, (name, anno_path, pred_res_path, image_root))`` — the constructor, ``load_data`` and
    ``__getitem__`` of the reference class, minus the decoded image (nothing on the refinement path reads it)."""

    def __init__(self, opt, dataset_info: Sequence[str]):
        name, anno_path, pred_res_path, image_root = dataset_info
        self.name, self.anno_path, self.pred_res_path = name, anno_path, pred_res_path
        self.image_root = osp.join(opt.data_root, image_root)
        self.opt, self.data_root, self.param_root = opt, opt.data_root, opt.param_root
        self.data_list: List[dict] = []
        self.num_add = 0
        self._sizes: Dict[str, Tuple[int, int]] = {}

    def load_data(self, world_size: int = 1):
        """opt_dataset.py:38-51: the number of samples is padded (with copies of sample 0) to a multiple of
        batchSize x world size, because the model needs full batches (optimize_model.py:185); the duplicates are
        dropped again by ``Evaluator.remove_redunc``."""
        data_list = load_anno_pred_data(self.data_root, self.anno_path, self.pred_res_path)
        bs = self.opt.batchSize * max(1, world_size)
        num_add = bs - len(data_list) % bs
        self.num_add = 0 if num_add == bs else num_add
        self.data_list = data_list + data_list[0:1] * self.num_add

    def __len__(self):
        return len(self.data_list)

    # ------------------------------------------------------------------ one sample, as the reference builds it
    def _ratio(self, img_path: str) -> float:
        """data_preprocess.py:45-60: the image is padded to a square and resized to inputSize; 2-D joints scale with it."""
        if img_path not in self._sizes:
            self._sizes[img_path] = image_size(osp.join(self.image_root, img_path))
        h, w = self._sizes[img_path]
        return self.opt.inputSize / h if h > w else self.opt.inputSize / w

    def _sample(self, index: int) -> Dict[str, np.ndarray]:
        anno = dict(self.data_list[index])
        anno.update(load_pkl(osp.join(self.param_root, anno["param_path"])))      # merge two dicts (:71-73)
        nj = self.opt.num_joints
        hand_type_array = hand_type_str2array(anno["hand_type"])
        hand_type_valid = np.array([anno["hand_type_valid"]], dtype=np.float32)
        joints_2d = np.array(anno["joints_2d"], dtype=np.float64) if "joints_2d" in anno else np.zeros((nj, 3))
        if joints_2d.shape[1] == 2:
            joints_2d = np.concatenate((joints_2d, np.ones((joints_2d.shape[0], 1), dtype=np.float32)), axis=1)
        joints_3d = np.array(anno["joints_3d"], dtype=np.float64) if "joints_3d" in anno else np.zeros((nj, 3))
        if joints_3d.shape[1] == 3:
            joints_3d = np.concatenate((joints_3d, np.ones((joints_3d.shape[0], 1), dtype=np.float32)), axis=1)
        scale_ratio = anno["scale"] if "scale" in anno else 1.0
        mano_pose, mano_betas = np.zeros((96,), np.float32), np.zeros((20,), np.float32)
        mano_params_weight = np.zeros((2,), np.float32)
        for i, hand in enumerate(("right", "left")):
            value = anno[f"{hand}_hand_param"]
            if value is not None:
                mano_pose[48 * i:48 * i + 48] = value["pose"]
                mano_betas[10 * i:10 * i + 10] = value["shape"]
                mano_params_weight[i] = 1
        if joints_3d[0, -1] > 0.0 and joints_3d[21, -1] > 0.0:
            hand_trans, tw = -joints_3d[0, :3] + joints_3d[21, :3], np.ones((1,), np.float32)
        else:
            hand_trans, tw = np.zeros((3,), np.float32), np.zeros((1,), np.float32)
        hand_trans = np.concatenate((hand_trans, tw)).reshape(1, 4)
        init_joints_2d, init_joints_3d = np.asarray(anno["pred_joints_2d"]), np.asarray(anno["pred_joints_3d"])
        score = np.ones((init_joints_2d.shape[0], 1))
        init_joints_2d = np.concatenate((init_joints_2d, score), axis=1)
        init_joints_3d = np.concatenate((init_joints_3d, score), axis=1)
        one = np.ones((1,), np.float32)
        init_hand_trans_j = np.concatenate((init_joints_3d[21, :3] - init_joints_3d[0, :3], one)).reshape(1, 4)
        init_hand_trans = np.concatenate((np.asarray(anno["pred_hand_trans"]), one)).reshape(1, 4)
        ratio, size = self._ratio(anno["img_path"]), float(self.opt.inputSize)

        def norm2d(j):                       # padding_and_resize + normalize_joints_2d, on the joints only
            j = np.array(j, dtype=np.float64)
            j[:, :2] *= ratio
            out = np.copy(j)
            out[:, 0] = (j[:, 0] / size) * 2.0 - 1.0
            out[:, 1] = (j[:, 1] / size) * 2.0 - 1.0
            return out
        return dict(joints_2d=norm2d(joints_2d), joints_3d=joints_3d, mano_pose=mano_pose, mano_betas=mano_betas,
                    mano_params_weight=mano_params_weight, hand_trans=hand_trans, hand_type_array=hand_type_array,
                    hand_type_valid=hand_type_valid, scale_ratio=np.float64(scale_ratio), index=np.int64(index),
                    init_cam=np.asarray(anno["pred_cam_params"]), init_shape_params=np.asarray(anno["pred_shape_params"]),
                    init_pose_params=np.asarray(anno["pred_pose_params"]), init_hand_trans=init_hand_trans,
                    init_joints_2d=norm2d(init_joints_2d), init_joints_3d=init_joints_3d, init_hand_trans_j=init_hand_trans_j)

    def __getitem__(self, index: int) -> Dict[str, torch.Tensor]:
        s = self._sample(index)
        out = {k: torch.from_numpy(np.asarray(v)).float() for k, v in s.items() if k not in ("scale_ratio", "index")}
        out["scale_ratio"] = torch.tensor(float(s["scale_ratio"]))
        out["index"] = torch.tensor(int(s["index"]))
        return out

    # ------------------------------------------------------------------ whole batches
    def batches(self, rank: int = 0, world_size: int = 1, pin: bool = True) -> Iterator[Dict[str, torch.Tensor]]:
        """Full batches of ``opt.batchSize`` samples for this rank, in order: rank r owns the contiguous block
        [r n / W, (r + 1) n / W) of the padded list.  Tensors are stacked per key (float32, ``index`` int64, as the
        default collate of the reference's DataLoader makes them) and pinned."""
        n, bs = len(self.data_list), self.opt.batchSize
        assert n % (bs * max(1, world_size)) == 0, "call load_data(world_size) first"
        per = n // max(1, world_size)
        for start in range(rank * per, (rank + 1) * per, bs):
            samples = [self._sample(i) for i in range(start, start + bs)]
            batch = {}
            for k in BATCH_KEYS:
                arr = np.stack([np.asarray(s[k]) for s in samples])
                if k == "index":
                    t = torch.from_numpy(arr.astype(np.int64))
                else:
                    t = torch.from_numpy(arr.astype(np.float32))
                batch[k] = t.pin_memory() if pin and torch.cuda.is_available() else t
            yield batch
