"""The caller of the refinement path: the per-dataset loop of /root/reference/src/optimize.py:40-102 over the
B200 path — batches from ``OPTDataset.batches`` (contiguous frame blocks per rank), ``OptimizeModel.run_pipelined``
(input and result copies overlap the refinement of the neighbouring batches), records in the reference's
``Evaluator`` format, one gather of the records at the end (torch.distributed object gather instead of the
reference's per-rank pickle files + barrier, optimize.py:78-89) and the result file
``evaluate_results/optimize/<dataset>.pkl`` (optimize.py:91-96)."""
from __future__ import annotations

import collections
import os.path as osp
from typing import Dict, Optional

import torch

from .evaluator import Evaluator
from .opt_dataset import OPTDataset
from .optimize_model import OptimizeModel

METRICS = ("mpjpe_3d", "inter_mpjpe_3d", "collision_ave", "collision_max")


def optimize_dataset(opt, dataset_info, out_dir: str = "evaluate_results/optimize", model: Optional[OptimizeModel] = None,
                     save_verts: bool = True) -> Dict[str, float]:
    """Refines every sample of one dataset and writes ``<out_dir>/<opt.opt_dataset or dataset name>.pkl`` (rank 0).
    Returns the four metrics src/optimize.py:98-102 prints."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(), dist.get_world_size()) if distributed else (0, 1)
    dataset = OPTDataset(opt, dataset_info)
    dataset.load_data(world_size=world)
    if model is None:
        model = OptimizeModel(opt)
    evaluator = Evaluator(opt, dataset, model)
    evaluator.clear()
    pending_idxs = collections.deque()       # run_pipelined reads one batch ahead of the results it hands out

    def feed():
        for batch in dataset.batches(rank, world):
            pending_idxs.append(batch["index"].numpy().copy())
            yield batch

    for res in model.run_pipelined(feed()):
        evaluator.update(pending_idxs.popleft(), res, save_verts=save_verts)
    if distributed:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(evaluator.pred_results, gathered, dst=0)
        if rank == 0:
            evaluator.clear()
            for part in gathered:
                evaluator.gather_pred(part)
    out = {}
    if rank == 0:
        evaluator.remove_redunc()
        name = getattr(opt, "opt_dataset", None) or dataset.name
        evaluator.save(osp.join(out_dir, f"{name}.pkl"))
        out = {m: float(getattr(evaluator, m)) for m in METRICS}
    return out
