"""The slice of pcdet/utils/common_utils.py that tools/test.py uses (logger, seeds, process-group init,
result merge).  merge_results_dist keeps the reference's signature (common_utils.py:202-223) but moves the
data with one collective instead of pickle files on a shared tmpdir."""
from __future__ import annotations

import logging
import os
import random

import numpy as np
import torch
import torch.distributed as dist


def create_logger(log_file=None, rank=0, log_level=logging.INFO):
    logger = logging.getLogger(__name__)
    logger.setLevel(log_level if rank == 0 else "ERROR")
    logger.handlers.clear()
    fmt = logging.Formatter("%(asctime)s  %(levelname)5s  %(message)s")
    console = logging.StreamHandler()
    console.setLevel(log_level if rank == 0 else "ERROR")
    console.setFormatter(fmt)
    logger.addHandler(console)
    if log_file is not None:
        fh = logging.FileHandler(filename=log_file)
        fh.setLevel(log_level if rank == 0 else "ERROR")
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    logger.propagate = False
    return logger


def set_random_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def init_dist_pytorch(tcp_port, local_rank, backend="nccl"):
    """one process per GPU; rendezvous from the torchrun / torch.distributed.launch environment."""
    n = max(torch.cuda.device_count(), 1)
    local_rank = int(os.environ.get("LOCAL_RANK", local_rank))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % n)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(tcp_port))
    dist.init_process_group(backend=backend)
    return n, dist.get_rank()


def get_dist_info(return_gpu_per_machine=False):
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    if return_gpu_per_machine:
        return rank, world, torch.cuda.device_count()
    return rank, world


def merge_results_dist(result_part, size, tmpdir=None):
    """every rank's list, re-interleaved in dataset order and cut to `size`; rank 0 gets the list, the
    others None (reference contract).  `tmpdir` is accepted and ignored."""
    rank, world = get_dist_info()
    if world == 1:
        return result_part[:size]
    parts = [None] * world
    dist.all_gather_object(parts, result_part)
    if rank != 0:
        return None
    ordered = []
    for res in zip(*parts):
        ordered.extend(res)
    return ordered[:size]


def rotate_points_along_z(points, angle):
    """(B, N, 3 + C), (B,) -> rotated copy (common_utils.py:35-57); host-side helper."""
    cosa, sina = torch.cos(angle), torch.sin(angle)
    zeros, ones = torch.zeros_like(angle), torch.ones_like(angle)
    rot = torch.stack((cosa, sina, zeros, -sina, cosa, zeros, zeros, zeros, ones), dim=1).view(-1, 3, 3).float()
    return torch.cat((torch.matmul(points[:, :, 0:3], rot), points[:, :, 3:]), dim=-1)
