"""build_network / load_data_to_gpu (reference pcdet/models/__init__.py:16-34)."""
import numpy as np
import torch

from .detectors import build_detector


def build_network(model_cfg, num_class, dataset):
    return build_detector(model_cfg=model_cfg, num_class=num_class, dataset=dataset)


def load_data_to_gpu(batch_dict):
    """every ndarray except bookkeeping keys -> float32 CUDA tensor, in place.  Pinned staging + one
    asynchronous copy per array instead of the reference's pageable synchronous `.cuda()`."""
    for key, val in batch_dict.items():
        if not isinstance(val, np.ndarray) or key in ("frame_id", "metadata", "calib"):
            continue
        t = torch.from_numpy(val)
        t = t.int() if key == "image_shape" else t.float()
        batch_dict[key] = t.pin_memory().cuda(non_blocking=True)
