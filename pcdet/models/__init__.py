"""build_network / load_data_to_gpu / model_fn_decorator (reference pcdet/models/__init__.py:16-50)."""
from collections import namedtuple

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


def model_fn_decorator():
    """pcdet/models/__init__.py:37-50: the `model_func(model, batch_dict)` tools/train.py hands to train_model --
    batch to the device, forward in training mode, `ret_dict['loss'].mean()`, global step advanced."""
    ModelReturn = namedtuple("ModelReturn", ["loss", "tb_dict", "disp_dict"])

    def model_func(model, batch_dict):
        load_data_to_gpu(batch_dict)
        ret_dict, tb_dict, disp_dict = model(batch_dict)
        loss = ret_dict["loss"].mean()
        (model if hasattr(model, "update_global_step") else model.module).update_global_step()
        return ModelReturn(loss, tb_dict, disp_dict)

    return model_func
