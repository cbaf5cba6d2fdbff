"""build_optimizer / build_scheduler of the reference's training tools (tools/train_utils/optimization/__init__.py:11-64)
for the optimizers the CAGroup3D configs name: adam, adamW (scannet / sunrgbd CAGroup3D.yaml: adamW, LR 0.001, step decay
at DECAY_STEP_LIST epochs), sgd.  `adam_onecycle` (fastai wrapper, outdoor configs) is outside the hot path."""
import torch.optim as optim
import torch.optim.lr_scheduler as lr_sched


def build_optimizer(model, optim_cfg):
    name = optim_cfg.OPTIMIZER
    if name == "adam":
        return optim.Adam(model.parameters(), lr=optim_cfg.LR, weight_decay=optim_cfg.WEIGHT_DECAY)
    if name == "adamW":
        return optim.AdamW(model.parameters(), lr=optim_cfg.LR, weight_decay=optim_cfg.WEIGHT_DECAY)
    if name == "sgd":
        return optim.SGD(model.parameters(), lr=optim_cfg.LR, weight_decay=optim_cfg.WEIGHT_DECAY, momentum=optim_cfg.MOMENTUM)
    raise NotImplementedError(f"OPTIMIZER {name}: only adam / adamW / sgd are on the CAGroup3D path")


def build_scheduler(optimizer, total_iters_each_epoch, total_epochs, last_epoch, optim_cfg):
    """-> (lr_scheduler, lr_warmup_scheduler=None).  The scheduler is stepped with the ITERATION count
    (train_utils.py:40 `lr_scheduler.step(accumulated_iter)`), so the decay epochs are converted to iterations."""
    steps = [e * total_iters_each_epoch for e in optim_cfg.DECAY_STEP_LIST]
    floor = optim_cfg.get("LR_CLIP", 0.0) / optim_cfg.LR

    def factor(it):
        f = 1.0
        for s in steps:
            if it >= s:
                f *= optim_cfg.LR_DECAY
        return max(f, floor)
    return lr_sched.LambdaLR(optimizer, factor, last_epoch=last_epoch), None
