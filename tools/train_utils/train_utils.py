"""Training loop of the reference's tools (tools/train_utils/train_utils.py:12-196) over the B200 training path.

Same call structure -- train_model -> train_one_epoch -> model_func(model, batch) -> backward -> clip_grad_norm_ ->
optimizer.step, `batch['cur_epoch']` set per iteration, `disp_dict.pop('cur_semantic_value')`, LambdaLR stepped with the
iteration count, one checkpoint per epoch with the reference's dict layout -- minus tqdm / tensorboardX (not installed
offline; the loss terms go to the logger every LOG_INTERVAL iterations like the reference's LogBuffer line).  The
DistributedDataParallel wrapper of tools/train.py:144 is `reducer` (cagroup3d_b200.dist.GradientAllReducer): gradients
live in flat buckets and a bucket's all-reduce is launched from the autograd hook of its last parameter, so NCCL overlaps
the rest of backward; `reducer.reduce()` waits for them before the clip.
"""
import glob
import os
import time

import torch
from torch.nn.utils import clip_grad_norm_

LOG_INTERVAL = 50


def train_one_epoch(model, optimizer, train_loader, model_func, lr_scheduler, accumulated_iter, optim_cfg, rank,
                    total_it_each_epoch, dataloader_iter, cur_epoch=None, logger=None, reducer=None, tb_log=None,
                    max_iters=None):
    if total_it_each_epoch == len(train_loader):
        dataloader_iter = iter(train_loader)
    sums, n_sum = {}, 0
    t_epoch = time.time()
    for cur_it in range(total_it_each_epoch):
        if max_iters is not None and cur_it >= max_iters:
            break
        try:
            batch = next(dataloader_iter)
        except StopIteration:
            dataloader_iter = iter(train_loader)
            batch = next(dataloader_iter)
        lr_scheduler.step(accumulated_iter)
        cur_lr = optimizer.param_groups[0]["lr"]
        model.train()
        if reducer is not None:
            reducer.zero_grad()                       # keeps p.grad pointing into the buckets
        else:
            optimizer.zero_grad()
        batch["cur_epoch"] = cur_epoch
        loss, tb_dict, disp_dict = model_func(model, batch)
        cur_semantic_value = disp_dict.pop("cur_semantic_value")
        loss.backward()
        if reducer is not None:
            reducer.reduce()
        clip_grad_norm_(model.parameters(), optim_cfg.GRAD_NORM_CLIP)
        optimizer.step()
        accumulated_iter += 1
        for k, v in disp_dict.items():
            sums[k] = sums.get(k, 0.0) + float(v)
        n_sum += 1
        if tb_log is not None and rank == 0:
            tb_log.add_scalar("train/loss", float(loss), accumulated_iter)
            tb_log.add_scalar("meta_data/learning_rate", cur_lr, accumulated_iter)
            for key, val in tb_dict.items():
                tb_log.add_scalar("train/" + key, val, accumulated_iter)
        if logger is not None and rank == 0 and ((cur_it + 1) % LOG_INTERVAL == 0 or cur_it + 1 == total_it_each_epoch
                                                 or (max_iters is not None and cur_it + 1 == max_iters)):
            info = "Epoch [{0:2d}][{1:4d}]/[{2:4d}] : lr: {3:10.3e}, sem_thr: {4:.2f}, ".format(
                cur_epoch + 1, cur_it + 1, total_it_each_epoch, cur_lr, cur_semantic_value)
            info += ", ".join(f"{k}: {v / n_sum:.4f}" for k, v in sums.items())
            info += f", {(time.time() - t_epoch) / (cur_it + 1):.3f} s/it"
            logger.info(info)
            sums, n_sum = {}, 0
    return accumulated_iter


def train_model(model, optimizer, train_loader, model_func, lr_scheduler, optim_cfg, start_epoch, total_epochs, start_iter,
                rank, tb_log, ckpt_save_dir, train_sampler=None, lr_warmup_scheduler=None, ckpt_save_interval=1,
                max_ckpt_save_num=50, merge_all_iters_to_one_epoch=False, logger=None, reducer=None, max_iters=None):
    accumulated_iter = start_iter
    total_it_each_epoch = len(train_loader)
    dataloader_iter = iter(train_loader)
    for cur_epoch in range(start_epoch, total_epochs):
        if train_sampler is not None:
            train_sampler.set_epoch(cur_epoch)
        cur_scheduler = lr_warmup_scheduler if (lr_warmup_scheduler is not None and cur_epoch < optim_cfg.WARMUP_EPOCH) \
            else lr_scheduler
        accumulated_iter = train_one_epoch(model, optimizer, train_loader, model_func, lr_scheduler=cur_scheduler,
                                           accumulated_iter=accumulated_iter, optim_cfg=optim_cfg, rank=rank,
                                           total_it_each_epoch=total_it_each_epoch, dataloader_iter=dataloader_iter,
                                           cur_epoch=cur_epoch, logger=logger, reducer=reducer, tb_log=tb_log,
                                           max_iters=max_iters)
        trained_epoch = cur_epoch + 1
        if trained_epoch % ckpt_save_interval == 0 and rank == 0:
            ckpt_list = glob.glob(str(ckpt_save_dir / "checkpoint_epoch_*.pth"))
            ckpt_list.sort(key=os.path.getmtime)
            for old in ckpt_list[:max(0, len(ckpt_list) - max_ckpt_save_num + 1)]:
                os.remove(old)
            save_checkpoint(checkpoint_state(model, optimizer, trained_epoch, accumulated_iter),
                            filename=ckpt_save_dir / ("checkpoint_epoch_%d" % trained_epoch))
    return accumulated_iter


def model_state_to_cpu(model_state):
    return type(model_state)((k, v.cpu()) for k, v in model_state.items())


def checkpoint_state(model=None, optimizer=None, epoch=None, it=None):
    """{'epoch', 'it', 'model_state', 'optimizer_state', 'version'} (train_utils.py:169-185): what
    load_params_from_file / load_params_with_optimizer read and what the reference's tools can load."""
    model_state = None
    if model is not None:
        model_state = model_state_to_cpu((model.module if hasattr(model, "module") else model).state_dict())
    try:
        import pcdet
        version = "pcdet+" + pcdet.__version__
    except Exception:
        version = "none"
    return {"epoch": epoch, "it": it, "model_state": model_state,
            "optimizer_state": optimizer.state_dict() if optimizer is not None else None, "version": version}


def save_checkpoint(state, filename="checkpoint"):
    torch.save(state, "{}.pth".format(filename))
