"""Training driver with the reference's CLI (tools/train.py:21-216), running the B200 training path.

    cd tools && python train.py --cfg_file cfgs/scannet_models/CAGroup3D.yaml --fix_random_seed [--epochs 1 --max_iters 20]
    torchrun --nproc-per-node 4 train.py --launcher pytorch --cfg_file cfgs/scannet_models/CAGroup3D.yaml --fix_random_seed

Thin by design: config -> dataloader -> build_network -> optimizer / scheduler -> resume from the newest checkpoint of
the output directory -> train_model (tools/train_utils/train_utils.py) -> repeat_eval_ckpt over the saved checkpoints.
Differences from the reference, all in the plumbing: the DistributedDataParallel wrapper (:144) is
cagroup3d_b200.dist.GradientAllReducer (flat fp32 buckets all-reduced from autograd hooks, overlapping backward);
`--sync_bn` is not offered (the reference trains CAGroup3D with per-GPU BatchNorm); tensorboardX is optional; `--max_iters`
(this repo) bounds the iterations per epoch for smoke runs on the synthetic dataset.
"""
import argparse
import datetime
import glob
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))

import torch  # noqa: E402

from pcdet.config import cfg, cfg_from_list, cfg_from_yaml_file, log_config_to_file  # noqa: E402
from pcdet.datasets import build_dataloader  # noqa: E402
from pcdet.models import build_network, model_fn_decorator  # noqa: E402
from pcdet.utils import common_utils  # noqa: E402
from train_utils.optimization import build_optimizer, build_scheduler  # noqa: E402
from train_utils.train_utils import train_model  # noqa: E402


def parse_config(argv=None):
    p = argparse.ArgumentParser(description="arg parser")
    p.add_argument("--cfg_file", type=str, default=None)
    p.add_argument("--batch_size", type=int, default=None)
    p.add_argument("--epochs", type=int, default=None)
    p.add_argument("--workers", type=int, default=4)
    p.add_argument("--extra_tag", type=str, default="default")
    p.add_argument("--ckpt", type=str, default=None)
    p.add_argument("--pretrained_model", type=str, default=None)
    p.add_argument("--launcher", choices=["none", "pytorch", "slurm"], default="none")
    p.add_argument("--tcp_port", type=int, default=18888)
    p.add_argument("--fix_random_seed", action="store_true", default=False)
    p.add_argument("--ckpt_save_interval", type=int, default=1)
    p.add_argument("--local_rank", type=int, default=0)
    p.add_argument("--max_ckpt_save_num", type=int, default=30)
    p.add_argument("--merge_all_iters_to_one_epoch", action="store_true", default=False)
    p.add_argument("--set", dest="set_cfgs", default=None, nargs=argparse.REMAINDER)
    p.add_argument("--max_waiting_mins", type=int, default=0)
    p.add_argument("--start_epoch", type=int, default=0)
    p.add_argument("--num_epochs_to_eval", type=int, default=0)
    p.add_argument("--save_to_file", action="store_true", default=False)
    p.add_argument("--max_iters", type=int, default=None, help="(this repo) iterations per epoch, for smoke runs")
    p.add_argument("--no_eval", action="store_true", default=False, help="(this repo) skip the evaluation after training")
    args = p.parse_args(argv)
    cfg_from_yaml_file(args.cfg_file, cfg)
    cfg.TAG = Path(args.cfg_file).stem
    cfg.EXP_GROUP_PATH = "/".join(args.cfg_file.split("/")[1:-1])
    if args.set_cfgs is not None:
        cfg_from_list(args.set_cfgs, cfg)
    return args, cfg


def main(argv=None):
    args, cfg_ = parse_config(argv)
    if args.launcher == "none":
        dist_train, total_gpus = False, 1
    else:
        total_gpus, cfg_.LOCAL_RANK = common_utils.init_dist_pytorch(args.tcp_port, args.local_rank, backend="nccl")
        dist_train = True
    if args.batch_size is None:
        args.batch_size = cfg_.OPTIMIZATION.BATCH_SIZE_PER_GPU
    else:
        assert args.batch_size % total_gpus == 0, "Batch size should match the number of gpus"
        args.batch_size = args.batch_size // total_gpus
    args.epochs = cfg_.OPTIMIZATION.NUM_EPOCHS if args.epochs is None else args.epochs
    assert args.fix_random_seed, "we must fix random seed."          # reference :72
    common_utils.set_random_seed(0)

    output_dir = cfg_.ROOT_DIR / "output" / cfg_.EXP_GROUP_PATH / cfg_.TAG / args.extra_tag
    ckpt_dir = output_dir / "ckpt"
    ckpt_dir.mkdir(parents=True, exist_ok=True)
    log_file = output_dir / ("log_train_%s.txt" % datetime.datetime.now().strftime("%Y%m%d-%H%M%S"))
    logger = common_utils.create_logger(log_file, rank=cfg_.LOCAL_RANK)
    logger.info("**********************Start logging**********************")
    logger.info("CUDA_VISIBLE_DEVICES=%s" % os.environ.get("CUDA_VISIBLE_DEVICES", "ALL"))
    if dist_train:
        logger.info("total_batch_size: %d" % (total_gpus * args.batch_size))
    for key, val in vars(args).items():
        logger.info("{:16} {}".format(key, val))
    log_config_to_file(cfg_, logger=logger)
    tb_log = None
    if cfg_.LOCAL_RANK == 0:
        try:
            from tensorboardX import SummaryWriter
            tb_log = SummaryWriter(log_dir=str(output_dir / "tensorboard"))
        except ImportError:
            pass

    train_set, train_loader, train_sampler = build_dataloader(
        dataset_cfg=cfg_.DATA_CONFIG, class_names=cfg_.CLASS_NAMES, batch_size=args.batch_size, dist=dist_train,
        workers=args.workers, logger=logger, training=True, merge_all_iters_to_one_epoch=args.merge_all_iters_to_one_epoch,
        total_epochs=args.epochs)
    model = build_network(model_cfg=cfg_.MODEL, num_class=len(cfg_.CLASS_NAMES), dataset=train_set)
    model.cuda()
    optimizer = build_optimizer(model, cfg_.OPTIMIZATION)

    start_epoch = it = 0
    last_epoch = -1
    if args.pretrained_model is not None:
        model.load_params_from_file(filename=args.pretrained_model, to_cpu=dist_train, logger=logger)
    if args.ckpt is not None:
        it, start_epoch = model.load_params_with_optimizer(args.ckpt, to_cpu=dist_train, optimizer=optimizer, logger=logger)
        last_epoch = start_epoch + 1
    else:
        ckpt_list = glob.glob(str(ckpt_dir / "*checkpoint_epoch_*.pth"))
        if ckpt_list:
            ckpt_list.sort(key=os.path.getmtime)
            it, start_epoch = model.load_params_with_optimizer(ckpt_list[-1], to_cpu=dist_train, optimizer=optimizer, logger=logger)
            last_epoch = start_epoch + 1
    model.train()
    from cagroup3d_b200 import dist as D
    reducer = D.GradientAllReducer(model.parameters()) if dist_train else None     # tools/train.py:144 (DDP)
    if last_epoch >= 0:
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", cfg_.OPTIMIZATION.LR)
    lr_scheduler, lr_warmup_scheduler = build_scheduler(optimizer, total_iters_each_epoch=len(train_loader), total_epochs=args.epochs,
                                                        last_epoch=last_epoch, optim_cfg=cfg_.OPTIMIZATION)
    logger.info("**********************Start training %s/%s(%s)**********************" % (cfg_.EXP_GROUP_PATH, cfg_.TAG, args.extra_tag))
    it = train_model(model, optimizer, train_loader, model_func=model_fn_decorator(), lr_scheduler=lr_scheduler,
                     optim_cfg=cfg_.OPTIMIZATION, start_epoch=start_epoch, total_epochs=args.epochs, start_iter=it,
                     rank=cfg_.LOCAL_RANK, tb_log=tb_log, ckpt_save_dir=ckpt_dir, train_sampler=train_sampler,
                     lr_warmup_scheduler=lr_warmup_scheduler, ckpt_save_interval=args.ckpt_save_interval,
                     max_ckpt_save_num=args.max_ckpt_save_num, merge_all_iters_to_one_epoch=args.merge_all_iters_to_one_epoch,
                     logger=logger, reducer=reducer, max_iters=args.max_iters)
    logger.info("**********************End training %s/%s(%s)**********************\n\n\n" % (cfg_.EXP_GROUP_PATH, cfg_.TAG, args.extra_tag))
    if args.no_eval:
        return it
    logger.info("**********************Start evaluation %s/%s(%s)**********************" % (cfg_.EXP_GROUP_PATH, cfg_.TAG, args.extra_tag))
    from test import repeat_eval_ckpt
    test_set, test_loader, sampler = build_dataloader(dataset_cfg=cfg_.DATA_CONFIG, class_names=cfg_.CLASS_NAMES, batch_size=1,
                                                      dist=dist_train, workers=args.workers, logger=logger, training=False)
    eval_output_dir = output_dir / "eval" / "eval_with_train"
    eval_output_dir.mkdir(parents=True, exist_ok=True)
    args.start_epoch = max(args.epochs - args.num_epochs_to_eval, 0)
    repeat_eval_ckpt(model, test_loader, args, eval_output_dir, logger, ckpt_dir, dist_test=dist_train)
    logger.info("**********************End evaluation %s/%s(%s)**********************" % (cfg_.EXP_GROUP_PATH, cfg_.TAG, args.extra_tag))
    return it


if __name__ == "__main__":
    main()
