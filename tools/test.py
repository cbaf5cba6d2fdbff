"""Inference driver with the reference's CLI (tools/test.py:21-54,138-203), running the B200 path.

    cd tools && python test.py --cfg_file cfgs/scannet_models/CAGroup3D.yaml --ckpt ../output/ckpt/checkpoint_epoch_10.pth
    torchrun --nproc-per-node 8 test.py --launcher pytorch --cfg_file ... --ckpt ...

`--ckpt` may be omitted (or name a missing file with --allow_random_init) to run seed-0 weights on the synthetic
dataset: there is no checkpoint or dataset offline.  The epoch id is parsed from the LAST integer of the checkpoint path
(reference :161-162) and drives the semantic threshold (cagroup3d.py:29-31).
"""
import argparse
import datetime
import os
import re
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from eval_utils import eval_utils  # noqa: E402
from pcdet.config import cfg, cfg_from_list, cfg_from_yaml_file, log_config_to_file  # noqa: E402
from pcdet.datasets import build_dataloader  # noqa: E402
from pcdet.models import build_network  # noqa: E402
from pcdet.utils import common_utils  # noqa: E402


def parse_config(argv=None):
    p = argparse.ArgumentParser(description="arg parser")
    p.add_argument("--cfg_file", type=str, default=None)
    p.add_argument("--batch_size", type=int, default=None)
    p.add_argument("--workers", type=int, default=4)
    p.add_argument("--extra_tag", type=str, default="default")
    p.add_argument("--ckpt", type=str, default=None)
    p.add_argument("--launcher", choices=["none", "pytorch", "slurm"], default="none")
    p.add_argument("--tcp_port", type=int, default=18888)
    p.add_argument("--local_rank", type=int, default=0)
    p.add_argument("--set", dest="set_cfgs", default=None, nargs=argparse.REMAINDER)
    p.add_argument("--max_waiting_mins", type=int, default=30)
    p.add_argument("--start_epoch", type=int, default=0)
    p.add_argument("--eval_tag", type=str, default="default")
    p.add_argument("--eval_all", action="store_true", default=False)
    p.add_argument("--ckpt_dir", type=str, default=None)
    p.add_argument("--save_to_file", action="store_true", default=False)
    p.add_argument("--allow_random_init", action="store_true", default=False,
                   help="(this repo) run seed-0 weights when no checkpoint file exists")
    args = p.parse_args(argv)
    cfg_from_yaml_file(args.cfg_file, cfg)
    cfg.TAG = Path(args.cfg_file).stem
    cfg.EXP_GROUP_PATH = "/".join(args.cfg_file.split("/")[1:-1])
    common_utils.set_random_seed(0)
    if args.set_cfgs is not None:
        cfg_from_list(args.set_cfgs, cfg)
    return args, cfg


def main(argv=None):
    args, cfg_ = parse_config(argv)
    if args.eval_all:
        raise NotImplementedError("--eval_all (checkpoint-directory polling, reference :101-135) is outside the hot path")
    if args.launcher == "none":
        dist_test, total_gpus = False, 1
    else:
        total_gpus, cfg_.LOCAL_RANK = common_utils.init_dist_pytorch(args.tcp_port, args.local_rank, backend="nccl")
        dist_test = True
    if args.batch_size is None:
        args.batch_size = cfg_.OPTIMIZATION.BATCH_SIZE_PER_GPU
    else:
        assert args.batch_size % total_gpus == 0, "Batch size should match the number of gpus"
        args.batch_size = args.batch_size // total_gpus

    output_dir = cfg_.ROOT_DIR / "output" / cfg_.EXP_GROUP_PATH / cfg_.TAG / args.extra_tag
    output_dir.mkdir(parents=True, exist_ok=True)
    nums = re.findall(r"\d+", args.ckpt) if args.ckpt is not None else []
    epoch_id = nums[-1] if nums else "no_number"
    eval_output_dir = output_dir / "eval" / ("epoch_%s" % epoch_id) / cfg_.DATA_CONFIG.DATA_SPLIT["test"] / args.eval_tag
    eval_output_dir.mkdir(parents=True, exist_ok=True)
    log_file = eval_output_dir / ("log_eval_%s.txt" % datetime.datetime.now().strftime("%Y%m%d-%H%M%S"))
    logger = common_utils.create_logger(log_file, rank=cfg_.LOCAL_RANK)
    logger.info("**********************Start logging**********************")
    logger.info("CUDA_VISIBLE_DEVICES=%s" % os.environ.get("CUDA_VISIBLE_DEVICES", "ALL"))
    if dist_test:
        logger.info("total_batch_size: %d" % (total_gpus * args.batch_size))
    for key, val in vars(args).items():
        logger.info("{:16} {}".format(key, val))
    log_config_to_file(cfg_, logger=logger)

    test_set, test_loader, sampler = build_dataloader(dataset_cfg=cfg_.DATA_CONFIG, class_names=cfg_.CLASS_NAMES,
                                                      batch_size=args.batch_size, dist=dist_test, workers=args.workers,
                                                      logger=logger, training=False)
    model = build_network(model_cfg=cfg_.MODEL, num_class=len(cfg_.CLASS_NAMES), dataset=test_set)
    with torch.no_grad():
        if args.ckpt is not None and os.path.isfile(args.ckpt):
            model.load_params_from_file(filename=args.ckpt, logger=logger, to_cpu=dist_test)
        elif args.ckpt is None or args.allow_random_init:
            logger.info("no checkpoint file: running the seed-0 initialisation (throughput / plumbing check only)")
            epoch_id = epoch_id if epoch_id != "no_number" else "10"
        else:
            raise FileNotFoundError(args.ckpt)
        model.cuda()
        return eval_utils.eval_one_epoch(cfg_, model, test_loader, int(epoch_id), logger, dist_test=dist_test,
                                         result_dir=eval_output_dir, save_to_file=args.save_to_file)


if __name__ == "__main__":
    main()
