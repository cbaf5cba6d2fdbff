"""Inference driver with the reference's CLI (tools/test.py:21-54,138-203), running the B200 path.

    cd tools && python test.py --cfg_file cfgs/scannet_models/CAGroup3D.yaml --ckpt ../output/ckpt/checkpoint_epoch_10.pth
    torchrun --nproc-per-node 8 test.py --launcher pytorch --cfg_file ... --ckpt ...

Like the reference it needs a checkpoint; `--allow_random_init` (this repo) runs the seed-0 weights on the synthetic
dataset when `--ckpt` is omitted or names a missing file: there is no checkpoint or dataset offline.  `--eval_all`
evaluates every not-yet-evaluated checkpoint of `--ckpt_dir` (reference :89-135).  The epoch id is parsed from the LAST integer of the checkpoint path
(reference :161-162) and drives the semantic threshold (cagroup3d.py:29-31).
"""
import argparse
import datetime
import glob
import os
import re
import sys
import time
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


def get_no_evaluated_ckpt(ckpt_dir, ckpt_record_file, args):
    """tools/test.py:72-86: the oldest checkpoint_epoch_*.pth of ckpt_dir whose epoch id is not in the record file (and is
    >= args.start_epoch, and is not an `optim` side file) -> (epoch_id, path) or (-1, None)."""
    ckpt_list = glob.glob(os.path.join(str(ckpt_dir), "*checkpoint_epoch_*.pth"))
    ckpt_list.sort(key=os.path.getmtime)
    evaluated = [float(x.strip()) for x in open(ckpt_record_file, "r").readlines() if x.strip()]
    for cur_ckpt in ckpt_list:
        num_list = re.findall(r"checkpoint_epoch_(.*).pth", cur_ckpt)
        if len(num_list) == 0:
            continue
        epoch_id = num_list[-1]
        if "optim" in epoch_id:
            continue
        if float(epoch_id) not in evaluated and int(float(epoch_id)) >= args.start_epoch:
            return epoch_id, cur_ckpt
    return -1, None


def repeat_eval_ckpt(model, test_loader, args, eval_output_dir, logger, ckpt_dir, dist_test=False, poll_seconds=30):
    """tools/test.py:89-135 (--eval_all, and the evaluation tools/train.py runs after training): evaluate every
    checkpoint of ckpt_dir that has not been evaluated yet, oldest first, recording the epoch ids in
    eval_list_<split>.txt; polls the directory until nothing new appeared for args.max_waiting_mins."""
    ckpt_record_file = Path(eval_output_dir) / ("eval_list_%s.txt" % cfg.DATA_CONFIG.DATA_SPLIT["test"])
    with open(ckpt_record_file, "a"):
        pass
    total_time, first_eval, results = 0, True, {}
    while True:
        cur_epoch_id, cur_ckpt = get_no_evaluated_ckpt(ckpt_dir, ckpt_record_file, args)
        if cur_epoch_id == -1 or int(float(cur_epoch_id)) < args.start_epoch:
            if total_time >= args.max_waiting_mins * 60 and (first_eval is False or args.max_waiting_mins == 0):
                break
            if cfg.LOCAL_RANK == 0:
                print("Wait %s seconds for next check (progress: %.1f / %d minutes): %s \r"
                      % (poll_seconds, total_time * 1.0 / 60, args.max_waiting_mins, ckpt_dir), end="", flush=True)
            time.sleep(poll_seconds)
            total_time += poll_seconds
            continue
        total_time, first_eval = 0, False
        model.load_params_from_file(filename=cur_ckpt, logger=logger, to_cpu=dist_test)
        model.cuda()
        cur_result_dir = Path(eval_output_dir) / ("epoch_%s" % cur_epoch_id) / cfg.DATA_CONFIG.DATA_SPLIT["test"]
        with torch.no_grad():
            results[cur_epoch_id] = eval_utils.eval_one_epoch(cfg, model, test_loader, int(float(cur_epoch_id)), logger,
                                                              dist_test=dist_test, result_dir=cur_result_dir,
                                                              save_to_file=args.save_to_file)
        with open(ckpt_record_file, "a") as f:
            print("%s" % cur_epoch_id, file=f)
        logger.info("Epoch %s has been evaluated" % cur_epoch_id)
    return results


def main(argv=None):
    args, cfg_ = parse_config(argv)
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
    if args.eval_all:                                     # reference :170-172,196-198
        ckpt_dir = Path(args.ckpt_dir) if args.ckpt_dir is not None else output_dir / "ckpt"
        eval_all_dir = output_dir / "eval" / "eval_all_default" / args.eval_tag
        eval_all_dir.mkdir(parents=True, exist_ok=True)
        return repeat_eval_ckpt(model, test_loader, args, eval_all_dir, logger, ckpt_dir, dist_test=dist_test)
    with torch.no_grad():
        if args.ckpt is not None and os.path.isfile(args.ckpt):
            model.load_params_from_file(filename=args.ckpt, logger=logger, to_cpu=dist_test)
        elif args.allow_random_init:
            logger.info("no checkpoint file: running the seed-0 initialisation (throughput / plumbing check only)")
            epoch_id = epoch_id if epoch_id != "no_number" else "10"
        else:
            raise FileNotFoundError("%s: the reference always evaluates a checkpoint; pass --allow_random_init to run the "
                                    "seed-0 weights (throughput / plumbing check only)" % args.ckpt)
        model.cuda()
        return eval_utils.eval_one_epoch(cfg_, model, test_loader, int(epoch_id), logger, dist_test=dist_test,
                                         result_dir=eval_output_dir, save_to_file=args.save_to_file)


if __name__ == "__main__":
    main()
