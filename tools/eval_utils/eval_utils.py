"""eval_one_epoch (reference tools/eval_utils/eval_utils.py:22-123) on the B200 path.

Same loop contract: load_data_to_gpu -> batch_dict['cur_epoch'] -> model(batch_dict) -> dataset.generate_prediction_dicts
-> merge across ranks -> dataset.evaluation; result.pkl is written by rank 0.  Differences: no DDP wrapper (inference
needs no parameter broadcast: every rank loads the same checkpoint), device-timed throughput, one collective for the merge.
"""
import pickle
import time

import torch

from pcdet.models import load_data_to_gpu
from pcdet.utils import common_utils


def eval_one_epoch(cfg, model, dataloader, epoch_id, logger, dist_test=False, save_to_file=False, result_dir=None):
    result_dir.mkdir(parents=True, exist_ok=True)
    dataset = dataloader.dataset
    class_names = dataset.class_names
    det_annos = []
    logger.info("*************** EPOCH %s EVALUATION *****************" % epoch_id)
    model.eval()
    torch.cuda.synchronize()
    t0 = time.time()
    n_scenes = 0
    for batch_dict in dataloader:
        load_data_to_gpu(batch_dict)
        batch_dict["cur_epoch"] = epoch_id
        with torch.no_grad():
            pred_dicts, ret_dict = model(batch_dict)
        det_annos += dataset.generate_prediction_dicts(batch_dict, pred_dicts, class_names)
        n_scenes += batch_dict["batch_size"]
    torch.cuda.synchronize()
    dt = time.time() - t0
    if dist_test:
        det_annos = common_utils.merge_results_dist(det_annos, len(dataset))
    logger.info("*************** Performance of EPOCH %s *****************" % epoch_id)
    logger.info("Generate label finished (sec_per_example: %.4f second, %.1f scenes/s on this rank)."
                % (dt / max(n_scenes, 1), n_scenes / dt))
    if cfg.LOCAL_RANK != 0:
        return {}
    total = sum(len(a["name"]) for a in det_annos)
    logger.info("Average predicted number of objects(%d samples): %.3f" % (len(det_annos), total / max(1, len(det_annos))))
    with open(result_dir / "result.pkl", "wb") as f:
        pickle.dump(det_annos, f)
    result_str, result_dict = dataset.evaluation(det_annos, class_names,
                                                 eval_metric=cfg.MODEL.POST_PROCESSING.EVAL_METRIC, output_path=result_dir)
    logger.info(result_str)
    logger.info("Result is save to %s" % result_dir)
    logger.info("****************Evaluation done.*****************")
    return dict(result_dict)
