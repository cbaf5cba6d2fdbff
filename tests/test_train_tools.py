"""tools/train.py plumbing: checkpoint dict layout + resume, result.pkl layout, the iteration-
stepped LR schedule, model_fn_decorator, the --eval_all checkpoint polling, the overlapped gradient all-reduce.
CPU tests use small stand-in modules; the `-m gpu` test runs tools/train.py itself on the synthetic dataset."""
import os
import pickle
import socket
import sys
import time
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tools")
if TOOLS not in sys.path:
    sys.path.insert(0, TOOLS)


def _tiny_detector():
    from cagroup3d_b200 import model_init
    return model_init.seeded_model(18, False, seed=3)


def test_checkpoint_layout_and_resume(tmp_path):
    """checkpoint_state / save_checkpoint (train_utils.py:169-196) -> load_params_with_optimizer / load_params_from_file
    (detector3d_template.py:337-419): the reference's key set, every tensor on the CPU, a bit-exact round trip incl. the
    optimizer moments, (it, epoch) returned for the resume."""
    from train_utils.train_utils import checkpoint_state, save_checkpoint
    m = _tiny_detector()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)
    p = next(m.parameters())
    p.grad = torch.ones_like(p)
    opt.step()
    m.update_global_step()
    state = checkpoint_state(m, opt, epoch=3, it=1234)
    assert set(state) == {"epoch", "it", "model_state", "optimizer_state", "version"}          # train_utils.py:185
    assert state["version"].startswith("pcdet+") and all(v.device.type == "cpu" for v in state["model_state"].values())
    assert "global_step" in state["model_state"] and "backbone_3d.conv1.0.kernel" in state["model_state"]
    save_checkpoint(state, filename=tmp_path / "checkpoint_epoch_3")
    f = str(tmp_path / "checkpoint_epoch_3.pth")
    assert os.path.isfile(f)
    m2 = type(m)(m.model_cfg, 18)
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-3, weight_decay=1e-4)
    it, epoch = m2.load_params_with_optimizer(f, to_cpu=True, optimizer=opt2)
    assert (it, epoch) == (1234, 3) and int(m2.global_step) == 1
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys() and all(torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) for k in s1)
    m3 = type(m)(m.model_cfg, 18)
    assert m3.load_params_from_file(f, to_cpu=True) == []                                   # inference-side loader, same file


def test_result_pkl_round_trip(tmp_path):
    """eval_utils.py:109-110 + scannet_dataset.py:102-127: result.pkl is a list of per-scene dicts with the reference's
    key set and array shapes; empty scenes keep the template's zero-length arrays."""
    from pcdet.datasets import SyntheticIndoorDataset
    names = [f"c{i}" for i in range(18)]
    preds = [{"pred_boxes": torch.rand(5, 7), "pred_scores": torch.rand(5), "pred_labels": torch.randint(0, 18, (5,))},
             {"pred_boxes": torch.zeros(0, 7), "pred_scores": torch.zeros(0), "pred_labels": torch.zeros(0, dtype=torch.long)}]
    annos = SyntheticIndoorDataset.generate_prediction_dicts({"frame_id": np.array([7, 8])}, preds, names)
    with open(tmp_path / "result.pkl", "wb") as f:
        pickle.dump(annos, f)
    back = pickle.load(open(tmp_path / "result.pkl", "rb"))
    keys = {"name", "labels_3d", "bbox", "dimensions", "location", "rotation_y", "scores_3d", "boxes_3d", "frame_id"}
    assert [set(a) for a in back] == [keys, keys]
    a = back[0]
    assert a["boxes_3d"].shape == (5, 7) and a["bbox"].shape == (5, 4) and a["dimensions"].shape == (5, 3)
    assert np.array_equal(a["location"], preds[0]["pred_boxes"][:, :3].numpy()) and a["frame_id"] == 7
    assert list(a["name"]) == [names[i] for i in preds[0]["pred_labels"].tolist()]
    assert back[1]["boxes_3d"].shape == (0, 7) and len(back[1]["scores_3d"]) == 0


def test_scheduler_steps_with_iterations():
    from pcdet.config import EasyDict
    from train_utils.optimization import build_optimizer, build_scheduler
    cfg = EasyDict(OPTIMIZER="adamW", LR=0.001, WEIGHT_DECAY=0.0001, DECAY_STEP_LIST=[7, 9], LR_DECAY=0.1, LR_CLIP=1e-7)
    lin = torch.nn.Linear(2, 2)
    opt = build_optimizer(lin, cfg)
    assert isinstance(opt, torch.optim.AdamW) and opt.param_groups[0]["weight_decay"] == 0.0001
    sched, warm = build_scheduler(opt, total_iters_each_epoch=100, total_epochs=10, last_epoch=-1, optim_cfg=cfg)
    assert warm is None
    lrs = {}
    for it in (0, 699, 700, 899, 900, 999):
        sched.step(it)                                                           # train_utils.py:40
        lrs[it] = opt.param_groups[0]["lr"]
    assert lrs[0] == lrs[699] == 0.001 and abs(lrs[700] - 1e-4) < 1e-12 and abs(lrs[900] - 1e-5) < 1e-12
    with pytest.raises(NotImplementedError):
        build_optimizer(lin, EasyDict(OPTIMIZER="adam_onecycle", LR=1, WEIGHT_DECAY=0))


def test_model_fn_decorator_contract(monkeypatch):
    import pcdet.models as M
    seen = {}

    class Fake(torch.nn.Module):
        steps = 0

        def update_global_step(self):
            Fake.steps += 1

        def forward(self, batch_dict):
            seen.update(batch_dict)
            return {"loss": torch.tensor([1.0, 3.0])}, {"loss_all": 2.0}, {"cur_semantic_value": 0.15}
    monkeypatch.setattr(M, "load_data_to_gpu", lambda b: b.__setitem__("on_device", True))
    r = M.model_fn_decorator()(Fake(), {"points": 1})
    assert float(r.loss) == 2.0 and r.tb_dict == {"loss_all": 2.0} and r.disp_dict["cur_semantic_value"] == 0.15
    assert seen["on_device"] and Fake.steps == 1 and r._fields == ("loss", "tb_dict", "disp_dict")


def test_eval_all_picks_unevaluated_checkpoints_oldest_first(tmp_path):
    import test as T
    rec = tmp_path / "eval_list_val.txt"
    rec.write_text("1\n")
    for e in (1, 2, 3):
        (tmp_path / f"checkpoint_epoch_{e}.pth").write_bytes(b"x")
        os.utime(tmp_path / f"checkpoint_epoch_{e}.pth", (time.time() + e, time.time() + e))
    (tmp_path / "checkpoint_epoch_3_optim.pth").write_bytes(b"x")
    args = types.SimpleNamespace(start_epoch=0)
    assert T.get_no_evaluated_ckpt(tmp_path, rec, args) == ("2", str(tmp_path / "checkpoint_epoch_2.pth"))
    rec.write_text("1\n2\n")
    assert T.get_no_evaluated_ckpt(tmp_path, rec, args)[0] == "3"
    args.start_epoch = 4
    assert T.get_no_evaluated_ckpt(tmp_path, rec, args) == (-1, None)


# ---- overlapped gradient all-reduce (gloo, world size 2) ------------------------------------------------------------
def _overlap_worker(rank, world, port, q):
    import torch.distributed as dist
    from cagroup3d_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(8, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 3))
    unused = torch.nn.Parameter(torch.ones(5))                       # never gets a gradient: reduced as zeros in reduce()
    params = list(m.parameters()) + [unused]
    red = D.GradientAllReducer(params, bucket_mb=0.01)                # several buckets
    ok = len(red.buckets) >= 3
    g = torch.Generator().manual_seed(100 + rank)
    for step in range(2):
        red.zero_grad()
        x = torch.randn((16, 8), generator=g)
        m(x).square().mean().backward()
        launched = red.launched_in_backward
        n = red.reduce()
        ok &= n == len(red.buckets) and launched >= (step + 1) * (len(red.buckets) - 1)       # all but the unused one's bucket
    # the averaged gradient equals the mean of the two ranks' local gradients
    m2 = torch.nn.Sequential(torch.nn.Linear(8, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 3))
    m2.load_state_dict(m.state_dict())
    want = None
    for r in range(world):
        gg = torch.Generator().manual_seed(100 + r)
        torch.randn((16, 8), generator=gg)
        x = torch.randn((16, 8), generator=gg)
        m2.zero_grad()
        m2(x).square().mean().backward()
        cur = [p.grad.clone() for p in m2.parameters()]
        want = cur if want is None else [a + b for a, b in zip(want, cur)]
    for p, w in zip(m.parameters(), want):
        ok &= torch.allclose(p.grad, w / world, rtol=1e-5, atol=1e-7)
    ok &= bool((unused.grad == 0).all())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gradient_allreduce_overlaps_backward_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=180) for _ in ps)
    [p.join(30) for p in ps]
    assert res == [(0, True), (1, True)]


@pytest.mark.gpu
def test_train_py_runs_saves_and_resumes(lib, tmp_path, monkeypatch):
    """tools/train.py on the synthetic ScanNet-shaped dataset: two iterations of epoch 1, checkpoint_epoch_1.pth in the
    reference's layout, then a second invocation resumes from it (epoch 2) and the evaluation loop picks both up."""
    import train as TR
    from pcdet.config import cfg
    monkeypatch.chdir(TOOLS)
    monkeypatch.setattr(cfg, "ROOT_DIR", tmp_path, raising=False)
    common = ["--cfg_file", "cfgs/scannet_models/CAGroup3D.yaml", "--fix_random_seed", "--batch_size", "2", "--workers", "0",
              "--max_iters", "2", "--extra_tag", "t", "--set", "DATA_CONFIG.SYNTHETIC.NUM_SCENES", "4",
              "DATA_CONFIG.SYNTHETIC.VOXELS", "3000", "DATA_CONFIG.REPEAT.train", "1"]
    it = TR.main(["--epochs", "1", "--no_eval"] + common)
    assert it == 2
    ck = tmp_path / "output" / "scannet_models" / "CAGroup3D" / "t" / "ckpt"
    state = torch.load(ck / "checkpoint_epoch_1.pth", map_location="cpu", weights_only=False)
    assert state["epoch"] == 1 and state["it"] == 2 and int(state["model_state"]["global_step"]) == 2
    assert all(torch.isfinite(v).all() for v in state["model_state"].values() if v.is_floating_point())
    it = TR.main(["--epochs", "2"] + common)                                    # resumes: only epoch 2 runs, then evaluates
    assert it == 4 and (ck / "checkpoint_epoch_2.pth").is_file()
    rec = tmp_path / "output" / "scannet_models" / "CAGroup3D" / "t" / "eval" / "eval_with_train" / "eval_list_val.txt"
    assert rec.read_text().split() == ["2"]                                     # num_epochs_to_eval 0 -> start_epoch = 2
