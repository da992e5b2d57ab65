"""Host logic of the data-parallel trainer (<pkg>/trainer.py) without a GPU: the batch plan is a pure function of
(seed, epoch), shards tile every global batch exactly once over a world-size-2 gloo group, a trailing batch smaller than
the world is dropped (every rank must take part in every step: the EMA finalize kernels barrier across ranks), and the
command line mirrors scripts/train_ablation.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_cli_mirrors_reference_script():
    import vqb200
    from vqb200 import trainer
    a = trainer.build_parser().parse_args([])
    # defaults of scripts/train_ablation.py:376-391
    assert (a.method, a.arch, a.epochs, a.batch_size, a.seed, a.window, a.patience, a.mode, a.resume, a.teacher_ckpt) == \
           ("hybrid", "transformer", 400, 256, [42], 64, -1, "teacher", False, None)
    a = trainer.build_parser().parse_args("--mode student --teacher_ckpt x.pth --seed 1 2 --window 10 --cuda_graph".split())
    assert a.mode == "student" and a.seed == [1, 2] and a.cuda_graph


def test_batch_plan_is_deterministic_and_covers_training_set():
    import vqb200
    from vqb200 import trainer
    tr, va = trainer.split_indices(1000, seed=5)
    tr2, va2 = trainer.split_indices(1000, seed=5)
    assert np.array_equal(tr, tr2) and np.array_equal(va, va2)
    assert len(tr) == 900 and len(va) == 100 and len(set(tr) | set(va)) == 1000
    b0 = trainer.epoch_batches(tr, 128, 5, 0)
    assert all(np.array_equal(x, y) for x, y in zip(b0, trainer.epoch_batches(tr, 128, 5, 0)))
    assert not np.array_equal(b0[0], trainer.epoch_batches(tr, 128, 5, 1)[0])          # reshuffled every epoch
    assert sorted(np.concatenate(b0).tolist()) == sorted(tr.tolist())                   # 7 full batches + a tail of 4
    assert [len(b) for b in b0] == [128] * 7 + [4]


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vqb200
    from vqb200 import trainer
    vqb200.dist.enable()                                        # CPU: NCCL-style transport, reason recorded
    assert vqb200.dist.peer_exchange() is None and "no CUDA" in vqb200.dist.peer_status()
    tr, _ = trainer.split_indices(515, seed=3)                  # 463 training windows
    plan = trainer.epoch_batches(tr, 2 * 33, 3, 0)              # global batch = batch_size x world; tail of 1 < world: dropped
    assert sum(len(b) for b in plan) == 462 and all(len(b) >= world for b in plan)
    seen = torch.zeros(515)
    for b in plan:
        mine = trainer._shard(b)
        assert abs(len(mine) - len(b) / world) <= 0.5
        seen[torch.from_numpy(mine)] += 1
    dist.all_reduce(seen)
    expect = torch.zeros(515)
    expect[torch.from_numpy(np.concatenate(plan))] = 1
    assert torch.equal(seen, expect)                            # every planned sample is trained on by exactly one rank
    t, = trainer._all_sum([float(rank + 1)], torch.device("cpu"))
    assert t == 3.0
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    vqb200.dist.disable()
    dist.destroy_process_group()


def test_two_rank_batch_sharding(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))
