"""Data-parallel training launcher with the command line of the reference's scripts/train_ablation.py.
One process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
      tools/train_ddp.py --mode teacher --arch transformer --method hybrid --window 10 --epochs 400 --batch_size 512
(or plain `python tools/train_ddp.py ...` on one GPU).  See <pkg>/trainer.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vqb200  # noqa: E402
from vqb200 import trainer  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(trainer.main())
