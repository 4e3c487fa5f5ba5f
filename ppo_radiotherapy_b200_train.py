#!/usr/bin/env python
"""Launcher of the device-resident PPO loop (the reference's `python train.py --config-file X --output-dir Y`,
train.py:285-324):

    python ppo_radiotherapy_b200_train.py --config-file cfg.yaml --output-dir out
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        ppo_radiotherapy_b200_train.py --config-file cfg.yaml --output-dir out
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from ppo_radiotherapy_b200.train import main  # noqa: E402

if __name__ == "__main__":
    main()
