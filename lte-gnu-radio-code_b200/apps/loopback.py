#!/usr/bin/env python3
"""On-device TX -> channel -> RX -> BER/EVM for a named numerology; one JSON line.  Under torchrun every rank
runs `--frames` frames of its own and the counters are all-reduced (NCCL).

    python apps/loopback.py --config ref --frames 100000 --sigma 0.05
    python -m torch.distributed.run --nproc-per-node 8 apps/loopback.py --config lte --multipath
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "python"))

CONFIGS = {
    "ref": dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2),            # reference default
    "ref16": dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4),
    "lte": dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6),    # LTE 20 MHz
}
PDP = [0.16, 0.79, 0.04, 0.01, 0.002]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ref", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--sigma", type=float, default=0.02, help="noise standard deviation per real dimension")
    ap.add_argument("--multipath", action="store_true", help="independent random 5-tap channel per frame")
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from ofdm_b200 import ModemConfig
    from ofdm_b200.loopback import Loopback
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    lb = Loopback(ModemConfig(snr=100.0, gate=0.5, noise_var=max(2 * a.sigma ** 2, 1e-6), **CONFIGS[a.config]))
    taps = lb.taps_random(a.frames, PDP, a.seed + 100 * rank) if a.multipath else None
    r = lb.run(a.frames, seed=a.seed + 1000 * rank, taps=taps, noise_sigma=a.sigma)
    if rank == 0:
        print(json.dumps(dict(config=a.config, n_gpus=world, frames_per_gpu=a.frames, sigma=a.sigma, multipath=a.multipath,
                              **r["summary"])))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
