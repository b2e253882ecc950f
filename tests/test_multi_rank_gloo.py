"""world_size-2 gloo test of the N>1 host path: frames are partitioned by rank with no data-path
exchange, each rank computes its own BER/EVM counters, and one all-reduce of the 8-entry vector
gives the same totals as a single process over all frames."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frames_and_counters(lo, hi):
    from oracle import ofdm_oracle as oo
    prm = oo.ModemParams(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2, snr=100, gate=0.5)
    cnt = np.zeros(8)
    for f in range(lo, hi):                       # every frame is generated from its own seed
        rng = np.random.default_rng(1000 + f)
        bits = rng.integers(0, 2, 3 * 120)
        tx = oo.tx_modulate(bits, prm, 1)
        rx = oo.add_awgn(np.convolve(tx, oo.reference_channel(8))[:len(tx)], 0.02, rng).astype(np.complex64)
        r = oo.rx_frames(rx, 1, prm)
        got = oo.hard_bits(r["eq"].reshape(-1), 2)
        c = oo.counters(got, bits, r["eq"].reshape(-1), oo.bits_to_symbols(bits, 2))
        cnt[:4] += c
        cnt[4] += r["locked"][0]
        cnt[5] += 1
    return cnt


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ofdm_b200 import shard
    lo, hi = shard.frame_block(n_frames, rank, world)
    cnt = torch.from_numpy(_frames_and_counters(lo, hi))
    shard.reduce_counters(cnt)
    if rank == 0:
        q.put((cnt.numpy().copy(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_counter_reduce_matches_single_process():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.pathsep.join(
        [root, os.path.join(root, "lte-gnu-radio-code_b200", "python"), os.environ.get("PYTHONPATH", "")])
    n_frames = 21
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, (lo, hi) = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (lo, hi) == (0, 11)
    single = _frames_and_counters(0, n_frames)
    assert np.allclose(total, single, rtol=1e-12, atol=0)
    assert total[5] == n_frames and total[1] == n_frames * 360
