#!/usr/bin/env python3
"""Headless equivalent of GRR/ofdm_chain.py:40-91 (TX block -> RX block -> null sink).

The reference flowgraph is a Qt GUI top_block; the GUI shell is out of scope, the wiring is not:
the same two blocks are instantiated with the same constructor arguments and connected
source -> sync block -> sink.  With GNU Radio installed the blocks can be dropped into the original
ofdm_chain.py unchanged; without it this file's 20-line scheduler moves the buffers.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "python"))


class null_sink:
    def work(self, input_items, output_items):
        return len(input_items[0])


class ofdm_chain:
    """source -> RX -> sink with fixed-size scheduler buffers."""

    def __init__(self, tx_block, rx_block, sink=None):
        self.tx, self.rx, self.sink = tx_block, rx_block, sink or null_sink()
        self.last_in = None
        self.last_out = None

    def run(self, n_calls, buffer_len):
        for _ in range(n_calls):
            a = np.zeros(buffer_len, np.complex64)
            b = np.zeros(buffer_len, np.complex64)
            na = self.tx.work(None, [a])
            nb = self.rx.work([a[:na]], [b[:na]])
            self.sink.work([b[:nb]], None)
            self.last_in, self.last_out = a[:na], b[:nb]


def run_loopback(num_ofdm_symb=24, n_calls=3, seed=1234, variant="utsa", snr=50):
    """BASELINE config 1.  Payload = default_rng(seed) bits, waveform synthesised by the CUDA
    transmitter, no channel; returns bit errors of the RX block's written output vs the payload.

    The receiver keeps the reference's cross-call state: from the second call on the
    utsa variant re-locks one frame later than the previous call did, so row i of its output holds
    frame i+shift and rows whose window falls off the buffer keep their previous content.  The
    expected payload order is tracked here from the block's own time_synch_ref.
    """
    import utsa_ofdm
    import RXOFDM
    from ofdm_b200 import ModemConfig
    nfft, cp, nsb, ndb, sd = 64, 16, 62, 60, [1, 3]
    L, fsyms = nfft + cp, sum(sd)
    cfg = ModemConfig(nfft=nfft, cp_len=cp, num_synch_bins=nsb, num_data_bins=ndb, m0=sd[0], d=sd[1], bps=2)
    n_frames = num_ofdm_symb // fsyms
    bits = np.random.default_rng(seed).integers(0, 2, n_frames * sd[1] * ndb * 2)
    per_frame = bits.reshape(n_frames, sd[1], ndb * 2)
    tx = utsa_ofdm.TxSignalTransmitter("", "", payload_bits=bits, config=cfg)
    if variant == "utsa":
        rx = utsa_ofdm.SynchAndChanEst(num_ofdm_symb, nfft, cp, nsb, sd, ndb, snr, 0.7, "/tmp/", "output_file", 0, 0)
    else:
        rx = RXOFDM.synch_and_chan_est(num_ofdm_symb, nfft, cp, nsb, sd, ndb, snr, "/tmp/", "output_file", 0, 0,
                                       zc_root=23)
    chain = ofdm_chain(tx, rx)
    buf = num_ofdm_symb * L
    content = [None] * n_frames                   # which payload frame each output row-group holds
    for _ in range(n_calls):
        chain.run(1, buf)
        if variant == "utsa":
            pos, d = int(rx.time_synch_ref[0]), int(rx.time_synch_ref[1])
            shift = int(round((pos - cp + d) / float(fsyms * L)))
            for i in range(n_frames):
                if pos + sd[0] * L * (i * fsyms + 1) + nfft - 1 <= buf and i + shift < n_frames:
                    content[i] = i + shift
        else:
            content = list(range(n_frames))
    n = rx.last_written
    out = chain.last_out[:n]
    hard = np.stack(((out.real < 0), (out.imag < 0)), axis=1).reshape(-1).astype(np.int64)
    if variant == "utsa":
        ref = np.concatenate([per_frame[c].reshape(-1) for c in content])
    else:
        ref = np.concatenate([per_frame[c][0] for c in content])
    return dict(bit_errors=int(np.count_nonzero(hard != ref[:hard.size])), bits=int(hard.size), written=int(n),
                tx_wave=chain.last_in.copy(), rx_out=chain.last_out.copy(), content=content)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--symbols", type=int, default=24)
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--variant", default="utsa", choices=["utsa", "rxofdm"])
    a = ap.parse_args()
    r = run_loopback(a.symbols, a.calls, variant=a.variant)
    print("bit errors %d / %d (written %d items)" % (r["bit_errors"], r["bits"], r["written"]))
