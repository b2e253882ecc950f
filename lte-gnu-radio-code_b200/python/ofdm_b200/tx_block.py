"""Transmitter source block.  The reference's live transmitter replays row 0 of a pickled complex
array (GRR/gr-TXOFDM/python/tx_signal_transmitter.py:15-27); the waveform itself was produced
offline by MultiAntennaSystem (GRR/LEGACY/gr-ofdm-rx/python/txrx_mod/MultiAntennaSystem.py:113-219).
This block can do either: replay a stored file, or synthesise the waveform on the GPU from payload
bits with the CUDA transmitter (keyword `payload_bits=` + `config=`)."""
import os
import pickle

import numpy as np

from .gr_compat import gr


def load_waveform(path):
    """(1, T) complex array from a .npy or a pickle (binary mode: the reference's text-mode open()
    fails on Python 3).  Only unpickle files you trust."""
    if path.endswith(".npy"):
        a = np.load(path)
    else:
        with open(path, "rb") as f:
            a = pickle.load(f, encoding="latin1")
    a = np.asarray(a)
    return a.reshape(1, -1) if a.ndim == 1 else a


class TxReplayBase(gr.sync_block):
    def __init__(self, pickle_directory, pickle_file, *, payload_bits=None, config=None, n_frames=None,
                 waveform=None, device=None):
        gr.sync_block.__init__(self, name="SimpleTx", in_sig=None, out_sig=[np.complex64])
        if waveform is not None:
            self.tx_data = np.asarray(waveform).reshape(1, -1)
        elif payload_bits is not None:
            self.tx_data = self._synthesise(np.asarray(payload_bits), config, n_frames, device)
        else:
            path = str(pickle_directory) + str(pickle_file)          # the reference concatenates (:20)
            if not os.path.exists(path):
                path = os.path.join(str(pickle_directory), str(pickle_file))
            self.tx_data = load_waveform(path)
        self._row = np.ascontiguousarray(self.tx_data[0, :], dtype=np.complex64)
        self._cursor = 0

    @staticmethod
    def _synthesise(bits, cfg, n_frames, device):
        import torch
        from .modem import Modem
        m = Modem(cfg, device=device)
        per_sym = cfg.num_data_bins * cfg.bps
        if n_frames is None:
            n_frames = bits.size // (cfg.d * per_sym)
        b = bits.reshape(n_frames * cfg.d, per_sym).astype(np.uint8)
        packed = np.packbits(b, axis=1)[:, :cfg.hard_stride].copy()
        out = m.tx_modulate(torch.from_numpy(packed).to(m.device), n_frames)
        return out.cpu().numpy().reshape(1, -1)

    # ---- throughput API (additive) ---------------------------------------------------------
    @staticmethod
    def modulate_batch(bits, config, n_frames=None, device=None):
        """bits -> waveform on the GPU, device tensors in and out: packed payload rows (uint8 CUDA tensor,
        n_frames*D rows of config.hard_stride bytes, MSB first) -> complex64 [n_frames*frame_len].  The arithmetic
        of MultiAntennaSystem.multi_ant_binary_map + multi_ant_symb_gen (MultiAntennaSystem.py:113-219), which is
        what produced the pickles the work() path replays."""
        from .modem import Modem
        key = (id(config), str(device))
        cache = TxReplayBase._modems
        if key not in cache:
            cache.clear()
            cache[key] = Modem(config, device=device)
        m = cache[key]
        if n_frames is None:
            n_frames = bits.numel() // (config.d * config.hard_stride)
        return m.tx_modulate(bits.reshape(-1), n_frames)

    _modems = {}

    def work(self, input_items, output_items):
        """Streams the stored waveform, wrapping around at its end.  The reference copies the whole
        array to out[0:T] on every call (and fails when T exceeds the buffer); the first
        T samples produced here are the same."""
        out = output_items[0]
        n, T = len(out), len(self._row)
        done = 0
        while done < n:
            k = min(n - done, T - self._cursor)
            out[done:done + k] = self._row[self._cursor:self._cursor + k]
            done += k
            self._cursor = (self._cursor + k) % T
        return len(output_items[0])
