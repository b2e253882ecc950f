"""OFDMReceiver.BitRecovery -- CUDA-backed drop-in for
GRR/LEGACY/gr-ofdm-rx/python/BitRecovery.py (ctor :36-59, work :66-189).  QPSK only, like the
reference.  The file/CSV/MAT side effects (:159-184) are not reproduced; the results of the last
call are kept on the instance (softbit0, softbit1, hardbit)."""
import numpy as np

from ofdm_b200.gr_compat import gr
from ofdm_b200.modem import Modem, ModemConfig


class BitRecovery(gr.sync_block):
    def __init__(self, modulation, directory_name, diagnostics, *, device=None):
        gr.sync_block.__init__(self, name="BitRecovery", in_sig=[np.complex64], out_sig=None)
        self.modulation = modulation
        self.directory_name = directory_name
        self.diagnostics = diagnostics
        self.K = 1.414213562373095
        self.count = 0
        self.modem = Modem(ModemConfig(), device=device)     # numerology is irrelevant to this block
        self.softbit0 = self.softbit1 = self.hardbit = None

    def work(self, input_items, output_items):
        import torch
        in0 = np.ascontiguousarray(input_items[0], dtype=np.complex64)
        if len(in0):
            x = torch.from_numpy(in0).to(self.modem.device)
            l0, l1, hb = self.modem.bit_recovery(x)
            self.softbit0 = l0.cpu().numpy()
            self.softbit1 = l1.cpu().numpy()
            self.hardbit = hb.cpu().numpy().astype(int)[:, np.newaxis]
        self.count += 1
        return len(input_items[0])
