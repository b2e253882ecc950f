"""ofdm_b200 -- B200 (sm_100a) implementation of the OFDM modem hot path, driven through the
C ABI in include/ofdm_b200.h.  The drop-in GNU Radio blocks live in the sibling packages
RXOFDM, TXOFDM and utsa_ofdm."""
from ._lib import OfdmError, LIB_PATH, lib
from .modem import Modem, ModemConfig, launch_count

__all__ = ["Modem", "ModemConfig", "OfdmError", "LIB_PATH", "lib", "launch_count"]
