"""OFDMReceiver -- the legacy sink block that turns equalised symbols into soft/hard bits
(GRC template: OFDMReceiver.BitRecovery(modulation, directory_name, diagnostics),
GRR/LEGACY/gr-ofdm-rx/grc/OFDMReceiver_BitRecovery.block.yml:24)."""
from .BitRecovery import BitRecovery

__all__ = ["BitRecovery"]
