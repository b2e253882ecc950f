"""OFDMReceiver -- the legacy receiver module (GRR/LEGACY/gr-ofdm-rx): the sink block that turns
equalised symbols into soft/hard bits (GRC template OFDMReceiver.BitRecovery(modulation,
directory_name, diagnostics), grc/OFDMReceiver_BitRecovery.block.yml:24) and the receiver with a
frequency-offset hypothesis bank (OFDMReceiver.SynchEstAndFO(case, fo_range, directory_name,
file_name_cest, diagnostics), grc/OFDMReceiver_SynchEstAndFO.block.yml:33-35) and its variant that also
despreads DSSS-coded data bins (OFDMReceiver.SynchEstFOAndDSSS, same arguments), and the pointer-tracking receiver
(OFDMReceiver.SynchronizeAndEstimate(case), grc/OFDMReceiver_SynchronizeAndEstimate.block.yml)."""
from .BitRecovery import BitRecovery
from .SynchEstAndFO import SynchEstAndFO
from .SynchEstFOAndDSSS import SynchEstFOAndDSSS
from .SynchronizeAndEstimate import SynchronizeAndEstimate

__all__ = ["BitRecovery", "SynchEstAndFO", "SynchEstFOAndDSSS", "SynchronizeAndEstimate"]
