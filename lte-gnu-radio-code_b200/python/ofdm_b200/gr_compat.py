"""GNU Radio shim.  When gnuradio is installed the blocks derive from gr.sync_block and can be
connected in a real flowgraph; otherwise a minimal stand-in keeps them importable and callable
(the reference's own offline harness does the same by commenting out gr.sync_block:
GRR/TEST/GNU_RADIO_OFFLINE/synch_and_chan_est.py:13-17,116-119)."""
try:
    from gnuradio import gr  # noqa: F401
    HAVE_GNURADIO = True
except Exception:  # pragma: no cover - depends on the host
    HAVE_GNURADIO = False

    class _SyncBlock:
        def __init__(self, name=None, in_sig=None, out_sig=None):
            self._name = name
            self.in_sig = in_sig
            self.out_sig = out_sig

        def name(self):
            return self._name

    class gr:  # noqa: N801 - mirrors the module name
        sync_block = _SyncBlock
