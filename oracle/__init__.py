"""CPU oracle for the OFDM modem hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in NumPy complex128 (the precision the
reference computes in), the arithmetic of the reference's modem path so that the CUDA path can be
checked against it.  Nothing under ``lte-gnu-radio-code_b200/`` may import it; the only callers
are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` (where it is the thing being *compared against*, never the product).

Parity status: PINNED for the reference's own configuration (N=64, CP=16, 62 Zadoff-Chu bins with
root 23, 60 QPSK data bins, frame pattern [1 synch, 3 data]) by the five data fixtures the
reference ships (tests/golden/, extracted by tests/golden/make_golden.py): the TX chain reproduces
the stored waveform to ~1e-15 and the RX chain recovers the stored payload with 0 bit errors.
UNPINNED ("our definition", no reference arithmetic or fixture exists): 16/64-QAM mapping and
max-log LLRs, comb pilots with frequency interpolation, ZF, any N other than 64, and the numeric
values of the QPSK LLRs (the reference stores no soft bits).  See DESIGN.md section 3.
"""
