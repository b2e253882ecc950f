#!/usr/bin/env python3
"""Extract the reference's in-tree data fixtures into plain .npy files WITHOUT unpickling.

The reference keeps its only golden data as pickled NumPy arrays:

  GNU-Radio-Repositories/TEST/GNU_RADIO_OFFLINE/Data/tx_bit_data_chan_type_Fading_SNR_100.pckl
  GNU-Radio-Repositories/TEST/GNU_RADIO_OFFLINE/Data/tx_data_online_chan_type_Fading_SNR_100.pckl
  GNU-Radio-Repositories/TEST/GNU_RADIO_OFFLINE/Data/tx_data_offline_chan_type_Fading_SNR_100.pckl
  GNU-Radio-Repositories/TEST/GNU_RADIO_OFFLINE/Output/_output_data.pckl
  GNU-Radio-Repositories/LEGACY/gr-ofdm-tx/python/tx_data_0.pckl

Unpickling untrusted files executes code, so this script never calls pickle.load.  It walks the
opcode stream with pickletools.genops (which executes nothing), checks that the only globals
named are the ones numpy's array reducer uses, and pulls shape / dtype / raw payload out of the
opcode arguments.  The payload is then viewed with np.frombuffer.

Run (in the build container, where /root/reference is mounted):
    python tests/golden/make_golden.py
Outputs go next to this script; they are committed so the GPU box (which has no /root/reference)
can run the golden tests.
"""
import hashlib
import os
import pickletools
import sys

import numpy as np

REF = os.environ.get("OFDM_REFERENCE_ROOT", "/root/reference")
GRR = os.path.join(REF, "GNU-Radio-Repositories")
HERE = os.path.dirname(os.path.abspath(__file__))

FIXTURES = {
    "tx_bits": "TEST/GNU_RADIO_OFFLINE/Data/tx_bit_data_chan_type_Fading_SNR_100.pckl",
    "tx_online": "TEST/GNU_RADIO_OFFLINE/Data/tx_data_online_chan_type_Fading_SNR_100.pckl",
    "rx_offline": "TEST/GNU_RADIO_OFFLINE/Data/tx_data_offline_chan_type_Fading_SNR_100.pckl",
    "chan_est_tim": "TEST/GNU_RADIO_OFFLINE/Output/_output_data.pckl",
    "tx_data_0": "LEGACY/gr-ofdm-tx/python/tx_data_0.pckl",
}

ALLOWED_GLOBALS = {
    ("numpy.core.multiarray", "_reconstruct"),
    ("numpy", "ndarray"),
    ("numpy", "dtype"),
    ("_codecs", "encode"),
}


def static_ndarray(path):
    """Return the single ndarray stored in `path`, reconstructed from opcode arguments only."""
    raw = open(path, "rb").read()
    strings = []          # unicode / bytes arguments in stream order
    ints = []             # (pos, value)
    globals_seen = []
    pending = []          # string-valued stack pushes, for STACK_GLOBAL (module, attr)
    memo = {}             # memo index -> last pushed value (strings only matter)
    last = None
    n_memo = 0
    for op, arg, pos in pickletools.genops(raw):
        name = op.name
        if name == "GLOBAL":
            mod, attr = arg.split(" ")
            globals_seen.append((mod, attr))
            last = None
        elif name == "STACK_GLOBAL":
            globals_seen.append((pending[-2], pending[-1]))
            last = None
        elif name in ("SHORT_BINUNICODE", "BINUNICODE", "UNICODE"):
            pending.append(arg)
            strings.append(arg)
            last = arg
        elif name in ("SHORT_BINBYTES", "BINBYTES", "BINBYTES8"):
            strings.append(arg)
            last = None
        elif name in ("BININT", "BININT1", "BININT2", "LONG1"):
            ints.append(arg)
            last = None
        elif name == "MEMOIZE":
            memo[n_memo] = last
            n_memo += 1
        elif name in ("BINPUT", "LONG_BINPUT"):
            memo[arg] = last
        elif name in ("BINGET", "LONG_BINGET"):
            last = memo.get(arg)
            if isinstance(last, str):
                pending.append(last)
        else:
            last = None
    for g in globals_seen:
        if g not in ALLOWED_GLOBALS:
            raise RuntimeError("%s: unexpected global %r - refusing to touch it" % (path, g))
    # dtype code is the short string right after the 'dtype' global: 'i4', 'c16', ...
    dtype_code = None
    for s in strings:
        if isinstance(s, str) and s in ("i4", "i8", "c16", "c8", "f8", "f4"):
            dtype_code = s
            break
    if dtype_code is None:
        raise RuntimeError("%s: dtype code not found" % path)
    # the payload is the largest bytes/str argument
    big = max(strings, key=lambda s: len(s))
    if isinstance(big, str):          # protocol-2 writes bytes as latin1 text through _codecs.encode
        payload = big.encode("latin1")
    else:
        payload = bytes(big)
    dt = np.dtype("<" + dtype_code)
    arr = np.frombuffer(payload, dtype=dt)
    # shape: the ints that follow the reducer's version int; the stream holds (ndim-entries) just
    # before the dtype.  All fixtures are (1, n): verify instead of guessing.
    n = arr.size
    if (1 not in ints) or (n not in ints):
        raise RuntimeError("%s: shape (1,%d) not confirmed by the opcode stream (ints=%r)" % (path, n, ints[:8]))
    return arr.reshape(1, n).copy(), hashlib.sha256(raw).hexdigest()


def main():
    manifest = []
    for key, rel in FIXTURES.items():
        path = os.path.join(GRR, rel)
        if not os.path.exists(path):
            print("missing:", path, file=sys.stderr)
            return 1
        arr, sha = static_ndarray(path)
        out = os.path.join(HERE, key + ".npy")
        if key == "tx_bits":
            assert set(np.unique(arr)) <= {0, 1}
            np.save(out, np.packbits(arr.astype(np.uint8), axis=1))   # (1, 2700) bytes, MSB first
            note = "packed bits (np.packbits), %d bits" % arr.size
        else:
            np.save(out, arr)
            note = "%s %r" % (arr.dtype, arr.shape)
        manifest.append("%s  sha256[:16]=%s  %s  <- %s" % (key + ".npy", sha[:16], note, rel))
        print(manifest[-1])
    with open(os.path.join(HERE, "MANIFEST.txt"), "w") as f:
        f.write("# written by make_golden.py; source = reference fixtures, extracted statically (no unpickling)\n")
        f.write("\n".join(manifest) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
