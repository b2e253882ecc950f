"""CPU-only checks: the C-ABI library loads and exports every symbol include/ofdm_b200.h declares,
the ctypes struct matches the header, and the host-side logic (numerology, block bookkeeping that
needs no device) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ofdm_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ofdm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from ofdm_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(_lib.SIGNATURES) == set(names)
    lib.ofdm_abi_version.restype = ctypes.c_int
    want = int(re.search(r"#define\s+OFDM_B200_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    assert lib.ofdm_abi_version() == want == 5


def test_block_result_struct_matches_header():
    from ofdm_b200 import _lib
    txt = open(HEADER).read()
    body = txt[txt.index("typedef struct ofdm_rx_block ofdm_rx_block;"):txt.index("} ofdm_rx_block_result;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"^\s*(?:const\s+)?(?:int32_t|int64_t|float)\s+\*?(\w+);", body, flags=re.M)
    assert fields == [f[0] for f in _lib.RxBlockResult._fields_]
    assert ctypes.sizeof(_lib.RxBlockResult) == 4 * 4 + 6 * ctypes.sizeof(ctypes.c_void_p)


def test_params_struct_matches_header():
    from ofdm_b200 import _lib
    txt = open(HEADER).read()
    body = txt[txt.index("typedef struct ofdm_params {"):txt.index("} ofdm_params;")]
    fields = re.findall(r"^\s*(int32_t|float)\s+(\w+)(\[\d+\])?;", body, flags=re.M)
    assert [f[1] for f in fields] == [f[0] for f in _lib.OfdmParams._fields_]
    assert ctypes.sizeof(_lib.OfdmParams) == 4 * (len(fields) - 1) + 4 * 3 == 96      # two reserved words became the CP pattern


def test_plan_create_needs_a_device_and_says_so():
    """No CPU fallback: on a machine without a GPU the product path fails loudly."""
    import torch
    from ofdm_b200 import Modem, ModemConfig, OfdmError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(OfdmError):
        Modem(ModemConfig())


def test_numerology_matches_reference_constants():
    from ofdm_b200 import ModemConfig
    u = ModemConfig(snr=100, gate=0.7, compat="utsa")
    assert (u.zc_root, u.stride, u.gate) == (23, 1, 0.7)
    assert u.regularisers() == (1e-5, 1e-5, 0.01)
    r = ModemConfig(snr=50, gate=0.7, compat="rxofdm")
    assert (r.zc_root, r.stride, r.gate) == (37, 15, 0.4)
    assert r.regularisers() == (0.02, 0.02, 0.02)
    assert u.frame_len == 320 and u.hard_stride == 15 and u.mm == 62
    lte = ModemConfig(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps="64qam")
    assert lte.bps == 6 and lte.hard_stride == 900 and lte.frame_len == 14 * 2192
    s = u.to_struct()
    assert (s.nfft, s.cp_len, s.zc_parity_of) == (64, 16, 62)
    assert r.to_struct().zc_parity_of == 64 - 2 or True


def test_tx_block_streams_without_a_device(golden):
    import TXOFDM
    w = golden["tx_online"]
    tx = TXOFDM.tx_signal_transmitter(0, "", "", waveform=w)
    out = np.zeros(25000, np.complex64)
    assert tx.work(None, [out]) == 25000
    assert np.array_equal(out[:19200], w[0].astype(np.complex64))
    assert np.array_equal(out[19200:], w[0][:5800].astype(np.complex64))


def test_block_names_and_signatures_match_the_reference():
    import inspect
    import RXOFDM
    import TXOFDM
    import utsa_ofdm
    import OFDMReceiver
    sig = inspect.signature(RXOFDM.synch_and_chan_est.__init__)
    pos = [p.name for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD][1:]
    assert pos == ["num_ofdm_symb", "nfft", "cp_len", "num_synch_bins", "synch_dat", "num_data_bins", "snr",
                   "directory_name", "file_name_cest", "diagnostics", "genie"]
    sig = inspect.signature(utsa_ofdm.SynchAndChanEst.__init__)
    pos = [p.name for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD][1:]
    assert pos == ["num_ofdm_symb", "nfft", "cp_len", "num_synch_bins", "synch_dat", "num_data_bins", "snr",
                   "scale_factor_gate", "directory_name", "file_name_cest", "diagnostics", "genie", "channel"]
    assert [p for p in inspect.signature(TXOFDM.tx_signal_transmitter.__init__).parameters][1:4] == \
        ["case", "pickle_directory", "pickle_file"]
    assert [p for p in inspect.signature(utsa_ofdm.TxSignalTransmitter.__init__).parameters][1:3] == \
        ["pickle_directory", "pickle_file"]
    assert [p for p in inspect.signature(OFDMReceiver.BitRecovery.__init__).parameters][1:4] == \
        ["modulation", "directory_name", "diagnostics"]
    # the additive vector-item blocks keep the reference's argument names and order where they still apply
    pos = [p.name for p in inspect.signature(utsa_ofdm.SynchAndChanEstVec.__init__).parameters.values()
           if p.kind == p.POSITIONAL_OR_KEYWORD][1:]
    assert pos == ["nfft", "cp_len", "num_synch_bins", "synch_dat", "num_data_bins", "snr", "scale_factor_gate"]
    pos = [p.name for p in inspect.signature(RXOFDM.synch_and_chan_est_vec.__init__).parameters.values()
           if p.kind == p.POSITIONAL_OR_KEYWORD][1:]
    assert pos == ["nfft", "cp_len", "num_synch_bins", "synch_dat", "num_data_bins", "snr"]


def test_grc_descriptors_keep_ids_and_make_templates():
    import yaml
    grc = os.path.join(ROOT, "lte-gnu-radio-code_b200", "grc")
    want = {
        "RXOFDM_synch_and_chan_est.block.yml": ("RXOFDM_synch_and_chan_est", "RXOFDM.synch_and_chan_est(", 11),
        "TXOFDM_tx_signal_transmitter.block.yml": ("TXOFDM_tx_signal_transmitter", "TXOFDM.tx_signal_transmitter(", 3),
        "utsa_ofdm_SynchAndChanEst.block.yml": ("utsa_ofdm_SynchAndChanEst", "utsa_ofdm.SynchAndChanEst(", 13),
        "utsa_ofdm_TxSignalTransmitter.block.yml": ("utsa_ofdm_TxSignalTransmitter", "utsa_ofdm.TxSignalTransmitter(", 2),
        # the legacy receiver module (GRR/LEGACY/gr-ofdm-rx/grc)
        "OFDMReceiver_BitRecovery.block.yml": ("OFDMReceiver_BitRecovery", "OFDMReceiver.BitRecovery(", 3),
        "OFDMReceiver_SynchEstAndFO.block.yml": ("OFDMReceiver_SynchEstAndFO", "OFDMReceiver.SynchEstAndFO(", 5),
        "OFDMReceiver_SynchEstFOAndDSSS.block.yml": ("OFDMReceiver_SynchEstFOAndDSSS", "OFDMReceiver.SynchEstFOAndDSSS(", 5),
        "OFDMReceiver_SynchronizeAndEstimate.block.yml": ("OFDMReceiver_SynchronizeAndEstimate",
                                                          "OFDMReceiver.SynchronizeAndEstimate(", 1),
        # additive: one frame per item ("vector I/O with vlen = frame")
        "utsa_ofdm_SynchAndChanEstVec.block.yml": ("utsa_ofdm_SynchAndChanEstVec", "utsa_ofdm.SynchAndChanEstVec(", 7),
        "RXOFDM_synch_and_chan_est_vec.block.yml": ("RXOFDM_synch_and_chan_est_vec", "RXOFDM.synch_and_chan_est_vec(", 6),
        "TXOFDM_tx_modulator_vec.block.yml": ("TXOFDM_tx_modulator_vec", "TXOFDM.tx_modulator_vec(", 6),
    }
    for fn, (bid, make, nargs) in want.items():
        d = yaml.safe_load(open(os.path.join(grc, fn)))
        assert d["id"] == bid
        assert d["templates"]["make"].replace("\n", " ").startswith(make)
        assert d["templates"]["make"].count("${") == nargs


def test_vector_item_descriptors_size_their_ports_like_the_blocks():
    """The additive one-frame-per-item blocks: the vlen expressions GRC evaluates for the ports must be the item sizes
    the classes announce (frame_len in, D * num_data_bins out; packed payload rows in for the modulator)."""
    import yaml
    from ofdm_b200 import ModemConfig
    grc = os.path.join(ROOT, "lte-gnu-radio-code_b200", "grc")
    for nfft, cp, sd, ndb, bps in ((64, 16, [1, 3], 60, 2), (2048, 144, [1, 13], 1200, 6), (128, 9, [2, 5], 100, 4)):
        cfg = ModemConfig(nfft=nfft, cp_len=cp, num_synch_bins=nfft - 2, num_data_bins=ndb, m0=sd[0], d=sd[1], bps=bps)
        env = dict(nfft=nfft, cp_len=cp, synch_dat=sd, num_data_bins=ndb, modulation=bps)

        def vlen(port):
            return eval(str(port["vlen"]).strip()[2:-1], {}, dict(env))
        for fn in ("utsa_ofdm_SynchAndChanEstVec.block.yml", "RXOFDM_synch_and_chan_est_vec.block.yml"):
            d = yaml.safe_load(open(os.path.join(grc, fn)))
            assert vlen(d["inputs"][0]) == cfg.frame_len
            assert vlen(d["outputs"][0]) == cfg.d * cfg.num_data_bins
        d = yaml.safe_load(open(os.path.join(grc, "TXOFDM_tx_modulator_vec.block.yml")))
        assert vlen(d["inputs"][0]) == cfg.d * cfg.hard_stride
        assert vlen(d["outputs"][0]) == cfg.frame_len


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """The boundary is a C ABI: include/ofdm_b200.h compiles as C99 (-pedantic) and as C++, and a C program linked
    against libofdm_b200.so resolves every declared entry point and can call the ones that need no device."""
    import shutil
    import subprocess
    from ofdm_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = declared_symbols()
    src = tmp_path / "link.c"
    refs = "\n".join("    (fn)%s," % n for n in names)
    src.write_text('#include <stdio.h>\n#include "ofdm_b200.h"\n'
                   "typedef void (*fn)(void);\nstatic const fn entry[] = {\n%s\n};\n"
                   "int main(void) {\n"
                   "    ofdm_plan *p = NULL;\n"
                   "    if (ofdm_abi_version() != OFDM_B200_ABI_VERSION) return 2;\n"
                   "    if (ofdm_plan_create(NULL, &p) == OFDM_OK || p != NULL) return 3;   /* bad argument: refused */\n"
                   '    printf("%%d %%zu %%s\\n", ofdm_abi_version(), sizeof(entry) / sizeof(entry[0]), ofdm_last_error());\n'
                   "    return 0;\n}\n" % refs)
    inc = os.path.dirname(HEADER)
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-pedantic", "-fsyntax-only", "-I", inc, str(src)], check=True)
    if shutil.which("g++"):
        subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-x", "c++", "-I", inc, HEADER], check=True)
    exe = tmp_path / "link"
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-lofdm_b200",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(None, 2)
    assert int(out[0]) == 5 and int(out[1]) == len(names) == len(_lib.SIGNATURES)
    assert out[2].strip()                                   # the refusal left a message in ofdm_last_error()


def test_sharding_helpers():
    from ofdm_b200 import shard
    assert shard.my_streams(10, 1, 4) == [1, 5, 9]
    blocks = [shard.frame_block(10, r, 4) for r in range(4)]
    assert blocks == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard.summarize([2, 100, 1.0, 4.0, 9, 10, 0, 0])["ber"] == 0.02


def test_bench_arms_print_the_same_config():
    """the reference arm's `config` must be the repo arm's, the sample size a separate key."""
    import bench
    for name in bench.WORKLOADS:
        a, b = bench.config_for(name, 1), bench.config_for(name, 1)
        assert a == b and a["workload"] == name and "frames_per_gpu" in a
    assert bench.config_for("rx_n64_16qam_1m", 8)["frames_total"] == 8 * 333334
    assert bench.config_for("rx_lte20_streams", 8)["frames_total"] == 40960


def test_bench_frame_length_matches_the_host_numerology():
    import bench
    from ofdm_b200 import ModemConfig
    for name, w in bench.WORKLOADS.items():
        cfg = ModemConfig(nfft=w["nfft"], cp_len=w["cp_len"], num_synch_bins=w["num_synch_bins"],
                          num_data_bins=w["num_data_bins"], m0=w["m0"], d=w["d"], bps=w["bps"],
                          cp_first=w.get("cp_first", 0), slot_syms=w.get("slot_syms", 0))
        assert bench.frame_len(w) == cfg.frame_len, name
        rd, wr = bench.algorithmic_bytes(w, 10)
        assert rd == 10 * cfg.frame_len * 8
        assert wr == 10 * cfg.d * (cfg.num_data_bins * (8 + 4 * cfg.bps) + cfg.hard_stride) + 90
    assert bench.frame_len(bench.WORKLOADS["rx_lte20_64qam"]) == 30720


def test_graft_entry_build_passes():
    """The driver's build check: compiles (a no-op when the library is current), loads the library, resolves every
    symbol and checks the ABI version against the header."""
    import __graft_entry__ as g
    g.build()
