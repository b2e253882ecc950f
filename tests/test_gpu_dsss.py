"""GPU parity of the DSSS despreading kernel and of the legacy block built on it
(GRR/LEGACY/gr-ofdm-rx/python/SynchEstFOAndDSSS.py:253-261, :392-397)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import ofdm_oracle as oo  # noqa: E402
from oracle.blocks_oracle import SynchEstFOAndDSSSOracle  # noqa: E402
from test_gpu_parity import make, rel_err, RTOL  # noqa: E402
from test_oracle_dsss import spread_waveform  # noqa: E402


@pytest.mark.parametrize("row_len,sf", [(36, 6), (48, 12), (120, 24), (12, 1), (250, 16), (3840, 7)])
def test_despread_kernel_matches_oracle(row_len, sf):
    m = make(dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2))
    rng = np.random.default_rng(sf)
    rows = (rng.standard_normal((37, row_len)) + 1j * rng.standard_normal((37, row_len))).astype(np.complex64)
    code = oo.dsss_code(sf, 37)
    got = m.dsss_despread(torch.from_numpy(rows).cuda(), torch.from_numpy(code.astype(np.complex64)).cuda()).cpu().numpy()
    ref = oo.dsss_despread(rows, code)
    assert got.shape == ref.shape == (37, row_len // sf)
    assert rel_err(got, ref) < RTOL
    assert m.dsss_despread(torch.zeros((0, row_len), dtype=torch.complex64, device="cuda"),
                           torch.from_numpy(code.astype(np.complex64)).cuda()).shape == (0, row_len // sf)


@pytest.mark.parametrize("case,fo_range", [(1, [0.0]), (4, [-3000.0, 0.0, 3000.0]), (8, [0.0, 2000.0])])
def test_dsss_block_matches_its_oracle(case, fo_range):
    import OFDMReceiver
    blk = OFDMReceiver.SynchEstFOAndDSSS(case, fo_range, "/tmp/", "cest", 0)
    ora = SynchEstFOAndDSSSOracle(case, fo_range)
    assert np.allclose(blk.SC, ora.SC)
    syms, rx = spread_waveform(ora, 9, 40 + case, tail=30)
    out_b = np.zeros(len(rx), np.complex64)
    out_o = np.zeros(len(rx), np.complex64)
    for call in range(2):
        assert blk.work([rx], [out_b]) == len(rx)
        assert ora.work(rx, out_o) == len(rx)
        assert np.array_equal(blk.time_synch_ref, ora.time_synch_ref), call
        assert blk.dmax_tmp_ind == ora.dmax_tmp_ind and blk.last_written == ora.last_written > 0
        assert rel_err(blk.est_data_freq, ora.est_data_freq) < RTOL
        assert rel_err(blk.est_data_freq_d, ora.est_data_freq_d) < RTOL
        n = ora.last_written
        assert rel_err(out_b[:n], out_o[:n]) < RTOL
    nloc = int(np.count_nonzero(ora.time_synch_ref[:, 0]))
    assert nloc >= 8
    if len(fo_range) == 1:                                      # no hypothesis quirks: the despread symbols are the transmitted ones
        got = blk.est_data_freq_d[:8]
        assert np.array_equal(oo.hard_bits(got.reshape(-1), 2), oo.hard_bits(syms[:8].reshape(-1), 2))
