"""Stateless NumPy restatement of the reference's OFDM modem arithmetic (TEST INFRASTRUCTURE ONLY).

Every function cites the reference file:line it follows, relative to
/root/reference/GNU-Radio-Repositories/ (abbreviated GRR/).  All arithmetic is complex128 like the
reference (NumPy promotes the complex64 stream).  Two styles are provided where the reference's
cost structure matters for the CPU baseline:

  * ``literal=True``  keeps the reference's dense ``diag`` matmuls and Python-level loops
    (this is what a GNU Radio thread running the reference block actually executes);
  * ``literal=False`` is the mathematically identical element-wise form.

Parity status is described in oracle/__init__.py (pinned for N=64/QPSK by the golden fixtures;
"our definition" for everything the reference has no arithmetic for).
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

TWO_PI = 2.0 * np.pi


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
@dataclass
class ModemParams:
    """Numerology of one modem instance.

    nfft, cp_len, num_synch_bins, synch_dat=[M0, D], num_data_bins are the reference block's ctor
    arguments (GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:17-19).  ``compat`` selects which of the
    two live RX variants' constants are reproduced:

      "utsa"   : ZC root 23, search stride 1, gate = scale_factor_gate, regulariser 10**(snr/20)
      "rxofdm" : ZC root 37, search stride CP-1, gate 0.4, regulariser = snr taken as linear
    """
    nfft: int = 64
    cp_len: int = 16
    num_synch_bins: int = 62
    num_data_bins: int = 60
    m0: int = 1                      # consecutive synch symbols per frame  (synch_dat[0])
    d: int = 3                       # data symbols per frame               (synch_dat[1])
    bps: int = 2                     # bits per data bin: 1 BPSK, 2 QPSK, 4 16-QAM, 6 64-QAM
    snr: float = 100.0               # the block's "snr" ctor argument, meaning depends on compat
    gate: float = 0.7                # scale_factor_gate (utsa); rxofdm hard-codes 0.4
    compat: str = "utsa"
    zc_root: Optional[int] = None    # default: 23 (utsa) / 37 (rxofdm)
    equalizer: str = "mmse"          # "mmse" (reference) or "zf" (extension: regulariser -> 0)
    sync_advance: int = 0            # framed mode: how many samples before the CP end the window opens
    # comb pilots in the data symbols (EXTENSION, our definition; the reference has ref_sigs = 0):
    pilot_spacing: int = 0           # 0 = none; else every pilot_spacing-th used bin (list order) is a pilot
    pilot_offset: int = -1           # first pilot position in the used-bin list; -1 -> pilot_spacing // 2
    pilot_root: int = 25             # Zadoff-Chu root of the pilot sequence
    # LTE "normal CP" pattern (EXTENSION, our definition; the reference assumes one CP length, rx_b_len =
    # nfft + cp_len, SynchAndChanEst.py:79): symbol k of a frame has a prefix of cp_first samples when
    # k % slot_syms == 0 and of cp_len samples otherwise (TS 36.211 5.6: 160/144 at 30.72 Msps, 7 symbols per slot).
    cp_first: int = 0                # 0 = every symbol uses cp_len
    slot_syms: int = 0

    def __post_init__(self):
        if self.zc_root is None:
            self.zc_root = 23 if self.compat == "utsa" else 37
        if self.compat == "rxofdm":
            self.gate = 0.4
        if self.pilot_spacing and self.pilot_offset < 0:
            self.pilot_offset = self.pilot_spacing // 2

    # derived -------------------------------------------------------------------------------
    @property
    def sym_len(self):               # rx_b_len, GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:79
        return self.nfft + self.cp_len

    @property
    def frame_syms(self):
        return self.m0 + self.d

    @property
    def cp_pattern(self):
        return self.cp_first > 0 and self.slot_syms > 0 and self.cp_first != self.cp_len

    def sym_cp(self, k):
        """prefix length of symbol k of a frame"""
        return self.cp_first if self.cp_pattern and k % self.slot_syms == 0 else self.cp_len

    def sym_start(self, k):
        """offset of the first prefix sample of symbol k inside its frame (k = frame_syms gives the frame length)"""
        return sum(self.nfft + self.sym_cp(i) for i in range(k))

    @property
    def frame_len(self):
        return self.sym_start(self.frame_syms)

    @property
    def mm(self):                    # MM, SynchAndChanEst.py:48-49
        return self.m0 * self.num_synch_bins

    @property
    def pilot_pos(self):
        """positions (in the used-bin list) that carry pilots"""
        if not self.pilot_spacing:
            return np.zeros(0, dtype=np.int64)
        return np.arange(self.pilot_offset, self.num_data_bins, self.pilot_spacing, dtype=np.int64)

    @property
    def data_pos(self):
        """positions (in the used-bin list) that carry payload"""
        m = np.ones(self.num_data_bins, dtype=bool)
        m[self.pilot_pos] = False
        return np.nonzero(m)[0]

    @property
    def n_payload_bins(self):
        return self.num_data_bins - len(self.pilot_pos)

    @property
    def stride(self):                # SynchAndChanEst.py:77 vs gr-RXOFDM/python/synch_and_chan_est.py:81
        return 1 if self.compat == "utsa" else self.cp_len - 1

    @property
    def reg_chanest(self):
        """x in  H = r*conj(zc) / (1 + 1/x).  utsa: SynchAndChanEst.py:99,180; rxofdm: :184."""
        return 10.0 ** (self.snr / 20.0) if self.compat == "utsa" else float(self.snr)

    @property
    def reg_eq(self):
        """x in  g = conj(H)/(|H|^2 + 1/x) for DATA symbols.  utsa :245; rxofdm :247."""
        if self.equalizer == "zf":
            return np.inf
        return 10.0 ** (self.snr / 20.0) if self.compat == "utsa" else float(self.snr)

    @property
    def reg_eq_synch(self):
        """x for the synch-symbol equaliser (raw snr in both variants: utsa :214, rxofdm :217)."""
        return float(self.snr)


# --------------------------------------------------------------------------------------------
# a1  bin maps
# --------------------------------------------------------------------------------------------
def signed_bins(n):
    """-n/2..-1, 1..n/2  (DC skipped).  SynchAndChanEst.py:38-39, :68-69; SDRScript.py:61."""
    h = int(n / 2)
    return list(range(-h, 0)) + list(range(1, h + 1))


def bins_used(n, nfft):
    """FFT indices of the used bins: (b + N) % N.  SynchAndChanEst.py:41,:70; MultiAntennaSystem.py:14."""
    return (np.array(signed_bins(n), dtype=np.int64) + nfft) % nfft


# --------------------------------------------------------------------------------------------
# a2  Zadoff-Chu
# --------------------------------------------------------------------------------------------
def zadoff_chu(mm, root, parity_of=None):
    """zc[k] = exp(-j*2pi/MM * p * k^2/2)  (even)  or  exp(-j*2pi/MM * p * k(k+1)/2)  (odd).

    utsa / TX simulator test the parity of MM (SynchAndChanEst.py:54-59, SynchSignal.py:24-30);
    gr-RXOFDM tests the parity of num_synch_bins instead (synch_and_chan_est.py:54-64): pass it as
    ``parity_of``.
    """
    k = np.arange(int(mm), dtype=np.float64)
    par = mm if parity_of is None else parity_of
    if par % 2 == 0:
        q = k * k / 2.0
    else:
        q = k * (k + 1.0) / 2.0
    return np.exp(-1j * (TWO_PI / mm) * root * q)


def zc_for(prm: ModemParams):
    par = prm.mm if prm.compat == "utsa" else prm.num_synch_bins
    return zadoff_chu(prm.mm, prm.zc_root, par)


# --------------------------------------------------------------------------------------------
# a3  bit -> symbol map
# --------------------------------------------------------------------------------------------
_QPSK_EIGHTHS = np.array([1.0, -1.0, 3.0, 5.0])       # MultiAntennaSystem.py:171-178


def pam_levels(bits_per_axis):
    """Gray-coded PAM amplitude for each bit pattern (MSB first), 3GPP TS 36.211 7.1 layout.

    1 bit : (1-2b0)
    2 bits: (1-2b0)*(2-(1-2b1))
    3 bits: (1-2b0)*(4-(1-2b1)*(2-(1-2b2)))
    Unpinned extension (the reference only has BPSK/QPSK).  Returned un-normalised.
    """
    n = 1 << bits_per_axis
    lev = np.zeros(n)
    for v in range(n):
        b = [(v >> (bits_per_axis - 1 - i)) & 1 for i in range(bits_per_axis)]
        if bits_per_axis == 1:
            a = 1 - 2 * b[0]
        elif bits_per_axis == 2:
            a = (1 - 2 * b[0]) * (2 - (1 - 2 * b[1]))
        else:
            a = (1 - 2 * b[0]) * (4 - (1 - 2 * b[1]) * (2 - (1 - 2 * b[2])))
        lev[v] = a
    return lev


QAM_NORM = {2: np.sqrt(2.0), 4: np.sqrt(10.0), 6: np.sqrt(42.0)}


def constellation(bps):
    """All 2**bps points indexed by the MSB-first decimal value of the bin's bits."""
    if bps == 1:
        return np.array([-1.0 + 0j, 1.0 + 0j])          # 2*b-1, MultiAntennaSystem.py:156-157
    if bps == 2:
        return np.exp(1j * TWO_PI / 8.0 * _QPSK_EIGHTHS)  # MultiAntennaSystem.py:171-178
    half = bps // 2
    lev = pam_levels(half) / QAM_NORM[bps]
    pts = np.zeros(1 << bps, dtype=complex)
    for v in range(1 << bps):
        # even-position bits (b0,b2,b4) drive I, odd-position bits (b1,b3,b5) drive Q -- the same
        # convention as the reference's QPSK (b0 -> sign of Re, b1 -> sign of Im).
        bi = bq = 0
        for i in range(half):
            bi = (bi << 1) | ((v >> (bps - 1 - 2 * i)) & 1)
            bq = (bq << 1) | ((v >> (bps - 2 - 2 * i)) & 1)
        pts[v] = lev[bi] + 1j * lev[bq]
    return pts


def bits_to_symbols(bits, bps):
    """bits: [..., n*bps] of 0/1 -> complex [..., n].  MSB first (MultiAntennaSystem.py:160-170)."""
    bits = np.asarray(bits)
    n = bits.shape[-1] // bps
    b = bits.reshape(bits.shape[:-1] + (n, bps)).astype(np.int64)
    w = 1 << np.arange(bps - 1, -1, -1)
    dec = (b * w).sum(-1)
    return constellation(bps)[dec]


# --------------------------------------------------------------------------------------------
# a3/a4  frequency-grid assembly (bit map + synch-symbol insertion)
# --------------------------------------------------------------------------------------------
def tx_freq_grid(bits, prm: ModemParams, n_frames):
    """[n_frames*(M0+D), N] frequency grid.  MultiAntennaSystem.py:135-183, SDRScript.py:79-85.

    ``bits`` is a flat 0/1 array of n_frames*D*ndb*bps payload bits.  Synch symbol m of a frame
    carries chunk m of the length-MM ZC sequence.  DEVIATION: the reference computes
    the chunk from synch_state but never advances it (MultiAntennaSystem.py:143-146), so with
    M0 > 1 it would repeat chunk 0 while its own receiver correlates against chunk m; for M0 = 1,
    the only case any fixture pins, both readings coincide.
    """
    N = prm.nfft
    zc = zadoff_chu(prm.mm, prm.zc_root)            # TX side always keys parity on MM (SynchSignal.py:24)
    sb = bins_used(prm.num_synch_bins, N)
    db = bins_used(prm.num_data_bins, N)
    per_sym = prm.n_payload_bins * prm.bps
    bits = np.asarray(bits).reshape(-1)
    assert bits.size == n_frames * prm.d * per_sym
    ppos, dpos, pv = prm.pilot_pos, prm.data_pos, pilot_values(prm)
    grid = np.zeros((n_frames * prm.frame_syms, N), dtype=complex)
    data_i = 0
    for f in range(n_frames):
        for s in range(prm.frame_syms):
            row = f * prm.frame_syms + s
            if s < prm.m0:
                grid[row, sb] = zc[s * prm.num_synch_bins:(s + 1) * prm.num_synch_bins]
            else:
                chunk = bits[data_i * per_sym:(data_i + 1) * per_sym]
                grid[row, db[dpos]] = bits_to_symbols(chunk, prm.bps)
                grid[row, db[ppos]] = pv
                data_i += 1
    return grid


# --------------------------------------------------------------------------------------------
# a5  IFFT + CP add + power normalise
# --------------------------------------------------------------------------------------------
def tx_symb_gen(grid, cp_len, cps=None):
    """Per symbol: ifft -> [last CP samples, body] -> unit energy/sample -> divide by sqrt(var).

    MultiAntennaSystem.py:189-219 (single antenna).  Returns the flat time waveform.  ``cps`` (EXTENSION)
    gives every symbol its own prefix length; the two normalisations then use that symbol's length.
    """
    n_sym, N = grid.shape
    cps = [cp_len] * n_sym if cps is None else list(cps)
    out = np.zeros(n_sym * N + sum(cps), dtype=complex)
    at = 0
    for s in range(n_sym):
        L = N + cps[s]
        body = np.fft.ifft(grid[s], N)
        t = np.concatenate((body[N - cps[s]:], body))
        energy = abs(np.dot(t, np.conj(t)))
        scale = np.sqrt(L / energy) if energy > 1e-30 else 1.0
        t = t * scale
        t = t * (1.0 / np.sqrt(np.var(t)))
        out[at:at + L] = t
        at += L
    return out


def pilot_values(prm: ModemParams):
    """Unit-modulus pilot sequence: Zadoff-Chu of length n_pilots (EXTENSION, our definition)."""
    n = len(prm.pilot_pos)
    return zadoff_chu(n, prm.pilot_root) if n else np.zeros(0, dtype=complex)


def tx_modulate(bits, prm: ModemParams, n_frames):
    cps = [prm.sym_cp(k) for k in range(prm.frame_syms)] * n_frames if prm.cp_pattern else None
    return tx_symb_gen(tx_freq_grid(bits, prm, n_frames), prm.cp_len, cps)


# --------------------------------------------------------------------------------------------
# a7  channel + AWGN (test bench)
# --------------------------------------------------------------------------------------------
REF_TAPS = np.array([0.3977, 0.7954 - 0.3977j, -0.1988, 0.0994, -0.0398])  # MultiAntennaSystem.py:64


def reference_channel(nfft, kind="Fading"):
    """N-tap zero-padded impulse response.  MultiAntennaSystem.py:79-90 ('AWGN' = one-sample delay)."""
    h = np.zeros(nfft, dtype=complex)
    if kind == "AWGN":
        h[1] = 1.0
    elif kind == "Ideal":
        h[0] = 1.0
    else:
        h[0:len(REF_TAPS)] = REF_TAPS / np.linalg.norm(REF_TAPS)
    return h


def apply_channel(tx, h):
    """np.convolve(tx, h): MultiAntennaSystem.py:221-231."""
    return np.convolve(tx, h)


def noise_variance(tx, prm: ModemParams, snr_db, snr_type="Digital"):
    """MultiAntennaSystem.py:236-246."""
    sig_pow = np.var(tx)
    if snr_type == "Digital":
        return (1.0 / (prm.num_data_bins * prm.bps)) * prm.sym_len * sig_pow * 10 ** (-snr_db / 10.0)
    return sig_pow * 10 ** (-snr_db / 10.0)


def add_awgn(x, noise_var, rng):
    """Complex AWGN of total variance noise_var.  MultiAntennaSystem.py:257-260 (seeded here)."""
    s = np.sqrt(noise_var / 2.0)
    return x + s * rng.standard_normal(x.shape) + 1j * s * rng.standard_normal(x.shape)


# --------------------------------------------------------------------------------------------
# a7 (test bench), batched form used by the benchmarks: per-frame FIR + counter-based AWGN.
# The reference draws its noise from NumPy's global generator (MultiAntennaSystem.py:257-260, unseeded), which no
# other implementation can reproduce; the batched generator therefore uses Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
# "Parallel random numbers: as easy as 1, 2, 3", SC'11 -- the Random123 reference algorithm) keyed by (seed) and
# counted by the absolute sample index, so the CUDA generator and this restatement produce the same stream.
# --------------------------------------------------------------------------------------------
PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4], key: (k0, k1) -> uint32 [..., 4].  Ten rounds of
    (c0, c1, c2, c3) <- (hi(M1*c2) ^ c1 ^ k0, lo(M1*c2), hi(M0*c0) ^ c3 ^ k1, lo(M0*c0)), key bumped by the Weyl
    constants between rounds."""
    c = np.asarray(counter, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = np.uint64(int(key[0]) & 0xFFFFFFFF), np.uint64(int(key[1]) & 0xFFFFFFFF)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(PHILOX_M0) * c0
        p1 = np.uint64(PHILOX_M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & m32
        hi1, lo1 = p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(PHILOX_W0)) & m32
        k1 = (k1 + np.uint64(PHILOX_W1)) & m32
    return np.stack((c0, c1, c2, c3), axis=-1).astype(np.uint32)


def philox_awgn(n, sigma, seed, first=0):
    """Complex noise samples first..first+n-1 of the stream `seed`: per-component standard deviation sigma.
    Sample i uses counter (i mod 2^32, i >> 32, 0, 0) and key (seed mod 2^32, seed >> 32); its first two output
    words become u1, u2 = (float32(word) + 0.5) * 2^-32 (formed in float32 like the CUDA generator), and the sample is
    the Box-Muller pair sigma * sqrt(-2 ln u1) * (cos 2 pi u2 + j sin 2 pi u2), evaluated here in double precision."""
    i = np.arange(first, first + n, dtype=np.uint64)
    ctr = np.zeros((n, 4), dtype=np.uint64)
    ctr[:, 0] = i & np.uint64(0xFFFFFFFF)
    ctr[:, 1] = i >> np.uint64(32)
    r = philox4x32_10(ctr, (int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF))
    scale = np.float32(2.0 ** -32)
    u1 = ((r[:, 0].astype(np.float32) + np.float32(0.5)) * scale).astype(np.float64)
    u2 = ((r[:, 1].astype(np.float32) + np.float32(0.5)) * scale).astype(np.float64)
    rad = float(sigma) * np.sqrt(-2.0 * np.log(u1))
    return rad * (np.cos(TWO_PI * u2) + 1j * np.sin(TWO_PI * u2))


def channel_awgn_frames(tx, n_frames, frame_len, taps, sigma=0.0, seed=0):
    """Batched test-bench channel: every frame is filtered on its own -- y[f, n] = sum_t h[f, t] x[f, n - t], the
    np.convolve of MultiAntennaSystem.py:221-231 truncated to the frame, with nothing leaking in from the previous
    frame -- then sigma * philox noise is added (the `additive_noise` step, :257-260, with a reproducible stream).
    taps: [n_taps] shared by every frame or [n_frames, n_taps] one row per frame."""
    x = np.asarray(tx, dtype=complex).reshape(n_frames, frame_len)
    h = np.asarray(taps, dtype=complex)
    y = np.empty_like(x)
    for f in range(n_frames):
        hf = h[f] if h.ndim == 2 else h
        y[f] = np.convolve(x[f], hf)[:frame_len]
    y = y.reshape(-1)
    if sigma > 0:
        y = y + philox_awgn(y.size, sigma, seed)
    return y


# --------------------------------------------------------------------------------------------
# a8-a10  one synchronisation trial
# --------------------------------------------------------------------------------------------
def delay_bank(prm: ModemParams):
    """del_mat_exp[(CP+1) x M0*nsb] = exp(j*2pi*d*bin/N), tiled M0 times.  SynchAndChanEst.py:74-75."""
    sb = bins_used(prm.num_synch_bins, prm.nfft)
    e = np.exp((1j * (TWO_PI / prm.nfft)) * np.outer(np.arange(prm.cp_len + 1), sb))
    return np.tile(e, (1, prm.m0))


def cfo_bank(fo_range, fs, nfft):
    """cfo[F x N] = exp(j*2pi*(1/fs)*fo*n): one window-local de-rotation phasor per frequency-offset
    hypothesis.  GRR/LEGACY/gr-ofdm-rx/python/SynchEstAndFO.py:196."""
    return np.exp(1j * 2 * np.pi * (1 / fs) * (np.outer(np.asarray(fo_range, dtype=float), list(range(nfft)))))


def _window(in0, a, N, rot=None, nu=None):
    """N samples from ``a`` (zero-padded past the end, like np.fft.fft(x, N)), times the CFO phasor.

    ``rot`` is one row of cfo_bank(): the reference restarts the phasor at every window (SynchEstAndFO.py:254,
    :332), so windows L samples apart keep a relative phase of 2*pi*nu*L.  EXTENSION (our definition):
    with ``nu`` (the row's frequency in cycles per sample) the phasor is continued from the buffer origin,
    exp(j*2pi*nu*(a+n)), which is what undoing a real carrier offset needs."""
    w = np.zeros(N, dtype=complex)
    seg = np.asarray(in0[a:a + N], dtype=complex)
    w[:len(seg)] = seg
    if rot is not None:
        w = w * rot
    if nu is not None:
        w = w * np.exp(2j * np.pi * ((nu * a) % 1.0))
    return w


def sync_window(in0, start, prm: ModemParams, rot=None, nu=None):
    """Gather the M0 FFT windows that start at ``start`` and pick the synch bins.

    SynchAndChanEst.py:145-154; with ``rot`` each window is first multiplied by one row of the CFO
    bank (SynchEstAndFO.py:252-256).  Returns the un-normalised M0*nsb vector.
    """
    N, L = prm.nfft, prm.sym_len
    sb = bins_used(prm.num_synch_bins, N)
    out = np.zeros(prm.mm, dtype=complex)
    for ll in range(prm.m0):
        a = L * ll + start
        spec = np.fft.fft(_window(in0, a, N, rot, nu), N)
        out[ll * prm.num_synch_bins:(ll + 1) * prm.num_synch_bins] = spec[sb]
    return out


def sync_trial(in0, start, prm: ModemParams, zc=None, bank=None, literal=False, rot=None, nu=None):
    """One trial of the sliding search at FFT-window start sample ``start``.

    Returns (dmax_ind, dmax_val, synchdat) where synchdat is the unit-power synch-bin vector.
    SynchAndChanEst.py:156-164 / synch_and_chan_est.py:158-168.
    """
    zc = zc_for(prm) if zc is None else zc
    bank = delay_bank(prm) if bank is None else bank
    sd0 = sync_window(in0, start, prm, rot, nu)
    if literal:
        p_est = np.sqrt(len(sd0) / sum(np.multiply(sd0, np.conj(sd0))))
        sd = p_est * sd0
        del_mat = np.matmul(np.matmul(bank, np.diag(sd)), np.conj(zc))
    else:
        p_est = np.sqrt(len(sd0) / np.sum(sd0 * np.conj(sd0)))
        sd = p_est * sd0
        del_mat = bank @ (sd * np.conj(zc))
    mag = np.abs(del_mat)
    return int(np.argmax(mag)), float(np.max(mag)), sd


def sync_search_time_domain(in0, first_start, stride, n_trials, prm: ModemParams, nus=None, continuous=False):
    """Every trial of the sliding search (sync_trial at first_start + t*stride) WITHOUT a transform per trial.

    Derivation (M0 windows s = 0..M0-1 at p + s*L, synch bins S, y_k = X_k conj(zc_k)):
        c[d; p] / p_est = sum_k X_k conj(zc_k) e^{+j2pi dk/N} = sum_n x[p+n] g[(d - n) mod N],
        g[m] = sum_{k in S} conj(zc_k) e^{+j2pi km/N}          (N times the inverse DFT of the scattered conj(ZC)).
    With h[m] = g[(-m) mod N] and the linear sliding correlation r(q) = sum_m x[q+m] h[m]:
        c[d; p] / p_est = r(p+d) + sum_{i=1..d} (x[p+d-i] - x[p+d-i+N]) h[N-i]
    i.e. for every position q = p+d the CP+1 delays are r(q) plus a PREFIX SUM over i.  The unit-power factor needs
    sum_{k in S} |X_k|^2 = N sum_n |x[p+n]|^2 - sum_{k not in S} |X_k|^2, the second term from a direct DFT of the
    few unused bins (DC and the band edge).  Same numbers as sync_trial up to rounding; this is the specification of
    the fast acquisition path, not what the reference executes.
    Returns (peak[n_trials], delay[n_trials]).  Windows past the end are zero-padded like np.fft.fft.

    With ``nus`` (cycles per sample of each row of cfo_bank(), SynchEstAndFO.py:196,:247-277) every hypothesis is
    evaluated and the first maximum of the peak wins; a third array, hyp[n_trials], is returned.  The window-local
    phasor e^{j2pi nu n} multiplies sample n of each window, so with u = e^{j2pi nu}
        c[d; p] / p_est = u^d * sum_s phi_s [ r_s(q_s) + sum_{i=1..d} ( x[q_s-i] h_s[N-i] u^{-i} - x[q_s-i+N] h_s[N-i] u^{N-i} ) ],
        r_s(q) = sum_m x[q+m] h_s[m] u^m,   q_s = p + s*L + d,
    (|u^d| = 1 drops out of the magnitude; phi_s = 1, or u^{s*L} up to a common factor when the phasor is continued
    from the buffer origin, ``continuous``), and the unused bins DC / band edge become the sliding sums
    sum_n x[p+n] (+-u)^n.
    """
    if nus is not None:
        return _sync_search_time_domain_cfo(in0, first_start, stride, n_trials, prm, nus, continuous)
    N, CP, L, M0, nsb = prm.nfft, prm.cp_len, prm.sym_len, prm.m0, prm.num_synch_bins
    assert nsb < N, "the +-N/2 alias case (num_synch_bins == nfft) keeps the transform path"
    sb = bins_used(nsb, N)
    zc = zc_for(prm)
    last = first_start + (n_trials - 1) * stride
    need = last + CP + (M0 - 1) * L + 2 * N + 1
    x = np.zeros(max(need, len(in0)), dtype=complex)
    x[:len(in0)] = np.asarray(in0, dtype=complex)
    n_q = last - first_start + CP + 1                                  # positions q = first_start .. last + CP
    num = np.zeros((n_q, CP + 1), dtype=complex)                        # num[q - first_start, d] = c[d; q-d] / p_est
    unused = np.setdiff1d(np.arange(N), sb)
    pw = np.zeros(last - first_start + 1)                                # sum over windows of sum_{k in S} |X_k|^2 at p
    nn = np.arange(N)
    for s in range(M0):
        Z = np.zeros(N, dtype=complex)
        Z[sb] = np.conj(zc[s * nsb:(s + 1) * nsb])
        g = N * np.fft.ifft(Z)
        h = g[(-nn) % N]
        xs = x[s * L:]
        # r(q) for q = first_start .. last + CP
        seg = xs[first_start:first_start + n_q + N - 1]
        r = np.correlate(seg, np.conj(h), mode="valid")                  # sum_m seg[q+m] * h[m]
        for qi in range(n_q):
            q = first_start + qi
            acc = r[qi]
            num[qi, 0] += acc
            for i in range(1, CP + 1):
                a = q - i
                dl = (xs[a] - xs[a + N]) if a >= 0 else (0.0 - xs[a + N] if a + N >= 0 else 0.0)
                acc = acc + dl * h[N - i]
                num[qi, i] += acc
        # window energy and the unused bins at every p
        e = np.abs(xs[first_start:first_start + (last - first_start) + N]) ** 2
        ce = np.concatenate(([0.0], np.cumsum(e)))
        energy = ce[N:N + last - first_start + 1] - ce[:last - first_start + 1]
        pw += N * energy
        for k in unused:
            ph = np.exp(-2j * np.pi * k * nn / N)
            for pi in range(last - first_start + 1):
                pw[pi] -= abs(np.dot(xs[first_start + pi:first_start + pi + N], ph)) ** 2
    peak = np.zeros(n_trials)
    delay = np.zeros(n_trials, dtype=np.int32)
    for t in range(n_trials):
        pi = t * stride
        mags = np.array([abs(num[pi + d, d]) for d in range(CP + 1)])
        p_est = np.sqrt(prm.mm / pw[pi]) if pw[pi] > 0 else 0.0
        delay[t] = int(np.argmax(mags))
        peak[t] = p_est * mags[delay[t]]
    return peak, delay


def _sync_search_time_domain_cfo(in0, first_start, stride, n_trials, prm: ModemParams, nus, continuous):
    """The hypothesis-bank form of sync_search_time_domain (see its docstring for the derivation)."""
    N, CP, L, M0, nsb = prm.nfft, prm.cp_len, prm.sym_len, prm.m0, prm.num_synch_bins
    assert nsb < N
    sb = bins_used(nsb, N)
    zc = zc_for(prm)
    last = first_start + (n_trials - 1) * stride
    span = last - first_start + 1
    need = last + CP + (M0 - 1) * L + 2 * N + 1
    x = np.zeros(max(need, len(in0)), dtype=complex)
    x[:len(in0)] = np.asarray(in0, dtype=complex)
    n_q = span + CP
    nn = np.arange(N)
    unused = np.setdiff1d(np.arange(N), sb)
    best_pk = np.full(n_trials, -1.0)
    best_dl = np.zeros(n_trials, dtype=np.int32)
    best_h = np.zeros(n_trials, dtype=np.int32)
    for hi, nu in enumerate(nus):
        u = np.exp(2j * np.pi * nu)
        num = np.zeros((n_q, CP + 1), dtype=complex)
        pw = np.zeros(span)
        for s in range(M0):
            Z = np.zeros(N, dtype=complex)
            Z[sb] = np.conj(zc[s * nsb:(s + 1) * nsb])
            g = N * np.fft.ifft(Z)
            h = g[(-nn) % N]
            phi = np.exp(2j * np.pi * ((nu * s * L) % 1.0)) if continuous else 1.0
            hm = phi * h * u ** nn                                       # main taps h_s[m] u^m
            hp = np.array([0.0] + [phi * h[N - i] * u ** (-i) for i in range(1, CP + 1)])     # prefix taps, index i
            xs = x[s * L:]
            seg = xs[first_start:first_start + n_q + N - 1]
            r = np.correlate(seg, np.conj(hm), mode="valid")
            for qi in range(n_q):
                q = first_start + qi
                acc = r[qi]
                num[qi, 0] += acc
                for i in range(1, CP + 1):
                    a = q - i
                    lo = xs[a] if a >= 0 else 0.0
                    hi_ = xs[a + N] if a + N >= 0 else 0.0
                    acc = acc + lo * hp[i] - hi_ * hm[N - i]
                    num[qi, i] += acc
            e = np.abs(xs[first_start:first_start + span - 1 + N]) ** 2
            ce = np.concatenate(([0.0], np.cumsum(e)))
            pw += N * (ce[N:N + span] - ce[:span])
            for k in unused:
                ph = np.exp(-2j * np.pi * k * nn / N) * u ** nn
                for pi in range(span):
                    pw[pi] -= abs(np.dot(xs[first_start + pi:first_start + pi + N], ph)) ** 2
        for t in range(n_trials):
            pi = t * stride
            mags = np.array([abs(num[pi + d, d]) for d in range(CP + 1)])
            p_est = np.sqrt(prm.mm / pw[pi]) if pw[pi] > 0 else 0.0
            d_i = int(np.argmax(mags))
            pk = p_est * mags[d_i]
            if pk > best_pk[t]:                                          # first maximum over the hypotheses
                best_pk[t], best_dl[t], best_h[t] = pk, d_i, hi
    return best_pk, best_dl, best_h


def sync_trial_cfo(in0, start, prm: ModemParams, cfo, zc=None, bank=None, literal=False, nus=None):
    """The trial repeated for every row of the CFO bank; the hypothesis with the largest peak wins
    (first one on ties, np.argmax).  SynchEstAndFO.py:247-277.

    Returns (fo_ind, dmax_ind, dmax_val, per-hypothesis list of (dmax_ind, dmax_val, synchdat)).
    """
    res = [sync_trial(in0, start, prm, zc, bank, literal, rot=cfo[h], nu=None if nus is None else nus[h])
           for h in range(len(cfo))]
    vals = np.array([r[1] for r in res])
    fo_ind = int(np.argmax(vals))
    return fo_ind, res[fo_ind][0], float(np.max(vals)), res


# --------------------------------------------------------------------------------------------
# a12/a13  LS channel estimate on the preamble
# --------------------------------------------------------------------------------------------
def chan_est(sd, dmax_ind, prm: ModemParams, zc=None, bank=None, literal=False):
    """De-rotate by the detected delay, divide out the ZC, average over the M0 repeats.

    Returns (chan_est[nsb], chan_est1[N], data_recov[MM]).  SynchAndChanEst.py:177-188.
    """
    zc = zc_for(prm) if zc is None else zc
    bank = delay_bank(prm) if bank is None else bank
    del_vec = bank[dmax_ind]
    if literal:
        data_recov = np.matmul(np.diag(del_vec), sd)
        tmp = np.divide(np.matmul(np.diag(data_recov), np.conj(zc)), (1.0 / prm.reg_chanest) + np.ones(len(zc)))
    else:
        data_recov = del_vec * sd
        tmp = data_recov * np.conj(zc) / (1.0 / prm.reg_chanest + 1.0)
    ce = tmp.reshape(prm.m0, prm.num_synch_bins).sum(axis=0) / float(prm.m0)
    ce1 = np.zeros(prm.nfft, dtype=complex)
    ce1[bins_used(prm.num_synch_bins, prm.nfft)] = ce
    return ce, ce1, data_recov


def chan_est_aux(ce, ce1, data_recov, prm: ModemParams):
    """Time-domain channel and equalised synch symbol.  SynchAndChanEst.py:202,:212-218.

    Returns (chan_est_tim[N], est_synch_freq[MM]).
    """
    tim = np.fft.ifft(ce1, prm.nfft)
    g = np.conj(ce) / (1.0 / prm.reg_eq_synch + ce * np.conj(ce))
    return tim, np.tile(g, prm.m0) * data_recov


# --------------------------------------------------------------------------------------------
# a14-a16  one data symbol
# --------------------------------------------------------------------------------------------
def demod_symbol(in0, ptr, delay, h_full, prm: ModemParams, literal=False, rot=None, nu=None):
    """FFT (CP skipped by the start pointer) -> data bins -> unit power -> de-rotate -> one-tap MMSE.

    SynchAndChanEst.py:226-248 / synch_and_chan_est.py:230-250.  A window that runs past the end
    of ``in0`` is zero-padded by np.fft.fft exactly like the reference.
    """
    N = prm.nfft
    db = bins_used(prm.num_data_bins, N)
    spec = np.fft.fft(_window(in0, ptr, N, rot, nu), N)    # rot: SynchEstAndFO.py:331-333
    fd = spec[db]
    p0 = np.sqrt(len(fd) / np.dot(fd, np.conj(fd)))
    fd = fd * p0
    hd = h_full[db]
    inv_reg = 0.0 if np.isinf(prm.reg_eq) else 1.0 / prm.reg_eq
    if prm.pilot_spacing:
        # EXTENSION (our definition): per-symbol LS estimate on the comb pilots, linear interpolation
        # across the used-bin list (held constant outside the outermost pilots), then the same one-tap
        # equaliser on the payload bins.  Timing (delay) still comes from the preamble.
        z = fd * np.exp(((1j * (TWO_PI / N)) * delay) * db)
        ppos, dpos = prm.pilot_pos, prm.data_pos
        hp = z[ppos] * np.conj(pilot_values(prm))
        jo = np.arange(prm.num_data_bins)
        hi = np.interp(jo, ppos, hp.real) + 1j * np.interp(jo, ppos, hp.imag)
        g = np.conj(hi) / (inv_reg + hi * np.conj(hi))
        return (g * z)[dpos]
    if literal:
        arg = [((1j * (TWO_PI / N)) * delay) * kk for kk in db]
        z = np.matmul(np.diag(fd), np.exp(arg))
        mag = np.matmul(np.diag(hd), np.conj(hd))
        g = np.divide(np.conj(hd), [inv_reg + v for v in mag])
        return np.matmul(np.diag(g), z)
    z = fd * np.exp(((1j * (TWO_PI / N)) * delay) * db)
    g = np.conj(hd) / (inv_reg + hd * np.conj(hd))
    return g * z


# --------------------------------------------------------------------------------------------
# framed receiver: one trial per frame at the nominal position (throughput mode)
# --------------------------------------------------------------------------------------------
def rx_frames(rx, n_frames, prm: ModemParams, frame_stride=None, first=0, literal=False, cfo=None, nus=None):
    """Loop-A body once per frame + loop-B body for all D data symbols of that frame.

    Frame f's synch FFT window opens at  first + f*frame_stride + CP - sync_advance ; data symbol n
    opens at that position + M0*L + n*L (with a CP pattern: at the end of its own prefix, less sync_advance).  This is the reference's per-lock arithmetic
    (SynchAndChanEst.py:143-219 then :221-248) applied with one candidate per frame, which is how
    a receiver with known framing runs it.  Returns dict(delay, peak, locked, h[N], eq[D,ndb]) per
    frame stacked on axis 0.  With a CFO bank (``cfo``, rows from cfo_bank()) every frame picks its own
    hypothesis on the preamble (SynchEstAndFO.py:247-277) and de-rotates its data windows with it
    (:331-333); ``hyp`` is added to the result.  ``nus`` (cycles per sample of each row) switches to the
    phase-continuous extension described at _window().
    """
    fs = prm.frame_len if frame_stride is None else frame_stride
    zc, bank = zc_for(prm), delay_bank(prm)
    assert not prm.cp_pattern or prm.m0 == 1, "the CP pattern is defined for one synch symbol per frame"
    # window of symbol k relative to the frame's synch window (uniform CP: k*L; with a CP pattern the prefixes differ)
    woff = [prm.sym_start(k) + prm.sym_cp(k) - prm.sym_cp(0) for k in range(prm.frame_syms)]
    delay = np.zeros(n_frames, dtype=np.int32)
    peak = np.zeros(n_frames)
    locked = np.zeros(n_frames, dtype=np.uint8)
    h = np.zeros((n_frames, prm.nfft), dtype=complex)
    eq = np.zeros((n_frames, prm.d, prm.n_payload_bins), dtype=complex)
    hyp = np.zeros(n_frames, dtype=np.int32)
    for f in range(n_frames):
        pos = first + f * fs + prm.sym_cp(0) - prm.sync_advance
        rot = nu = None
        if cfo is None:
            d_i, d_v, sd = sync_trial(rx, pos, prm, zc, bank, literal)
        else:
            hyp[f], d_i, d_v, res = sync_trial_cfo(rx, pos, prm, cfo, zc, bank, literal, nus)
            sd = res[hyp[f]][2]
            rot = cfo[hyp[f]]
            nu = None if nus is None else nus[hyp[f]]
        delay[f], peak[f] = d_i, d_v
        locked[f] = 1 if d_v > prm.gate * prm.mm else 0
        _, h[f], _ = chan_est(sd, d_i, prm, zc, bank, literal)
        for n in range(prm.d):
            eq[f, n] = demod_symbol(rx, pos + woff[prm.m0 + n], d_i, h[f], prm, literal, rot=rot, nu=nu)
    out = dict(delay=delay, peak=peak, locked=locked, h=h, eq=eq)
    if cfo is not None:
        out["hyp"] = hyp
    return out


# --------------------------------------------------------------------------------------------
# DSSS despreading of the equalised data bins (legacy receiver)
# --------------------------------------------------------------------------------------------
def dsss_code(sf, root=37):
    """Spreading code SC: a Zadoff-Chu sequence of length ``sf`` (k^2 for even, k(k+1) for odd lengths).
    GRR/LEGACY/gr-ofdm-rx/python/SynchEstFOAndDSSS.py:253-261 (root = self.p = 37, :195)."""
    k = np.arange(sf)
    xx = k * k if sf % 2 == 0 else k * (k + 1)
    return np.exp((-1j * (TWO_PI / sf) * root / 2.0) * xx)


def dsss_despread(rows, code):
    """rows[..., ndb] -> [..., ndb // sf]: each group of sf consecutive data bins times conj(code), averaged.
    SynchEstFOAndDSSS.py:392-397 (bins past the last whole group are ignored, :358)."""
    rows = np.asarray(rows)
    sf = len(code)
    ng = rows.shape[-1] // sf
    g = rows[..., :ng * sf].reshape(rows.shape[:-1] + (ng, sf))
    return np.mean(g * np.conj(code), axis=-1)


def dsss_spread(symbols, code):
    """Transmit-side counterpart (our definition; the spreading transmitter is not in the reference tree): every
    symbol occupies sf consecutive data bins, multiplied by the code chips."""
    return (np.asarray(symbols)[..., None] * code).reshape(np.asarray(symbols).shape[:-1] + (-1,))


# --------------------------------------------------------------------------------------------
# a18  demapping
# --------------------------------------------------------------------------------------------
def bit_recovery(data_in):
    """The legacy QPSK soft/hard demapper.  GRR/LEGACY/gr-ofdm-rx/python/BitRecovery.py:77-157.

    Returns (llrp0[2n], llrp1[2n], hardbit[2n]).  sigma is estimated over the WHOLE buffer (:102).
    """
    x = np.asarray(data_in, dtype=complex).reshape(-1)
    pts = np.exp((1j * (TWO_PI / 8.0)) * _QPSK_EIGHTHS)            # :45-52
    dist = np.abs(x[:, None] - pts[None, :])                       # :86-88
    near = np.argmin(dist, axis=1)
    dmin = dist[np.arange(len(x)), near]
    ez = x - pts[near]                                             # :93-98
    sigma = 0.7071067811865476 * np.mean(np.abs(dmin))            # :102
    dfact = 1.0 / (sigma * sigma)
    K = 1.414213562373095                                          # :57
    n = len(x)
    llr0 = np.zeros(2 * n)
    llr1 = np.zeros(2 * n)
    same_r = -0.5 * dfact * np.abs(ez.real)
    other_r = -0.5 * dfact * (K - np.abs(ez.real))
    same_i = -0.5 * dfact * np.abs(ez.imag)
    other_i = -0.5 * dfact * (K - np.abs(ez.imag))
    re, im = x.real, x.imag
    # quadrant tests in the reference's order, first match wins (:106-125)
    q1 = (re >= 0) & (im >= 0)
    q2 = ~q1 & (re <= 0) & (im >= 0)
    q3 = ~q1 & ~q2 & (re <= 0) & (im <= 0)
    q4 = ~q1 & ~q2 & ~q3 & (re >= 0) & (im <= 0)
    re_pos = q1 | q4          # Re bit is 0 -> llrp0 gets the "same" metric
    im_pos = q1 | q2
    any_q = q1 | q2 | q3 | q4  # (NaN inputs match nothing and stay 0)
    llr0[0::2] = np.where(any_q, np.where(re_pos, same_r, other_r), 0.0)
    llr1[0::2] = np.where(any_q, np.where(re_pos, other_r, same_r), 0.0)
    llr0[1::2] = np.where(any_q, np.where(im_pos, same_i, other_i), 0.0)
    llr1[1::2] = np.where(any_q, np.where(im_pos, other_i, same_i), 0.0)
    hard = (0.5 * (np.sign(llr1 - llr0) + 1.0)).astype(int)        # :155-156
    return llr0, llr1, hard


def maxlog_llr(x, bps, scale=1.0):
    """Exact max-log LLR per bit, brute force over the whole constellation (OUR DEFINITION).

    L_b = scale * ( min_{s: b=0} |x-s|^2 - min_{s: b=1} |x-s|^2 ),  L_b > 0  <=>  bit 1.
    ``scale`` = 1/noise_variance.  Bit order is the mapper's (MSB first).  Returns [..., n*bps].
    """
    x = np.asarray(x, dtype=complex)
    pts = constellation(bps)
    d2 = np.abs(x[..., None] - pts) ** 2
    out = np.zeros(x.shape + (bps,))
    idx = np.arange(1 << bps)
    for b in range(bps):
        one = ((idx >> (bps - 1 - b)) & 1).astype(bool)
        out[..., b] = scale * (d2[..., ~one].min(-1) - d2[..., one].min(-1))
    return out.reshape(x.shape[:-1] + (x.shape[-1] * bps,))


def hard_bits(x, bps):
    """Bits of the nearest constellation point (== sign of the max-log LLR).  [..., n*bps] uint8."""
    x = np.asarray(x, dtype=complex)
    pts = constellation(bps)
    near = np.argmin(np.abs(x[..., None] - pts), axis=-1)
    sh = np.arange(bps - 1, -1, -1)
    b = ((near[..., None] >> sh) & 1).astype(np.uint8)
    return b.reshape(x.shape[:-1] + (x.shape[-1] * bps,))


# --------------------------------------------------------------------------------------------
# counters (BER / EVM) -- the only quantity that is reduced across GPUs
# --------------------------------------------------------------------------------------------
def counters(bits_rx, bits_tx, eq, ref_syms):
    """[bit_errors, bits, sum|e|^2, sum|ref|^2] as float64."""
    bits_rx = np.asarray(bits_rx).reshape(-1)
    bits_tx = np.asarray(bits_tx).reshape(-1)
    e = np.asarray(eq).reshape(-1) - np.asarray(ref_syms).reshape(-1)
    return np.array([float(np.count_nonzero(bits_rx != bits_tx)), float(bits_tx.size),
                     float(np.sum(np.abs(e) ** 2)), float(np.sum(np.abs(ref_syms) ** 2))])
