"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on
the same seeded inputs, against the committed reference fixtures and the reference's golden
vectors, plus size-independent properties at larger sizes.

Tolerances (BASELINE.json north_star): per-channel relative RMS error <= 1e-5 and
max-abs error <= 1e-4 normalised to sqrt(input power), for the float32 channelizer outputs."""
import os

import numpy as np
import pytest

from conftest import TOL_MAX_ABS, TOL_REL_RMS, cplx, max_abs_norm, rel_rms_per_channel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import liquid_dsp_b200
    assert liquid_dsp_b200.device_count() > 0, "GPU tests need a CUDA device"
    return liquid_dsp_b200


def check_channels(y, want, x):
    assert rel_rms_per_channel(y, want).max() < TOL_REL_RMS
    assert max_abs_norm(y, want, x) < TOL_MAX_ABS


# ------------------------------------------------------------------ firpfbch2 analysis (headline path)
@pytest.mark.parametrize("M,m,nfr", [(8, 5, 80), (16, 5, 80), (32, 5, 64), (64, 5, 64), (6, 2, 50), (12, 3, 41),
                                     (256, 4, 48), (1024, 7, 96), (4096, 9, 40), (2048, 3, 24), (128, 1, 33)])
def test_pfbch2_analysis_vs_oracle(L, orc, M, m, nfr):
    x = orc.synth_input(nfr * M // 2, seed=0xA0 + M)
    want = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(x)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    got = q.execute_block(x)
    check_channels(got, want, x)


@pytest.mark.parametrize("M,m", [(8, 5), (16, 5), (64, 3), (6, 2), (256, 4)])
def test_pfbch2_vs_reference_fixtures(L, refout, M, m):
    x = refout["pfbch2_a_M%d_m%d_x" % (M, m)]
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    check_channels(q.execute_block(x), refout["pfbch2_a_M%d_m%d_y" % (M, m)], x)
    X = refout["pfbch2_s_M%d_m%d_x" % (M, m)]
    want = refout["pfbch2_s_M%d_m%d_y" % (M, m)]
    qs = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    got = qs.execute_block(X)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5


def test_pfbch2_survey_fingerprints(L, refout):
    x = refout["pfbch2_a_M16_m5_lcg_x"]
    y = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, 16, 5, 60.0).execute_block(x)
    assert abs(y[0, 0] - (-7.94109947e-05)) < 1e-8
    assert abs(y[21, 1] - (1.93961821e-02 - 8.78997445e-02j)) < 1e-6
    assert abs(y[79, 2] - (-7.58994371e-02 - 8.95354375e-02j)) < 1e-6


def test_pfbch2_single_frame_execute_equals_block_and_state_carry(L, orc):
    """k x execute == execute_block(k), bit for bit, across odd splits (flag and history carried)"""
    M, m = 32, 4
    x = orc.synth_input(37 * M // 2, seed=5)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    whole = q.execute_block(x)
    q.reset()
    one = np.stack([q.execute(x[t * M // 2:(t + 1) * M // 2]) for t in range(37)])
    assert np.array_equal(whole, one)
    q.reset()
    parts = np.concatenate([q.execute_block(x[:5 * M // 2]), q.execute_block(x[5 * M // 2:18 * M // 2]),
                            q.execute_block(x[18 * M // 2:])])
    assert np.array_equal(whole, parts)


def test_pfbch2_get_set_state_resumes(L, orc):
    M, m = 64, 3
    x = orc.synth_input(40 * M // 2, seed=6)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    whole = q.execute_block(x)
    q.reset()
    a = q.execute_block(x[:13 * M // 2])
    hist, flag = q.get_state()
    assert flag == 1 and hist.size == q.get_state_len()
    q2 = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    q2.set_state(hist, flag)
    b = q2.execute_block(x[13 * M // 2:])
    assert np.array_equal(whole, np.concatenate([a, b]))


@pytest.mark.parametrize("M", [8, 16, 32, 64])
def test_pfbch2_autotest_roundtrip(L, M):
    """firpfbch2_crcf_autotest.c:27-99 run on the GPU objects: analysis -> synthesis reconstructs
    the input delayed by 2Mm - M/2 + 1 within 1e-3, interleaving single-frame execute calls"""
    m, As = 5, 60.0
    n = M * 8 * m
    s, x = 1, np.empty(n, np.complex64)
    for i in range(n):
        s = (s * 524287) % 1031
        x[i] = np.float32(s) / np.float32(1031) - np.float32(0.5)
    qa = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, As)
    qs = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, As)
    y = np.empty(n, np.complex64)
    for i in range(0, n, M // 2):
        Y = qa.execute(x[i:i + M // 2])
        y[i:i + M // 2] = qs.execute(Y)
    delay = 2 * M * m - M // 2 + 1
    assert np.abs(y[:delay]).max() < 1e-3
    assert np.abs(y[delay:] - x[:n - delay]).max() < 1e-3


@pytest.mark.parametrize("M,m,nfr", [(16, 3, 60), (64, 2, 40), (10, 2, 33), (256, 4, 30)])
def test_pfbch2_synthesis_vs_oracle(L, orc, M, m, nfr):
    X = orc.synth_input(nfr * M, seed=0xB0 + M)
    want = orc.Firpfbch2(orc.SYNTHESIZER, M, m, As=60.0).execute_frames(X)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    got = np.concatenate([q.execute_block(X[:7 * M]), q.execute_block(X[7 * M:])])
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5


@pytest.mark.parametrize("M,m,nfr,cut", [(256, 2, 333, 65), (256, 4, 700, 128), (1024, 5, 301, 77),
                                          (1024, 7, 520, 260), (1024, 9, 200, 100)])
def test_pfbch2_synthesis_register_window_vs_oracle(L, orc, M, m, nfr, cut):
    """long blocks take the register-window synthesis FIR (pfb_kernels.cu, 4m in {8,16,20,28}; m=9 stays on
    the generic kernel): odd and even starting parity, state carried across blocks, ragged chunk tail"""
    X = orc.synth_input(nfr * M, seed=0xC0 + M + m)
    want = orc.Firpfbch2(orc.SYNTHESIZER, M, m, As=60.0).execute_frames(X)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    got = np.concatenate([q.execute_block(X[:cut * M]), q.execute_block(X[cut * M:])])
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5
    q.reset()
    one = q.execute_block(X)
    assert np.abs(one - got).max() / np.abs(want).max() < 2e-6


def test_pfbch2_large_block_linearity_and_time_invariance(L, orc):
    """size-independent properties at a large size (2^22 samples, M=1024, m=7): linearity, and
    invariance to a shift by a whole number of frame pairs once the history is flushed"""
    M, m = 1024, 7
    n = 1 << 22
    a = orc.synth_input(n, seed=21)
    b = orc.synth_input(n, seed=22)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    ya = q.execute_block(a); q.reset()
    yb = q.execute_block(b); q.reset()
    yab = q.execute_block((a + 2 * b).astype(np.complex64)); q.reset()
    assert np.abs(yab - (ya + 2 * yb)).max() < 2e-6 * 3
    shift = 2 * 10                       # frames (even: same parity)
    ys = q.execute_block(a[shift * M // 2:])
    skip = 4 * m                          # frames needed to flush the zero history
    assert np.abs(ys[skip:] - ya[shift + skip:]).max() < 1e-6
    # a segment deep inside the stream against the oracle started 2mM samples earlier (SURVEY 8c trick)
    f0 = 5000
    start = f0 * M // 2 - 2 * m * M
    seg = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(a[start:start + (4 * m + 64) * M // 2])
    check_channels(ya[f0:f0 + 64], seg[4 * m:4 * m + 64], a[:4096])


def test_pfbch2_segment_form_equals_stateful(L, orc):
    """the stateless halo form used by the multi-GPU time sharding reproduces the stateful stream"""
    import ctypes as C
    M, m = 256, 4
    nfr = 64
    x = orc.synth_input(nfr * M // 2, seed=31)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    whole = q.execute_block(x)
    H = q.get_state_len()
    cut = 22 * M // 2                     # segment boundary (even frame), inside the stream
    xin = np.concatenate([np.zeros(H, np.complex64), x]).astype(np.complex64)
    d_in = L.lib.liquid_cuda_malloc(xin.nbytes)
    d_out = L.lib.liquid_cuda_malloc((nfr - 22) * M * 8)
    L.lib.liquid_cuda_memcpy_h2d(d_in, xin.ctypes.data_as(C.c_void_p), xin.nbytes)
    d_x = d_in + 8 * (H + cut)
    q.execute_segment_dev(d_x - 8 * H, d_x, 22, nfr - 22, d_out, 0)
    L.lib.liquid_cuda_sync()
    got = np.empty(((nfr - 22), M), np.complex64)
    L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_out, got.nbytes)
    L.lib.liquid_cuda_free(d_in); L.lib.liquid_cuda_free(d_out)
    assert np.array_equal(got, whole[22:])


# ------------------------------------------------------------------ BASELINE full sizes
@pytest.mark.parametrize("M,m,log2n", [(1024, 7, 30), (4096, 9, 30)])
def test_pfbch2_full_size_spot_checks_vs_oracle(L, orc, M, m, log2n):
    """BASELINE configs 3 and 5 at the size one GPU runs in bench.py (2^30 input samples, 8 GiB in,
    16 GiB out, device-resident): blocks of 48 frames at the stream start, around every kind of internal
    boundary of the fused launch (the first chunk boundaries between CTAs / clusters, the middle one, the
    last one), at pseudo-random places and at the very end are pulled back and checked against the oracle,
    which is fed from 2mM samples before each block (the state is only the last 2mM - M/2 samples,
    SURVEY.md 8c)."""
    import ctypes as C
    n = 1 << log2n
    M2 = M // 2
    n_frames = n // M2
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    H = q.get_state_len()
    d_in = L.lib.liquid_cuda_malloc(8 * (H + n))
    d_out = L.lib.liquid_cuda_malloc(8 * n_frames * M)
    assert d_in and d_out
    zeros = np.zeros(H, np.complex64)
    L.lib.liquid_cuda_memcpy_h2d(d_in, zeros.ctypes.data_as(C.c_void_p), zeros.nbytes)
    seed = 0x5EED0000 + M
    L.synth_input_dev(d_in + 8 * H, n, seed=seed)
    q.execute_segment_dev(d_in, d_in + 8 * H, 0, n_frames, d_out, 0)
    L.lib.liquid_cuda_sync()
    chunk = int(L.lib.liquid_cuda_last_chunk_frames())
    assert chunk > 0, "the fused kernel did not run"
    n_chunks = -(-n_frames // chunk)
    nb = 48
    starts = {0, n_frames - nb}
    for c in (1, 2, n_chunks // 2, n_chunks - 1):
        if 0 < c < n_chunks:
            starts.add(c * chunk - nb // 2)              # straddles the boundary between chunks c-1 and c
    rng = np.random.default_rng(M)
    starts.update(int(v) & ~1 for v in rng.integers(4 * m, n_frames - nb, size=6))     # even: frame parity 0
    oracle = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0)
    got = np.empty((nb, M), np.complex64)
    for f0 in sorted(starts):
        L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_out + 8 * f0 * M, got.nbytes)
        assert f0 % 2 == 0 and (f0 == 0 or f0 >= 4 * m)
        lead = min(4 * m, f0)                            # frames of run-in for the oracle (4m frames = 2mM samples)
        x = orc.synth_input((lead + nb) * M2, seed=seed, first=(f0 - lead) * M2)
        oracle.reset()
        want = oracle.execute_frames(x)[lead:]
        check_channels(got, want, x)
    L.lib.liquid_cuda_free(d_in)
    L.lib.liquid_cuda_free(d_out)


# ------------------------------------------------------------------ firpfbch
@pytest.mark.parametrize("M,m,nfr", [(4, 3, 40), (16, 4, 64), (10, 2, 33), (64, 5, 40), (256, 5, 48), (1024, 2, 20)])
def test_pfbch_analysis_and_synthesis_vs_oracle(L, orc, M, m, nfr):
    x = orc.synth_input(nfr * M, seed=0xC0 + M)
    for typ_o, typ_l in ((orc.ANALYZER, L.LIQUID_ANALYZER), (orc.SYNTHESIZER, L.LIQUID_SYNTHESIZER)):
        want = orc.Firpfbch(typ_o, M, kaiser=(m, 60.0)).execute_frames(x)
        q = L.firpfbch_crcf.create_kaiser(typ_l, M, m, 60.0)
        got = np.concatenate([q.execute_block(x[:9 * M]), q.execute_block(x[9 * M:])])
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-5
        if typ_o == orc.ANALYZER:
            assert rel_rms_per_channel(got, want).max() < TOL_REL_RMS


@pytest.mark.parametrize("M,m,nfr", [(64, 2, 1500), (100, 3, 900), (1024, 2, 700), (512, 5, 600), (2048, 4, 530)])
def test_pfbch_synthesis_register_window_vs_oracle(L, orc, monkeypatch, M, m, nfr):
    """long blocks take the register-window synthesis FIR (pfb_kernels.cu, p = 2m taps in registers): against the
    oracle, state carried over an awkward cut, and bit for bit against the plain FIR kernel; complex taps too"""
    X = orc.synth_input(nfr * M, seed=0xD0 + M)
    want = orc.Firpfbch(orc.SYNTHESIZER, M, kaiser=(m, 60.0)).execute_frames(X)
    outs = {}
    for plain in (False, True):
        if plain:
            monkeypatch.setenv("LIQUID_CUDA_SYNTH_NO_WINDOW", "1")
        else:
            monkeypatch.delenv("LIQUID_CUDA_SYNTH_NO_WINDOW", raising=False)
        q = L.firpfbch_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
        cut = (nfr // 2 + 7) * M
        outs[plain] = np.concatenate([q.execute_block(X[:cut]), q.execute_block(X[cut:])])
    monkeypatch.delenv("LIQUID_CUDA_SYNTH_NO_WINDOW", raising=False)
    assert np.abs(outs[False] - want).max() / np.abs(want).max() < 1e-5
    assert np.array_equal(outs[False].view(np.uint32), outs[True].view(np.uint32))
    if M <= 512:
        rng = np.random.default_rng(M)
        h = ((rng.standard_normal(M * 2 * m) + 1j * rng.standard_normal(M * 2 * m)) / (2 * m)).astype(np.complex64)
        got = L.firpfbch_cccf.create(L.LIQUID_SYNTHESIZER, M, 2 * m, h).execute_block(X)
        o = orc.FirpfbchCccf(orc.SYNTHESIZER, M, 2 * m, h).execute_frames(X)
        assert np.abs(got - o).max() / np.abs(o).max() < 1e-5


def test_pfbch_hM3_prototype_vs_reference_fixture(L, refrn2):
    """firpfbch_crcf / firpfbch_cccf create_rnyquist with the harris-Moerder prototype (Parks-McClellan designs on the
    host) against the compiled reference's channelizer outputs"""
    for name in sorted(k for k in refrn2 if k.startswith("pfbch_hM3_") and k.endswith("_args")):
        M, m, beta = refrn2[name]
        M, m = int(M), int(m)
        x = refrn2["pfbch_hM3_M%d_m%d_x" % (M, m)]
        for kind, typ in (("a", L.LIQUID_ANALYZER), ("s", L.LIQUID_SYNTHESIZER)):
            want = refrn2["pfbch_hM3_%s_M%d_m%d_y" % (kind, M, m)]
            got = L.firpfbch_crcf.create_rnyquist(typ, M, m, np.float32(beta), L.LIQUID_FIRFILT_hM3).execute_block(x)
            assert np.abs(got - want).max() / np.abs(want).max() < 1e-5, (M, kind)
            got = L.firpfbch_cccf.create_rnyquist(typ, M, m, np.float32(beta), L.LIQUID_FIRFILT_hM3).execute_block(x)
            assert np.abs(got - want).max() / np.abs(want).max() < 1e-5, (M, kind, "cccf")


@pytest.mark.parametrize("M,m", [(4, 3), (16, 4), (10, 2), (64, 5)])
def test_pfbch_vs_reference_fixtures(L, refout, M, m):
    x = refout["pfbch_a_M%d_m%d_x" % (M, m)]
    for kind, typ in (("a", L.LIQUID_ANALYZER), ("s", L.LIQUID_SYNTHESIZER)):
        want = refout["pfbch_%s_M%d_m%d_y" % (kind, M, m)]
        got = L.firpfbch_crcf.create_kaiser(typ, M, m, 60.0).execute_block(x)
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-5


def test_pfbch_rrc_vs_reference_fixture(L, refout):
    x = refout["pfbch_rrc_M8_m3_x"]
    for kind, typ in (("a", L.LIQUID_ANALYZER), ("s", L.LIQUID_SYNTHESIZER)):
        want = refout["pfbch_rrc_%s_M8_m3_y" % kind]
        got = L.firpfbch_crcf.create_rnyquist(typ, 8, 3, 0.4, L.LIQUID_FIRFILT_RRC).execute_block(x)
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-5


def test_pfbch_analyzer_autotest_equivalence(L, orc):
    """firpfbch_crcf_analyzer_autotest.c:29-152 on the GPU objects (firpfbch + firfilt), tol 1e-4"""
    M, p, S = 4, 5, 40
    rng = np.random.default_rng(9)
    h = (rng.integers(0, 4, M * p) - 1.5).astype(np.float32)
    x = (0.1 * np.sqrt(0.5) * ((rng.integers(0, 4, M * S) - 1.5) + 1j * (rng.integers(0, 4, M * S) - 1.5))).astype(np.complex64)
    q = L.firpfbch_crcf.create(L.LIQUID_ANALYZER, M, p, h)
    Y0 = np.stack([q.analyzer_execute(x[i * M:(i + 1) * M]) for i in range(S)])
    f = L.firfilt_crcf.create(h)
    j = np.arange(M * S)
    for c in range(M):
        f.reset()
        mixed = (x * np.exp(-2j * np.pi * c * j / M)).astype(np.complex64)
        y = f.execute_block(mixed)
        assert np.abs(Y0[:, c] - y[M - 1::M]).max() < 1e-4


def test_pfbch_push_run_alignment(L, orc):
    """analyzer_push / analyzer_run(k) (firpfbch.c:369-409; sandbox/firpfbch_analysis_alignment_test.c)"""
    import ctypes as C
    M, m = 8, 2
    x = orc.synth_input(5 * M + 3, seed=41)
    q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    o = orc.Firpfbch(orc.ANALYZER, M, kaiser=(m, 60.0))
    for i, v in enumerate(x):
        q.analyzer_push(v)
        orc.lib().orc_firpfbch_analyzer_push(o.h, orc.c_cf(v.real, v.imag))
        if i % 5 == 4:
            for k in (0, 3, M - 1):
                want = np.empty(M, np.complex64)
                orc.lib().orc_firpfbch_analyzer_run(o.h, k, want)
                got = q.analyzer_run(k)
                assert np.abs(got - want).max() < 1e-5 * max(1.0, np.abs(want).max())


# ------------------------------------------------------------------ firfilt
@pytest.mark.parametrize("tag", ["h4x8", "h7x16", "h13x32", "h23x64"])
def test_firfilt_reference_golden_vectors(L, kat, tag):
    case = kat["firfilt_crcf"]["cases"][tag]
    y = L.firfilt_crcf.create(np.array(case["h"], np.float32)).execute_block(cplx(case["x"]))
    assert np.abs(y - cplx(case["y"])).max() < kat["firfilt_crcf"]["tol"]
    # push / execute, sample by sample (firfilt_runtest.c:82-90)
    q = L.firfilt_crcf.create(np.array(case["h"], np.float32))
    for xv, yv in zip(cplx(case["x"]), cplx(case["y"])):
        q.push(xv)
        assert abs(q.execute() - yv) < kat["firfilt_crcf"]["tol"]


def test_firfilt_vs_reference_fixture_and_oracle(L, orc, refout):
    x = refout["firfilt_k64_x"]
    q = L.firfilt_crcf.create_kaiser(64, 0.25, 60.0, 0.0)
    assert q.get_length() == 64
    y = np.concatenate([q.execute_block(x[:100]), q.execute_block(x[100:101]), q.execute_block(x[101:])])
    assert np.abs(y - refout["firfilt_k64_y"]).max() < 2e-6
    q.reset(); q.set_scale(0.375)
    assert np.abs(q.execute_block(x) - refout["firfilt_k64_scale_y"]).max() < 2e-6
    # in place (x == y), firfilt.c:340-341
    q.reset(); q.set_scale(1.0)
    buf = x.copy()
    q.execute_block(buf, out=buf)
    assert np.abs(buf - refout["firfilt_k64_y"]).max() < 2e-6
    for n_taps in (1, 2, 5, 31, 200, 1025):
        h = np.random.default_rng(n_taps).standard_normal(n_taps).astype(np.float32)
        xx = orc.synth_input(5000, seed=n_taps)
        want = orc.Firfilt(h=h).execute_block(xx)
        got = L.firfilt_crcf.create(h).execute_block(xx)
        assert np.abs(got - want).max() < 2e-6 * np.sqrt(n_taps) * max(1.0, np.abs(want).max())


def test_firfilt_streams_dev(L, orc):
    import ctypes as C
    n_streams, n = 7, 3000
    x = orc.synth_input(n_streams * n, seed=51).reshape(n_streams, n)
    q = L.firfilt_crcf.create_kaiser(64, 0.25, 60.0, 0.0)
    d_x = L.lib.liquid_cuda_malloc(x.nbytes); d_y = L.lib.liquid_cuda_malloc(x.nbytes)
    L.lib.liquid_cuda_memcpy_h2d(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes)
    q.execute_streams_dev(d_x, n_streams, n, d_y, 0)
    L.lib.liquid_cuda_sync()
    got = np.empty_like(x)
    L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_y, x.nbytes)
    L.lib.liquid_cuda_free(d_x); L.lib.liquid_cuda_free(d_y)
    o = orc.Firfilt(kaiser=(64, 0.25, 60.0, 0.0))
    for s in range(n_streams):
        o.reset()
        assert np.abs(got[s] - o.execute_block(x[s])).max() < 2e-6


def test_firfilt_full_size_spot_checks_vs_oracle(L, orc):
    """BASELINE config 2 at full size (64-tap Kaiser low-pass, 1024 streams x 2^20 samples, 8 GiB in and
    out, device-resident): blocks at stream starts (zero state), at warp-tile boundaries (multiples of
    512 outputs), at pseudo-random places and at stream ends, in the first, a middle and the last stream"""
    import ctypes as C
    n_streams, n, nt = 1024, 1 << 20, 64
    q = L.firfilt_crcf.create_kaiser(nt, 0.25, 60.0, 0.0)
    d_x = L.lib.liquid_cuda_malloc(8 * n_streams * n)
    d_y = L.lib.liquid_cuda_malloc(8 * n_streams * n)
    assert d_x and d_y
    seed = 0x5EED0C02
    L.synth_input_dev(d_x, n_streams * n, seed=seed)
    q.execute_streams_dev(d_x, n_streams, n, d_y, 0)
    L.lib.liquid_cuda_sync()
    o = orc.Firfilt(kaiser=(nt, 0.25, 60.0, 0.0))
    nb = 1100                                             # spans two warp-tile boundaries
    got = np.empty(nb, np.complex64)
    rng = np.random.default_rng(2)
    for s in (0, 1, 517, n_streams - 1):
        offs = {0, 512 - 300, 4096 - 300, n - nb} | {int(v) for v in rng.integers(nt, n - nb, size=3)}
        for o0 in sorted(offs):
            L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_y + 8 * (s * n + o0), got.nbytes)
            lead = min(nt - 1, o0)                        # samples of run-in: the filter's memory, zero at the stream start
            x = orc.synth_input(lead + nb, seed=seed, first=s * n + o0 - lead)
            o.reset()
            want = o.execute_block(x)[lead:]
            assert np.abs(got - want).max() < 2e-6, (s, o0)
    L.lib.liquid_cuda_free(d_x)
    L.lib.liquid_cuda_free(d_y)


@pytest.mark.parametrize("n_taps", [1, 2, 7, 16, 17, 31, 32, 33, 40, 63, 64, 65, 96, 97, 128, 129, 192, 193, 256])
def test_firfilt_fast_fir_form_vs_direct_and_oracle(L, orc, monkeypatch, n_taps):
    """1 - 256 taps take the 2-parallel fast-FIR kernel (instantiated for 16 / 32 / 64 / 96 / 128 / 192 / 256 taps) with the transposed
    store path: against the oracle and the direct register-blocked kernel (up to 64 taps) or the tiled kernel (above), on
    streams whose length is not a multiple of the 512-output warp tile, state carried"""
    import ctypes as C
    rng = np.random.default_rng(n_taps)
    h = (rng.standard_normal(n_taps) / np.sqrt(n_taps)).astype(np.float32)
    n_streams, n = 5, 512 * 9 + 130
    x = orc.synth_input(n_streams * n, seed=0x61 + n_taps).reshape(n_streams, n)
    outs = {}
    for direct in (False, True):
        if direct:
            monkeypatch.setenv("LIQUID_CUDA_FIRFILT_DIRECT", "1")
            if n_taps > 64:
                monkeypatch.setenv("LIQUID_CUDA_NO_REGBLOCK", "1")
        q = L.firfilt_crcf.create(h)
        d_x = L.lib.liquid_cuda_malloc(x.nbytes); d_y = L.lib.liquid_cuda_malloc(x.nbytes)
        L.lib.liquid_cuda_memcpy_h2d(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes)
        q.execute_streams_dev(d_x, n_streams, n, d_y, 0)
        L.lib.liquid_cuda_sync()
        got = np.empty_like(x)
        L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_y, x.nbytes)
        L.lib.liquid_cuda_free(d_x); L.lib.liquid_cuda_free(d_y)
        # one stateful stream cut at an awkward place
        q2 = L.firfilt_crcf.create(h)
        cut = 512 * 3 + 77
        st = np.concatenate([q2.execute_block(x[0][:cut]), q2.execute_block(x[0][cut:])])
        outs[direct] = (got, st)
    monkeypatch.delenv("LIQUID_CUDA_FIRFILT_DIRECT")
    monkeypatch.delenv("LIQUID_CUDA_NO_REGBLOCK", raising=False)
    o = orc.Firfilt(h=h)
    for s in range(n_streams):
        o.reset()
        want = o.execute_block(x[s])
        assert np.abs(outs[False][0][s] - want).max() < 2e-6
        assert np.abs(outs[True][0][s] - want).max() < 2e-6
    assert np.abs(outs[False][1] - outs[False][0][0]).max() < 2e-6
    assert np.abs(outs[False][0] - outs[True][0]).max() < 2e-6


def test_firfilt_freqresponse_groupdelay(L):
    h = np.array([0.1, 0.5, 1.0, 0.5, 0.1], np.float32)
    q = L.firfilt_crcf.create(h)
    H = q.freqresponse(0.1)
    want = (h[::-1] * np.exp(2j * np.pi * 0.1 * np.arange(5))).sum()
    assert abs(H - want) < 1e-5
    assert abs(q.groupdelay(0.1) - 2.0) < 1e-4


# ------------------------------------------------------------------ dotprod / firpfb / fft
def test_dotprod_reference_kats(L, kat):
    for case in kat["dotprod_crcf"]["cases"]:
        h, x, want = np.array(case["h"], np.float32), cplx(case["x"]), complex(*case["y"])
        assert abs(L.dotprod_crcf.run(h, x) - want) < kat["dotprod_crcf"]["tol"]
        assert abs(L.dotprod_crcf(h).execute(x) - want) < kat["dotprod_crcf"]["tol"]


def test_dotprod_struct_vs_ordinal(L):
    """dotprod_crcf_autotest.c:160-166 (n = 1..512), tolerance 1e-4 relative to the vector norm"""
    rng = np.random.default_rng(3)
    for n in [1, 2, 3, 4, 7, 8, 15, 16, 31, 32, 33, 63, 64, 100, 255, 256, 512]:
        h = rng.standard_normal(n).astype(np.float32)
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        want = (h.astype(np.float64) * x.astype(np.complex128)).sum()
        assert abs(L.dotprod_crcf(h).execute(x) - want) < 1e-4 * max(1.0, np.sqrt(n))


@pytest.mark.parametrize("n,stride", [(4, 4), (8, 8), (10, 12), (16, 16), (30, 30), (32, 32), (34, 40), (64, 64), (64, 65), (100, 128), (128, 128), (254, 256), (256, 256), (1000, 1000)])
def test_dotprod_batch_dev_vs_numpy(L, orc, n, stride):
    """the batched device entry point: one thread (n < 8), 2 / 4 / 8 / 16 lanes (8 <= n < 256, aligned) or one warp per
    vector; ragged vector counts, strides longer than the vector, an odd stride (the unaligned path)"""
    import ctypes as C
    rng = np.random.default_rng(n + stride)
    h = rng.standard_normal(n).astype(np.float32)
    for n_vec in (1, 7, 1031):
        x = orc.synth_input(n_vec * stride, seed=0xD0 + n)
        d_x = L.lib.liquid_cuda_malloc(x.nbytes); d_y = L.lib.liquid_cuda_malloc(8 * n_vec)
        L.lib.liquid_cuda_memcpy_h2d(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes)
        L.dotprod_crcf(h).execute_batch_dev(d_x, n_vec, stride, d_y, 0)
        L.lib.liquid_cuda_sync()
        got = np.empty(n_vec, np.complex64)
        L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_y, got.nbytes)
        L.lib.liquid_cuda_free(d_x); L.lib.liquid_cuda_free(d_y)
        want = (x.reshape(n_vec, stride)[:, :n].astype(np.complex128) * h.astype(np.float64)).sum(axis=1)
        assert np.abs(got - want).max() < 1e-5 * np.sqrt(n), (n, stride, n_vec)


def test_firpfb_reference_kat_and_oracle(L, orc, kat):
    k = kat["firpfb"]
    q = L.firpfb_crcf(k["M"], np.array(k["h"], np.float32))
    for v in k["x"]:
        q.push(v)
    for i, want in enumerate(k["y"]):
        got = q.execute(i)
        assert abs(got.real - want) < k["tol"] and abs(got.imag) < k["tol"]
    # batched push+execute against the oracle
    M, m = 8, 4
    qo = orc.Firpfb(M, kaiser=(m, 0.3, 50.0))
    ql = L.firpfb_crcf.create_kaiser(M, m, 0.3, 50.0)
    ql.set_scale(0.5); qo.set_scale(0.5)
    x = orc.synth_input(300, seed=61)
    idx = (np.arange(300) * 5) % M
    want = np.empty(300, np.complex64)
    for i in range(300):
        qo.push(x[i]); want[i] = qo.execute(int(idx[i]))
    got = np.concatenate([ql.execute_block(x[:123], idx[:123]), ql.execute_block(x[123:], idx[123:])])
    assert np.abs(got - want).max() < 2e-6


def test_firpfb_rnyquist_and_drnyquist_vs_reference_fixture(L, refr2):
    """firpfb_crcf_create_rnyquist / _create_drnyquist (firpfb.c:146-239): push + execute against the compiled
    reference's outputs on the same input, sample by sample and as one block"""
    tags = sorted(k[:-5] for k in refr2 if k.startswith("pfbd_") and k.endswith("_args"))
    assert len(tags) == 8
    for tag in tags:
        ftype, M, k, m, beta = refr2[tag + "_args"]
        x, want = refr2[tag + "_x"], refr2[tag + "_y"]
        make = L.firpfb_crcf.create_drnyquist if "_drnyquist_" in tag else L.firpfb_crcf.create_rnyquist
        q = make(int(ftype), int(M), int(k), int(m), np.float32(beta))
        idx = (7 * np.arange(x.size) + 3) % int(M)
        got = q.execute_block(x, idx)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() < 3e-6 * scale, tag
        q.reset()
        for n in range(40):
            q.push(x[n])
            assert abs(q.execute(int(idx[n])) - want[n]) < 3e-6 * scale, (tag, n)
        q.destroy()


@pytest.mark.parametrize("M,m", [(8, 4), (32, 6), (33, 3), (64, 11)])
def test_firpfb_long_blocks_tiled_kernel(L, orc, monkeypatch, M, m):
    """long blocks take the shared-memory tiled kernel: bit for bit against the plain kernel (arbitrary filter index per
    sample, state carried over an awkward cut), and the head of the stream against the oracle run sample by sample"""
    n = 40000
    x = orc.synth_input(n, seed=0xE0 + M)
    idx = (np.random.default_rng(M).integers(0, M, n)).astype(np.uint32)
    outs = {}
    for plain in (False, True):
        if plain:
            monkeypatch.setenv("LIQUID_CUDA_FIRPFB_SIMPLE", "1")
        else:
            monkeypatch.delenv("LIQUID_CUDA_FIRPFB_SIMPLE", raising=False)
        q = L.firpfb_crcf.create_kaiser(M, m, 0.4, 60.0)
        outs[plain] = np.concatenate([q.execute_block(x[:17001], idx[:17001]), q.execute_block(x[17001:], idx[17001:])])
    monkeypatch.delenv("LIQUID_CUDA_FIRPFB_SIMPLE", raising=False)
    assert np.array_equal(outs[False].view(np.uint32), outs[True].view(np.uint32))
    # the one-output-per-thread tile kernel the four-output one replaced
    monkeypatch.setenv("LIQUID_CUDA_FIRPFB_TILE1", "1")
    q = L.firpfb_crcf.create_kaiser(M, m, 0.4, 60.0)
    tile1 = np.concatenate([q.execute_block(x[:17001], idx[:17001]), q.execute_block(x[17001:], idx[17001:])])
    monkeypatch.delenv("LIQUID_CUDA_FIRPFB_TILE1")
    assert np.array_equal(outs[False].view(np.uint32), tile1.view(np.uint32))
    qo = orc.Firpfb(M, kaiser=(m, 0.4, 60.0))
    for i in range(1500):
        qo.push(x[i])
        assert abs(outs[False][i] - qo.execute(int(idx[i]))) < 2e-6


def test_fft_reference_golden_vectors(L, kat):
    """fft_runtest.c:30-66: forward vs golden, inverse round trip, every size the reference ships"""
    for n, case in kat["fft"]["cases"].items():
        x, want = cplx(case["x"]), cplx(case["y"])
        y = L.fft_run(x, L.LIQUID_FFT_FORWARD)
        assert np.abs(y - want).max() < kat["fft"]["tol"], n
        z = L.fft_run(y, L.LIQUID_FFT_BACKWARD) / np.float32(int(n))
        assert np.abs(z - x).max() < kat["fft"]["tol"], n


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_fft_pow2_vs_numpy_and_plan_api(L, orc, n):
    x = orc.synth_input(n * 3, seed=70 + n)
    xin, yout = x[:n].copy(), np.empty(n, np.complex64)
    p = L.fftplan(n, xin, yout, L.LIQUID_FFT_FORWARD)
    p.execute()
    want = np.fft.fft(xin.astype(np.complex128))
    assert np.abs(yout - want).max() / np.abs(want).max() < 2e-6
    yb = p.execute_batch(x)                      # 3 transforms
    wantb = np.fft.fft(x.reshape(3, n).astype(np.complex128), axis=1).reshape(-1)
    assert np.abs(yb - wantb).max() / np.abs(wantb).max() < 2e-6
    back = L.fft_run(yout, L.LIQUID_FFT_BACKWARD)
    assert np.abs(back / n - xin).max() < 1e-5


@pytest.mark.parametrize("n", [6, 12, 18, 20, 30, 100, 300, 360, 1000, 1155, 1200, 2197, 3000, 6000, 7000, 12288])
def test_fft_composite_sizes_mixed_radix_vs_numpy(L, orc, monkeypatch, n):
    """composite sizes take the mixed-radix Stockham kernel (fft_mixed_radix.c:128-191 in the reference; factors 4, 2, 3,
    5, 7, 11, 13): forward and backward against numpy in double, several transforms per CTA and a ragged batch, and
    against the table-driven DFT the sizes ran on before"""
    nb = 5
    x = orc.synth_input(n * nb, seed=170 + n)
    xin, yout = x[:n].copy(), np.empty(n, np.complex64)
    for d, np_fft in ((L.LIQUID_FFT_FORWARD, np.fft.fft), (L.LIQUID_FFT_BACKWARD, lambda a, axis=-1: np.fft.ifft(a, axis=axis) * n)):
        p = L.fftplan(n, xin, yout, d)
        yb = p.execute_batch(x)
        want = np_fft(x.reshape(nb, n).astype(np.complex128), axis=1).reshape(-1)
        assert np.abs(yb - want).max() / np.abs(want).max() < 3e-6
    if n <= 3000:
        monkeypatch.setenv("LIQUID_CUDA_FFT_NO_MIXED", "1")
        yd = L.fftplan(n, xin, yout, L.LIQUID_FFT_BACKWARD).execute_batch(x)
        monkeypatch.delenv("LIQUID_CUDA_FFT_NO_MIXED")
        assert np.abs(yb - yd).max() / np.abs(yd).max() < 3e-6


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_fft_tile_and_quad_transforms_long_ragged_batches(L, orc, monkeypatch, n):
    """n = 4 ... 32 run one thread per transform behind a per-warp staging tile (fft_small_kernel),
    n = 64 ... 512 run the tile transform (tile_fft.cuh), n = 1024 ... 4096 the quad transform (quad_fft.cuh): batches
    long enough that an instance walks several tiles (both tile buffers, the register prefetch) and ragged against the tile
    size, both directions against numpy in double and against the shared-memory Stockham kernel they replace"""
    per_tile = 32 if n <= 32 else 8 if n <= 512 else 4096 // n
    for nb in (1, per_tile + 1, (10_000_000 // n) | 1):      # the long one: 10 M points, an odd number of transforms
        x = orc.synth_input(n * nb, seed=270 + n + nb)
        xin, yout = x[:n].copy(), np.empty(n, np.complex64)
        for d, np_fft in ((L.LIQUID_FFT_FORWARD, np.fft.fft), (L.LIQUID_FFT_BACKWARD, lambda a, axis=-1: np.fft.ifft(a, axis=axis) * n)):
            p = L.fftplan(n, xin, yout, d)
            yb = p.execute_batch(x)
            want = np_fft(x.reshape(nb, n).astype(np.complex128), axis=1).reshape(-1)
            assert np.abs(yb - want).max() / np.abs(want).max() < 2e-6, (n, nb, d)
            monkeypatch.setenv("LIQUID_CUDA_FFT_NO_TILE", "1")
            ys = p.execute_batch(x)
            monkeypatch.delenv("LIQUID_CUDA_FFT_NO_TILE")
            assert np.abs(yb - ys).max() / np.abs(want).max() < 2e-6, (n, nb, d)


@pytest.mark.parametrize("n", [16384, 25600])
def test_fft_large_sizes_on_the_table_driven_path(L, orc, n):
    """sizes that fit neither the shared-memory FFTs nor the mixed-radix buffers still transform (O(n^2), 200 KB of
    shared memory opted in): the limit the error message names is real"""
    x = orc.synth_input(n, seed=9 + n)
    y = L.fft_run(x, L.LIQUID_FFT_FORWARD)
    want = np.fft.fft(x.astype(np.complex128))
    assert np.abs(y - want).max() / np.abs(want).max() < 1e-5


def test_fft_reference_fixtures(L, refout):
    for n in (128, 256, 1024, 4096):
        x = refout["fft_%d_x" % n]
        for d, key in ((L.LIQUID_FFT_FORWARD, "fwd"), (L.LIQUID_FFT_BACKWARD, "bwd")):
            want = refout["fft_%d_%s" % (n, key)]
            assert np.abs(L.fft_run(x, d) - want).max() / np.abs(want).max() < 2e-6


def test_synth_input_dev_matches_host_generator(L, orc):
    import ctypes as C
    n = 100000
    d = L.lib.liquid_cuda_malloc(n * 8)
    L.synth_input_dev(d, n, seed=0x5EED0000 + 3, first=12345)
    L.lib.liquid_cuda_sync()
    got = np.empty(n, np.complex64)
    L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d, n * 8)
    L.lib.liquid_cuda_free(d)
    assert np.array_equal(got, orc.synth_input(n, seed=0x5EED0000 + 3, first=12345))


# ------------------------------------------------------------------ fused M=1024 kernel specifics
@pytest.mark.parametrize("m", [2, 4, 5, 7])
def test_pfbch2_fused_1024_vs_oracle_odd_splits(L, orc, m):
    """the fused kernel (M=1024) across: odd start parity, frame counts that are not a multiple of
    the 4-frame group, several CTAs, state carried between calls"""
    M = 1024
    nfr = 33 + 100 + 515
    x = orc.synth_input(nfr * M // 2, seed=0xF0 + m)
    want = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(x)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    got = np.concatenate([q.execute_block(x[:33 * 512]), q.execute_block(x[33 * 512:133 * 512]),
                          q.execute_block(x[133 * 512:])])
    check_channels(got, want, x)
    # one call over everything gives the same stream
    q.reset()
    whole = q.execute_block(x)
    check_channels(whole, want, x)
    assert np.abs(whole - got).max() < 1e-6


def test_pfbch2_fused_matches_generic_path(L, orc, monkeypatch):
    M, m = 1024, 7
    x = orc.synth_input(3000 * M // 2, seed=0xFE)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    fused = q.execute_block(x)
    q.reset()
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")
    generic = q.execute_block(x)
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert rel_rms_per_channel(fused, generic).max() < 2e-6
    assert np.abs(fused - generic).max() < 1e-6


# ------------------------------------------------------------------ tile kernels, M = 64 ... 512 (pfbch2_tile.cu)
@pytest.mark.parametrize("M,m", [(64, 2), (64, 4), (64, 5), (64, 7), (128, 2), (128, 5), (256, 2), (256, 4), (256, 7),
                                 (512, 2), (512, 5), (512, 7)])
def test_pfbch2_tile_vs_oracle_odd_splits_and_generic(L, orc, monkeypatch, M, m):
    """the reference's own benchmark shapes (firpfbch2_crcf_benchmark.c:70-82) and longer filters on the tile kernel:
    odd start parity, frame counts that are not a multiple of the 8-frame tile, several instances with ragged
    tile counts, state carried between calls; and against the shape-generic path"""
    nfr = 65 + 200 + 1283
    x = orc.synth_input(nfr * M // 2, seed=0x71 + M + m)
    want = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(x)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    l0 = L.launch_count()
    got = np.concatenate([q.execute_block(x[:65 * M // 2]), q.execute_block(x[65 * M // 2:265 * M // 2]),
                          q.execute_block(x[265 * M // 2:])])
    check_channels(got, want, x)
    q.reset()
    whole = q.execute_block(x)
    check_channels(whole, want, x)
    assert np.abs(whole - got).max() < 1e-6
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")
    q.reset()
    generic = q.execute_block(x)
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert rel_rms_per_channel(whole, generic).max() < 2e-6
    assert np.abs(whole - generic).max() < 1e-6


@pytest.mark.parametrize("M,m", [(64, 2), (64, 4), (64, 5), (128, 2), (128, 5), (256, 2), (256, 4), (256, 5),
                                 (512, 2), (512, 4), (512, 5), (1024, 2), (1024, 4)])
def test_pfbch2_tile_synthesizer_vs_oracle_odd_splits_and_generic(L, orc, monkeypatch, M, m):
    """the reference's synthesizer benchmark shapes (firpfbch2_crcf_benchmark.c:84-90, m = 2) and longer filters on
    the fused tile synthesizer: odd start parity, ragged tiles, several instances (warm-up tiles rebuild the
    windows), state handed from the fused kernel to the two-kernel path and back, get/set_state"""
    nfr = 65 + 9 + 200 + 1283
    X = orc.synth_input(nfr * M, seed=0x91 + M + m)
    want = orc.Firpfbch2(orc.SYNTHESIZER, M, m, As=60.0).execute_frames(X)
    scale = np.abs(want).max()
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    cuts = [0, 65, 74, 274, nfr]                      # 9 frames: below 4m for m >= 4 (two-kernel path in between)
    got = np.concatenate([q.execute_block(X[a * M:b * M]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.abs(got - want).max() / scale < 1e-5
    hist, flag = q.get_state()
    q.reset()
    whole = q.execute_block(X)
    assert np.abs(whole - want).max() / scale < 1e-5
    assert np.abs(whole - got).max() / scale < 2e-6
    hist2, flag2 = q.get_state()
    assert flag == flag2 and np.abs(hist - hist2).max() / max(np.abs(hist).max(), 1e-30) < 2e-6
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")
    q.reset()
    generic = q.execute_block(X)
    hist3, _ = q.get_state()
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert np.abs(whole - generic).max() / scale < 2e-6
    assert np.abs(hist - hist3).max() / max(np.abs(hist).max(), 1e-30) < 2e-6


# ------------------------------------------------------------------ the smallest shapes, M = 4 ... 32 (pfbch2_small.cu)
@pytest.mark.parametrize("M,m", [(4, 2), (4, 5), (8, 2), (8, 4), (16, 2), (16, 5), (32, 2), (32, 4)])
def test_pfbch2_small_analyzer_vs_oracle_odd_splits_and_generic(L, orc, monkeypatch, M, m):
    """the low end of the reference's benchmark shapes (firpfbch2_crcf_benchmark.c:70-71: M = 4, 16): several instances
    per warp, odd start parity, ragged tiles, state carried, single-frame execute calls; against the oracle and the
    shape-generic path"""
    nfr = 33 + 200 + 5000 + 3
    x = orc.synth_input(nfr * M // 2, seed=0x41 + M + m)
    want = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(x)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    cuts = [0, 33, 233, 5233, 5234, nfr]
    got = np.concatenate([q.execute_block(x[a * M // 2:b * M // 2]) for a, b in zip(cuts[:-1], cuts[1:])])
    check_channels(got, want, x)
    q.reset()
    whole = q.execute_block(x)
    assert np.array_equal(whole.view(np.uint32), got.view(np.uint32))          # independent of the block cut
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")
    q.reset()
    generic = q.execute_block(x)
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert np.abs(whole - generic).max() < 1e-6


@pytest.mark.parametrize("M,m", [(4, 2), (4, 4), (8, 2), (16, 2), (16, 4), (32, 2), (32, 4)])
def test_pfbch2_small_synthesizer_vs_oracle_odd_splits_and_generic(L, orc, monkeypatch, M, m):
    """synthesizer at M = 4 ... 32 (firpfbch2_crcf_benchmark.c:84-85): warm-up tiles, ragged tiles, odd parity, state handed
    between the fused kernel and the two-kernel path"""
    nfr = 65 + 5 + 200 + 5000
    X = orc.synth_input(nfr * M, seed=0x51 + M + m)
    want = orc.Firpfbch2(orc.SYNTHESIZER, M, m, As=60.0).execute_frames(X)
    scale = np.abs(want).max()
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    cuts = [0, 65, 70, 270, nfr]                      # 5 frames: below 4m (two-kernel path in between)
    got = np.concatenate([q.execute_block(X[a * M:b * M]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.abs(got - want).max() / scale < 1e-5
    hist, flag = q.get_state()
    q.reset()
    whole = q.execute_block(X)
    assert np.abs(whole - want).max() / scale < 1e-5
    assert np.abs(whole - got).max() / scale < 2e-6
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")
    q.reset()
    generic = q.execute_block(X)
    hist3, flag3 = q.get_state()
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert np.abs(whole - generic).max() / scale < 2e-6
    assert flag == flag3 and np.abs(hist - hist3).max() / max(np.abs(hist).max(), 1e-30) < 2e-6


@pytest.mark.parametrize("M,p", [(64, 4), (64, 10), (128, 8), (256, 10), (256, 4), (512, 10), (512, 4), (64, 6), (128, 12), (256, 6), (512, 12)])
def test_pfbch_cccf_complex_taps_tile_analyzer(L, orc, monkeypatch, M, p):
    """firpfbch_cccf analysis with complex taps on the tile kernel (two packed FFMA2 per tap): against the oracle with the
    state carried over awkward cuts (incl. single frames), and against the shape-generic complex-tap kernels"""
    rng = np.random.default_rng(M + p)
    h = ((rng.standard_normal(M * p) + 1j * rng.standard_normal(M * p)) / p).astype(np.complex64)
    nfr = 8 * 41 + 5
    x = orc.synth_input(nfr * M, seed=0xCC + M + p)
    want = orc.FirpfbchCccf(orc.ANALYZER, M, p, h).execute_frames(x)
    q = L.firpfbch_cccf.create(L.LIQUID_ANALYZER, M, p, h)
    cuts = [0, 1, 2, 75, 76, 200, nfr]
    got = np.concatenate([q.execute_block(x[a * M:b * M]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5
    q.reset()
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")
    generic = q.execute_block(x)
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert np.abs(got - generic).max() / np.abs(want).max() < 2e-6


# ------------------------------------------------------------------ random block cuts across every dedicated kernel
@pytest.mark.parametrize("ms,seed", [((2, 4, 5), 20261019), ((3, 6), 20261020), ((7,), 20261021)])
def test_channelizers_random_cuts_vs_oracle(L, orc, ms, seed):
    """every shape with a dedicated kernel (M = 4 ... 1024; m = 2 ... 7; at m = 7 only the firpfbch2 analyzers have one), analysis and synthesis of both channelizers, fed
    in seeded random block sizes that include single frames, blocks below the synthesizers' 4m-frame threshold (two-kernel
    path in between, state handed over) and ragged tiles: one stream of outputs whatever the cuts"""
    rng = np.random.default_rng(seed)
    for M in (4, 8, 16, 32, 64, 128, 256, 512, 1024):
        for m in ms:
            nfr = int(rng.integers(300, 700)) if M >= 256 else int(rng.integers(1500, 4000))
            sizes = []
            left = nfr
            while left > 0:
                k = int(rng.choice([1, 1, 2, 3, 7, 8, 9, 15, 16, 17, 40, 64, 65, 200, 1000]))
                k = min(k, left)
                sizes.append(k)
                left -= k
            cuts = np.concatenate([[0], np.cumsum(sizes)])
            for kind in ("a2", "s2", "a1", "s1"):
                if kind[1] == "2":
                    typ = L.LIQUID_ANALYZER if kind[0] == "a" else L.LIQUID_SYNTHESIZER
                    per_in = M // 2 if kind[0] == "a" else M
                    x = orc.synth_input(nfr * per_in, seed=int(rng.integers(1 << 30)))
                    want = orc.Firpfbch2(orc.ANALYZER if kind[0] == "a" else orc.SYNTHESIZER, M, m, As=60.0).execute_frames(x)
                    q = L.firpfbch2_crcf.create_kaiser(typ, M, m, 60.0)
                else:
                    typ = L.LIQUID_ANALYZER if kind[0] == "a" else L.LIQUID_SYNTHESIZER
                    per_in = M
                    x = orc.synth_input(nfr * per_in, seed=int(rng.integers(1 << 30)))
                    want = orc.Firpfbch(orc.ANALYZER if kind[0] == "a" else orc.SYNTHESIZER, M, kaiser=(m, 60.0)).execute_frames(x)
                    q = L.firpfbch_crcf.create_kaiser(typ, M, m, 60.0)
                got = np.concatenate([np.asarray(q.execute_block(x[a * per_in:b * per_in])).reshape(-1)
                                      for a, b in zip(cuts[:-1], cuts[1:])])
                q.destroy()
                want = np.asarray(want).reshape(-1)
                err = np.abs(got - want).max() / np.abs(want).max()
                assert err < 1e-5, (kind, M, m, err)


# ------------------------------------------------------------------ root-Nyquist Kaiser prototypes
def test_pfbch_rnyquist_kaiser_prototypes_vs_reference(L, orc, refrn):
    """firpfbch_crcf_create_rnyquist with ARKAISER / RKAISER against the compiled reference's outputs and
    the oracle; firfilt / firpfb accept the same prototypes"""
    for name in sorted(k for k in refrn if k.startswith("pfbch_") and k.endswith("_args")):
        M, m, beta, ftype = refrn[name]
        M, m, ftype, beta = int(M), int(m), int(ftype), np.float32(beta)
        proto = name.split("_")[1]
        x = refrn[name[:-5] + "_x"]
        for typ, key, otyp in ((L.LIQUID_ANALYZER, "a", orc.ANALYZER), (L.LIQUID_SYNTHESIZER, "s", orc.SYNTHESIZER)):
            want = refrn["pfbch_%s_%s_M%d_m%d_y" % (proto, key, M, m)]
            got = L.firpfbch_crcf.create_rnyquist(typ, M, m, beta, ftype).execute_block(x)
            assert np.abs(got - want).max() <= 4e-6 * np.abs(want).max(), (name, key)
            worc = orc.Firpfbch(otyp, M, rnyquist=(m, beta, ftype)).execute_frames(x)
            assert np.abs(got - worc).max() <= 2e-6 * np.abs(worc).max(), (name, key)
    # firfilt with an RKAISER prototype == firfilt with the oracle's taps
    x = orc.synth_input(3000, seed=0xAB)
    h = orc.firdes_rkaiser(4, 5, np.float32(0.3))
    got = L.firfilt_crcf.create_rnyquist(L.LIQUID_FIRFILT_RKAISER, 4, 5, 0.3).execute_block(x)
    want = orc.Firfilt(h=h).execute_block(x)
    assert np.abs(got - want).max() < 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("M,m,n_streams,log2n", [(16, 4, 1, 28), (256, 5, 32, 22)])
def test_pfbch_full_size_spot_checks_vs_oracle(L, orc, M, m, n_streams, log2n):
    """BASELINE configs 1 and 4 at full size (M=16: one 2^28-sample stream; M=256: 32 streams x 2^22 -- one
    GPU's share of config 4), device-resident, synthesis followed by analysis for M=256: blocks at stream
    starts, at multiples of 16 and 256 frames (tile / instance granularity of the fused kernels), at
    pseudo-random places and at stream ends against the oracle fed from p frames earlier"""
    import ctypes as C
    n = 1 << log2n
    nf = n // M
    p = 2 * m
    seed = 0x5EED0C00 + M
    d_x = L.lib.liquid_cuda_malloc(8 * n_streams * n)
    d_t = L.lib.liquid_cuda_malloc(8 * n_streams * n)
    d_y = L.lib.liquid_cuda_malloc(8 * n_streams * n)
    L.synth_input_dev(d_x, n_streams * n, seed=seed)
    qa = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    qs = L.firpfbch_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    stages = [("a", qa, d_x, d_y, orc.ANALYZER)]
    if M == 256:      # round trip: synthesise the channel frames, then analyse the synthesised stream
        stages = [("s", qs, d_x, d_t, orc.SYNTHESIZER), ("a", qa, d_t, d_y, orc.ANALYZER)]
    nb = 40
    buf_in = np.empty(((p + nb), M), np.complex64)
    got = np.empty((nb, M), np.complex64)
    rng = np.random.default_rng(M)
    for key, q, d_in, d_out, otyp in stages:
        q.execute_streams_dev(d_in, n_streams, nf, d_out, 0)
        L.lib.liquid_cuda_sync()
        o = orc.Firpfbch(otyp, M, kaiser=(m, 60.0))
        for s in sorted({0, n_streams // 2, n_streams - 1}):
            starts = {0, 16 * 7 - 20, 256 * 3 - 20, nf - nb} | {int(v) for v in rng.integers(p, nf - nb, size=3)}
            for f0 in sorted(starts):
                lead = min(p, f0)                         # frames of run-in: the channelizer's memory is p frames
                # the stage's own input as it is on the device (for "a" after "s" that is the synthesised stream)
                src = buf_in[: lead + nb]
                L.lib.liquid_cuda_memcpy_d2h(src.ctypes.data_as(C.c_void_p), d_in + 8 * (s * n + (f0 - lead) * M), src.nbytes)
                L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_out + 8 * (s * n + f0 * M), got.nbytes)
                o.reset()
                want = o.execute_frames(src.reshape(-1))[lead:]
                if lead < p and f0 > 0:
                    continue
                assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), (key, s, f0)
                if otyp == orc.ANALYZER:
                    check_channels(got, want, src.reshape(-1))
    for d in (d_x, d_t, d_y):
        L.lib.liquid_cuda_free(d)


@pytest.mark.parametrize("M,m", [(2048, 3), (2048, 6), (2048, 8), (4096, 9), (1024, 7)])
def test_pfbch2_two_kernel_synthesizer_register_fir(L, orc, M, m):
    """synthesizer shapes without a fused kernel (batched IDFT, then the register-window FIR instantiated for 4m = 8 ... 36):
    against the oracle, state carried over an awkward cut and over blocks below the 64-frame threshold"""
    nfr = 64 * 3 + 17
    X = orc.synth_input(nfr * M, seed=0x77 + M + m)
    want = orc.Firpfbch2(orc.SYNTHESIZER, M, m, As=60.0).execute_frames(X)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    cuts = [0, 70, 71, 75, 150, nfr]
    got = np.concatenate([q.execute_block(X[a * M:b * M]) for a, b in zip(cuts[:-1], cuts[1:])])
    q.destroy()
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5, (M, m)


@pytest.mark.parametrize("M,m", [(2048, 2), (2048, 7), (4096, 3), (1024, 9), (256, 8), (512, 1), (64, 9), (128, 8), (192, 3)])
def test_column_walking_mac_fallback(L, orc, monkeypatch, M, m):
    """shapes without a dedicated analyzer, long blocks: column-walking MAC kernel + tile / quad transform.  Both
    channelizers against the oracle over awkward cuts (odd frame counts move the firpfbch2 parity; blocks below 64 frames
    take the one-kernel path), and the MAC bit for bit against the element-per-thread kernel"""
    nfr = 64 * 3 + 37
    x2 = orc.synth_input(nfr * M // 2, seed=0xB1 + M + m)
    x1 = orc.synth_input(nfr * M, seed=0xB2 + M + m)
    cuts = [0, 65, 66, 131, 140, nfr]
    for kind, x, per in (("pfbch2", x2, M // 2), ("pfbch", x1, M)):
        mk = (lambda: L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)) if kind == "pfbch2" else \
             (lambda: L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0))
        want = (orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0) if kind == "pfbch2" else orc.Firpfbch(orc.ANALYZER, M, kaiser=(m, 60.0))).execute_frames(x)
        q = mk()
        got = np.concatenate([np.asarray(q.execute_block(x[a * per:b * per])).reshape(-1) for a, b in zip(cuts[:-1], cuts[1:])])
        check_channels(got.reshape(nfr, M), np.asarray(want).reshape(nfr, M), x)
        monkeypatch.setenv("LIQUID_CUDA_MAC_PLAIN", "1")
        monkeypatch.setenv("LIQUID_CUDA_TWO_KERNEL", "1")
        q2 = mk()
        plain = np.concatenate([np.asarray(q2.execute_block(x[a * per:b * per])).reshape(-1) for a, b in zip(cuts[:-1], cuts[1:])])
        monkeypatch.delenv("LIQUID_CUDA_MAC_PLAIN")
        q3 = mk()
        cols = np.concatenate([np.asarray(q3.execute_block(x[a * per:b * per])).reshape(-1) for a, b in zip(cuts[:-1], cuts[1:])])
        monkeypatch.delenv("LIQUID_CUDA_TWO_KERNEL")
        # the same transform after two MAC kernels with the same fmaf sequence: identical bits
        assert np.array_equal(plain.view(np.uint32), cols.view(np.uint32)), (kind, M, m)
        assert np.abs(got - plain).max() / np.abs(np.asarray(want)).max() < 2e-6


# ------------------------------------------------------------------ one-kernel shape-generic analyzers
@pytest.mark.parametrize("M,m,nfr", [(8, 5, 37), (64, 3, 50), (256, 4, 33), (2048, 3, 19), (512, 6, 21)])
def test_generic_one_kernel_analyzers_equal_two_kernel_path(L, orc, monkeypatch, M, m, nfr):
    """power-of-two M without a dedicated fused kernel: MAC + FFT in one kernel (shared memory) gives the
    same bits as polyphase-MAC kernel + batched FFT kernel, for firpfbch2 and firpfbch analysis, and
    matches the oracle"""
    x2 = orc.synth_input(nfr * M // 2, seed=0x61 + M)
    x1 = orc.synth_input(nfr * M, seed=0x62 + M)
    q2 = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    q1 = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")          # keep the dedicated fused kernels out of it
    one2, one1 = q2.execute_block(x2), q1.execute_block(x1)
    q2.reset()
    q1.reset()
    monkeypatch.setenv("LIQUID_CUDA_TWO_KERNEL", "1")
    two2, two1 = q2.execute_block(x2), q1.execute_block(x1)
    monkeypatch.delenv("LIQUID_CUDA_TWO_KERNEL")
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert np.array_equal(one2, two2) and np.array_equal(one1, two1)
    check_channels(one2, orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(x2), x2)
    check_channels(one1, orc.Firpfbch(orc.ANALYZER, M, kaiser=(m, 60.0)).execute_frames(x1), x1)


# ------------------------------------------------------------------ firpfbch_cccf
def test_pfbch_cccf_vs_reference_and_oracle(L, orc, refrn):
    """complex taps: generic complex-tap kernels against the compiled reference's outputs and the oracle,
    in one block and split across calls (state carry)"""
    for name in sorted(k for k in refrn if k.startswith("cccf_") and k.endswith("_h")):
        M, p = (int(v[1:]) for v in name.split("_")[1:3])
        h, x = refrn[name], refrn["cccf_M%d_p%d_x" % (M, p)]
        for typ, key, otyp in ((L.LIQUID_ANALYZER, "a", orc.ANALYZER), (L.LIQUID_SYNTHESIZER, "s", orc.SYNTHESIZER)):
            want = refrn["cccf_M%d_p%d_%s_y" % (M, p, key)]
            q = L.firpfbch_cccf.create(typ, M, p, h)
            got = q.execute_block(x)
            assert np.abs(got - want).max() <= 4e-6 * np.abs(want).max(), (M, p, key)
            q.reset()
            parts = np.concatenate([q.execute_block(x[:7 * M]), q.execute_block(x[7 * M:8 * M]), q.execute_block(x[8 * M:])])
            assert np.array_equal(parts, got)
            worc = orc.FirpfbchCccf(otyp, M, p, h).execute_frames(x)
            assert np.abs(got - worc).max() <= 4e-6 * np.abs(worc).max()
            # one frame at a time through the reference's own entry points
            q.reset()
            fn = q.analyzer_execute if typ == L.LIQUID_ANALYZER else q.synthesizer_execute
            single = np.stack([fn(x[t * M:(t + 1) * M]) for t in range(5)])
            assert np.array_equal(single, got[:5])


def test_pfbch_cccf_real_taps_take_the_crcf_path(L, orc, monkeypatch):
    """create_kaiser / create_rnyquist (real prototypes) equal firpfbch_crcf bit for bit (same kernels), and
    the generic complex-tap kernels agree with them to float rounding; push/run alignment works too"""
    M, m = 256, 5
    x = orc.synth_input(16 * 40 * M // 16, seed=0xCD)
    for typ in (L.LIQUID_ANALYZER, L.LIQUID_SYNTHESIZER):
        a = L.firpfbch_cccf.create_kaiser(typ, M, m, 60.0).execute_block(x)
        b = L.firpfbch_crcf.create_kaiser(typ, M, m, 60.0).execute_block(x)
        assert np.array_equal(a, b)
        monkeypatch.setenv("LIQUID_CUDA_CCCF_GENERIC", "1")
        g = L.firpfbch_cccf.create_kaiser(typ, M, m, 60.0).execute_block(x)
        monkeypatch.delenv("LIQUID_CUDA_CCCF_GENERIC")
        assert np.abs(g - b).max() <= 2e-6 * np.abs(b).max()
    a = L.firpfbch_cccf.create_rnyquist(L.LIQUID_ANALYZER, 8, 3, 0.4, L.LIQUID_FIRFILT_ARKAISER).execute_block(x[:320])
    b = L.firpfbch_crcf.create_rnyquist(L.LIQUID_ANALYZER, 8, 3, 0.4, L.LIQUID_FIRFILT_ARKAISER).execute_block(x[:320])
    assert np.array_equal(a, b)
    # analyzer_push / analyzer_run(k) with complex taps against the oracle-free identity run(0) after M pushes
    rng = np.random.default_rng(3)
    M, p = 8, 4
    h = ((rng.standard_normal(M * p) + 1j * rng.standard_normal(M * p)) / p).astype(np.complex64)
    q = L.firpfbch_cccf.create(L.LIQUID_ANALYZER, M, p, h)
    xs = orc.synth_input(3 * M, seed=0xCE)
    want = L.firpfbch_cccf.create(L.LIQUID_ANALYZER, M, p, h).execute_block(xs)
    for t in range(3):
        for v in xs[t * M:(t + 1) * M]:
            q.analyzer_push(v)
        assert np.array_equal(q.analyzer_run(0), want[t])


# ------------------------------------------------------------------ firdecim / firinterp
def test_firdecim_firinterp_vs_reference_fixture_and_oracle(L, orc, refrn):
    """Kaiser objects and random taps against the compiled reference's outputs; state carried across calls
    of awkward sizes; a long block against the oracle"""
    from test_oracle import _mr_cases
    for tag, M, m in _mr_cases(refrn):
        x, h = refrn[tag + "_x"], refrn[tag + "_h"]
        for make_l, key in ((lambda: L.firdecim_crcf.create_kaiser(M, m, 60.0), "_decim_kaiser"),
                            (lambda: L.firdecim_crcf.create(M, h), "_decim_h")):
            want = refrn[tag + key]
            q = make_l()
            got = np.concatenate([q.execute_block(x[:7 * M]), q.execute_block(x[7 * M:8 * M]), q.execute_block(x[8 * M:])])
            assert np.abs(got - want).max() < 4e-6 * max(1.0, np.abs(want).max()), (tag, key)
            q.clear()
            assert np.abs(q.execute(x[:M]) - want[0]) < 4e-6 * max(1.0, np.abs(want).max())
        xi = x[:300]
        for make_l, key in ((lambda: L.firinterp_crcf.create_kaiser(M, m, 60.0), "_interp_kaiser"),
                            (lambda: L.firinterp_crcf.create(M, h), "_interp_h")):
            want = refrn[tag + key]
            q = make_l()
            got = np.concatenate([q.execute_block(xi[:5]), q.execute_block(xi[5:6]), q.execute_block(xi[6:])])
            assert np.abs(got - want).max() < 4e-6 * max(1.0, np.abs(want).max()), (tag, key)
            q.reset()
            assert np.abs(q.execute(xi[0]) - want[:M]).max() < 4e-6 * max(1.0, np.abs(want).max())
    M, m = 8, 6
    x = orc.synth_input(M * 40000 + 3 * M, seed=0x3A)
    assert np.abs(L.firdecim_crcf.create_kaiser(M, m, 60.0).execute_block(x) - orc.Firdecim(M, kaiser=(m, 60.0)).execute_block(x)).max() < 4e-6
    assert np.abs(L.firinterp_crcf.create_kaiser(M, m, 60.0).execute_block(x[:50000])
                  - orc.Firinterp(M, kaiser=(m, 60.0)).execute_block(x[:50000])).max() < 4e-6
    # root-Nyquist prototype path
    got = L.firdecim_crcf.create_prototype(L.LIQUID_FIRFILT_ARKAISER, 4, 5, 0.3).execute_block(x[:4000])
    want = orc.Firdecim(4, h=orc.firdes_arkaiser(4, 5, np.float32(0.3))).execute_block(x[:4000])
    assert np.abs(got - want).max() < 4e-6 * np.abs(want).max()


@pytest.mark.parametrize("M,m", [(2, 9), (3, 4), (4, 5), (5, 3), (8, 6), (16, 7), (64, 3), (300, 2)])
def test_firinterp_firdecim_long_blocks_vs_oracle(L, orc, M, m):
    """the shared-memory tiled kernels (long blocks) for several rates and lengths incl. odd ones, state carried
    over an awkward cut, against the oracle"""
    n_in = 60000 if M <= 16 else 6000
    x = orc.synth_input(n_in * max(M, 1), seed=0x600 + M)
    rng = np.random.default_rng(M)
    h = rng.standard_normal(2 * M * m + 1).astype(np.float32)
    q, o = L.firinterp_crcf.create(M, h), orc.Firinterp(M, h=h)
    cut = n_in // 3 + 1
    got = np.concatenate([q.execute_block(x[:cut]), q.execute_block(x[cut:n_in])])
    want = o.execute_block(x[:n_in])
    assert np.abs(got - want).max() < 4e-6 * max(1.0, np.abs(want).max())
    qd, od = L.firdecim_crcf.create(M, h), orc.Firdecim(M, h=h)
    cutd = (n_in // 3 + 1) * M
    got = np.concatenate([qd.execute_block(x[:cutd]), qd.execute_block(x[cutd:])])
    want = od.execute_block(x)
    assert np.abs(got - want).max() < 4e-6 * max(1.0, np.abs(want).max())


# ------------------------------------------------------------------ resamp2 / msresamp2
def test_resamp2_vs_reference_fixture(L, orc, refrn2):
    """all five modes and the mixed-mode run (shared windows, toggle) against the compiled reference's outputs;
    the prototype bit for bit against the oracle; the per-call entry points; recreate"""
    from test_oracle import _hb_cases, hb_check
    for tag, m, f0, As in _hb_cases(refrn2):
        q = L.resamp2_crcf(m, f0, As)
        assert np.array_equal(q.get_taps()[1], orc.Resamp2(m, f0, As).taps()), tag
        assert q.get_delay() == 2 * m - 1
        hb_check(q, orc.run_resamp2_mixed, refrn2, tag, 1e-6)
        # per-call API on a fresh state, continuing with blocks
        q.clear()
        x = refrn2[tag + "_x"]
        want = refrn2[tag + "_decim"]
        got = [q.decim_execute(x[0:2]), q.decim_execute(x[2:4])]
        got = np.concatenate([np.array(got, np.complex64), q.decim_execute_block(x[4:9 * 2]), q.decim_execute_block(x[18:])])
        assert np.abs(got - want).max() < 1e-6, tag
        q.clear()
        y = np.concatenate([q.interp_execute(x[0]), q.interp_execute(x[1]), q.interp_execute_block(x[2:])])
        assert np.abs(y - refrn2[tag + "_interp"]).max() < 1e-6, tag
        q.clear()
        pairs = [q.filter_execute(v) for v in x[:5]]
        y0, y1 = q.filter_execute_block(x[5:])
        got = np.concatenate([np.array([p[0] for p in pairs], np.complex64), y0, np.array([p[1] for p in pairs], np.complex64), y1])
        assert np.abs(got - refrn2[tag + "_filter"]).max() < 1e-6, tag
        q.clear()
        y = np.concatenate([q.analyzer_execute(x[0:2]), q.analyzer_execute_block(x[2:])])
        assert np.abs(y - refrn2[tag + "_analyzer"]).max() < 1e-6, tag
        q.clear()
        y = np.concatenate([q.synthesizer_execute(x[0:2]), q.synthesizer_execute_block(x[2:])])
        assert np.abs(y - refrn2[tag + "_synthesizer"]).max() < 1e-6, tag
    q = L.resamp2_crcf(4, 0.0, 60.0)
    q.recreate(4, 0.1, 70.0)
    assert np.array_equal(q.get_taps()[1], orc.Resamp2(4, 0.1, 70.0).taps())
    q.recreate(7, 0.1, 70.0)
    assert np.abs(q.decim_execute_block(refrn2["hb_m7_f+010_x"]) - refrn2["hb_m7_f+010_decim"]).max() < 1e-6


def test_msresamp2_vs_reference_fixture_and_oracle(L, orc, refrn2):
    from test_oracle import _ms_cases
    for tag, typ, S, fc, As, delay in _ms_cases(refrn2):
        q = L.msresamp2_crcf(typ, S, fc, 0.0, As)
        assert q.get_delay() == delay, tag
        assert q.stage_semilengths() == orc.Msresamp2(typ, S, fc, 0.0, As).stage_m(), tag
        x, want = refrn2[tag + "_x"], refrn2[tag + "_y"]
        M = 1 << S
        unit = 1 if typ == L.LIQUID_RESAMP_INTERP else M
        cut = 7 * unit
        got = np.concatenate([q.execute(x[:unit]), q.execute_block(x[unit:cut]), q.execute_block(x[cut:])])
        assert np.abs(got - want).max() < 2e-6 * max(1.0, np.abs(want).max()), tag
        q.reset()
        assert np.abs(q.execute_block(x) - want).max() < 2e-6 * max(1.0, np.abs(want).max()), tag
    # long blocks against the oracle, state carried
    x = orc.synth_input(1 << 20, seed=0x77)
    for typ, S, n in ((L.LIQUID_RESAMP_DECIM, 4, 1 << 20), (L.LIQUID_RESAMP_INTERP, 3, 1 << 16)):
        q = L.msresamp2_crcf(typ, S, 0.4, 0.0, 60.0)
        o = orc.Msresamp2(typ, S, 0.4, 0.0, 60.0)
        cut = (n // 3) & ~15
        got = np.concatenate([q.execute_block(x[:cut]), q.execute_block(x[cut:n])])
        want = o.execute_block(x[:n])
        assert np.abs(got - want).max() < 4e-6 * max(1.0, np.abs(want).max())


def test_msresamp2_fused_cascade_equals_stage_by_stage(L, orc, monkeypatch):
    """the one-kernel decimating cascade against the stage-by-stage form of the same library, bit for bit: several
    stage counts, blocks of awkward sizes (state carried through every stage's windows), and against the oracle"""
    x = orc.synth_input(3 << 20, seed=0x51)
    for S, fc, As in ((2, 0.3, 60.0), (3, 0.45, 80.0), (4, 0.4, 60.0), (6, 0.2, 70.0), (8, 0.1, 60.0)):
        M = 1 << S
        cuts = [0, 40000 * M // 16 * 16 // M * M, (40000 * M // 16 * 16 // M + 9001) * M, len(x) // M * M]
        cuts = [min(c, len(x) // M * M) for c in cuts]
        outs = {}
        for staged in (False, True):
            if staged:
                monkeypatch.setenv("LIQUID_CUDA_MSRESAMP2_STAGED", "1")
            else:
                monkeypatch.delenv("LIQUID_CUDA_MSRESAMP2_STAGED", raising=False)
            q = L.msresamp2_crcf(L.LIQUID_RESAMP_DECIM, S, fc, 0.0, As)
            outs[staged] = np.concatenate([q.execute_block(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a])
        monkeypatch.delenv("LIQUID_CUDA_MSRESAMP2_STAGED", raising=False)
        assert np.array_equal(outs[False].view(np.uint32), outs[True].view(np.uint32)), S
        if S <= 4:
            want = orc.Msresamp2(orc.RESAMP_DECIM, S, fc, 0.0, As).execute_block(x[:cuts[-1]])
            assert np.abs(outs[False] - want).max() < 4e-6 * max(1.0, np.abs(want).max()), S


def test_resamp2_long_blocks_vs_oracle(L, orc):
    """2^21 samples through each mode, two blocks with an odd split for the filter mode"""
    x = orc.synth_input(1 << 21, seed=0x99)
    m = 9
    for mode in ("analyzer", "synthesizer", "decim", "interp"):
        q, o = L.resamp2_crcf(m, 0.05, 60.0), orc.Resamp2(m, 0.05, 60.0)
        n = len(x) if mode != "interp" else len(x) // 2
        cut = 333334
        fn = getattr(q, mode + "_execute_block")
        got = np.concatenate([fn(x[:cut]), fn(x[cut:n])])
        assert np.abs(got - getattr(o, mode + "_execute")(x[:n])).max() < 2e-6, mode
    # the modes in turn on ONE object, long blocks (the windows each mode leaves are the next one's history)
    q, o = L.resamp2_crcf(m, 0.05, 60.0), orc.Resamp2(m, 0.05, 60.0)
    pos = 0
    for mode, ln in (("decim", 100000), ("interp", 50001), ("analyzer", 70002), ("interp", 40000), ("synthesizer", 60000),
                     ("decim", 40000)):
        seg = x[pos:pos + ln]
        pos += ln
        assert np.abs(getattr(q, mode + "_execute_block")(seg) - getattr(o, mode + "_execute")(seg)).max() < 2e-6, mode
    q, o = L.resamp2_crcf(m, 0.05, 60.0), orc.Resamp2(m, 0.05, 60.0)
    a0, a1 = q.filter_execute_block(x[:777777])
    b0, b1 = q.filter_execute_block(x[777777:])
    w0, w1 = o.filter_execute(x)
    assert np.abs(np.concatenate([a0, b0]) - w0).max() < 2e-6
    assert np.abs(np.concatenate([a1, b1]) - w1).max() < 2e-6


# ------------------------------------------------------------------ asgramcf
def test_asgramcf_vs_reference_fixture(L, orc, refrn2):
    from test_oracle import _asg_cases, asgram_check
    for tag, nfft, offset, scale in _asg_cases(refrn2):
        asgram_check(L.asgramcf(nfft), refrn2, tag, offset, scale)
    # push / reset / a second spectrum on the same object against the oracle
    q, o = L.asgramcf(64), orc.Asgramcf(64)
    x = (orc.synth_input(200, seed=3) * np.float32(0.2) + np.exp(-2j * np.pi * 0.3 * np.arange(200))).astype(np.complex64)
    for v in x[:70]:
        q.push(v)
    o.write(x[:70])
    q.write(x[70:])
    o.write(x[70:])
    (ga, gv, gf), (wa, wv, wf) = q.execute(), o.execute()
    assert sum(a != b for a, b in zip(ga, wa)) <= 1 and abs(gv - wv) < 1e-3 and gf == wf
    q.reset()
    o.reset()
    q.write(x[:64])
    o.write(x[:64])
    (ga, gv, gf), (wa, wv, wf) = q.execute(), o.execute()
    assert sum(a != b for a, b in zip(ga, wa)) <= 1 and abs(gv - wv) < 1e-3 and gf == wf


# ------------------------------------------------------------------ ofdmframegen
def test_ofdmframegen_vs_reference_fixture_and_oracle(L, orc, refrn2, monkeypatch):
    """whole frames against the compiled reference (default / custom allocation, power-of-two and other M, no
    taper), symbol-by-symbol and block calls mixed on one object (pilot sequence and overlap carried), the
    one-kernel path against the scratch path, and long blocks against the oracle"""
    from test_oracle import _ofdm_cases
    for tag, M, cp, tl, p in _ofdm_cases(refrn2):
        x, want = refrn2[tag + "_x"], refrn2[tag + "_frame"]
        for generic in (False, True):
            if generic:
                monkeypatch.setenv("LIQUID_CUDA_OFDM_GENERIC", "1")
            else:
                monkeypatch.delenv("LIQUID_CUDA_OFDM_GENERIC", raising=False)
            q = L.ofdmframegen(M, cp, tl, p)
            assert np.array_equal(q.sctype(), orc.Ofdmframegen(M, cp, tl, p).sctype())
            got = np.concatenate([q.write_S0a(), q.write_S0b(), q.write_S1(), q.writesymbol(x[:M]), q.writesymbols(x[M:3 * M]),
                                  q.writesymbols(x[3 * M:]), q.writetail()])
            assert got.shape == want.shape and np.abs(got - want).max() < 3e-6 * np.abs(want).max(), (tag, generic)
            q.reset()
            assert np.abs(orc.ofdm_frame(q, x) - want).max() < 3e-6 * np.abs(want).max(), (tag, generic)
        monkeypatch.delenv("LIQUID_CUDA_OFDM_GENERIC", raising=False)
    assert np.array_equal(L.ofdmframe_init_default_sctype(120), orc.Ofdmframegen(120, 8, 2).sctype())
    for M, cp, tl, nsym in ((64, 16, 4, 5000), (256, 64, 16, 3000), (2048, 512, 64, 300), (200, 40, 10, 700)):
        x = orc.synth_input(M * nsym, seed=0xF0 + M)
        q, o = L.ofdmframegen(M, cp, tl), orc.Ofdmframegen(M, cp, tl)
        cut = (nsym // 3 + 1) * M
        got = np.concatenate([q.writesymbols(x[:cut]), q.writesymbols(x[cut:]), q.writetail()])
        want = np.concatenate([o.writesymbols(x), o.writetail()])
        assert np.abs(got - want).max() < 4e-6 * np.abs(want).max(), M


# ------------------------------------------------------------------ fftfilt
def test_fftfilt_vs_reference_fixture_and_oracle(L, orc, refrn2, monkeypatch):
    """the compiled reference's outputs (power-of-two and other block sizes, set_scale), state carried over calls of
    one and of several blocks, the shared-memory kernel against the scratch-buffer path, long inputs against the
    oracle's overlap-add and against firfilt"""
    from test_oracle import _ff_cases
    for tag, n, scale in _ff_cases(refrn2):
        x, h, want = refrn2[tag + "_x"], refrn2[tag + "_h"], refrn2[tag + "_y"]
        for generic in (False, True):
            if generic:
                monkeypatch.setenv("LIQUID_CUDA_FFTFILT_GENERIC", "1")
            else:
                monkeypatch.delenv("LIQUID_CUDA_FFTFILT_GENERIC", raising=False)
            q = L.fftfilt_crcf(h, n)
            assert q.get_length() == len(h)
            if scale is not None:
                q.set_scale(scale)
            got = np.concatenate([q.execute(x[:n]), q.execute_block(x[n:3 * n]), q.execute_block(x[3 * n:])])
            assert np.abs(got - want).max() < 2e-6 * np.abs(want).max(), (tag, generic)
            q.reset()
            assert np.abs(q.execute_block(x) - want).max() < 2e-6 * np.abs(want).max(), (tag, generic)
        monkeypatch.delenv("LIQUID_CUDA_FFTFILT_GENERIC", raising=False)
    x = orc.synth_input(1 << 20, seed=0x41)
    for h_len, n in ((129, 128), (40, 128), (257, 256), (100, 256), (513, 512), (77, 512), (1025, 1024), (300, 2048), (2049, 2048), (200, 500)):
        h = orc.firdes_kaiser(h_len, 0.15, 60.0)
        q = L.fftfilt_crcf(h, n)
        nb = len(x) // n
        cut = (nb // 3) * n
        got = np.concatenate([q.execute_block(x[:cut]), q.execute_block(x[cut:nb * n])])
        want = orc.Fftfilt(h, n).execute_blocks(x[:nb * n])
        assert np.abs(got - want).max() < 4e-6 * np.abs(want).max(), (h_len, n)
        if n in (128, 256, 512, 1024, 2048):
            # 2n = 256 / 512 run the chained tile transforms, 2n = 1024 ... 4096 the chained quad transforms: against the
            # shared-memory Stockham kernel they replaced
            monkeypatch.setenv("LIQUID_CUDA_FFTFILT_STOCKHAM", "1")
            q2 = L.fftfilt_crcf(h, n)
            old = np.concatenate([q2.execute_block(x[:cut]), q2.execute_block(x[cut:nb * n])])
            monkeypatch.delenv("LIQUID_CUDA_FFTFILT_STOCKHAM")
            assert np.abs(got - old).max() < 4e-6 * np.abs(want).max(), (h_len, n)
        if h_len <= 300:
            ff = L.firfilt_crcf.create(h).execute_block(x[:nb * n])
            assert np.abs(got - ff).max() < 4e-6 * np.abs(want).max(), (h_len, n)


# ------------------------------------------------------------------ spgramcf
def test_spgramcf_vs_reference_fixture_and_oracle(L, orc, refrn):
    """every entry point against the compiled reference's outputs (power-of-two and other nfft), plus a
    long block against the oracle: many transforms per call, several fold chunks, alpha < 1"""
    from test_oracle import _spgram_cases, _spgram_run, lin_close, spgram_check
    for tag, nfft, wl, beta in _spgram_cases(refrn):
        q = L.spgramcf.create_kaiser(nfft, wl, beta)
        spgram_check(_spgram_run(q, refrn[tag + "_x"], nfft, wl), refrn, tag)
    nfft, wl = 256, 160
    x = (orc.synth_input(3000 * nfft, seed=0x59) * np.float32(0.2)
         + np.exp(2j * np.pi * 0.11 * np.arange(3000 * nfft)).astype(np.complex64)).astype(np.complex64)
    q, o = L.spgramcf.create_kaiser(nfft, wl, 8.0), orc.Spgramcf(nfft, window_len=wl, beta=8.0)
    for part, alpha in ((x[:1000], 0.5), (x[1000:400000], 0.02), (x[400000:400077], 0.02), (x[400077:], 0.001)):
        q.accumulate_psd(part, alpha)
        o.accumulate_psd(part, alpha)
        assert q.num_transforms() == o.num_transforms()
        assert lin_close(q.write_accumulation(), o.write_accumulation(), rel=1e-5)
    assert lin_close(q.estimate_psd(x), o.estimate_psd(x), rel=1e-5)
    # push one sample at a time == write
    q.reset(); o.reset()
    for v in x[:wl + 3]:
        q.push(v)
    o.write(x[:wl + 3])
    assert np.abs(q.execute() - o.execute()).max() < 2e-6 * np.abs(o.execute()).max()


@pytest.mark.parametrize("nfft,wl", [(64, 64), (128, 50), (256, 128), (512, 301), (1024, 512), (2048, 1100), (4096, 2048)])
def test_spgramcf_tile_and_quad_transforms(L, orc, monkeypatch, nfft, wl):
    """nfft = 64 ... 4096 fold their transforms on the tile / quad transforms: several chunks per call, ragged tiles, the
    exponential average (order matters) and the plain sum, against the oracle and against the shared-memory kernel"""
    from test_oracle import lin_close
    n = 700 * nfft + 37
    x = (orc.synth_input(n, seed=0x5A + nfft) * np.float32(0.2)
         + np.exp(2j * np.pi * 0.07 * np.arange(n)).astype(np.complex64)).astype(np.complex64)
    q, o = L.spgramcf.create_kaiser(nfft, wl, 7.0), orc.Spgramcf(nfft, window_len=wl, beta=7.0)
    for part, alpha in ((x[:3 * nfft], 0.5), (x[3 * nfft:n // 2], 0.05), (x[n // 2:n // 2 + 11], 0.05), (x[n // 2 + 11:], 0.004)):
        q.accumulate_psd(part, alpha)
        o.accumulate_psd(part, alpha)
        assert q.num_transforms() == o.num_transforms()
        assert lin_close(q.write_accumulation(), o.write_accumulation(), rel=2e-5), (nfft, alpha)
    est = q.estimate_psd(x)
    assert lin_close(est, o.estimate_psd(x), rel=2e-5)
    monkeypatch.setenv("LIQUID_CUDA_SPGRAM_NO_TILE", "1")
    q2 = L.spgramcf.create_kaiser(nfft, wl, 7.0)
    est2 = q2.estimate_psd(x)
    monkeypatch.delenv("LIQUID_CUDA_SPGRAM_NO_TILE")
    assert lin_close(est, est2, rel=2e-5)


# ------------------------------------------------------------------ cluster-of-8 M=4096 kernel specifics
@pytest.mark.parametrize("m", [2, 3, 4, 5, 6, 7, 8, 9])
def test_pfbch2_cluster_4096_vs_oracle_odd_splits(L, orc, m):
    """the cluster kernel (M=4096, one frame per 8 CTAs, DSMEM exchange) across: odd start parity,
    frame counts that are not a multiple of the 4-frame group, several clusters, state carried
    between calls"""
    M = 4096
    nfr = 33 + 61 + 170
    x = orc.synth_input(nfr * M // 2, seed=0xC0 + m)
    want = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(x)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    got = np.concatenate([q.execute_block(x[:33 * 2048]), q.execute_block(x[33 * 2048:94 * 2048]),
                          q.execute_block(x[94 * 2048:])])
    check_channels(got, want, x)
    q.reset()
    whole = q.execute_block(x)
    check_channels(whole, want, x)
    assert np.abs(whole - got).max() < 1e-6


def test_pfbch2_cluster_matches_generic_path(L, orc, monkeypatch):
    """many groups per cluster (ring wrap-around, credits) against the two-kernel generic path"""
    M, m = 4096, 9
    x = orc.synth_input(3000 * M // 2, seed=0xCE)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    launches0 = L.launch_count()
    fused = q.execute_block(x)
    q.reset()
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")
    generic = q.execute_block(x)
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert rel_rms_per_channel(fused, generic).max() < 2e-6
    assert np.abs(fused - generic).max() < 1e-6
    assert L.launch_count() > launches0


# ------------------------------------------------------------------ fused firpfbch analysis kernels
@pytest.mark.parametrize("M,m,nfr", [(16, 4, 16 * 37 + 5), (16, 2, 16 * 300), (256, 5, 16 * 21 + 3), (256, 2, 16 * 40),
                                     (64, 2, 1500), (64, 5, 8 * 77 + 3), (128, 4, 8 * 90 + 1), (128, 2, 1111),
                                     (512, 2, 8 * 61 + 5), (512, 5, 8 * 40 + 7), (512, 4, 333),
                                     (1024, 2, 4 * 177 + 3), (1024, 5, 4 * 50 + 1), (1024, 4, 301),
                                     (4, 2, 5003), (4, 5, 777), (8, 2, 3001), (8, 4, 16 * 40 + 9), (32, 2, 2000), (32, 5, 16 * 33 + 5)])
def test_pfbch_fused_analysis_vs_oracle_and_generic(L, orc, monkeypatch, M, m, nfr):
    """fused M=16 / M=256 analyzers and the tile analyzers for M = 64 / 128 / 512 / 1024 (the reference's own benchmark shapes,
    firpfbch_crcf_benchmark.c:80-85): many tiles and instances, a tail that is not a multiple of the
    tile, state carried across calls; against the oracle and against the generic kernels"""
    x = orc.synth_input(nfr * M, seed=0xD0 + M + m)
    want = orc.Firpfbch(orc.ANALYZER, M, kaiser=(m, 60.0)).execute_frames(x)
    q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    cut = 16 * 5 + 7
    got = np.concatenate([q.execute_block(x[:cut * M]), q.execute_block(x[cut * M:])])
    assert rel_rms_per_channel(got, want).max() < TOL_REL_RMS
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5
    q.reset()
    monkeypatch.setenv("LIQUID_CUDA_NO_FUSED", "1")
    generic = q.execute_block(x)
    monkeypatch.delenv("LIQUID_CUDA_NO_FUSED")
    assert np.abs(got - generic).max() / np.abs(want).max() < 2e-6


@pytest.mark.parametrize("M,m", [(16, 4), (256, 5), (64, 3), (64, 2), (512, 5), (128, 4), (1024, 2), (4, 2), (8, 4), (32, 2)])
def test_pfbch_streams_dev(L, orc, M, m):
    """stateless multi-stream analyzer form (BASELINE config 4 shape): every stream from the reset state"""
    import ctypes as C
    n_streams, nfr = 5, 16 * 9 + 4
    x = orc.synth_input(n_streams * nfr * M, seed=0xE0 + M).reshape(n_streams, nfr * M)
    q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    d_x = L.lib.liquid_cuda_malloc(x.nbytes); d_y = L.lib.liquid_cuda_malloc(x.nbytes)
    L.lib.liquid_cuda_memcpy_h2d(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes)
    q.execute_streams_dev(d_x, n_streams, nfr, d_y, 0)
    L.lib.liquid_cuda_sync()
    got = np.empty_like(x)
    L.lib.liquid_cuda_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_y, x.nbytes)
    L.lib.liquid_cuda_free(d_x); L.lib.liquid_cuda_free(d_y)
    for s in range(n_streams):
        want = orc.Firpfbch(orc.ANALYZER, M, kaiser=(m, 60.0)).execute_frames(x[s])
        assert np.abs(got[s].reshape(nfr, M) - want).max() / np.abs(want).max() < 1e-5


@pytest.mark.parametrize("M,m,nfr", [(256, 5, 16 * 40 + 9), (256, 2, 16 * 33), (64, 3, 50), (1024, 2, 4 * 77 + 3), (1024, 5, 203),
                                     (1024, 4, 1300), (64, 2, 1500), (64, 5, 8 * 77 + 3), (128, 4, 700), (512, 2, 8 * 61 + 5),
                                     (512, 5, 333)])
def test_pfbch_synthesis_fused_and_streams(L, orc, M, m, nfr):
    """fused M=256 / M=1024 synthesizers (many tiles / instances, tail, state carry) and the multi-stream form"""
    import ctypes as C
    X = orc.synth_input(nfr * M, seed=0xC5 + M + m)
    want = orc.Firpfbch(orc.SYNTHESIZER, M, kaiser=(m, 60.0)).execute_frames(X)
    q = L.firpfbch_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    cut = 16 * 2 + 5
    got = np.concatenate([q.execute_block(X[:cut * M]), q.execute_block(X[cut * M:cut * M + 3 * M]),
                          q.execute_block(X[(cut + 3) * M:])])
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5
    n_streams = 3
    xs = orc.synth_input(n_streams * nfr * M, seed=0xC9).reshape(n_streams, nfr * M)
    d_x = L.lib.liquid_cuda_malloc(xs.nbytes); d_y = L.lib.liquid_cuda_malloc(xs.nbytes)
    L.lib.liquid_cuda_memcpy_h2d(d_x, xs.ctypes.data_as(C.c_void_p), xs.nbytes)
    q.execute_streams_dev(d_x, n_streams, nfr, d_y, 0)
    L.lib.liquid_cuda_sync()
    out = np.empty_like(xs)
    L.lib.liquid_cuda_memcpy_d2h(out.ctypes.data_as(C.c_void_p), d_y, xs.nbytes)
    L.lib.liquid_cuda_free(d_x); L.lib.liquid_cuda_free(d_y)
    for s in range(n_streams):
        w = orc.Firpfbch(orc.SYNTHESIZER, M, kaiser=(m, 60.0)).execute_frames(xs[s])
        assert np.abs(out[s].reshape(nfr, M) - w).max() / np.abs(w).max() < 1e-5


# ------------------------------------------------------------------ randomized shapes (the 8f objects)
def test_randomized_shapes_vs_oracle(L, orc):
    """seeded random shapes and block cuts for the rate changers, fftfilt and ofdmframegen against the oracle: sizes
    around the thresholds where the tiled / fused kernels take over, odd lengths, one-unit blocks"""
    rng = np.random.default_rng(20261023)

    def cuts(n_units, k=3):
        c = sorted(set([0, n_units] + [int(v) for v in rng.integers(0, n_units + 1, k)]))
        return list(zip(c[:-1], c[1:]))

    for _ in range(12):
        M, m = int(rng.integers(2, 40)), int(rng.integers(1, 8))
        n = int(rng.choice([1, 7, 250, 1100, 9000]))
        h = rng.standard_normal(int(rng.integers(M, 2 * M * m + 2))).astype(np.float32)
        x = orc.synth_input(n * M, seed=int(rng.integers(1 << 30)))
        q, o = L.firdecim_crcf.create(M, h), orc.Firdecim(M, h=h)
        got = np.concatenate([q.execute_block(x[a * M:b * M]) for a, b in cuts(n)])
        want = o.execute_block(x)
        assert np.abs(got - want).max() < 5e-6 * max(1.0, np.abs(want).max()), ("firdecim", M, len(h), n)
        q, o = L.firinterp_crcf.create(M, h), orc.Firinterp(M, h=h)
        got = np.concatenate([q.execute_block(x[a:b]) for a, b in cuts(n)])
        want = o.execute_block(x[:n])
        assert np.abs(got - want).max() < 5e-6 * max(1.0, np.abs(want).max()), ("firinterp", M, len(h), n)
    for _ in range(8):
        m, f0 = int(rng.integers(2, 20)), float(rng.uniform(-0.5, 0.5))
        n = int(rng.choice([2, 30, 8190, 8194, 40000]))
        x = orc.synth_input(2 * n, seed=int(rng.integers(1 << 30)))
        for mode in ("decim", "interp", "analyzer", "synthesizer"):
            q, o = L.resamp2_crcf(m, f0, 60.0), orc.Resamp2(m, f0, 60.0)
            per = 1 if mode == "interp" else 2
            got = np.concatenate([getattr(q, mode + "_execute_block")(x[a * per:b * per]) for a, b in cuts(n)])
            want = getattr(o, mode + "_execute")(x[:n * per])
            assert np.abs(got - want).max() < 3e-6, ("resamp2", mode, m, n)
    for _ in range(6):
        S, typ = int(rng.integers(1, 7)), int(rng.integers(0, 2))
        fc, As = float(rng.uniform(0.05, 0.45)), float(rng.choice([40.0, 60.0, 90.0]))
        n = int(rng.choice([1, 5, 300, 5000]))
        q, o = L.msresamp2_crcf(typ, S, fc, 0.0, As), orc.Msresamp2(typ, S, fc, 0.0, As)
        unit = 1 if typ == L.LIQUID_RESAMP_INTERP else (1 << S)
        x = orc.synth_input(n * unit, seed=int(rng.integers(1 << 30)))
        got = np.concatenate([q.execute_block(x[a * unit:b * unit]) for a, b in cuts(n)])
        want = o.execute_block(x)
        assert np.abs(got - want).max() < 5e-6 * max(1.0, np.abs(want).max()), ("msresamp2", typ, S, n)
    for _ in range(8):
        n = int(rng.choice([1, 2, 5, 64, 100, 512, 1000, 4096]))
        h_len = int(rng.integers(1, n + 2))
        nb = int(rng.choice([1, 3, 40]))
        h = rng.standard_normal(h_len).astype(np.float32)
        x = orc.synth_input(n * nb, seed=int(rng.integers(1 << 30)))
        q, o = L.fftfilt_crcf(h, n), orc.Fftfilt(h, n)
        got = np.concatenate([q.execute_block(x[a * n:b * n]) for a, b in cuts(nb, 2)])
        want = o.execute_blocks(x)
        assert np.abs(got - want).max() < 6e-6 * max(1.0, np.abs(want).max()), ("fftfilt", h_len, n, nb)
    for _ in range(8):
        M = int(rng.choice([8, 12, 16, 50, 64, 128, 2048]))
        cp = int(rng.integers(0, M // 2 + 1))
        tl = int(rng.integers(0, cp + 1))
        nsym = int(rng.choice([1, 2, 17, 300]))
        try:
            o = orc.Ofdmframegen(M, cp, tl)
        except Exception:       # (an allocation with fewer than two pilots: the reference rejects it too)
            continue
        if not o.h:
            continue
        x = orc.synth_input(M * nsym, seed=int(rng.integers(1 << 30)))
        q = L.ofdmframegen(M, cp, tl)
        got = np.concatenate([q.write_S0a(), q.write_S0b(), q.write_S1()] + [q.writesymbols(x[a * M:b * M]) for a, b in cuts(nsym, 2)]
                             + [q.writetail()])
        want = orc.ofdm_frame(o, x)
        assert np.abs(got - want).max() < 4e-6 * max(1.0, np.abs(want).max()), ("ofdm", M, cp, tl, nsym)


# ------------------------------------------------------------------ the reference's own example programs on the drop-in
REF_EXAMPLES_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "examples")


@pytest.mark.parametrize("name", ["firpfbch2_crcf_example", "firpfbch_crcf_analysis_example", "firpfbch_crcf_synthesis_example",
                                  "firdecim_crcf_example", "msresamp2_crcf_example", "resamp2_crcf_decim_example",
                                  "firfilt_crcf_example", "firinterp_crcf_example", "resamp2_crcf_interp_example",
                                  "resamp2_crcf_filter_example", "resamp2_crcf_example"])
def test_reference_example_program_unmodified_on_the_drop_in(L, name, tmp_path):
    """a binary of the reference's own example (built by oracle/Makefile from /root/reference/examples, unmodified,
    linked against the compiled reference) run twice: as it is, and with libliquid_cuda.so preloaded so that the
    hot-path symbols resolve to this library -- every number it writes to its .m file must agree"""
    import re
    import subprocess
    exe = os.path.join(REF_EXAMPLES_DIR, name)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/examples not built (needs /root/reference at build time)")
    outs = []
    for tag, preload in (("ref", None), ("cuda", L.LIB_PATH)):
        d = tmp_path / tag
        d.mkdir()
        env = dict(os.environ)
        env["LIQUID_CUDA_VERBOSE"] = "1"                  # the library announces itself on stderr when it is in charge
        if preload:
            env["LD_PRELOAD"] = preload
        r = subprocess.run([exe], cwd=str(d), env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (tag, r.stderr[-400:])
        assert ("liquid_cuda: active on device" in r.stderr) == bool(preload), (tag, r.stderr[-400:])
        text = (d / (name + ".m")).read_text()
        nums = [float(v) for v in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", text)]
        outs.append(np.array(nums))
    ref, got = outs
    assert ref.shape == got.shape and ref.size > 50
    assert np.abs(ref - got).max() <= 2e-5 * max(1.0, np.abs(ref).max()), name


# ------------------------------------------------------------------ ofdmframesync (receive side)
@pytest.mark.parametrize("tag", ["osync_M64", "osync_M256", "osync_M48", "osync_M1024"])
def test_ofdmframesync_rows_vs_reference_fixture(L, refr2, tag):
    """ofdmframegen -> channel (delay, gain, carrier offset, noise) -> ofdmframesync: every row the callback receives
    against the rows the compiled reference's callback received on the same samples, fed in the same uneven chunks
    (acquisition on the host, the payload symbols in one device launch per execute() call)"""
    M, cp, taper, nsym, cfo_ref = refr2[tag + "_args"]
    M, cp, taper, nsym = int(M), int(cp), int(taper), int(nsym)
    y, chunks, want = refr2[tag + "_y"], [int(c) for c in refr2[tag + "_chunks"]], refr2[tag + "_rows"]
    rows = []
    q = L.ofdmframesync(M, cp, taper, None, lambda row, p: rows.append(row) or 0)
    pos = k = 0
    while pos < y.size:
        c = chunks[k % len(chunks)]
        k += 1
        q.execute(y[pos: pos + c])
        pos += c
    assert len(rows) >= nsym, (len(rows), nsym)
    got = np.array(rows[:nsym])
    scale = np.abs(want).max()
    assert np.abs(got - want).max() < 2e-4 * scale, float(np.abs(got - want).max() / scale)
    # (the fixture's carrier offset is read after the three all-noise symbols that follow the frame, whose pilot
    # phases are arbitrary: only its neighbourhood is comparable)
    assert abs(q.get_cfo() - cfo_ref) < 1e-2
    q.destroy()


def test_ofdmframesync_callback_ends_frames(L, refr2):
    """two frames in one stream; the callback returns 1 after the last data symbol of each: the rows after it are dropped,
    the object goes back to acquisition inside the same execute() call and finds the second frame"""
    M, cp, taper, nsym, _ = [int(v) for v in refr2["osync_two_args"]]
    y, chunks, want = refr2["osync_two_y"], [int(c) for c in refr2["osync_two_chunks"]], refr2["osync_two_rows"]
    rows = []

    def cb(row, p):
        rows.append(row)
        return 1 if len(rows) % nsym == 0 else 0
    q = L.ofdmframesync(M, cp, taper, None, cb)
    pos = k = 0
    while pos < y.size:
        c = chunks[k % len(chunks)]
        k += 1
        q.execute(y[pos: pos + c])
        pos += c
    assert len(rows) == len(want) == 2 * nsym
    got = np.array(rows)
    assert np.abs(got - want).max() < 2e-4 * np.abs(want).max()
    assert not q.is_frame_open()
    q.destroy()
