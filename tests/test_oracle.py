"""Pin the CPU oracle (oracle/liquid_oracle.c): against the reference's own golden vectors,
against committed outputs of the compiled reference, against the live compiled reference when
oracle/_ref is present, against float64 closed forms, and through the reference's autotest
properties.  CPU only."""
import numpy as np
import pytest

from conftest import TOL_MAX_ABS, TOL_REL_RMS, cplx, max_abs_norm, rel_rms_per_channel
import closed_form as CF


# ---------------------------------------------------------------- reference golden vectors
def test_kat_besseli0f(orc, kat):
    for z, want in kat["besseli0f"]["points"]:
        assert abs(orc.lib().orc_besseli0f(z) - want) < kat["besseli0f"]["tol"]


def test_kat_firdes_rrcos(orc, kat):
    k = kat["firdes_rrcos"]
    h = orc.firdes_rrcos(k["k"], k["m"], k["beta"])
    assert np.abs(h - np.array(k["h"])).max() < k["tol"]


def test_kat_dotprod_crcf(orc, kat):
    for case in kat["dotprod_crcf"]["cases"]:
        want = complex(*case["y"])
        for order in ("run", "run4", "sse"):
            got = orc.dotprod_crcf(case["h"], cplx(case["x"]), order)
            assert abs(got - want) < kat["dotprod_crcf"]["tol"]


@pytest.mark.parametrize("tag", ["h4x8", "h7x16", "h13x32", "h23x64"])
def test_kat_firfilt_crcf(orc, kat, tag):
    case = kat["firfilt_crcf"]["cases"][tag]
    y = orc.Firfilt(h=case["h"]).execute_block(cplx(case["x"]))
    assert np.abs(y - cplx(case["y"])).max() < kat["firfilt_crcf"]["tol"]


def test_kat_fft_all_sizes(orc, kat):
    for n, case in kat["fft"]["cases"].items():
        x, want = cplx(case["x"]), cplx(case["y"])
        y = orc.fft(x, orc.FFT_FORWARD)
        assert np.abs(y - want).max() < kat["fft"]["tol"], n
        z = orc.fft(y, orc.FFT_BACKWARD) / np.float32(int(n))
        assert np.abs(z - x).max() < kat["fft"]["tol"], n


def test_kat_firpfb(orc, kat):
    k = kat["firpfb"]
    q = orc.Firpfb(k["M"], h=k["h"])
    for v in k["x"]:
        q.push(v)
    for i, want in enumerate(k["y"]):
        got = q.execute(i)
        assert abs(got.real - want) < k["tol"] and abs(got.imag) < k["tol"]


def test_msb_index(orc):
    for x in [0, 1, 2, 3, 4, 7, 8, 14, 15, 16, 18, 1023, 1024, 2**31]:
        assert orc.lib().orc_msb_index(x) == x.bit_length()


# ---------------------------------------------------------------- committed reference outputs
def test_fixture_kaiser_taps(orc, refout):
    names = [k for k in refout if k.startswith("kaiser_") and not k.endswith("_args")]
    assert len(names) >= 6
    for name in names:
        n, fc, As = refout[name + "_args"]
        h = orc.firdes_kaiser(int(n), np.float32(fc), np.float32(As))
        # float32 design; the compiled reference contracts a few FMAs -> allow 4 ulp of the peak
        assert np.abs(h - refout[name]).max() <= 5e-7, name


def test_fixture_rrcos_taps(orc, refout):
    for name in [k for k in refout if k.startswith("rrcos_") and not k.endswith("_args")]:
        k, m, beta = refout[name + "_args"]
        h = orc.firdes_rrcos(int(k), int(m), np.float32(beta))
        assert np.abs(h - refout[name]).max() <= 1e-6, name


@pytest.mark.parametrize("M,m", [(8, 5), (16, 5), (64, 3), (6, 2), (256, 4)])
def test_fixture_pfbch2(orc, refout, M, m):
    for kind, typ in (("a", orc.ANALYZER), ("s", orc.SYNTHESIZER)):
        x = refout["pfbch2_%s_M%d_m%d_x" % (kind, M, m)]
        want = refout["pfbch2_%s_M%d_m%d_y" % (kind, M, m)]
        y = orc.Firpfbch2(typ, M, m, As=60.0).execute_frames(x)
        if kind == "a":
            assert max_abs_norm(y, want, x) < 2e-6
            assert rel_rms_per_channel(y, want).max() < 2e-6
        else:
            assert np.abs(y - want).max() / np.abs(want).max() < 2e-6


@pytest.mark.parametrize("M,m", [(4, 3), (16, 4), (10, 2), (64, 5)])
def test_fixture_pfbch(orc, refout, M, m):
    x = refout["pfbch_a_M%d_m%d_x" % (M, m)]
    for kind, typ in (("a", orc.ANALYZER), ("s", orc.SYNTHESIZER)):
        want = refout["pfbch_%s_M%d_m%d_y" % (kind, M, m)]
        y = orc.Firpfbch(typ, M, kaiser=(m, 60.0)).execute_frames(x)
        assert np.abs(y - want).max() / np.abs(want).max() < 2e-6


def test_fixture_pfbch_rrc(orc, refout):
    x = refout["pfbch_rrc_M8_m3_x"]
    for kind, typ in (("a", orc.ANALYZER), ("s", orc.SYNTHESIZER)):
        want = refout["pfbch_rrc_%s_M8_m3_y" % kind]
        y = orc.Firpfbch(typ, 8, rrc=(3, 0.4)).execute_frames(x)
        assert np.abs(y - want).max() / np.abs(want).max() < 2e-6


def test_fixture_firfilt(orc, refout):
    x = refout["firfilt_k64_x"]
    q = orc.Firfilt(kaiser=(64, 0.25, 60.0, 0.0))
    assert np.abs(q.execute_block(x) - refout["firfilt_k64_y"]).max() < 1e-6
    q.reset()
    q.set_scale(0.375)
    assert np.abs(q.execute_block(x) - refout["firfilt_k64_scale_y"]).max() < 1e-6


@pytest.mark.parametrize("n", [128, 256, 1024, 4096])
def test_fixture_fft(orc, refout, n):
    x = refout["fft_%d_x" % n]
    for d, key in ((orc.FFT_FORWARD, "fwd"), (orc.FFT_BACKWARD, "bwd")):
        want = refout["fft_%d_%s" % (n, key)]
        y = orc.fft(x, d)
        assert np.abs(y - want).max() / np.abs(want).max() < 2e-6


def test_fixture_survey_fingerprints(orc, refout):
    """SURVEY.md section 8c fingerprints of firpfbch2 analysis on the autotest LCG input"""
    y = orc.Firpfbch2(orc.ANALYZER, 16, 5, As=60.0).execute_frames(refout["pfbch2_a_M16_m5_lcg_x"])
    assert abs(y[0, 0] - (-7.94109947e-05)) < 1e-9
    assert abs(y[21, 1] - (1.93961821e-02 - 8.78997445e-02j)) < 1e-7
    assert abs(y[79, 2] - (-7.58994371e-02 - 8.95354375e-02j)) < 1e-7
    assert np.abs(y - refout["pfbch2_a_M16_m5_lcg_y"]).max() < 1e-7


# ---------------------------------------------------------------- live compiled reference
def _need_ref(orc):
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built here")


@pytest.mark.parametrize("M,m,nfr", [(32, 5, 64), (1024, 7, 40), (12, 3, 30)])
def test_live_ref_pfbch2_analysis(orc, M, m, nfr):
    _need_ref(orc)
    x = orc.synth_input(nfr * M // 2, seed=77 + M)
    y = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(x)
    want = orc.ref_firpfbch2_frames(orc.ANALYZER, M, m, 60.0, x)
    assert rel_rms_per_channel(y, want).max() < 2e-6
    assert max_abs_norm(y, want, x) < 1e-6


def test_live_ref_scalar_math(orc):
    _need_ref(orc)
    R, L = orc.ref(), orc.lib()
    for z in [0.05, 0.5, 1.0, 2.5, 5.65, 9.99, 10.0, 33.3]:
        assert abs(L.orc_lngammaf(z) - R.liquid_lngammaf(z)) <= 4e-6 * max(1.0, abs(R.liquid_lngammaf(z)))
        assert abs(L.orc_besseli0f(z) - R.liquid_besseli0f(z)) <= 4e-6 * R.liquid_besseli0f(z)
    for x in [0.0, 0.001, 0.0099, 0.01, 0.3, 1.0, 2.5, -7.25]:
        assert abs(L.orc_sincf(x) - R.sincf(x)) <= 2e-7
    for As in [10.0, 21.0, 35.0, 50.0, 60.0, 80.0, -60.0]:
        assert L.orc_kaiser_beta_As(As) == pytest.approx(R.kaiser_beta_As(As), rel=1e-6)


def test_live_ref_dotprod_orders(orc):
    _need_ref(orc)
    import ctypes as C
    R = orc.ref()
    rng = np.random.default_rng(5)
    for n in [1, 2, 3, 7, 14, 18, 31, 32, 33, 64, 100, 512]:
        h = rng.standard_normal(n).astype(np.float32)
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        want = np.zeros(1, np.complex64)
        q = C.c_void_p(R.dotprod_crcf_create(h, n))
        R.dotprod_crcf_execute(q, x, want)
        R.dotprod_crcf_destroy(q)
        got = orc.dotprod_crcf(h, x, "sse")
        assert abs(got - want[0]) <= 2e-6 * np.sqrt(n) * 4


# ---------------------------------------------------------------- float64 closed forms
def test_closed_form_pfbch2_analysis(orc):
    M, m = 16, 3
    h = orc.firdes_kaiser(2 * M * m + 1, 1.0 / M, 60.0)
    hs = (h.astype(np.float32) * np.float32(M) / np.float32(np.cumsum(h, dtype=np.float32)[-1]))
    x = orc.synth_input(30 * M // 2, seed=3)
    y = orc.Firpfbch2(orc.ANALYZER, M, m, h=hs[: 2 * M * m]).execute_frames(x)
    want = CF.pfbch2_analysis(hs, M, m, x)
    assert rel_rms_per_channel(y, want).max() < TOL_REL_RMS / 10
    assert max_abs_norm(y, want, x) < TOL_MAX_ABS / 10


def test_closed_form_pfbch2_synthesis(orc):
    M, m = 8, 2
    rng = np.random.default_rng(1)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    X = orc.synth_input(20 * M, seed=4)
    y = orc.Firpfbch2(orc.SYNTHESIZER, M, m, h=h).execute_frames(X).reshape(-1)
    want = CF.pfbch2_synthesis(h, M, m, X)
    assert np.abs(y - want).max() / np.abs(want).max() < 1e-5


def test_closed_form_pfbch(orc):
    M, p = 12, 5
    rng = np.random.default_rng(2)
    h = rng.standard_normal(M * p).astype(np.float32)
    x = orc.synth_input(25 * M, seed=5)
    ya = orc.Firpfbch(orc.ANALYZER, M, p=p, h=h).execute_frames(x)
    assert np.abs(ya - CF.pfbch_analysis(h, M, p, x)).max() / np.abs(ya).max() < 1e-5
    ys = orc.Firpfbch(orc.SYNTHESIZER, M, p=p, h=h).execute_frames(x).reshape(-1)
    assert np.abs(ys - CF.pfbch_synthesis(h, M, p, x)).max() / np.abs(ys).max() < 1e-5


def test_firfilt_is_convolution(orc):
    rng = np.random.default_rng(3)
    h = rng.standard_normal(23).astype(np.float32)
    x = orc.synth_input(500, seed=6)
    y = orc.Firfilt(h=h).execute_block(x)
    want = np.convolve(x.astype(np.complex128), h.astype(np.float64))[: x.size]
    assert np.abs(y - want).max() < 1e-5


# ---------------------------------------------------------------- reference autotest properties
@pytest.mark.parametrize("M", [8, 16, 32, 64])
def test_autotest_pfbch2_roundtrip(orc, M):
    """firpfbch2_crcf_autotest.c:27-99: analysis -> synthesis reconstructs x delayed by
    2Mm - M/2 + 1, tolerance 1e-3"""
    m, As = 5, 60.0
    n = M * 8 * m
    s, x = 1, np.empty(n, np.complex64)
    for i in range(n):
        s = (s * 524287) % 1031
        x[i] = np.float32(s) / np.float32(1031) - np.float32(0.5)
    Y = orc.Firpfbch2(orc.ANALYZER, M, m, As=As).execute_frames(x)
    y = orc.Firpfbch2(orc.SYNTHESIZER, M, m, As=As).execute_frames(Y).reshape(-1)
    delay = 2 * M * m - M // 2 + 1
    assert np.abs(y[:delay]).max() < 1e-3
    assert np.abs(y[delay:] - x[: n - delay]).max() < 1e-3


def test_autotest_pfbch_analyzer_equals_mixer_filter_decimate(orc):
    """firpfbch_crcf_analyzer_autotest.c:29-152: channel c == mix down by c/M, FIR h, keep
    every M-th; tolerance 1e-4"""
    M, p, S = 4, 5, 40
    rng = np.random.default_rng(9)
    h = (rng.integers(0, 4, M * p) - 1.5).astype(np.float32)
    x = (0.1 * np.sqrt(0.5) * ((rng.integers(0, 4, M * S) - 1.5) + 1j * (rng.integers(0, 4, M * S) - 1.5))).astype(np.complex64)
    Y0 = orc.Firpfbch(orc.ANALYZER, M, p=p, h=h).execute_frames(x)
    j = np.arange(M * S)
    for c in range(M):
        mixed = (x * np.exp(-2j * np.pi * c * j / M)).astype(np.complex64)
        y = orc.Firfilt(h=h).execute_block(mixed)
        assert np.abs(Y0[:, c] - y[M - 1::M]).max() < 1e-4


def test_state_carry_and_reset(orc):
    """k x execute == one block of k frames; reset returns to the zero state"""
    M, m = 16, 4
    x = orc.synth_input(40 * M // 2, seed=8)
    q = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0)
    whole = q.execute_frames(x)
    q.reset()
    parts = np.concatenate([q.execute_frames(x[: 7 * M // 2]), q.execute_frames(x[7 * M // 2:])])
    assert np.array_equal(whole, parts)


def test_synth_input_is_deterministic_and_bounded(orc):
    a = orc.synth_input(1000, seed=11)
    b = orc.synth_input(500, seed=11, first=500)
    assert np.array_equal(a[500:], b)
    assert a.real.min() >= -0.5 and a.real.max() < 0.5 and abs(a.real.mean()) < 0.05
    assert 0.07 < np.mean(np.abs(a) ** 2) < 0.25


# ------------------------------------------------------------------ root-Nyquist Kaiser prototypes
def _rn_cases(refrn):
    for name in sorted(k for k in refrn if k.startswith("args_")):
        k, m, beta, dt = refrn[name]
        yield name[5:], int(k), int(m), np.float32(beta), np.float32(dt)


def test_fixture_rkaiser_helpers(orc, refrn):
    for df, N, want in refrn["req_As"]:
        assert orc.lib().orc_estimate_req_filter_As(df, int(N)) == np.float32(want)
    for m, b, want in refrn["approx_rho"]:
        assert orc.lib().orc_rkaiser_approximate_rho(int(m), b) == np.float32(want)


def test_fixture_arkaiser_rkaiser_taps_bit_exact(orc, refrn):
    """the restatement equals the reference (baseline x86-64 build, no FMA contraction) bit for bit,
    including the rho found by the 14-round parabolic search of rkaiser"""
    for tag, k, m, beta, dt in _rn_cases(refrn):
        assert np.array_equal(orc.firdes_arkaiser(k, m, beta, dt), refrn["taps_arkaiser_" + tag]), tag
        assert np.array_equal(orc.firdes_rkaiser(k, m, beta, dt), refrn["taps_rkaiser_" + tag]), tag
        rms, peak = orc.filter_isi(refrn["taps_rkaiser_" + tag], k, m)
        assert np.allclose([rms, peak], refrn["isi_rkaiser_" + tag], rtol=1e-6)


def test_rkaiser_properties(orc):
    """what the reference's autotests check for its root-Nyquist designs (src/filter/tests/
    firdes_autotest.c: ISI of the matched pair below a threshold): energy k, small ISI, and rkaiser is
    no worse than arkaiser"""
    for k, m, beta in [(2, 9, 0.3), (4, 7, 0.25), (10, 5, 0.5)]:
        ha, hr = orc.firdes_arkaiser(k, m, beta), orc.firdes_rkaiser(k, m, beta)
        for h in (ha, hr):
            assert abs(float((h.astype(np.float64) ** 2).sum()) - k) < 1e-4 * k
        isi_a, isi_r = orc.filter_isi(ha, k, m)[0], orc.filter_isi(hr, k, m)[0]
        assert isi_r <= isi_a * 1.0001 and isi_r < 10 ** (-30 / 20)


def test_live_reference_rkaiser(orc):
    """against the compiled reference of THIS host (may contract FMAs): taps agree to float rounding;
    the rkaiser search is allowed the resulting 1e-5 wobble of its optimum"""
    if not orc.ref_available():
        pytest.skip("compiled reference not present")
    import ctypes as C
    R = orc.ref()
    for k, m, beta in [(2, 3, 0.3), (16, 7, 0.25), (64, 4, 0.9)]:
        for name, fn, tol in (("arkaiser", orc.firdes_arkaiser, 2e-6), ("rkaiser", orc.firdes_rkaiser, 5e-5)):
            want = np.empty(2 * k * m + 1, np.float32)
            getattr(R, "liquid_firdes_" + name)(k, m, beta, 0.0, want)
            assert np.abs(fn(k, m, np.float32(beta)) - want).max() <= tol * np.abs(want).max(), (name, k, m)


def test_fixture_pfbch_rnyquist_channelizers(orc, refrn):
    for name in sorted(k for k in refrn if k.startswith("pfbch_") and k.endswith("_args")):
        M, m, beta, ftype = refrn[name]
        M, m, ftype = int(M), int(m), int(ftype)
        stem = name[:-5]                                  # pfbch_<proto>_M<M>_m<m>
        proto = stem.split("_")[1]
        x = refrn[stem + "_x"]
        for typ, key in ((orc.ANALYZER, "a"), (orc.SYNTHESIZER, "s")):
            want = refrn["pfbch_%s_%s_M%d_m%d_y" % (proto, key, M, m)]
            got = orc.Firpfbch(typ, M, rnyquist=(m, np.float32(beta), ftype)).execute_frames(x)
            assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max(), (stem, key)


# ------------------------------------------------------------------ firpfbch_cccf
def _cccf_cases(refrn):
    for name in sorted(k for k in refrn if k.startswith("cccf_") and k.endswith("_h")):
        M, p = (int(v[1:]) for v in name.split("_")[1:3])
        yield M, p, refrn[name], refrn["cccf_M%d_p%d_x" % (M, p)]


def test_fixture_pfbch_cccf(orc, refrn):
    for M, p, h, x in _cccf_cases(refrn):
        for typ, key in ((orc.ANALYZER, "a"), (orc.SYNTHESIZER, "s")):
            want = refrn["cccf_M%d_p%d_%s_y" % (M, p, key)]
            got = orc.FirpfbchCccf(typ, M, p, h).execute_frames(x)
            assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max(), (M, p, key)


def test_pfbch_cccf_real_taps_equal_crcf(orc):
    """with real taps the cccf channelizer is the crcf one"""
    M, m = 16, 3
    h = orc.firdes_kaiser(2 * M * m + 1, 0.5 / M, 60.0)[: 2 * M * m]
    x = orc.synth_input(30 * M, seed=0xCC)
    for typ in (orc.ANALYZER, orc.SYNTHESIZER):
        a = orc.FirpfbchCccf(typ, M, 2 * m, h.astype(np.complex64)).execute_frames(x)
        b = orc.Firpfbch(typ, M, p=2 * m, h=h).execute_frames(x)
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()


# ------------------------------------------------------------------ spgramcf
def _spgram_cases(refrn):
    for name in sorted(k for k in refrn if k.startswith("spgram_") and k.endswith("_args")):
        nfft, wl, beta = refrn[name]
        yield name[:-5], int(nfft), int(wl), np.float32(beta)


def _spgram_run(q, x, nfft, wl):
    q.write(x[: wl + 5])
    X, psd = q.execute(), q.execute_psd()
    q.reset()
    q.accumulate_psd(x[: 2 * nfft + 3], 0.3)
    acc1 = q.write_accumulation()
    q.accumulate_psd(x[2 * nfft + 3:], 0.3)
    acc2 = q.write_accumulation()
    est = q.estimate_psd(x)
    return X, psd, acc1, acc2, est


def lin_close(got_db, want_db, rel=2e-6):
    """dB spectra compared as linear power relative to the strongest bin: a float32 FFT is accurate to
    ~1e-7 of the LARGEST component, so a bin 50 dB down is only good to ~1e-4 dB while its absolute error
    in power stays far below rel * peak"""
    g, w = 10.0 ** (np.asarray(got_db, np.float64) / 10), 10.0 ** (np.asarray(want_db, np.float64) / 10)
    return np.abs(g - w).max() <= rel * w.max()


def spgram_check(res, refrn, tag):
    X, psd, acc1, acc2, est = res
    assert np.abs(X - refrn[tag + "_X"]).max() <= 2e-6 * np.abs(refrn[tag + "_X"]).max()
    for got, key in ((psd, "_psd"), (acc1, "_acc1"), (acc2, "_acc2"), (est, "_est")):
        assert lin_close(got, refrn[tag + key]), (tag, key)


def test_fixture_spgramcf(orc, refrn):
    for tag, nfft, wl, beta in _spgram_cases(refrn):
        q = orc.Spgramcf(nfft, window_len=wl, beta=beta)
        spgram_check(_spgram_run(q, refrn[tag + "_x"], nfft, wl), refrn, tag)


def test_spgramcf_properties(orc):
    """default object = Kaiser(nfft/2, beta 10); the window scaling sqrt(2) / (rms(w) sqrt(nfft)) puts white
    input of power P at 2 P window_len / nfft per bin on average (= P for the default window); the first
    accumulate_psd call after a reset runs with alpha = 1 whatever was asked (spgram.c:291-292), so it
    leaves the LAST periodogram"""
    nfft = 128
    x = orc.synth_input(40 * nfft, seed=77)                   # uniform +-0.5 per component: power 1/6
    q = orc.Spgramcf(nfft)
    est = q.estimate_psd(x)
    mean_lin = np.mean(10 ** (est / 10))
    assert abs(mean_lin - 1.0 / 6.0) < 0.02
    q.reset()
    q.accumulate_psd(x[:5 * nfft], 0.1)
    first = q.write_accumulation()
    q2 = orc.Spgramcf(nfft)
    q2.write(x[:5 * nfft])                                    # same buffer contents as after the last transform
    assert np.abs(first - q2.execute_psd()).max() < 1e-3


# ------------------------------------------------------------------ firdecim / firinterp
def _mr_cases(refrn):
    for name in sorted(k for k in refrn if k.startswith("mr_") and k.endswith("_x")):
        M, m = (int(v[1:]) for v in name.split("_")[1:3])
        yield name[:-2], M, m


def test_fixture_firdecim_firinterp(orc, refrn):
    for tag, M, m in _mr_cases(refrn):
        x, h = refrn[tag + "_x"], refrn[tag + "_h"]
        assert np.abs(orc.Firdecim(M, kaiser=(m, 60.0)).execute_block(x) - refrn[tag + "_decim_kaiser"]).max() < 2e-6
        got = orc.Firdecim(M, h=h).execute_block(x)
        assert np.abs(got - refrn[tag + "_decim_h"]).max() < 2e-6 * np.abs(refrn[tag + "_decim_h"]).max()
        assert np.abs(orc.Firinterp(M, kaiser=(m, 60.0)).execute_block(x[:300]) - refrn[tag + "_interp_kaiser"]).max() < 2e-6
        got = orc.Firinterp(M, h=h).execute_block(x[:300])
        assert np.abs(got - refrn[tag + "_interp_h"]).max() < 2e-6 * np.abs(refrn[tag + "_interp_h"]).max()


def test_firdecim_is_subsampled_firfilt_and_firinterp_is_filtered_upsampling(orc):
    """the identities the two objects stand for: decimation = every M-th output of the full-rate filter;
    interpolation = the full-rate filter over the zero-stuffed input"""
    M, m = 4, 3
    h = orc.firdes_kaiser(2 * M * m + 1, 0.5 / M, 60.0)[: 2 * M * m]
    x = orc.synth_input(M * 100, seed=5)
    full = orc.Firfilt(h=h).execute_block(x)
    assert np.abs(orc.Firdecim(M, h=h).execute_block(x) - full[::M]).max() < 1e-6
    up = np.zeros(100 * M, np.complex64)
    up[::M] = x[:100]
    assert np.abs(orc.Firinterp(M, h=h).execute_block(x[:100]) - orc.Firfilt(h=h).execute_block(up)).max() < 1e-6


# ------------------------------------------------------------------ resamp2 / msresamp2
HB_MODES = ("analyzer", "synthesizer", "decim", "interp")


def _hb_cases(fx):
    for name in sorted(k for k in fx if k.startswith("hb_") and k.endswith("_args")):
        m, f0, As = fx[name]
        yield name[:-5], int(m), float(f0), float(As)


def _ms_cases(fx):
    for name in sorted(k for k in fx if k.startswith("ms_") and k.endswith("_args")):
        typ, S, fc, As, delay = fx[name]
        yield name[:-5], int(typ), int(S), float(fc), float(As), float(delay)


def hb_check(q, run_mixed, fx, tag, tol):
    """every mode of a half-band object (fresh state each) and the mixed-mode run against the fixture"""
    x = fx[tag + "_x"]
    y0, y1 = q.filter_execute_block(x) if hasattr(q, "filter_execute_block") else q.filter_execute(x)
    assert np.abs(np.concatenate([y0, y1]) - fx[tag + "_filter"]).max() < tol, (tag, "filter")
    for mode in HB_MODES:
        q.clear()
        fn = getattr(q, mode + "_execute_block", None) or getattr(q, mode + "_execute")
        assert np.abs(fn(x) - fx[tag + "_" + mode]).max() < tol, (tag, mode)
    q.clear()
    assert np.abs(run_mixed(q, x) - fx[tag + "_mixed"]).max() < tol, (tag, "mixed")


def test_fixture_resamp2(orc, refrn2):
    for tag, m, f0, As in _hb_cases(refrn2):
        hb_check(orc.Resamp2(m, f0, As), orc.run_resamp2_mixed, refrn2, tag, 3e-7)


def test_fixture_msresamp2_and_filter_len(orc, refrn2):
    for (df, As), want in zip(refrn2["len_points"], refrn2["len_values"]):
        assert orc.lib().orc_estimate_req_filter_len(float(df), float(As)) == want
    for tag, typ, S, fc, As, delay in _ms_cases(refrn2):
        q = orc.Msresamp2(typ, S, fc, 0.0, As)
        assert q.get_delay() == delay
        want = refrn2[tag + "_y"]
        assert np.abs(q.execute_block(refrn2[tag + "_x"]) - want).max() < 1e-6 * max(1.0, np.abs(want).max()), tag


def test_resamp2_properties(orc):
    """what the object stands for: the two bands of the filter mode add up to the delayed input; the decimator is
    every second (odd) output of the low band, times two; the interpolator is the low band of the zero-stuffed
    input, times two"""
    m = 6
    x = orc.synth_input(800, seed=77)
    q = orc.Resamp2(m, 0.0, 60.0)
    y0, y1 = q.filter_execute(x)
    assert np.abs((y0 + y1)[2 * m:] - x[:-2 * m]).max() < 1e-6
    q.clear()
    d = q.decim_execute(x)
    assert np.abs(d - 2 * y0[1::2]).max() < 2e-6
    q.clear()
    up = q.interp_execute(x[:300])               # interpolation = half-band filtering of the zero-stuffed input (x2)
    z = np.zeros(600, np.complex64)
    z[::2] = x[:300]
    q.clear()
    z0, _ = q.filter_execute(z)
    assert np.abs(up - 2 * z0).max() < 2e-6


# ------------------------------------------------------------------ fftfilt
def _ff_cases(fx):
    for name in sorted(k for k in fx if k.startswith("ff_") and k.endswith("_args")):
        h_len, n, scale = fx[name]
        yield name[:-5], int(n), (None if scale < 0 else float(scale))


def test_fixture_fftfilt(orc, refrn2):
    for tag, n, scale in _ff_cases(refrn2):
        q = orc.Fftfilt(refrn2[tag + "_h"], n)
        if scale is not None:
            q.set_scale(scale)
        want = refrn2[tag + "_y"]
        assert np.abs(q.execute_blocks(refrn2[tag + "_x"]) - want).max() < 1e-6 * np.abs(want).max(), tag


def test_fftfilt_is_firfilt(orc):
    """what the object stands for: the causal linear convolution, however the blocks are cut"""
    h = orc.firdes_kaiser(57, 0.2, 60.0)
    x = orc.synth_input(64 * 20, seed=31)
    want = orc.Firfilt(h=h).execute_block(x)
    for n in (56, 64, 80):
        got = orc.Fftfilt(h, n).execute_blocks(x[: len(x) // n * n])
        assert np.abs(got - want[: len(got)]).max() < 2e-6


# ------------------------------------------------------------------ asgramcf
def _asg_cases(fx):
    for name in sorted(k for k in fx if k.startswith("asg_") and k.endswith("_args")):
        nfft, offset, scale = fx[name]
        yield name[:-5], int(nfft), (None if np.isnan(offset) else float(offset)), (None if np.isnan(scale) else float(scale))


def asgram_check(q, fx, tag, offset, scale):
    """characters and peak of an asgramcf object against the fixture; a character may move by one level where
    the bin sits within float rounding of a threshold (at most 2 % of the bins)"""
    levels = " .,-+*&NM#"
    if offset is not None:
        q.set_scale(offset, scale)
    q.write(fx[tag + "_x"])
    got, pv, pf = q.execute()
    want = bytes(fx[tag + "_ascii"]).decode("ascii")
    assert len(got) == len(want)
    diff = [(a, b) for a, b in zip(got, want) if a != b]
    assert len(diff) <= max(1, len(want) // 50), (tag, got, want)
    assert all(abs(levels.index(a) - levels.index(b)) == 1 for a, b in diff), (tag, diff)
    assert abs(pv - fx[tag + "_peak"][0]) < 1e-3 and pf == fx[tag + "_peak"][1], tag


def test_fixture_asgramcf(orc, refrn2):
    for tag, nfft, offset, scale in _asg_cases(refrn2):
        asgram_check(orc.Asgramcf(nfft), refrn2, tag, offset, scale)


# ------------------------------------------------------------------ ofdmframegen
def ofdm_custom_alloc(M):
    p = np.full(M, 2, np.uint8)
    p[0] = p[M // 2] = 0
    p[1] = p[M - 1] = p[3] = 1
    return p


def _ofdm_cases(fx):
    for name in sorted(k for k in fx if k.startswith("ofdm_") and k.endswith("_args")):
        M, cp, tl, custom = (int(v) for v in fx[name])
        yield name[:-5], M, cp, tl, (ofdm_custom_alloc(M) if custom else None)


def test_fixture_ofdmframegen(orc, refrn2):
    for tag, M, cp, tl, p in _ofdm_cases(refrn2):
        want = refrn2[tag + "_frame"]
        got = orc.ofdm_frame(orc.Ofdmframegen(M, cp, tl, p), refrn2[tag + "_x"])
        assert got.shape == want.shape and np.abs(got - want).max() < 2e-6 * np.abs(want).max(), tag


def test_ofdmframegen_properties(orc):
    """what the generator stands for: the default allocation (guard band, DC null, pilot spacing); the body of a
    data symbol is the scaled IDFT of the mapped subcarriers with BPSK pilots; the cyclic prefix repeats the end of
    the body; consecutive symbols overlap-add to a constant-envelope taper (taper[i] + taper[tl-1-i] = 1)"""
    M, cp, tl = 64, 16, 8
    q = orc.Ofdmframegen(M, cp, tl)
    p = q.sctype()
    assert p[0] == 0 and np.all(p[M // 2 - 6:M // 2 + 7] == 0) and (p == 1).sum() == 6 and (p == 2).sum() == 44
    x = orc.synth_input(M * 3, seed=11)
    y = q.writesymbols(x).reshape(3, M + cp)
    g = 1.0 / np.sqrt((p != 0).sum())
    for s in range(3):
        X = np.fft.fft(y[s, cp:]) / M                      # the reference transform is unscaled
        assert np.abs(X[p == 2] - x[s * M:(s + 1) * M][p == 2] * g).max() < 1e-5
        assert np.allclose(np.abs(X[p == 1]), g, atol=1e-5) and np.abs(X[p == 1].imag).max() < 1e-5
        assert np.abs(X[p == 0]).max() < 1e-5
        assert np.abs(y[s, tl:cp] - y[s, M + tl:M + cp]).max() < 1e-6
    t = np.sin(np.pi / 2 * (np.arange(tl) + 0.5) / tl) ** 2
    assert np.abs(y[1, :tl] - (y[1, M:M + tl] * t + y[0, cp:cp + tl] * t[::-1])).max() < 1e-6
