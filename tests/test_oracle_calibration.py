"""HMMER-free known-answer test of the oracle's profile parsing, configuration, MSV byte arithmetic and
Forward recurrence.

Every profile in ITSx_db carries hmmbuild's own calibration of exactly the scoring systems the search
uses: `STATS LOCAL MSV mu lambda`, `STATS LOCAL FORWARD tau lambda`.  hmmbuild derives them as
  lambda = ln 2 + 1.44 / (M * H)        H = mean match-state relative entropy (bits) vs f = 0.25
  MSV mu    = ML Gumbel location at fixed lambda over random sequences of L = 200
  Forward tau: complete-Gumbel ML fit over random sequences of L = 100, tail mass 0.04
Re-deriving them from the oracle must land on the file's values (lambda exactly, mu/tau within the
sampling noise of hmmbuild's own N = 200 simulation, ~0.3-0.5 bit); an error in tbm/tjb, the -3 nat
correction, the entry distribution or the length model would show up as a multi-bit offset.
"""
import os

import numpy as np
import pytest

from conftest import HMM_DIR

LN2 = 0.69314718055994529


@pytest.fixture(scope="module")
def db(oracle):
    return oracle.ProfileDB([os.path.join(HMM_DIR, "A.hmm"), os.path.join(HMM_DIR, "M.hmm")], None)


def _entropy_bits(mat):
    p = mat[1:].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(p > 0, p * np.log2(p / 0.25), 0.0)
    return t.sum(1).mean()


@pytest.mark.parametrize("p", [0, 1, 2, 50, 150, 300])
def test_lambda_exact(db, p):
    if p >= db.n:
        pytest.skip("profile index beyond the loaded set")
    mat, t, compo = db.raw(p)
    M = db.M[p]
    lam = LN2 + 1.44 / (M * _entropy_bits(mat))
    ev = db.evparam(p)
    assert abs(lam - ev[1]) < 2e-5          # MSV lambda
    assert abs(lam - ev[5]) < 2e-5          # Forward lambda


def _random_codes(rng, n, L):
    return rng.integers(0, 4, size=(n, L), dtype=np.uint8)


@pytest.mark.parametrize("p", [0, 2, 200])
def test_msv_mu(db, oracle, p):
    rng = np.random.default_rng(100 + p)
    ev = db.evparam(p)
    lam = float(ev[1])
    L = 200
    nullsc = oracle.lib().ora_nullsc(L)
    xs = []
    for dsq in _random_codes(rng, 3000, L):
        sc, ovf = db.msv_score(p, np.ascontiguousarray(dsq))
        assert not ovf
        xs.append((sc - nullsc) / LN2)
    xs = np.array(xs, np.float64)
    mu = -np.log(np.mean(np.exp(-lam * xs))) / lam
    assert abs(mu - ev[0]) < 0.5, (mu, ev[0])


@pytest.mark.parametrize("p", [0, 2, 200])
def test_forward_tau(db, oracle, p):
    rng = np.random.default_rng(200 + p)
    ev = db.evparam(p)
    lam = float(ev[5])
    L = 100
    nullsc = oracle.lib().ora_nullsc(L)
    xs = np.array([(db.forward_score(p, np.ascontiguousarray(d)) - nullsc) / LN2
                   for d in _random_codes(rng, 3000, L)], np.float64)
    # complete-data Gumbel ML fit (Newton on lambda), as esl_gumbel_FitComplete
    lg = np.pi / np.sqrt(6 * xs.var())
    for _ in range(100):
        e = np.exp(-lg * xs)
        fx = 1.0 / lg - xs.mean() + (xs * e).sum() / e.sum()
        dfx = ((xs * e).sum() / e.sum()) ** 2 - (xs * xs * e).sum() / e.sum() - 1.0 / lg ** 2
        step = fx / dfx
        lg -= step
        if abs(step) < 1e-9:
            break
    mug = -np.log(np.mean(np.exp(-lg * xs))) / lg
    tailp = 0.04
    tau = mug - np.log(-np.log(1 - tailp)) / lg + np.log(tailp) / lam
    assert abs(tau - ev[4]) < 0.6, (tau, ev[4])


@pytest.mark.parametrize("p", [0, 2, 200])
def test_viterbi_filter_mu(db, oracle, p):
    """`STATS LOCAL VITERBI mu lambda` is hmmbuild's calibration of the 16-bit Viterbi filter itself (ML Gumbel
    location at fixed lambda, random sequences of L = 200): the oracle's restatement of that filter must land on it.
    (The stage never runs on the reference's path, F1 == F2; this pins the word profile for the day it does.)"""
    rng = np.random.default_rng(300 + p)
    ev = db.evparam(p)
    lam = float(ev[3])
    assert abs(lam - float(ev[1])) < 1e-6                       # Viterbi and MSV share lambda
    L = 200
    nullsc = oracle.lib().ora_nullsc(L)
    xs = []
    for dsq in _random_codes(rng, 3000, L):
        sc, ovf = db.viterbi_filter(p, np.ascontiguousarray(dsq))
        assert not ovf
        xs.append((sc - nullsc) / LN2)
    xs = np.array(xs, np.float64)
    mu = -np.log(np.mean(np.exp(-lam * xs))) / lam
    assert abs(mu - ev[2]) < 0.5, (mu, ev[2])


def test_viterbi_filter_below_forward(db):
    """The best path scores no more than the sum over paths (up to the filter's 1/500-bit quantisation and its flat
    3-nat length correction), and approaches it on a sequence that carries a strong match."""
    import synth
    rng = np.random.default_rng(12)
    for p in (0, 5, 120):
        for L in (60, 200, 420):
            dsq = np.ascontiguousarray(rng.integers(0, 4, size=L, dtype=np.uint8))
            v, ovf = db.viterbi_filter(p, dsq)
            assert not ovf and v <= db.forward_score(p, dsq) + 0.15, (p, L)
    seq, off, which, cfg = synth.make_config("c2_small", scale=0.05)
    m = os.path.join(synth.HMM_DIR, cfg["hmm_file"])
    from oracle import oracle as O
    d2 = O.ProfileDB([m], [cfg["left_prefix"]])
    close = 0
    for r in range(40):
        dsq = O.digitize(seq[off[r]:off[r + 1]].tobytes())
        best = max(range(d2.n), key=lambda q: d2.forward_score(q, dsq))
        f = d2.forward_score(best, dsq)
        v, ovf = d2.viterbi_filter(best, dsq)
        assert ovf or v <= f + 0.15
        close += (not ovf) and f - v < 6.0 and f > 10.0
    assert close >= 20


def test_forward_equals_backward(db):
    rng = np.random.default_rng(9)
    for p in (0, 5, 120):
        for L in (30, 180, 400):
            dsq = np.ascontiguousarray(rng.integers(0, 4, size=L, dtype=np.uint8))
            f, b = db.forward_score(p, dsq), db.backward_score(p, dsq)
            assert abs(f - b) < 2e-3 * max(1.0, abs(f)), (p, L, f, b)


def test_short_models_present(oracle):
    """G.hmm holds the two profiles with M < 45 (25 and 11 match states)."""
    g = oracle.ProfileDB([os.path.join(HMM_DIR, "G.hmm")], ["1_"])
    assert sorted(set(g.M)) == [11, 25, 45]


def test_msv_profile_bytes(db):
    cost, sc = db.msv_profile(0)
    assert sc["base"] == 190 and sc["tec"] == 3          # byteify(ln 0.5) = round(3) third-bits
    assert cost.shape == (db.M[0] + 1, 16)
    assert 0 < sc["bias"] < 60
    assert sc["tbm"] == int(round(-(3.0 / LN2) * np.log(2.0 / (45 * 46))))


# ---- multidomain resolver (stochastic-traceback ensemble + clustering) ---------------------------------------
def test_rng_leapfrog(oracle):
    """x <- 69069 x + 1 (mod 2^32); the closed-form jump the trace substreams use equals stepping."""
    x0 = oracle.rng_state0(42)
    assert 0 < x0 < 2 ** 32
    x = x0
    for k in range(1, 2000):
        x = (x * 69069 + 1) & 0xFFFFFFFF
        if k in (1, 2, 3, 17, 255, 256, 1023, 1999):
            assert oracle.rng_jump(x0, k) == x
    a = oracle.rng_jump(x0, 1 << 20)
    assert oracle.rng_jump(a, 1 << 20) == oracle.rng_jump(x0, 2 << 20)
    assert oracle.rng_jump(x0, 0) == x0


def test_multidomain_resolution(oracle):
    """A flagged region is replaced by its cluster envelopes: they lie inside the region, are ordered by start,
    carry the flag, and the run is deterministic; with the switch off the region stays one envelope."""
    import synth
    seq, off, which, cfg = synth.make_config("c2_small", scale=0.05)
    paths = [os.path.join(synth.HMM_DIR, cfg["hmm_file"])]
    db = oracle.ProfileDB(paths, [cfg["left_prefix"], cfg["right_prefix"]])
    found = 0
    for r in range(len(off) - 1):
        dsq = oracle.digitize(seq[off[r]:off[r + 1]].tobytes())
        for p in range(db.n):
            pr0, d0 = db.pair_run(p, dsq, oracle.default_params(1, 0))
            if not pr0.nmultidomain:
                continue
            pr1, d1 = db.pair_run(p, dsq, oracle.default_params(1, 1))
            pr2, d2 = db.pair_run(p, dsq, oracle.default_params(1, 1))
            assert [tuple(x) for x in d1[["ienv", "jenv"]].tolist()] == [tuple(x) for x in d2[["ienv", "jenv"]].tolist()]
            assert pr1.fwdsc == pr0.fwdsc and pr1.nmultidomain == pr0.nmultidomain
            regions = [(int(d["ienv"]), int(d["jenv"])) for d in d0 if d["is_multidomain"]]
            simple0 = [(int(d["ienv"]), int(d["jenv"])) for d in d0 if not d["is_multidomain"]]
            simple1 = [(int(d["ienv"]), int(d["jenv"])) for d in d1 if not d["is_multidomain"]]
            assert simple0 == simple1
            for d in d1:
                if d["is_multidomain"]:
                    assert any(a <= d["ienv"] <= d["jenv"] <= b for a, b in regions)
            starts = [int(d["ienv"]) for d in d1]
            assert starts == sorted(starts)
            found += 1
        if found >= 12:
            break
    assert found >= 12
