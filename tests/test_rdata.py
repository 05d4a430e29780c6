"""causalstump_b200/rdata.py: the RDX2 / XDR workspace reader.  The round trip below builds a
workspace with a minimal writer so it runs anywhere; the reference's own files are read when /root/reference exists."""
import glob
import gzip
import os
import struct

import numpy as np
import pytest

from causalstump_b200 import hill, rdata

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ihdp_covariates.npz")


# ---- a minimal RDX2 writer (test only) -------------------------------------------------------------
def _i(v):
    return struct.pack(">i", v)


def _chars(s):
    b = s.encode()
    return _i(0x40009) + _i(len(b)) + b


def _sym(name):
    return _i(1) + _chars(name)


def _strs(xs, attr=b""):
    return _i(16 | (0x200 if attr else 0)) + _i(len(xs)) + b"".join(_chars(x) for x in xs) + attr


def _reals(xs, attr=b""):
    xs = np.asarray(xs, dtype=">f8")
    return _i(14 | (0x200 if attr else 0)) + _i(len(xs)) + xs.tobytes() + attr


def _ints(xs, attr=b""):
    xs = np.asarray(xs, dtype=">i4")
    return _i(13 | (0x200 if attr else 0)) + _i(len(xs)) + xs.tobytes() + attr


def _pairlist(items):  # [(name, payload)] -> tagged LISTSXP chain terminated by NILVALUE
    out = b""
    for name, payload in items:
        out += _i(2 | 0x400) + _sym(name) + payload
    return out + _i(254)


def _workspace():
    X = np.arange(12, dtype=np.float64).reshape(4, 3, order="F") / 7.0
    df_cols = [_reals([0.5, -1.25, 3.0]), _ints([1, rdata.NA_INTEGER, 2])]
    df = (_i(19 | 0x200 | 0x100) + _i(2) + b"".join(df_cols) +
          _pairlist([("names", _strs(["bw", "first"])), ("class", _strs(["data.frame"])), ("row.names", _ints([rdata.NA_INTEGER, -3]))]))
    mat = _reals(X.ravel(order="F"), _pairlist([("dim", _ints([4, 3])),
                                                ("dimnames", _i(19) + _i(2) + _i(254) + _strs(["a", "b", "c"]))]))
    body = _pairlist([("p", _reals([25.0])), ("Trt", _ints([1, 0, 0, 1])), ("covs", _strs(["bw", "b.head"])), ("X", df), ("M", mat)])
    raw = b"RDX2\nX\n" + _i(2) + _i(0x030400) + _i(0x020300) + body
    return raw, X


def test_round_trip_of_a_synthetic_workspace():
    raw, X = _workspace()
    for blob in (raw, gzip.compress(raw)):
        ws = rdata.read_rdata(blob)
        assert list(ws) == ["p", "Trt", "covs", "X", "M"]
        assert ws["p"][0] == 25.0 and list(ws["Trt"]) == [1, 0, 0, 1] and ws["covs"].value == ["bw", "b.head"]
        df, names = rdata.as_matrix(ws["X"])
        assert names == ["bw", "first"] and df.shape == (3, 2) and df[0, 0] == 0.5 and np.isnan(df[1, 1]) and df[2, 1] == 2.0
        M, cn = rdata.as_matrix(ws["M"])
        assert cn == ["a", "b", "c"] and np.array_equal(M, X)
    with pytest.raises(ValueError):
        rdata.read_rdata(b"RDA2\nA\n")


def test_ihdp_fixture_is_consistent_with_the_synthetic_generator():
    G = np.load(GOLD)
    X, z = G["X"], G["Trt"]
    assert X.shape == (747, 25) and z.sum() == 139 and list(G["names"][:3]) == ["bw", "b.head", "preterm"]
    assert np.all(np.abs(X[:, :6].mean(axis=0)) < 1e-12)                      # standardised continuous columns
    means = X[:, 6:].mean(axis=0) - np.where(np.arange(19) == 7, 1.0, 0.0)    # `first` is coded {1,2}
    assert np.max(np.abs(means - np.array(hill._BIN_MEANS))) < 0.006           # marginals hill.hill_covariates draws from
    gp = dict(zip(G["metrics"], np.nanmean(G["gp_results"], axis=1)))
    assert abs(gp["PEHE.all.train"] - 1.399) < 1e-3 and abs(gp["RMSE.all.train"] - 0.877) < 1e-3   # SURVEY.md section 6 band
    d = hill.hill_surface_b(replicate=1, X=X, z=z)
    assert d["X"].shape == (747, 25) and d["z"].sum() == 139 and np.isfinite(d["y"]).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="reference workspace files not present")
def test_reads_every_reference_workspace_file():
    files = sorted(glob.glob("/root/reference/test/*.RData")) + sorted(glob.glob("/root/reference/*.RData"))
    assert len(files) == 5
    G = np.load(GOLD)
    for f in files:
        ws = rdata.read_rdata(f)
        X, names = rdata.as_matrix(ws["X"])
        assert X.shape == (747, 25) and np.asarray(ws["Trt"]).sum() == 139
        R, _ = rdata.as_matrix(ws["results"])
        assert R.shape[1] == 100 and R.shape[0] >= 22   # metrics x methods per replicate (the scripts store different arms)
        assert ws["update.results"].kind == "closure"
    ws = rdata.read_rdata(files[0])
    assert np.array_equal(rdata.as_matrix(ws["X"])[0], G["X"]) and list(rdata.as_matrix(ws["X"])[1]) == list(G["names"])


def test_update_results_mirrors_the_reference_metric_block():
    """replicates.update_results = test/sec5-3_Hill2011.R:104-152, including the `[z.train=1]` quirk."""
    from causalstump_b200 import replicates as rp
    rng = np.random.default_rng(0)
    n = 60
    y1, y0 = rng.normal(size=n) + 4.0, rng.normal(size=n)
    z = (rng.random(n) < 0.3).astype(float)
    y = np.where(z == 1, y1, y0)
    tau = y1 - y0
    th = tau + np.linspace(-0.5, 0.5, n)
    ci = np.column_stack([th - 0.3, th + 0.3])
    test = np.array([3, 7, 11, 50])
    m = rp.update_results(test, y + 0.2, th, ci, (3.0, 5.5), (3.9, 4.0), y, z, y1, y0)
    assert tuple(m) != () and set(m) == set(rp.METRICS)
    tr = np.setdiff1d(np.arange(n), test)
    assert abs(m["PEHE.all.train"] - np.sqrt(np.mean((tau - th)[tr] ** 2))) < 1e-15
    assert abs(m["PEHE.tt.train"] - abs((tau - th)[tr][0])) < 1e-15          # first element, not the treated
    assert abs(m["PEHE.tt.test"] - abs((tau - th)[test][0])) < 1e-15
    assert abs(m["RMSE.all.test"] - 0.2) < 1e-12 and abs(m["SE.ate.train"] - m["E.ate.train"] ** 2) < 1e-18
    assert abs(m["coverage.ite.train"] - np.mean(np.abs((tau - th)[tr]) < 0.3)) < 1e-15
    assert m["range.ate.train"] == 2.5 and m["coverage.ate.train"] in (0.0, 1.0)
    assert abs(m["E.att.test"] - (np.mean(tau[test][z[test] == 1]) - np.mean(th[test][z[test] == 1]))) < 1e-15 or np.isnan(m["E.att.test"])
