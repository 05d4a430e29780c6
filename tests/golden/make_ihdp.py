"""Regenerates tests/golden/ihdp_covariates.npz: the real IHDP covariates (747 x 25), the treatment indicator
(139 treated) and per-metric summaries of the stored 100-replicate `results` matrix, decoded from the
reference's workspace file test/H11_100_27_august.RData by causalstump_b200/rdata.py (no R needed).

    python tests/golden/make_ihdp.py      # needs /root/reference (build container only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from causalstump_b200 import rdata  # noqa: E402

SRC = "/root/reference/test/H11_100_27_august.RData"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ihdp_covariates.npz")


def build():
    ws = rdata.read_rdata(SRC)
    X, names = rdata.as_matrix(ws["X"])
    trt = np.asarray(ws["Trt"], dtype=np.float64)
    R, _ = rdata.as_matrix(ws["results"])   # 220 metric rows x 100 replicates (update.results, sec5-3_Hill2011.R:104-152)
    assert X.shape == (747, 25) and trt.sum() == 139 and R.shape == (220, 100)
    # row r of `results` = metric r % 22 of method r // 22 (result.item is a 22 x 10 matrix, sec5-3_Hill2011.R:163-166)
    metrics = ["PEHE.all.train", "PEHE.tt.train", "RMSE.all.train", "RMSE.tt.train", "E.ate.train", "E.att.train", "SE.ate.train",
               "SE.att.train", "coverage.ite.train", "coverage.ate.train", "range.ate.train", "PEHE.all.test", "PEHE.tt.test",
               "RMSE.all.test", "RMSE.tt.test", "E.ate.test", "E.att.test", "SE.ate.test", "SE.att.test", "coverage.ite.test",
               "coverage.ate.test", "range.ate.test"]
    methods = ["BART", "CF", "VT-RF", "CF-RF", "GP", "BCF", "BART.PS", "GP.PS", "TP2.PS", "TP2000.PS"]
    gp = R[22 * methods.index("GP"):22 * (methods.index("GP") + 1)]     # 22 x 100: the stored GP arm of the reference
    np.savez_compressed(OUT, X=X, names=np.array(names), Trt=trt, metrics=np.array(metrics), methods=np.array(methods),
                        gp_results=gp, results_shape=np.array(R.shape))
    print("reference GP arm, mean over 100 replicates:", dict(zip(metrics, np.round(np.nanmean(gp, axis=1), 3))))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    build()
