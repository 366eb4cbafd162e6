"""Regenerate tests/golden/*.npz from the reference's bundled datasets.

Run once in the dev container (``/root/reference`` is not present on the GPU box):

    python tests/golden/make_fixtures.py

Sources (read-only): /root/reference/data/cell_lines_small.RData, cell_lines.rda,
pbmc_stim.RData (documented in /root/reference/R/data.R).  Covariates are integer-coded
exactly as ``as.factor()`` does in /root/reference/R/ui.R:210-221 (levels = sorted unique
strings), so level ``b`` here is row ``b`` of the reference's ``phi``.

The PBMC fixture follows the preprocessing of /root/reference/vignettes/Seurat.Rmd:86-99
(log-normalise, variable genes, scale, PCA) in plain numpy but keeps 50 PCs (BASELINE.json
config 2: "pbmc_stim, 50 PCs, 1 covariate 'stim', K=50").  The shipped file only holds a
2 x 1000-cell subsample of the Kang et al. data, so N = 2000 here.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rdata import r_factor_or_strings, r_list, read_rdata  # noqa: E402

REF = "/root/reference/data"


def cell_lines(fname, key, out):
    top = r_list(read_rdata(os.path.join(REF, fname))[key])
    meta = r_list(top["meta_data"])
    pcs = r_list(top["scaled_pcs"])
    Z = np.stack([np.asarray(pcs[f"X{j + 1}"].value, dtype=np.float64) for j in range(len(pcs))], axis=1)
    ds, ds_lv = r_factor_or_strings(meta["dataset"])
    ct, ct_lv = r_factor_or_strings(meta["cell_type"])
    np.savez_compressed(
        os.path.join(HERE, out), scaled_pcs=Z, dataset=ds, dataset_levels=np.array(ds_lv),
        cell_type=ct, cell_type_levels=np.array(ct_lv))
    print(out, Z.shape, ds_lv, np.bincount(ds), ct_lv, np.bincount(ct))


def pbmc(out, n_pcs=50, n_var=2000):
    d = read_rdata(os.path.join(REF, "pbmc_stim.RData"))
    mats = []
    for name in ("pbmc.ctrl", "pbmc.stim"):  # Seurat.Rmd:86 cbind(pbmc.stim, pbmc.ctrl); order is irrelevant
        a = d[name].attr
        nr, nc = (int(x) for x in a["Dim"].value)
        i, p, x = a["i"].value, a["p"].value, a["x"].value
        M = np.zeros((nr, nc), dtype=np.float64)
        for c in range(nc):
            M[i[p[c]:p[c + 1]], c] = x[p[c]:p[c + 1]]
        mats.append(M)
    X = np.concatenate(mats, axis=1)                       # genes x cells
    stim = np.concatenate([np.zeros(mats[0].shape[1]), np.ones(mats[1].shape[1])]).astype(np.int32)
    X = np.log1p(X / X.sum(axis=0, keepdims=True) * 1e4)   # NormalizeData
    mu, var = X.mean(axis=1), X.var(axis=1, ddof=1)
    disp = np.where(mu > 0, var / np.maximum(mu, 1e-12), 0.0)
    keep = np.argsort(-disp, kind="stable")[:n_var]        # simple dispersion ranking
    Xs = X[keep]
    Xs = (Xs - Xs.mean(axis=1, keepdims=True)) / np.maximum(Xs.std(axis=1, ddof=1, keepdims=True), 1e-12)
    Xs = np.clip(Xs, -10, 10)                              # ScaleData clips at 10
    U, S, Vt = np.linalg.svd(Xs, full_matrices=False)
    Z = (Vt[:n_pcs].T * S[:n_pcs])                         # cells x PCs
    np.savez_compressed(os.path.join(HERE, out), pcs=Z.astype(np.float64), stim=stim,
                        stim_levels=np.array(["ctrl", "stim"]))
    print(out, Z.shape, np.bincount(stim))


if __name__ == "__main__":
    cell_lines("cell_lines_small.RData", "cell_lines_small", "cell_lines_small.npz")
    cell_lines("cell_lines.rda", "cell_lines", "cell_lines.npz")
    pbmc("pbmc_stim_pcs.npz")
