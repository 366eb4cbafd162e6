"""CPU tests of the oracle (oracle/harmony_oracle.cpp).

The reference holds no golden vectors for this path (its tests pin invariants only),
so the oracle is pinned three ways: (1) the reference's own test invariants
(/root/reference/tests/testthat/test_integration.R, test_two_variable.R) on the reference's own
fixtures; (2) agreement with an independent float64 numpy restatement of the vignette's formulas;
(3) fp32 instance vs fp64 instance.
"""
import numpy as np
import pytest

from harmony_b200 import harmony_options, prepare_inputs
from helpers import load_cell_lines, make_perms, make_Y0, rel_l2, run_oracle, setup_args, synthetic
from numpy_restatement import NumpyHarmony


def _prep(Z, meta, vars_use, **kw):
    return prepare_inputs(Z, meta, vars_use, **kw)


def test_integration_invariants_cell_lines_small():
    # test_integration.R:5-26
    Z, meta = load_cell_lines(small=True)
    a = _prep(Z, meta, "dataset", theta=1, nclust=50, options=harmony_options(max_iter_cluster=10))
    o, _, _ = run_oracle(a, make_Y0(Z, 50, 1), 5)
    N, d, K = 300, 20, 50
    assert o.get("Y").T.shape == (d, K)
    assert o.get("Z_corr").T.shape == (d, N) and o.get("Z_orig").T.shape == (d, N)
    R = o.get("R")
    assert R.T.shape == (K, N)
    assert R.min() >= 0 and R.max() <= 1
    np.testing.assert_allclose(R.sum(axis=1), 1.0, atol=1e-5)
    assert np.all(np.isfinite(o.get("Z_corr")))


def _chi2(o):
    O, E = o.get("O"), o.get("E")
    return float((((O - E) ** 2) / E).sum())


def test_theta_decreases_chi2_one_covariate():
    # test_integration.R:29-41
    Z, meta = load_cell_lines(small=True)
    a0 = _prep(Z, meta, "dataset", theta=0, nclust=20)
    a1 = _prep(Z, meta, "dataset", theta=1, nclust=5)
    o0, _, _ = run_oracle(a0, make_Y0(Z, 20, 2), 2)
    o1, _, _ = run_oracle(a1, make_Y0(Z, 5, 2), 2)
    assert _chi2(o0) > _chi2(o1)


def test_two_variable_invariants_and_chi2():
    # test_two_variable.R:5-55 (the arma::inv multi-covariate branch)
    Z, meta = load_cell_lines(small=False)
    a = _prep(Z, meta, ["cell_type", "dataset"], theta=[1, 1], nclust=50,
              options=harmony_options(max_iter_cluster=10))
    o, _, _ = run_oracle(a, make_Y0(Z, 50, 3), 10)
    assert o.get("O").shape[0] == 5 and o.get("E").shape[0] == 5
    R = o.get("R")
    assert R.min() >= 0 and R.max() <= 1
    np.testing.assert_allclose(R.sum(axis=1), 1.0, atol=1e-5)
    assert np.all(np.isfinite(o.get("Z_corr")))
    lo = _prep(Z, meta, ["cell_type", "dataset"], theta=[0, 0], nclust=20)
    hi = _prep(Z, meta, ["cell_type", "dataset"], theta=[2, 2], nclust=20)
    Y0 = make_Y0(Z, 20, 4)
    olo, _, _ = run_oracle(lo, Y0, 2)
    ohi, _, _ = run_oracle(hi, Y0, 2)
    assert _chi2(olo) > _chi2(ohi)


def test_E_column_sums_equal_batch_sizes():
    # doc/detailedWalkthrough.html soft expectation: colSums(E) = N_b
    Z, meta = load_cell_lines(small=False)
    a = _prep(Z, meta, "dataset", nclust=5)
    o, _, _ = run_oracle(a, make_Y0(Z, 5, 5), 1)
    np.testing.assert_allclose(o.get("E").sum(axis=1), [846, 824, 700], rtol=1e-4)


@pytest.mark.parametrize("case", ["small_1cov", "full_2cov", "synth_3cov", "fixed_lambda"])
def test_oracle_matches_numpy_restatement(case):
    if case == "small_1cov":
        Z, meta = load_cell_lines(small=True)
        a = _prep(Z, meta, "dataset", nclust=5)
    elif case == "full_2cov":
        Z, meta = load_cell_lines(small=False)
        a = _prep(Z, meta, ["cell_type", "dataset"], nclust=20)
    elif case == "synth_3cov":
        Z, meta = synthetic(1500, 12, [3, 6, 2], seed=7)
        a = _prep(Z, meta, ["cov0", "cov1", "cov2"], nclust=12, theta=[2, 1, 0.5])
    else:
        Z, meta = load_cell_lines(small=False)
        a = _prep(Z, meta, ["dataset"], nclust=10, lambda_=1.0)
    K, T, N = a["K"], a["max_iter_kmeans"], Z.shape[0]
    Y0 = make_Y0(Z, K, 11)
    n_iter = 3
    perms = make_perms(N, n_iter * T, 5).reshape(n_iter, T, N)
    o64, _, _ = run_oracle(dict(a, epsilon_harmony=-np.inf), Y0, n_iter, double=True, perms=perms)
    o32, _, _ = run_oracle(dict(a, epsilon_harmony=-np.inf), Y0, n_iter, double=False, perms=perms)
    s = setup_args(a)
    ref = NumpyHarmony(s["Z"], s["phi_i"], s["B_vec"], s["sigma"], s["theta"], s["lambda_"], s["alpha"], T, K,
                       s["block_size"], s["batch_proportion_cutoff"])
    ref.init_cluster(Y0)
    for it in range(n_iter):
        ref.cluster(perms[it])
        ref.moe_correct_ridge()
    # fp64 oracle vs independent fp64 numpy: tight
    assert rel_l2(o64.get("Z_corr"), ref.Z_corr) < 1e-9
    assert np.abs(o64.get("R") - ref.R).max() < 1e-9
    assert np.abs(o64.get("Y") - ref.Y).max() < 1e-9
    assert np.abs(o64.get("O").T - ref.O).max() < 1e-7
    np.testing.assert_allclose(o64.trace("objective_kmeans"), ref.obj_kmeans, rtol=1e-6)
    # fp32 oracle (the reference's arithmetic) vs truth
    assert rel_l2(o32.get("Z_corr"), ref.Z_corr) < 2e-4
    am32, am64 = o32.get("R").argmax(axis=1), ref.R.argmax(axis=1)
    assert (am32 != am64).mean() < 0.01


def test_small_n_guards():
    # harmony.cpp:83-91
    from oracle.oracle import OracleHarmony
    Z, meta = synthetic(5, 4, [2], seed=1)
    a = prepare_inputs(Z, meta, "cov0", nclust=2)
    o = OracleHarmony()
    with pytest.raises(RuntimeError, match="less than 6 cells"):
        o.setup(**setup_args(a))
    Z, meta = synthetic(30, 4, [2], seed=1)
    a = prepare_inputs(Z, meta, "cov0", nclust=2)
    o = OracleHarmony()
    o.setup(**setup_args(a))
    assert o.L.ho_warned_small(o.h) == 1
    o.init_cluster_cpp(make_Y0(Z, 2, 0))
    o.cluster_cpp(make_perms(30, 4, 0))     # block_size 0.2 -> 5 blocks of 6
    o.moe_correct_ridge_cpp()
    assert np.all(np.isfinite(o.get("Z_corr")))


def test_convergence_semantics():
    # harmony.cpp:190-200: the numerator is signed
    Z, meta = load_cell_lines(small=True)
    a = _prep(Z, meta, "dataset", nclust=5)
    o, _, iters = run_oracle(a, make_Y0(Z, 5, 1), 10)
    oh = o.trace("objective_harmony")
    assert len(oh) == iters + 1
    assert len(o.trace("kmeans_rounds")) == iters
    assert len(o.trace("objective_kmeans")) == 1 + 4 * iters
    last = (oh[-2] - oh[-1]) / abs(oh[-2])
    assert (last < 1e-2) == (iters < 10 or o.check_convergence(1))


def _separated(N=900, d=8, seed=3):
    """Batches that each hold a cell type of their own (plus one shared type): many (cluster, level) pairs fall
    below batch_proportion_cutoff, so the level-filter / subset / skip branches of harmony.cpp:358-547 all fire."""
    rng = np.random.default_rng(seed)
    batch = rng.integers(0, 3, N)
    own = rng.random(N) < 0.7
    ctype = np.where(own, batch, 3)
    M = rng.standard_normal((4, d)) * 3.0
    Z = M[ctype] + 0.3 * rng.standard_normal((N, d)) + 0.4 * rng.standard_normal((3, d))[batch]
    donor = batch * 2 + rng.integers(0, 2, N)
    return Z, {"batch": batch, "donor": donor}


@pytest.mark.parametrize("vars_use", [["batch"], ["batch", "donor"]])
def test_level_filter_branches_match_numpy(vars_use):
    Z, meta = _separated()
    a = _prep(Z, meta, vars_use, nclust=12)
    K, T, N = a["K"], a["max_iter_kmeans"], Z.shape[0]
    Y0 = make_Y0(Z, K, 5)
    perms = make_perms(N, 2 * T, 9).reshape(2, T, N)
    o64, _, _ = run_oracle(dict(a, epsilon_harmony=-np.inf), Y0, 2, double=True, perms=perms)
    s = setup_args(a)
    ref = NumpyHarmony(s["Z"], s["phi_i"], s["B_vec"], s["sigma"], s["theta"], s["lambda_"], s["alpha"], T, K,
                       s["block_size"], s["batch_proportion_cutoff"])
    ref.init_cluster(Y0)
    for it in range(2):
        ref.cluster(perms[it])
        ref.moe_correct_ridge()
    # the filter really fires: some (cluster, level) pairs are below the cutoff, some clusters keep every level
    avg = ref.O / ref.N_b[None, :]
    assert (avg <= 1e-5).any() and (avg > 1e-5).all(axis=1).any()
    assert rel_l2(o64.get("Z_corr"), ref.Z_corr) < 1e-9
    assert np.abs(o64.get("Y") - ref.Y).max() < 1e-9


@pytest.mark.parametrize("N,block_size", [(41, 0.05), (1000, 1.0), (1003, 0.01), (977, 0.3)])
def test_block_geometry_edge_cases(N, block_size):
    """n_blocks = my_ceil(1/block_size), cells_per_block = unsigned(N*block_size) in float, last block takes the
    remainder (harmony.cpp:279-300): one block, 100 blocks, a remainder larger than a block."""
    Z, meta = synthetic(N, 6, [3], seed=N)
    a = _prep(Z, meta, "cov0", nclust=4, options=harmony_options(block_size=block_size))
    T = a["max_iter_kmeans"]
    Y0 = make_Y0(Z, 4, 1)
    perms = make_perms(N, T, 2).reshape(1, T, N)
    o64, _, _ = run_oracle(dict(a, epsilon_harmony=-np.inf), Y0, 1, double=True, perms=perms)
    s = setup_args(a)
    ref = NumpyHarmony(s["Z"], s["phi_i"], s["B_vec"], s["sigma"], s["theta"], s["lambda_"], s["alpha"], T, 4,
                       s["block_size"], s["batch_proportion_cutoff"])
    ref.init_cluster(Y0)
    ref.cluster(perms[0])
    ref.moe_correct_ridge()
    assert np.abs(o64.get("R") - ref.R).max() < 1e-9
    assert rel_l2(o64.get("Z_corr"), ref.Z_corr) < 1e-9
    np.testing.assert_allclose(o64.get("R").sum(axis=1), 1.0, atol=1e-12)


@pytest.mark.parametrize("double", [False, True])
def test_oracle_reproduces_the_tables_printed_by_the_reference_vignette(double):
    """Golden values from the reference itself: the rendered vignette doc/detailedWalkthrough.html (:656-708) prints
    round(O), round(E) and the cluster x cell-type counts left by init_cluster_cpp on data(cell_lines) with
    nclust = 5.  From the matching k-means centroids (tests/golden/make_vignette_fixture.py) the oracle must
    reproduce all 30 integers of O and E; the cell-type counts (10 integers) agree to +-1 — the vignette's own
    table is off by one against its O row sums (454 vs 158 + 0 + 295), its 10-iteration arma::kmeans was not
    fully converged."""
    from helpers import GOLDEN
    from oracle.oracle import OracleHarmony
    import os
    g = np.load(os.path.join(GOLDEN, "vignette_walkthrough.npz"))
    Z, meta = load_cell_lines(small=False)
    a = prepare_inputs(Z, meta, "dataset", nclust=5, theta=1.0)
    assert np.allclose(a["sigma"], float(g["sigma"]))
    o = OracleHarmony(double=double)
    o.setup(**setup_args(a))
    o.init_cluster_cpp(g["Y"])
    O = np.asarray(o.get("O")).T      # K x B like harmonyObj$O
    E = np.asarray(o.get("E")).T
    assert np.array_equal(np.round(O), g["O_init"])
    assert np.array_equal(np.round(E), g["E_init"])
    R = np.asarray(o.get("R"))
    R = R if R.shape[0] == Z.shape[0] else R.T
    ct = np.stack([(meta["cell_type"] == lv) for lv in ("jurkat", "t293")], axis=1).astype(np.float64)
    assert np.abs(np.round(R.T @ ct) - g["celltype_init"]).max() <= 1


def _vignette_clustered_state(impl, double, seed):
    """init_cluster_cpp from the vignette's centroids, max_iter_kmeans <- 10, cluster_cpp() with the centroid step
    of harmony.cpp:235-238 switched on.  Returns O (K x B), R (N x K), number of rounds."""
    import os
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "vignette_walkthrough.npz"))
    Z, meta = load_cell_lines(small=False)
    a = prepare_inputs(Z, meta, "dataset", nclust=5, theta=1.0)
    perms = make_perms(Z.shape[0], 10, seed)
    if impl == "numpy":
        h = NumpyHarmony(a["Z"], a["phi_i"], a["B_vec"], a["sigma"], a["theta"], a["lambda_"], a["alpha"], 10, a["K"],
                         a["block_size"], a["batch_proportion_cutoff"])
        h.legacy_centroid_step = True
        h.init_cluster(g["Y"])
        h.cluster(perms)
        return g, meta, h.O, h.R, len(h.obj_kmeans) - 1
    from oracle.oracle import OracleHarmony
    o = OracleHarmony(double=double)
    args = setup_args(a)
    args["max_iter_kmeans"] = 10
    o.setup(**args)
    o.set_legacy_centroid_step(True)
    o.init_cluster_cpp(g["Y"])
    o.cluster_cpp(perms)
    R = np.asarray(o.get("R"))
    return g, meta, np.asarray(o.get("O")).T, (R if R.shape[0] == Z.shape[0] else R.T), int(o.trace("kmeans_rounds")[-1])


@pytest.mark.parametrize("impl,double", [("numpy", True), ("oracle", False), ("oracle", True)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_restatements_reproduce_the_vignette_tables_after_cluster_cpp(impl, double, seed):
    """Golden values for update_R + compute_objective + the convergence window: after `max_iter_kmeans <- 10;
    cluster_cpp()` the reference's vignette prints round(O) (doc/detailedWalkthrough.html:733-739), the cluster x
    cell-type counts (:769-775) and the per-cluster error rates (:786).  It was rendered while STEP 1 of
    harmony.cpp:235-238 (centroid update) was still active; with that step on, both restatements give exactly the
    25 printed integers for any update order — and only because the loop stops after 5 rounds like the reference
    (10 rounds would be off by one)."""
    g, meta, O, R, rounds = _vignette_clustered_state(impl, double, seed)
    assert rounds == 5
    assert np.array_equal(np.round(O), g["O_clustered"])
    ct = np.stack([(meta["cell_type"] == lv) for lv in ("jurkat", "t293")], axis=1).astype(np.float64)
    counts = R.T @ ct
    assert np.array_equal(np.round(counts), g["celltype_clustered"])
    err = (counts / counts.sum(axis=1, keepdims=True)).min(axis=1) * 100.0
    assert np.abs(err - g["error_rate_clustered"]).max() < (6e-4 if double else 2e-3)   # printed with 3 decimals
    # round((E / O)^theta, 2) of the same state (:844-850): 15 values from 0.35 to 9.4e8, i.e. down to the 1e-7
    # tails of exp(-dist / sigma); they move by ~1 % with the last digits of the centroids, which are not R's
    E = np.outer(R.sum(axis=0), np.bincount(np.unique(meta["dataset"], return_inverse=True)[1]) / R.shape[0])
    assert np.abs(E / O / g["e_over_o_clustered"] - 1.0).max() < 0.03
