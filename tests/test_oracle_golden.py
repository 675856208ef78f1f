"""Pins oracle/diffmm_oracle.py against vectors produced by the unmodified reference
(tests/golden/*.npz, written by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import diffmm_oracle as O
from conftest import load_golden, params_of


def _sched_from(args):
    ns, nmin, nmax, steps = args
    return O.make_schedule(ns, nmin, nmax, int(steps))


@pytest.mark.parametrize("name", ["tiktok", "sports", "wide"])
def test_schedule(name):
    g = load_golden("schedule")
    s = _sched_from(g[f"{name}.args"])
    for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
              "sqrt_one_minus_alphas_cumprod", "posterior_variance", "posterior_mean_coef1",
              "posterior_mean_coef2"):
        np.testing.assert_allclose(s[k], g[f"{name}.{k}"], rtol=1e-14, atol=0, err_msg=k)
    np.testing.assert_allclose(O.snr(s, np.arange(len(s["betas"]))), g[f"{name}.snr"], rtol=1e-13)


def test_schedule_tiktok_probe_constants():
    # coef1=[1,.9614,.4908,.3298,.2485]; w=[1,9612.5,189.4,64.8,32.7]
    s = O.make_schedule(0.5, 1e-4, 0.02, 5)
    np.testing.assert_allclose(s["posterior_mean_coef1"], [1, .9614, .4908, .3298, .2485], atol=5e-5)
    np.testing.assert_allclose(O.snr_weight(s, np.arange(5)), [1, 9612.5, 189.4, 64.8, 32.7], rtol=2e-3)


def test_denoise_forward():
    g = load_golden("denoise_forward")
    p = params_of(g)
    np.testing.assert_allclose(O.denoise_forward(p, g["x_t"], g["t"]), g["out_nogate"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(O.denoise_forward(p, g["x_t"], g["t"], g["feat"]), g["out_gate"], rtol=2e-5, atol=2e-6)


def test_q_sample():
    g = load_golden("q_sample")
    s = O.make_schedule(0.5, 1e-4, 0.02, 5)
    np.testing.assert_allclose(O.forward_cal_xt(s, g["x0"], g["t"], noise=g["noise"]), g["xt_explicit"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(O.forward_cal_xt(s, g["x0"], g["t"], randn=g["randn"]), g["xt_default"], rtol=1e-6, atol=1e-7)


def test_training_losses_value():
    g = load_golden("training_losses")
    s = O.make_schedule(0.5, 1e-4, 0.02, 5)
    got = O.training_losses(s, params_of(g), g["x0"], g["i_embs"], g["feat"], g["t"], g["noise"],
                            float(g["reg"]), float(g["sim_weight"]))
    assert got.dtype == np.float64          # the reference returns fp64 (SURVEY §7 "fp64 tail")
    np.testing.assert_allclose(got, g["losses"], rtol=2e-5)


def test_generate_view():
    g = load_golden("generate_view")
    tl = load_golden("training_losses")
    s = O.make_schedule(0.5, 1e-4, 0.02, 5)
    p = params_of(tl)
    np.testing.assert_allclose(O.generate_view(s, p, g["x0"], 0), g["view0"], rtol=5e-5, atol=5e-6)
    np.testing.assert_allclose(O.generate_view(s, p, g["x0"], 2, randn=g["randn"]), g["view2"], rtol=5e-5, atol=5e-6)


def test_topk_edges_match_reference_sets():
    g = load_golden("rebuild")
    got = O.topk_edges(g["scores"], g["deg"])
    ref_u, ref_i = g["edge_u"], g["edge_i"]
    assert sum(len(r) for r in got) == len(ref_u)
    for u, idxs in enumerate(got):
        want = np.sort(ref_i[ref_u == u])
        np.testing.assert_array_equal(idxs, want)      # no ties in this fixture -> identical sets
    assert len(got[1]) == 0                              # degree-0 user emits nothing


def test_topk_tie_break_is_value_desc_index_asc():
    s = np.array([[1, 3, 3, 3, 0, 3]], dtype=np.float32)
    np.testing.assert_array_equal(O.topk_edges(s, [2])[0], [1, 2])
    np.testing.assert_array_equal(O.topk_edges(s, [5])[0], [0, 1, 2, 3, 5])
    np.testing.assert_array_equal(O.topk_edges(-np.zeros((1, 4), np.float32), [3])[0], [0, 1, 2])


def _canon_coo(idx, val):
    order = np.lexsort((idx[1], idx[0]))
    return idx[0][order], idx[1][order], val[order]


@pytest.mark.parametrize("which", ["adj", "bi"])
def test_normalized_adj(which):
    g = load_golden("rebuild")
    U, I = 40, 120
    if which == "adj":
        indptr, indices, vals = O.normalized_adj_csr(g["edge_u"], g["edge_i"], U, I)
    else:
        indptr, indices, vals = O.normalized_adj_csr(g["trn_u"], g["trn_i"], U, I)
    r, c, v = _canon_coo(g[f"{which}_idx"], g[f"{which}_val"])
    rows = np.repeat(np.arange(U + I), np.diff(indptr))
    np.testing.assert_array_equal(rows, r)
    np.testing.assert_array_equal(indices, c)
    np.testing.assert_array_equal(vals, v)               # bit-exact fp32 values


@pytest.mark.parametrize("tag,mods", [("3", ["image", "text", "audio"]), ("2", ["image", "text"])])
def test_gcn_mm(tag, mods):
    g = load_golden(f"gcn_mm_{tag}")
    U, I = 40, 120
    adj = O.normalized_adj_csr(g["trn_u"], g["trn_i"], U, I)
    madj = [O.normalized_adj_csr(g[f"adj_u.{m}"], g[f"adj_i.{m}"], U, I) for m in mods]
    out = O.gcn_mm(g["u_embs"], g["i_embs"], [g[f"feat.{m}"] for m in mods], [g[f"lin_w.{m}"] for m in mods],
                   [g[f"lin_b.{m}"] for m in mods], g["modal_weight"], adj, madj,
                   float(g["modal_adj_weight"]), float(g["residual_weight"]))
    np.testing.assert_allclose(out["final"], g["final"], rtol=2e-5, atol=2e-6)
    for k, m in enumerate(mods):
        np.testing.assert_allclose(out["modal"][k], g[f"z.{m}"], rtol=2e-5, atol=2e-6)


def test_losses():
    g = load_golden("losses")
    np.testing.assert_allclose(O.info_nce(g["v1"], g["v2"], g["idx"], float(g["temp"])), g["infonce"], rtol=1e-5)
    np.testing.assert_allclose(O.bpr_loss(g["u"], g["p"], g["n"]), g["bpr"], rtol=1e-5)
    np.testing.assert_allclose(O.l2_reg_loss(float(g["l2_reg"]), [g["v1"], g["v2"]]), g["l2"], rtol=1e-5)
    with pytest.raises(ValueError):
        O.info_nce(g["v1"], g["v2"][:, :32], g["idx"], 0.2)


def test_cl_propagate():
    g = load_golden("cl_propagate")
    adj = O.normalized_adj_csr(g["trn_u"], g["trn_i"], 40, 120)
    mean, l1 = O.cl_propagate(adj, g["u"], g["i"], list(g["rand"]), float(g["noise_degree"]))
    np.testing.assert_allclose(l1, g["layer1"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(mean, g["mean"], rtol=2e-5, atol=2e-6)
