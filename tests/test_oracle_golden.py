"""The oracle restatement against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import mfm_oracle as O
from helpers import load_golden, golden_params, rel_l2, tiny_case, tiny_kl_case, tiny_kl_ef_case

TOL = 2e-5   # fp32 op-order noise between two torch formulations of the same math


@pytest.mark.parametrize("head,od", [("l1", 1), ("ce", 3), ("l1", 4)])
def test_tiny_full_step(head, od):
    g, configs, P, x, y, noise, T, n = tiny_case(head, od)
    # init_params reproduces torch.manual_seed(seed); MFM(*configs) bit-exactly
    P2 = O.init_params(configs, int(g["meta"][0]))
    assert list(P2) == list(P)
    for k in P:
        assert torch.equal(P[k], P2[k]), k
    newP, losses, G, out = O.train_step(P, x, y, configs, noise, {}, head=head)
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(out[k], g["lat/" + k]) < TOL, k
    for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
        assert rel_l2(out[k], g[k]) < TOL, k
    for k, v in losses.items():
        assert abs(v - float(g["loss/" + k])) <= TOL * abs(float(g["loss/" + k])) + 1e-7, k
    n_grads = 0
    for k in P:
        if "g/" + k in g:
            assert rel_l2(G[k], g["g/" + k]) < TOL, k
            n_grads += 1
        else:
            assert G[k] is None and k in O.UNUSED_PARAMS
    assert n_grads == 86
    for k in P:
        assert rel_l2(newP[k], g["p1/" + k]) < TOL, k


def test_mosi_b32_digest():
    """BASELINE configs[0]: MOSI shapes, best_acc hyper-parameters, B=32, T=20."""
    g = load_golden("mosi_b32.npz")
    seed, T, n, data_seed, noise_seed, od = [int(v) for v in g["meta"]]
    configs = O.best_acc_configs(dropout=False)
    P = O.init_params(configs, seed)
    assert sum(v.numel() for v in P.values()) == 717719
    x, y = O.synthetic_batch(configs, T, n, data_seed)
    noise = O.draw_mmd_noise(configs, n, noise_seed)
    newP, losses, G, out = O.train_step(P, x, y, configs, noise, {}, head="l1")
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(out[k], g["lat/" + k]) < TOL, k
    assert rel_l2(out["y_hat"], g["y_hat"]) < TOL
    assert rel_l2(out["x_a_hat"], g["x_a_hat"]) < TOL
    assert rel_l2(out["x_l_hat"][0], g["x_l_hat_t0"]) < TOL
    assert rel_l2(out["x_l_hat"][-1], g["x_l_hat_tlast"]) < TOL
    for k, v in losses.items():
        assert abs(v - float(g["loss/" + k])) <= TOL * abs(float(g["loss/" + k])), k
    names = [str(s) for s in g["grad_names"]]
    assert len(names) == 86
    norms = np.array([float(G[k].double().norm()) for k in names])
    assert np.allclose(norms, g["grad_norms"], rtol=1e-4, atol=1e-9)
    for k in ("last_to_zy_fc1.weight", "encoder_a.lstm.weight_ih", "decoder_v.lstm.weight_hh",
              "mfn_encoder.gamma1_fc1.bias"):
        assert rel_l2(G[k], g["g/" + k]) < 1e-4, k
    dn = np.array([float((newP[k] - P[k]).double().norm()) for k in names])
    assert np.allclose(dn, g["p1_minus_p0_norms"], rtol=1e-3)


def test_tiny_kl_full_step():
    """MFM_KL (mfm_model.py:662-764) restated in oracle.mfm_kl_forward, against the unmodified reference's MFM_KL."""
    g, configs, P, x, y, T, n = tiny_kl_case()
    P2 = O.init_params(configs, int(g["meta"][0]), variant="kl")
    assert list(P2) == list(P) and len(P) == 104
    for k in P:
        assert torch.equal(P[k], P2[k]), k
    newP, losses, G, out = O.train_step(P, x, y, configs, None, {}, head="l1", variant="kl")
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(out[k], g["lat/" + k]) < TOL, k
    for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
        assert rel_l2(out[k], g[k]) < TOL, k
    for k, v in losses.items():
        assert abs(v - float(g["loss/" + k])) <= TOL * abs(float(g["loss/" + k])) + 1e-7, k
    for k in P:
        if "g/" + k in g:
            assert rel_l2(G[k], g["g/" + k]) < TOL, k
        else:
            assert G[k] is None and k in O.UNUSED_PARAMS
    for k in P:
        assert rel_l2(newP[k], g["p1/" + k]) < TOL, k


def test_tiny_kl_ef_full_step():
    """MFM_KL_EF (mfm_model.py:557-660) restated in oracle.mfm_kl_ef_forward, against the unmodified reference's class."""
    g, configs, P, x, y, T, n = tiny_kl_ef_case()
    P2 = O.init_params(configs, int(g["meta"][0]), variant="kl_ef")
    assert list(P2) == list(P) and len(P) == 78
    for k in P:
        assert torch.equal(P[k], P2[k]), k
    newP, losses, G, out = O.train_step(P, x, y, configs, None, {}, head="l1", variant="kl_ef")
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(out[k], g["lat/" + k]) < TOL, k
    for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
        assert rel_l2(out[k], g[k]) < TOL, k
    for k, v in losses.items():
        assert abs(v - float(g["loss/" + k])) <= TOL * abs(float(g["loss/" + k])) + 1e-7, k
    for k in P:
        assert rel_l2(G[k], g["g/" + k]) < TOL, k
        assert rel_l2(newP[k], g["p1/" + k]) < TOL, k


@pytest.mark.parametrize("variant,ntensors", [("m_a", 70), ("m_b", 52), ("m_c", 60), ("m_d", 32)])
def test_tiny_ablation_full_step(variant, ntensors):
    """M_A .. M_D (mfm_model.py:201-467) restated in oracle.ablation_forward, against the unmodified reference's classes:
    same init for the same seed, and one train step (mfm_mosi.py:677-697) -- outputs, losses, gradients, post-Adam."""
    from helpers import tiny_ablation_case
    g, configs, P, x, y, noise, T, n = tiny_ablation_case(variant)
    P2 = O.init_params(configs, int(g["meta"][0]), variant=variant)
    assert list(P2) == list(P) and len(P) == ntensors
    for k in P:
        assert torch.equal(P[k], P2[k]), k
    newP, losses, G, out = O.train_step(P, x, y, configs, noise, {}, head="l1", variant=variant)
    for k in ("zl", "za", "zv", "zy"):
        if "lat/" + k in g:
            assert rel_l2(out[k], g["lat/" + k]) < TOL, k
    for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
        assert rel_l2(out[k], g[k]) < TOL, k
    for k, v in losses.items():
        assert abs(v - float(g["loss/" + k])) <= TOL * abs(float(g["loss/" + k])) + 1e-7, k
    for k in P:
        if "g/" + k in g:
            assert rel_l2(G[k], g["g/" + k]) < TOL, k
        else:
            assert G[k] is None, k
        assert rel_l2(newP[k], g["p1/" + k]) < TOL, k


def test_baselines_of_the_mosi_script():
    """EFLSTM and MFN-with-head (test_mosi.py:130-157, 158-265) restated in oracle.eflstm_forward / mfn_baseline_forward, against
    outputs and gradients of the reference's own classes (oracle/make_golden.py section 1e)."""
    g = load_golden("tiny_baselines.npz")
    configs = O.tiny_configs(output_dim=1)
    x, y = torch.from_numpy(g["x"].copy()), torch.from_numpy(g["y"].copy())
    Fn = torch.nn.functional
    Pg = {"mfn_encoder." + k[len("mfn/p/"):]: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in g.items() if k.startswith("mfn/p/")}
    assert len(Pg) == 32
    out = O.mfn_baseline_forward(x, Pg, configs)
    Fn.l1_loss(out.squeeze(1), y).backward()
    assert rel_l2(out.detach(), g["mfn/out"]) < TOL
    for k, p in Pg.items():
        assert rel_l2(p.grad, g["mfn/g/" + k[len("mfn_encoder."):]]) < TOL, k
    Pe = {k[len("ef/p/"):]: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in g.items() if k.startswith("ef/p/")}
    out = O.eflstm_forward(x, Pe)
    Fn.l1_loss(out.squeeze(1), y).backward()
    assert rel_l2(out.detach(), g["ef/out"]) < TOL
    for k, p in Pe.items():
        assert rel_l2(p.grad, g["ef/g/" + k]) < TOL, k


def test_tiny_missing_full_step():
    """MFM_missing (mfm_model.py:766-885) through train_mfm_missing's step (mfm_mosi.py:957-982), restated in
    oracle.mfm_missing_forward / mfm_missing_losses, against the unmodified reference's class."""
    from helpers import tiny_missing_case
    g, configs, P, x, y, noise, T, n = tiny_missing_case()
    P2 = O.init_params(configs, int(g["meta"][0]), variant="missing")
    assert list(P2) == list(P) and len(P) == 126
    for k in P:
        assert torch.equal(P[k], P2[k]), k
    newP, losses, G, out = O.train_step(P, x, y, configs, noise, {}, variant="missing")
    for s in O.MISSING_PASSES:
        for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
            assert rel_l2(out[k + s], g[k + s]) < TOL, k + s
    for k in ("total", "disc", "gen", "mmd", "missing"):
        assert abs(losses[k] - float(g["loss/" + k])) <= TOL * abs(float(g["loss/" + k])) + 1e-7, k
    for k in P:
        if "g/" + k in g:
            assert rel_l2(G[k], g["g/" + k]) < TOL, k
        else:
            assert G[k] is None, k
        assert rel_l2(newP[k], g["p1/" + k]) < TOL, k


def test_seq2seq_and_basic_missing_restatements():
    """seq2seq / basic_missing (mfm_model.py:887-1017) restated in the oracle, against outputs, loss and gradients of the
    reference's own classes under the losses of train_seq2seq / train_basic_missing (make_golden.py section 1g)."""
    g = load_golden("tiny_toy_missing.npz")
    configs = O.tiny_configs(output_dim=1)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    x, y = torch.from_numpy(g["x"].copy()), torch.from_numpy(g["y"].copy())
    n = x.shape[1]
    Fn = torch.nn.functional
    for tag, fwd, nseed, sizes in (("s2s", O.seq2seq_forward, 61, (c["zv_size"], c["za_size"], c["zl_size"])),
                                   ("bm", O.basic_missing_forward, 62, (c["zy_size"],) * 3)):
        P = {k[len(tag) + 3:]: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in g.items() if k.startswith(tag + "/p/")}
        torch.manual_seed(nseed)
        noise = [torch.randn(n, k) for k in sizes]
        o = fwd(x, P, configs, noise)
        if tag == "s2s":
            loss = c["lda_xl"] * Fn.mse_loss(o["x_l_hat_nol"], x[:, :, :d_l]) + c["lda_xa"] * Fn.mse_loss(o["x_a_hat_noa"], x[:, :, d_l:d_l + d_a]) \
                + c["lda_xv"] * Fn.mse_loss(o["x_v_hat_nov"], x[:, :, d_l + d_a:]) + c["lda_mmd"] * o["mmd"]
        else:
            loss = sum(Fn.l1_loss(o[k].squeeze(1), y) for k in ("y_hat_nol", "y_hat_noa", "y_hat_nov")) + c["lda_mmd"] * o["mmd"]
        loss.backward()
        assert abs(float(loss) - float(g[tag + "/loss"])) < TOL * abs(float(g[tag + "/loss"]))
        for k, p in P.items():
            assert rel_l2(p.grad, g["%s/g/%s" % (tag, k)]) < TOL, (tag, k)


def test_oracle_ops_match_their_published_definitions():
    """The torch ops the oracle is built from against oracle/numpy_restatement.py (documented LSTMCell gate order i,f,g,o; mean
    reductions; Adam with eps outside the bias-corrected root; the reference's own MMD with its [n,m,dim] tensor), in float64."""
    from oracle import numpy_restatement as N
    rs = np.random.RandomState(0)
    T, n, d, h = 4, 5, 7, 3
    P = {"e.lstm.weight_ih": rs.randn(4 * h, d), "e.lstm.weight_hh": rs.randn(4 * h, h), "e.lstm.bias_ih": rs.randn(4 * h),
         "e.lstm.bias_hh": rs.randn(4 * h), "e.fc1.weight": rs.randn(h, h), "e.fc1.bias": rs.randn(h)}
    Pt = {k: torch.from_numpy(v) for k, v in P.items()}
    x = rs.randn(T, n, d)
    hh, cc = rs.randn(n, h), rs.randn(n, h)
    h2, c2 = O.lstm_cell(torch.from_numpy(x[0]), torch.from_numpy(hh), torch.from_numpy(cc), Pt, "e.lstm")
    h2n, c2n = N.lstm_cell(x[0], hh, cc, P["e.lstm.weight_ih"], P["e.lstm.weight_hh"], P["e.lstm.bias_ih"], P["e.lstm.bias_hh"])
    assert np.allclose(h2.numpy(), h2n, atol=1e-12) and np.allclose(c2.numpy(), c2n, atol=1e-12)
    # the same through torch's own nn.LSTMCell (what the reference calls)
    cell = torch.nn.LSTMCell(d, h).double()
    with torch.no_grad():
        for leaf in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            getattr(cell, leaf).copy_(Pt["e.lstm." + leaf])
        h3, c3 = cell(torch.from_numpy(x[0]), (torch.from_numpy(hh), torch.from_numpy(cc)))
    assert np.allclose(h3.numpy(), h2n, atol=1e-12) and np.allclose(c3.numpy(), c2n, atol=1e-12)
    z = O.encoder_lstm(torch.from_numpy(x), Pt, "e")
    zn = N.encoder_lstm(x, *(P["e." + k] for k in ("lstm.weight_ih", "lstm.weight_hh", "lstm.bias_ih", "lstm.bias_hh", "fc1.weight", "fc1.bias")))
    assert np.allclose(z.numpy(), zn, atol=1e-12)
    a, b = rs.randn(6, 4), rs.randn(6, 4)
    Fn = torch.nn.functional
    assert abs(float(torch.nn.MSELoss()(torch.from_numpy(a), torch.from_numpy(b))) - N.mse_loss(a, b)) < 1e-14
    assert abs(float(torch.nn.L1Loss()(torch.from_numpy(a), torch.from_numpy(b))) - N.l1_loss(a, b)) < 1e-14
    lab = rs.randint(0, 4, 6)
    assert abs(float(torch.nn.CrossEntropyLoss()(torch.from_numpy(a), torch.from_numpy(lab))) - N.cross_entropy(a, lab)) < 1e-13
    assert np.allclose(Fn.softmax(torch.from_numpy(a), dim=1).numpy(), N.softmax_rows(a), atol=1e-14)
    zz, gg = rs.randn(9, 5), rs.randn(9, 5)
    assert abs(float(O.loss_mmd(torch.from_numpy(zz), torch.from_numpy(gg))) - N.loss_mmd(zz, gg)) < 1e-13
    assert np.allclose(O.compute_kernel(torch.from_numpy(zz), torch.from_numpy(gg)).numpy(), N.compute_kernel(zz, gg), atol=1e-14)
    # Adam: three steps of torch.optim.Adam, of the oracle's adam_step and of the published update
    p0 = rs.randn(5, 3)
    grads = [rs.randn(5, 3) * (10.0 ** -k) for k in range(3)]
    tp = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([tp])
    Po, st = {"w": torch.from_numpy(p0.copy())}, {}
    pn, mn, vn = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for t, g_ in enumerate(grads, 1):
        tp.grad = torch.from_numpy(g_.copy())
        opt.step()
        Po = O.adam_step(Po, {"w": torch.from_numpy(g_.copy())}, st)
        pn, mn, vn = N.adam_step(pn, g_, mn, vn, t)
        assert np.allclose(tp.detach().numpy(), pn, atol=1e-14) and np.allclose(Po["w"].numpy(), pn, atol=1e-14)
