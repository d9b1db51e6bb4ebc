"""Everything either side of the MFM step, on the GPU through the C ABI (this file sorts last: the hot path's own parity tests
run first).  The ablation models M_A .. M_D (mfm_model.py:201-467; train_mfm_ablation, mfm_mosi.py:640-767), MFM_missing (:766-885;
train_mfm_missing), train_mfm_test_zeros, the MOSI script's baselines (test_mosi.py:130-265), seq2seq / basic_missing and the
module-level functions, the classification script's own MFM (mfm_mosi_acc.py:311-394) and the trainer's input staging: golden
vectors of the unmodified reference classes, train-mode steps at MOSI shapes against the oracle with the masks replayed, and the
entry points.  Tolerance 1e-3 relative, fp32 (BASELINE.json north_star)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import mfm_oracle as O
from helpers import rel_l2, tiny_ablation_case

pytestmark = pytest.mark.gpu

TOL = 1e-3
VARIANTS = ["m_a", "m_b", "m_c", "m_d"]


@pytest.mark.parametrize("variant", VARIANTS)
def test_ablation_golden_module_and_trainer(variant):
    """The drop-in class under torch autograd against the golden vectors (same init for the same seed; outputs, latents, loss,
    every gradient), then one fused trainer step against the reference's post-Adam parameters."""
    from factorized_b200.ablations import ABLATION_MODELS
    from factorized_b200.train import MFMTrainer
    g, configs, P, x, y, noise, T, n = tiny_ablation_case(variant)
    torch.manual_seed(int(g["meta"][0]))
    model = ABLATION_MODELS[variant](*configs).cuda().eval()
    sd = model.state_dict()
    assert list(sd) == list(P)
    for k, v in sd.items():
        assert torch.equal(v.cpu(), P[k]), k
    xd, yd = x.cuda(), y.cuda()
    torch.manual_seed(int(g["meta"][4]))                             # fixes loss_MMD's draws (CPU generator, reference order)
    decoded, mmd, missing = model.forward(xd)
    if variant == "m_d":
        assert mmd == 0.0 and isinstance(mmd, float)
        assert torch.equal(decoded[0], xd[:, :, :configs[0]["input_dims"][0]])
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    Fn = torch.nn.functional
    for i, k in enumerate(("x_l_hat", "x_a_hat", "x_v_hat", "y_hat")):
        assert rel_l2(decoded[i], g[k]) < TOL, k
    gen = c["lda_xl"] * Fn.mse_loss(decoded[0], xd[:, :, :d_l]) + c["lda_xa"] * Fn.mse_loss(decoded[1], xd[:, :, d_l:d_l + d_a]) \
        + c["lda_xv"] * Fn.mse_loss(decoded[2], xd[:, :, d_l + d_a:])
    loss = Fn.l1_loss(decoded[3].squeeze(1), yd) + gen + c["lda_mmd"] * mmd + missing
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    lat = model.latents
    assert sorted(lat) == sorted(k[4:] for k in g if k.startswith("lat/"))
    for k in lat:
        assert rel_l2(lat[k], g["lat/" + k]) < TOL, k
    bad = {}
    for k, p in model.named_parameters():
        if "g/" + k in g:
            e = rel_l2(p.grad, g["g/" + k])
            if not e < TOL:
                bad[k] = e
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    assert not bad, bad
    # fused trainer, CUDA graph (dropout is 0 in the tiny configuration: its train-mode step equals the golden step)
    torch.manual_seed(int(g["meta"][0]))
    model2 = ABLATION_MODELS[variant](*configs).cuda()
    tr = MFMTrainer(model2, T, n, head="l1", use_graph=True)
    real_randn = tr.ops.randn
    try:
        tr.ops.randn = lambda *a, **k: None                           # keep the injected Gaussian samples
        for k in tr.eng.mmd_slots:
            tr.noise[k].copy_(noise[k])
        tr.x.copy_(xd)
        tr.y.copy_(yd.reshape(-1))
        tr.step_device()
        torch.cuda.synchronize()
    finally:
        tr.ops.randn = real_randn
    assert abs(float(tr.eng.loss_buf[8]) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    sd2 = model2.state_dict()
    bad = {}
    for k in P:
        d_ref = g["p1/" + k] - P[k].numpy()
        if float(np.abs(d_ref).max()) == 0.0:
            assert torch.equal(sd2[k].cpu(), P[k]), k
            continue
        e = rel_l2(sd2[k].cpu() - P[k], d_ref)
        if not e < 5e-3:
            bad[k] = e
    assert not bad, bad


@pytest.mark.parametrize("variant", VARIANTS)
def test_ablation_train_mode_step_at_mosi_shapes(variant):
    """MOSI shapes (best_acc dims, T=20), train mode with every dropout on, CUDA graph; the masks the step drew are replayed
    in the oracle.  Batch 192: the recurrences take their tensor-core path, the projections the tcgen05 GEMM."""
    from factorized_b200.ablations import ABLATION_MODELS
    from factorized_b200.train import MFMTrainer
    from oracle.rng_replay import train_masks_and_branches
    configs = O.best_acc_configs(dropout=True)
    configs[0]["type"] = variant
    T, n = 20, 192
    torch.manual_seed(123)
    model = ABLATION_MODELS[variant](*configs).cuda().train()
    P = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    tr = MFMTrainer(model, T, n, head="l1", use_graph=True, seed=77)
    x, y = O.synthetic_batch(configs, T, n, 5)
    lb = tr.step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    masks, br = train_masks_and_branches(tr.eng, tr.rng.cpu())
    noise = [tr.noise[k].cpu() if k in tr.eng.mmd_slots else None for k in range(4)]
    del O.RELU_REPLAY_VIOLATIONS[:]
    newP, losses, Go, outo = O.train_step(P, x, y, configs, noise, {}, head="l1", train=True, masks=masks, branches=br,
                                          variant=variant)
    assert not O.RELU_REPLAY_VIOLATIONS, O.RELU_REPLAY_VIOLATIONS[:5]
    lbc = lb.cpu()
    assert abs(float(lbc[0]) - losses["disc"]) < TOL * abs(losses["disc"])
    assert abs(float(lbc[8]) - losses["total"]) < TOL * abs(losses["total"])
    for i, k in enumerate(("mse_l", "mse_a", "mse_v")):
        assert abs(float(lbc[1 + i]) - losses[k]) <= TOL * abs(losses[k]), k
    bad = {k: rel_l2(tr.G[k], go) for k, go in Go.items() if go is not None and not rel_l2(tr.G[k], go) < TOL}
    assert not bad, bad
    sd = model.state_dict()
    bad = {k: rel_l2(sd[k].cpu(), newP[k]) for k in P if not rel_l2(sd[k].cpu(), newP[k]) < 1e-4}
    assert not bad, bad
    # a second step replays the captured graph
    loss1 = float(lbc[8])
    lb = tr.step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    assert np.isfinite(float(lb[8])) and float(lb[8]) != loss1


def test_train_mfm_ablation_entry_point(tmp_path):
    """train_mfm_ablation (mfm_mosi.py:640-767): model selected by config['type'], the epoch loop of train_mfm."""
    import factorized_b200 as F
    configs = O.tiny_configs()
    configs[0].update(type="m_b", batchsize=8, num_epochs=3)
    rs = np.random.RandomState(3)
    D, T = sum(configs[0]["input_dims"]), 5
    Xtr, ytr = rs.randn(40, T, D).astype(np.float32), rs.randn(40).astype(np.float32)
    Xva, yva = rs.randn(16, T, D).astype(np.float32), rs.randn(16).astype(np.float32)
    Xte, yte = rs.randn(24, T, D).astype(np.float32), rs.randn(24).astype(np.float32)
    np.random.seed(1)
    torch.manual_seed(2)
    out = F.train_mfm_ablation(Xtr, ytr, Xva, yva, Xte, yte, configs, verbose=False, save_dir=str(tmp_path))
    assert type(out["model"]).__name__ == "M_B"
    assert len(out["history"]) == 3 and all(np.isfinite(v) for h in out["history"] for v in h[1:])
    assert out["predictions"].shape == (24,) and np.isfinite(out["scores"]["mae"])
    configs[0]["type"] = "mfm"
    with pytest.raises(ValueError):
        F.train_mfm_ablation(Xtr, ytr, Xva, yva, Xte, yte, configs, verbose=False, save_dir=str(tmp_path))


def test_baselines_of_the_mosi_script_vs_reference_golden():
    """The baselines of the reference's MOSI script (test_mosi.py): MFN with its out_fc1/out_fc2 head (:158-265) and the
    early-fusion LSTM (:130-157) -- output and every gradient against the fixture made from the reference's own classes, eval
    mode; same initial weights for the same seed."""
    from factorized_b200 import baselines
    from helpers import load_golden
    g = load_golden("tiny_baselines.npz")
    configs = O.tiny_configs()
    x, y = torch.from_numpy(g["x"].copy()), torch.from_numpy(g["y"].copy())
    n = x.shape[1]
    Fn = torch.nn.functional
    torch.manual_seed(17)
    mfn = baselines.MFN(*configs).cuda().eval()
    for k, v in mfn.state_dict().items():
        assert torch.equal(v.cpu(), torch.from_numpy(g["mfn/p/" + k])), k
    out = mfn.forward(x.cuda())
    assert out.shape == (n, 1)
    Fn.l1_loss(out.squeeze(1), y.cuda()).backward()
    assert rel_l2(out.detach(), g["mfn/out"]) < TOL
    bad = {k: rel_l2(p.grad, g["mfn/g/" + k]) for k, p in mfn.named_parameters() if not rel_l2(p.grad, g["mfn/g/" + k]) < TOL}
    assert not bad, bad
    torch.manual_seed(5)
    ef = baselines.EFLSTM(x.shape[2], 6, 1, 0.3).cuda().eval()
    for k, v in ef.state_dict().items():
        assert torch.equal(v.cpu(), torch.from_numpy(g["ef/p/" + k])), k
    out = ef.forward(x.cuda())
    Fn.l1_loss(out.squeeze(1), y.cuda()).backward()
    assert rel_l2(out.detach(), g["ef/out"]) < TOL
    bad = {k: rel_l2(p.grad, g["ef/g/" + k]) for k, p in ef.named_parameters() if not rel_l2(p.grad, g["ef/g/" + k]) < TOL}
    assert not bad, bad


def test_train_mfm_test_zeros_entry_point(tmp_path):
    """train_mfm_test_zeros (mfm_mosi.py:505-638): MFM trained as train_mfm does, then three test-set predictions with one
    modality zeroed.  The predictions must be the oracle's forward of the TRAINED weights on the same zeroed inputs."""
    import factorized_b200 as F
    rs = np.random.RandomState(7)
    configs = O.tiny_configs()
    configs[0].update(batchsize=8, num_epochs=2, type="kl")             # the type is ignored: the reference builds MFM (:516)
    T, D = 4, sum(configs[0]["input_dims"])
    d_l, d_a, d_v = configs[0]["input_dims"]
    Xtr, ytr = rs.randn(32, T, D).astype(np.float32), rs.randn(32).astype(np.float32)
    Xva, yva = rs.randn(16, T, D).astype(np.float32), rs.randn(16).astype(np.float32)
    Xte, yte = rs.randn(20, T, D).astype(np.float32), rs.randn(20).astype(np.float32)
    np.random.seed(3)
    torch.manual_seed(4)
    out = F.train_mfm_test_zeros(Xtr, ytr, Xva, yva, Xte, yte, configs, verbose=False, save_dir=str(tmp_path))
    assert type(out["model"]).__name__ == "MFM"
    P = OrderedDict((k, v.detach().cpu()) for k, v in out["model"].state_dict().items())
    Xt = torch.from_numpy(np.ascontiguousarray(np.swapaxes(Xte, 0, 1)))
    noise = O.draw_mmd_noise(configs, 20, 1)
    for i, (tag, lo, hi) in enumerate((("nol", 0, d_l), ("noa", d_l, d_l + d_a), ("nov", d_l + d_a, D))):
        Xz = Xt.clone()
        Xz[:, :, lo:hi] = 0.0
        ref = O.mfm_forward(Xz, P, configs, noise)
        assert rel_l2(out["predictions_" + tag], ref["y_hat"].squeeze(1)) < TOL, tag
        rec = float(torch.nn.functional.mse_loss(ref[("x_l_hat", "x_a_hat", "x_v_hat")[i]], Xt[:, :, lo:hi]))
        assert abs(out["recon_" + tag] - rec) < TOL * rec, tag
        assert abs(out["scores_" + tag]["mae"] - float(np.mean(np.abs(out["predictions_" + tag] - yte)))) < 1e-6
    assert not np.allclose(out["predictions_nol"], out["predictions"])


MISSING_OUT = ["%s%s" % (k, s) for s in ("", "_nol", "_noa", "_nov") for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat")]


def test_mfm_missing_golden_module_and_trainer():
    """MFM_missing (mfm_model.py:766-885) on the GPU: the drop-in module under torch autograd through train_mfm_missing's loss
    (mfm_mosi.py:962-982) against the golden vectors of the unmodified reference -- sixteen decoded tensors, loss, every
    gradient -- then the fused trainer step against the reference's post-Adam parameters."""
    from helpers import tiny_missing_case
    from factorized_b200.missing import MFM_missing
    from factorized_b200.train import MFMTrainer
    g, configs, P, x, y, noise, T, n = tiny_missing_case()
    torch.manual_seed(int(g["meta"][0]))
    model = MFM_missing(*configs).cuda().eval()
    sd = model.state_dict()
    assert list(sd) == list(P)
    for k, v in sd.items():
        assert torch.equal(v.cpu(), P[k]), k
    xd, yd = x.cuda(), y.cuda()
    torch.manual_seed(int(g["meta"][4]))
    dec, nol, noa, nov, mmd, missing = model.forward(xd)
    flat = list(dec) + list(nol) + list(noa) + list(nov)
    for k, t in zip(MISSING_OUT, flat):
        assert rel_l2(t, g[k]) < TOL, k
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    Fn = torch.nn.functional
    x_l, x_a, x_v = xd[:, :, :d_l], xd[:, :, d_l:d_l + d_a], xd[:, :, d_l + d_a:]
    gen = c["lda_xl"] * Fn.mse_loss(dec[0], x_l) + c["lda_xa"] * Fn.mse_loss(dec[1], x_a) + c["lda_xv"] * Fn.mse_loss(dec[2], x_v) \
        + c["lda_xl"] * Fn.mse_loss(nol[0], x_l) + c["lda_xa"] * Fn.mse_loss(noa[1], x_a) + c["lda_xv"] * Fn.mse_loss(noa[2], x_v)
    disc = sum(Fn.l1_loss(d[3].squeeze(1), yd) for d in (dec, nol, noa, nov))
    loss = disc + gen + c["lda_mmd"] * mmd + missing
    loss.backward()
    assert abs(float(missing.detach()) - float(g["loss/missing"])) < TOL * float(g["loss/missing"])
    assert abs(float(loss.detach()) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    bad = {}
    for k, p in model.named_parameters():
        if "g/" + k in g:
            e = rel_l2(p.grad, g["g/" + k])
            if not e < TOL:
                bad[k] = e
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    assert not bad, bad
    torch.manual_seed(int(g["meta"][0]))
    model2 = MFM_missing(*configs).cuda()
    tr = MFMTrainer(model2, T, n, head="l1", use_graph=True)
    real_randn = tr.ops.randn
    try:
        tr.ops.randn = lambda *a, **k: None
        for k in range(4):
            tr.noise[k].copy_(noise[k])
        tr.x.copy_(xd)
        tr.y.copy_(yd.reshape(-1))
        tr.step_device()
        torch.cuda.synchronize()
    finally:
        tr.ops.randn = real_randn
    assert abs(float(tr.eng.loss_buf[8]) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    sd2 = model2.state_dict()
    bad = {}
    for k in P:
        d_ref = g["p1/" + k] - P[k].numpy()
        if float(np.abs(d_ref).max()) == 0.0:
            assert torch.equal(sd2[k].cpu(), P[k]), k
            continue
        e = rel_l2(sd2[k].cpu() - P[k], d_ref)
        if not e < 5e-3:
            bad[k] = e
    assert not bad, bad


def test_mfm_missing_train_mode_step_at_mosi_shapes_and_entry_point(tmp_path):
    """MOSI shapes, train mode with every dropout on, CUDA graph, each of the four generative passes drawing its own masks; replayed
    in the oracle.  Then train_mfm_missing (mfm_mosi.py:918-1105) end to end on a small problem."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    from oracle.rng_replay import train_masks_and_branches
    configs = O.best_acc_configs(dropout=True)
    T, n = 20, 128
    torch.manual_seed(123)
    model = F.MFM_missing(*configs).cuda().train()
    P = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    tr = MFMTrainer(model, T, n, head="l1", use_graph=True, seed=55)
    x, y = O.synthetic_batch(configs, T, n, 5)
    lb = tr.step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    masks, br = train_masks_and_branches(tr.eng, tr.rng.cpu())
    noise = [t.cpu() for t in tr.noise]
    del O.RELU_REPLAY_VIOLATIONS[:]
    newP, losses, Go, outo = O.train_step(P, x, y, configs, noise, {}, train=True, masks=masks, branches=br, variant="missing")
    assert not O.RELU_REPLAY_VIOLATIONS, O.RELU_REPLAY_VIOLATIONS[:5]
    lbc = lb.cpu()
    assert abs(float(lbc[0]) - losses["disc"]) < TOL * abs(losses["disc"])
    assert abs(float(lbc[9]) - losses["missing"]) < TOL * abs(losses["missing"])
    assert abs(float(lbc[8]) - losses["total"]) < TOL * abs(losses["total"])
    bad = {k: rel_l2(tr.G[k], go) for k, go in Go.items() if go is not None and not rel_l2(tr.G[k], go) < TOL}
    assert not bad, bad
    sd = model.state_dict()
    bad = {k: rel_l2(sd[k].cpu(), newP[k]) for k in P if not rel_l2(sd[k].cpu(), newP[k]) < 1e-4}
    assert not bad, bad
    # the entry point
    cfg = O.tiny_configs()
    cfg[0].update(batchsize=8, num_epochs=2)
    rs = np.random.RandomState(5)
    D, T2 = sum(cfg[0]["input_dims"]), 4
    Xtr, ytr = rs.randn(32, T2, D).astype(np.float32), rs.randn(32).astype(np.float32)
    Xva, yva = rs.randn(16, T2, D).astype(np.float32), rs.randn(16).astype(np.float32)
    Xte, yte = rs.randn(24, T2, D).astype(np.float32), rs.randn(24).astype(np.float32)
    np.random.seed(1)
    torch.manual_seed(2)
    out = F.train_mfm_missing(Xtr, ytr, Xva, yva, Xte, yte, cfg, verbose=False, save_dir=str(tmp_path))
    assert type(out["model"]).__name__ == "MFM_missing" and len(out["history"]) == 2
    assert all(np.isfinite(v) for h in out["history"] for v in h[1:]) and len(out["recon"]) == 12
    Pm = OrderedDict((k, v.detach().cpu()) for k, v in out["model"].state_dict().items())
    Xt = torch.from_numpy(np.ascontiguousarray(np.swapaxes(Xte, 0, 1)))
    ref = O.mfm_missing_forward(Xt, Pm, cfg, O.draw_mmd_noise(cfg, 24, 1))
    for s in ("", "_nol", "_noa", "_nov"):
        assert rel_l2(out["predictions" + s], ref["y_hat" + s].squeeze(1)) < TOL, s
        assert np.isfinite(out["scores" + s]["mae"])
    d_l = cfg[0]["input_dims"][0]
    rec = float(torch.nn.functional.mse_loss(ref["x_l_hat_nol"], Xt[:, :, :d_l]))
    assert abs(out["recon"]["x_l_hat_nol"] - rec) < TOL * rec


def test_seq2seq_basic_missing_and_module_level_functions():
    """The rest of the reference's mfm_model surface (mfm_mosi.py:30 imports it all): seq2seq / basic_missing (mfm_model.py:887-1017)
    against the fixture made from the reference's own classes -- outputs, loss, every gradient -- and compute_kernel / loss_MMD /
    loss_KLD (:14-38) against the oracle."""
    import mfm_model as shim
    from helpers import load_golden
    g = load_golden("tiny_toy_missing.npz")
    configs = O.tiny_configs(output_dim=1)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    x, y = torch.from_numpy(g["x"].copy()).cuda(), torch.from_numpy(g["y"].copy()).cuda()
    Fn = torch.nn.functional
    for tag, cls, nseed in (("s2s", shim.seq2seq, 61), ("bm", shim.basic_missing, 62)):
        torch.manual_seed(int(g[tag + "/seed"][0]))
        m = cls(*configs).cuda().eval()
        for k, v in m.state_dict().items():
            assert torch.equal(v.cpu(), torch.from_numpy(g["%s/p/%s" % (tag, k)])), (tag, k)
        torch.manual_seed(nseed)                                       # loss_MMD draws on the CPU generator, reference order
        res = m.forward(x)
        if tag == "s2s":
            outs = dict(x_l_hat_nol=res[0][0], x_a_hat_noa=res[1][0], x_v_hat_nov=res[2][0])
            loss = c["lda_xl"] * Fn.mse_loss(res[0][0], x[:, :, :d_l]) + c["lda_xa"] * Fn.mse_loss(res[1][0], x[:, :, d_l:d_l + d_a]) \
                + c["lda_xv"] * Fn.mse_loss(res[2][0], x[:, :, d_l + d_a:]) + c["lda_mmd"] * res[3]
        else:
            outs = dict(y_hat_nol=res[0], y_hat_noa=res[1], y_hat_nov=res[2])
            loss = sum(Fn.l1_loss(r.squeeze(1), y) for r in res[:3]) + c["lda_mmd"] * res[3]
        loss.backward()
        for k, t in outs.items():
            assert rel_l2(t.detach(), g["%s/%s" % (tag, k)]) < TOL, (tag, k)
        assert abs(float(res[-1].detach()) - float(g[tag + "/mmd"])) < TOL * abs(float(g[tag + "/mmd"]))
        assert abs(float(loss.detach()) - float(g[tag + "/loss"])) < TOL * abs(float(g[tag + "/loss"]))
        bad = {k: rel_l2(p.grad, g["%s/g/%s" % (tag, k)]) for k, p in m.named_parameters()
               if not rel_l2(p.grad, g["%s/g/%s" % (tag, k)]) < TOL}
        assert not bad, (tag, bad)
    # module-level functions
    gen = torch.Generator().manual_seed(3)
    a, b = torch.randn(37, 11, generator=gen), torch.randn(29, 11, generator=gen)
    assert rel_l2(shim.compute_kernel(a.cuda(), b.cuda()), O.compute_kernel(a, b)) < 1e-5
    mu, lv = torch.randn(19, 7, generator=gen), 0.3 * torch.randn(19, 7, generator=gen)
    mu_r, lv_r = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    (2.5 * O.loss_kld(mu_r, lv_r)).backward()
    mu_g, lv_g = mu.cuda().requires_grad_(True), lv.cuda().requires_grad_(True)
    kl = shim.loss_KLD(mu_g, lv_g)
    (2.5 * kl).backward()
    assert abs(float(kl.detach()) - float(O.loss_kld(mu, lv))) < 1e-5 * abs(float(O.loss_kld(mu, lv)))
    assert rel_l2(mu_g.grad, mu_r.grad) < 1e-5 and rel_l2(lv_g.grad, lv_r.grad) < 1e-5
    z = torch.randn(33, 9, generator=gen)
    torch.manual_seed(8)
    noise = torch.randn(33, 9)
    z_r = z.clone().requires_grad_(True)
    (0.7 * O.loss_mmd(z_r, noise)).backward()
    z_g = z.cuda().requires_grad_(True)
    torch.manual_seed(8)
    mm = shim.loss_MMD(z_g)
    (0.7 * mm).backward()
    assert abs(float(mm.detach()) - float(O.loss_mmd(z, noise))) < 1e-4 * abs(float(O.loss_mmd(z, noise)))
    assert rel_l2(z_g.grad, z_r.grad) < 1e-4


def test_mfm_of_the_classification_script_returns_differentiable_latents():
    """mfm_mosi_acc.py carries its own MFM (:311-394): output_dim 2, forward returns (zl, za, zv, zy, x_l_hat, x_a_hat, x_v_hat,
    y_hat), and ITS loop applies loss_MMD to the latents (:441) before the cross-entropy step (:446-448).  The drop-in class run
    through that loop body against the oracle's CE train step: loss and every gradient."""
    from factorized_b200 import mfm_acc
    from factorized_b200.functional import loss_MMD
    configs = O.tiny_configs(output_dim=2)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    T, n = 5, 9
    x, y = O.synthetic_batch(configs, T, n, 8, "ce")
    cfg_in = [dict(v) for v in configs]
    cfg_in[0]["output_dim"] = 7                                      # ignored: the script hard-codes 2
    torch.manual_seed(77)
    model = mfm_acc.MFM(*cfg_in).cuda().eval()
    P = O.init_params(configs, 77)
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), P[k]), k
    xd, yd = x.cuda(), y.cuda()
    torch.manual_seed(5)
    zl, za, zv, zy, x_l_hat, x_a_hat, x_v_hat, y_hat = model.forward(xd)
    assert y_hat.shape == (n, 2)
    Fn = torch.nn.functional
    mmd = c["lda_mmd"] * (loss_MMD(zl) + loss_MMD(za) + loss_MMD(zv) + loss_MMD(zy))
    gen = c["lda_xl"] * Fn.mse_loss(x_l_hat, xd[:, :, :d_l]) + c["lda_xa"] * Fn.mse_loss(x_a_hat, xd[:, :, d_l:d_l + d_a]) \
        + c["lda_xv"] * Fn.mse_loss(x_v_hat, xd[:, :, d_l + d_a:])
    loss = Fn.cross_entropy(y_hat, yd.long()) + gen + mmd
    loss.backward()
    noise = O.draw_mmd_noise(configs, n, 5)
    _, losses, Go, outo = O.train_step(P, x, y, configs, noise, {}, head="ce")
    assert abs(float(loss.detach()) - losses["total"]) < TOL * abs(losses["total"])
    assert rel_l2(zy.detach(), outo["zy"]) < TOL and rel_l2(y_hat.detach(), outo["y_hat"]) < TOL
    bad = {}
    for k, p in model.named_parameters():
        if Go[k] is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        elif not rel_l2(p.grad, Go[k]) < TOL:
            bad[k] = rel_l2(p.grad, Go[k])
    assert not bad, bad


def test_input_staging_feeds_every_step_its_own_batch():
    """SURVEY section 8 f2, the input staging of MFMTrainer.step: host batches travel on a copy stream into one of two device
    staging buffers while the previous step computes (the reference does a blocking pageable copy per step, mfm_mosi.py:428-429).
    Six distinct batches are queued back to back without a synchronisation -- pinned, then one pageable -- and every step's
    losses must equal those of an identically seeded trainer fed from device tensors."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    configs = O.tiny_configs()
    T, n, steps = 5, 16, 6
    gen = torch.Generator().manual_seed(9)
    D = sum(configs[0]["input_dims"])
    xs = [(1.0 + 0.5 * i) * torch.randn(T, n, D, generator=gen) for i in range(steps)]
    ys = [torch.randn(n, generator=gen) + i for i in range(steps)]
    trainers = []
    for _ in range(2):
        torch.manual_seed(31)
        trainers.append(MFMTrainer(F.MFM(*configs).cuda(), T, n, head="l1", use_graph=True, seed=5))
    host, dev = trainers
    xp = [x.clone().pin_memory() if i < steps - 1 else x.clone() for i, x in enumerate(xs)]     # the last one pageable
    yp = [y.clone().pin_memory() if i < steps - 1 else y.clone() for i, y in enumerate(ys)]
    got = [host.step(xp[i], yp[i]).clone() for i in range(steps)]          # no sync in between: the double buffer is live
    torch.cuda.synchronize()
    want = []
    for i in range(steps):
        want.append(dev.step(xs[i].cuda(), ys[i].cuda()).clone())
        torch.cuda.synchronize()
    for i in range(steps):
        for slot in (0, 1, 2, 3, 8):
            a, b = float(got[i][slot]), float(want[i][slot])
            assert abs(a - b) <= 1e-4 * abs(b), (i, slot, a, b)
    assert len({round(float(w[8]), 3) for w in want}) == steps             # the batches are distinguishable by their loss
    sh, sd = host.model.state_dict(), dev.model.state_dict()
    assert max(rel_l2(sh[k], sd[k]) for k in sh) < 1e-4
