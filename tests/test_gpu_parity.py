"""Parity of the CUDA path (through the C ABI) against the golden vectors of the unmodified reference and
against the oracle on the same seeded inputs.  Tolerance: 1e-3 relative (BASELINE.json north_star), fp32."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import mfm_oracle as O
from helpers import load_golden, rel_l2, tiny_case, tiny_kl_case, tiny_kl_ef_case

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(params=["simt_fp32", "tcgen05_bf16x3"])
def gemm_path(request):
    """Run a test once per GEMM math path: exact fp32 CUDA cores, and tcgen05 with split-bf16 operands."""
    from factorized_b200.cuda_ops import CudaOps, PATH_SIMT_FP32, PATH_TC_BF16X3
    ops = CudaOps()
    old = ops.get_gemm_path()
    ops.set_gemm_path(PATH_SIMT_FP32 if request.param == "simt_fp32" else PATH_TC_BF16X3)
    yield request.param
    ops.set_gemm_path(old)


def cuda_engine_step(configs, P, x, y, noise, T, n, head):
    from factorized_b200.engine import Engine
    from factorized_b200.cuda_ops import CudaOps
    ops = CudaOps()
    dev = torch.device("cuda")
    Pd = OrderedDict((k, v.to(dev)) for k, v in P.items())
    eng = Engine(configs, T, n, dev, ops, head=head)
    out = eng.forward(Pd, x.to(dev).contiguous(), [t.to(dev) for t in noise], train=False)
    dX, dY = eng.losses(y.to(dev))
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in Pd.items())
    eng.backward(Pd, G, dX, dY, eng.dm.lda_mmd)
    torch.cuda.synchronize()
    return eng, out, G


@pytest.mark.parametrize("head,od", [("l1", 1), ("ce", 3), ("l1", 4)])
def test_tiny_golden(head, od):
    g, configs, P, x, y, noise, T, n = tiny_case(head, od)
    eng, out, G = cuda_engine_step(configs, P, x, y, noise, T, n, head)
    report = {}
    for k in ("zl", "za", "zv", "zy"):
        report[k] = rel_l2(out[k], g["lat/" + k])
    for k in ("x_l_hat", "x_a_hat", "x_v_hat"):
        report[k] = rel_l2(out[k].view(T, n, -1), g[k])
    report["y_hat"] = rel_l2(out["y_hat"], g["y_hat"])
    lb = eng.loss_buf.cpu()
    for i, k in ((0, "disc"), (1, "mse_l"), (2, "mse_a"), (3, "mse_v"), (8, "total")):
        report["loss." + k] = abs(float(lb[i]) - float(g["loss/" + k])) / abs(float(g["loss/" + k]))
    mmd_w = float(lb[4:8].sum()) * configs[0]["lda_mmd"]
    report["loss.mmd"] = abs(mmd_w - float(g["loss/mmd"])) / abs(float(g["loss/mmd"]))
    for k in P:
        if "g/" + k in g:
            report["grad." + k] = rel_l2(G[k], g["g/" + k])
    bad = {k: v for k, v in report.items() if not (v < TOL)}
    assert not bad, bad


def relu_branches(eng):
    """The ReLU branch decisions the CUDA run took (dropout off: output > 0 <=> branch taken), keyed like
    oracle.relu's sites, so the oracle's backward runs through the same branches (see oracle.relu)."""
    ws, dm = eng.ws, eng.dm
    T, n = dm.T, dm.B
    pos = lambda t: (t > 0).float().cpu()
    br = dict(att1=pos(ws["H1"]).view(T, n, -1), att2=pos(ws["H2"]).view(T, n, -1),
              gamma1=pos(ws["U1"]).view(T, n, -1), gamma2=pos(ws["U2"]).view(T, n, -1),
              fy1=pos(ws["F1y"]), fy=pos(ws["FY"]), y1=pos(ws["Y1"]))
    for m, tag in enumerate("lav"):
        br["f%s1" % tag] = pos(ws["F1_%d" % m])
        br["f%s" % tag] = pos(ws["EMB%d" % m][:, dm.fy:])
    return br


def _compare_to_oracle(configs, T, n, head, seed=123, data_seed=1234, noise_seed=999):
    P = O.init_params(configs, seed)
    x, y = O.synthetic_batch(configs, T, n, data_seed, head)
    noise = O.draw_mmd_noise(configs, n, noise_seed)
    eng, out, G = cuda_engine_step(configs, P, x, y, noise, T, n, head)
    del O.RELU_REPLAY_VIOLATIONS[:]
    newP, losses, Go, outo = O.train_step(P, x, y, configs, noise, {}, head=head, branches=relu_branches(eng))
    assert not O.RELU_REPLAY_VIOLATIONS, O.RELU_REPLAY_VIOLATIONS[:5]
    report = {}
    for k in ("zl", "za", "zv", "zy", "y_hat"):
        report[k] = rel_l2(out[k], outo[k])
    for k in ("x_l_hat", "x_a_hat", "x_v_hat"):
        report[k] = rel_l2(out[k].view(T, n, -1), outo[k])
    lb = eng.loss_buf.cpu()
    for i, k in ((0, "disc"), (1, "mse_l"), (2, "mse_a"), (3, "mse_v"), (8, "total")):
        report["loss." + k] = abs(float(lb[i]) - losses[k]) / abs(losses[k])
    report["loss.mmd"] = abs(float(lb[4:8].sum()) * configs[0]["lda_mmd"] - losses["mmd"]) / abs(losses["mmd"])
    for k, go in Go.items():
        if go is not None:
            report["grad." + k] = rel_l2(G[k], go)
    # one fused Adam step on the flat buffers against the oracle's post-step parameters
    from factorized_b200.cuda_ops import CudaOps
    ops = CudaOps()
    names = [k for k in P if Go[k] is not None]
    fp = torch.cat([P[k].reshape(-1) for k in names]).cuda()
    fg = torch.cat([G[k].reshape(-1) for k in names])
    m, v = torch.zeros_like(fp), torch.zeros_like(fp)
    st = torch.tensor([1e-3, 0, 0, 0], dtype=torch.float32).cuda()
    ops.adam(fp, fg, m, v, st)
    ref_delta = torch.cat([(newP[k] - P[k]).reshape(-1) for k in names])
    got_delta = fp.cpu() - torch.cat([P[k].reshape(-1) for k in names])
    report["adam.delta"] = rel_l2(got_delta, ref_delta)
    return report


def test_mosi_b32_golden_digest(gemm_path):
    """BASELINE configs[0] shapes against the digest the unmodified reference produced.  Gradients are compared
    on the exact-fp32 path only: without ReLU branch replay (the digest cannot provide it) a knife-edge branch
    flip under the 1e-5-level tensor-core arithmetic would be a property of the data, not of the kernels."""
    g = load_golden("mosi_b32.npz")
    seed, T, n, data_seed, noise_seed, od = [int(v) for v in g["meta"]]
    configs = O.best_acc_configs(dropout=False)
    P = O.init_params(configs, seed)
    x, y = O.synthetic_batch(configs, T, n, data_seed)
    noise = O.draw_mmd_noise(configs, n, noise_seed)
    eng, out, G = cuda_engine_step(configs, P, x, y, noise, T, n, "l1")
    report = {k: rel_l2(out[k], g["lat/" + k]) for k in ("zl", "za", "zv", "zy")}
    report["y_hat"] = rel_l2(out["y_hat"], g["y_hat"])
    report["x_a_hat"] = rel_l2(out["x_a_hat"].view(T, n, -1), g["x_a_hat"])
    report["x_l_hat_t0"] = rel_l2(out["x_l_hat"].view(T, n, -1)[0], g["x_l_hat_t0"])
    report["x_l_hat_tlast"] = rel_l2(out["x_l_hat"].view(T, n, -1)[-1], g["x_l_hat_tlast"])
    lb = eng.loss_buf.cpu()
    for i, k in ((0, "disc"), (1, "mse_l"), (2, "mse_a"), (3, "mse_v"), (8, "total")):
        report["loss." + k] = abs(float(lb[i]) - float(g["loss/" + k])) / abs(float(g["loss/" + k]))
    if gemm_path == "simt_fp32":
        names = [str(s) for s in g["grad_names"]]
        norms = np.array([float(G[k].double().norm()) for k in names])
        report["grad_norms"] = float(np.max(np.abs(norms - g["grad_norms"]) / (g["grad_norms"] + 1e-12)))
        for k in ("last_to_zy_fc1.weight", "encoder_a.lstm.weight_ih", "decoder_v.lstm.weight_hh", "mfn_encoder.gamma1_fc1.bias"):
            report["grad." + k] = rel_l2(G[k], g["g/" + k])
    bad = {k: v for k, v in report.items() if not (v < TOL)}
    assert not bad, bad


@pytest.mark.parametrize("name,input_dims,T,n,head,od", [
    ("mosi_b256", (300, 5, 20), 20, 256, "l1", 1),          # BASELINE configs[1]
    ("mosei_b64", (300, 74, 35), 50, 64, "l1", 1),          # configs[2] shapes, one rank's shard
    ("mosei_b512", (300, 74, 35), 50, 512, "l1", 1),        # configs[2] at its full batch on one GPU
    ("iemocap_b256", (300, 74, 35), 20, 256, "ce", 4),      # configs[3]
    ("pom_b96", (300, 43, 43), 100, 96, "l1", 16),          # configs[4] shapes (ragged batch)
    ("pom_b1024", (300, 43, 43), 100, 1024, "l1", 16),      # configs[4] at its full per-GPU batch
])
def test_full_step_vs_oracle(gemm_path, name, input_dims, T, n, head, od):
    configs = O.best_acc_configs(input_dims=input_dims, output_dim=od, dropout=False)
    report = _compare_to_oracle(configs, T, n, head)
    bad = {k: v for k, v in report.items() if not (v < TOL)}
    worst = max(report, key=report.get)
    print("%s [%s]: worst %s = %.3g" % (name, gemm_path, worst, report[worst]))
    assert not bad, bad


def test_forward_only_inference_skips_the_mmd():
    """evaluate() / predict() of the reference call forward on the whole set under no_grad and discard the MMD
    (mfm_mosi.py:445-465).  With ``eval_skip_mmd`` the O(n^2) statistic is not launched: same decoded outputs and latents,
    mmd reads 0, fewer kernels; in train mode or with grad enabled the flag changes nothing."""
    import factorized_b200 as F
    from factorized_b200.mfm_model import _ops
    g, configs, P, x, y, noise, T, n = tiny_case("l1", 1)
    torch.manual_seed(int(g["meta"][0]))
    model = F.MFM(*configs).cuda().eval()
    xd = x.cuda()
    with torch.no_grad():
        n0 = _ops().launches
        dec_a, mmd_a, _ = model.forward(xd)
        full = _ops().launches - n0
        lat_a = {k: v.clone() for k, v in model.latents.items()}
        model.eval_skip_mmd = True
        n0 = _ops().launches
        dec_b, mmd_b, _ = model.forward(xd)
        lean = _ops().launches - n0
    assert float(mmd_a) > 0.0 and float(mmd_b) == 0.0
    assert lean < full - 30, (lean, full)
    for a, b in zip(dec_a, dec_b):
        assert torch.equal(a, b)
    for k in lat_a:
        assert torch.equal(lat_a[k], model.latents[k])
    decoded, mmd_c, _ = model.forward(xd)                   # grad enabled: the statistic is part of the loss again
    assert float(mmd_c) > 0.0


def test_dropin_module_autograd_and_trainer():
    """The nn.Module boundary: MFM.forward + torch losses + loss.backward() (the reference's own loop body,
    mfm_mosi.py:430-441) against the oracle; then MFMTrainer's fused step against the same."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    g, configs, P, x, y, noise, T, n = tiny_case("l1", 1)
    torch.manual_seed(int(g["meta"][0]))
    model = F.MFM(*configs).cuda().eval()
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), P[k]), k
    torch.manual_seed(int(g["meta"][4]))                   # same CPU generator state as the reference run
    xd, yd = x.cuda(), y.cuda()
    decoded, mmd, missing = model.forward(xd)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    Fn = torch.nn.functional
    gen = c["lda_xl"] * Fn.mse_loss(decoded[0], xd[:, :, :d_l]) + c["lda_xa"] * Fn.mse_loss(decoded[1], xd[:, :, d_l:d_l + d_a]) \
        + c["lda_xv"] * Fn.mse_loss(decoded[2], xd[:, :, d_l + d_a:])
    loss = Fn.l1_loss(decoded[3].squeeze(1), yd) + gen + c["lda_mmd"] * mmd + missing
    loss.backward()
    assert missing == 0.0
    assert abs(float(loss) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(model.latents[k], g["lat/" + k]) < TOL
    bad = {}
    for k, p in model.named_parameters():
        if "g/" + k in g:
            e = rel_l2(p.grad, g["g/" + k])
            if not e < TOL:
                bad[k] = e
        else:
            assert p.grad is None, k
    assert not bad, bad
    # whole-module pickle round trip (mfm_mosi.py:477,481)
    import io
    b = io.BytesIO()
    torch.save(model, b)
    b.seek(0)
    m2 = torch.load(b, weights_only=False)
    torch.manual_seed(int(g["meta"][4]))
    d2, _, _ = m2.forward(xd)
    assert torch.equal(d2[3], decoded[3])

    # fused trainer step: same math, noise injected, no dropout in this config
    torch.manual_seed(int(g["meta"][0]))
    model2 = F.MFM(*configs).cuda()
    tr = MFMTrainer(model2, T, n, head="l1", use_graph=False)
    tr.x.copy_(xd)
    tr.y.copy_(yd.reshape(-1))
    for k in range(4):
        tr.noise[k].copy_(noise[k])
    tr.ops.randn = lambda *a, **kw: None                  # keep the injected noise for the parity check
    tr.step_device()
    torch.cuda.synchronize()
    assert abs(float(tr.eng.loss_buf[8]) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    sd = model2.state_dict()
    bad = {k: rel_l2(sd[k].cpu() - P[k], g["p1/" + k] - P[k].numpy()) for k in P if "g/" + k in g}
    bad = {k: v for k, v in bad.items() if not v < 5e-3}
    assert not bad, bad


def test_trainer_graph_replay_trains():
    """CUDA-graph replay of the fused step: loss decreases on a fixed batch and matches the eager schedule."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    configs = O.best_acc_configs(dropout=True)
    T, n = 20, 64
    x, y = O.synthetic_batch(configs, T, n, 7)
    res = []
    for use_graph in (False, True):
        torch.manual_seed(5)
        model = F.MFM(*configs).cuda()
        tr = MFMTrainer(model, T, n, use_graph=use_graph, seed=77)
        ls = []
        for i in range(6):
            lb = tr.step(x, y)
            ls.append(float(lb[8]))
        res.append(ls)
        assert tr.launches_per_step > 50
    assert res[0][-1] < res[0][0]
    assert np.allclose(res[0], res[1], rtol=2e-3), res


from oracle.rng_replay import train_masks_and_branches as _train_masks_and_branches   # noqa: E402


@pytest.mark.parametrize("name,input_dims,T,n,head,od,use_graph", [
    ("mosi_b256_eager", (300, 5, 20), 20, 256, "l1", 1, False),
    ("mosi_b256_graph", (300, 5, 20), 20, 256, "l1", 1, True),
    ("mosi_b2048_graph", (300, 5, 20), 20, 2048, "l1", 1, True),       # the configuration bench.py measures
    ("iemocap_b256_graph", (300, 74, 35), 20, 256, "ce", 4, True),
])
def test_train_mode_fused_step_vs_oracle(name, input_dims, T, n, head, od, use_graph):
    """The MEASURED configuration against the oracle: MFMTrainer.step in train mode (all nine dropouts active, device
    RNG for masks and MMD noise, CUDA graph), two consecutive steps.  The step's masks / noise / ReLU branches are
    replayed in oracle.train_step (mfm_mosi.py:427-441); losses, latents, every gradient and the post-Adam parameters
    must agree to 1e-3."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    configs = O.best_acc_configs(input_dims=input_dims, output_dim=od, dropout=True)
    torch.manual_seed(123)
    model = F.MFM(*configs).cuda().train()
    P = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    tr = MFMTrainer(model, T, n, head=head, use_graph=use_graph, seed=4242)
    state, report = {}, {}
    for step in (1, 2):
        x, y = O.synthetic_batch(configs, T, n, 1000 + step, head)
        lb = tr.step(x.cuda(), y.cuda())
        torch.cuda.synchronize()
        rng = tr.rng.cpu()
        assert int(rng[1]) == step
        masks, br = _train_masks_and_branches(tr.eng, rng)
        noise = [t.detach().cpu().clone() for t in tr.noise]
        del O.RELU_REPLAY_VIOLATIONS[:]
        newP, losses, Go, outo = O.train_step(P, x, y, configs, noise, state, head=head, train=True, masks=masks, branches=br)
        assert not O.RELU_REPLAY_VIOLATIONS, O.RELU_REPLAY_VIOLATIONS[:5]
        lbc = lb.cpu()
        tag = "s%d." % step
        for i, k in ((0, "disc"), (1, "mse_l"), (2, "mse_a"), (3, "mse_v"), (8, "total")):
            report[tag + "loss." + k] = abs(float(lbc[i]) - losses[k]) / abs(losses[k])
        report[tag + "loss.mmd"] = abs(float(lbc[4:8].sum()) * configs[0]["lda_mmd"] - losses["mmd"]) / abs(losses["mmd"])
        ws = tr.eng.ws
        for k, b in (("zl", "Z0"), ("za", "Z1"), ("zv", "Z2"), ("zy", "ZY"), ("y_hat", "Yhat")):
            report[tag + k] = rel_l2(ws[b], outo[k])
        for k, go in Go.items():
            if go is not None:
                report[tag + "grad." + k] = rel_l2(tr.G[k], go)
        sd = model.state_dict()
        num = sum(float((sd[k].cpu().double() - newP[k].double()).pow(2).sum()) for k in P if Go[k] is not None)
        den = sum(float((newP[k].double() - P[k].double()).pow(2).sum()) for k in P if Go[k] is not None)
        report[tag + "adam.delta"] = (num / den) ** 0.5
        P = newP
    bad = {k: v for k, v in report.items() if not (v < (5e-3 if k.endswith("adam.delta") else TOL))}
    worst = max((k for k in report if not k.endswith("adam.delta")), key=report.get)
    print("%s: worst %s = %.3g, adam.delta %.3g / %.3g" % (name, worst, report[worst], report["s1.adam.delta"], report["s2.adam.delta"]))
    assert not bad, bad


def test_standalone_modules_vs_oracle():
    import factorized_b200 as F
    configs = O.tiny_configs()
    T, n = 5, 7
    x, _ = O.synthetic_batch(configs, T, n, 3)
    P = O.init_params(configs, 11)
    # encoderLSTM on a strided modality slice, with grad to the input
    enc = F.encoderLSTM(3, 3).cuda()
    enc.load_state_dict({k[len("encoder_a."):]: v for k, v in P.items() if k.startswith("encoder_a.")})
    xa = x[:, :, 7:10].clone().requires_grad_(True)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    z_ref = O.encoder_lstm(xa, Pg, "encoder_a")
    z_ref.square().sum().backward()
    xg = x.cuda()[:, :, 7:10]
    xg.requires_grad_(True)
    z = enc.forward(xg)
    z.square().sum().backward()
    assert rel_l2(z, z_ref) < TOL
    assert rel_l2(enc.lstm.weight_ih.grad, Pg["encoder_a.lstm.weight_ih"].grad) < TOL
    assert rel_l2(enc.lstm.bias_hh.grad, Pg["encoder_a.lstm.bias_hh"].grad) < TOL
    assert rel_l2(xg.grad, xa.grad) < TOL
    # decoderLSTM
    dec = F.decoderLSTM(7, 3).cuda()
    dec.load_state_dict({k[len("decoder_a."):]: v for k, v in P.items() if k.startswith("decoder_a.")})
    emb = torch.randn(n, 7, generator=torch.Generator().manual_seed(1))
    e1 = emb.clone().requires_grad_(True)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    r_ref = O.decoder_lstm(e1, T, Pg, "decoder_a")
    r_ref.square().sum().backward()
    e2 = emb.cuda().requires_grad_(True)
    r = dec.forward(e2, T)
    r.square().sum().backward()
    assert rel_l2(r, r_ref) < TOL and rel_l2(e2.grad, e1.grad) < TOL
    assert rel_l2(dec.lstm.weight_ih.grad, Pg["decoder_a.lstm.weight_ih"].grad) < TOL
    assert rel_l2(dec.lstm.weight_hh.grad, Pg["decoder_a.lstm.weight_hh"].grad) < TOL
    assert rel_l2(dec.fc1.weight.grad, Pg["decoder_a.fc1.weight"].grad) < TOL
    # MFN
    mfn = F.MFN(*configs).cuda().eval()
    mfn.load_state_dict({k[len("mfn_encoder."):]: v for k, v in P.items() if k.startswith("mfn_encoder.")})
    Pg = {k: v.clone().requires_grad_(k not in O.UNUSED_PARAMS) for k, v in P.items()}
    last_ref = O.mfn_encoder(x, Pg, configs)
    last_ref.square().sum().backward()
    last = mfn.forward(x.cuda())
    last.square().sum().backward()
    assert rel_l2(last, last_ref) < TOL
    bad = {}
    for k, p in mfn.named_parameters():
        gr = Pg["mfn_encoder." + k].grad
        if gr is None:
            assert p.grad is None
        elif not rel_l2(p.grad, gr) < TOL:
            bad[k] = rel_l2(p.grad, gr)
    assert not bad, bad


def test_train_mfm_entry_point(tmp_path):
    """The drop-in for the reference's train_mfm (mfm_mosi.py:386-503) against the oracle's restatement of the same loop:
    shuffle once with numpy's global RNG (:387-389), time-major batches, the tail dropped (:423), Adam at the default lr
    (:403), the epoch's mean discriminative loss (:442-443), whole-set validation L1 (:449-455), save-best / reload /
    predict / score.  Dropout is off in this configuration, so the only difference is the MMD noise stream (device RNG
    here, CPU generator there), which perturbs the trajectory far below the tolerance."""
    import factorized_b200 as F
    rs = np.random.RandomState(0)
    configs = O.tiny_configs()
    configs[0].update(batchsize=16, num_epochs=2)
    T, D, bs = 5, sum(configs[0]["input_dims"]), 16

    def make(n):
        X = rs.randn(n, T, D).astype(np.float32)
        return X, (2.0 * X[:, :, 0].mean(1)).astype(np.float32)
    Xtr, ytr = make(bs * 12 + 5)                       # 5 samples fall off the last batch, like the reference
    Xva, yva = make(40)
    Xte, yte = make(48)
    np.random.seed(11)
    torch.manual_seed(123)
    out = F.train_mfm(Xtr, ytr, Xva, yva, Xte, yte, configs, verbose=False, save_dir=str(tmp_path))

    # the oracle's statement of the same two epochs
    np.random.seed(11)
    p = np.random.permutation(Xtr.shape[0])
    Xs, ys = torch.from_numpy(np.ascontiguousarray(np.swapaxes(Xtr[p], 0, 1))), torch.from_numpy(ytr[p])
    Xv = torch.from_numpy(np.ascontiguousarray(np.swapaxes(Xva, 0, 1)))
    P, state, hist = O.init_params(configs, 123), {}, []
    for ep in range(2):
        acc = 0.0
        for b in range(Xs.shape[1] // bs):
            noise = O.draw_mmd_noise(configs, bs, 1000 * ep + b)
            P, losses, _, _ = O.train_step(P, Xs[:, b * bs:(b + 1) * bs].contiguous(), ys[b * bs:(b + 1) * bs], configs, noise, state)
            acc += losses["disc"]
        vo = O.mfm_forward(Xv, P, configs, O.draw_mmd_noise(configs, Xv.shape[1], 5))
        hist.append((acc / (Xs.shape[1] // bs), float(torch.nn.functional.l1_loss(vo["y_hat"].squeeze(1), torch.from_numpy(yva)))))

    assert len(out["history"]) == 2 and os.path.exists(out["checkpoint"])
    for (ep, tl, vl), (otl, ovl) in zip(out["history"], hist):
        assert abs(tl - otl) < 1e-3 * abs(otl), (ep, tl, otl)
        assert abs(vl - ovl) < 1e-3 * abs(ovl), (ep, vl, ovl)
    assert out["predictions"].shape == (48,) and np.isfinite(out["predictions"]).all()
    # the reference's score block (mfm_mosi.py:483-498): mae, corr, mult_acc, weighted F-score, confusion matrix, report, accuracy
    assert set(out["scores"]) == {"mae", "corr", "mult_acc", "binary_acc", "mult_f_score", "confusion_matrix", "classification_report"}
    cm = np.asarray(out["scores"]["confusion_matrix"])
    assert cm.sum() == 48 and abs(np.trace(cm) / 48.0 - out["scores"]["binary_acc"]) < 1e-9
    assert abs(out["best_valid"] - min(h[2] for h in out["history"])) < 1e-7
    reloaded = torch.load(out["checkpoint"], weights_only=False)       # whole-module pickle, as the reference saves it
    assert sorted(reloaded.state_dict()) == sorted(out["model"].state_dict())


def test_mfm_kl_variant_golden_module_and_trainer(gemm_path):
    """MFM_KL (mfm_model.py:662-764; what train_mfm builds for config['type'] == 'kl', mfm_mosi.py:398-399) on the GPU:
    the drop-in module under torch autograd against the golden vectors of the unmodified reference, then the fused
    trainer step against the same."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    g, configs, P, x, y, T, n = tiny_kl_case()
    torch.manual_seed(int(g["meta"][0]))
    model = F.MFM_KL(*configs).cuda().eval()
    sd = model.state_dict()
    assert len(sd) == 104
    for k, v in sd.items():
        assert torch.equal(v.cpu(), P[k]), k                      # same construction order -> same init stream
    xd, yd = x.cuda(), y.cuda()
    decoded, kld, missing = model.forward(xd)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    Fn = torch.nn.functional
    gen = c["lda_xl"] * Fn.mse_loss(decoded[0], xd[:, :, :d_l]) + c["lda_xa"] * Fn.mse_loss(decoded[1], xd[:, :, d_l:d_l + d_a]) \
        + c["lda_xv"] * Fn.mse_loss(decoded[2], xd[:, :, d_l + d_a:])
    loss = Fn.l1_loss(decoded[3].squeeze(1), yd) + gen + c["lda_mmd"] * kld + missing
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    assert abs(float(kld.detach()) * c["lda_mmd"] - float(g["loss/mmd"])) < TOL * abs(float(g["loss/mmd"]))
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(model.latents[k], g["lat/" + k]) < TOL
    bad = {}
    for k, p in model.named_parameters():
        if "g/" + k in g:
            e = rel_l2(p.grad, g["g/" + k])
            if not e < TOL:
                bad[k] = e
        else:
            assert p.grad is None, k
    assert not bad, bad
    # fused trainer
    torch.manual_seed(int(g["meta"][0]))
    model2 = F.MFM_KL(*configs).cuda()
    tr = MFMTrainer(model2, T, n, head="l1", use_graph=False)
    tr.x.copy_(xd)
    tr.y.copy_(yd.reshape(-1))
    tr.step_device()
    torch.cuda.synchronize()
    assert abs(float(tr.eng.loss_buf[8]) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    sd2 = model2.state_dict()
    bad = {k: rel_l2(sd2[k].cpu() - P[k], g["p1/" + k] - P[k].numpy()) for k in P if "g/" + k in g}
    bad = {k: v for k, v in bad.items() if not v < 5e-3}
    assert not bad, bad


def test_mfm_kl_train_mode_step_and_train_mfm_dispatch(tmp_path):
    """MFM_KL at MOSI shapes in train mode (dropout masks replayed) against the oracle, and train_mfm's dispatch on
    config['type'] (mfm_mosi.py:398-399)."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    from oracle.rng_replay import train_masks_and_branches
    configs = O.best_acc_configs(dropout=True)
    configs[0]["type"] = "kl"
    T, n = 20, 128
    torch.manual_seed(123)
    model = F.MFM_KL(*configs).cuda().train()
    P = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    tr = MFMTrainer(model, T, n, head="l1", use_graph=True, seed=99)
    x, y = O.synthetic_batch(configs, T, n, 5)
    lb = tr.step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    masks, br = train_masks_and_branches(tr.eng, tr.rng.cpu())
    del O.RELU_REPLAY_VIOLATIONS[:]
    newP, losses, Go, outo = O.train_step(P, x, y, configs, None, {}, head="l1", train=True, masks=masks, branches=br, variant="kl")
    assert not O.RELU_REPLAY_VIOLATIONS
    lbc = lb.cpu()
    assert abs(float(lbc[8]) - losses["total"]) < TOL * abs(losses["total"])
    assert abs(float(lbc[4:8].sum()) * configs[0]["lda_mmd"] - losses["mmd"]) < TOL * abs(losses["mmd"])
    bad = {k: rel_l2(tr.G[k], go) for k, go in Go.items() if go is not None and not rel_l2(tr.G[k], go) < TOL}
    assert not bad, bad
    # train_mfm builds MFM_KL for type == 'kl'
    rs = np.random.RandomState(0)
    cfg = O.tiny_configs()
    cfg[0].update(batchsize=8, num_epochs=1, type="kl")
    D = sum(cfg[0]["input_dims"])
    Xtr, ytr = rs.randn(40, 5, D).astype(np.float32), rs.randn(40).astype(np.float32)
    out = F.train_mfm(Xtr, ytr, Xtr[:16], ytr[:16], Xtr[:16], ytr[:16], cfg, verbose=False, save_dir=str(tmp_path))
    assert type(out["model"]).__name__ == "MFM_KL" and np.isfinite(out["history"][0][1])


@pytest.mark.gpu
def test_mfm_kl_ef_variant_golden_module_and_trainer():
    """MFM_KL_EF (mfm_model.py:557-660: one early-fusion encoder cell over the whole input instead of the MFN) on the GPU: the
    drop-in module under torch autograd against the golden vectors of the unmodified reference, then the fused trainer."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    g, configs, P, x, y, T, n = tiny_kl_ef_case()
    torch.manual_seed(int(g["meta"][0]))
    model = F.MFM_KL_EF(*configs).cuda().eval()
    sd = model.state_dict()
    assert list(sd) == list(P)
    for k, v in sd.items():
        assert torch.equal(v.cpu(), P[k]), k                      # same construction order -> same init stream
    xd, yd = x.cuda(), y.cuda()
    decoded, kld, missing = model.forward(xd)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    Fn = torch.nn.functional
    gen = c["lda_xl"] * Fn.mse_loss(decoded[0], xd[:, :, :d_l]) + c["lda_xa"] * Fn.mse_loss(decoded[1], xd[:, :, d_l:d_l + d_a]) \
        + c["lda_xv"] * Fn.mse_loss(decoded[2], xd[:, :, d_l + d_a:])
    loss = Fn.l1_loss(decoded[3].squeeze(1), yd) + gen + c["lda_mmd"] * kld + missing
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(model.latents[k], g["lat/" + k]) < TOL
    bad = {k: rel_l2(p.grad, g["g/" + k]) for k, p in model.named_parameters() if not rel_l2(p.grad, g["g/" + k]) < TOL}
    assert not bad, bad
    # fused trainer (dropout is 0 in the tiny configuration: its train-mode step equals the reference's golden step)
    torch.manual_seed(int(g["meta"][0]))
    model2 = F.MFM_KL_EF(*configs).cuda()
    tr = MFMTrainer(model2, T, n, head="l1", use_graph=False)
    tr.x.copy_(xd)
    tr.y.copy_(yd.reshape(-1))
    tr.step_device()
    torch.cuda.synchronize()
    assert abs(float(tr.eng.loss_buf[8]) - float(g["loss/total"])) < TOL * abs(float(g["loss/total"]))
    sd2 = model2.state_dict()
    bad = {k: rel_l2(sd2[k].cpu() - P[k], g["p1/" + k] - P[k].numpy()) for k in P}
    bad = {k: v for k, v in bad.items() if not v < 5e-3}
    assert not bad, bad


def test_mfm_kl_ef_train_mode_step_at_mosi_shapes():
    """MFM_KL_EF at MOSI shapes (early-fusion cell h = 120: the CUDA-core recurrence serves it beside the tensor-core encoder
    cells), train mode, CUDA graph, dropout masks replayed in the oracle."""
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    from oracle.rng_replay import train_masks_and_branches
    configs = O.best_acc_configs(dropout=True)
    T, n = 20, 96
    torch.manual_seed(123)
    model = F.MFM_KL_EF(*configs).cuda().train()
    P = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    tr = MFMTrainer(model, T, n, head="l1", use_graph=True, seed=99)
    x, y = O.synthetic_batch(configs, T, n, 5)
    lb = tr.step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    masks, br = train_masks_and_branches(tr.eng, tr.rng.cpu())
    del O.RELU_REPLAY_VIOLATIONS[:]
    newP, losses, Go, outo = O.train_step(P, x, y, configs, None, {}, head="l1", train=True, masks=masks, branches=br, variant="kl_ef")
    assert not O.RELU_REPLAY_VIOLATIONS
    lbc = lb.cpu()
    assert abs(float(lbc[8]) - losses["total"]) < TOL * abs(losses["total"])
    bad = {k: rel_l2(tr.G[k], go) for k, go in Go.items() if go is not None and not rel_l2(tr.G[k], go) < TOL}
    assert not bad, bad
    # post-Adam parameters: the first Adam step is lr * g / (|g| + eps), i.e. sign-like, so entries with |g| ~ eps (the text
    # cell's input weights have thousands) turn a 1e-7 gradient difference into a visible step difference: looser bound on
    # the step, tight bound on the parameters themselves
    sd = model.state_dict()
    bad = {k: rel_l2(sd[k].cpu() - P[k], newP[k] - P[k]) for k in P if Go[k] is not None}
    bad = {k: v for k, v in bad.items() if not v < 3e-2}
    assert not bad, bad
    bad = {k: rel_l2(sd[k].cpu(), newP[k]) for k in P if not rel_l2(sd[k].cpu(), newP[k]) < 1e-4}
    assert not bad, bad


def test_eflstm_baseline_vs_oracle():
    """The early-fusion LSTM baseline (test_mosi.py:130-157) on the encoder kernels: forward, loss and every gradient against the
    oracle's restatement, eval mode (the dropout of the [N, h] head is torch's own)."""
    import factorized_b200 as F
    torch.manual_seed(5)
    T, n, d, h, od = 9, 37, 25, 48, 1
    model = F.EFLSTM(d, h, od, 0.3).cuda().eval()
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    assert sorted(P) == sorted(["lstm.weight_ih", "lstm.weight_hh", "lstm.bias_ih", "lstm.bias_hh", "fc1.weight", "fc1.bias",
                                "fc2.weight", "fc2.bias"])
    x = torch.randn(T, n, d)
    y = torch.randn(n)
    out_ref = O.eflstm_forward(x, P)
    torch.nn.functional.l1_loss(out_ref.squeeze(1), y).backward()
    out = model.forward(x.cuda())
    torch.nn.functional.l1_loss(out.squeeze(1), y.cuda()).backward()
    assert rel_l2(out, out_ref) < TOL
    bad = {k: rel_l2(p.grad, P[k].grad) for k, p in model.named_parameters() if not rel_l2(p.grad, P[k].grad) < TOL}
    assert not bad, bad
