"""Generate tests/golden/*.npz from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python oracle/make_golden.py``

It imports ``/root/reference/mfm_model.py`` under Python 3 with
``torch.Tensor.cuda`` neutralised (the reference hard-codes ``.cuda()``,
mfm_model.py:29,51-52,76-77,147-153; there is no GPU here), runs the
reference ``MFM`` through the py3 restatement of the 25-line train step
(mfm_mosi.py:427-441; CE head mfm_mosi_acc.py:441-452), and stores inputs,
outputs, losses, gradients and post-Adam parameters.  It also checks the
oracle restatement (oracle/mfm_oracle.py) against the live reference and
refuses to write fixtures if they disagree.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import mfm_oracle as O

REF = "/root/reference"


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; golden vectors can only be made in the build container")
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import mfm_model as ref  # noqa
    return ref


def run_reference(ref, configs, seed, T, n, head, data_seed, noise_seed, variant="mfm"):
    torch.manual_seed(seed)
    cls = dict(mfm=ref.MFM, kl=ref.MFM_KL, kl_ef=ref.MFM_KL_EF, m_a=ref.M_A, m_b=ref.M_B, m_c=ref.M_C, m_d=ref.M_D)[variant]
    model = cls(*configs).eval()                                                       # eval(): the 9 dropouts become identity
    params0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, y = O.synthetic_batch(configs, T, n, data_seed, head)
    lat = {}
    hooks = {}
    for k, owner, attr in (("zl", "encoder_l", "fc1"), ("za", "encoder_a", "fc1"), ("zv", "encoder_v", "fc1"),
                           ("zy", "last_to_zy_fc1", None)):
        if hasattr(model, owner):                    # the ablation models lack some of the four latents
            hooks[k] = getattr(model, owner) if attr is None else getattr(getattr(model, owner), attr)
    if variant in ("kl", "kl_ef"):                   # the latents are the means: one more Linear after the encoders
        hooks = dict(zl=model.last_to_zl_fc1, za=model.last_to_za_fc1, zv=model.last_to_zv_fc1, zy=model.last_to_zy_fc1)
    for k, m in hooks.items():
        m.register_forward_hook(lambda mod, i, o, k=k: lat.__setitem__(k, o.detach().clone()))
    opt = torch.optim.Adam(model.parameters())     # mfm_mosi.py:403
    opt.zero_grad()
    torch.manual_seed(noise_seed)                   # fixes loss_MMD's four randn draws
    decoded, mmd, missing = model.forward(x)
    x_l_hat, x_a_hat, x_v_hat, y_hat = decoded
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    F = torch.nn.functional
    yh = y_hat.squeeze(1) if y_hat.shape[1] == 1 else y_hat
    mse = [F.mse_loss(x_l_hat, x[:, :, :d_l]), F.mse_loss(x_a_hat, x[:, :, d_l:d_l + d_a]),
           F.mse_loss(x_v_hat, x[:, :, d_l + d_a:])]
    gen = c["lda_xl"] * mse[0] + c["lda_xa"] * mse[1] + c["lda_xv"] * mse[2]
    disc = F.l1_loss(yh, y) if head == "l1" else F.cross_entropy(yh, y.long())
    mmd_w = c["lda_mmd"] * mmd
    loss = disc + gen + mmd_w + missing
    loss.backward()
    grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in model.named_parameters()}
    opt.step()
    params1 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return dict(params0=params0, params1=params1, grads=grads, x=x, y=y, lat=lat,
                x_l_hat=x_l_hat.detach(), x_a_hat=x_a_hat.detach(), x_v_hat=x_v_hat.detach(), y_hat=y_hat.detach(),
                losses=dict(total=float(loss), disc=float(disc), gen=float(gen), mmd=float(mmd_w),
                            mse_l=float(mse[0]), mse_a=float(mse[1]), mse_v=float(mse[2])))


def check_oracle(r, configs, seed, n, head, noise_seed, tag, variant="mfm"):
    """oracle restatement vs live reference; returns the max abs diff seen."""
    P = O.init_params(configs, seed, variant=variant)
    worst = 0.0
    for k, v in r["params0"].items():
        dd = float((P[k] - v).abs().max())
        worst = max(worst, dd)
    assert worst == 0.0, "init_params does not reproduce the reference's init (max diff %g)" % worst
    noise = O.draw_mmd_noise(configs, n, noise_seed, variant=variant)
    newP, losses, G, out = O.train_step(P, r["x"], r["y"], configs, noise, {}, head=head, variant=variant)
    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-30))
    errs = {}
    for k in r["lat"]:
        errs[k] = rel(out[k], r["lat"][k])
    for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
        errs[k] = rel(out[k], r[k])
    for k, v in r["losses"].items():
        errs["loss." + k] = abs(losses[k] - v) / (abs(v) + 1e-30)
    gmax = 0.0
    for k, g in r["grads"].items():
        if g is None:
            assert G[k] is None, k
            continue
        gmax = max(gmax, rel(G[k], g))
    errs["grads(max rel-L2)"] = gmax
    pmax = max(rel(newP[k], v) for k, v in r["params1"].items())
    errs["params1(max rel-L2)"] = pmax
    w = max(errs.values())
    print("[%s] oracle vs live reference: worst rel err %.3g" % (tag, w))
    for k, v in errs.items():
        if v > 1e-6:
            print("    %-24s %.3g" % (k, v))
    assert w < 2e-5, "oracle restatement disagrees with the reference"
    return w


def write_fixture(path, r, meta):
    blob = dict(meta=np.array(meta), x=r["x"].numpy(), y=r["y"].numpy())
    for k, v in r["params0"].items():
        blob["p0/" + k] = v.numpy()
    for k, v in r["params1"].items():
        blob["p1/" + k] = v.numpy()
    for k, v in r["grads"].items():
        if v is not None:
            blob["g/" + k] = v.numpy()
    for k, v in r["lat"].items():
        blob["lat/" + k] = v.numpy()
    for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
        blob[k] = r[k].numpy()
    for k, v in r["losses"].items():
        blob["loss/" + k] = np.float64(v)
    np.savez_compressed(path, **blob)


def ablations(ref, outdir):
    """(1d) the ablation models of train_mfm_ablation (mfm_mosi.py:651-658): M_A, M_B, M_C, M_D (mfm_model.py:201-467)."""
    for i, variant in enumerate(O.ABLATIONS):
        configs = O.tiny_configs(output_dim=1)
        configs[0]["type"] = variant
        seed, T, n, data_seed, noise_seed = 700 + i, 4 + (i % 2), 6 + i, 21 + i, 90 + i
        r = run_reference(ref, configs, seed, T, n, "l1", data_seed, noise_seed, variant=variant)
        check_oracle(r, configs, seed, n, "l1", noise_seed, "tiny_%s/l1/out1" % variant, variant=variant)
        write_fixture(os.path.join(outdir, "tiny_%s_l1_out1.npz" % variant), r, [seed, T, n, data_seed, noise_seed, 1])


def missing(ref, outdir):
    """(1f) MFM_missing (mfm_model.py:766-885) through the step of train_mfm_missing (mfm_mosi.py:957-982)."""
    configs = O.tiny_configs(output_dim=1)
    seed, T, n, data_seed, noise_seed = 808, 4, 7, 31, 95
    torch.manual_seed(seed)
    model = ref.MFM_missing(*configs).eval()
    params0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, y = O.synthetic_batch(configs, T, n, data_seed, "l1")
    opt = torch.optim.Adam(model.parameters())                                         # mfm_mosi.py:931
    opt.zero_grad()
    torch.manual_seed(noise_seed)
    decoded, decoded_nol, decoded_noa, decoded_nov, mmd, missing_loss = model.forward(x)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    l1, l2 = torch.nn.L1Loss(), torch.nn.MSELoss()
    gen = c["lda_xl"] * l2(decoded[0], x_l) + c["lda_xa"] * l2(decoded[1], x_a) + c["lda_xv"] * l2(decoded[2], x_v) \
        + c["lda_xl"] * l2(decoded_nol[0], x_l) + c["lda_xa"] * l2(decoded_noa[1], x_a) + c["lda_xv"] * l2(decoded_noa[2], x_v)
    disc = l1(decoded[3].squeeze(1), y) + l1(decoded_nol[3].squeeze(1), y) + l1(decoded_noa[3].squeeze(1), y) \
        + l1(decoded_nov[3].squeeze(1), y)
    loss = disc + gen + c["lda_mmd"] * mmd + missing_loss                                # :977-982
    loss.backward()
    grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in model.named_parameters()}
    opt.step()
    params1 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # oracle restatement vs the live class
    P = O.init_params(configs, seed, variant="missing")
    assert list(P) == list(params0) and all(torch.equal(P[k], params0[k]) for k in P), "init differs"
    noise = O.draw_mmd_noise(configs, n, noise_seed)
    newP, losses, G, out = O.train_step(P, x, y, configs, noise, {}, variant="missing")
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    w = abs(losses["total"] - float(loss)) / abs(float(loss))
    w = max(w, abs(losses["missing"] - float(missing_loss)) / abs(float(missing_loss)))
    blob = dict(meta=np.array([seed, T, n, data_seed, noise_seed, 1]), x=x.numpy(), y=y.numpy())
    for sfx, dec in zip(O.MISSING_PASSES, (decoded, decoded_nol, decoded_noa, decoded_nov)):
        for name, t in zip(("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"), dec):
            w = max(w, rel(out[name + sfx], t.detach()))
            blob[name + sfx] = t.detach().numpy()
    for k, g in grads.items():
        if g is None:
            assert G[k] is None, k
        else:
            w = max(w, rel(G[k], g))
            blob["g/" + k] = g.numpy()
    for k in params0:
        w = max(w, rel(newP[k], params1[k]))
        blob["p0/" + k] = params0[k].numpy()
        blob["p1/" + k] = params1[k].numpy()
    print("[tiny_missing] oracle vs live reference: worst rel err %.3g" % w)
    assert w < 2e-5, "oracle restatement disagrees with the reference"
    for k, v in dict(total=loss, disc=disc, gen=gen, mmd=c["lda_mmd"] * mmd, missing=missing_loss).items():
        blob["loss/" + k] = np.float64(float(v))
    np.savez_compressed(os.path.join(outdir, "tiny_missing_l1_out1.npz"), **blob)


def toy_missing_models(ref, outdir):
    """(1g) seq2seq and basic_missing (mfm_model.py:887-1017) with the losses of train_seq2seq / train_basic_missing
    (mfm_mosi.py:819-823, :1153-1157): reconstruction MSEs (resp. three L1 label terms) + lda_mmd * MMD."""
    configs = O.tiny_configs(output_dim=1)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    T, n = 4, 8
    x, y = O.synthetic_batch(configs, T, n, 41, "l1")
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    blob = dict(x=x.numpy(), y=y.numpy(), noise_seed=np.array([61, 62]))
    l1, l2 = torch.nn.L1Loss(), torch.nn.MSELoss()
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    for tag, cls, seed, nseed in (("s2s", ref.seq2seq, 901, 61), ("bm", ref.basic_missing, 902, 62)):
        torch.manual_seed(seed)
        m = cls(*configs).eval()
        torch.manual_seed(nseed)
        res = m.forward(x)
        mmd = res[-1]
        if tag == "s2s":
            outs = dict(x_l_hat_nol=res[0][0], x_a_hat_noa=res[1][0], x_v_hat_nov=res[2][0])
            loss = c["lda_xl"] * l2(res[0][0], x_l) + c["lda_xa"] * l2(res[1][0], x_a) + c["lda_xv"] * l2(res[2][0], x_v) \
                + c["lda_mmd"] * mmd
            sizes = (c["zv_size"], c["za_size"], c["zl_size"])
        else:
            outs = dict(y_hat_nol=res[0], y_hat_noa=res[1], y_hat_nov=res[2])
            loss = l1(res[0].squeeze(1), y) + l1(res[1].squeeze(1), y) + l1(res[2].squeeze(1), y) + c["lda_mmd"] * mmd
            sizes = (c["zy_size"],) * 3
        loss.backward()
        P = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        torch.manual_seed(nseed)
        noise = [torch.randn(n, k) for k in sizes]
        o = (O.seq2seq_forward if tag == "s2s" else O.basic_missing_forward)(x, P, configs, noise)
        if tag == "s2s":
            lo = c["lda_xl"] * l2(o["x_l_hat_nol"], x_l) + c["lda_xa"] * l2(o["x_a_hat_noa"], x_a) \
                + c["lda_xv"] * l2(o["x_v_hat_nov"], x_v) + c["lda_mmd"] * o["mmd"]
        else:
            lo = sum(l1(o[k].squeeze(1), y) for k in ("y_hat_nol", "y_hat_noa", "y_hat_nov")) + c["lda_mmd"] * o["mmd"]
        lo.backward()
        w = abs(float(lo) - float(loss)) / abs(float(loss))
        for k, t in outs.items():
            w = max(w, rel(o[k].detach(), t.detach()))
            blob["%s/%s" % (tag, k)] = t.detach().numpy()
        for k, p in m.named_parameters():
            w = max(w, rel(P[k].grad, p.grad))
            blob["%s/p/%s" % (tag, k)] = p.detach().numpy()
            blob["%s/g/%s" % (tag, k)] = p.grad.numpy()
        blob["%s/loss" % tag] = np.float64(float(loss))
        blob["%s/mmd" % tag] = np.float64(float(mmd))
        blob["%s/seed" % tag] = np.array([seed])
        print("[%s] oracle vs live reference: worst rel err %.3g" % (cls.__name__, w))
        assert w < 2e-5
    np.savez_compressed(os.path.join(outdir, "tiny_toy_missing.npz"), **blob)


def reference_baseline_classes():
    """EFLSTM and MFN of /root/reference/test_mosi.py (:130-157, :158-265).  That script cannot be imported (Python-2 prints,
    argparse and data loading at module level), so only its two class statements are executed -- read from the file at run time,
    unmodified -- in a namespace holding what the script's own imports bind (torch, nn, F)."""
    import ast
    import torch.nn as nn
    import torch.nn.functional as F
    src = open(os.path.join(REF, "test_mosi.py")).read().split("\n")
    starts = [i for i, ln in enumerate(src) if ln.startswith("class EFLSTM") or ln.startswith("class MFN")]
    assert len(starts) == 2
    ns = dict(torch=torch, nn=nn, F=F)
    for a in starts:
        b = a + 1
        while b < len(src) and (src[b].strip() == "" or src[b][0] in "\t "):
            b += 1
        code = "\n".join(src[a:b])
        ast.parse(code)
        exec(compile(code, "test_mosi.py:%d" % (a + 1), "exec"), ns)
    return ns["EFLSTM"], ns["MFN"]


def baselines(ref, outdir):
    """(1e) the MOSI script's baselines: the oracle restatements (eflstm_forward, mfn_baseline_forward) against the live classes,
    and a fixture with their outputs and gradients."""
    EF, MFNB = reference_baseline_classes()
    configs = O.tiny_configs(output_dim=1)
    T, n = 5, 9
    x, y = O.synthetic_batch(configs, T, n, 4)
    blob = dict(x=x.numpy(), y=y.numpy())
    Fn = torch.nn.functional
    # MFN with its output head
    torch.manual_seed(17)
    m = MFNB(*configs).eval()
    out = m.forward(x)
    Fn.l1_loss(out.squeeze(1), y).backward()
    P = {"mfn_encoder." + k: v.detach().clone() for k, v in m.state_dict().items()}
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    o2 = O.mfn_baseline_forward(x, Pg, configs)
    Fn.l1_loss(o2.squeeze(1), y).backward()
    w = float((o2 - out).abs().max())
    for k, p in m.named_parameters():
        w = max(w, float((Pg["mfn_encoder." + k].grad - p.grad).norm() / (p.grad.norm() + 1e-30)))
        blob["mfn/p/" + k] = p.detach().numpy()
        blob["mfn/g/" + k] = p.grad.numpy()
    blob["mfn/out"] = out.detach().numpy()
    print("[baseline MFN] oracle vs live reference: worst err %.3g" % w)
    assert w < 2e-5
    # early-fusion LSTM
    D, h = sum(configs[0]["input_dims"]), 6
    torch.manual_seed(5)
    e = EF(D, h, 1, 0.3).eval()
    out = e.forward(x)
    Fn.l1_loss(out.squeeze(1), y).backward()
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in e.state_dict().items()}
    o2 = O.eflstm_forward(x, Pg)
    Fn.l1_loss(o2.squeeze(1), y).backward()
    w = float((o2 - out).abs().max())
    for k, p in e.named_parameters():
        w = max(w, float((Pg[k].grad - p.grad).norm() / (p.grad.norm() + 1e-30)))
        blob["ef/p/" + k] = p.detach().numpy()
        blob["ef/g/" + k] = p.grad.numpy()
    blob["ef/out"] = out.detach().numpy()
    print("[baseline EFLSTM] oracle vs live reference: worst err %.3g" % w)
    assert w < 2e-5
    np.savez_compressed(os.path.join(outdir, "tiny_baselines.npz"), **blob)


def main():
    ref = import_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    if "--ablations-only" in sys.argv:               # leaves the other fixtures' files untouched
        ablations(ref, outdir)
        return
    if "--baselines-only" in sys.argv:
        baselines(ref, outdir)
        return
    if "--missing-only" in sys.argv:
        missing(ref, outdir)
        return
    if "--toy-only" in sys.argv:
        toy_missing_models(ref, outdir)
        return
    ablations(ref, outdir)
    baselines(ref, outdir)
    missing(ref, outdir)
    toy_missing_models(ref, outdir)

    # ---- (1) tiny awkward config, everything stored, L1 head and CE head -------------
    for head, od in (("l1", 1), ("ce", 3), ("l1", 4)):
        configs = O.tiny_configs(output_dim=od)
        seed, T, n, data_seed, noise_seed = 321, 4, 6, 11, 77
        r = run_reference(ref, configs, seed, T, n, head, data_seed, noise_seed)
        check_oracle(r, configs, seed, n, head, noise_seed, "tiny/%s/out%d" % (head, od))
        blob = dict(meta=np.array([seed, T, n, data_seed, noise_seed, od]), x=r["x"].numpy(), y=r["y"].numpy())
        for k, v in r["params0"].items():
            blob["p0/" + k] = v.numpy()
        for k, v in r["params1"].items():
            blob["p1/" + k] = v.numpy()
        for k, v in r["grads"].items():
            if v is not None:
                blob["g/" + k] = v.numpy()
        for k, v in r["lat"].items():
            blob["lat/" + k] = v.numpy()
        for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
            blob[k] = r[k].numpy()
        for k, v in r["losses"].items():
            blob["loss/" + k] = np.float64(v)
        np.savez_compressed(os.path.join(outdir, "tiny_%s_out%d.npz" % (head, od)), **blob)

    # ---- (1b) MFM_KL (the variant train_mfm dispatches to for config['type'] == 'kl', mfm_mosi.py:398-399) ----
    configs = O.tiny_configs(output_dim=1)
    configs[0]["type"] = "kl"
    seed, T, n, data_seed, noise_seed = 321, 4, 6, 11, 77
    r = run_reference(ref, configs, seed, T, n, "l1", data_seed, noise_seed, variant="kl")
    check_oracle(r, configs, seed, n, "l1", noise_seed, "tiny_kl/l1/out1", variant="kl")
    blob = dict(meta=np.array([seed, T, n, data_seed, noise_seed, 1]), x=r["x"].numpy(), y=r["y"].numpy())
    for k, v in r["params0"].items():
        blob["p0/" + k] = v.numpy()
    for k, v in r["params1"].items():
        blob["p1/" + k] = v.numpy()
    for k, v in r["grads"].items():
        if v is not None:
            blob["g/" + k] = v.numpy()
    for k, v in r["lat"].items():
        blob["lat/" + k] = v.numpy()
    for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
        blob[k] = r[k].numpy()
    for k, v in r["losses"].items():
        blob["loss/" + k] = np.float64(v)
    np.savez_compressed(os.path.join(outdir, "tiny_kl_l1_out1.npz"), **blob)

    # ---- (1c) MFM_KL_EF (mfm_model.py:557-660): the early-fusion variant of the same family ----
    configs = O.tiny_configs(output_dim=1)
    seed, T, n, data_seed, noise_seed = 654, 5, 7, 13, 78
    r = run_reference(ref, configs, seed, T, n, "l1", data_seed, noise_seed, variant="kl_ef")
    check_oracle(r, configs, seed, n, "l1", noise_seed, "tiny_kl_ef/l1/out1", variant="kl_ef")
    blob = dict(meta=np.array([seed, T, n, data_seed, noise_seed, 1]), x=r["x"].numpy(), y=r["y"].numpy())
    for k, v in r["params0"].items():
        blob["p0/" + k] = v.numpy()
    for k, v in r["params1"].items():
        blob["p1/" + k] = v.numpy()
    for k, v in r["grads"].items():
        if v is not None:
            blob["g/" + k] = v.numpy()
    for k, v in r["lat"].items():
        blob["lat/" + k] = v.numpy()
    for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat"):
        blob[k] = r[k].numpy()
    for k, v in r["losses"].items():
        blob["loss/" + k] = np.float64(v)
    np.savez_compressed(os.path.join(outdir, "tiny_kl_ef_l1_out1.npz"), **blob)

    # ---- (2) BASELINE configs[0]: MOSI shapes, best_acc dims, B=32, T=20 --------------
    # parameters are regenerated from the seed (2.9 MB otherwise); the fixture keeps
    # outputs, latents, losses, per-parameter gradient norms and a few raw slices.
    configs = O.best_acc_configs(dropout=False)
    seed, T, n, data_seed, noise_seed = 123, 20, 32, 1234, 999
    r = run_reference(ref, configs, seed, T, n, "l1", data_seed, noise_seed)
    check_oracle(r, configs, seed, n, "l1", noise_seed, "mosi_b32")
    blob = dict(meta=np.array([seed, T, n, data_seed, noise_seed, 1]))
    for k, v in r["lat"].items():
        blob["lat/" + k] = v.numpy()
    blob["y_hat"] = r["y_hat"].numpy()
    blob["x_a_hat"] = r["x_a_hat"].numpy()
    blob["x_l_hat_t0"] = r["x_l_hat"][0].numpy()
    blob["x_l_hat_tlast"] = r["x_l_hat"][-1].numpy()
    blob["x_v_hat_tlast"] = r["x_v_hat"][-1].numpy()
    for k, v in r["losses"].items():
        blob["loss/" + k] = np.float64(v)
    names = [k for k, v in r["grads"].items() if v is not None]
    blob["grad_names"] = np.array(names)
    blob["grad_norms"] = np.array([float(r["grads"][k].double().norm()) for k in names])
    blob["grad_sums"] = np.array([float(r["grads"][k].double().sum()) for k in names])
    blob["p0_norms"] = np.array([float(r["params0"][k].double().norm()) for k in names])
    blob["p1_minus_p0_norms"] = np.array([float((r["params1"][k] - r["params0"][k]).double().norm()) for k in names])
    blob["g/last_to_zy_fc1.weight"] = r["grads"]["last_to_zy_fc1.weight"].numpy()
    blob["g/encoder_a.lstm.weight_ih"] = r["grads"]["encoder_a.lstm.weight_ih"].numpy()
    blob["g/decoder_v.lstm.weight_hh"] = r["grads"]["decoder_v.lstm.weight_hh"].numpy()
    blob["g/mfn_encoder.gamma1_fc1.bias"] = r["grads"]["mfn_encoder.gamma1_fc1.bias"].numpy()
    np.savez_compressed(os.path.join(outdir, "mosi_b32.npz"), **blob)
    for f in sorted(os.listdir(outdir)):
        print("%-28s %8.1f KB" % (f, os.path.getsize(os.path.join(outdir, f)) / 1024))


if __name__ == "__main__":
    main()
