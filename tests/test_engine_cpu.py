"""Host schedule + hand-derived backward (factorized_b200.engine) checked against
the golden vectors / oracle autograd with the torch statement of the primitive
set injected (tests/emu_ops.py).  CPU only: this tests the HOST LOGIC, the CUDA
kernels are tested against the same statements under -m gpu."""
from collections import OrderedDict

import pytest
import torch

from oracle import mfm_oracle as O
from factorized_b200.engine import Engine
from emu_ops import EmuOps
from helpers import rel_l2, tiny_case, tiny_kl_case, tiny_kl_ef_case


def run_engine(configs, P, x, y, noise, T, n, head, dtype=torch.float32, train=False, rng=None):
    P = OrderedDict((k, v.to(dtype)) for k, v in P.items())
    eng = Engine(configs, T, n, "cpu", EmuOps(), head=head)
    if dtype != torch.float32:
        pytest.skip("engine workspace is fp32")
    out = eng.forward(P, x.contiguous(), noise, train=train, rng=rng)
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    return eng, out, G


@pytest.mark.parametrize("head,od", [("l1", 1), ("ce", 3), ("l1", 4)])
def test_engine_matches_reference_golden(head, od):
    g, configs, P, x, y, noise, T, n = tiny_case(head, od)
    eng, out, G = run_engine(configs, P, x, y, noise, T, n, head)
    tol = 1e-4
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(out[k], g["lat/" + k]) < tol, k
    for k, d in (("x_l_hat", 0), ("x_a_hat", 1), ("x_v_hat", 2)):
        assert rel_l2(out[k].view(T, n, -1), g[k]) < tol, k
    assert rel_l2(out["y_hat"], g["y_hat"]) < tol
    lb = eng.loss_buf
    assert abs(float(lb[0]) - float(g["loss/disc"])) < tol * abs(float(g["loss/disc"])) + 1e-7
    assert abs(float(lb[1]) - float(g["loss/mse_l"])) < tol * float(g["loss/mse_l"])
    assert abs(float(lb[8]) - float(g["loss/total"])) < tol * abs(float(g["loss/total"]))
    bad = []
    for k in P:
        if "g/" + k in g:
            e = rel_l2(G[k], g["g/" + k])
            if e > 2e-4:
                bad.append((k, e))
        else:
            assert float(G[k].abs().max()) == 0.0
    assert not bad, bad


def test_engine_dropout_masks_replay():
    """train=True: the counter-based masks the schedule applies are replayed
    through the oracle (explicit keep-masks) and forward+backward must agree."""
    from emu_ops import keep_mask
    from factorized_b200 import engine as E
    g, configs, P, x, y, noise, T, n = tiny_case("l1", 1)
    configs = [dict(c) for c in configs]
    configs[0].update(zy_to_fy_dropout=0.3, zl_to_fl_dropout=0.2, za_to_fa_dropout=0.5, zv_to_fv_dropout=0.4,
                      fy_to_y_dropout=0.25)
    for c, p in zip(configs[1:5], (0.5, 0.3, 0.2, 0.4)):
        c["drop"] = p
    rng = torch.tensor([12345, 7], dtype=torch.int64)
    eng, out, G = run_engine(configs, P, x, y, noise, T, n, "l1", train=True, rng=rng)
    c = configs[0]
    nn1, nn2, g1, g2 = configs[1:5]
    masks = dict(
        att1=keep_mask(rng, E.SITE_ATT1, nn1["drop"], T * n, nn1["shapes"]).view(T, n, -1),
        att2=keep_mask(rng, E.SITE_ATT2, nn2["drop"], T * n, nn2["shapes"]).view(T, n, -1),
        gamma1=keep_mask(rng, E.SITE_G1, g1["drop"], T * n, g1["shapes"]).view(T, n, -1),
        gamma2=keep_mask(rng, E.SITE_G2, g2["drop"], T * n, g2["shapes"]).view(T, n, -1),
        fy=keep_mask(rng, E.SITE_FY, c["zy_to_fy_dropout"], n, c["fy_size"]),
        fl=keep_mask(rng, E.SITE_FL, c["zl_to_fl_dropout"], n, c["fl_size"]),
        fa=keep_mask(rng, E.SITE_FA, c["za_to_fa_dropout"], n, c["fa_size"]),
        fv=keep_mask(rng, E.SITE_FV, c["zv_to_fv_dropout"], n, c["fv_size"]),
        y=keep_mask(rng, E.SITE_Y, c["fy_to_y_dropout"], n, c["fy_size"]),
    )
    newP, losses, Go, outo = O.train_step(P, x, y, configs, noise, {}, head="l1", train=True, masks=masks)
    assert rel_l2(out["zy"], outo["zy"]) < 1e-4
    assert rel_l2(out["y_hat"], outo["y_hat"]) < 1e-4
    assert abs(float(eng.loss_buf[8]) - losses["total"]) < 1e-4 * abs(losses["total"])
    bad = [(k, rel_l2(G[k], Go[k])) for k in P if Go[k] is not None and rel_l2(G[k], Go[k]) > 3e-4]
    assert not bad, bad


def test_engine_schedule_variants_agree():
    """The trainer's variants of the schedule (MMD joined only in backward, total summed there; reconstruction MSE fused
    into the decoder heads' GEMM) and the experiment switches that split the last backward recurrence into two launches (by
    cells, or in time with the carried state handed on) give the same losses and gradients as the default."""
    g, configs, P, x, y, noise, T, n = tiny_case("l1", 1)
    ref_eng, _, Gref = run_engine(configs, P, x, y, noise, T, n, "l1")
    for variant in ("defer_mmd_join", "split_last_recurrence", "fuse_mse", "time_split"):
        eng = Engine(configs, T, n, "cpu", EmuOps(), head="l1")
        setattr(eng, variant, True)
        eng.forward(OrderedDict(P), x.contiguous(), noise)
        dX, dY = eng.losses(y)
        G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
        eng.backward(OrderedDict(P), G, dX, dY, eng.dm.lda_mmd)
        assert abs(float(eng.loss_buf[8]) - float(ref_eng.loss_buf[8])) < 1e-6 * abs(float(ref_eng.loss_buf[8])), variant
        bad = [(k, rel_l2(G[k], Gref[k])) for k in P if float(Gref[k].abs().max()) > 0 and rel_l2(G[k], Gref[k]) > 1e-6]
        assert not bad, (variant, bad)


def test_engine_timeline_marks_are_inert_without_a_buffer():
    g, configs, P, x, y, noise, T, n = tiny_case("l1", 1)
    eng, out, G = run_engine(configs, P, x, y, noise, T, n, "l1")
    assert eng.stamps is None and eng.stamp_names == []


def test_rng_streams_of_consecutive_steps_are_unrelated():
    """with key = seed + step*C the step-(s+1) mask was the step-s mask shifted by one element.  The stream key
    now passes every input through its own mixing round; masks of consecutive steps / sites agree only by chance."""
    from emu_ops import keep_mask
    m = [keep_mask(torch.tensor([123, s]), 3, 0.5, 64, 128) for s in range(4)]
    for s in range(3):
        for shift in (0, 1, 2):
            a, b = m[s].view(-1)[shift:4096 + shift], m[s + 1].view(-1)[:4096]
            assert abs(float((a == b).float().mean()) - 0.5) < 0.05, (s, shift)
    a, b = keep_mask(torch.tensor([123, 5]), 3, 0.5, 64, 128), keep_mask(torch.tensor([123, 5]), 4, 0.5, 64, 128)
    assert abs(float((a == b).float().mean()) - 0.5) < 0.05


def test_engine_kl_variant_matches_reference_golden():
    """MFM_KL (the variant train_mfm dispatches to for config['type'] == 'kl', mfm_mosi.py:398-399): schedule + hand-derived
    backward against the golden vectors of the unmodified reference's MFM_KL."""
    g, configs, P, x, y, T, n = tiny_kl_case()
    P = OrderedDict(P)
    eng = Engine(configs, T, n, "cpu", EmuOps(), head="l1", variant="kl")
    out = eng.forward(P, x.contiguous(), [torch.zeros(1, 1)] * 4)
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    tol = 1e-4
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(out[k], g["lat/" + k]) < tol, k
    assert rel_l2(out["y_hat"], g["y_hat"]) < tol
    lb = eng.loss_buf
    kld_w = float(lb[4:8].sum()) * configs[0]["lda_mmd"]
    assert abs(kld_w - float(g["loss/mmd"])) < tol * abs(float(g["loss/mmd"]))
    assert abs(float(lb[8]) - float(g["loss/total"])) < tol * abs(float(g["loss/total"]))
    bad = []
    for k in P:
        if "g/" + k in g:
            e = rel_l2(G[k], g["g/" + k])
            if e > 2e-4:
                bad.append((k, e))
        else:
            assert float(G[k].abs().max()) == 0.0, k
    assert not bad, bad


def test_engine_kl_ef_variant_matches_reference_golden():
    """MFM_KL_EF (mfm_model.py:557-660: one early-fusion encoder cell instead of the MFN): schedule + hand-derived backward
    against the golden vectors of the unmodified reference class."""
    g, configs, P, x, y, T, n = tiny_kl_ef_case()
    P = OrderedDict(P)
    assert "ef_encoder.lstm.weight_ih" in P and not any(k.startswith("mfn_encoder") for k in P)
    eng = Engine(configs, T, n, "cpu", EmuOps(), head="l1", variant="kl_ef")
    out = eng.forward(P, x.contiguous(), [torch.zeros(1, 1)] * 4)
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    tol = 1e-4
    for k in ("zl", "za", "zv", "zy"):
        assert rel_l2(out[k], g["lat/" + k]) < tol, k
    assert rel_l2(out["y_hat"], g["y_hat"]) < tol
    lb = eng.loss_buf
    assert abs(float(lb[4:8].sum()) * configs[0]["lda_mmd"] - float(g["loss/mmd"])) < tol * abs(float(g["loss/mmd"]))
    assert abs(float(lb[8]) - float(g["loss/total"])) < tol * abs(float(g["loss/total"]))
    bad = []
    for k in P:
        e = rel_l2(G[k], g["g/" + k])
        if e > 2e-4:
            bad.append((k, e))
    assert not bad, bad


def check_ablation_against_golden(g, eng, out, G, P, configs, T, n, tol=1e-4, gtol=2e-4):
    for k in ("zl", "za", "zv", "zy"):
        if "lat/" + k in g:
            assert rel_l2(out[k], g["lat/" + k]) < tol, k
        else:
            assert out[k] is None, k
    for k in ("x_l_hat", "x_a_hat", "x_v_hat"):
        assert rel_l2(out[k].reshape(T, n, -1), g[k]) < tol, k
    assert rel_l2(out["y_hat"], g["y_hat"]) < tol
    lb = eng.loss_buf
    assert abs(float(lb[0]) - float(g["loss/disc"])) < tol * abs(float(g["loss/disc"])) + 1e-7
    for i, k in enumerate(("mse_l", "mse_a", "mse_v")):
        assert abs(float(lb[1 + i]) - float(g["loss/" + k])) <= tol * float(g["loss/" + k]), k
    assert abs(float(lb[4:8].sum()) * configs[0]["lda_mmd"] - float(g["loss/mmd"])) <= tol * abs(float(g["loss/mmd"])) + 1e-9
    assert abs(float(lb[8]) - float(g["loss/total"])) < tol * abs(float(g["loss/total"]))
    bad = []
    for k in P:
        if "g/" + k in g:
            e = rel_l2(G[k], g["g/" + k])
            if e > gtol:
                bad.append((k, e))
        else:
            assert float(G[k].abs().max()) == 0.0, k
    assert not bad, bad


@pytest.mark.parametrize("variant", ["m_a", "m_b", "m_c", "m_d"])
@pytest.mark.parametrize("fused", [False, True])
def test_engine_ablation_variants_match_reference_golden(variant, fused):
    """M_A .. M_D (mfm_model.py:201-467, the models of train_mfm_ablation mfm_mosi.py:651-658): schedule + hand-derived
    backward of factorized_b200.ablations against the golden vectors of the unmodified reference classes, in the module's
    configuration and in the trainer's (MSE fused into the decoder heads, MMD joined in backward)."""
    from helpers import tiny_ablation_case
    from factorized_b200.ablations import AblationEngine
    g, configs, P, x, y, noise, T, n = tiny_ablation_case(variant)
    P = OrderedDict(P)
    eng = AblationEngine(configs, T, n, "cpu", EmuOps(), head="l1", variant=variant)
    eng.fuse_mse = eng.defer_mmd_join = fused
    out = eng.forward(P, x.contiguous(), noise)
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    if fused:
        out = dict(out)
        for k in ("x_l_hat", "x_a_hat", "x_v_hat"):       # the fused heads never write x_hat: compare the losses and gradients
            if out[k] is None:
                out[k] = torch.from_numpy(g[k].copy())
    check_ablation_against_golden(g, eng, out, G, P, configs, T, n)


@pytest.mark.parametrize("variant", ["m_a", "m_b", "m_c", "m_d"])
def test_engine_ablation_dropout_masks_replay(variant):
    """train=True with every dropout on: the counter-based masks the schedule applies, replayed through the oracle."""
    from emu_ops import keep_mask
    from factorized_b200 import engine as E
    from helpers import tiny_ablation_case
    from factorized_b200.ablations import AblationEngine
    g, configs, P, x, y, noise, T, n = tiny_ablation_case(variant)
    configs = [dict(c) for c in configs]
    configs[0].update(zy_to_fy_dropout=0.3, zl_to_fl_dropout=0.2, za_to_fa_dropout=0.5, zv_to_fv_dropout=0.4,
                      fy_to_y_dropout=0.25)
    for c, p in zip(configs[1:5], (0.5, 0.3, 0.2, 0.4)):
        c["drop"] = p
    rng = torch.tensor([4321, 3], dtype=torch.int64)
    P = OrderedDict(P)
    eng = AblationEngine(configs, T, n, "cpu", EmuOps(), head="l1", variant=variant)
    out = eng.forward(P, x.contiguous(), noise, train=True, rng=rng)
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    c = configs[0]
    nn1, nn2, g1, g2 = configs[1:5]
    masks = dict(
        att1=keep_mask(rng, E.SITE_ATT1, nn1["drop"], T * n, nn1["shapes"]).view(T, n, -1),
        att2=keep_mask(rng, E.SITE_ATT2, nn2["drop"], T * n, nn2["shapes"]).view(T, n, -1),
        gamma1=keep_mask(rng, E.SITE_G1, g1["drop"], T * n, g1["shapes"]).view(T, n, -1),
        gamma2=keep_mask(rng, E.SITE_G2, g2["drop"], T * n, g2["shapes"]).view(T, n, -1),
        fy=keep_mask(rng, E.SITE_FY, c["zy_to_fy_dropout"], n, c["fy_size"]),
        fl=keep_mask(rng, E.SITE_FL, c["zl_to_fl_dropout"], n, c["fl_size"]),
        fa=keep_mask(rng, E.SITE_FA, c["za_to_fa_dropout"], n, c["fa_size"]),
        fv=keep_mask(rng, E.SITE_FV, c["zv_to_fv_dropout"], n, c["fv_size"]),
        y=keep_mask(rng, E.SITE_Y, c["fy_to_y_dropout"], n, c["fy_size"]),
    )
    onoise = [None if v.shape[0] != n else v for v in noise]
    newP, losses, Go, outo = O.train_step(P, x, y, configs, onoise, {}, head="l1", train=True, masks=masks, variant=variant)
    assert rel_l2(out["y_hat"], outo["y_hat"]) < 1e-4
    assert abs(float(eng.loss_buf[8]) - losses["total"]) < 1e-4 * abs(losses["total"])
    bad = [(k, rel_l2(G[k], Go[k])) for k in P if Go[k] is not None and rel_l2(G[k], Go[k]) > 3e-4]
    assert not bad, bad
    assert all(float(G[k].abs().max()) == 0.0 for k in P if Go[k] is None)


@pytest.mark.parametrize("variant", ["m_a", "m_b", "m_c", "m_d"])
def test_ablation_modules_init_and_trainer_step_match_reference_golden(variant):
    """The drop-in classes M_A .. M_D draw the reference's initial weights for the same seed (same construction order, same
    state-dict keys), and one fused trainer step on them (host logic, primitives injected) lands on the reference's
    post-Adam parameters."""
    from helpers import tiny_ablation_case
    from factorized_b200.ablations import ABLATION_MODELS, AblationEngine
    from factorized_b200.train import MFMTrainer
    g, configs, P, x, y, noise, T, n = tiny_ablation_case(variant)
    torch.manual_seed(int(g["meta"][0]))
    model = ABLATION_MODELS[variant](*configs).eval()
    sd = model.state_dict()
    assert list(sd) == list(P)
    for k in P:
        assert torch.equal(sd[k], P[k]), k
    tr = MFMTrainer(model, T, n, head="l1", _test_ops=EmuOps())
    assert isinstance(tr.eng, AblationEngine) and tr.eng.abl == variant
    drawn = []
    tr.ops.randn = lambda out, rng, site: drawn.append(site)        # keep the injected noise
    for k in tr.eng.mmd_slots:
        tr.noise[k].copy_(noise[k])
    lb = tr.step(x, y)
    assert len(drawn) == len(tr.eng.mmd_slots)
    assert abs(float(lb[8]) - float(g["loss/total"])) < 1e-4 * abs(float(g["loss/total"]))
    worst = 0.0
    for k, v in model.state_dict().items():
        d_ref, d_got = torch.from_numpy(g["p1/" + k]) - P[k], v - P[k]
        if float(d_ref.norm()) == 0.0:
            assert float(d_got.abs().max()) == 0.0, k
            continue
        worst = max(worst, float((d_got - d_ref).norm() / d_ref.norm()))
    assert worst < 2e-3, worst


MISSING_OUT = ["%s%s" % (k, s) for s in ("", "_nol", "_noa", "_nov") for k in ("x_l_hat", "x_a_hat", "x_v_hat", "y_hat")]


@pytest.mark.parametrize("deferred", [False, True])
def test_engine_missing_variant_matches_reference_golden(deferred):
    """MFM_missing (mfm_model.py:766-885) through train_mfm_missing's step (mfm_mosi.py:957-982): schedule + hand-derived
    backward of factorized_b200.missing against the golden vectors of the unmodified reference class -- all sixteen decoded
    tensors, the losses, all 122 gradients."""
    from helpers import tiny_missing_case
    from factorized_b200.missing import MissingEngine
    g, configs, P, x, y, noise, T, n = tiny_missing_case()
    P = OrderedDict(P)
    eng = MissingEngine(configs, T, n, "cpu", EmuOps(), head="l1")
    eng.defer_mmd_join = deferred
    out = eng.forward(P, x.contiguous(), noise)
    dX, dY = eng.losses(y)
    assert sorted(dX) == [(0, 0), (0, 1), (0, 2), (1, 0), (2, 1), (2, 2)] and len(dY) == 4
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    tol = 1e-4
    for k in MISSING_OUT:
        got = out[k] if k.startswith("y_hat") else out[k].reshape(T, n, -1)
        assert rel_l2(got, g[k]) < tol, k
    lb = eng.loss_buf
    for slot, k, w in ((0, "disc", 1.0), (9, "missing", 1.0), (8, "total", 1.0)):
        assert abs(float(lb[slot]) - float(g["loss/" + k])) < tol * abs(float(g["loss/" + k])), k
    assert abs(float(lb[4:8].sum()) * configs[0]["lda_mmd"] - float(g["loss/mmd"])) < tol * abs(float(g["loss/mmd"]))
    c = configs[0]
    gen = c["lda_xl"] * float(lb[1]) + c["lda_xa"] * float(lb[2]) + c["lda_xv"] * float(lb[3])
    assert abs(gen - float(g["loss/gen"])) < tol * float(g["loss/gen"])
    d_l = c["input_dims"][0]
    assert abs(float(lb[10]) - float(((torch.from_numpy(g["x_l_hat"]) - x[:, :, :d_l]) ** 2).mean())) < tol * float(lb[10])
    bad = []
    for k in P:
        if "g/" + k in g:
            e = rel_l2(G[k], g["g/" + k])
            if e > 2e-4:
                bad.append((k, e))
        else:
            assert float(G[k].abs().max()) == 0.0, k
    assert not bad, bad


def missing_masks(rng, configs, n):
    """The keep-masks of the MFM_missing schedule: the MFN sites as in MFM, the generative half's sites once per pass
    (site + 32 p), keyed for oracle.mfm_missing_forward ("<site>@p")."""
    from emu_ops import keep_mask
    from factorized_b200 import engine as E
    c = configs[0]
    masks = {}
    for p in range(4):
        sfx = "" if p == 0 else "@%d" % p
        masks["fy" + sfx] = keep_mask(rng, E.SITE_FY + 32 * p, c["zy_to_fy_dropout"], n, c["fy_size"])
        masks["fl" + sfx] = keep_mask(rng, E.SITE_FL + 32 * p, c["zl_to_fl_dropout"], n, c["fl_size"])
        masks["fa" + sfx] = keep_mask(rng, E.SITE_FA + 32 * p, c["za_to_fa_dropout"], n, c["fa_size"])
        masks["fv" + sfx] = keep_mask(rng, E.SITE_FV + 32 * p, c["zv_to_fv_dropout"], n, c["fv_size"])
        masks["y" + sfx] = keep_mask(rng, E.SITE_Y + 32 * p, c["fy_to_y_dropout"], n, c["fy_size"])
    return masks


def test_engine_missing_dropout_masks_replay():
    """train=True with every dropout on: each of the four passes draws its own masks; replayed through the oracle."""
    from emu_ops import keep_mask
    from factorized_b200 import engine as E
    from helpers import tiny_missing_case
    from factorized_b200.missing import MissingEngine
    g, configs, P, x, y, noise, T, n = tiny_missing_case()
    configs = [dict(c) for c in configs]
    configs[0].update(zy_to_fy_dropout=0.3, zl_to_fl_dropout=0.2, za_to_fa_dropout=0.5, zv_to_fv_dropout=0.4,
                      fy_to_y_dropout=0.25)
    for c, p in zip(configs[1:5], (0.5, 0.3, 0.2, 0.4)):
        c["drop"] = p
    rng = torch.tensor([999, 11], dtype=torch.int64)
    P = OrderedDict(P)
    eng = MissingEngine(configs, T, n, "cpu", EmuOps(), head="l1")
    out = eng.forward(P, x.contiguous(), noise, train=True, rng=rng)
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    nn1, nn2, g1, g2 = configs[1:5]
    masks = missing_masks(rng, configs, n)
    masks.update(
        att1=keep_mask(rng, E.SITE_ATT1, nn1["drop"], T * n, nn1["shapes"]).view(T, n, -1),
        att2=keep_mask(rng, E.SITE_ATT2, nn2["drop"], T * n, nn2["shapes"]).view(T, n, -1),
        gamma1=keep_mask(rng, E.SITE_G1, g1["drop"], T * n, g1["shapes"]).view(T, n, -1),
        gamma2=keep_mask(rng, E.SITE_G2, g2["drop"], T * n, g2["shapes"]).view(T, n, -1))
    assert not torch.equal(masks["fy"], masks["fy@1"])
    newP, losses, Go, outo = O.train_step(P, x, y, configs, noise, {}, train=True, masks=masks, variant="missing")
    for k in ("y_hat", "y_hat_nol", "y_hat_noa", "y_hat_nov", "x_a_hat_noa"):
        got = out[k] if k.startswith("y_hat") else out[k].reshape(T, n, -1)
        assert rel_l2(got, outo[k]) < 1e-4, k
    assert abs(float(eng.loss_buf[8]) - losses["total"]) < 1e-4 * abs(losses["total"])
    bad = [(k, rel_l2(G[k], Go[k])) for k in P if Go[k] is not None and rel_l2(G[k], Go[k]) > 3e-4]
    assert not bad, bad


def test_missing_module_init_and_trainer_step_match_reference_golden():
    from helpers import tiny_missing_case
    from factorized_b200.missing import MFM_missing, MissingEngine
    from factorized_b200.train import MFMTrainer
    g, configs, P, x, y, noise, T, n = tiny_missing_case()
    torch.manual_seed(int(g["meta"][0]))
    model = MFM_missing(*configs).eval()
    sd = model.state_dict()
    assert list(sd) == list(P) and len(P) == 126
    for k in P:
        assert torch.equal(sd[k], P[k]), k
    tr = MFMTrainer(model, T, n, head="l1", _test_ops=EmuOps())
    assert isinstance(tr.eng, MissingEngine)
    tr.ops.randn = lambda *a, **k: None
    for k in range(4):
        tr.noise[k].copy_(noise[k])
    lb = tr.step(x, y)
    assert abs(float(lb[8]) - float(g["loss/total"])) < 1e-4 * abs(float(g["loss/total"]))
    worst = 0.0
    for k, v in model.state_dict().items():
        d_ref, d_got = torch.from_numpy(g["p1/" + k]) - P[k], v - P[k]
        if float(d_ref.norm()) == 0.0:
            assert float(d_got.abs().max()) == 0.0, k
            continue
        worst = max(worst, float((d_got - d_ref).norm() / d_ref.norm()))
    assert worst < 2e-3, worst


@pytest.mark.parametrize("variant", ["m_a", "m_b", "m_c", "m_d", "missing"])
def test_rng_replay_helper_reads_the_variant_engines(variant):
    """oracle.rng_replay.train_masks_and_branches (what the GPU train-mode tests and bench.py's parity check use to replay a CUDA
    step's dropout masks and ReLU branches in the oracle) against the ablation and MFM_missing schedules, on the test double."""
    from helpers import tiny_ablation_case, tiny_missing_case
    from oracle.rng_replay import train_masks_and_branches
    from factorized_b200.ablations import make_engine
    g, configs, P, x, y, noise, T, n = tiny_missing_case() if variant == "missing" else tiny_ablation_case(variant)
    configs = [dict(c) for c in configs]
    configs[0].update(zy_to_fy_dropout=0.3, zl_to_fl_dropout=0.2, za_to_fa_dropout=0.5, zv_to_fv_dropout=0.4,
                      fy_to_y_dropout=0.25)
    for c, p in zip(configs[1:5], (0.5, 0.3, 0.2, 0.4)):
        c["drop"] = p
    rng = torch.tensor([31337, 5], dtype=torch.int64)
    P = OrderedDict(P)
    eng = make_engine(configs, T, n, "cpu", EmuOps(), head="l1", variant=variant)
    eng.forward(P, x.contiguous(), noise, train=True, rng=rng)
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    masks, br = train_masks_and_branches(eng, rng)
    onoise = [None if v.shape[0] != n else v for v in noise]
    del O.RELU_REPLAY_VIOLATIONS[:]
    newP, losses, Go, outo = O.train_step(P, x, y, configs, onoise, {}, head="l1", train=True, masks=masks, branches=br,
                                          variant=variant)
    assert not O.RELU_REPLAY_VIOLATIONS, O.RELU_REPLAY_VIOLATIONS[:3]
    assert abs(float(eng.loss_buf[8]) - losses["total"]) < 1e-4 * abs(losses["total"])
    bad = [(k, rel_l2(G[k], Go[k])) for k in P if Go[k] is not None and rel_l2(G[k], Go[k]) > 3e-4]
    assert not bad, bad


def test_engine_backward_with_external_latent_gradients():
    """Engine.backward(d_latents=...): the MFM of mfm_mosi_acc.py returns its latents and the loop applies loss_MMD to them itself
    (:394, :441).  Forward without the in-step MMD + the MMD gradient formed outside and handed in must reproduce the golden
    step's gradients."""
    g, configs, P, x, y, noise, T, n = tiny_case("l1", 1)
    P = OrderedDict(P)
    eng = Engine(configs, T, n, "cpu", EmuOps(), head="l1")
    eng.want_mmd = False
    out = eng.forward(P, x.contiguous(), [None] * 4)
    assert float(eng.loss_buf[4:8].abs().sum()) == 0.0
    dX, dY = eng.losses(y)
    lda = configs[0]["lda_mmd"]
    dlat = []
    for k, key in enumerate(("zl", "za", "zv", "zy")):
        z = out[key].clone().requires_grad_(True)
        (lda * O.loss_mmd(z, noise[k])).backward()
        dlat.append(z.grad)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, 0.0, d_latents=dlat)
    bad = [(k, rel_l2(G[k], g["g/" + k])) for k in P if "g/" + k in g and rel_l2(G[k], g["g/" + k]) > 2e-4]
    assert not bad, bad


def _random_configs(seed):
    """Small configurations drawn from the shape of the reference's search space (mfm_mosi.py:1304-1351): unequal modality
    widths, cell sizes, latent / factor sizes, memory and hidden-layer widths; every dropout on."""
    import random
    r = random.Random(seed)
    pick = lambda lo, hi: r.randint(lo, hi)
    config = dict(input_dims=[pick(5, 9), pick(2, 4), pick(3, 6)], h_dims=[pick(3, 7), pick(2, 5), pick(2, 6)],
                  zy_size=pick(2, 6), zl_size=pick(2, 7), za_size=pick(2, 4), zv_size=pick(3, 8),
                  fy_size=pick(2, 5), fl_size=pick(2, 6), fa_size=pick(2, 4), fv_size=pick(2, 5), memsize=pick(3, 9),
                  zy_to_fy_dropout=0.3, zl_to_fl_dropout=0.2, za_to_fa_dropout=0.4, zv_to_fv_dropout=0.1, fy_to_y_dropout=0.25,
                  lda_mmd=r.choice([0.5, 1.0, 2.0]), lda_xl=r.choice([0.1, 1.0]), lda_xa=r.choice([0.01, 0.5]), lda_xv=r.choice([0.5, 2.0]),
                  missing=0, windowsize=2, batchsize=4, num_epochs=1, lr=0.01, momentum=0.9, output_dim=1, type="mfm")
    nn_ = lambda: dict(shapes=pick(4, 12), drop=r.choice([0.2, 0.5]))
    return [config, nn_(), nn_(), nn_(), nn_(), nn_()], pick(2, 5), pick(3, 9)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("variant", ["mfm", "kl", "kl_ef", "m_a", "m_b", "m_c", "m_d", "missing"])
def test_every_schedule_on_random_configurations(variant, seed):
    """All eight schedules (MFM, MFM_KL, MFM_KL_EF, M_A..M_D, MFM_missing) on randomly drawn small configurations, train mode with
    every dropout on: losses and all gradients against the oracle's autograd with the masks and ReLU branches replayed."""
    from oracle.rng_replay import train_masks_and_branches
    from factorized_b200.ablations import make_engine
    configs, T, n = _random_configs(100 * seed + len(variant))
    configs[0]["type"] = variant
    P = OrderedDict(O.init_params(configs, seed, variant=variant))
    x, y = O.synthetic_batch(configs, T, n, seed + 10)
    noise = O.draw_mmd_noise(configs, n, seed + 20, variant=variant if variant.startswith("m_") else "mfm")
    rng = torch.tensor([77 + seed, 2], dtype=torch.int64)
    eng = make_engine(configs, T, n, "cpu", EmuOps(), head="l1", variant=variant)
    eng.forward(P, x.contiguous(), [torch.zeros(1, 1) if v is None else v for v in noise], train=True, rng=rng)
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    masks, br = train_masks_and_branches(eng, rng)
    del O.RELU_REPLAY_VIOLATIONS[:]
    _, losses, Go, _ = O.train_step(P, x, y, configs, noise, {}, head="l1", train=True, masks=masks, branches=br, variant=variant)
    assert not O.RELU_REPLAY_VIOLATIONS
    assert abs(float(eng.loss_buf[8]) - losses["total"]) < 1e-4 * abs(losses["total"]), (float(eng.loss_buf[8]), losses["total"])
    bad = [(k, rel_l2(G[k], Go[k])) for k in P if Go[k] is not None and float(Go[k].abs().max()) > 0 and rel_l2(G[k], Go[k]) > 5e-4]
    assert not bad, bad
    assert all(float(G[k].abs().max()) == 0.0 for k in P if Go[k] is None)


@pytest.mark.parametrize("T,n", [(1, 3), (3, 1), (1, 1)])
@pytest.mark.parametrize("variant", ["mfm", "missing", "m_c"])
def test_degenerate_sequence_length_and_batch(variant, T, n):
    """One time step (no c_{t-1} block from a previous step, no second half of the attention gradient) and a batch of one row
    (the MMD of a single sample): the schedules against the oracle's autograd."""
    from factorized_b200.ablations import make_engine
    configs = O.tiny_configs()
    configs[0]["type"] = variant
    P = OrderedDict(O.init_params(configs, 9, variant=variant))
    x, y = O.synthetic_batch(configs, T, n, 12)
    noise = O.draw_mmd_noise(configs, n, 13, variant=variant if variant.startswith("m_") else "mfm")
    eng = make_engine(configs, T, n, "cpu", EmuOps(), head="l1", variant=variant)
    eng.forward(P, x.contiguous(), [torch.zeros(1, 1) if v is None else v for v in noise])
    dX, dY = eng.losses(y)
    G = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    eng.backward(P, G, dX, dY, eng.dm.lda_mmd)
    _, losses, Go, _ = O.train_step(P, x, y, configs, noise, {}, head="l1", variant=variant)
    assert abs(float(eng.loss_buf[8]) - losses["total"]) < 1e-4 * abs(losses["total"])
    bad = [(k, rel_l2(G[k], Go[k])) for k in P if Go[k] is not None and float(Go[k].abs().max()) > 1e-12 and rel_l2(G[k], Go[k]) > 5e-4]
    assert not bad, bad
