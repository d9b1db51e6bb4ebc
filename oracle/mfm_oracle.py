"""CPU oracle for the MFM training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this file.  The product
(``factorized_b200``) never does; it fails loudly without its CUDA library.

This is a functional, plain-PyTorch-on-CPU fp32 (or fp64) restatement of the
reference's Multimodal Factorization Model.  The reference
(pliang279/factorized) is Python calling torch ops; its arithmetic lives in the
third-party dependency PyTorch ("PyTorch 0.4.0", /root/reference/README.md:19;
installed here: torch 2.11).  The op semantics this file restates are the
published ones: LSTMCell gate order i,f,g,o with
``c' = sig(f)*c + sig(i)*tanh(g); h' = sig(o)*tanh(c')``; Linear = x W^T + b;
MSELoss/L1Loss/CrossEntropyLoss with reduction='mean'; Adam with bias
correction, eps outside the sqrt(v_hat).

Pinning: the reference ships no tests or golden vectors (SURVEY.md section 4,
8c), so the pin is the reference itself, imported unmodified from
/root/reference in the build container by ``oracle/make_golden.py``; the
vectors it produced are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks this restatement against them.

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor

# --------------------------------------------------------------------------
# configuration helpers
# --------------------------------------------------------------------------

def best_acc_configs(input_dims=(300, 5, 20), output_dim=1, dropout=True):
    """The only fixed MFM hyper-parameter set in the reference
    (mfm_mosi.py:1239-1286, function best_acc)."""
    dr = (lambda p: p) if dropout else (lambda p: 0.0)
    config = dict(
        input_dims=list(input_dims), h_dims=[88, 64, 48],
        zy_size=32, zl_size=32, za_size=8, zv_size=80,
        fy_size=16, fl_size=88, fa_size=8, fv_size=8, memsize=64,
        zy_to_fy_dropout=dr(0.0), zl_to_fl_dropout=dr(0.2), za_to_fa_dropout=dr(0.2),
        zv_to_fv_dropout=dr(0.7), fy_to_y_dropout=dr(0.0),
        lda_mmd=1.0, lda_xl=1.0, lda_xa=0.01, lda_xv=0.5,
        missing=0, windowsize=2, batchsize=32, num_epochs=30, lr=0.01, momentum=0.9,
        output_dim=output_dim, type="mfm",
    )
    nn_ = lambda s: dict(shapes=s, drop=dr(0.5))
    return [config, nn_(128), nn_(128), nn_(128), nn_(128), nn_(64)]


def tiny_configs(output_dim=1):
    """A deliberately awkward small configuration (non-multiples of 8/16
    everywhere) used for golden vectors and edge-case tests."""
    config = dict(
        input_dims=[7, 3, 5], h_dims=[6, 5, 4],
        zy_size=5, zl_size=6, za_size=3, zv_size=7,
        fy_size=4, fl_size=6, fa_size=3, fv_size=2, memsize=9,
        zy_to_fy_dropout=0.0, zl_to_fl_dropout=0.0, za_to_fa_dropout=0.0,
        zv_to_fv_dropout=0.0, fy_to_y_dropout=0.0,
        lda_mmd=0.7, lda_xl=1.0, lda_xa=0.3, lda_xv=0.5,
        missing=0, windowsize=2, batchsize=6, num_epochs=1, lr=0.01, momentum=0.9,
        output_dim=output_dim, type="mfm",
    )
    nn_ = lambda s: dict(shapes=s, drop=0.0)
    return [config, nn_(10), nn_(11), nn_(12), nn_(13), nn_(6)]


def param_shapes(configs, variant: str = "mfm") -> "OrderedDict[str, Tuple[int, ...]]":
    """state_dict names and shapes in the reference's construction order
    (mfm_model.py:491-520 for MFM; :43-44 encoderLSTM; :67-68 decoderLSTM;
    :116-137 MFN).  90 tensors.  variant "kl": MFM_KL (mfm_model.py:683-721), 104 tensors.  variant "kl_ef": MFM_KL_EF
    (mfm_model.py:579-619): the MFN encoder is replaced by ONE early-fusion encoderLSTM over the concatenated input."""
    config, nn1, nn2, g1, g2, out = configs
    d = config["input_dims"]
    hm = config["h_dims"]
    z = [config["zl_size"], config["za_size"], config["zv_size"]]
    f = [config["fl_size"], config["fa_size"], config["fv_size"]]
    fy, zy, mem = config["fy_size"], config["zy_size"], config["memsize"]
    H = sum(hm)
    shapes: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def lstm(prefix, din, h):
        shapes[prefix + ".weight_ih"] = (4 * h, din)
        shapes[prefix + ".weight_hh"] = (4 * h, h)
        shapes[prefix + ".bias_ih"] = (4 * h,)
        shapes[prefix + ".bias_hh"] = (4 * h,)

    def lin(prefix, din, dout):
        shapes[prefix + ".weight"] = (dout, din)
        shapes[prefix + ".bias"] = (dout,)

    if variant in ABLATIONS:
        return _ablation_shapes(configs, variant, lstm, lin, shapes)
    for m, tag in enumerate("lav"):
        lstm("encoder_%s.lstm" % tag, d[m], z[m])
        lin("encoder_%s.fc1" % tag, z[m], z[m])
    if variant == "missing":                                  # MFM_missing, mfm_model.py:792-798: six cross-modal encoders
        for name, (a, b), zo in MISSING_ENCODERS(config):
            lstm(name + ".lstm", d[a] + d[b], zo)
            lin(name + ".fc1", zo, zo)
    for m, tag in enumerate("lav"):
        hd = fy + f[m]
        lstm("decoder_%s.lstm" % tag, hd, hd)
        lin("decoder_%s.fc1" % tag, hd, d[m])
    if variant == "kl_ef":                                    # mfm_model.py:587-599
        ef = sum(z)
        lstm("ef_encoder.lstm", sum(d), ef)
        lin("ef_encoder.fc1", ef, ef)
        lin("last_to_zy_fc1", ef, zy)
        lin("last_to_logvarzy_fc1", ef, zy)
        for m, tag in enumerate("lav"):
            lin("last_to_z%s_fc1" % tag, z[m], z[m])
        for m, tag in enumerate("lav"):
            lin("last_to_logvarz%s_fc1" % tag, z[m], z[m])
        lin("zy_to_fy_fc1", zy, fy)
        lin("zy_to_fy_fc2", fy, fy)
        for m, tag in enumerate("lav"):
            lin("z%s_to_f%s_fc1" % (tag, tag), z[m], f[m])
            lin("z%s_to_f%s_fc2" % (tag, tag), f[m], f[m])
        lin("fy_to_y_fc1", fy, fy)
        lin("fy_to_y_fc2", fy, config["output_dim"])
        return shapes
    for m, tag in enumerate("lav"):
        lstm("mfn_encoder.lstm_%s" % tag, d[m], hm[m])
    att_in = H * config["windowsize"]
    gam_in = att_in + mem
    lin("mfn_encoder.att1_fc1", att_in, nn1["shapes"])
    lin("mfn_encoder.att1_fc2", nn1["shapes"], att_in)
    lin("mfn_encoder.att2_fc1", att_in, nn2["shapes"])
    lin("mfn_encoder.att2_fc2", nn2["shapes"], mem)
    lin("mfn_encoder.gamma1_fc1", gam_in, g1["shapes"])
    lin("mfn_encoder.gamma1_fc2", g1["shapes"], mem)
    lin("mfn_encoder.gamma2_fc1", gam_in, g2["shapes"])
    lin("mfn_encoder.gamma2_fc2", g2["shapes"], mem)
    lin("mfn_encoder.out_fc1", H + mem, out["shapes"])          # constructed, never used (mfm_model.py:136-137)
    lin("mfn_encoder.out_fc2", out["shapes"], config["output_dim"])
    lin("last_to_zy_fc1", H + mem, zy)
    if variant == "kl":                                       # mfm_model.py:696-704
        lin("last_to_logvarzy_fc1", H + mem, zy)
        for m, tag in enumerate("lav"):
            lin("last_to_z%s_fc1" % tag, z[m], z[m])
        for m, tag in enumerate("lav"):
            lin("last_to_logvarz%s_fc1" % tag, z[m], z[m])
    lin("zy_to_fy_fc1", zy, fy)
    lin("zy_to_fy_fc2", fy, fy)
    for m, tag in enumerate("lav"):
        lin("z%s_to_f%s_fc1" % (tag, tag), z[m], f[m])
        lin("z%s_to_f%s_fc2" % (tag, tag), f[m], f[m])
    lin("fy_to_y_fc1", fy, fy)
    lin("fy_to_y_fc2", fy, config["output_dim"])
    return shapes


ABLATIONS = ("m_a", "m_b", "m_c", "m_d")


def MISSING_ENCODERS(config):
    """The six cross-modal encoders of MFM_missing in construction order (mfm_model.py:792-798): (name, the two modalities whose
    columns it reads, output size).  encoder_XY_to_Z infers the latent of the missing modality Z from the other two."""
    zl, za, zv, zy = config["zl_size"], config["za_size"], config["zv_size"], config["zy_size"]
    return [("encoder_la_to_v", (0, 1), zv), ("encoder_lv_to_a", (0, 2), za), ("encoder_av_to_l", (1, 2), zl),
            ("encoder_la_to_y", (0, 1), zy), ("encoder_lv_to_y", (0, 2), zy), ("encoder_av_to_y", (1, 2), zy)]


def _ablation_shapes(configs, variant, lstm, lin, shapes):
    """Construction order of the ablation models M_A (mfm_model.py:223-242), M_B (:293-315), M_C (:367-380) and
    M_D (:427-443) -- the models train_mfm_ablation builds (mfm_mosi.py:651-658)."""
    config, nn1, nn2, g1, g2, out = configs
    d, hm = config["input_dims"], config["h_dims"]
    z = [config["zl_size"], config["za_size"], config["zv_size"]]
    f = [config["fl_size"], config["fa_size"], config["fv_size"]]
    fy, zy, mem, od = config["fy_size"], config["zy_size"], config["memsize"], config["output_dim"]
    H = sum(hm)

    def mfn():
        for m, tag in enumerate("lav"):
            lstm("mfn_encoder.lstm_%s" % tag, d[m], hm[m])
        att_in = H * config["windowsize"]
        gam_in = att_in + mem
        lin("mfn_encoder.att1_fc1", att_in, nn1["shapes"])
        lin("mfn_encoder.att1_fc2", nn1["shapes"], att_in)
        lin("mfn_encoder.att2_fc1", att_in, nn2["shapes"])
        lin("mfn_encoder.att2_fc2", nn2["shapes"], mem)
        lin("mfn_encoder.gamma1_fc1", gam_in, g1["shapes"])
        lin("mfn_encoder.gamma1_fc2", g1["shapes"], mem)
        lin("mfn_encoder.gamma2_fc1", gam_in, g2["shapes"])
        lin("mfn_encoder.gamma2_fc2", g2["shapes"], mem)
        lin("mfn_encoder.out_fc1", H + mem, out["shapes"])
        lin("mfn_encoder.out_fc2", out["shapes"], od)
        lin("last_to_zy_fc1", H + mem, zy)

    def decoders(hd):
        for m, tag in enumerate("lav"):
            lstm("decoder_%s.lstm" % tag, hd[m], hd[m])
            lin("decoder_%s.fc1" % tag, hd[m], d[m])

    def mlp(src, dst, zin, fout):
        lin("z%s_to_f%s_fc1" % (src, dst), zin, fout)
        lin("z%s_to_f%s_fc2" % (src, dst), fout, fout)

    if variant == "m_a":                       # ONE encoder over the whole input; the three decoders read cat(fy, fl)
        lstm("encoder_l.lstm", sum(d), z[0])
        lin("encoder_l.fc1", z[0], z[0])
        decoders([fy + f[0]] * 3)
        mfn()
        mlp("y", "y", zy, fy)
        mlp("l", "l", z[0], f[0])
        lin("fy_to_y_fc1", fy, fy)
        lin("fy_to_y_fc2", fy, od)
    elif variant == "m_b":                     # no MFN / no z_y: the label is predicted from cat(fl, fa, fv)
        for m, tag in enumerate("lav"):
            lstm("encoder_%s.lstm" % tag, d[m], z[m])
            lin("encoder_%s.fc1" % tag, z[m], z[m])
        decoders(f)
        for m, tag in enumerate("lav"):
            mlp(tag, tag, z[m], f[m])
        lin("fy_to_y_fc1", sum(f), fy)
        lin("fy_to_y_fc2", fy, od)
    elif variant == "m_c":                     # MFN only: the three decoders read fy
        decoders([fy] * 3)
        mfn()
        mlp("y", "y", zy, fy)
        lin("fy_to_y_fc1", fy, fy)
        lin("fy_to_y_fc2", fy, od)
    else:                                      # m_d: purely discriminative (no decoders, no regulariser)
        for m, tag in enumerate("lav"):
            lstm("encoder_%s.lstm" % tag, d[m], z[m])
            lin("encoder_%s.fc1" % tag, z[m], z[m])
        for m, tag in enumerate("lav"):
            mlp(tag, tag, z[m], f[m])
        lin("fs_to_y", sum(f), od)
    return shapes


UNUSED_PARAMS = ("mfn_encoder.out_fc1.weight", "mfn_encoder.out_fc1.bias",
                 "mfn_encoder.out_fc2.weight", "mfn_encoder.out_fc2.bias")


def init_params(configs, seed: int, dtype=torch.float32, variant: str = "mfm") -> "OrderedDict[str, Tensor]":
    """Parameters as ``torch.manual_seed(seed); MFM(*configs)`` would draw them.

    The reference constructs nn.LSTMCell / nn.Linear in the order of
    ``param_shapes``; torch's reset_parameters draws, per module and in
    parameter order, U(-1/sqrt(h), 1/sqrt(h)) for every LSTMCell tensor and
    kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for Linear
    weight then bias.  We instantiate the same torch modules so the RNG stream
    is consumed identically (checked against the reference in
    oracle/make_golden.py)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    shapes = param_shapes(configs, variant)
    names = list(shapes)
    i = 0
    while i < len(names):
        n = names[i]
        if n.endswith(".weight_ih"):
            h4, din = shapes[n]
            cell = nn.LSTMCell(din, h4 // 4)
            pre = n[: -len(".weight_ih")]
            for leaf in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                out[pre + "." + leaf] = getattr(cell, leaf).detach().clone().to(dtype)
            i += 4
        else:
            dout, din = shapes[n]
            lin = nn.Linear(din, dout)
            pre = n[: -len(".weight")]
            out[pre + ".weight"] = lin.weight.detach().clone().to(dtype)
            out[pre + ".bias"] = lin.bias.detach().clone().to(dtype)
            i += 2
    return out


# --------------------------------------------------------------------------
# primitive restatements
# --------------------------------------------------------------------------

def linear(x: Tensor, P: Dict[str, Tensor], name: str) -> Tensor:
    """nn.Linear: y = x W^T + b."""
    return x @ P[name + ".weight"].t() + P[name + ".bias"]


def lstm_cell(x: Tensor, h: Tensor, c: Tensor, P: Dict[str, Tensor], name: str):
    """nn.LSTMCell (gate order i,f,g,o) as called at mfm_model.py:56,83,85,167-169."""
    gates = x @ P[name + ".weight_ih"].t() + P[name + ".bias_ih"] \
        + h @ P[name + ".weight_hh"].t() + P[name + ".bias_hh"]
    hs = h.shape[1]
    i = torch.sigmoid(gates[:, 0 * hs:1 * hs])
    f = torch.sigmoid(gates[:, 1 * hs:2 * hs])
    g = torch.tanh(gates[:, 2 * hs:3 * hs])
    o = torch.sigmoid(gates[:, 3 * hs:4 * hs])
    c2 = f * c + i * g
    h2 = o * torch.tanh(c2)
    return h2, c2


def dropout(x: Tensor, p: float, train: bool, mask: Optional[Tensor] = None) -> Tensor:
    """nn.Dropout.  The oracle takes an explicit keep-mask (1/0) when given so
    CUDA-generated masks can be replayed; otherwise identity unless train and
    p>0 (then torch's own generator, which no other implementation can match)."""
    if mask is not None:
        return x * mask / (1.0 - p)
    if train and p > 0.0:
        return torch.nn.functional.dropout(x, p, True)
    return x


RELU_REPLAY_VIOLATIONS: List[str] = []


def relu(x: Tensor, branches, key: str, t: Optional[int] = None) -> Tensor:
    """F.relu, optionally with the branch decision (x > 0) REPLAYED from another implementation.

    ReLU's derivative is discontinuous: an input that the reference computes as +1e-7 and another
    correct fp32 implementation as -1e-7 yields equal forward values (to 1e-7) but gradients that differ
    by a whole row.  Parity tests therefore replay the branch mask the CUDA run took (like dropout masks)
    and separately assert that any disagreement with the oracle's own decision happens only where
    |x| < 1e-3, i.e. on the knife edge.  With branches=None this is exactly torch.relu."""
    if branches is None or key not in branches:
        return torch.relu(x)
    m = branches[key] if t is None else branches[key][t]
    m = m.to(x.dtype)
    own = (x > 0).to(x.dtype)
    m = torch.where(m < 0, own, m)              # -1 = "not observed" (the unit was dropped): the oracle decides itself
    bad = (own != m) & (x.abs() > 1e-3)
    if bool(bad.any()):
        RELU_REPLAY_VIOLATIONS.append("%s[t=%s]: %d branch decisions differ with |x| > 1e-3" % (key, t, int(bad.sum())))
    return x * m


def encoder_lstm(x: Tensor, P, prefix: str) -> Tensor:
    """encoderLSTM.forward, mfm_model.py:47-62: zero state, T cell steps,
    fc1 on the last hidden state, no activation."""
    T, n, _ = x.shape
    hs = P[prefix + ".lstm.weight_hh"].shape[1]
    h = x.new_zeros(n, hs)
    c = x.new_zeros(n, hs)
    for t in range(T):
        h, c = lstm_cell(x[t], h, c, P, prefix + ".lstm")
    return linear(h, P, prefix + ".fc1")


def eflstm_forward(x: Tensor, P, train=False, p_drop=0.0, keep=None) -> Tensor:
    """EFLSTM.forward, test_mosi.py:139-157 (the early-fusion LSTM baseline): zero state, T cell steps over the whole input
    row, ``fc2(dropout(relu(fc1(h_T))))``.  P holds lstm.*, fc1.*, fc2.*."""
    T, n, _ = x.shape
    hs = P["lstm.weight_hh"].shape[1]
    h = x.new_zeros(n, hs)
    c = x.new_zeros(n, hs)
    for t in range(T):
        h, c = lstm_cell(x[t], h, c, P, "lstm")
    out = torch.relu(linear(h, P, "fc1"))
    if train and p_drop > 0.0:
        out = out * keep / (1.0 - p_drop)
    return linear(out, P, "fc2")


def decoder_lstm(emb: Tensor, T: int, P, prefix: str) -> Tensor:
    """decoderLSTM.forward, mfm_model.py:72-91: step 0 eats the embedding,
    step t>0 eats the previous hidden state; fc1 over all T hiddens."""
    n, hs = emb.shape
    h = emb.new_zeros(n, hs)
    c = emb.new_zeros(n, hs)
    outs = []
    for t in range(T):
        inp = emb if t == 0 else outs[-1]
        h, c = lstm_cell(inp, h, c, P, prefix + ".lstm")
        outs.append(h)
    hs_all = torch.stack(outs, 0)
    return linear(hs_all, P, prefix + ".fc1")


def mfn_encoder(x: Tensor, P, configs, train=False, masks=None, branches=None) -> Tensor:
    """MFN.forward, mfm_model.py:140-199 (the memory fusion network that
    produces the input of last_to_zy_fc1).  ``masks`` optionally maps
    'att1','att2','gamma1','gamma2' -> [T,n,shapes] keep-masks."""
    config, nn1, nn2, g1, g2, _ = configs
    d_l, d_a, d_v = config["input_dims"]
    T, n, _ = x.shape
    xs = (x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:])
    pre = "mfn_encoder."
    hcs = []
    for m, tag in enumerate("lav"):
        hsz = P[pre + "lstm_%s.weight_hh" % tag].shape[1]
        hcs.append([x.new_zeros(n, hsz), x.new_zeros(n, hsz)])
    mem = x.new_zeros(n, config["memsize"])
    mk = (lambda k, t: None if (masks is None or masks.get(k) is None) else masks[k][t])
    for t in range(T):
        prev_cs = torch.cat([hc[1] for hc in hcs], 1)                       # :163-165,171
        for m, tag in enumerate("lav"):
            hcs[m][0], hcs[m][1] = lstm_cell(xs[m][t], hcs[m][0], hcs[m][1], P, pre + "lstm_%s" % tag)  # :167-169
        new_cs = torch.cat([hc[1] for hc in hcs], 1)                        # :172
        c_star = torch.cat([prev_cs, new_cs], 1)                            # :173
        a = relu(linear(c_star, P, pre + "att1_fc1"), branches, "att1", t)
        a = dropout(a, nn1["drop"], train, mk("att1", t))
        attention = torch.softmax(linear(a, P, pre + "att1_fc2"), dim=1)    # :174
        attended = attention * c_star                                        # :175
        b = relu(linear(attended, P, pre + "att2_fc1"), branches, "att2", t)
        b = dropout(b, nn2["drop"], train, mk("att2", t))
        c_hat = torch.tanh(linear(b, P, pre + "att2_fc2"))                  # :176
        both = torch.cat([attended, mem], 1)                                 # :177
        u1 = dropout(relu(linear(both, P, pre + "gamma1_fc1"), branches, "gamma1", t), g1["drop"], train, mk("gamma1", t))
        u2 = dropout(relu(linear(both, P, pre + "gamma2_fc1"), branches, "gamma2", t), g2["drop"], train, mk("gamma2", t))
        gamma1 = torch.sigmoid(linear(u1, P, pre + "gamma1_fc2"))           # :178
        gamma2 = torch.sigmoid(linear(u2, P, pre + "gamma2_fc2"))           # :179
        mem = gamma1 * mem + gamma2 * c_hat                                  # :180
    return torch.cat([hcs[0][0], hcs[1][0], hcs[2][0], mem], 1)             # :194-198


def mfn_baseline_forward(x: Tensor, P, configs, train=False, masks=None, branches=None, prefix="mfn_encoder.") -> Tensor:
    """The MFN baseline of the reference's MOSI script (test_mosi.py:158-265): the MFN of mfm_model.py plus the output head
    ``out_fc2(out_dropout(relu(out_fc1(last_hs))))`` (:264) that mfm_model.MFN constructs and never uses."""
    last = mfn_encoder(x, P, configs, train, masks, branches)
    mk = None if masks is None else masks.get("out")
    h = dropout(relu(linear(last, P, prefix + "out_fc1"), branches, "out1"), configs[5]["drop"], train, mk)
    return linear(h, P, prefix + "out_fc2")


def compute_kernel(x: Tensor, y: Tensor) -> Tensor:
    """mfm_model.py:14-23: exp(-mean_k((x_ik-y_jk)^2)/dim) = exp(-|x_i-y_j|^2/dim^2)."""
    dim = x.shape[1]
    diff = x.unsqueeze(1) - y.unsqueeze(0)
    return torch.exp(-(diff.pow(2).mean(2) / float(dim)))


def loss_mmd(z: Tensor, gauss: Tensor) -> Tensor:
    """mfm_model.py:25-34 with the Gaussian sample injected instead of drawn
    (the reference draws torch.randn(z.size()) on the CPU default generator)."""
    return compute_kernel(gauss, gauss).mean() + compute_kernel(z, z).mean() \
        - 2.0 * compute_kernel(gauss, z).mean()


def draw_mmd_noise(configs, n: int, seed: int, dtype=torch.float32, variant: str = "mfm") -> List[Tensor]:
    """The four draws loss_MMD makes inside MFM.forward, in order zl, za, zv, zy
    (mfm_model.py:536), after torch.manual_seed(seed).  The ablation models regularise fewer latents (M_A zl, zy :255;
    M_B zl, za, zv :328; M_C zy :392; M_D none :456): the list keeps four slots, None where nothing is drawn."""
    c = configs[0]
    torch.manual_seed(seed)
    sizes = (c["zl_size"], c["za_size"], c["zv_size"], c["zy_size"])
    used = dict(m_a=(0, 3), m_b=(0, 1, 2), m_c=(3,), m_d=()).get(variant, (0, 1, 2, 3))
    return [torch.randn(n, k).to(dtype) if i in used else None for i, k in enumerate(sizes)]


def factor_mlp(z: Tensor, P, name: str, p: float, train: bool, mask=None, branches=None, key="") -> Tensor:
    """relu(fc2(drop(relu(fc1(z))))), mfm_model.py:539-542."""
    h = dropout(relu(linear(z, P, name + "_fc1"), branches, key + "1"), p, train, mask)
    return relu(linear(h, P, name + "_fc2"), branches, key)


def mfm_forward(x: Tensor, P, configs, noise: Sequence[Tensor], train=False, masks=None, branches=None):
    """MFM.forward, mfm_model.py:522-555.  Returns a dict with the reference's
    outputs plus the latents the reference computes but does not return."""
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    T = x.shape[0]
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    zl = encoder_lstm(x_l, P, "encoder_l")                                   # :530
    za = encoder_lstm(x_a, P, "encoder_a")
    zv = encoder_lstm(x_v, P, "encoder_v")
    mfn_last = mfn_encoder(x, P, configs, train, masks, branches)            # :534
    zy = linear(mfn_last, P, "last_to_zy_fc1")                               # :535
    mmd = loss_mmd(zl, noise[0]) + loss_mmd(za, noise[1]) + loss_mmd(zv, noise[2]) + loss_mmd(zy, noise[3])  # :536
    mk = (lambda k: None if masks is None else masks.get(k))
    fy = factor_mlp(zy, P, "zy_to_fy", config["zy_to_fy_dropout"], train, mk("fy"), branches, "fy")
    fl = factor_mlp(zl, P, "zl_to_fl", config["zl_to_fl_dropout"], train, mk("fl"), branches, "fl")
    fa = factor_mlp(za, P, "za_to_fa", config["za_to_fa_dropout"], train, mk("fa"), branches, "fa")
    fv = factor_mlp(zv, P, "zv_to_fv", config["zv_to_fv_dropout"], train, mk("fv"), branches, "fv")
    x_l_hat = decoder_lstm(torch.cat([fy, fl], 1), T, P, "decoder_l")       # :544-551
    x_a_hat = decoder_lstm(torch.cat([fy, fa], 1), T, P, "decoder_a")
    x_v_hat = decoder_lstm(torch.cat([fy, fv], 1), T, P, "decoder_v")
    y1 = dropout(relu(linear(fy, P, "fy_to_y_fc1"), branches, "y1"), config["fy_to_y_dropout"], train, mk("y"))
    y_hat = linear(y1, P, "fy_to_y_fc2")                                     # :552
    return dict(x_l_hat=x_l_hat, x_a_hat=x_a_hat, x_v_hat=x_v_hat, y_hat=y_hat, mmd=mmd,
                zl=zl, za=za, zv=zv, zy=zy, fy=fy, fl=fl, fa=fa, fv=fv, mfn_last=mfn_last)


def loss_kld(mu: Tensor, logvar: Tensor) -> Tensor:
    """mfm_model.py:36-38: -0.5 * sum(1 + logvar - mu^2 - exp(logvar)) -- a SUM over all elements, not a mean."""
    return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())


def mfm_kl_forward(x: Tensor, P, configs, train=False, masks=None, branches=None):
    """MFM_KL.forward, mfm_model.py:723-764: the encoders' outputs pass one more Linear to the means (z) and another to
    the log-variances; there is NO sampling (the means feed the factor MLPs); the regulariser is the KL term, returned in
    the slot MFM uses for the MMD ("mmd" below) and weighted by lda_mmd in the train step (mfm_mosi.py:433)."""
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    T = x.shape[0]
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    lasts = [encoder_lstm(x_l, P, "encoder_l"), encoder_lstm(x_a, P, "encoder_a"), encoder_lstm(x_v, P, "encoder_v")]    # :732-734
    zs = [linear(lasts[m], P, "last_to_z%s_fc1" % tag) for m, tag in enumerate("lav")]                                  # :735-737
    lvs = [linear(lasts[m], P, "last_to_logvarz%s_fc1" % tag) for m, tag in enumerate("lav")]                           # :738-740
    mfn_last = mfn_encoder(x, P, configs, train, masks, branches)                                                        # :742
    zy = linear(mfn_last, P, "last_to_zy_fc1")
    lvy = linear(mfn_last, P, "last_to_logvarzy_fc1")                                                                    # :743-744
    kld = loss_kld(zs[0], lvs[0]) + loss_kld(zs[1], lvs[1]) + loss_kld(zs[2], lvs[2]) + loss_kld(zy, lvy)               # :746
    zl, za, zv = zs
    mk = (lambda k: None if masks is None else masks.get(k))
    fy = factor_mlp(zy, P, "zy_to_fy", config["zy_to_fy_dropout"], train, mk("fy"), branches, "fy")
    fl = factor_mlp(zl, P, "zl_to_fl", config["zl_to_fl_dropout"], train, mk("fl"), branches, "fl")
    fa = factor_mlp(za, P, "za_to_fa", config["za_to_fa_dropout"], train, mk("fa"), branches, "fa")
    fv = factor_mlp(zv, P, "zv_to_fv", config["zv_to_fv_dropout"], train, mk("fv"), branches, "fv")
    x_l_hat = decoder_lstm(torch.cat([fy, fl], 1), T, P, "decoder_l")
    x_a_hat = decoder_lstm(torch.cat([fy, fa], 1), T, P, "decoder_a")
    x_v_hat = decoder_lstm(torch.cat([fy, fv], 1), T, P, "decoder_v")
    y1 = dropout(relu(linear(fy, P, "fy_to_y_fc1"), branches, "y1"), config["fy_to_y_dropout"], train, mk("y"))
    y_hat = linear(y1, P, "fy_to_y_fc2")
    return dict(x_l_hat=x_l_hat, x_a_hat=x_a_hat, x_v_hat=x_v_hat, y_hat=y_hat, mmd=kld,
                zl=zl, za=za, zv=zv, zy=zy, lvl=lvs[0], lva=lvs[1], lvv=lvs[2], lvy=lvy, fy=fy, fl=fl, fa=fa, fv=fv,
                mfn_last=mfn_last)


def mfm_kl_ef_forward(x: Tensor, P, configs, train=False, masks=None, branches=None):
    """MFM_KL_EF.forward, mfm_model.py:621-660: MFM_KL with the MFN encoder replaced by one early-fusion encoderLSTM over
    the whole input (``ef_encoder``, hidden size zl+za+zv); z_y and its log-variance are Linears of its output."""
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    T = x.shape[0]
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    lasts = [encoder_lstm(x_l, P, "encoder_l"), encoder_lstm(x_a, P, "encoder_a"), encoder_lstm(x_v, P, "encoder_v")]    # :629-631
    zs = [linear(lasts[m], P, "last_to_z%s_fc1" % tag) for m, tag in enumerate("lav")]                                  # :632-634
    lvs = [linear(lasts[m], P, "last_to_logvarz%s_fc1" % tag) for m, tag in enumerate("lav")]                           # :635-637
    ef_last = encoder_lstm(x, P, "ef_encoder")                                                                           # :639
    zy = linear(ef_last, P, "last_to_zy_fc1")
    lvy = linear(ef_last, P, "last_to_logvarzy_fc1")                                                                     # :640-641
    kld = loss_kld(zs[0], lvs[0]) + loss_kld(zs[1], lvs[1]) + loss_kld(zs[2], lvs[2]) + loss_kld(zy, lvy)               # :643
    zl, za, zv = zs
    mk = (lambda k: None if masks is None else masks.get(k))
    fy = factor_mlp(zy, P, "zy_to_fy", config["zy_to_fy_dropout"], train, mk("fy"), branches, "fy")
    fl = factor_mlp(zl, P, "zl_to_fl", config["zl_to_fl_dropout"], train, mk("fl"), branches, "fl")
    fa = factor_mlp(za, P, "za_to_fa", config["za_to_fa_dropout"], train, mk("fa"), branches, "fa")
    fv = factor_mlp(zv, P, "zv_to_fv", config["zv_to_fv_dropout"], train, mk("fv"), branches, "fv")
    x_l_hat = decoder_lstm(torch.cat([fy, fl], 1), T, P, "decoder_l")
    x_a_hat = decoder_lstm(torch.cat([fy, fa], 1), T, P, "decoder_a")
    x_v_hat = decoder_lstm(torch.cat([fy, fv], 1), T, P, "decoder_v")
    y1 = dropout(relu(linear(fy, P, "fy_to_y_fc1"), branches, "y1"), config["fy_to_y_dropout"], train, mk("y"))
    y_hat = linear(y1, P, "fy_to_y_fc2")
    return dict(x_l_hat=x_l_hat, x_a_hat=x_a_hat, x_v_hat=x_v_hat, y_hat=y_hat, mmd=kld,
                zl=zl, za=za, zv=zv, zy=zy, lvl=lvs[0], lva=lvs[1], lvv=lvs[2], lvy=lvy, fy=fy, fl=fl, fa=fa, fv=fv)


def ablation_forward(x: Tensor, P, configs, noise: Sequence[Tensor], variant: str, train=False, masks=None, branches=None):
    """M_A / M_B / M_C / M_D .forward (mfm_model.py:244-269, 317-343, 382-403, 445-467): MFM with parts removed.  Returns
    the same dict as mfm_forward (latents a model does not have are absent; "mmd" is the python float 0.0 for M_D)."""
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    T = x.shape[0]
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    mk = (lambda k: None if masks is None else masks.get(k))
    out = {}

    def fmlp(zk, name, pkey, key):
        return factor_mlp(zk, P, name, config[pkey], train, mk(key), branches, key)

    def y_head(fin):
        y1 = dropout(relu(linear(fin, P, "fy_to_y_fc1"), branches, "y1"), config["fy_to_y_dropout"], train, mk("y"))
        return linear(y1, P, "fy_to_y_fc2")

    if variant in ("m_a", "m_c"):
        if variant == "m_a":
            out["zl"] = zl = encoder_lstm(x, P, "encoder_l")                                  # :252 (the WHOLE input)
        out["mfn_last"] = mfn_last = mfn_encoder(x, P, configs, train, masks, branches)       # :253 / :390
        out["zy"] = zy = linear(mfn_last, P, "last_to_zy_fc1")
        if variant == "m_a":
            out["mmd"] = loss_mmd(zl, noise[0]) + loss_mmd(zy, noise[3])                      # :255
        else:
            out["mmd"] = loss_mmd(zy, noise[3])                                               # :392
        out["fy"] = fy = fmlp(zy, "zy_to_fy", "zy_to_fy_dropout", "fy")
        emb = fy
        if variant == "m_a":
            out["fl"] = fl = fmlp(zl, "zl_to_fl", "zl_to_fl_dropout", "fl")
            emb = torch.cat([fy, fl], 1)                                                      # :260
        out["x_l_hat"] = decoder_lstm(emb, T, P, "decoder_l")                                 # :263-265 / :397-399
        out["x_a_hat"] = decoder_lstm(emb, T, P, "decoder_a")
        out["x_v_hat"] = decoder_lstm(emb, T, P, "decoder_v")
        out["y_hat"] = y_head(fy)
        return out
    out["zl"] = zl = encoder_lstm(x_l, P, "encoder_l")                                        # :325-327 / :453-455
    out["za"] = za = encoder_lstm(x_a, P, "encoder_a")
    out["zv"] = zv = encoder_lstm(x_v, P, "encoder_v")
    out["fl"] = fl = fmlp(zl, "zl_to_fl", "zl_to_fl_dropout", "fl")
    out["fa"] = fa = fmlp(za, "za_to_fa", "za_to_fa_dropout", "fa")
    out["fv"] = fv = fmlp(zv, "zv_to_fv", "zv_to_fv_dropout", "fv")
    fs = torch.cat([fl, fa, fv], 1)
    if variant == "m_b":
        out["mmd"] = loss_mmd(zl, noise[0]) + loss_mmd(za, noise[1]) + loss_mmd(zv, noise[2])  # :328
        out["x_l_hat"] = decoder_lstm(fl, T, P, "decoder_l")                                  # :336-338
        out["x_a_hat"] = decoder_lstm(fa, T, P, "decoder_a")
        out["x_v_hat"] = decoder_lstm(fv, T, P, "decoder_v")
        out["y_hat"] = y_head(fs)                                                             # :339-340
    else:
        out["mmd"] = 0.0                                                                      # :456
        out["x_l_hat"], out["x_a_hat"], out["x_v_hat"] = x_l, x_a, x_v                        # :465 the inputs themselves
        out["y_hat"] = linear(fs, P, "fs_to_y")                                               # :464
    return out


def seq2seq_forward(x: Tensor, P, configs, noise: Sequence[Tensor], train=False):
    """seq2seq.forward, mfm_model.py:929-958: every modality reconstructed from the other two.  ``noise``: the Gaussian samples
    of the three loss_MMD calls in the reference's order (zv_nov, za_noa, zl_nol; :942).  Dropout off (train=False) only."""
    assert not train
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    T = x.shape[0]
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    zv_nov = encoder_lstm(torch.cat([x_l, x_a], 2), P, "encoder_la_to_v")
    za_noa = encoder_lstm(torch.cat([x_l, x_v], 2), P, "encoder_lv_to_a")
    zl_nol = encoder_lstm(torch.cat([x_a, x_v], 2), P, "encoder_av_to_l")
    mmd = loss_mmd(zv_nov, noise[0]) + loss_mmd(za_noa, noise[1]) + loss_mmd(zl_nol, noise[2])
    fl = factor_mlp(zl_nol, P, "zl_to_fl", 0.0, False)
    fa = factor_mlp(za_noa, P, "za_to_fa", 0.0, False)
    fv = factor_mlp(zv_nov, P, "zv_to_fv", 0.0, False)
    return dict(x_l_hat_nol=decoder_lstm(fl, T, P, "decoder_l"), x_a_hat_noa=decoder_lstm(fa, T, P, "decoder_a"),
                x_v_hat_nov=decoder_lstm(fv, T, P, "decoder_v"), mmd=mmd)


def basic_missing_forward(x: Tensor, P, configs, noise: Sequence[Tensor], train=False):
    """basic_missing.forward, mfm_model.py:998-1017: the label from two modalities.  ``noise`` in the order zy_nov, zy_noa,
    zy_nol (:1011).  Dropout off only."""
    assert not train
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    zy_nov = encoder_lstm(torch.cat([x_l, x_a], 2), P, "encoder_la_to_y")
    zy_noa = encoder_lstm(torch.cat([x_l, x_v], 2), P, "encoder_lv_to_y")
    zy_nol = encoder_lstm(torch.cat([x_a, x_v], 2), P, "encoder_av_to_y")
    mmd = loss_mmd(zy_nov, noise[0]) + loss_mmd(zy_noa, noise[1]) + loss_mmd(zy_nol, noise[2])
    head = lambda z, n: linear(torch.relu(linear(z, P, n + "_fc1")), P, n + "_fc2")
    return dict(y_hat_nol=head(zy_nol, "zy_nol_to_y"), y_hat_noa=head(zy_noa, "zy_noa_to_y"),
                y_hat_nov=head(zy_nov, "zy_nov_to_y"), mmd=mmd)


MISSING_PASSES = ("", "_nol", "_noa", "_nov")


def mfm_missing_forward(x: Tensor, P, configs, noise: Sequence[Tensor], train=False, masks=None, branches=None):
    """MFM_missing.forward, mfm_model.py:827-885: MFM plus six cross-modal encoders that infer the latents of a missing modality
    (and z_y) from the other two, an MSE between inferred and true latents (``missing``), and FOUR passes through the shared
    generative half -- all present, language / acoustic / visual inferred.  Output keys carry the reference's suffixes
    ("x_l_hat_nol", ...).  Dropout: every pass draws its own masks; masks / branches of pass p > 0 are keyed "<site>@p"."""
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    T = x.shape[0]
    xm = [x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]]
    zl = encoder_lstm(xm[0], P, "encoder_l")                                  # :836-838
    za = encoder_lstm(xm[1], P, "encoder_a")
    zv = encoder_lstm(xm[2], P, "encoder_v")
    mfn_last = mfn_encoder(x, P, configs, train, masks, branches)             # :839
    zy = linear(mfn_last, P, "last_to_zy_fc1")
    inf = {}
    for name, (a, b), _ in MISSING_ENCODERS(config):                          # :843-850
        inf[name] = encoder_lstm(torch.cat([xm[a], xm[b]], dim=2), P, name)
    zv_nov, za_noa, zl_nol = inf["encoder_la_to_v"], inf["encoder_lv_to_a"], inf["encoder_av_to_l"]
    zy_nov, zy_noa, zy_nol = inf["encoder_la_to_y"], inf["encoder_lv_to_y"], inf["encoder_av_to_y"]
    mmd = loss_mmd(zl, noise[0]) + loss_mmd(za, noise[1]) + loss_mmd(zv, noise[2]) + loss_mmd(zy, noise[3])   # :852
    F = torch.nn.functional
    missing = F.mse_loss(zv_nov, zv) + F.mse_loss(za_noa, za) + F.mse_loss(zl_nol, zl) \
        + F.mse_loss(zy_nov, zy) + F.mse_loss(zy_noa, zy) + F.mse_loss(zy_nol, zy)                           # :853-858
    out = dict(mmd=mmd, missing=missing, zl=zl, za=za, zv=zv, zy=zy, zl_nol=zl_nol, za_noa=za_noa, zv_nov=zv_nov,
               zy_nol=zy_nol, zy_noa=zy_noa, zy_nov=zy_nov)

    def decode(p, zl_, za_, zv_, zy_):                                        # :860-875
        sfx = "" if p == 0 else "@%d" % p
        mk = (lambda k: None if masks is None else masks.get(k + sfx))
        fy = factor_mlp(zy_, P, "zy_to_fy", config["zy_to_fy_dropout"], train, mk("fy"), branches, "fy" + sfx)
        fl = factor_mlp(zl_, P, "zl_to_fl", config["zl_to_fl_dropout"], train, mk("fl"), branches, "fl" + sfx)
        fa = factor_mlp(za_, P, "za_to_fa", config["za_to_fa_dropout"], train, mk("fa"), branches, "fa" + sfx)
        fv = factor_mlp(zv_, P, "zv_to_fv", config["zv_to_fv_dropout"], train, mk("fv"), branches, "fv" + sfx)
        s = MISSING_PASSES[p]
        out["x_l_hat" + s] = decoder_lstm(torch.cat([fy, fl], 1), T, P, "decoder_l")
        out["x_a_hat" + s] = decoder_lstm(torch.cat([fy, fa], 1), T, P, "decoder_a")
        out["x_v_hat" + s] = decoder_lstm(torch.cat([fy, fv], 1), T, P, "decoder_v")
        y1 = dropout(relu(linear(fy, P, "fy_to_y_fc1"), branches, "y1" + sfx), config["fy_to_y_dropout"], train, mk("y"))
        out["y_hat" + s] = linear(y1, P, "fy_to_y_fc2")

    decode(0, zl, za, zv, zy)                                                 # :876-883
    decode(1, zl_nol, za, zv, zy_nol)
    decode(2, zl, za_noa, zv, zy_noa)
    decode(3, zl, za, zv_nov, zy_nov)
    return out


def mfm_missing_losses(out: Dict[str, Tensor], x: Tensor, y: Tensor, configs) -> Dict[str, Tensor]:
    """Loss of train_mfm_missing's step, mfm_mosi.py:962-982.  Six reconstruction terms -- the three of the all-present pass,
    x_l of the language-inferred pass and x_a AND x_v of the acoustic-inferred pass (the reference reads x_v_hat_noa, :976; the
    visual-inferred pass contributes its label only) -- four L1 label terms, the MMD and the latent-matching MSE."""
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    F = torch.nn.functional
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    mse_l = F.mse_loss(out["x_l_hat"], x_l) + F.mse_loss(out["x_l_hat_nol"], x_l)
    mse_a = F.mse_loss(out["x_a_hat"], x_a) + F.mse_loss(out["x_a_hat_noa"], x_a)
    mse_v = F.mse_loss(out["x_v_hat"], x_v) + F.mse_loss(out["x_v_hat_noa"], x_v)
    gen = config["lda_xl"] * mse_l + config["lda_xa"] * mse_a + config["lda_xv"] * mse_v
    disc = sum(F.l1_loss(out["y_hat" + s].squeeze(1), y) for s in MISSING_PASSES)
    mmd = config["lda_mmd"] * out["mmd"]
    total = disc + gen + mmd + out["missing"]
    return dict(total=total, disc=disc, gen=gen, mmd=mmd, mse_l=mse_l, mse_a=mse_a, mse_v=mse_v, missing=out["missing"])


def mfm_losses(out: Dict[str, Tensor], x: Tensor, y: Tensor, configs, head: str = "l1") -> Dict[str, Tensor]:
    """Loss assembly of the train step: mfm_mosi.py:432-439 (L1 head) and
    mfm_mosi_acc.py:441-451 / mfm_moud.py:495-508 (cross-entropy head)."""
    config = configs[0]
    d_l, d_a, d_v = config["input_dims"]
    F = torch.nn.functional
    x_l, x_a, x_v = x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:]
    mse_l = F.mse_loss(out["x_l_hat"], x_l)
    mse_a = F.mse_loss(out["x_a_hat"], x_a)
    mse_v = F.mse_loss(out["x_v_hat"], x_v)
    gen = config["lda_xl"] * mse_l + config["lda_xa"] * mse_a + config["lda_xv"] * mse_v
    y_hat = out["y_hat"].squeeze(1) if out["y_hat"].shape[1] == 1 else out["y_hat"]
    if head == "l1":
        disc = F.l1_loss(y_hat, y)
    elif head == "ce":
        disc = F.cross_entropy(y_hat, y.long())
    else:
        raise ValueError(head)
    mmd = config["lda_mmd"] * out["mmd"]
    total = disc + gen + mmd + 0.0                                           # missing_loss == 0.0 (mfm_model.py:537)
    return dict(total=total, disc=disc, gen=gen, mmd=mmd, mse_l=mse_l, mse_a=mse_a, mse_v=mse_v)


def adam_step(P, G, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam defaults as used at mfm_mosi.py:403 (no weight decay,
    no amsgrad).  Parameters without a gradient are skipped, like torch."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    b1, b2 = betas
    for k, g in G.items():
        if g is None:
            continue
        m = state.setdefault("m." + k, torch.zeros_like(P[k]))
        v = state.setdefault("v." + k, torch.zeros_like(P[k]))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        P[k] = P[k] - (lr / (1 - b1 ** t)) * (m / denom)
    return P


def train_step(P, x, y, configs, noise, state, head="l1", lr=1e-3, train=False, masks=None, branches=None, variant="mfm"):
    """One iteration of the inner loop of train_mfm (mfm_mosi.py:427-442):
    forward, loss, backward, Adam.  Returns (new params, losses, grads, fwd).  variant "kl": the model is MFM_KL
    (train_mfm's own dispatch, mfm_mosi.py:398-399; `noise` is unused)."""
    Pg = OrderedDict((k, v.detach().clone().requires_grad_(k not in UNUSED_PARAMS)) for k, v in P.items())
    if variant == "kl":
        out = mfm_kl_forward(x, Pg, configs, train=train, masks=masks, branches=branches)
    elif variant == "kl_ef":
        out = mfm_kl_ef_forward(x, Pg, configs, train=train, masks=masks, branches=branches)
    elif variant in ABLATIONS:                                # train_mfm_ablation runs the same step (mfm_mosi.py:677-697)
        out = ablation_forward(x, Pg, configs, noise, variant, train=train, masks=masks, branches=branches)
    elif variant == "missing":                                # train_mfm_missing (mfm_mosi.py:918-982): its own loss
        out = mfm_missing_forward(x, Pg, configs, noise, train=train, masks=masks, branches=branches)
    else:
        out = mfm_forward(x, Pg, configs, noise, train=train, masks=masks, branches=branches)
    losses = mfm_missing_losses(out, x, y, configs) if variant == "missing" else mfm_losses(out, x, y, configs, head)
    losses["total"].backward()
    G = OrderedDict((k, (None if v.grad is None else v.grad.detach().clone())) for k, v in Pg.items())
    newP = adam_step(OrderedDict((k, v.detach().clone()) for k, v in P.items()), G, state, lr=lr)
    return newP, {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in losses.items()}, G, \
        {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}


def synthetic_batch(configs, T: int, n: int, seed: int, head="l1", dtype=torch.float32):
    """Synthetic inputs of the benchmark shape."""
    g = torch.Generator().manual_seed(seed)
    D = sum(configs[0]["input_dims"])
    x = torch.randn(T, n, D, generator=g).to(dtype)
    od = configs[0]["output_dim"]
    if head == "ce":
        y = torch.randint(0, od, (n,), generator=g)
    elif od == 1:
        y = torch.randn(n, generator=g).to(dtype)
    else:
        y = torch.randn(n, od, generator=g).to(dtype)
    return x, y
