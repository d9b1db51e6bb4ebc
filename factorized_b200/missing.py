"""MFM_missing (/root/reference/mfm_model.py:766-885) and the step of ``train_mfm_missing`` (mfm_mosi.py:918-1105) on the same
kernels: MFM plus six cross-modal encoders that infer the latents of a missing modality (and z_y) from the other two, an MSE that
pulls inferred and true latents together, and FOUR passes through the shared generative half (factor MLPs, decoders, label head):
all modalities present, then language / acoustic / visual replaced by its inferred latent.

``MissingEngine`` is the host schedule.  The input side is MFM's (three encoders, the MFN block, the MMD streams -- the engine's
own ``_forward_mfn_head`` / ``_backward_mfn``) plus one more recurrence launch for the six cross-modal cells; the generative half
runs once per pass on per-pass buffers, the four decoder cells of a modality in ONE recurrence launch (same weights, four
states), and the weight gradients of the shared layers accumulate over the passes on the side streams.  Like the engine it does
no arithmetic itself and has no CPU path.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from . import engine as E
from .engine import ACT_RELU, SITE_FL, SITE_FY, SITE_Y, TAGS
from .mfm_model import MFM, MFN, UNUSED, _ops, _require_cuda, decoderLSTM, encoderLSTM, eng_generation

PASSES = ("", "_nol", "_noa", "_nov")          # the reference's suffixes: all present, language / acoustic / visual inferred
SITE_PASS = 32                                 # dropout site of pass p = site + 32 p: every pass draws its own masks
# reconstructions that enter train_mfm_missing's loss (mfm_mosi.py:971-976): (pass, modality).  The acoustic-inferred pass
# contributes x_a AND x_v (the reference reads x_v_hat_noa), the visual-inferred pass its label only.
LOSS_RECON = ((0, 0), (0, 1), (0, 2), (1, 0), (2, 1), (2, 2))


def cross_encoders(dm):
    """(parameter prefix, the two modalities whose columns it reads, output size, name of the inferred latent) in the reference's
    construction order (mfm_model.py:792-798)."""
    return [("encoder_la_to_v", (0, 1), dm.z[2], "zv_nov"), ("encoder_lv_to_a", (0, 2), dm.z[1], "za_noa"),
            ("encoder_av_to_l", (1, 2), dm.z[0], "zl_nol"), ("encoder_la_to_y", (0, 1), dm.zy, "zy_nov"),
            ("encoder_lv_to_y", (0, 2), dm.zy, "zy_noa"), ("encoder_av_to_y", (1, 2), dm.zy, "zy_nol")]


# latent fed to the factor MLP of (l, a, v, y) in pass p (mfm_model.py:876-882); "z0".."z2", "zy" are the true latents
PASS_LATENTS = (("z0", "z1", "z2", "zy"), ("zl_nol", "z1", "z2", "zy_nol"), ("z0", "za_noa", "z2", "zy_noa"),
                ("z0", "z1", "zv_nov", "zy_nov"))
# the latent-matching terms F.mse_loss(inferred, true) (:853-858)
MATCH = (("zv_nov", "z2"), ("za_noa", "z1"), ("zl_nol", "z0"), ("zy_nov", "zy"), ("zy_noa", "zy"), ("zy_nol", "zy"))


class MissingEngine(E.Engine):
    """One (T, B) instance of the MFM_missing schedule with its HBM workspace.
    loss_buf: 0 sum of the four L1 label terms, 1..3 the reconstruction MSEs per modality (two terms each, LOSS_RECON), 4..7 MMD
    parts, 8 total, 9 the latent-matching loss, 10 the all-present text reconstruction MSE (train_mfm_missing's epoch loss)."""

    def __init__(self, configs, T: int, B: int, device, ops, head: str = "l1"):
        if head != "l1":
            raise ValueError("MFM_missing is trained with the L1 label loss only (mfm_mosi.py:977-980)")
        super().__init__(configs, T, B, device, ops, head=head, variant="mfm")
        self.cross = cross_encoders(self.dm)

    def _lat(self, key):
        return self.ws[dict(z0="Z0", z1="Z1", z2="Z2", zy="ZY").get(key, "Zx_" + key)]

    # -- forward -------------------------------------------------------------------
    def forward(self, P: Dict[str, torch.Tensor], x: torch.Tensor, noise: Sequence[torch.Tensor],
                train: bool = False, rng: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """MFM_missing.forward (mfm_model.py:827-885).  Returns views into the workspace keyed like the reference's variables
        ("x_l_hat", "x_l_hat_nol", ..., "y_hat_nov" with x_*_hat as [T*B, d]), the ten latents, the MMD parts in loss_buf[4:8]
        and the latent-matching loss in loss_buf[9]."""
        dm, ops, buf = self.dm, self.ops, self.buf
        T, B, H = dm.T, dm.B, dm.H
        TB = T * B
        if tuple(x.shape) != (T, B, dm.D) or not x.is_contiguous() or x.dtype != torch.float32:
            raise ValueError("x must be contiguous fp32 [T=%d,B=%d,D=%d], got %s" % (T, B, dm.D, tuple(x.shape)))
        self.train = bool(train)
        self.x = self.x_in = x
        self.noise = list(noise)
        self.rng = rng
        X2 = x.view(TB, dm.D)
        pad4 = lambda n: (n + 3) // 4 * 4
        xs = [buf("Xp%d" % m, TB, pad4(dm.d[m]))[:, :dm.d[m]] for m in range(3)]
        self.xs = xs
        # cat(x_a, x_b) of the cross-modal encoders (:843-850), packed with 16 B-aligned rows
        xq = {ab: buf("Xq%d%d" % ab, TB, pad4(dm.d[ab[0]] + dm.d[ab[1]]))[:, :dm.d[ab[0]] + dm.d[ab[1]]]
              for ab in ((0, 1), (0, 2), (1, 2))}
        self.xq = xq
        drop = (lambda p, site: (p, site) if (train and p > 0.0) else None)
        pre = self.pre
        self.mark("fwd:start")

        # (0) aligned copies of x, (1) input projections of the three encoders and the three MFN cells
        def project(m):
            def run():
                tag = TAGS[m]
                src = X2[:, dm.off[m]:dm.off[m] + dm.d[m]]
                ops.copy2d(src, xs[m])
                for (a, b), dst in xq.items():
                    if m == a:
                        ops.copy2d(src, dst[:, :dm.d[a]])
                    elif m == b:
                        ops.copy2d(src, dst[:, dm.d[a]:])
                e, n = "encoder_%s.lstm" % tag, pre + "lstm_%s" % tag
                ops.gemm("nt", xs[m], P[e + ".weight_ih"], buf("GxE%d" % m, TB, 4 * dm.z[m]), bias=P[e + ".bias_ih"],
                         bias2=P[e + ".bias_hh"])
                ops.gemm("nt", xs[m], P[n + ".weight_ih"], buf("GxN%d" % m, TB, 4 * dm.hm[m]), bias=P[n + ".bias_ih"],
                         bias2=P[n + ".bias_hh"])
            return run

        self._par([project(0), project(1), project(2)])
        self._par([(lambda name=name, ab=ab, z=z, key=key: ops.gemm(
            "nt", xq[ab], P[name + ".lstm.weight_ih"], buf("GxX_" + key, TB, 4 * z), bias=P[name + ".lstm.bias_ih"],
            bias2=P[name + ".lstm.bias_hh"])) for name, ab, z, key in self.cross])
        self.mark("fwd:projections")

        # (2) recurrences: MFM's six cells in one launch, the six cross-modal cells in a second
        Hall = buf("Hall", (T + 1) * B, H)
        CS2 = buf("CS2", (T + 2) * B, 2 * H)
        Call, Cdup = CS2[:(T + 1) * B, H:], CS2[B:, :H]
        self.ws_views = dict(Call=Call)
        cells = []
        for m, tag in enumerate(TAGS):
            cells.append(dict(T=T, B=B, h=dm.z[m], gx=self.ws["GxE%d" % m], gx_steps=T, bias_rest=None,
                              W=P["encoder_%s.lstm.weight_hh" % tag], hs=buf("hsE%d" % m, (T + 1) * B, dm.z[m]),
                              cs=buf("csE%d" % m, (T + 1) * B, dm.z[m]), gates=buf("gatesE%d" % m, TB, 4 * dm.z[m])))
        for m, tag in enumerate(TAGS):
            o = dm.hoff[m]
            cells.append(dict(T=T, B=B, h=dm.hm[m], gx=self.ws["GxN%d" % m], gx_steps=T, bias_rest=None,
                              W=P[pre + "lstm_%s.weight_hh" % tag], hs=Hall[:, o:o + dm.hm[m]], cs=Call[:, o:o + dm.hm[m]],
                              cs_dup=Cdup[:, o:o + dm.hm[m]], gates=buf("gatesN%d" % m, TB, 4 * dm.hm[m])))
        ops.lstm_fwd(cells)
        ops.lstm_fwd([dict(T=T, B=B, h=z, gx=self.ws["GxX_" + key], gx_steps=T, bias_rest=None, W=P[name + ".lstm.weight_hh"],
                           hs=buf("hsX_" + key, (T + 1) * B, z), cs=buf("csX_" + key, (T + 1) * B, z),
                           gates=buf("gatesX_" + key, TB, 4 * z)) for name, ab, z, key in self.cross])
        self.mark("fwd:lstm enc+mfn")

        # (3) latents: z_m = fc1(h_T) with their MMD on the auxiliary streams; the inferred latents on the main stream
        Z = [buf("Z%d" % m, B, dm.z[m]) for m in range(3)]
        ops.zero(self.mmd_acc.view(torch.float32))
        self._z_ready = []
        for m, tag in enumerate(TAGS):
            with self._aux(m):
                ops.gemm("nt", self.ws["hsE%d" % m][TB:], P["encoder_%s.fc1.weight" % tag], Z[m], bias=P["encoder_%s.fc1.bias" % tag])
                self._z_ready.append(self._aux_event(m))
                if self.want_mmd:
                    self._mmd(m, Z[m])
        for name, ab, z, key in self.cross:
            ops.gemm("nt", self.ws["hsX_" + key][TB:], P[name + ".fc1.weight"], buf("Zx_" + key, B, z), bias=P[name + ".fc1.bias"])

        # (4)-(7) the MFN block, z_y and its MMD: the engine's own
        ZY = self._forward_mfn_head(P, CS2, Hall, drop, rng)
        for ev in self._z_ready:
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
        self._z_ready = None

        # (7b) the latent-matching loss (:853-858); its gradients are formed in backward (they scale with dLoss/dmissing)
        ops.zero(self.loss_buf[9:11])
        for inf, true in MATCH:
            a, b = self._lat(inf), self._lat(true)
            ops.mse_fwd_bwd(a, b, 1.0 / float(a.numel()), 0.0, self.loss_buf[9:10], None)

        # (8)-(11) the generative half, once per pass (:860-883)
        lat = [[self._lat(k) for k in PASS_LATENTS[p]] for p in range(4)]
        FY = [buf("FY@%d" % p, B, dm.fy) for p in range(4)]
        EMB = [[buf("EMB%d@%d" % (m, p), B, dm.hd[m]) for m in range(3)] for p in range(4)]
        Xhat = [[buf("Xhat%d@%d" % (m, p), TB, dm.d[m]) for m in range(3)] for p in range(4)]
        Yhat = [buf("Yhat@%d" % p, B, dm.out) for p in range(4)]

        def mlp_y():
            for p in range(4):
                F1 = buf("F1y@%d" % p, B, dm.fy)
                ops.gemm("nt", lat[p][3], P["zy_to_fy_fc1.weight"], F1, bias=P["zy_to_fy_fc1.bias"], act=ACT_RELU,
                         drop=drop(dm.p_fy, SITE_FY + SITE_PASS * p), rng=rng)
                ops.gemm("nt", F1, P["zy_to_fy_fc2.weight"], FY[p], bias=P["zy_to_fy_fc2.bias"], act=ACT_RELU)

        def mlp_m(m):
            def run():
                nm = "z%s_to_f%s" % (TAGS[m], TAGS[m])
                for p in range(4):
                    F1 = buf("F1_%d@%d" % (m, p), B, dm.f[m])
                    ops.gemm("nt", lat[p][m], P[nm + "_fc1.weight"], F1, bias=P[nm + "_fc1.bias"], act=ACT_RELU,
                             drop=drop(dm.p_f[m], SITE_FL + m + SITE_PASS * p), rng=rng)
                    ops.gemm("nt", F1, P[nm + "_fc2.weight"], EMB[p][m][:, dm.fy:], bias=P[nm + "_fc2.bias"], act=ACT_RELU)
                d_, hd = "decoder_%s.lstm" % TAGS[m], dm.hd[m]
                ops.add(P[d_ + ".weight_ih"], P[d_ + ".weight_hh"], buf("Wm%d" % m, 4 * hd, hd))
                ops.add(P[d_ + ".bias_ih"].view(1, -1), P[d_ + ".bias_hh"].view(1, -1), buf("bsumD%d" % m, 1, 4 * hd))
            return run

        self.mark("fwd:zy")
        self._par([mlp_y, mlp_m(0), mlp_m(1), mlp_m(2)])
        self.mark("fwd:factor MLPs")

        def decoder(m):
            def run():
                tag = TAGS[m]
                d_, hd = "decoder_%s.lstm" % tag, dm.hd[m]
                cells = []
                for p in range(4):
                    ops.copy2d(FY[p], EMB[p][m][:, :dm.fy])
                    G0 = buf("G0_%d@%d" % (m, p), B, 4 * hd)
                    ops.gemm("nt", EMB[p][m], P[d_ + ".weight_ih"], G0, bias=P[d_ + ".bias_ih"], bias2=P[d_ + ".bias_hh"])
                    cells.append(dict(T=T, B=B, h=hd, gx=G0, gx_steps=1, bias_rest=self.ws["bsumD%d" % m].view(-1),
                                      W=self.ws["Wm%d" % m], hs=buf("hsD%d@%d" % (m, p), (T + 1) * B, hd),
                                      cs=buf("csD%d@%d" % (m, p), (T + 1) * B, hd), gates=buf("gatesD%d@%d" % (m, p), TB, 4 * hd)))
                ops.lstm_fwd(cells)                              # the four passes of this decoder: same weights, four states
                for p in range(4):
                    ops.gemm("nt", self.ws["hsD%d@%d" % (m, p)][B:], P["decoder_%s.fc1.weight" % tag], Xhat[p][m],
                             bias=P["decoder_%s.fc1.bias" % tag])
            return run

        def head():
            for p in range(4):
                Y1 = buf("Y1@%d" % p, B, dm.fy)
                ops.gemm("nt", FY[p], P["fy_to_y_fc1.weight"], Y1, bias=P["fy_to_y_fc1.bias"], act=ACT_RELU,
                         drop=drop(dm.p_y, SITE_Y + SITE_PASS * p), rng=rng)
                ops.gemm("nt", Y1, P["fy_to_y_fc2.weight"], Yhat[p], bias=P["fy_to_y_fc2.bias"])

        self._par([decoder(0), decoder(1), decoder(2), head])
        self.mark("fwd:decoders+head")
        if not self.defer_mmd_join:
            self._join_aux()
        out = dict(zl=Z[0], za=Z[1], zv=Z[2], zy=ZY, mmd_parts=self.loss_buf[4:8], missing=self.loss_buf[9:10])
        for name, ab, z, key in self.cross:
            out[key] = self.ws["Zx_" + key]
        for p, sfx in enumerate(PASSES):
            for m, tag in enumerate(TAGS):
                out["x_%s_hat%s" % (tag, sfx)] = Xhat[p][m]
            out["y_hat" + sfx] = Yhat[p]
        return out

    # -- losses (train_mfm_missing's step) ---------------------------------------------
    def losses(self, y: torch.Tensor):
        """mfm_mosi.py:962-982.  Returns (dX, dY): dX[(p, m)] for the six reconstructions in the loss, dY[p] for the four
        label heads."""
        dm, ops, buf = self.dm, self.ops, self.buf
        TB = dm.T * dm.B
        if not self.defer_mmd_join:
            self._join_aux()
        ops.zero(self.loss_buf[0:4])
        dX, dY = {}, []

        def recon(m):
            def run():
                n = float(TB * dm.d[m])
                for p, mm in LOSS_RECON:
                    if mm != m:
                        continue
                    dX[(p, m)] = buf("dXhat%d@%d" % (m, p), TB, dm.d[m])
                    ops.mse_fwd_bwd(self.ws["Xhat%d@%d" % (m, p)], self.xs[m], 1.0 / n, 2.0 * dm.lda[m] / n,
                                    self.loss_buf[1 + m:2 + m], dX[(p, m)])
                    if m == 0 and p == 0:                     # the epoch loss of train_mfm_missing is this term alone (:984)
                        ops.copy2d(self.loss_buf[1:2].view(1, 1), self.loss_buf[10:11].view(1, 1))
            return run

        def disc():
            for p in range(4):
                d = buf("dYhat@%d" % p, dm.B, dm.out)
                ops.l1_fwd_bwd(self.ws["Yhat@%d" % p], y.view(dm.B, dm.out), 1.0 / (dm.B * dm.out), self.loss_buf[0:1], d)
                dY.append(d)

        self._par([recon(0), recon(1), recon(2), disc])
        if not self.defer_mmd_join:
            self._total()
        return dX, dY

    def _total(self):
        dm, ops = self.dm, self.ops
        ops.loss_total(self.loss_buf, dm.lda[0], dm.lda[1], dm.lda[2], dm.lda_mmd)
        ops.copy2d(self.loss_buf[9:10].view(1, 1), self.loss_buf[8:9].view(1, 1), accumulate=True)     # + missing_loss (:981)

    # -- backward ------------------------------------------------------------------
    def backward(self, P: Dict[str, torch.Tensor], G: Dict[str, torch.Tensor], dXhat, dYhat, mmd_scale: float,
                 mmd_scale_dev: Optional[torch.Tensor] = None, missing_scale: float = 1.0):
        """Adjoint of ``forward``.  ``dXhat[(p, m)]`` [T*B, d_m] for the reconstructions that carry a gradient (absent keys: the
        decoder pass is skipped), ``dYhat[p]`` [B, out]; d(loss)/d(mmd) = mmd_scale x mmd_scale_dev; d(loss)/d(missing) =
        missing_scale.  Parameter gradients are ACCUMULATED into ``G``."""
        dm, ops, buf, ws = self.dm, self.ops, self.buf, self.ws
        T, B, H, mem = dm.T, dm.B, dm.H, dm.mem
        TB = T * B
        relu_scale = (lambda p: 1.0 / (1.0 - p) if (self.train and p > 0.0) else 1.0)

        def wgrad(dY, A, name):
            self._wgrad_gemm(dY, A, G[name], accumulate=True)

        def bgrad(dY, name):
            ops.colsum(dY, G[name])

        def lin_bwd(dY, A, name, dA=None, accumulate=False, mask=None, mask_scale=1.0):
            self._wgrad_gemm(dY, A, G[name + ".weight"], accumulate=True, colsum_out=G[name + ".bias"])
            if dA is not None:
                ops.gemm("nn", dY, P[name + ".weight"], dA, accumulate=accumulate, mask=mask, mask_scale=mask_scale)

        # (7') MMD gradients on the auxiliary streams; the attention's concatenated weights ride along
        lat4 = [ws["Z0"], ws["Z1"], ws["Z2"], ws["ZY"]]
        dmmd = [buf("dZmmd%d" % k, B, lat4[k].shape[1]) for k in range(4)]
        with self._aux(0):
            self._build_wcat(P)
        self._wcat_ready = True
        for k in range(4):
            with self._aux(k):
                ops.zero(dmmd[k])
                rc, t12 = ws["mmd_rc%d" % k], ws["mmd_t12_%d" % k]
                ops.mmd_combine(lat4[k], rc[:B], rc[B:], t12[:B], t12[B:], mmd_scale, dmmd[k], mmd_scale_dev)

        # gradient accumulators of the ten latents
        keys = ["z0", "z1", "z2", "zy"] + [c[3] for c in self.cross]
        dlat = {k: buf("dLat_" + k, B, self._lat(k).shape[1]) for k in keys}
        for k in keys:
            ops.zero(dlat[k])
        dFY = [buf("dFY@%d" % p, B, dm.fy) for p in range(4)]
        dEMB = {}

        def mlp2_bwd(tagp, df, f, F1, zin, nm, p, dz_acc):
            dpre = buf("dpre_" + tagp, f.shape[0], f.shape[1])
            ops.relu_bwd(df, f, dpre)
            dF1 = buf("dF1_" + tagp, F1.shape[0], F1.shape[1])
            lin_bwd(dpre, F1, nm + "_fc2", dF1, mask=F1, mask_scale=relu_scale(p))
            dz = buf("dz_" + tagp, zin.shape[0], zin.shape[1])
            lin_bwd(dF1, zin, nm + "_fc1", dz)
            return dz

        def head_bwd():
            for p in range(4):
                dY1 = buf("dY1@%d" % p, B, dm.fy)
                lin_bwd(dYhat[p], ws["Y1@%d" % p], "fy_to_y_fc2", dY1, mask=ws["Y1@%d" % p], mask_scale=relu_scale(dm.p_y))
                lin_bwd(dY1, ws["FY@%d" % p], "fy_to_y_fc1", dFY[p])

        dz_m = {}

        def decoder_bwd(m):
            def run():
                tag, hd = TAGS[m], dm.hd[m]
                d_ = "decoder_%s.lstm" % tag
                used = [p for p in range(4) if (p, m) in dXhat]
                if not used:
                    return
                cells = []
                for p in used:
                    dHd = buf("dHd%d@%d" % (m, p), TB, hd)
                    lin_bwd(dXhat[(p, m)], ws["hsD%d@%d" % (m, p)][B:], "decoder_%s.fc1" % tag, dHd)
                    cells.append(dict(T=T, B=B, h=hd, gates=ws["gatesD%d@%d" % (m, p)], cs=ws["csD%d@%d" % (m, p)],
                                      W=ws["Wm%d" % m], dh_all=dHd, dh_last=None, dc_ext=None,
                                      dG=buf("dGD%d@%d" % (m, p), TB, 4 * hd), dc_scratch=buf("dcSD%d@%d" % (m, p), B, hd)))
                ops.lstm_bwd(cells)
                # as in Engine.backward: for t >= 1 the input IS h_{t-1}, so dW_ih and dW_hh share dG^T h_prev (and the bias
                # gradients the same column sums).  Here the shared part is summed over the passes first, added to the input
                # gradients once, then each pass's step-0 product follows; all on ONE side stream, in order.
                key = G[d_ + ".weight_ih"].data_ptr()
                for p, c in zip(used, cells):
                    self._wgrad_gemm(c["dG"], ws["hsD%d@%d" % (m, p)][:TB], G[d_ + ".weight_hh"], accumulate=True,
                                     colsum_out=G[d_ + ".bias_hh"], stream_key=key)
                self._on_side(key, lambda: (ops.copy2d(G[d_ + ".weight_hh"], G[d_ + ".weight_ih"], accumulate=True),
                                            ops.copy2d(G[d_ + ".bias_hh"].view(1, -1), G[d_ + ".bias_ih"].view(1, -1), accumulate=True)))
                for p, c in zip(used, cells):
                    self._wgrad_gemm(c["dG"][:B], ws["EMB%d@%d" % (m, p)], G[d_ + ".weight_ih"], accumulate=True, stream_key=key)
                for p, c in zip(used, cells):
                    dEMB[(p, m)] = buf("dEMB%d@%d" % (m, p), B, hd)
                    ops.gemm("nn", c["dG"][:B], P[d_ + ".weight_ih"], dEMB[(p, m)])
                    dz_m[(p, m)] = mlp2_bwd("f%d@%d" % (m, p), dEMB[(p, m)][:, dm.fy:], ws["EMB%d@%d" % (m, p)][:, dm.fy:],
                                            ws["F1_%d@%d" % (m, p)], self._lat(PASS_LATENTS[p][m]), "z%s_to_f%s" % (tag, tag),
                                            dm.p_f[m], None)
            return run

        self.mark("bwd:start")
        self._par([decoder_bwd(0), decoder_bwd(1), decoder_bwd(2), head_bwd])
        self.mark("bwd:decoder chains")
        for (p, m), dz in sorted(dz_m.items()):
            ops.copy2d(dz, dlat[PASS_LATENTS[p][m]], accumulate=True)
        for p in range(4):
            for m in range(3):
                if (p, m) in dEMB:
                    ops.copy2d(dEMB[(p, m)][:, :dm.fy], dFY[p], accumulate=True)
            dz = mlp2_bwd("fy@%d" % p, dFY[p], ws["FY@%d" % p], ws["F1y@%d" % p], self._lat(PASS_LATENTS[p][3]), "zy_to_fy",
                          dm.p_fy, None)
            ops.copy2d(dz, dlat[PASS_LATENTS[p][3]], accumulate=True)
        self.mark("bwd:mlp y")

        # (7b') latent matching: d/d inferred = s (inferred - true), d/d true = s (true - inferred), s = 2 missing_scale / n
        if missing_scale != 0.0:
            for inf, true in MATCH:
                a, b = self._lat(inf), self._lat(true)
                s = 2.0 * float(missing_scale) / float(a.numel())
                t = buf("dMatch_" + inf, a.shape[0], a.shape[1])
                ops.mse_fwd_bwd(a, b, 0.0, s, self.loss_buf[9:10], t)
                ops.copy2d(t, dlat[inf], accumulate=True)
                ops.mse_fwd_bwd(b, a, 0.0, s, self.loss_buf[9:10], t)
                ops.copy2d(t, dlat[true], accumulate=True)

        self._join_aux()
        self.mark("bwd:join mmd")
        if self.defer_mmd_join:
            self._total()
        for k, key in enumerate(("z0", "z1", "z2", "zy")):
            ops.copy2d(dmmd[k], dlat[key], accumulate=True)

        # (3'') the six cross-modal encoders: heads, one recurrence launch, weight gradients
        xcells = []
        for name, ab, z, key in self.cross:
            dh = buf("dhX_" + key, B, z)
            lin_bwd(dlat[key], ws["hsX_" + key][TB:], name + ".fc1", dh)
            xcells.append(dict(T=T, B=B, h=z, gates=ws["gatesX_" + key], cs=ws["csX_" + key], W=P[name + ".lstm.weight_hh"],
                               dh_all=None, dh_last=dh, dc_ext=None, dG=buf("dGX_" + key, TB, 4 * z),
                               dc_scratch=buf("dcSX_" + key, B, z)))
        ops.lstm_bwd(xcells)
        for i, ((name, ab, z, key), c) in enumerate(zip(self.cross, xcells)):
            nm = name + ".lstm"
            self._wgrad_pair(c["dG"], self.xq[ab], G[nm + ".weight_ih"], G[nm + ".bias_ih"], ws["hsX_" + key][:TB],
                             G[nm + ".weight_hh"], G[nm + ".bias_hh"], index=i)

        # (6') last_to_zy_fc1 over cat(h_T, mem_T); (3') the encoder heads; then the engine's MFN adjoint with the encoder cells
        dZY = dlat["zy"]
        Wzy, Gzy = P["last_to_zy_fc1.weight"], G["last_to_zy_fc1.weight"]
        Hall, mems = ws["Hall"], ws["mems"]
        self._wgrad_gemm(dZY, Hall[TB:], Gzy[:, :H], accumulate=True)
        self._wgrad_gemm(dZY, mems[TB:], Gzy[:, H:], accumulate=True)
        bgrad(dZY, "last_to_zy_fc1.bias")
        dHlast, dmemT = buf("dHlast", B, H), buf("dmemT", B, mem)
        ops.gemm("nn", dZY, Wzy[:, :H], dHlast)
        ops.gemm("nn", dZY, Wzy[:, H:], dmemT)
        enc_cells = []
        for m, tag in enumerate(TAGS):
            dhE = buf("dhE%d" % m, B, dm.z[m])
            lin_bwd(dlat["z%d" % m], ws["hsE%d" % m][TB:], "encoder_%s.fc1" % tag, dhE)
            enc_cells.append(dict(T=T, B=B, h=dm.z[m], gates=ws["gatesE%d" % m], cs=ws["csE%d" % m],
                                  W=P["encoder_%s.lstm.weight_hh" % tag], dh_all=None, dh_last=dhE, dc_ext=None,
                                  dG=buf("dGE%d" % m, TB, 4 * dm.z[m]), dc_scratch=buf("dcSE%d" % m, B, dm.z[m])))
        self._backward_mfn(P, G, dHlast, dmemT, enc_cells, wgrad, bgrad, lin_bwd, relu_scale)
        self.mark("bwd:lstm enc+mfn")
        self._join_side()
        self.mark("bwd:join wgrads")


# ---------------------------------------------------------------------------------------------------------------------
# the drop-in module
# ---------------------------------------------------------------------------------------------------------------------

_OUT_KEYS = ["x_%s_hat%s" % (t, s) if t != "y" else "y_hat" + s for s in PASSES for t in ("l", "a", "v", "y")]


class _MissingFunction(torch.autograd.Function):
    """forward = MissingEngine.forward, backward = MissingEngine.backward; inputs (x, 4 noise tensors, *params); outputs the 16
    decoded tensors in the reference's order, the MMD and the latent-matching loss."""

    @staticmethod
    def forward(ctx, module, x, n0, n1, n2, n3, *params):
        T, B, _ = x.shape
        eng = module._engine(T, B, x.device)
        P = OrderedDict(zip(module._param_names, params))
        rng = module._rng_state(x.device)
        if module.training:
            _ops().rng_tick(rng)
        eng.want_mmd = not module.__dict__.get("_skip_mmd_now", False)
        if not eng.want_mmd:
            _ops().zero(eng.loss_buf[4:8])
        try:
            out = eng.forward(P, x, [n0, n1, n2, n3], train=module.training, rng=rng)
        finally:
            eng.want_mmd = True
        ctx.eng, ctx.P, ctx.gen = eng, P, eng_generation(eng, bump=True)
        dm = eng.dm
        res = []
        for k in _OUT_KEYS:
            t = out[k]
            res.append(t.clone() if k.startswith("y_hat") else t.reshape(T, B, -1).clone())
        module._latents = {k: out[k].clone() for k in ("zl", "za", "zv", "zy", "zl_nol", "za_noa", "zv_nov", "zy_nol", "zy_noa", "zy_nov")}
        return tuple(res) + (eng.loss_buf[4:8].sum(), eng.loss_buf[9].clone())

    @staticmethod
    def backward(ctx, *grads):
        eng, P = ctx.eng, ctx.P
        if eng_generation(eng) != ctx.gen:
            raise RuntimeError("MFM_missing backward called after another forward of the same (T,B) shape overwrote the "
                               "kernel workspace; run backward before the next forward")
        dm = eng.dm
        TB = dm.T * dm.B
        dX, dY = {}, []
        for p in range(4):
            for m in range(3):
                g = grads[4 * p + m]
                if g is not None:
                    dX[(p, m)] = g.contiguous().view(TB, dm.d[m])
            g = grads[4 * p + 3]
            dY.append(torch.zeros(dm.B, dm.out, dtype=torch.float32, device=eng.device) if g is None
                      else g.contiguous().view(dm.B, dm.out))
        dmmd, dmiss = grads[16], grads[17]
        sizes = [p.numel() for p in P.values()]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=eng.device)
        _ops().zero(flat)
        G, o = OrderedDict(), 0
        for (k, p), n in zip(P.items(), sizes):
            G[k] = flat[o:o + n].view(p.shape)
            o += n
        # the coefficient of the latent-matching loss is read on the host (one D2H sync; the fused trainer has none: it is 1)
        ms = 0.0 if dmiss is None else float(dmiss)
        if dmmd is None:
            eng.backward(P, G, dX, dY, 0.0, missing_scale=ms)
        else:
            eng.backward(P, G, dX, dY, 1.0, mmd_scale_dev=dmmd.contiguous().view(1), missing_scale=ms)
        return (None, None, None, None, None, None) + tuple(G.values())


class MFM_missing(MFM):
    """mfm_model.py:766-885.  ``forward(x[T,N,D]) -> (decoded, decoded_nol, decoded_noa, decoded_nov, mmd_loss, missing_loss)``,
    each ``decoded*`` = [x_l_hat, x_a_hat, x_v_hat, y_hat]."""
    _variant = "missing"

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        nn.Module.__init__(self)
        [self.d_l, self.d_a, self.d_v] = config["input_dims"]
        [self.dh_l, self.dh_a, self.dh_v] = config["h_dims"]
        d_l, d_a, d_v = self.d_l, self.d_a, self.d_v
        zy, zl, za, zv = config["zy_size"], config["zl_size"], config["za_size"], config["zv_size"]
        fy, fl, fa, fv = config["fy_size"], config["fl_size"], config["fa_size"], config["fv_size"]
        last_mfn_size = self.dh_l + self.dh_a + self.dh_v + config["memsize"]
        # construction order fixes the init RNG stream (mfm_model.py:788-825)
        self.encoder_l = encoderLSTM(d_l, zl)
        self.encoder_a = encoderLSTM(d_a, za)
        self.encoder_v = encoderLSTM(d_v, zv)
        self.encoder_la_to_v = encoderLSTM(d_l + d_a, zv)
        self.encoder_lv_to_a = encoderLSTM(d_l + d_v, za)
        self.encoder_av_to_l = encoderLSTM(d_a + d_v, zl)
        self.encoder_la_to_y = encoderLSTM(d_l + d_a, zy)
        self.encoder_lv_to_y = encoderLSTM(d_l + d_v, zy)
        self.encoder_av_to_y = encoderLSTM(d_a + d_v, zy)
        self.decoder_l = decoderLSTM(fy + fl, d_l)
        self.decoder_a = decoderLSTM(fy + fa, d_a)
        self.decoder_v = decoderLSTM(fy + fv, d_v)
        self.mfn_encoder = MFN(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)
        self.last_to_zy_fc1 = nn.Linear(last_mfn_size, zy)
        self.zy_to_fy_fc1 = nn.Linear(zy, fy)
        self.zy_to_fy_fc2 = nn.Linear(fy, fy)
        self.zy_to_fy_dropout = nn.Dropout(config["zy_to_fy_dropout"])
        self.zl_to_fl_fc1 = nn.Linear(zl, fl)
        self.zl_to_fl_fc2 = nn.Linear(fl, fl)
        self.zl_to_fl_dropout = nn.Dropout(config["zl_to_fl_dropout"])
        self.za_to_fa_fc1 = nn.Linear(za, fa)
        self.za_to_fa_fc2 = nn.Linear(fa, fa)
        self.za_to_fa_dropout = nn.Dropout(config["za_to_fa_dropout"])
        self.zv_to_fv_fc1 = nn.Linear(zv, fv)
        self.zv_to_fv_fc2 = nn.Linear(fv, fv)
        self.zv_to_fv_dropout = nn.Dropout(config["zv_to_fv_dropout"])
        self.fy_to_y_fc1 = nn.Linear(fy, fy)
        self.fy_to_y_fc2 = nn.Linear(fy, config["output_dim"])
        self.fy_to_y_dropout = nn.Dropout(config["fy_to_y_dropout"])
        self._cfg = [dict(config), dict(NN1Config), dict(NN2Config), dict(gamma1Config), dict(gamma2Config), dict(outConfig)]
        self._param_names = [k for k, _ in self.named_parameters() if k not in UNUSED]
        self.mmd_noise = "cpu"
        self.dropout_seed = 123

    def _engine(self, T, B, device):
        engs = self.__dict__.setdefault("_engines", {})
        key = (int(T), int(B), str(device))
        if key not in engs:
            if len(engs) >= 4:
                engs.pop(next(iter(engs)))
            engs[key] = MissingEngine(self._cfg, T, B, device, _ops(), head="l1")
        return engs[key]

    def forward(self, x):
        _require_cuda(x, "MFM_missing.forward")
        if x.dim() != 3:
            raise ValueError("MFM_missing.forward expects x[T,N,D]")
        if x.requires_grad:
            raise RuntimeError("MFM_missing.forward: gradient w.r.t. the input is not provided (the reference never asks for it)")
        x = x.contiguous().float()
        noise = self.draw_mmd_noise(x.shape[1], x.device)
        pd = dict(self.named_parameters())
        params = [pd[k] for k in self._param_names]
        self.__dict__["_skip_mmd_now"] = bool(getattr(self, "eval_skip_mmd", False) and not self.training and not torch.is_grad_enabled())
        res = _MissingFunction.apply(self, x, *noise, *params)
        dec = [list(res[4 * p:4 * p + 4]) for p in range(4)]
        return dec[0], dec[1], dec[2], dec[3], res[16], res[17]
