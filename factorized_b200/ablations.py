"""The ablation models of the reference's ``train_mfm_ablation`` (mfm_mosi.py:640-767): M_A, M_B, M_C, M_D
(/root/reference/mfm_model.py:201-467) -- MFM with parts removed -- on the same kernels.

=====  ===================================  =========  ==========================  ==========================  ==============
model  encoders                             MFN / z_y  regulariser (MMD of)        decoders read               label head
=====  ===================================  =========  ==========================  ==========================  ==============
M_A    ONE, over the whole input (z_l)      yes        z_l, z_y        (:255)      cat(fy, fl), all three      fy_to_y on fy
M_B    three, one per modality              no         z_l, z_a, z_v   (:328)      f_l / f_a / f_v             fy_to_y on cat(fl, fa, fv)
M_C    none                                 yes        z_y             (:392)      fy, all three               fy_to_y on fy
M_D    three, one per modality              no         none (0.0)      (:456)      none: "decoded" = inputs    fs_to_y (one Linear)
=====  ===================================  =========  ==========================  ==========================  ==============

``AblationEngine`` is the host schedule (forward + hand-derived backward) of these four topologies.  It issues the same
primitives as ``engine.Engine`` and reuses its MFN block (``_forward_mfn_head`` / ``_backward_mfn``), its MMD streams, its
weight-gradient side streams and its loss heads unchanged; only the wiring between the blocks differs.  Like the engine it does
no arithmetic itself and has no CPU path.  The module classes keep the reference's constructor signatures, submodule names and
construction order (same initial weights for the same seed, same state-dict keys).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from . import engine as E
from .engine import ACT_RELU, SITE_FL, SITE_FY, SITE_Y, TAGS
from .mfm_model import MFM, MFN, UNUSED, _ops, decoderLSTM, encoderLSTM

VARIANTS = ("m_a", "m_b", "m_c", "m_d")


class AblationEngine(E.Engine):
    """One (T, B) instance of the schedule of M_A / M_B / M_C / M_D with its HBM workspace."""

    def __init__(self, configs, T: int, B: int, device, ops, head: str = "l1", variant: str = "m_a"):
        if variant not in VARIANTS:
            raise ValueError(variant)
        super().__init__(configs, T, B, device, ops, head=head, variant="mfm")
        dm = self.dm
        self.abl = variant
        self.has_mfn = variant in ("m_a", "m_c")
        self.has_dec = variant != "m_d"
        # encoders: (latent slot k, parameter prefix, modality whose columns of x it reads or None for the whole row, size)
        if variant == "m_a":
            self.enc = [(0, "encoder_l", None, dm.z[0])]
        elif variant == "m_c":
            self.enc = []
        else:
            self.enc = [(m, "encoder_%s" % TAGS[m], m, dm.z[m]) for m in range(3)]
        # width of the decoders' recurrent state = width of their step-0 input (mfm_model.py:225-227, 297-299, 367-369)
        self.hd = dict(m_a=[dm.fy + dm.f[0]] * 3, m_b=list(dm.f), m_c=[dm.fy] * 3, m_d=[0, 0, 0])[variant]
        self.mmd_slots = dict(m_a=(0, 3), m_b=(0, 1, 2), m_c=(3,), m_d=())[variant]   # latents with an MMD term

    # -- forward -------------------------------------------------------------------
    def forward(self, P: Dict[str, torch.Tensor], x: torch.Tensor, noise: Sequence[torch.Tensor],
                train: bool = False, rng: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """M_A.forward (mfm_model.py:244-269), M_B (:317-343), M_C (:382-403), M_D (:445-467).  Same contract as
        ``Engine.forward``; ``noise[k]`` is read only for the latents in ``mmd_slots``; latents a model lacks come back None."""
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
        xs = [buf("Xp%d" % m, TB, (dm.d[m] + 3) // 4 * 4)[:, :dm.d[m]] for m in range(3)]
        self.xs = xs
        drop = (lambda p, site: (p, site) if (train and p > 0.0) else None)
        pre = self.pre
        self.mark("fwd:start")
        if self.fuse_mse:
            ops.zero(self.loss_buf[0:4])

        # (0) aligned per-modality copies of x (MSE targets, and the inputs of the per-modality cells); (1) input projections
        def project(m):
            def run():
                ops.copy2d(X2[:, dm.off[m]:dm.off[m] + dm.d[m]], xs[m])
                for k, name, mod, z in self.enc:
                    if mod == m:
                        ops.gemm("nt", xs[m], P[name + ".lstm.weight_ih"], buf("GxE%d" % k, TB, 4 * z),
                                 bias=P[name + ".lstm.bias_ih"], bias2=P[name + ".lstm.bias_hh"])
                if self.has_mfn:
                    n = pre + "lstm_%s" % TAGS[m]
                    ops.gemm("nt", xs[m], P[n + ".weight_ih"], buf("GxN%d" % m, TB, 4 * dm.hm[m]),
                             bias=P[n + ".bias_ih"], bias2=P[n + ".bias_hh"])
            return run

        def project_whole(k, name, z):                     # M_A's encoder reads the whole row of x (:252)
            return lambda: ops.gemm("nt", X2, P[name + ".lstm.weight_ih"], buf("GxE%d" % k, TB, 4 * z),
                                    bias=P[name + ".lstm.bias_ih"], bias2=P[name + ".lstm.bias_hh"])

        self._par([project(0), project(1), project(2)] + [project_whole(k, name, z) for k, name, mod, z in self.enc if mod is None])
        self.mark("fwd:projections")

        # (2) all recurrences of the input side in one launch
        cells = []
        for k, name, mod, z in self.enc:
            cells.append(dict(T=T, B=B, h=z, gx=self.ws["GxE%d" % k], gx_steps=T, bias_rest=None, W=P[name + ".lstm.weight_hh"],
                              hs=buf("hsE%d" % k, (T + 1) * B, z), cs=buf("csE%d" % k, (T + 1) * B, z),
                              gates=buf("gatesE%d" % k, TB, 4 * z)))
        Hall = CS2 = None
        if self.has_mfn:                                       # cell histories in the attention's cat(c_{t-1}, c_t) layout
            Hall = buf("Hall", (T + 1) * B, H)
            CS2 = buf("CS2", (T + 2) * B, 2 * H)
            Call, Cdup = CS2[:(T + 1) * B, H:], CS2[B:, :H]
            self.ws_views = dict(Call=Call)
            for m, tag in enumerate(TAGS):
                o = dm.hoff[m]
                cells.append(dict(T=T, B=B, h=dm.hm[m], gx=self.ws["GxN%d" % m], gx_steps=T, bias_rest=None,
                                  W=P[pre + "lstm_%s.weight_hh" % tag], hs=Hall[:, o:o + dm.hm[m]],
                                  cs=Call[:, o:o + dm.hm[m]], cs_dup=Cdup[:, o:o + dm.hm[m]],
                                  gates=buf("gatesN%d" % m, TB, 4 * dm.hm[m])))
        ops.lstm_fwd(cells)
        self.mark("fwd:lstm enc+mfn")

        # (3) z = fc1(h_T) and its MMD, off the main stream
        Z = [None, None, None]
        ops.zero(self.mmd_acc.view(torch.float32))
        self._z_ready = []
        for k, name, mod, z in self.enc:
            Z[k] = buf("Z%d" % k, B, z)
            with self._aux(k):
                ops.gemm("nt", self.ws["hsE%d" % k][TB:], P[name + ".fc1.weight"], Z[k], bias=P[name + ".fc1.bias"])
                self._z_ready.append(self._aux_event(k))
                if self.want_mmd and k in self.mmd_slots:
                    self._mmd(k, Z[k])

        # (4)-(7) the MFN block, z_y and its MMD: the engine's own
        ZY = self._forward_mfn_head(P, CS2, Hall, drop, rng) if self.has_mfn else None
        for ev in self._z_ready:
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
        self._z_ready = None

        # (8) factor MLPs relu(fc2(drop(relu(fc1(z)))))
        fsum = sum(dm.f)
        FY = buf("FY", B, dm.fy) if self.has_mfn else None
        FS = buf("FS", B, fsum) if not self.has_mfn else None          # cat(fl, fa, fv) (:339, :463)
        if self.abl == "m_a":
            EMB = [buf("EMBA", B, self.hd[0])] * 3                        # cat(fy, fl), shared by the three decoders (:260)
            fdst = {0: EMB[0][:, dm.fy:]}
        elif self.abl == "m_b":
            EMB = [buf("EMB%d" % m, B, dm.f[m]) for m in range(3)]
            fdst = {m: EMB[m] for m in range(3)}
        elif self.abl == "m_c":
            EMB, fdst = [FY] * 3, {}
        else:
            EMB = None
            fdst = {m: buf("Fm%d" % m, B, dm.f[m]) for m in range(3)}
        self.fdst = fdst

        def mlp_y():
            ops.gemm("nt", ZY, P["zy_to_fy_fc1.weight"], buf("F1y", B, dm.fy), bias=P["zy_to_fy_fc1.bias"], act=ACT_RELU,
                     drop=drop(dm.p_fy, SITE_FY), rng=rng)
            ops.gemm("nt", self.ws["F1y"], P["zy_to_fy_fc2.weight"], FY, bias=P["zy_to_fy_fc2.bias"], act=ACT_RELU)

        def mlp_m(k):
            def run():
                nm = "z%s_to_f%s" % (TAGS[k], TAGS[k])
                F1 = buf("F1_%d" % k, B, dm.f[k])
                ops.gemm("nt", Z[k], P[nm + "_fc1.weight"], F1, bias=P[nm + "_fc1.bias"], act=ACT_RELU,
                         drop=drop(dm.p_f[k], SITE_FL + k), rng=rng)
                ops.gemm("nt", F1, P[nm + "_fc2.weight"], fdst[k], bias=P[nm + "_fc2.bias"], act=ACT_RELU)
                if FS is not None:
                    o = sum(dm.f[:k])
                    ops.copy2d(fdst[k], FS[:, o:o + dm.f[k]])
            return run

        def merged_weights(m):                              # the decoders' W_ih + W_hh (parameters only)
            def run():
                d_, hd = "decoder_%s.lstm" % TAGS[m], self.hd[m]
                ops.add(P[d_ + ".weight_ih"], P[d_ + ".weight_hh"], buf("Wm%d" % m, 4 * hd, hd))
                ops.add(P[d_ + ".bias_ih"].view(1, -1), P[d_ + ".bias_hh"].view(1, -1), buf("bsumD%d" % m, 1, 4 * hd))
            return run

        self.mark("fwd:zy")
        self._par(([mlp_y] if self.has_mfn else []) + [mlp_m(k) for k in sorted(fdst)]
                  + ([merged_weights(m) for m in range(3)] if self.has_dec else []))
        if self.abl == "m_a":
            ops.copy2d(FY, EMB[0][:, :dm.fy])
        self.mark("fwd:factor MLPs")

        # (9) decoders, (10) reconstructions, (11) label head
        Xhat = [None] * 3
        if not self.has_dec:
            Xhat = list(xs)                                  # M_D: "decoded" are the inputs themselves (:465)
            for m in range(3):
                self.ws["Xhat%d" % m] = xs[m]
        elif not self.fuse_mse:
            Xhat = [buf("Xhat%d" % m, TB, dm.d[m]) for m in range(3)]
        dXf = [buf("dXhat%d" % m, TB, dm.d[m]) for m in range(3)] if (self.fuse_mse and self.has_dec) else None
        Yhat = buf("Yhat", B, dm.out)

        def decoder(m):
            def run():
                tag = TAGS[m]
                d_, hd = "decoder_%s.lstm" % tag, self.hd[m]
                G0 = buf("G0_%d" % m, B, 4 * hd)
                ops.gemm("nt", EMB[m], P[d_ + ".weight_ih"], G0, bias=P[d_ + ".bias_ih"], bias2=P[d_ + ".bias_hh"])
                cell = dict(T=T, B=B, h=hd, gx=G0, gx_steps=1, bias_rest=self.ws["bsumD%d" % m].view(-1), W=self.ws["Wm%d" % m],
                            hs=buf("hsD%d" % m, (T + 1) * B, hd), cs=buf("csD%d" % m, (T + 1) * B, hd),
                            gates=buf("gatesD%d" % m, TB, 4 * hd))
                ops.lstm_fwd([cell])
                if self.fuse_mse:
                    n = float(TB * dm.d[m])
                    ops.gemm_mse(self.ws["hsD%d" % m][B:], P["decoder_%s.fc1.weight" % tag], P["decoder_%s.fc1.bias" % tag],
                                 xs[m], 1.0 / n, 2.0 * dm.lda[m] / n, self.loss_buf[1 + m:2 + m], dXf[m])
                else:
                    ops.gemm("nt", self.ws["hsD%d" % m][B:], P["decoder_%s.fc1.weight" % tag], Xhat[m],
                             bias=P["decoder_%s.fc1.bias" % tag])
            return run

        def head():
            if self.abl == "m_d":                             # y_hat = fs_to_y(cat(fl, fa, fv)) (:463-464)
                ops.gemm("nt", FS, P["fs_to_y.weight"], Yhat, bias=P["fs_to_y.bias"])
                return
            src = FY if self.has_mfn else FS                  # M_B: fy_to_y_fc1 reads cat(fl, fa, fv) (:339-340)
            Y1 = buf("Y1", B, dm.fy)
            ops.gemm("nt", src, P["fy_to_y_fc1.weight"], Y1, bias=P["fy_to_y_fc1.bias"], act=ACT_RELU,
                     drop=drop(dm.p_y, SITE_Y), rng=rng)
            ops.gemm("nt", Y1, P["fy_to_y_fc2.weight"], Yhat, bias=P["fy_to_y_fc2.bias"])

        self._par(([decoder(0), decoder(1), decoder(2)] if self.has_dec else []) + [head])
        self.mark("fwd:decoders+head")
        if not self.defer_mmd_join:
            self._join_aux()
        return dict(x_l_hat=Xhat[0], x_a_hat=Xhat[1], x_v_hat=Xhat[2], y_hat=Yhat, zl=Z[0], za=Z[1], zv=Z[2], zy=ZY,
                    fy=FY, mmd_parts=self.loss_buf[4:8])

    # -- backward ------------------------------------------------------------------
    def backward(self, P: Dict[str, torch.Tensor], G: Dict[str, torch.Tensor], dXhat: Sequence[torch.Tensor],
                 dYhat: torch.Tensor, mmd_scale: float, mmd_scale_dev: Optional[torch.Tensor] = None,
                 d_mfn_last: Optional[torch.Tensor] = None):
        """Adjoint of ``forward`` (same contract as ``Engine.backward``)."""
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

        # (7') the MMD terms' gradients, one combine per regularised latent, on the auxiliary streams
        lat = {k: ws["Z%d" % k] for k, _, _, _ in self.enc}
        if self.has_mfn:
            lat[3] = ws["ZY"]
            with self._aux(0):
                self._build_wcat(P)
            self._wcat_ready = True
        dmmd = {}
        for k in self.mmd_slots:
            dmmd[k] = buf("dZmmd%d" % k, B, lat[k].shape[1])
            with self._aux(k):
                ops.zero(dmmd[k])
                rc, t12 = ws["mmd_rc%d" % k], ws["mmd_t12_%d" % k]
                ops.mmd_combine(lat[k], rc[:B], rc[B:], t12[:B], t12[B:], mmd_scale, dmmd[k], mmd_scale_dev)

        fsum = sum(dm.f)
        dFY = buf("dFY", B, dm.fy) if self.has_mfn else None
        dFS = buf("dFS", B, fsum) if not self.has_mfn else None
        dZ = {k: buf("dZ%d" % k, B, z) for k, _, _, z in self.enc}
        dEMB = [buf("dEMB%d" % m, B, self.hd[m]) for m in range(3)] if self.has_dec else None

        def head_bwd():
            if self.abl == "m_d":
                lin_bwd(dYhat, ws["FS"], "fs_to_y", dFS)
                return
            dY1 = buf("dY1", B, dm.fy)
            lin_bwd(dYhat, ws["Y1"], "fy_to_y_fc2", dY1, mask=ws["Y1"], mask_scale=relu_scale(dm.p_y))
            if self.has_mfn:
                lin_bwd(dY1, ws["FY"], "fy_to_y_fc1", dFY)
            else:
                lin_bwd(dY1, ws["FS"], "fy_to_y_fc1", dFS)

        def decoder_bwd(m):
            def run():
                tag, hd = TAGS[m], self.hd[m]
                d_ = "decoder_%s.lstm" % tag
                emb = ws["EMBA"] if self.abl == "m_a" else (ws["FY"] if self.abl == "m_c" else ws["EMB%d" % m])
                dHd = buf("dHd%d" % m, TB, hd)
                lin_bwd(dXhat[m], ws["hsD%d" % m][B:], "decoder_%s.fc1" % tag, dHd)
                dG = buf("dGD%d" % m, TB, 4 * hd)
                ops.lstm_bwd([dict(T=T, B=B, h=hd, gates=ws["gatesD%d" % m], cs=ws["csD%d" % m], W=ws["Wm%d" % m],
                                   dh_all=dHd, dh_last=None, dc_ext=None, dG=dG, dc_scratch=buf("dcSD%d" % m, B, hd))])
                # as in Engine.backward: for t >= 1 the input IS h_{t-1}, so dW_ih and dW_hh share dG^T h_prev
                key = G[d_ + ".weight_ih"].data_ptr()
                self._wgrad_gemm(dG, ws["hsD%d" % m][:TB], G[d_ + ".weight_hh"], accumulate=True, colsum_out=G[d_ + ".bias_hh"],
                                 stream_key=key)
                self._on_side(key, lambda: (ops.copy2d(G[d_ + ".weight_hh"], G[d_ + ".weight_ih"], accumulate=True),
                                            ops.copy2d(G[d_ + ".bias_hh"].view(1, -1), G[d_ + ".bias_ih"].view(1, -1), accumulate=True)))
                self._wgrad_gemm(dG[:B], emb, G[d_ + ".weight_ih"], accumulate=True, stream_key=key)
                ops.gemm("nn", dG[:B], P[d_ + ".weight_ih"], dEMB[m])
            return run

        def mlp2_bwd(df, f, F1, zin, nm, p, dz):
            dpre = buf("dpre_" + nm, f.shape[0], f.shape[1])
            ops.relu_bwd(df, f, dpre)
            dF1 = buf("dF1_" + nm, F1.shape[0], F1.shape[1])
            lin_bwd(dpre, F1, nm + "_fc2", dF1, mask=F1, mask_scale=relu_scale(p))
            lin_bwd(dF1, zin, nm + "_fc1", dz)

        self.mark("bwd:start")
        self._par(([decoder_bwd(0), decoder_bwd(1), decoder_bwd(2)] if self.has_dec else []) + [head_bwd])
        self.mark("bwd:decoder chains")
        # gradients of the factors: the label head's share plus every decoder that read them
        df = {}
        if self.abl == "m_a":
            ops.copy2d(dEMB[1], dEMB[0], accumulate=True)
            ops.copy2d(dEMB[2], dEMB[0], accumulate=True)
            ops.copy2d(dEMB[0][:, :dm.fy], dFY, accumulate=True)
            df[0] = dEMB[0][:, dm.fy:]
        elif self.abl == "m_c":
            for m in range(3):
                ops.copy2d(dEMB[m], dFY, accumulate=True)
        else:
            for m in range(3):
                o = sum(dm.f[:m])
                if self.abl == "m_b":
                    ops.copy2d(dFS[:, o:o + dm.f[m]], dEMB[m], accumulate=True)
                    df[m] = dEMB[m]
                else:
                    df[m] = dFS[:, o:o + dm.f[m]]
        dZY = None
        if self.has_mfn:
            dZY = buf("dZY", B, dm.zy)
            mlp2_bwd(dFY, ws["FY"], ws["F1y"], ws["ZY"], "zy_to_fy", dm.p_fy, dZY)
        self._par([(lambda k=k: mlp2_bwd(df[k], self.fdst[k], ws["F1_%d" % k], ws["Z%d" % k], "z%s_to_f%s" % (TAGS[k], TAGS[k]),
                                         dm.p_f[k], dZ[k])) for k in sorted(df)])
        self.mark("bwd:mlp y")
        self._join_aux()
        self.mark("bwd:join mmd")
        if self.defer_mmd_join:
            ops.loss_total(self.loss_buf, dm.lda[0], dm.lda[1], dm.lda[2], dm.lda_mmd)
        for k in self.mmd_slots:
            ops.copy2d(dmmd[k], dZY if k == 3 else dZ[k], accumulate=True)

        # (3') encoder heads and cells: they hang off the latents only -- one launch, then their weight gradients
        enc_cells = []
        for k, name, mod, z in self.enc:
            dhE = buf("dhE%d" % k, B, z)
            lin_bwd(dZ[k], ws["hsE%d" % k][TB:], name + ".fc1", dhE)
            enc_cells.append(dict(T=T, B=B, h=z, gates=ws["gatesE%d" % k], cs=ws["csE%d" % k], W=P[name + ".lstm.weight_hh"],
                                  dh_all=None, dh_last=dhE, dc_ext=None, dG=buf("dGE%d" % k, TB, 4 * z),
                                  dc_scratch=buf("dcSE%d" % k, B, z)))
        if enc_cells:
            ops.lstm_bwd(enc_cells)
            X2 = self.x_in.view(TB, dm.D)
            for i, ((k, name, mod, z), c) in enumerate(zip(self.enc, enc_cells)):
                nm = name + ".lstm"
                self._wgrad_pair(c["dG"], X2 if mod is None else self.xs[mod], G[nm + ".weight_ih"], G[nm + ".bias_ih"],
                                 ws["hsE%d" % k][:TB], G[nm + ".weight_hh"], G[nm + ".bias_hh"], index=3 + i)

        if self.has_mfn:
            # (6') last_to_zy_fc1 over cat(h_T, mem_T), then the engine's MFN adjoint (memory recurrence, attention, cells)
            Wzy, Gzy = P["last_to_zy_fc1.weight"], G["last_to_zy_fc1.weight"]
            Hall, mems = ws["Hall"], ws["mems"]
            self._wgrad_gemm(dZY, Hall[TB:], Gzy[:, :H], accumulate=True)
            self._wgrad_gemm(dZY, mems[TB:], Gzy[:, H:], accumulate=True)
            bgrad(dZY, "last_to_zy_fc1.bias")
            dHlast, dmemT = buf("dHlast", B, H), buf("dmemT", B, mem)
            ops.gemm("nn", dZY, Wzy[:, :H], dHlast)
            ops.gemm("nn", dZY, Wzy[:, H:], dmemT)
            self._backward_mfn(P, G, dHlast, dmemT, [], wgrad, bgrad, lin_bwd, relu_scale)
        self.mark("bwd:lstm enc+mfn")
        self._join_side()
        self.mark("bwd:join wgrads")


def make_engine(configs, T, B, device, ops, head="l1", variant="mfm"):
    """The schedule object of a model variant."""
    if variant in VARIANTS:
        return AblationEngine(configs, T, B, device, ops, head=head, variant=variant)
    if variant == "missing":
        from .missing import MissingEngine
        return MissingEngine(configs, T, B, device, ops, head=head)
    return E.Engine(configs, T, B, device, ops, head=head, variant=variant)


# ---------------------------------------------------------------------------------------------------------------------
# module classes (drop-in for mfm_model.M_A .. M_D)
# ---------------------------------------------------------------------------------------------------------------------

class _Ablation(MFM):
    """Shared plumbing: the autograd function, engine cache, RNG state and pickling rules of ``MFM``."""
    _variant = "m_a"

    def _finish(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        [self.d_l, self.d_a, self.d_v] = config["input_dims"]
        [self.dh_l, self.dh_a, self.dh_v] = config["h_dims"]
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
            engs[key] = AblationEngine(self._cfg, T, B, device, _ops(), head="l1", variant=self._variant)
        return engs[key]

    def draw_mmd_noise(self, n, device):
        """The Gaussian samples of the model's loss_MMD calls, drawn in the reference's order (z_l, z_a, z_v, z_y minus the
        latents the model lacks); unused slots hold a placeholder."""
        c = self._cfg[0]
        sizes = (c["zl_size"], c["za_size"], c["zv_size"], c["zy_size"])
        used = dict(m_a=(0, 3), m_b=(0, 1, 2), m_c=(3,), m_d=())[self._variant]
        out = []
        for i, k in enumerate(sizes):
            if i not in used:
                out.append(torch.zeros(1, 1, device=device))
            elif self.mmd_noise == "cpu":
                out.append(torch.randn(n, k).to(device))
            else:
                out.append(torch.randn(n, k, device=device))
        return out


def _sizes(config):
    return (config["zy_size"], config["zl_size"], config["za_size"], config["zv_size"],
            config["fy_size"], config["fl_size"], config["fa_size"], config["fv_size"])


class M_A(_Ablation):
    """mfm_model.py:201-269: one encoder over the whole input, MFN, the three decoders read cat(fy, fl)."""
    _variant = "m_a"

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        nn.Module.__init__(self)
        d_l, d_a, d_v = config["input_dims"]
        zy, zl, za, zv, fy, fl, fa, fv = _sizes(config)
        last_mfn_size = sum(config["h_dims"]) + config["memsize"]
        # construction order fixes the init RNG stream (mfm_model.py:223-242)
        self.encoder_l = encoderLSTM(d_l + d_a + d_v, zl)
        self.decoder_l = decoderLSTM(fy + fl, d_l)
        self.decoder_a = decoderLSTM(fy + fl, d_a)
        self.decoder_v = decoderLSTM(fy + fl, d_v)
        self.mfn_encoder = MFN(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)
        self.last_to_zy_fc1 = nn.Linear(last_mfn_size, zy)
        self.zy_to_fy_fc1 = nn.Linear(zy, fy)
        self.zy_to_fy_fc2 = nn.Linear(fy, fy)
        self.zy_to_fy_dropout = nn.Dropout(config["zy_to_fy_dropout"])
        self.zl_to_fl_fc1 = nn.Linear(zl, fl)
        self.zl_to_fl_fc2 = nn.Linear(fl, fl)
        self.zl_to_fl_dropout = nn.Dropout(config["zl_to_fl_dropout"])
        self.fy_to_y_fc1 = nn.Linear(fy, fy)
        self.fy_to_y_fc2 = nn.Linear(fy, config["output_dim"])
        self.fy_to_y_dropout = nn.Dropout(config["fy_to_y_dropout"])
        self._finish(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)


class M_B(_Ablation):
    """mfm_model.py:271-343: no MFN and no z_y; decoders read f_l / f_a / f_v, the label comes from cat(fl, fa, fv)."""
    _variant = "m_b"

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        nn.Module.__init__(self)
        d_l, d_a, d_v = config["input_dims"]
        zy, zl, za, zv, fy, fl, fa, fv = _sizes(config)
        # construction order: mfm_model.py:293-315
        self.encoder_l = encoderLSTM(d_l, zl)
        self.encoder_a = encoderLSTM(d_a, za)
        self.encoder_v = encoderLSTM(d_v, zv)
        self.decoder_l = decoderLSTM(fl, d_l)
        self.decoder_a = decoderLSTM(fa, d_a)
        self.decoder_v = decoderLSTM(fv, d_v)
        self.zl_to_fl_fc1 = nn.Linear(zl, fl)
        self.zl_to_fl_fc2 = nn.Linear(fl, fl)
        self.zl_to_fl_dropout = nn.Dropout(config["zl_to_fl_dropout"])
        self.za_to_fa_fc1 = nn.Linear(za, fa)
        self.za_to_fa_fc2 = nn.Linear(fa, fa)
        self.za_to_fa_dropout = nn.Dropout(config["za_to_fa_dropout"])
        self.zv_to_fv_fc1 = nn.Linear(zv, fv)
        self.zv_to_fv_fc2 = nn.Linear(fv, fv)
        self.zv_to_fv_dropout = nn.Dropout(config["zv_to_fv_dropout"])
        self.fy_to_y_fc1 = nn.Linear(fl + fa + fv, fy)
        self.fy_to_y_fc2 = nn.Linear(fy, config["output_dim"])
        self.fy_to_y_dropout = nn.Dropout(config["fy_to_y_dropout"])
        self._finish(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)


class M_C(_Ablation):
    """mfm_model.py:345-403: the MFN only; the three decoders read fy."""
    _variant = "m_c"

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        nn.Module.__init__(self)
        d_l, d_a, d_v = config["input_dims"]
        zy, zl, za, zv, fy, fl, fa, fv = _sizes(config)
        last_mfn_size = sum(config["h_dims"]) + config["memsize"]
        # construction order: mfm_model.py:367-380
        self.decoder_l = decoderLSTM(fy, d_l)
        self.decoder_a = decoderLSTM(fy, d_a)
        self.decoder_v = decoderLSTM(fy, d_v)
        self.mfn_encoder = MFN(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)
        self.last_to_zy_fc1 = nn.Linear(last_mfn_size, zy)
        self.zy_to_fy_fc1 = nn.Linear(zy, fy)
        self.zy_to_fy_fc2 = nn.Linear(fy, fy)
        self.zy_to_fy_dropout = nn.Dropout(config["zy_to_fy_dropout"])
        self.fy_to_y_fc1 = nn.Linear(fy, fy)
        self.fy_to_y_fc2 = nn.Linear(fy, config["output_dim"])
        self.fy_to_y_dropout = nn.Dropout(config["fy_to_y_dropout"])
        self._finish(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)


class M_D(_Ablation):
    """mfm_model.py:405-467: purely discriminative -- three encoders, three factor MLPs, one Linear.  ``forward`` returns
    the INPUT slices as "decoded" and the python float 0.0 as mmd_loss, as the reference does (:456, :465)."""
    _variant = "m_d"

    def __init__(self, config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig):
        nn.Module.__init__(self)
        d_l, d_a, d_v = config["input_dims"]
        zy, zl, za, zv, fy, fl, fa, fv = _sizes(config)
        # construction order: mfm_model.py:427-443
        self.encoder_l = encoderLSTM(d_l, zl)
        self.encoder_a = encoderLSTM(d_a, za)
        self.encoder_v = encoderLSTM(d_v, zv)
        self.zl_to_fl_fc1 = nn.Linear(zl, fl)
        self.zl_to_fl_fc2 = nn.Linear(fl, fl)
        self.zl_to_fl_dropout = nn.Dropout(config["zl_to_fl_dropout"])
        self.za_to_fa_fc1 = nn.Linear(za, fa)
        self.za_to_fa_fc2 = nn.Linear(fa, fa)
        self.za_to_fa_dropout = nn.Dropout(config["za_to_fa_dropout"])
        self.zv_to_fv_fc1 = nn.Linear(zv, fv)
        self.zv_to_fv_fc2 = nn.Linear(fv, fv)
        self.zv_to_fv_dropout = nn.Dropout(config["zv_to_fv_dropout"])
        self.fs_to_y = nn.Linear(fl + fa + fv, config["output_dim"])
        self._finish(config, NN1Config, NN2Config, gamma1Config, gamma2Config, outConfig)

    def forward(self, x):
        decoded, _, missing_loss = MFM.forward(self, x)
        d_l, d_a = self.d_l, self.d_a
        return [x[:, :, :d_l], x[:, :, d_l:d_l + d_a], x[:, :, d_l + d_a:], decoded[3]], 0.0, missing_loss


ABLATION_MODELS = dict(m_a=M_A, m_b=M_B, m_c=M_C, m_d=M_D)
