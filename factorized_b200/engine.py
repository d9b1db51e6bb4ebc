"""Host-side schedule of the MFM training step over the CUDA primitive set.

The reference runs ``MFM.forward`` (mfm_model.py:522-555) as ~20 000 tiny
torch ops per step and lets autograd replay them.  Here the step is a fixed
schedule of a few dozen hand-written sm_100a kernels (``csrc/``, reached
through the C ABI in ``include/mfm_b200.h``): every time-parallel contraction
is hoisted out of the recurrences into large GEMMs over all T*B rows, each
recurrence (9 LSTM cells, the MFN memory) runs as one kernel with the
timestep loop inside, and the backward pass is the hand-derived adjoint of the
same schedule.  This file only decides *which* kernel runs on *which* buffer;
it does no arithmetic itself and has no CPU path: ``ops`` is the CUDA binding
(``factorized_b200.cuda_ops.CudaOps``).  Tests may inject another object with
the same primitive interface to check this schedule against autograd on a
machine without a GPU.

Buffer naming: ``[T*B, n]`` matrices are time-major row blocks (row t*B+b);
state histories have T+1 blocks with block 0 the zero initial state.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import torch

ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3

# dropout sites (index mixed into the counter-based RNG seed)
SITE_ATT1, SITE_ATT2, SITE_G1, SITE_G2, SITE_FY, SITE_FL, SITE_FA, SITE_FV, SITE_Y = range(1, 10)

TAGS = "lav"


class Dims:
    """Sizes read from the reference's six config dicts (mfm_model.py:472-489,96-114)."""

    def __init__(self, configs, T: int, B: int, head: str = "l1"):
        config, nn1, nn2, g1, g2, out = configs
        if config.get("windowsize", 2) != 2:
            raise ValueError("windowsize must be 2: MFN concatenates exactly (c_{t-1}, c_t) (mfm_model.py:171-173)")
        self.T, self.B = int(T), int(B)
        self.d = [int(v) for v in config["input_dims"]]
        self.D = sum(self.d)
        self.off = [0, self.d[0], self.d[0] + self.d[1]]
        self.hm = [int(v) for v in config["h_dims"]]
        self.H = sum(self.hm)
        self.hoff = [0, self.hm[0], self.hm[0] + self.hm[1]]
        self.z = [int(config["zl_size"]), int(config["za_size"]), int(config["zv_size"])]
        self.zy = int(config["zy_size"])
        self.f = [int(config["fl_size"]), int(config["fa_size"]), int(config["fv_size"])]
        self.fy = int(config["fy_size"])
        self.hd = [self.fy + f for f in self.f]
        self.mem = int(config["memsize"])
        self.a1, self.a2 = int(nn1["shapes"]), int(nn2["shapes"])
        self.g1, self.g2 = int(g1["shapes"]), int(g2["shapes"])
        self.out = int(config["output_dim"])
        self.p_att1, self.p_att2 = float(nn1["drop"]), float(nn2["drop"])
        self.p_g1, self.p_g2 = float(g1["drop"]), float(g2["drop"])
        self.p_f = [float(config["zl_to_fl_dropout"]), float(config["za_to_fa_dropout"]), float(config["zv_to_fv_dropout"])]
        self.p_fy = float(config["zy_to_fy_dropout"])
        self.p_y = float(config["fy_to_y_dropout"])
        self.lda = [float(config.get("lda_xl", 1.0)), float(config.get("lda_xa", 1.0)), float(config.get("lda_xv", 1.0))]
        self.lda_mmd = float(config.get("lda_mmd", 1.0))
        if head not in ("l1", "ce"):
            raise ValueError(head)
        self.head = head


class Engine:
    """One (T, B) instance of the schedule with its HBM workspace."""

    def __init__(self, configs, T: int, B: int, device, ops, head: str = "l1", mfn_only: bool = False,
                 mfn_prefix: str = "mfn_encoder.", variant: str = "mfm"):
        self.dm = Dims(configs, T, B, head)
        if variant not in ("mfm", "kl", "kl_ef"):
            raise ValueError(variant)
        # "kl": MFM_KL (mfm_model.py:662-764) -- the encoder outputs pass one more Linear to the means (the latents z) and
        # another to the log-variances, and the regulariser is loss_KLD instead of loss_MMD; everything else is MFM
        self.kl = variant in ("kl", "kl_ef")   # KL regulariser on (mean, log-variance) heads instead of the MMD
        self.ef = variant == "kl_ef"           # MFM_KL_EF (mfm_model.py:557-660): ONE early-fusion encoder cell instead of the MFN
        self.mfn_only = bool(mfn_only)          # standalone MFN module: only steps (1,2,4,5) on the MFN cells
        self.pre = mfn_prefix
        self.device = torch.device(device)
        self.ops = ops
        self.ws: Dict[str, torch.Tensor] = {}
        self.train = False
        self.loss_buf = torch.zeros(16, dtype=torch.float32, device=self.device)
        self.mmd_acc = torch.zeros(4, dtype=torch.float64, device=self.device)     # the four MMD terms, summed in double
        self.use_side_stream = True
        self._side = None
        self._side_used = set()   # side streams that received work in this schedule
        self._enc_stream = None          # the encoder cells' backward runs beside the MFN backward
        self._enc_used = False
        self._aux_streams = {}           # the MMD / KL terms of the four latents run on their own streams, side by side
        self._aux_used = set()
        self._pool = []                  # streams for _par branches
        self._z_ready = None             # events: encoder latents written (they are produced on the auxiliary streams)
        self._mmd_pending = False        # MMD accumulators (double) not yet folded into loss_buf[4:8]
        self._wcat_ready = False         # backward() built WcatAtt on an auxiliary stream
        self.n_side = max(1, int(os.environ.get("MFM_SIDE_STREAMS", "6")))    # weight-gradient streams
        self.n_aux = max(1, int(os.environ.get("MFM_AUX_STREAMS", "4")))      # MMD / KL streams (one per latent)
        self.want_mmd = True             # False: forward-only inference skips the O(B^2) MMD (its parts read 0)
        self.fused_dcext = os.environ.get("MFM_FUSED_DCEXT", "1") == "1"   # see _backward_mfn (0: separate gather pass)
        self.time_split = os.environ.get("MFM_TIME_SPLIT", "0") == "1"     # experiment, see _backward_mfn (default: one launch)
        self.stamps, self.stamp_names = None, []      # debug timeline (mark)
        # two launches for the last backward recurrence (heavy cells first, their weight gradients start early): measured
        # SLOWER (3.67 vs 3.56 ms/step: the gradient GEMMs delay the second launch), kept as an experiment switch
        self.split_last_recurrence = os.environ.get("MFM_SPLIT_LAST", "0") == "1"
        self.defer_mmd_join = False      # the fused trainer joins the MMD stream in losses() instead of at the end of forward
        # Fused trainer: the decoders' reconstruction head (fc1, mfm_model.py:88-90) produces the MSE terms and their
        # gradients in its GEMM epilogue (mfm_gemm_mse); x_hat never touches HBM and losses() launches no MSE kernel.
        self.fuse_mse = False
        # loss_buf: 0 disc, 1..3 mse_l/a/v, 4..7 mmd per latent (unweighted), 8 total (weighted)

    # -- debug timeline ---------------------------------------------------------------
    def mark(self, name: str):
        """Milestone on the current stream (only when a stamp buffer is attached: scripts/step_timeline.py)."""
        if self.stamps is not None:
            if name not in self.stamp_names:
                self.stamp_names.append(name)
            self.ops.stamp(self.stamps, self.stamp_names.index(name))

    # -- workspace -----------------------------------------------------------------
    def buf(self, name: str, *shape) -> torch.Tensor:
        t = self.ws.get(name)
        if t is None:
            t = torch.zeros(*shape, dtype=torch.float32, device=self.device)
            self.ws[name] = t
        return t

    def workspace_bytes(self) -> int:
        return sum(t.numel() * 4 for t in self.ws.values())

    # -- forward -------------------------------------------------------------------
    def forward(self, P: Dict[str, torch.Tensor], x: torch.Tensor, noise: Sequence[torch.Tensor],
                train: bool = False, rng: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """MFM.forward (mfm_model.py:522-555).  ``x`` is [T,B,D] contiguous fp32,
        ``noise`` the four Gaussian samples of loss_MMD (order zl,za,zv,zy).
        Returns views into the workspace: x_l_hat,x_a_hat,x_v_hat [T*B,d], y_hat
        [B,out], latents; MMD parts go to loss_buf[4:8]."""
        dm, ops, buf = self.dm, self.ops, self.buf
        T, B, H, mem = dm.T, dm.B, dm.H, dm.mem
        TB = T * B
        if tuple(x.shape) != (T, B, dm.D) or not x.is_contiguous() or x.dtype != torch.float32:
            raise ValueError("x must be contiguous fp32 [T=%d,B=%d,D=%d], got %s" % (T, B, dm.D, tuple(x.shape)))
        self.train = bool(train)
        self.x = x
        self.noise = list(noise)
        self.rng = rng
        self.x_in = x
        X2 = x.view(TB, dm.D)
        # (0) split x into per-modality matrices whose rows start 16 B aligned (leading dimension padded to a
        #     multiple of 4 floats): the reference layout concatenates the modalities on the last axis (:523-525), row
        #     pitch 4*D bytes (1300 B on MOSI), which TMA cannot describe.  One pass over x; the
        #     nine GEMM reads and three MSE reads of x that follow use the aligned copies.
        full = not self.mfn_only
        xs = [buf("Xp%d" % m, TB, (dm.d[m] + 3) // 4 * 4)[:, :dm.d[m]] for m in range(3)]
        self.xs = xs
        drop = (lambda p, site: (p, site) if (train and p > 0.0) else None)

        # (1) hoisted input projections  G_x = X W_ih^T + b_ih + b_hh   (encoders :56, MFN :167-169); one branch per modality.
        #     The encoder cell and the MFN cell of a modality read the same x: their G_x are column blocks of ONE matrix, and
        #     when the trainer has laid the two input weights (and bias vectors) out back to back the two projections are one
        #     GEMM with N = 4(z + h) -- x is streamed once (text: 93 -> 57 us).
        def adjacent(a, b):
            return (a.is_contiguous() and b.is_contiguous() and a.dim() == b.dim() and a.shape[1:] == b.shape[1:]
                    and a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()
                    and a.data_ptr() + a.numel() * 4 == b.data_ptr())

        def joined(a, b):
            shape = (a.shape[0] + b.shape[0],) + tuple(a.shape[1:])
            return torch.as_strided(a, shape, a.stride())

        def project(m):
            def run():
                tag = TAGS[m]
                e, n = "encoder_%s.lstm" % tag, self.pre + "lstm_%s" % tag
                ops.copy2d(X2[:, dm.off[m]:dm.off[m] + dm.d[m]], xs[m])
                ze, zn = (4 * dm.z[m] if full else 0), (0 if self.ef else 4 * dm.hm[m])
                gcat = buf("GxCat%d" % m, TB, ze + zn)
                if full:
                    self.ws["GxE%d" % m] = gcat[:, :ze]
                if not self.ef:
                    self.ws["GxN%d" % m] = gcat[:, ze:]
                if full and not self.ef and all(
                        adjacent(P[e + leaf], P[n + leaf]) for leaf in (".weight_ih", ".bias_ih", ".bias_hh")):
                    ops.gemm("nt", xs[m], joined(P[e + ".weight_ih"], P[n + ".weight_ih"]), gcat,
                             bias=joined(P[e + ".bias_ih"], P[n + ".bias_ih"]), bias2=joined(P[e + ".bias_hh"], P[n + ".bias_hh"]))
                    return
                if full:
                    ops.gemm("nt", xs[m], P[e + ".weight_ih"], gcat[:, :ze], bias=P[e + ".bias_ih"], bias2=P[e + ".bias_hh"])
                if not self.ef:
                    ops.gemm("nt", xs[m], P[n + ".weight_ih"], gcat[:, ze:], bias=P[n + ".bias_ih"], bias2=P[n + ".bias_hh"])
            return run

        def project_ef():                                      # the early-fusion cell reads the whole row of x (:639)
            ops.gemm("nt", X2, P["ef_encoder.lstm.weight_ih"], buf("GxEF", TB, 4 * hef),
                     bias=P["ef_encoder.lstm.bias_ih"], bias2=P["ef_encoder.lstm.bias_hh"])

        self.mark("fwd:start")
        if self.fuse_mse and full:
            ops.zero(self.loss_buf[0:4])                       # the decoder heads accumulate the MSE terms during forward
        hef = sum(dm.z)
        self._par([project(0), project(1), project(2)] + ([project_ef] if self.ef else []))
        self.mark("fwd:projections")

        # (2) six recurrences in one launch: h W_hh^T + G_x[t] -> gates -> (h, c)
        Hall = buf("Hall", (T + 1) * B, H)
        # The attention reads cat(c_{t-1}, c_t) (:171-173).  The cell histories are written straight into that layout:
        # row block i of CS2 = [ c of history block i-1 | c of history block i ]; the kernels write every c block to the
        # right half of its own row block (cs) and to the left half of the next one (cs_dup).  cStar is then the
        # contiguous view CS2[B : (T+1)B] -- no copy.
        CS2 = buf("CS2", (T + 2) * B, 2 * H)
        Call = CS2[:(T + 1) * B, H:]
        Cdup = CS2[B:, :H]
        self.ws_views = dict(Call=Call)
        cells = []
        for m, tag in enumerate(TAGS if full else ""):
            cells.append(dict(T=T, B=B, h=dm.z[m], gx=self.ws["GxE%d" % m], gx_steps=T, bias_rest=None,
                              W=P["encoder_%s.lstm.weight_hh" % tag],
                              hs=buf("hsE%d" % m, (T + 1) * B, dm.z[m]), cs=buf("csE%d" % m, (T + 1) * B, dm.z[m]),
                              gates=buf("gatesE%d" % m, TB, 4 * dm.z[m])))
        if self.ef:
            cells.append(dict(T=T, B=B, h=hef, gx=self.ws["GxEF"], gx_steps=T, bias_rest=None,
                              W=P["ef_encoder.lstm.weight_hh"], hs=buf("hsEF", (T + 1) * B, hef),
                              cs=buf("csEF", (T + 1) * B, hef), gates=buf("gatesEF", TB, 4 * hef)))
        for m, tag in enumerate("" if self.ef else TAGS):
            o = dm.hoff[m]
            cells.append(dict(T=T, B=B, h=dm.hm[m], gx=self.ws["GxN%d" % m], gx_steps=T, bias_rest=None,
                              W=P[self.pre + "lstm_%s.weight_hh" % tag],
                              hs=Hall[:, o:o + dm.hm[m]], cs=Call[:, o:o + dm.hm[m]], cs_dup=Cdup[:, o:o + dm.hm[m]],
                              gates=buf("gatesN%d" % m, TB, 4 * dm.hm[m])))
        ops.lstm_fwd(cells)
        self.mark("fwd:lstm enc+mfn")

        # (3) z_m = fc1(h_T), no activation (:60-61), and (7) the MMD of the three encoder latents: nothing on the MFN
        #     path needs them, so they leave the main stream here and overlap the attention / memory chain
        Z = [buf("Z%d" % m, B, dm.z[m]) for m in range(3)] if full else []
        if full:
            if self.kl:
                ops.zero(self.loss_buf[4:8])
            else:
                ops.zero(self.mmd_acc.view(torch.float32))
            self._z_ready = []
            for m, tag in enumerate(TAGS):
                with self._aux(m):
                    zlast = buf("Zlast%d" % m, B, dm.z[m]) if self.kl else Z[m]
                    ops.gemm("nt", self.ws["hsE%d" % m][TB:], P["encoder_%s.fc1.weight" % tag], zlast,
                             bias=P["encoder_%s.fc1.bias" % tag])
                    if self.kl:                                   # means and log-variances (mfm_model.py:735-740), KL term (:746)
                        lv = buf("LV%d" % m, B, dm.z[m])
                        ops.gemm("nt", zlast, P["last_to_z%s_fc1.weight" % tag], Z[m], bias=P["last_to_z%s_fc1.bias" % tag])
                        ops.gemm("nt", zlast, P["last_to_logvarz%s_fc1.weight" % tag], lv, bias=P["last_to_logvarz%s_fc1.bias" % tag])
                        ops.kld_fwd(Z[m], lv, self.loss_buf[4 + m:5 + m])
                    self._z_ready.append(self._aux_event(m))
                    if not self.kl and self.want_mmd:
                        self._mmd(m, Z[m])

        if self.ef:
            ZY = self._forward_ef_head(P)
        else:
            ZY = self._forward_mfn_head(P, CS2, Hall, drop, rng)
            if self.mfn_only:
                return ZY
        for ev in (self._z_ready or []):                   # the factor MLPs read Z, produced on the auxiliary streams
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
        self._z_ready = None

        # (8) factor MLPs (:539-542) written straight into the decoder inputs cat(fy, f_m) (:544-546).
        #     The four MLPs, and below the three decoders, are independent of each other: branches run side by side.
        FY = buf("FY", B, dm.fy)
        F1y = buf("F1y", B, dm.fy)
        EMB = [buf("EMB%d" % m, B, dm.hd[m]) for m in range(3)]
        F1s = [buf("F1_%d" % m, B, dm.f[m]) for m in range(3)]

        def mlp_y():
            ops.gemm("nt", ZY, P["zy_to_fy_fc1.weight"], F1y, bias=P["zy_to_fy_fc1.bias"], act=ACT_RELU,
                     drop=drop(dm.p_fy, SITE_FY), rng=rng)
            ops.gemm("nt", F1y, P["zy_to_fy_fc2.weight"], FY, bias=P["zy_to_fy_fc2.bias"], act=ACT_RELU)

        def mlp_m(m):
            def run():
                nm = "z%s_to_f%s" % (TAGS[m], TAGS[m])
                ops.gemm("nt", Z[m], P[nm + "_fc1.weight"], F1s[m], bias=P[nm + "_fc1.bias"], act=ACT_RELU,
                         drop=drop(dm.p_f[m], SITE_FL + m), rng=rng)
                ops.gemm("nt", F1s[m], P[nm + "_fc2.weight"], EMB[m][:, dm.fy:], bias=P[nm + "_fc2.bias"], act=ACT_RELU)
                # the decoder's merged recurrent weight depends on the parameters only: it rides along here
                d_ = "decoder_%s.lstm" % TAGS[m]
                hd = dm.hd[m]
                ops.add(P[d_ + ".weight_ih"], P[d_ + ".weight_hh"], buf("Wm%d" % m, 4 * hd, hd))
                ops.add(P[d_ + ".bias_ih"].view(1, -1), P[d_ + ".bias_hh"].view(1, -1), buf("bsumD%d" % m, 1, 4 * hd))
            return run

        self.mark("fwd:zy")
        self._par([mlp_y, mlp_m(0), mlp_m(1), mlp_m(2)])
        self.mark("fwd:factor MLPs")

        # (9) decoders (:72-91): step 0 eats the embedding; for t>=1 the input IS h_{t-1}, so the two
        #     gate GEMMs collapse into one with W_ih + W_hh.   (10) reconstructions x_hat = fc1(all hiddens) (:88-90)
        # (11) discriminative head (:552) -- depends on fy only, a fifth branch
        Xhat = [None] * 3 if self.fuse_mse else [buf("Xhat%d" % m, TB, dm.d[m]) for m in range(3)]
        dXf = [buf("dXhat%d" % m, TB, dm.d[m]) for m in range(3)] if self.fuse_mse else None
        Y1 = buf("Y1", B, dm.fy)
        Yhat = buf("Yhat", B, dm.out)

        def decoder(m):
            def run():
                tag = TAGS[m]
                d_ = "decoder_%s.lstm" % tag
                hd = dm.hd[m]
                ops.copy2d(FY, EMB[m][:, :dm.fy])
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
            ops.gemm("nt", FY, P["fy_to_y_fc1.weight"], Y1, bias=P["fy_to_y_fc1.bias"], act=ACT_RELU,
                     drop=drop(dm.p_y, SITE_Y), rng=rng)
            ops.gemm("nt", Y1, P["fy_to_y_fc2.weight"], Yhat, bias=P["fy_to_y_fc2.bias"])

        self._par([decoder(0), decoder(1), decoder(2), head])
        self.mark("fwd:decoders+head")
        if not self.defer_mmd_join:
            self._join_aux()
        return dict(x_l_hat=Xhat[0], x_a_hat=Xhat[1], x_v_hat=Xhat[2], y_hat=Yhat,
                    zl=Z[0], za=Z[1], zv=Z[2], zy=ZY, fy=FY, mmd_parts=self.loss_buf[4:8])

    def _forward_mfn_head(self, P, CS2, Hall, drop, rng):
        """Steps (4)-(7) of forward for the MFN variants: attention, memory recurrence, z_y and its regulariser.  Returns z_y
        (or, for a standalone MFN, its output dict)."""
        dm, ops, buf = self.dm, self.ops, self.buf
        T, B, H, mem = dm.T, dm.B, dm.H, dm.mem
        TB = T * B
        # (4) MFN attention for all T at once (:171-176); only gamma*_fc1's memory columns are sequential
        pre = self.pre
        cStar = CS2[B:(T + 1) * B]
        self.ws_views["cStar"] = cStar
        H1 = buf("H1", TB, dm.a1)
        ops.gemm("nt", cStar, P[pre + "att1_fc1.weight"], H1, bias=P[pre + "att1_fc1.bias"], act=ACT_RELU,
                 drop=drop(dm.p_att1, SITE_ATT1), rng=rng)
        Att = buf("Att", TB, 2 * H)
        ops.gemm("nt", H1, P[pre + "att1_fc2.weight"], Att, bias=P[pre + "att1_fc2.bias"])
        Attended = buf("Attended", TB, 2 * H)
        ops.softmax_gate_fwd(Att, cStar, Attended)
        self.mark("fwd:att1+gate")
        H2 = buf("H2", TB, dm.a2)
        cHat = buf("cHat", TB, mem)
        G1pre = buf("G1pre", TB, dm.g1)
        G2pre = buf("G2pre", TB, dm.g2)
        Wg1, Wg2 = P[pre + "gamma1_fc1.weight"], P[pre + "gamma2_fc1.weight"]

        def att2():
            ops.gemm("nt", Attended, P[pre + "att2_fc1.weight"], H2, bias=P[pre + "att2_fc1.bias"], act=ACT_RELU,
                     drop=drop(dm.p_att2, SITE_ATT2), rng=rng)
            ops.gemm("nt", H2, P[pre + "att2_fc2.weight"], cHat, bias=P[pre + "att2_fc2.bias"], act=ACT_TANH)

        self._par([att2,
                   lambda: ops.gemm("nt", Attended, Wg1[:, :2 * H], G1pre, bias=P[pre + "gamma1_fc1.bias"]),
                   lambda: ops.gemm("nt", Attended, Wg2[:, :2 * H], G2pre, bias=P[pre + "gamma2_fc1.bias"])])

        self.mark("fwd:att2+gamma")
        # (5) the memory recurrence (:177-180), T steps in one kernel
        mems = buf("mems", (T + 1) * B, mem)
        ops.mfn_mem_fwd(dict(
            T=T, B=B, mem=mem, g1=dm.g1, g2=dm.g2, G1pre=G1pre, G2pre=G2pre, cHat=cHat,
            W1m=Wg1[:, 2 * H:], W2m=Wg2[:, 2 * H:],
            W12=P[pre + "gamma1_fc2.weight"], b12=P[pre + "gamma1_fc2.bias"],
            W22=P[pre + "gamma2_fc2.weight"], b22=P[pre + "gamma2_fc2.bias"],
            mems=mems, U1=buf("U1", TB, dm.g1), U2=buf("U2", TB, dm.g2),
            Gam1=buf("Gam1", TB, mem), Gam2=buf("Gam2", TB, mem),
            drop1=drop(dm.p_g1, SITE_G1), drop2=drop(dm.p_g2, SITE_G2), rng=rng))

        self.mark("fwd:mem recurrence")
        if self.mfn_only:                                     # MFN.forward returns cat(h_T^l,h_T^a,h_T^v,mem_T) (:194-198)
            last = buf("mfn_last", B, H + mem)
            ops.copy2d(Hall[TB:], last[:, :H])
            ops.copy2d(mems[TB:], last[:, H:])
            return dict(mfn_last=last)          # (MFN.forward's own return value)

        # (6) zy = last_to_zy_fc1(cat(h_T^l, h_T^a, h_T^v, mem_T))  (:194-198, :535)
        ZY = buf("ZY", B, dm.zy)
        Wzy = P["last_to_zy_fc1.weight"]
        ops.gemm("nt", Hall[TB:], Wzy[:, :H], ZY, bias=P["last_to_zy_fc1.bias"])
        ops.gemm("nt", mems[TB:], Wzy[:, H:], ZY, accumulate=True)

        # (7) MMD of z_y (the encoder latents went out in step 3); same auxiliary stream, off the critical path
        with self._aux(3):
            if self.kl:                                           # log-variance of z_y and its KL term (mfm_model.py:744,746)
                Wlv = P["last_to_logvarzy_fc1.weight"]
                LVY = buf("LVY", B, dm.zy)
                ops.gemm("nt", Hall[TB:], Wlv[:, :H], LVY, bias=P["last_to_logvarzy_fc1.bias"])
                ops.gemm("nt", mems[TB:], Wlv[:, H:], LVY, accumulate=True)
                ops.kld_fwd(ZY, LVY, self.loss_buf[7:8])
            else:
                if self.want_mmd:
                    self._mmd(3, ZY)
        return ZY

    def _forward_ef_head(self, P):
        """MFM_KL_EF (mfm_model.py:639-643): z_y and its log-variance are Linears of the early-fusion encoder's output."""
        dm, ops, buf = self.dm, self.ops, self.buf
        B, TB, hef = dm.B, dm.T * dm.B, sum(dm.z)
        EFlast = buf("EFlast", B, hef)
        ops.gemm("nt", self.ws["hsEF"][TB:], P["ef_encoder.fc1.weight"], EFlast, bias=P["ef_encoder.fc1.bias"])
        ZY = buf("ZY", B, dm.zy)
        ops.gemm("nt", EFlast, P["last_to_zy_fc1.weight"], ZY, bias=P["last_to_zy_fc1.bias"])
        with self._aux(3):
            LVY = buf("LVY", B, dm.zy)
            ops.gemm("nt", EFlast, P["last_to_logvarzy_fc1.weight"], LVY, bias=P["last_to_logvarzy_fc1.bias"])
            ops.kld_fwd(ZY, LVY, self.loss_buf[7:8])
        return ZY

    # -- losses (the fused training path) ------------------------------------------
    def losses(self, y: torch.Tensor):
        """Loss terms of mfm_mosi.py:432-439 (L1) / mfm_mosi_acc.py:441-451 (CE) and
        their gradients w.r.t. the forward outputs, written to dXhat*/dYhat.
        loss_buf[0..3] = disc, mse_l, mse_a, mse_v (means); loss_buf[8] = total."""
        dm, ops, buf = self.dm, self.ops, self.buf
        TB = dm.T * dm.B
        # loss_buf[4:8] (MMD parts) come from the auxiliary stream.  Only the scalar total needs their VALUES (no
        # gradient does), so the fused trainer does not wait for the MMD here: backward() sums the total after it has
        # joined the auxiliary stream anyway (the loss gradients do not depend on it).
        if not self.defer_mmd_join:
            self._join_aux()
        if not self.fuse_mse:
            ops.zero(self.loss_buf[0:4])
        dX = [buf("dXhat%d" % m, TB, dm.d[m]) for m in range(3)]
        dY = buf("dYhat", dm.B, dm.out)

        def mse(m):
            if self.fuse_mse:                      # already done by the decoder heads' GEMM epilogue
                return None
            def run():
                n = float(TB * dm.d[m])
                ops.mse_fwd_bwd(self.ws["Xhat%d" % m], self.xs[m], 1.0 / n, 2.0 * dm.lda[m] / n,
                                self.loss_buf[1 + m:2 + m], dX[m])
            return run

        def disc():
            if dm.head == "l1":
                ops.l1_fwd_bwd(self.ws["Yhat"], y.view(dm.B, dm.out), 1.0 / (dm.B * dm.out), self.loss_buf[0:1], dY)
            else:
                ops.ce_fwd_bwd(self.ws["Yhat"], y, 1.0 / dm.B, self.loss_buf[0:1], dY)

        self.mark("loss:join mmd")
        self._par([mse(0), mse(1), mse(2), disc])
        self.mark("loss:heads")
        if not self.defer_mmd_join:
            ops.loss_total(self.loss_buf, dm.lda[0], dm.lda[1], dm.lda[2], dm.lda_mmd)
        return dX, dY

    # -- weight-gradient GEMMs run on a side stream ---------------------------------
    def _on_side(self, key, fn, index=None):
        """Run fn on the side stream selected by key (or, explicitly, by index), ordered after everything issued so far on the
        current stream."""
        if self.device.type != "cuda" or not self.use_side_stream:
            fn()
            return
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(self.n_side)]
        i = ((key >> 8) if index is None else index) % len(self._side)
        st = self._side[i]
        ev = torch.cuda.Event()
        ev.record(main)
        st.wait_event(ev)
        with torch.cuda.stream(st):
            fn()
        self._side_used.add(i)

    def _wgrad_pair(self, dY, A1, G1, b1, A2, G2, b2=None, index=None):
        """dW1 += dY^T A1, dW2 += dY^T A2 and the bias gradient(s) = column sums of dY, in one launch on a side stream: the
        two weight gradients of a cell share dY, which is then read from HBM once.  b2 (same sums) is copied from b1."""
        ops = self.ops

        def run():
            ops.gemm_tn_pair(dY, A1, G1, b1, A2, G2)
            if b2 is not None:
                ops.copy2d(b1.view(1, -1), b2.view(1, -1), accumulate=True)
        self._on_side(G1.data_ptr(), run, index)

    def _wgrad_gemm(self, dY, A, Gout, stream_key=None, **kw):
        """dW += dY^T A (TN GEMM, optionally with the fused bias gradient).  Weight gradients only read stashes and
        write the flat gradient buffer, so they do not belong on the critical path of the backward chain: they are
        issued on a side stream that forks from the main stream right after dY is produced and joins before the
        optimizer step (inside a CUDA graph this becomes a parallel branch)."""
        ops = self.ops
        if self.device.type != "cuda" or not self.use_side_stream:
            ops.gemm("tn", dY, A, Gout, **kw)
            return
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(self.n_side)]
        # independent weight gradients also overlap each other; two GEMMs into the same tensor share a stream
        i = ((stream_key if stream_key is not None else Gout.data_ptr()) >> 8) % len(self._side)
        st = self._side[i]
        ev = torch.cuda.Event()
        ev.record(main)
        st.wait_event(ev)
        with torch.cuda.stream(st):
            ops.gemm("tn", dY, A, Gout, **kw)
        self._side_used.add(i)

    def _join_side(self):
        if self._side is not None and self._side_used:
            # only the streams that received work in this schedule: inside a graph capture, waiting on a stream that was never
            # forked from the capturing stream is an error (cudaErrorStreamCaptureIsolation)
            main = torch.cuda.current_stream(self.device)
            for i in sorted(self._side_used):
                main.wait_stream(self._side[i])
            self._side_used = set()
        if self._enc_used:
            torch.cuda.current_stream(self.device).wait_stream(self._enc_stream)
            self._enc_used = False

    # -- independent small branches run side by side -------------------------------------
    def _par(self, thunks):
        """Run independent branches concurrently: branch 0 stays on the current stream, the others fork onto pool
        streams and are joined before returning.  The tails of forward and backward are dozens of batch-row-sized
        kernels (factor MLPs, heads, decoder input projections) of a few CTAs each; serialised they cost ~10 us apiece
        on the critical path, side by side they cost one chain.  Sequential on the CPU test double."""
        thunks = [t for t in thunks if t is not None]
        if self.device.type != "cuda" or not self.use_side_stream or len(thunks) <= 1:
            for t in thunks:
                t()
            return
        main = torch.cuda.current_stream(self.device)
        while len(self._pool) < len(thunks) - 1:
            self._pool.append(torch.cuda.Stream(device=self.device, priority=-1))
        ev = torch.cuda.Event()
        ev.record(main)
        used = []
        for i, t in enumerate(thunks[1:]):
            st = self._pool[i]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                t()
            used.append(st)
        thunks[0]()
        for st in used:
            main.wait_stream(st)

    # -- the MMD statistic is off the critical path: it runs on its own stream -------
    class _Aux:
        """``with eng._aux(k):`` runs the body on auxiliary stream k, forked from the main stream at entry."""

        def __init__(self, eng, k):
            self.eng, self.k, self.ctx = eng, k % eng.n_aux, None

        def __enter__(self):
            e = self.eng
            if e.device.type != "cuda" or not e.use_side_stream:
                return self
            st = e._aux_streams.get(self.k)
            if st is None:
                st = e._aux_streams[self.k] = torch.cuda.Stream(device=e.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(e.device))
            st.wait_event(ev)
            self.ctx = torch.cuda.stream(st)
            self.ctx.__enter__()
            e._aux_used.add(self.k)
            return self

        def __exit__(self, *a):
            if self.ctx is not None:
                self.ctx.__exit__(*a)
            return False

    def _aux(self, k=0):
        return Engine._Aux(self, k)

    def _aux_event(self, k=0):
        """Event at the current point of auxiliary stream k (None on the CPU test double)."""
        if self.device.type != "cuda" or not self.use_side_stream:
            return None
        ev = torch.cuda.Event()
        ev.record(self._aux_streams[k % self.n_aux])
        return ev

    def _mmd(self, k, zk):
        """loss_MMD of latent k against its Gaussian sample (:25-34, :536) on the current (auxiliary) stream.
        Pair matrices S = X Y^T on the tensor cores, K = exp(-(|x|^2+|y|^2-2S)/dim^2) in place; K(z,z) and K(g,z) stay
        resident (no [B,B,dim] tensor).  Everything of the gradient that does not depend on dLoss/dMMD follows at once:
        d/dz = (2c/B^2) [ (rowsum Kzz - colsum Kgz) * z - Kzz Z + Kgz^T G ], c = -2/dim^2 -- the column sums and the two
        products; backward() only combines them.  So neither direction of the MMD sits on the step's critical path."""
        ops, buf, B = self.ops, self.buf, self.dm.B
        self._mmd_pending = True
        gk, dim = self.noise[k], zk.shape[1]
        nz, ng = buf("mmd_nz%d" % k, B), buf("mmd_ng%d" % k, B)
        Kzz, Kgz, Kgg = buf("Kzz%d" % k, B, B), buf("Kgz%d" % k, B, B), buf("Kgg%d" % k, B, B)
        inv_bb = 1.0 / (float(B) * float(B))
        slot = self.mmd_acc[k:k + 1]                   # double: the three means cancel to O(1/B)
        ops.rownorm2(zk, nz)
        ops.rownorm2(gk, ng)
        ops.gemm("nt", zk, zk, Kzz)
        ops.mmd_kexp64(Kzz, nz, nz, dim, inv_bb, slot)
        ops.gemm("nt", gk, zk, Kgz)                   # rows index the Gaussian sample, columns the latent
        ops.mmd_kexp64(Kgz, ng, nz, dim, -2.0 * inv_bb, slot)
        ops.gemm("nt", gk, gk, Kgg)
        ops.mmd_kexp64(Kgg, ng, ng, dim, inv_bb, slot)
        rc, t12 = buf("mmd_rc%d" % k, 2 * B), buf("mmd_t12_%d" % k, 2 * B, dim)
        ops.zero(rc)
        ops.colsum(Kzz, rc[:B])                          # K(z,z) is symmetric: column sums == row sums
        ops.colsum(Kgz, rc[B:])
        ops.zero(t12)                                    # accumulate form lets the GEMM split K = B over CTAs
        ops.gemm("nn", Kzz, zk, t12[:B], accumulate=True)
        ops.gemm("tn", Kgz, gk, t12[B:], accumulate=True)

    def _join_aux(self):
        """Join the auxiliary streams; the four double MMD accumulators are folded into the fp32 loss buffer here."""
        if self._aux_used:
            main = torch.cuda.current_stream(self.device)
            for k in sorted(self._aux_used):
                main.wait_stream(self._aux_streams[k])
            self._aux_used = set()
        if self._mmd_pending:
            self.ops.mmd_fold(self.mmd_acc, self.loss_buf[4:8])
            self._mmd_pending = False

    # -- backward ------------------------------------------------------------------
    def backward(self, P: Dict[str, torch.Tensor], G: Dict[str, torch.Tensor], dXhat: Sequence[torch.Tensor],
                 dYhat: torch.Tensor, mmd_scale: float, mmd_scale_dev: Optional[torch.Tensor] = None,
                 d_mfn_last: Optional[torch.Tensor] = None, d_latents: Optional[Sequence[torch.Tensor]] = None):
        """Adjoint of ``forward``.  ``dXhat[m]`` [T*B,d_m] and ``dYhat`` [B,out] are
        d(loss)/d(output); d(loss)/d(mmd) = ``mmd_scale`` (host float) times the optional
        device scalar ``mmd_scale_dev``.  Parameter gradients are ACCUMULATED into ``G``
        (same names as ``P``), which the caller zeroes.  ``d_latents`` (zl, za, zv, zy): gradients a caller formed on the
        latents OUTSIDE the step (the MFM of mfm_mosi_acc.py returns them and its loop applies loss_MMD itself, :394, :441);
        they take the place of the regulariser's own gradient."""
        dm, ops, buf, ws = self.dm, self.ops, self.buf, self.ws
        T, B, H, mem = dm.T, dm.B, dm.H, dm.mem
        TB = T * B
        relu_scale = (lambda p: 1.0 / (1.0 - p) if (self.train and p > 0.0) else 1.0)

        def wgrad(dY, A, name):            # dW[N,K] += dY[M,N]^T A[M,K]
            self._wgrad_gemm( dY, A, G[name], accumulate=True)

        def bgrad(dY, name):
            ops.colsum(dY, G[name])

        def lin_bwd(dY, A, name, dA=None, accumulate=False, mask=None, mask_scale=1.0):
            """y = A W^T + b: dW and db in one GEMM (db rides along as a ones column), and (optionally)
            dA = (dY W) [* relu/dropout mask of A]."""
            self._wgrad_gemm( dY, A, G[name + ".weight"], accumulate=True, colsum_out=G[name + ".bias"])
            if dA is not None:
                ops.gemm("nn", dY, P[name + ".weight"], dA, accumulate=accumulate, mask=mask, mask_scale=mask_scale)

        if self.mfn_only:
            dHlast, dmemT = d_mfn_last[:, :H], d_mfn_last[:, H:]
            self._backward_mfn(P, G, dHlast, dmemT, [], wgrad, bgrad, lin_bwd, relu_scale)
            self._join_side()
            return

        # (7') MMD: the column sums and products were formed right after the forward kernels (_mmd); what is left is one
        #      combine per latent, scaled by dLoss/dMMD = ``mmd_scale`` x ``mmd_scale_dev``
        lat = [ws["Z0"], ws["Z1"], ws["Z2"], ws["ZY"]]
        dmmd = [buf("dZmmd%d" % k, B, lat[k].shape[1]) for k in range(4)]
        dLV = [buf("dLV%d" % k, B, lat[k].shape[1]) for k in range(4)] if self.kl else None
        if not self.ef:                                      # (parameters only: off the main stream, joined with the MMD streams)
            with self._aux(0):
                self._build_wcat(P)
            self._wcat_ready = True
        for k in range(4):
            with self._aux(k):
                ops.zero(dmmd[k])
                if d_latents is not None:
                    ops.copy2d(d_latents[k], dmmd[k])
                elif self.kl:                                     # d KLD / d mu and / d logvar, scaled by dLoss/dKLD
                    ops.kld_bwd(lat[k], ws["LV%d" % k] if k < 3 else ws["LVY"], mmd_scale, dmmd[k], dLV[k], mmd_scale_dev)
                else:
                    rc, t12 = ws["mmd_rc%d" % k], ws["mmd_t12_%d" % k]
                    ops.mmd_combine(lat[k], rc[:B], rc[B:], t12[:B], t12[B:], mmd_scale, dmmd[k], mmd_scale_dev)

        # (11') head, (10') + (9') + (8') one chain per decoder: reconstruction head, recurrence, weight gradients, embedding,
        #       factor MLP.  The four chains touch disjoint buffers (the decoders' shares of dFY are added after the join).
        dY1 = buf("dY1", B, dm.fy)
        dFY = buf("dFY", B, dm.fy)
        dZ = [buf("dZ%d" % m, B, dm.z[m]) for m in range(3)]
        dZY = buf("dZY", B, dm.zy)
        dEMB = [buf("dEMB%d" % m, B, dm.hd[m]) for m in range(3)]

        def mlp2_bwd(df, f, F1, zin, nm, p, dz):
            dpre = buf("dpre_" + nm, f.shape[0], f.shape[1])
            ops.relu_bwd(df, f, dpre)
            dF1 = buf("dF1_" + nm, F1.shape[0], F1.shape[1])
            lin_bwd(dpre, F1, nm + "_fc2", dF1, mask=F1, mask_scale=relu_scale(p))
            lin_bwd(dF1, zin, nm + "_fc1", dz)

        def head_bwd():
            lin_bwd(dYhat, ws["Y1"], "fy_to_y_fc2", dY1, mask=ws["Y1"], mask_scale=relu_scale(dm.p_y))
            lin_bwd(dY1, ws["FY"], "fy_to_y_fc1", dFY)

        def decoder_bwd(m):
            def run():
                tag = TAGS[m]
                hd = dm.hd[m]
                d_ = "decoder_%s.lstm" % tag
                dHd = buf("dHd%d" % m, TB, hd)
                lin_bwd(dXhat[m], ws["hsD%d" % m][B:], "decoder_%s.fc1" % tag, dHd)
                dG = buf("dGD%d" % m, TB, 4 * hd)
                ops.lstm_bwd([dict(T=T, B=B, h=hd, gates=ws["gatesD%d" % m], cs=ws["csD%d" % m], W=ws["Wm%d" % m],
                                   dh_all=dHd, dh_last=None, dc_ext=None, dG=dG, dc_scratch=buf("dcSD%d" % m, B, hd))])
                hprev = ws["hsD%d" % m][:TB]                    # h_{t-1}; block 0 is the zero state
                # for t >= 1 the input IS h_{t-1} (:85): dW_ih and dW_hh share the product dG^T h_prev (and the bias
                # gradients are the same column sums), so it is computed once and added to the second gradient; step 0's
                # input is the embedding (:83).  All three updates of dW_ih stay on one side stream (keyed by its pointer).
                key = G[d_ + ".weight_ih"].data_ptr()
                self._wgrad_gemm(dG, hprev, G[d_ + ".weight_hh"], accumulate=True, colsum_out=G[d_ + ".bias_hh"], stream_key=key)
                self._on_side(key, lambda: (ops.copy2d(G[d_ + ".weight_hh"], G[d_ + ".weight_ih"], accumulate=True),
                                            ops.copy2d(G[d_ + ".bias_hh"].view(1, -1), G[d_ + ".bias_ih"].view(1, -1), accumulate=True)))
                self._wgrad_gemm(dG[:B], ws["EMB%d" % m], G[d_ + ".weight_ih"], accumulate=True, stream_key=key)
                ops.gemm("nn", dG[:B], P[d_ + ".weight_ih"], dEMB[m])
                mlp2_bwd(dEMB[m][:, dm.fy:], ws["EMB%d" % m][:, dm.fy:], ws["F1_%d" % m], ws["Z%d" % m],
                         "z%s_to_f%s" % (tag, tag), dm.p_f[m], dZ[m])
            return run

        self.mark("bwd:start")
        self._par([decoder_bwd(0), decoder_bwd(1), decoder_bwd(2), head_bwd])
        self.mark("bwd:decoder chains")
        for m in range(3):
            ops.copy2d(dEMB[m][:, :dm.fy], dFY, accumulate=True)
        mlp2_bwd(dFY, ws["FY"], ws["F1y"], ws["ZY"], "zy_to_fy", dm.p_fy, dZY)

        self.mark("bwd:mlp y")
        self._join_aux()
        self.mark("bwd:join mmd")
        if self.defer_mmd_join:              # fused trainer: the total (loss_buf[8]) is summed now that the MMD parts exist
            ops.loss_total(self.loss_buf, dm.lda[0], dm.lda[1], dm.lda[2], dm.lda_mmd)
        dlat = dZ + [dZY]
        for k in range(4):
            ops.copy2d(dmmd[k], dlat[k], accumulate=True)

        dHlast = dmemT = None
        if self.ef:
            # (6') MFM_KL_EF: z_y and its log-variance are Linears of the early-fusion encoder's output (:640-641)
            hef = sum(dm.z)
            dEF = buf("dEFlast", B, hef)
            lin_bwd(dZY, ws["EFlast"], "last_to_zy_fc1", dEF)
            lin_bwd(dLV[3], ws["EFlast"], "last_to_logvarzy_fc1", dEF, accumulate=True)
            dhEF = buf("dhEF", B, hef)
            lin_bwd(dEF, ws["hsEF"][TB:], "ef_encoder.fc1", dhEF)
        else:
            # (6') last_to_zy_fc1 over cat(h_T, mem_T)
            Wzy = P["last_to_zy_fc1.weight"]
            Gzy = G["last_to_zy_fc1.weight"]
            Hall, mems = ws["Hall"], ws["mems"]
            self._wgrad_gemm( dZY, Hall[TB:], Gzy[:, :H], accumulate=True)
            self._wgrad_gemm( dZY, mems[TB:], Gzy[:, H:], accumulate=True)
            bgrad(dZY, "last_to_zy_fc1.bias")
            dHlast = buf("dHlast", B, H)
            dmemT = buf("dmemT", B, mem)
            ops.gemm("nn", dZY, Wzy[:, :H], dHlast)
            ops.gemm("nn", dZY, Wzy[:, H:], dmemT)
            if self.kl:                                               # the log-variance head of z_y reads the same cat(h_T, mem_T)
                Wlv, Glv = P["last_to_logvarzy_fc1.weight"], G["last_to_logvarzy_fc1.weight"]
                self._wgrad_gemm(dLV[3], Hall[TB:], Glv[:, :H], accumulate=True)
                self._wgrad_gemm(dLV[3], mems[TB:], Glv[:, H:], accumulate=True)
                bgrad(dLV[3], "last_to_logvarzy_fc1.bias")
                ops.gemm("nn", dLV[3], Wlv[:, :H], dHlast, accumulate=True)
                ops.gemm("nn", dLV[3], Wlv[:, H:], dmemT, accumulate=True)

        enc_cells = []
        dhE = [buf("dhE%d" % m, B, dm.z[m]) for m in range(3)]

        def enc_head_bwd(m):
            def run():
                tag = TAGS[m]
                dzl = dZ[m]
                if self.kl:                                       # z = last_to_z_fc1(zlast), logvar = last_to_logvarz_fc1(zlast)
                    dzl = buf("dZlast%d" % m, B, dm.z[m])
                    lin_bwd(dZ[m], ws["Zlast%d" % m], "last_to_z%s_fc1" % tag, dzl)
                    lin_bwd(dLV[m], ws["Zlast%d" % m], "last_to_logvarz%s_fc1" % tag, dzl, accumulate=True)
                lin_bwd(dzl, ws["hsE%d" % m][TB:], "encoder_%s.fc1" % tag, dhE[m])
            return run
        self._par([enc_head_bwd(m) for m in range(3)])
        for m, tag in enumerate(TAGS):                       # (3') encoder heads
            dhl = dhE[m]
            enc_cells.append(dict(T=T, B=B, h=dm.z[m], gates=ws["gatesE%d" % m], cs=ws["csE%d" % m],
                                  W=P["encoder_%s.lstm.weight_hh" % tag], dh_all=None, dh_last=dhl, dc_ext=None,
                                  dG=buf("dGE%d" % m, TB, 4 * dm.z[m]), dc_scratch=buf("dcSE%d" % m, B, dm.z[m])))
        # The encoder cells hang off the latents only: their recurrence and weight gradients do not wait for the attention /
        # memory chain.  They leave the main stream here (one launch for the three cells, then their weight gradients) and
        # run under the MFN backward; only the three MFN cells remain in the last recurrence launch of the step.
        if self.device.type == "cuda" and self.use_side_stream and not self.split_last_recurrence:
            if self._enc_stream is None:
                self._enc_stream = torch.cuda.Stream(device=self.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._enc_stream.wait_event(ev)
            with torch.cuda.stream(self._enc_stream):
                ops.lstm_bwd(enc_cells)
                for m, c in enumerate(enc_cells):
                    nm = "encoder_%s.lstm" % TAGS[m]
                    ops.gemm_tn_pair(c["dG"], self.xs[m], G[nm + ".weight_ih"], G[nm + ".bias_ih"], ws["hsE%d" % m][:TB],
                                     G[nm + ".weight_hh"])
                    ops.copy2d(G[nm + ".bias_ih"].view(1, -1), G[nm + ".bias_hh"].view(1, -1), accumulate=True)
            self._enc_used = True
            enc_cells = []
        if self.ef:                                          # the early-fusion cell is the last recurrence of the step
            hef = sum(dm.z)
            cell = dict(T=T, B=B, h=hef, gates=ws["gatesEF"], cs=ws["csEF"], W=P["ef_encoder.lstm.weight_hh"], dh_all=None,
                        dh_last=dhEF, dc_ext=None, dG=buf("dGEF", TB, 4 * hef), dc_scratch=buf("dcSEF", B, hef))
            ops.lstm_bwd(enc_cells + [cell])
            X2 = self.x_in.view(TB, dm.D)
            jobs = [(c, "encoder_%s.lstm" % TAGS[m], self.xs[m], ws["hsE%d" % m][:TB]) for m, c in enumerate(enc_cells)]
            jobs.append((cell, "ef_encoder.lstm", X2, ws["hsEF"][:TB]))
            for i, (c, nm, xin, hs) in enumerate(jobs):
                self._wgrad_pair(c["dG"], xin, G[nm + ".weight_ih"], G[nm + ".bias_ih"], hs, G[nm + ".weight_hh"],
                                 G[nm + ".bias_hh"], index=i)
        else:
            self._backward_mfn(P, G, dHlast, dmemT, enc_cells, wgrad, bgrad, lin_bwd, relu_scale)
        self.mark("bwd:lstm enc+mfn")
        self._join_side()
        self.mark("bwd:join wgrads")

    def _build_wcat(self, P):
        """The row-concatenated weights of the three consumers of `attended` (gamma1_fc1, gamma2_fc1, att2_fc1): the data
        gradient of `attended` is then one GEMM.  Depends on the parameters only."""
        dm, ops, pre = self.dm, self.ops, self.pre
        Wcat = self.buf("WcatAtt", dm.g1 + dm.g2 + dm.a2, 2 * dm.H)
        ops.copy2d(P[pre + "gamma1_fc1.weight"][:, :2 * dm.H], Wcat[:dm.g1])
        ops.copy2d(P[pre + "gamma2_fc1.weight"][:, :2 * dm.H], Wcat[dm.g1:dm.g1 + dm.g2])
        ops.copy2d(P[pre + "att2_fc1.weight"], Wcat[dm.g1 + dm.g2:])

    def _backward_mfn(self, P, G, dHlast, dmemT, enc_cells, wgrad, bgrad, lin_bwd, relu_scale):
        """Adjoint of steps (5),(4),(2),(1): memory recurrence, attention MLPs, then the MFN cells together
        with any encoder cells handed in (one launch), then all input-side weight gradients."""
        dm, ops, buf, ws = self.dm, self.ops, self.buf, self.ws
        T, B, H, mem = dm.T, dm.B, dm.H, dm.mem
        TB = T * B
        Hall, Call, mems = ws["Hall"], self.ws_views["Call"], ws["mems"]

        # (5') memory recurrence, reversed
        pre = self.pre
        Wg1, Wg2 = P[pre + "gamma1_fc1.weight"], P[pre + "gamma2_fc1.weight"]
        # dU1 | dU2 | dH2 are column blocks of one matrix: the data gradient of `attended` (three products in the
        # reference's autograd) becomes ONE GEMM against the row-concatenated weights, not three read-modify-write passes
        dUcat = buf("dUcat", TB, dm.g1 + dm.g2 + dm.a2)
        dU1, dU2, dH2 = dUcat[:, :dm.g1], dUcat[:, dm.g1:dm.g1 + dm.g2], dUcat[:, dm.g1 + dm.g2:]
        dP1, dP2 = buf("dP1", TB, mem), buf("dP2", TB, mem)
        dPc = buf("dPc", TB, mem)
        ops.mfn_mem_bwd(dict(
            T=T, B=B, mem=mem, g1=dm.g1, g2=dm.g2, cHat=ws["cHat"], mems=mems, U1=ws["U1"], U2=ws["U2"],
            Gam1=ws["Gam1"], Gam2=ws["Gam2"], W1m=Wg1[:, 2 * H:], W2m=Wg2[:, 2 * H:],
            W12=P[pre + "gamma1_fc2.weight"], W22=P[pre + "gamma2_fc2.weight"],
            scale1=relu_scale(dm.p_g1), scale2=relu_scale(dm.p_g2),
            dmem_last=dmemT, dU1=dU1, dU2=dU2, dP1=dP1, dP2=dP2, dPc=dPc))
        self.mark("bwd:mem recurrence")
        self._wgrad_gemm( dP1, ws["U1"], G[pre + "gamma1_fc2.weight"], accumulate=True, colsum_out=G[pre + "gamma1_fc2.bias"])
        self._wgrad_gemm( dP2, ws["U2"], G[pre + "gamma2_fc2.weight"], accumulate=True, colsum_out=G[pre + "gamma2_fc2.bias"])
        Attended, cStar, Att = ws["Attended"], self.ws_views["cStar"], ws["Att"]
        dAtt = buf("dAttended", TB, 2 * H)
        for (dU, nm) in ((dU1, "gamma1_fc1"), (dU2, "gamma2_fc1")):     # the two column blocks of gamma*_fc1 share dU
            Gw = G[pre + nm + ".weight"]
            self._wgrad_pair(dU, Attended, Gw[:, :2 * H], G[pre + nm + ".bias"], mems[:TB], Gw[:, 2 * H:])

        # (4') attention MLPs, time-parallel
        lin_bwd(dPc, ws["H2"], pre + "att2_fc2", dH2, mask=ws["H2"], mask_scale=relu_scale(dm.p_att2))
        lin_bwd(dH2, Attended, pre + "att2_fc1")
        Wcat = buf("WcatAtt", dm.g1 + dm.g2 + dm.a2, 2 * H)
        if not self._wcat_ready:
            self._build_wcat(P)
        self._wcat_ready = False
        ops.gemm("nn", dUcat, Wcat, dAtt)
        dL = buf("dL", TB, 2 * H)
        dcStar = buf("dcStar", TB, 2 * H)
        self.mark("bwd:att2+dAtt")
        ops.softmax_gate_bwd(dAtt, Att, cStar, dL, dcStar)
        dH1 = buf("dH1", TB, dm.a1)
        lin_bwd(dL, ws["H1"], pre + "att1_fc2", dH1, mask=ws["H1"], mask_scale=relu_scale(dm.p_att1))
        lin_bwd(dH1, cStar, pre + "att1_fc1", dcStar, accumulate=True)
        # c_t enters cStar twice: as "new" at step t and as "prev" at step t+1.  The recurrence kernel adds the two halves
        # itself (dc_ext + dc_ext2: two column blocks of the 2H-wide gradient, L2-prefetched a step ahead by its idle issuer
        # warps); the separate gather pass it replaces (two strided copies, 165 MB) cost 1.7 % of the step.  (With round 1's
        # recurrence kernel, which had no prefetch, the gather pass was the faster of the two.)
        gather = not self.fused_dcext
        if gather:
            dCext = buf("dCext", TB, H)                      # row block t = grad wrt c of cell step t
            ops.copy2d(dcStar[:, H:], dCext)
            if T > 1:
                ops.copy2d(dcStar[B:, :H], dCext[:TB - B], accumulate=True)
            dc1, dc2 = dCext, None
        else:                                                # the recurrence adds the two halves itself (dc_ext + dc_ext2)
            dc1, dc2 = dcStar[:, H:], (dcStar[B:, :H] if T > 1 else None)

        self.mark("bwd:att1+dCext")
        # (2') the recurrences reversed and (1') the weight gradients of the 6 input-side cells, all T at once
        #      (dW_ih = dG^T x and dW_hh = dG^T h_prev share dG; b_ih and b_hh share its column sums).
        #      One launch for all cells (optionally two, heavy cells first -- see split_last_recurrence).
        todo = []                                            # (cell, weight-gradient job)
        for m, tag in enumerate(TAGS):
            o = dm.hoff[m]
            cell = dict(T=T, B=B, h=dm.hm[m], gates=ws["gatesN%d" % m], cs=Call[:, o:o + dm.hm[m]],
                        W=P[pre + "lstm_%s.weight_hh" % tag], dh_all=None,
                        dh_last=dHlast[:, o:o + dm.hm[m]], dc_ext=dc1[:, o:o + dm.hm[m]],
                        dc_ext2=(None if dc2 is None else dc2[:, o:o + dm.hm[m]]),
                        dG=buf("dGN%d" % m, TB, 4 * dm.hm[m]), dc_scratch=buf("dcSN%d" % m, B, dm.hm[m]))
            todo.append((cell, (pre + "lstm_%s" % tag, "dGN%d" % m, m, Hall[:TB, o:o + dm.hm[m]])))
        for m, c in enumerate(enc_cells):
            todo.append((c, ("encoder_%s.lstm" % TAGS[m], "dGE%d" % m, m, ws["hsE%d" % m][:TB])))
        todo.sort(key=lambda cj: -cj[0]["h"])
        if self.time_split and T >= 4 and not self.split_last_recurrence:
            # The step ends when the weight gradients of these cells are done, and they cannot start before dG exists.  With
            # this switch the recurrence is launched in two halves in TIME -- steps [t0, T), then [0, t0), the carried
            # (dh, dc) handed from one to the other (mfm_lstm_cell.dh_out / dc_out / dc_last) -- and the gradient GEMMs over
            # the late half's rows run under the early half.  Measured (same box, A/B): 2.273 -> 2.298 ms, i.e. SLOWER -- the
            # gradient GEMMs' CTAs hold the SMs the second recurrence launch needs (one 200 KB CTA per SM), and each launch
            # pays its own weight-image prologue.  Kept as an experiment switch; the kernels' carry path is tested.
            t0 = T // 2
            carry = [(buf("dhCarry_" + job[1], B, c["h"]), buf("dcCarry_" + job[1], B, c["h"])) for c, job in todo]

            def sub(c, lo, hi, i):
                d = dict(c)
                d.update(T=hi - lo, gates=c["gates"][lo * B:hi * B], cs=c["cs"][lo * B:(hi + 1) * B], dG=c["dG"][lo * B:hi * B])
                for k in ("dh_all", "dc_ext"):
                    if c.get(k) is not None:
                        d[k] = c[k][lo * B:hi * B]
                if c.get("dc_ext2") is not None:             # block t applies at step t < T-1
                    d["dc_ext2"] = c["dc_ext2"][lo * B:min(hi, T - 1) * B]
                    d["dc_ext2_full"] = hi < T
                if hi == T:
                    d.update(dh_out=carry[i][0], dc_out=carry[i][1])
                else:
                    d.update(dh_last=carry[i][0], dc_last=carry[i][1])
                return d
            for lo, hi in ((t0, T), (0, t0)):
                ops.lstm_bwd([sub(c, lo, hi, i) for i, (c, _) in enumerate(todo)])
                for i, (_, (nm, dGn, m, hs)) in enumerate(todo):      # both halves of a cell on ONE side stream, in order
                    rows = slice(lo * B, hi * B)
                    self._wgrad_pair(ws[dGn][rows], self.xs[m][rows], G[nm + ".weight_ih"], G[nm + ".bias_ih"], hs[rows],
                                     G[nm + ".weight_hh"], G[nm + ".bias_hh"] if lo == 0 else None, index=i)
            return
        half = (len(todo) + 1) // 2 if (self.split_last_recurrence and len(todo) > 3) else len(todo)
        for part in (todo[:half], todo[half:]):
            if not part:
                continue
            ops.lstm_bwd([c for c, _ in part])
            # the step ends when these finish: one side stream each (the pointer hash could put two on one stream)
            for i, (_, (nm, dGn, m, hs)) in enumerate(part):
                self._wgrad_pair(ws[dGn], self.xs[m], G[nm + ".weight_ih"], G[nm + ".bias_ih"], hs, G[nm + ".weight_hh"],
                                 G[nm + ".bias_hh"], index=i)
