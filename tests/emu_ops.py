"""Torch-CPU statement of the CUDA primitive set -- TEST INFRASTRUCTURE ONLY.

Each method states, in plain torch fp32/fp64, what the kernel of the same name
in ``factorized_b200/csrc`` must compute.  It serves two purposes:
 * on a machine without a GPU, injecting it into ``Engine`` lets the host
   schedule + hand-derived backward be checked against autograd of the oracle;
 * on the GPU box, each CUDA primitive is compared against it on random inputs.
The product never imports this file.
"""
import torch

from oracle.rng_replay import M32, _fmix32, site_seed, rng_bits, keep_mask   # noqa: F401  (one statement of the RNG)


class EmuOps:
    def __init__(self, dtype=torch.float32):
        self.dtype = dtype
        self.launches = 0

    # ---- GEMM with fused epilogue ----
    def gemm(self, mode, A, B, C, bias=None, bias2=None, act=0, accumulate=False, mask=None, mask_scale=1.0,
             drop=None, rng=None, colsum_out=None):
        self.launches += 1
        if colsum_out is not None:
            assert mode == "tn"
            colsum_out += A.sum(0)
        if mode == "nt":
            v = A @ B.t()
        elif mode == "nn":
            v = A @ B
        elif mode == "tn":
            v = A.t() @ B
        else:
            raise ValueError(mode)
        assert v.shape == C.shape, (mode, A.shape, B.shape, C.shape)
        if bias is not None:
            v = v + bias
        if bias2 is not None:
            v = v + bias2
        if act == 1:
            v = torch.relu(v)
        elif act == 2:
            v = torch.tanh(v)
        elif act == 3:
            v = torch.sigmoid(v)
        if drop is not None:
            p, site = drop
            v = v * keep_mask(rng, site, p, v.shape[0], v.shape[1]) / (1.0 - p)
        if mask is not None:
            v = v * (mask > 0).to(v.dtype) * mask_scale
        if accumulate:
            C += v
        else:
            C.copy_(v)

    def gemm_mse(self, A, W, bias, x, loss_scale, grad_scale, slot, dxhat, xhat=None):
        self.launches += 1
        v = A @ W.t() + bias
        r = v - x
        slot[0] += loss_scale * (r * r).sum()
        dxhat.copy_(grad_scale * r)
        if xhat is not None:
            xhat.copy_(v)

    def gemm_tn_pair(self, dY, A1, C1, colsum1, A2, C2):
        self.gemm("tn", dY, A1, C1, accumulate=True, colsum_out=colsum1)
        self.gemm("tn", dY, A2, C2, accumulate=True)

    # ---- LSTM recurrences ----
    def lstm_fwd(self, cells):
        self.launches += 1
        for c in cells:
            T, B, h = c["T"], c["B"], c["h"]
            hs, cs, gates, W = c["hs"], c["cs"], c["gates"], c["W"]
            hs[:B] = 0
            cs[:B] = 0
            dup = c.get("cs_dup")
            if dup is not None:
                dup[:B] = 0
            for t in range(T):
                hp, cp = hs[t * B:(t + 1) * B], cs[t * B:(t + 1) * B]
                pre = hp @ W.t()
                if t < c["gx_steps"]:
                    pre = pre + c["gx"][t * B:(t + 1) * B]
                else:
                    pre = pre + c["bias_rest"]
                i = torch.sigmoid(pre[:, :h])
                f = torch.sigmoid(pre[:, h:2 * h])
                g = torch.tanh(pre[:, 2 * h:3 * h])
                o = torch.sigmoid(pre[:, 3 * h:])
                cn = f * cp + i * g
                hn = o * torch.tanh(cn)
                gates[t * B:(t + 1) * B] = torch.cat([i, f, g, o], 1)
                hs[(t + 1) * B:(t + 2) * B] = hn
                cs[(t + 1) * B:(t + 2) * B] = cn
                if dup is not None:
                    dup[(t + 1) * B:(t + 2) * B] = cn

    def lstm_bwd(self, cells):
        self.launches += 1
        for c in cells:
            T, B, h = c["T"], c["B"], c["h"]
            gates, cs, W, dG = c["gates"], c["cs"], c["W"], c["dG"]
            dh = torch.zeros(B, h, dtype=gates.dtype)
            dc = torch.zeros(B, h, dtype=gates.dtype)
            if c.get("dc_last") is not None:                 # time-split recurrence: carried dc comes in
                dc = dc + c["dc_last"]
            for t in range(T - 1, -1, -1):
                r = slice(t * B, (t + 1) * B)
                if c["dh_all"] is not None:
                    dh = dh + c["dh_all"][r]
                if c["dh_last"] is not None and t == T - 1:
                    dh = dh + c["dh_last"]
                g4 = gates[r]
                i, f, g, o = g4[:, :h], g4[:, h:2 * h], g4[:, 2 * h:3 * h], g4[:, 3 * h:]
                cprev, cnew = cs[r], cs[(t + 1) * B:(t + 2) * B]
                tc = torch.tanh(cnew)
                dc = dc + dh * o * (1 - tc * tc)
                if c["dc_ext"] is not None:
                    dc = dc + c["dc_ext"][r]
                if c.get("dc_ext2") is not None and (t < T - 1 or c.get("dc_ext2_full")):
                    dc = dc + c["dc_ext2"][r]
                d_o = dh * tc * o * (1 - o)
                d_i = dc * g * i * (1 - i)
                d_f = dc * cprev * f * (1 - f)
                d_g = dc * i * (1 - g * g)
                dg4 = torch.cat([d_i, d_f, d_g, d_o], 1)
                dG[r] = dg4
                dh = dg4 @ W
                dc = dc * f
            if c.get("dh_out") is not None:                  # ... and goes out
                c["dh_out"].copy_(dh)
                c["dc_out"].copy_(dc)

    # ---- MFN memory recurrence ----
    def mfn_mem_fwd(self, a):
        self.launches += 1
        T, B, mem = a["T"], a["B"], a["mem"]
        mems = a["mems"]
        mems[:B] = 0
        for t in range(T):
            r = slice(t * B, (t + 1) * B)
            mp = mems[r]
            u1 = torch.relu(a["G1pre"][r] + mp @ a["W1m"].t())
            u2 = torch.relu(a["G2pre"][r] + mp @ a["W2m"].t())
            if a["drop1"] is not None:
                p, site = a["drop1"]
                u1 = u1 * keep_mask(a["rng"], site, p, B, a["g1"], row0=t * B) / (1 - p)
            if a["drop2"] is not None:
                p, site = a["drop2"]
                u2 = u2 * keep_mask(a["rng"], site, p, B, a["g2"], row0=t * B) / (1 - p)
            ga1 = torch.sigmoid(u1 @ a["W12"].t() + a["b12"])
            ga2 = torch.sigmoid(u2 @ a["W22"].t() + a["b22"])
            a["U1"][r] = u1
            a["U2"][r] = u2
            a["Gam1"][r] = ga1
            a["Gam2"][r] = ga2
            mems[(t + 1) * B:(t + 2) * B] = ga1 * mp + ga2 * a["cHat"][r]

    def mfn_mem_bwd(self, a):
        self.launches += 1
        T, B = a["T"], a["B"]
        dmem = a["dmem_last"].clone()
        for t in range(T - 1, -1, -1):
            r = slice(t * B, (t + 1) * B)
            mp = a["mems"][r]
            ga1, ga2, ch = a["Gam1"][r], a["Gam2"][r], a["cHat"][r]
            dp1 = dmem * mp * ga1 * (1 - ga1)
            dp2 = dmem * ch * ga2 * (1 - ga2)
            a["dP1"][r] = dp1
            a["dP2"][r] = dp2
            a["dPc"][r] = dmem * ga2 * (1 - ch * ch)
            du1 = (dp1 @ a["W12"]) * (a["U1"][r] > 0).to(dp1.dtype) * a["scale1"]
            du2 = (dp2 @ a["W22"]) * (a["U2"][r] > 0).to(dp1.dtype) * a["scale2"]
            a["dU1"][r] = du1
            a["dU2"][r] = du2
            dmem = dmem * ga1 + du1 @ a["W1m"] + du2 @ a["W2m"]

    # ---- attention gate ----
    def softmax_gate_fwd(self, L, cstar, attended):
        self.launches += 1
        att = torch.softmax(L, dim=1)
        L.copy_(att)
        attended.copy_(att * cstar)

    def softmax_gate_bwd(self, dAttended, att, cstar, dL, dcstar):
        self.launches += 1
        da = dAttended * cstar
        dL.copy_(att * (da - (da * att).sum(1, keepdim=True)))
        dcstar.copy_(dAttended * att)

    # ---- MMD ----
    @staticmethod
    def _k(x, y):
        dim = x.shape[1]
        d2 = ((x.unsqueeze(1) - y.unsqueeze(0)) ** 2).sum(2)
        return torch.exp(-d2 / float(dim * dim))

    def mmd_fwd(self, z, g, out):
        self.launches += 1
        out[0] = self._k(g, g).mean() + self._k(z, z).mean() - 2.0 * self._k(g, z).mean()

    def mmd_bwd(self, z, g, scale, dz, scale_dev=None):
        self.launches += 1
        if scale_dev is not None:
            scale = scale * float(scale_dev)
        n, dim = z.shape
        c = -2.0 / float(dim * dim)
        kzz = self._k(z, z)
        kgz = self._k(g, z)                                    # [i over g, j over z]
        dzz = (2.0 / (n * n)) * c * ((kzz.sum(1, keepdim=True) * z) - kzz @ z)
        dgz = (-2.0 / (n * n)) * c * ((kgz.sum(0).unsqueeze(1) * z) - kgz.t() @ g)
        dz += scale * (dzz + dgz)

    def rownorm2(self, x, out):
        self.launches += 1
        out.copy_((x * x).sum(1))

    def mmd_kexp(self, S, nx, ny, dim, weight, slot):
        self.launches += 1
        d2 = (nx.view(-1, 1) + ny.view(1, -1) - 2.0 * S).clamp_min(0.0)
        S.copy_(torch.exp(-d2 / float(dim * dim)))
        slot[0] += weight * S.sum()

    def mmd_kexp64(self, S, nx, ny, dim, weight, acc):
        self.launches += 1
        d2 = (nx.view(-1, 1) + ny.view(1, -1) - 2.0 * S).clamp_min(0.0)
        S.copy_(torch.exp(-d2 / float(dim * dim)))
        acc[0] += weight * S.double().sum()

    def mmd_fold(self, acc, slots):
        self.launches += 1
        slots.copy_(acc.to(slots.dtype))

    def mmd_combine(self, z, rs, cs, t1, t2, scale, dz, scale_dev=None):
        self.launches += 1
        if scale_dev is not None:
            scale = scale * float(scale_dev)
        n, dim = z.shape
        coef = scale * 2.0 * (-2.0 / float(dim * dim)) / float(n * n)
        dz += coef * ((rs - cs).view(-1, 1) * z - t1 + t2)

    # ---- small elementwise / reductions ----
    def copy2d(self, src, dst, accumulate=False):
        self.launches += 1
        if accumulate:
            dst += src
        else:
            dst.copy_(src)

    def add(self, a, b, out):
        self.launches += 1
        out.copy_(a + b)

    def zero(self, t):
        self.launches += 1
        t.zero_()

    def colsum(self, A, out):
        self.launches += 1
        out += A.sum(0)

    def relu_bwd(self, dy, y, out):
        self.launches += 1
        out.copy_(dy * (y > 0).to(dy.dtype))

    def mse_fwd_bwd(self, xhat, x, loss_scale, grad_scale, slot, dxhat):
        self.launches += 1
        r = xhat - x
        slot[0] += loss_scale * (r * r).sum()
        if dxhat is not None:
            dxhat.copy_(grad_scale * r)

    def l1_fwd_bwd(self, yhat, y, scale, slot, dy):
        self.launches += 1
        r = yhat - y
        slot[0] += scale * r.abs().sum()
        dy.copy_(scale * torch.sign(r))

    def ce_fwd_bwd(self, yhat, y, scale, slot, dy):
        self.launches += 1
        lp = torch.log_softmax(yhat, dim=1)
        idx = y.long().view(-1, 1)
        slot[0] += -scale * lp.gather(1, idx).sum()
        p = torch.exp(lp)
        p.scatter_add_(1, idx, -torch.ones_like(p[:, :1]))
        dy.copy_(scale * p)

    def kld_fwd(self, mu, logvar, slot):
        self.launches += 1
        slot[0] += -0.5 * (1 + logvar - mu * mu - torch.exp(logvar)).sum()

    def kld_bwd(self, mu, logvar, scale, dmu, dlogvar, scale_dev=None):
        self.launches += 1
        if scale_dev is not None:
            scale = scale * float(scale_dev)
        dmu += scale * mu
        dlogvar.copy_(scale * 0.5 * (torch.exp(logvar) - 1.0))

    def loss_total(self, lb, l0, l1, l2, lmmd):
        self.launches += 1
        lb[8] = lb[0] + l0 * lb[1] + l1 * lb[2] + l2 * lb[3] + lmmd * (lb[4] + lb[5] + lb[6] + lb[7])

    def adam(self, p, g, m, v, state, grad_scale=1.0, betas=(0.9, 0.999), eps=1e-8):
        """state: float tensor [lr, step, scratch, scratch]; step is incremented here."""
        self.launches += 1
        state[1] += 1
        t = float(state[1])
        lr = float(state[0])
        b1, b2 = betas
        gg = g * grad_scale
        m.mul_(b1).add_(gg, alpha=1 - b1)
        v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
        denom = v.sqrt() / (1 - b2 ** t) ** 0.5 + eps
        p -= (lr / (1 - b1 ** t)) * (m / denom)

    def rng_tick(self, rng):
        self.launches += 1
        rng[1] += 1

    def randn(self, out, rng, site):
        self.launches += 1
        n = out.numel()
        i = torch.arange(n, dtype=torch.int64)
        ss = site_seed(rng, site)
        h1 = rng_bits(ss, 2 * i)
        h2 = rng_bits(ss, 2 * i + 1)
        u1 = ((h1 >> 8).to(torch.float32) + 1.0) * (1.0 / 16777216.0)
        u2 = (h2 >> 8).to(torch.float32) * (1.0 / 16777216.0)
        out.view(-1).copy_(torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * torch.pi * u2))
