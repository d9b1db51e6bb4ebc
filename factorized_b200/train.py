"""The fused MFM training step and the ``train_mfm`` entry point.

``MFMTrainer.step`` is the inner loop body of the reference's ``train_mfm``
(mfm_mosi.py:427-442: zero_grad, forward, L1/CE + sum(lambda*MSE) + lambda*MMD,
backward, Adam) as ONE fixed kernel schedule: forward with the loss heads fused,
the hand-derived backward into a flat gradient buffer, an optional NCCL
all-reduce of that buffer when the batch is sharded over ranks, and a fused
flat-buffer Adam.  The schedule is captured into a CUDA graph on first use and
replayed afterwards.

``train_mfm`` keeps the reference signature
``(X_train, y_train, X_valid, y_valid, X_test, y_test, configs)`` and epoch
logic (shuffle once, time-major swap, Adam default lr, ReduceLROnPlateau on the
validation L1, save-best / reload, score) -- mfm_mosi.py:386-503.
"""
from __future__ import annotations

import os
import random
import tempfile
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from .mfm_model import MFM, MFM_KL, UNUSED, _ops

SITE_NOISE = 20   # RNG sites 20..23: the four MMD Gaussian samples


class MFMTrainer:
    """Owns flat parameter / gradient / Adam buffers for one MFM module and runs fused steps.

    The module's nn.Parameters are re-pointed at views of the flat parameter buffer, so
    ``state_dict`` / ``torch.save(model)`` keep working and see the trained weights.
    """

    def __init__(self, model: MFM, T: int, B: int, head: str = "l1", lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, use_graph: bool = True, process_group=None, seed: int = 123, _test_ops=None,
                 distributed: bool = True):
        dev = next(model.parameters()).device
        if dev.type != "cuda" and _test_ops is None:
            raise RuntimeError("MFMTrainer: model must be on a CUDA device (model.to('cuda')); no CPU path exists")
        self.model, self.dev = model, dev
        # _test_ops: tests inject a statement of the primitive set to check the multi-rank host logic without a GPU
        self.ops = _test_ops if _test_ops is not None else _ops()
        if _test_ops is not None:
            use_graph = False
        self.T, self.B = int(T), int(B)
        self.betas, self.eps = betas, eps
        self.pg = process_group
        self.world = 1
        # distributed=False: a single-rank trainer inside an initialised process group (bench.py's parity check runs on rank 0
        # only -- a collective there would wait for ranks that never call it)
        if distributed and (process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
        pd = dict(model.named_parameters())
        self.names = [k for k in pd if k not in UNUSED]
        # The encoder cell and the MFN cell of a modality read the same input: with their input weights (and bias vectors) ADJACENT
        # in the flat buffer the two input projections are one GEMM over a [4(z+h), d] view (engine.forward step 1).  Order only;
        # a tensor whose size is not a multiple of the alignment leaves a gap and the engine then keeps two GEMMs.
        if getattr(model, "_variant", "mfm") in ("mfm", "kl") and "mfn_encoder.lstm_l.weight_ih" in pd:
            moved = []
            for leaf in ("weight_ih", "bias_ih", "bias_hh"):
                for tag in "lav":
                    moved += ["encoder_%s.lstm.%s" % (tag, leaf), "mfn_encoder.lstm_%s.%s" % (tag, leaf)]
            self.names = moved + [k for k in self.names if k not in moved]
        # every tensor starts 256 B aligned inside the flat buffers (the cp.async-staged GEMM needs 16 B-aligned
        # operands); the gaps stay zero in all four buffers, so Adam and the all-reduce leave them zero
        ALIGN = 64
        offs, n = {}, 0
        for k in self.names:
            offs[k] = n
            n += (pd[k].numel() + ALIGN - 1) // ALIGN * ALIGN
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.P: Dict[str, torch.Tensor] = OrderedDict()
        self.G: Dict[str, torch.Tensor] = OrderedDict()
        with torch.no_grad():
            for k in self.names:
                p, o = pd[k], offs[k]
                view = self.flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self.P[k] = view
                self.G[k] = self.flat_g[o:o + p.numel()].view(p.shape)
        self.variant = getattr(model, "_variant", "mfm")
        from .ablations import make_engine
        self.eng = make_engine(model._cfg, T, B, dev, self.ops, head=head, variant=self.variant)
        self.eng.defer_mmd_join = True
        self.eng.fuse_mse = True
        dm = self.eng.dm
        self.x = torch.zeros(T, B, dm.D, dtype=torch.float32, device=dev)
        if head == "ce":
            self.y = torch.zeros(B, dtype=torch.int64, device=dev)
        else:
            self.y = torch.zeros(B * dm.out, dtype=torch.float32, device=dev)
        self.noise = [torch.zeros(B, k, dtype=torch.float32, device=dev) for k in (dm.z[0], dm.z[1], dm.z[2], dm.zy)]
        # every rank draws its own dropout masks and MMD Gaussian samples (independent shards, like independent
        # reference processes): the rank is folded into the stream seed here, not left to the caller
        rank = torch.distributed.get_rank(process_group) if self.world > 1 else 0
        self.rng = torch.tensor([(int(seed) + rank * 0x9E3779B1) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
        self.adam_state = torch.tensor([lr, 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        self.use_graph = use_graph
        self._copy_stream = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self.steps_done = 0

    # ------------------------------------------------------------------
    def set_lr(self, lr: float):
        self.adam_state[0:1].fill_(float(lr))

    def _schedule(self):
        """One training step on the static buffers (x, y)."""
        self._schedule_compute()
        self._schedule_update()

    def _schedule_compute(self):
        """forward + losses + backward into the flat gradient buffer (no communication)."""
        ops, eng = self.ops, self.eng
        ops.rng_tick(self.rng)
        if self.variant not in ("kl", "kl_ef"):
            for k in getattr(eng, "mmd_slots", range(4)):     # loss_MMD's Gaussian samples (mfm_model.py:26)
                ops.randn(self.noise[k], self.rng, SITE_NOISE + k)
        eng.forward(self.P, self.x, self.noise, train=True, rng=self.rng)
        dX, dY = eng.losses(self.y)
        ops.zero(self.flat_g)
        eng.backward(self.P, self.G, dX, dY, eng.dm.lda_mmd)

    def _schedule_update(self):
        """[all-reduce of the flat gradient buffer over NCCL/NVLink] + fused Adam (1/world folded in)."""
        ops = self.ops
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.pg)
        ops.adam(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.adam_state, grad_scale=1.0 / self.world,
                 betas=self.betas, eps=self.eps)

    def _run(self):
        if not self.use_graph:
            n0 = self.ops.launches
            self._schedule()
            self.launches_per_step = self.ops.launches - n0
            return
        if self.graph is None:
            # warm up on a side stream (lazy module loads, cudaFuncSetAttribute, NCCL init), then capture
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            snap = [t.clone() for t in (self.flat_p, self.flat_m, self.flat_v, self.adam_state, self.rng)]
            with torch.cuda.stream(s):
                self._schedule()
            torch.cuda.current_stream(self.dev).wait_stream(s)
            torch.cuda.synchronize(self.dev)
            for t, c in zip((self.flat_p, self.flat_m, self.flat_v, self.adam_state, self.rng), snap):
                t.copy_(c)
            # single rank: the whole step is one graph.  Multi-rank: the compute part is a graph, the NCCL
            # all-reduce + Adam (3 launches) stay eager on the same stream -- the collective is not captured.
            g = torch.cuda.CUDAGraph()
            n0 = self.ops.launches
            # the capture stream (and the engine's branch pool) are HIGH priority, the weight-gradient and MMD streams normal:
            # kernel nodes inherit it, so when SMs free up the critical chain's CTAs are placed before the background work's
            with torch.cuda.graph(g, stream=torch.cuda.Stream(device=self.dev, priority=-1)):
                if self.world > 1:
                    self._schedule_compute()
                else:
                    self._schedule()
            self.launches_per_step = self.ops.launches - n0 + (2 if self.world > 1 else 0)
            self.graph = g
        self.graph.replay()
        if self.world > 1:
            self._schedule_update()

    def step_device(self):
        """Run one step on whatever is in self.x / self.y (already on the device)."""
        self._run()
        self.steps_done += 1

    def step(self, x, y):
        """x: [T,B,D] fp32, y: targets; host (ideally pinned) or device tensors.  Returns the device loss buffer
        (index 0 = discriminative loss, 1..3 = MSE l/a/v, 4..7 = MMD parts, 8 = total); reading it syncs.

        Host batches travel on a dedicated copy stream into one of two device staging buffers, so the H2D copy of
        batch n+1 overlaps the compute of batch n (the reference does a blocking pageable copy per step,
        mfm_mosi.py:428-429).  The call itself never synchronises."""
        if x.is_cuda or self.dev.type != "cuda":
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y.reshape(self.y.shape), non_blocking=True)
            self.step_device()
            return self.eng.loss_buf
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._stage = [(torch.empty_like(self.x), torch.empty_like(self.y)) for _ in range(2)]
            self._consumed = [None, None]
            self._stage_idx = 0
        i = self._stage_idx
        self._stage_idx ^= 1
        cs, main = self._copy_stream, torch.cuda.current_stream(self.dev)
        sx, sy = self._stage[i]
        if self._consumed[i] is not None:
            cs.wait_event(self._consumed[i])                 # staging buffer i was read by step n-2
        with torch.cuda.stream(cs):
            sx.copy_(x, non_blocking=True)
            sy.copy_(y.reshape(sy.shape), non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(cs)
        main.wait_event(landed)
        self.x.copy_(sx)
        self.y.copy_(sy)
        done = torch.cuda.Event()
        done.record(main)
        self._consumed[i] = done
        self.step_device()
        return self.eng.loss_buf


def _to_time_major(X):
    return np.ascontiguousarray(np.swapaxes(np.asarray(X, dtype=np.float32), 0, 1))


def score(predictions, y_test, head: str = "l1"):
    """The reference's score block (mfm_mosi.py:483-498) as a dict instead of prints."""
    scores = {}
    y_hat, y_test = np.asarray(predictions), np.asarray(y_test)
    if head == "l1" and y_hat.ndim == 1:
        scores["mae"] = float(np.mean(np.absolute(y_hat - y_test)))                            # :484
        scores["corr"] = float(np.corrcoef(y_hat, y_test)[0][1])                               # :486
        scores["mult_acc"] = round(float(np.mean(np.round(y_hat) == np.round(y_test))), 5)     # :488
        true_label, predicted_label = (y_test >= 0), (y_hat >= 0)                              # :492-493
        scores["binary_acc"] = float(np.mean(predicted_label == true_label))                   # :498 accuracy_score
        try:                                                                                   # :490,495-497 (sklearn, as the reference)
            from sklearn.metrics import classification_report, confusion_matrix, f1_score
            scores["mult_f_score"] = round(float(f1_score(np.round(y_hat), np.round(y_test), average="weighted")), 5)
            scores["confusion_matrix"] = confusion_matrix(true_label, predicted_label).tolist()
            scores["classification_report"] = classification_report(true_label, predicted_label, digits=5, zero_division=0)
        except ImportError:
            pass
    elif head == "ce":                                   # mfm_moud.py:421-428 / mfm_you.py:399-406: argmax, confusion matrix, report
        predicted_label = np.argmax(y_hat, axis=1)
        scores["acc"] = float(np.mean(predicted_label == y_test))
        try:
            from sklearn.metrics import classification_report, confusion_matrix
            scores["confusion_matrix"] = confusion_matrix(y_test, predicted_label).tolist()
            scores["classification_report"] = classification_report(y_test, predicted_label, digits=5, zero_division=0)
        except ImportError:
            pass
    return scores


def train_mfm_test_zeros(X_train, y_train, X_valid, y_valid, X_test, y_test, configs, verbose: bool = True,
                         save_dir: Optional[str] = None):
    """Drop-in for the reference's train_mfm_test_zeros (mfm_mosi.py:505-638): train MFM exactly as train_mfm does, then predict
    the test set three times with one modality zeroed out -- language, acoustic, visual (:571-578) -- and score each (:632-638).
    Returns train_mfm's dict plus ``predictions_nol/noa/nov``, ``scores_nol/noa/nov`` and the reconstruction errors of the zeroed
    modality the reference prints (:592-595)."""
    configs = [dict(c) for c in configs]
    configs[0]["type"] = "mfm"                                        # (:516 builds MFM whatever the type)
    out = train_mfm(X_train, y_train, X_valid, y_valid, X_test, y_test, configs, head="l1", verbose=verbose, save_dir=save_dir)
    model = out["model"].eval()
    d_l, d_a, d_v = configs[0]["input_dims"]
    dev = next(model.parameters()).device
    X = torch.from_numpy(_to_time_major(X_test)).to(dev)
    spans = dict(nol=(0, d_l), noa=(d_l, d_l + d_a), nov=(d_l + d_a, d_l + d_a + d_v))
    with torch.no_grad():
        for i, (tag, (lo, hi)) in enumerate(spans.items()):
            Xz = X.clone()
            Xz[:, :, lo:hi] = 0.0
            decoded, _, _ = model.forward(Xz)
            y_hat = decoded[3].squeeze(1).cpu().numpy()
            out["predictions_" + tag] = y_hat
            out["scores_" + tag] = score(y_hat, y_test)
            out["recon_" + tag] = float(torch.nn.functional.mse_loss(decoded[i], X[:, :, lo:hi]))
            if verbose:
                print("scoring y_hat_" + tag, out["scores_" + tag])
    return out


def train_mfm_missing(X_train, y_train, X_valid, y_valid, X_test, y_test, configs, verbose: bool = True,
                      save_dir: Optional[str] = None):
    """Drop-in for the reference's train_mfm_missing (mfm_mosi.py:918-1105): MFM_missing trained with the four-pass loss
    (:962-982, fused step of ``missing.MissingEngine``); the epoch's training figure is the mean all-present text reconstruction
    MSE (:984-985); validation is the FULL loss of the whole validation set in eval mode, MMD included (:987-1021), which drives
    ReduceLROnPlateau and save-best; the test set is predicted with all modalities and with each one inferred (:1023-1060) and
    every prediction scored (:1097-1105).  Returns a dict with the model, the four prediction vectors and their scores, the
    twelve reconstruction errors the reference prints, and the history."""
    from .missing import MFM_missing, PASSES
    config = configs[0]
    p = np.random.permutation(X_train.shape[0])                     # :919-921
    X_train, y_train = np.asarray(X_train)[p], np.asarray(y_train)[p]
    Xt, Xv, Xte = _to_time_major(X_train), _to_time_major(X_valid), _to_time_major(X_test)
    dev = torch.device("cuda")
    model = MFM_missing(*configs).to(dev)                            # :929
    model.mmd_noise = "cuda"
    d_l, d_a, d_v = config["input_dims"]
    T, total_n = Xt.shape[0], Xt.shape[1]
    bs = int(config["batchsize"])
    num_batches = total_n // bs                                      # :951
    lr = 1e-3                                                        # optim.Adam(model.parameters()), :931
    trainer = MFMTrainer(model, T, bs, head="l1", lr=lr)
    sched_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(sched_opt, "min")   # :945
    ytr = np.asarray(y_train, dtype=np.float32)
    Xpin = torch.empty((max(num_batches, 1), T, bs, Xt.shape[2]), dtype=torch.float32).pin_memory()
    ypin = torch.empty((max(num_batches, 1), bs), dtype=torch.float32).pin_memory()
    for b in range(num_batches):
        Xpin[b].copy_(torch.from_numpy(Xt[:, b * bs:(b + 1) * bs]))
        ypin[b].copy_(torch.from_numpy(ytr[b * bs:(b + 1) * bs]))
    Fn = torch.nn.functional

    def full_loss(out, bx, by):                                      # :962-982 / :1003-1019
        dec, nol, noa, nov, mmd, missing = out
        x_l, x_a, x_v = bx[:, :, :d_l], bx[:, :, d_l:d_l + d_a], bx[:, :, d_l + d_a:]
        gen = config["lda_xl"] * Fn.mse_loss(dec[0], x_l) + config["lda_xa"] * Fn.mse_loss(dec[1], x_a) \
            + config["lda_xv"] * Fn.mse_loss(dec[2], x_v) + config["lda_xl"] * Fn.mse_loss(nol[0], x_l) \
            + config["lda_xa"] * Fn.mse_loss(noa[1], x_a) + config["lda_xv"] * Fn.mse_loss(noa[2], x_v)
        disc = sum(Fn.l1_loss(d[3].squeeze(1), by) for d in (dec, nol, noa, nov))
        return disc + gen + config["lda_mmd"] * mmd + missing

    def evaluate(X, y):
        model.eval()
        with torch.no_grad():
            bx = torch.from_numpy(X).to(dev)
            by = torch.from_numpy(np.asarray(y, dtype=np.float32)).to(dev)
            return float(full_loss(model.forward(bx), bx, by))

    def predict(X):
        model.eval()
        with torch.no_grad():
            bx = torch.from_numpy(X).to(dev)
            out = model.forward(bx)
            xm = (bx[:, :, :d_l], bx[:, :, d_l:d_l + d_a], bx[:, :, d_l + d_a:])
            recon = {"x_%s_hat%s" % (t, s): float(Fn.mse_loss(out[p][m], xm[m]))
                     for p, s in enumerate(PASSES) for m, t in enumerate("lav")}
            return [out[p][3].squeeze(1).cpu().numpy() for p in range(4)], recon

    best_valid = 999999.0
    save_dir = save_dir or tempfile.mkdtemp(prefix="res_mfm2_")
    path = os.path.join(save_dir, "mfn_%d.pt" % random.randint(0, 100000))
    history = []
    for epoch in range(int(config["num_epochs"])):
        model.train()
        acc = torch.zeros((), dtype=torch.float32, device=dev)
        for b in range(num_batches):
            lb = trainer.step(Xpin[b], ypin[b])
            acc += lb[10]                                            # l2_loss(x_l_hat, x_l), :984
        train_loss = float(acc) / max(num_batches, 1)
        valid_loss = evaluate(Xv, y_valid)
        scheduler.step(valid_loss)
        trainer.set_lr(sched_opt.param_groups[0]["lr"])
        history.append((epoch, train_loss, valid_loss))
        if valid_loss <= best_valid:
            best_valid = valid_loss
            torch.save(model, path)
            if verbose:
                print(epoch, train_loss, valid_loss, "saving model")
        elif verbose:
            print(epoch, train_loss, valid_loss)
    if os.path.exists(path):
        model = torch.load(path, weights_only=False)
    preds, recon = predict(Xte)
    out = dict(model=model, history=history, best_valid=best_valid, checkpoint=path, recon=recon)
    for p, s in enumerate(PASSES):
        out["predictions" + s] = preds[p]
        out["scores" + s] = score(preds[p], y_test)
        if verbose:
            print("scoring y_hat" + s, out["scores" + s])
    return out


def train_mfm_ablation(X_train, y_train, X_valid, y_valid, X_test, y_test, configs, head: str = "l1", verbose: bool = True,
                       save_dir: Optional[str] = None):
    """Drop-in for the reference's train_mfm_ablation (mfm_mosi.py:640-767): ``config['type']`` in m_a / m_b / m_c / m_d
    selects M_A .. M_D (:651-658); the epoch loop, the step and the scores are train_mfm's (:660-767 repeat :403-503)."""
    from .ablations import ABLATION_MODELS
    kind = configs[0].get("type")
    if kind not in ABLATION_MODELS:
        raise ValueError("train_mfm_ablation: config['type'] must be one of %s, got %r" % (sorted(ABLATION_MODELS), kind))
    return train_mfm(X_train, y_train, X_valid, y_valid, X_test, y_test, configs, head=head, verbose=verbose,
                     save_dir=save_dir, _model_cls=ABLATION_MODELS[kind])


def train_mfm(X_train, y_train, X_valid, y_valid, X_test, y_test, configs, head: str = "l1", verbose: bool = True,
              save_dir: Optional[str] = None, _model_cls=None):
    """Drop-in for the reference's train_mfm (mfm_mosi.py:386-503); CE head: mfm_mosi_acc.py:396-503.
    X_* are numpy [n,T,D]; y_* [n] (or [n,out]).  Returns a dict with the trained model and scores."""
    config = configs[0]
    p = np.random.permutation(X_train.shape[0])                     # :387-389
    X_train, y_train = np.asarray(X_train)[p], np.asarray(y_train)[p]
    Xt, Xv, Xte = _to_time_major(X_train), _to_time_major(X_valid), _to_time_major(X_test)    # :391-393
    dev = torch.device("cuda")
    model = (_model_cls or (MFM_KL if config.get("type", "mfm") == "kl" else MFM))(*configs).to(dev)      # :398-401,414
    model.mmd_noise = "cuda"
    model.eval_skip_mmd = True            # evaluate / predict discard the MMD of the whole-set forward (:448,:460)
    T, total_n = Xt.shape[0], Xt.shape[1]
    bs = int(config["batchsize"])
    num_batches = total_n // bs                                      # :423 (py2 integer division: drops the tail)
    lr = 1e-3                                                        # optim.Adam(model.parameters()) default, :403
    trainer = MFMTrainer(model, T, bs, head=head, lr=lr)
    sched_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)   # carrier for ReduceLROnPlateau only
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(sched_opt, "min")   # :417
    ydt = np.int64 if head == "ce" else np.float32
    # Pinned staging, one CONTIGUOUS [T,bs,D] block per batch: a strided slice Xt[:, b*bs:(b+1)*bs] of the time-major array
    # is not a legal source for one async H2D copy (torch would stage it through pageable memory and block)
    ytr = np.asarray(y_train, dtype=ydt)
    Xpin = torch.empty((max(num_batches, 1), T, bs, Xt.shape[2]), dtype=torch.float32).pin_memory()
    ypin = torch.empty((max(num_batches, 1), bs) + tuple(ytr.shape[1:]), dtype=torch.from_numpy(ytr[:1]).dtype).pin_memory()
    for b in range(num_batches):
        Xpin[b].copy_(torch.from_numpy(Xt[:, b * bs:(b + 1) * bs]))
        ypin[b].copy_(torch.from_numpy(ytr[b * bs:(b + 1) * bs]))

    def evaluate(X, y):                                              # :445-455
        model.eval()
        with torch.no_grad():
            bx = torch.from_numpy(X).to(dev)
            decoded, _, _ = model.forward(bx)
            y_hat = decoded[3]
            y_hat = y_hat.squeeze(1) if y_hat.shape[1] == 1 else y_hat
            by = torch.from_numpy(np.asarray(y, dtype=ydt)).to(dev)
            if head == "ce":
                return torch.nn.functional.cross_entropy(y_hat, by).item()
            return torch.nn.functional.l1_loss(y_hat, by).item()

    def predict(X):                                                  # :457-465
        model.eval()
        with torch.no_grad():
            decoded, _, _ = model.forward(torch.from_numpy(X).to(dev))
            y_hat = decoded[3]
            return (y_hat.squeeze(1) if y_hat.shape[1] == 1 else y_hat).cpu().numpy()

    best_valid = 999999.0
    save_dir = save_dir or tempfile.mkdtemp(prefix="res_mfm2_")
    path = os.path.join(save_dir, "mfn_%d.pt" % random.randint(0, 100000))
    history = []
    for epoch in range(int(config["num_epochs"])):
        model.train()
        acc = torch.zeros((), dtype=torch.float32, device=dev)
        for b in range(num_batches):
            lb = trainer.step(Xpin[b], ypin[b])
            acc += lb[0]                                             # device-side; the reference syncs here (:442)
        train_loss = float(acc) / max(num_batches, 1)
        valid_loss = evaluate(Xv, y_valid)
        scheduler.step(valid_loss)
        trainer.set_lr(sched_opt.param_groups[0]["lr"])
        history.append((epoch, train_loss, valid_loss))
        if valid_loss <= best_valid:
            best_valid = valid_loss
            torch.save(model, path)                                  # :477 whole-module pickle
            if verbose:
                print(epoch, train_loss, valid_loss, "saving model")
        elif verbose:
            print(epoch, train_loss, valid_loss)
    if os.path.exists(path):
        model = torch.load(path, weights_only=False)                 # :481
    y_hat = predict(Xte)
    scores = score(y_hat, y_test, head)
    if verbose:
        print(scores)
    return dict(model=model, scores=scores, history=history, best_valid=best_valid, checkpoint=path, predictions=y_hat)
