#!/usr/bin/env python
"""bench.py -- MFM train samples/sec on synthetic CMU-MOSI shapes (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config mosi|mosei|iemocap|pom] [--batch B] [--scaling weak|strong]
                  [--impl ours|reference]

One "step" = one MFM training step (forward, L1 + sum(lambda*MSE) + lambda*MMD, backward, [all-reduce], Adam) on one
synthetic batch [T=20, B, D=325] per GPU, best_acc hyper-parameters (mfm_mosi.py:1239-1286), dropout ACTIVE
(model.train(), like the reference's loop).  Weak scaling: B per GPU is fixed, value = N*B*K / max-over-ranks time.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)



def gate_train_flops_per_sample(configs, T):
    """3 * sum over the 9 LSTM cells of T*2*4h*(d_in+h), counted on the reference's two-GEMM formulation."""
    c = configs[0]
    d, hm = c["input_dims"], c["h_dims"]
    z = [c["zl_size"], c["za_size"], c["zv_size"]]
    hd = [c["fy_size"] + f for f in (c["fl_size"], c["fa_size"], c["fv_size"])]
    f = 0
    for m in range(3):
        f += T * 8 * z[m] * (d[m] + z[m]) + T * 8 * hm[m] * (d[m] + hm[m]) + T * 8 * hd[m] * (2 * hd[m])
    return 3.0 * f


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], bf16=j["bf16_tflops"], bf16_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


def cpu_reference_leg(configs, T, B, steps, warmup, head="l1"):
    """The reference's own CPU path for this workload, timed on this box's host cores: the oracle port of
    MFM + the 25-line train step (forward, loss, backward, Adam) in torch fp32 on all host threads.
    (/root/reference cannot travel to the GPU box; the oracle is pinned to it by tests/golden.)"""
    import torch
    from oracle import mfm_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    P = O.init_params(configs, 123)
    x, y = O.synthetic_batch(configs, T, B, 1234, head)
    state = {}
    times = []
    for i in range(warmup + steps):
        noise = O.draw_mmd_noise(configs, B, 999 + i)
        t0 = time.perf_counter()
        P, losses, _, _ = O.train_step(P, x, y, configs, noise, state, head=head, train=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return dict(value=B / sec, unit="samples/s", cores=torch.get_num_threads(), kind="port",
                sample="%d warm-up + %d timed steps of batch %d (T=%d), oracle port of mfm_mosi.py:427-441, torch fp32, "
                       "%.2f s/step" % (warmup, steps, B, T, sec)), sec


def profile_primitives(trainer, reps=3):
    # aggregates over `reps` eager steps; callers divide by reps
    """Per-primitive device time inside one eager step (CUDA events on the launching stream), used to name the
    dominant kernel and its achieved FLOP rate.  Not part of the timed region."""
    import torch
    ops = trainer.ops
    rec = []
    shapes = []

    def wrap(name, fn):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            fl = 0.0
            by = 0.0
            tag = name
            if name == "gemm":
                C = a[3]
                Kd = a[1].shape[1] if a[0] in ("nt", "nn") else a[1].shape[0]
                fl = 2.0 * C.shape[0] * C.shape[1] * Kd
                # algorithmic bytes of one launch: read A and B once, write C once (read it too when accumulating into a
                # full-size C; a split-K weight gradient's C is tiny either way)
                by = 4.0 * (a[1].numel() + a[2].numel() + C.numel() * (2 if k.get("accumulate") else 1))
                tag = "gemm_" + a[0]
                streamed = C.shape[0] * C.shape[1] * Kd >= (1 << 20)   # the tcgen05 path (mfm_set_gemm_tc_min_work default)
                shapes.append(("%s %dx%dx%d%s" % (a[0], C.shape[0], C.shape[1], Kd, " acc" if k.get("accumulate") else ""),
                               e0, e1, by, streamed))
            elif name == "gemm_tn_pair":                      # C1 += dY^T A1 (+ column sums), C2 += dY^T A2: dY read once
                dY, A1, C1, _cs, A2, C2 = a[:6]
                fl = 2.0 * dY.shape[1] * (A1.shape[1] + A2.shape[1]) * dY.shape[0]
                by = 4.0 * (dY.numel() + A1.numel() + A2.numel() + 2 * (C1.numel() + C2.numel()))
                tag = "gemm_tn"
                shapes.append(("tn-pair %dx(%d+%d)x%d acc" % (dY.shape[1], A1.shape[1], A2.shape[1], dY.shape[0]), e0, e1, by,
                               fl / 2 >= (1 << 20)))
            elif name in ("lstm_fwd", "lstm_bwd"):
                fl = sum(2.0 * c["T"] * c["B"] * 4 * c["h"] * c["h"] for c in a[0])
                tag = name + ("_dec" if (a[0][0].get("gx_steps", 0) == 1 or a[0][0].get("dh_all") is not None) else "_enc_mfn")
            elif name in ("mfn_mem_fwd", "mfn_mem_bwd"):
                d = a[0]
                fl = 2.0 * d["T"] * d["B"] * 2 * (d["mem"] * (d["g1"] + d["g2"]))
            elif name in ("mmd_fwd", "mmd_bwd"):
                Bn, dim = a[0].shape
                fl = (9.0 if name == "mmd_fwd" else 8.0) * Bn * Bn * dim
            rec.append((tag, e0, e1, fl, by))
            return r
        return inner
    names = ["gemm", "gemm_tn_pair", "lstm_fwd", "lstm_bwd", "mfn_mem_fwd", "mfn_mem_bwd", "softmax_gate_fwd", "softmax_gate_bwd", "mmd_fwd",
             "mmd_bwd", "copy2d", "add", "zero", "colsum", "relu_bwd", "mse_fwd_bwd", "l1_fwd_bwd", "ce_fwd_bwd",
             "loss_total", "adam", "randn", "rng_tick", "rownorm2", "mmd_kexp", "mmd_kexp64", "mmd_fold", "mmd_combine"]
    orig = {n: getattr(ops, n) for n in names}
    agg = {}
    side = trainer.eng.use_side_stream
    trainer.eng.use_side_stream = False       # per-kernel timing needs everything on the timed stream
    world, trainer.world = trainer.world, 1   # rank 0 profiles alone: no collective may be issued here
    try:
        for n in names:
            setattr(ops, n, wrap(n, orig[n]))
        trainer._schedule()                       # one untimed pass
        torch.cuda.synchronize()
        rec.clear()
        shapes.clear()
        for _ in range(reps):
            # The eager pass is host-bound (~30 us of Python per primitive): on an idle GPU every event bracket would also
            # hold the host's launch latency (tensor-map encoding, two launches).  A spinning kernel in front of the pass keeps
            # the GPU busy until the whole pass is queued, so the brackets measure device time only.
            torch.cuda._sleep(int(4.0e7))
            trainer._schedule()
            torch.cuda.synchronize()
        for tag, e0, e1, fl, by in rec:
            a = agg.setdefault(tag, [0.0, 0.0, 0, 0.0])
            a[0] += e0.elapsed_time(e1)
            a[1] += fl
            a[2] += 1
            a[3] += by
    finally:
        trainer.eng.use_side_stream = side
        trainer.world = world
        for n in names:
            try:
                delattr(ops, n)
            except AttributeError:
                pass
    by_shape = {}
    stream_ms, stream_bytes, stream_n = 0.0, 0.0, 0
    for key, e0, e1, by, streamed in shapes:
        v = by_shape.setdefault(key, [0.0, 0, by])
        dt = e0.elapsed_time(e1)
        v[0] += dt / reps
        v[1] += 1
        if streamed:
            stream_ms += dt
            stream_bytes += by
            stream_n += 1
    # [total ms per step, launches per step, algorithmic bytes per launch]
    agg["_gemm_shapes"] = {k: [round(v[0], 4), v[1] // reps, v[2]] for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:40]}
    agg["_gemm_streamed"] = dict(ms_per_step=stream_ms / reps, bytes_per_step=stream_bytes / reps, launches_per_step=stream_n // reps)
    return agg


def parity_check(configs, T, B, head, dev, use_graph):
    """One fixed-seed TRAIN-MODE step of the measured configuration against the oracle (the CPU restatement of the
    reference, oracle/): a fresh model (seed 123) and trainer of the same shape as the timed one take two steps, and the
    third is compared -- its dropout masks / MMD noise / ReLU branches are replayed in oracle.train_step
    (mfm_mosi.py:427-441) from the trainer's parameters.  Returns the worst relative error over losses, latents and all
    gradients.  Rank 0, outside the timed region; the oracle is the checker only.  (The timed trainer itself is not used:
    after hundreds of steps on four synthetic batches its gradients are a badly conditioned target -- measured with
    scripts/parity_probe.py: relative errors grow from 2e-5 to 3e-4 over 500 steps on one batch on the tcgen05 path and
    from 4e-6 to 3e-5 on the exact-fp32 path.)"""
    import torch
    from collections import OrderedDict
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    from oracle import mfm_oracle as O
    from oracle.rng_replay import train_masks_and_branches
    torch.manual_seed(123)
    model = F.MFM(*configs).to(dev).train()
    trainer = MFMTrainer(model, T, B, head=head, use_graph=use_graph, seed=2024, distributed=False)
    for i in range(2):
        xw, yw = O.synthetic_batch(configs, T, B, 77 + i, head)
        trainer.step(xw.to(dev), yw.to(dev))
    torch.cuda.synchronize()
    P = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    x, y = O.synthetic_batch(configs, T, B, 4321, head)
    lb = trainer.step(x.to(trainer.dev), y.to(trainer.dev))
    torch.cuda.synchronize()
    masks, br = train_masks_and_branches(trainer.eng, trainer.rng.cpu())
    noise = [t.detach().cpu().clone() for t in trainer.noise]
    del O.RELU_REPLAY_VIOLATIONS[:]
    _, losses, Go, outo = O.train_step(P, x, y, configs, noise, {}, head=head, train=True, masks=masks, branches=br)
    rel = lambda a, b: float((a.detach().cpu().double() - b.double()).norm() / (b.double().norm() + 1e-30))
    lbc = lb.cpu()
    rep = {}
    for i, k in ((0, "disc"), (1, "mse_l"), (2, "mse_a"), (3, "mse_v"), (8, "total")):
        rep["loss." + k] = abs(float(lbc[i]) - losses[k]) / abs(losses[k])
    rep["loss.mmd"] = abs(float(lbc[4:8].sum()) * configs[0]["lda_mmd"] - losses["mmd"]) / abs(losses["mmd"])
    ws = trainer.eng.ws
    for k, b in (("zl", "Z0"), ("za", "Z1"), ("zv", "Z2"), ("zy", "ZY"), ("y_hat", "Yhat")):
        rep[k] = rel(ws[b], outo[k])
    for k, go in Go.items():
        if go is not None:
            rep["grad." + k] = rel(trainer.G[k], go)
    worst = max(rep, key=rep.get)
    top = sorted(rep, key=rep.get, reverse=True)[:6]
    return dict(worst_rel=rep[worst], worst=worst, n_compared=len(rep), top={k: float("%.3g" % rep[k]) for k in top}, relu_replay_violations=len(O.RELU_REPLAY_VIOLATIONS),
                tolerance=1e-3, passed=bool(rep[worst] < 1e-3 and not O.RELU_REPLAY_VIOLATIONS),
                config="fresh model (seed 123), third train-mode step (9 dropouts, device RNG), batch %d, T=%d, %s head, %s; masks / "
                       "noise / ReLU branches replayed in oracle.train_step" % (B, T, head, "CUDA graph" if use_graph else "eager"))


def top_gemm_shape(dm):
    """The streamed GEMM shape with the most algorithmic bytes per launch in the step: the attention products over
    [T*B, 2H] (att1_fc2: [T*B,a1] x [2H,a1]^T -> [T*B,2H]).  Chosen from the dimensions, NOT from a timing, so the ncu
    DRAM traffic recorded for it (profiles/r2_gemm_tcp_ncu.json) can be looked up deterministically."""
    M, N, K = dm.T * dm.B, 2 * dm.H, dm.a1
    return "nt %dx%dx%d" % (M, N, K), 4.0 * (M * K + N * K + M * N)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="mosi", choices=["mosi", "mosei", "iemocap", "pom"],
                    help="BASELINE.json workload (default: the headline, MOSI shapes batch 2048 per GPU)")
    ap.add_argument("--batch", type=int, default=0, help="batch (per GPU for weak scaling, global for strong); 0 = the workload's own")
    ap.add_argument("--scaling", default="", choices=["", "weak", "strong"],
                    help="weak: the batch is per GPU; strong: the batch is global and split over the ranks")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from factorized_b200.configs import WORKLOADS, best_acc_configs
    wl = WORKLOADS[args.config]
    T, head = wl["T"], wl["head"]
    scaling = args.scaling or ("strong" if wl["batch_is"] == "global" else "weak")
    batch = args.batch or wl["batch"]
    if scaling == "strong":
        if batch % max(world, 1):
            raise SystemExit("strong scaling: global batch %d is not divisible by %d ranks" % (batch, world))
        B, global_batch = batch // max(world, 1), batch
    else:
        B, global_batch = batch, batch * max(world, 1)
    configs = best_acc_configs(input_dims=wl["input_dims"], output_dim=wl["out"], dropout=True)
    gate_fl = gate_train_flops_per_sample(configs, T)
    workload = "%s; batch %d per GPU (global %d, %s scaling), best_acc hyper-parameters (mfm_mosi.py:1239-1286), fp32 storage, " \
               "dropout active, %s head" % (wl["desc"], B, global_batch, scaling, "L1" if head == "l1" else "cross-entropy")
    metric = "MFM train samples/sec (MOSI seq_len=20)" if args.config == "mosi" else "MFM train samples/sec (%s shapes, seq_len=%d)" % (args.config, T)
    D = sum(wl["input_dims"])
    in_mb = T * B * D * 4 / 2 ** 20
    base = dict(metric=metric, unit="samples/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, higher_is_better=True, scaling=scaling, vs_baseline=None, data="synthetic",
                config=dict(workload=workload, name=args.config, batch_per_gpu=B, global_batch=global_batch, seq_len=T,
                            parallelism="dp%d" % max(world, 1),
                            l2="4 distinct input batches rotate (%.0f MB) and the step streams its whole stash workspace (see "
                               "workspace_mb), together > 126 MB L2; no explicit flush" % (4 * in_mb)))
    DTYPE = "f32 (fp32 storage; contractions as bf16x3 split tcgen05 MMA with f32 accumulate, ~2^-16)"

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        from oracle import ref_runner
        if ref_runner.available():           # the UNMODIFIED reference model (oracle/_ref, built by oracle/build_ref.py)
            sec, threads = ref_runner.time_train_steps(configs, T, B, steps, warm, head)
            cb = dict(value=B / sec, unit="samples/s", cores=threads, kind="reference",
                      sample="%d warm-up + %d timed steps of batch %d (T=%d): the unmodified reference MFM (oracle/_ref, "
                             ".cuda() neutralised) under the py3 restatement of mfm_mosi.py:427-441, torch fp32, %.2f s/step"
                             % (warm, steps, B, T, sec))
        else:
            cb, sec = cpu_reference_leg(configs, T, B, steps, warm, head)
        line = dict(base, impl="reference", value=cb["value"], ms_per_step=sec * 1e3, dtype="f32", cpu_baseline=cb,
                    steps=steps, warmup=warm,
                    e2e=dict(value=cb["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0, n_gpus=args.gpus)
        print(json.dumps(line))
        return

    import torch
    import factorized_b200 as F
    from factorized_b200.train import MFMTrainer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(123)
    model = F.MFM(*configs).to(dev).train()
    trainer = MFMTrainer(model, T, B, head=head, use_graph=not args.no_graph, seed=123)
    gen = torch.Generator().manual_seed(1234 + rank)
    NB = 4

    def make_y():
        if head == "ce":
            return torch.randint(0, wl["out"], (B,), generator=gen)
        return torch.randn(B * wl["out"], generator=gen)
    xs_host = [torch.randn(T, B, D, generator=gen).pin_memory() for _ in range(NB)]
    ys_host = [make_y().pin_memory() for _ in range(NB)]
    xs_dev = [t.to(dev) for t in xs_host]
    ys_dev = [t.to(dev) for t in ys_host]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t)
        return ms

    # ---- (A) inputs resident in HBM -------------------------------------------------------------------------------
    for i in range(args.warmup):
        trainer.x.copy_(xs_dev[i % NB]); trainer.y.copy_(ys_dev[i % NB]); trainer.step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as cs:
        e0.record()
        for i in range(args.steps):
            trainer.x.copy_(xs_dev[i % NB]); trainer.y.copy_(ys_dev[i % NB]); trainer.step_device()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        # ---- (B) end to end through the public API: pinned host batch -> H2D -> step -> D2H loss, every step ----------
        loss_host = torch.zeros(args.steps, 16).pin_memory()
        for i in range(min(args.warmup, 3)):
            trainer.step(xs_host[i % NB], ys_host[i % NB])
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            lb = trainer.step(xs_host[i % NB], ys_host[i % NB])
            loss_host[i].copy_(lb, non_blocking=True)
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        if ms + ms_e2e < 1500.0:             # short timed regions: keep the sampler up until it has seen the GPU under load
            t_end = time.time() + 1.2
            while time.time() < t_end:
                trainer.step_device()
            torch.cuda.synchronize()
    clocks = cs.summary()
    final_loss = float(loss_host[-1][8])
    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    pk = peaks()

    # ---- (C) the other scaling mode, briefly: a weak run also reports the STRONG-scaling rate of the same global batch ------
    # (every rank runs this identically: the second trainer's all-reduce is a collective)
    other = None
    if world > 1 and scaling == "weak" and batch % world == 0 and batch // world >= 16:
        Bs = batch // world
        torch.manual_seed(123)
        model_s = F.MFM(*configs).to(dev).train()
        tr_s = MFMTrainer(model_s, T, Bs, head=head, use_graph=not args.no_graph, seed=123)
        xs_s = [t[:, :Bs].contiguous() for t in xs_dev[:2]]
        ys_s = [t.view(B, -1)[:Bs].reshape(-1).contiguous() for t in ys_dev[:2]]
        for i in range(max(args.warmup, 3)):
            tr_s.x.copy_(xs_s[i % 2]); tr_s.y.copy_(ys_s[i % 2]); tr_s.step_device()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(args.steps):
            tr_s.x.copy_(xs_s[i % 2]); tr_s.y.copy_(ys_s[i % 2]); tr_s.step_device()
        g1.record()
        barrier()
        ms_s = max_over_ranks(g0.elapsed_time(g1))
        other = dict(scaling="strong", global_batch=batch, per_gpu_batch=Bs, ms_per_step=ms_s / args.steps,
                     value=world * Bs * args.steps / (ms_s * 1e-3), unit="samples/s",
                     note="same workload with the GLOBAL batch held at %d (split over the ranks); %d timed steps after the main run" % (batch, args.steps))
        del tr_s, model_s

    # ---- replicas must hold identical parameters after the same number of steps (the all-reduce keeps them in sync) -----
    replicas = None
    if world > 1:
        chk = trainer.flat_p.double().sum().view(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        torch.distributed.all_gather(allc, chk)
        vals = [float(v) for v in allc]
        replicas = dict(param_checksums_equal=bool(max(vals) - min(vals) <= 1e-9 * max(1.0, abs(vals[0]))), checksum=vals[0])

    # ---- dominant kernel + roofline (rank 0, outside the timed region) --------------------------------------------------
    roof, kernels, parity = None, None, None
    if rank == 0:
        agg = profile_primitives(trainer)
        gemm_shapes = agg.pop("_gemm_shapes")
        streamed = agg.pop("_gemm_streamed")
        tot = sum(v[0] for v in agg.values())
        kernels = {k: dict(ms_per_step=round(v[0] / 3, 4), share=round(v[0] / tot, 4), launches=v[2] // 3,
                           tflops=(round(v[1] / (v[0] * 1e-3) / 1e12, 3) if v[1] else None),
                           gbs=(round(v[3] / (v[0] * 1e-3) / 1e9, 1) if v[3] else None)) for k, v in
                   sorted(agg.items(), key=lambda kv: -kv[1][0])}
        # The dominant kernel class is the streamed tcgen05 GEMM (gemm_ps_kernel for activations x weights, gemm_tcp_kernel for
        # the weight gradients and the masked / accumulating epilogues: every GEMM of M*N*K >= 2^20).  It is
        # HBM-bound by design (N, K <= 512 against T*B rows): achieved = algorithmic bytes of those launches / their
        # CUDA-event time in one eager pass (per call: the weight pre-split kernel, when there is one, is inside the
        # bracket).  `traffic` = DRAM bytes per launch that ncu --set full measured for the largest streamed shape of this
        # workload (chosen from the dimensions, not from a timing; profiles/r2_gemm_tcp_ncu.json).
        gshare = streamed["ms_per_step"] / (tot / 3)
        ach = streamed["bytes_per_step"] / (streamed["ms_per_step"] * 1e-3) / 1e9 if streamed["ms_per_step"] else 0.0
        top_key, top_bytes = top_gemm_shape(trainer.eng.dm)
        tms, tn, tby = gemm_shapes.get(top_key, (None, None, top_bytes))
        traffic, traffic_src = None, "profiles/r2_gemm_tcp_ncu.json has no entry for '%s' (run scripts/gemm_prof.py under ncu)" % top_key
        try:
            nj = json.load(open(os.path.join(ROOT, "profiles", "r2_gemm_tcp_ncu.json")))
            if top_key in nj:
                traffic, traffic_src = nj[top_key]["dram_bytes"], "ncu --set full, profiles/r2_gemm_tcp_ncu.json"
        except Exception as ex:
            traffic_src = "profiles/r2_gemm_tcp_ncu.json unreadable: %s" % type(ex).__name__
        # the recurrence kernels on the TENSOR roof: recurrent gate-GEMM FLOPs of the launches (2*4h*h per
        # row and step; the input-projection half of the gate GEMM is hoisted into the streamed GEMMs above) / their time
        lstm = {}
        for k in ("lstm_fwd_enc_mfn", "lstm_bwd_enc_mfn", "lstm_fwd_dec", "lstm_bwd_dec"):
            if k in kernels and kernels[k]["tflops"]:
                lstm[k] = dict(us_per_step=round(kernels[k]["ms_per_step"] * 1e3, 1), tflops=kernels[k]["tflops"],
                               frac_of_bf16_peak=round(kernels[k]["tflops"] / pk["bf16"], 5))
        roof = dict(bound="hbm", kernel="gemm_ps_kernel + gemm_tcp_kernel (streamed tcgen05 GEMMs: the %d launches/step with "
                                        "M*N*K >= 2^20; the top shape below runs on the persistent gemm_ps_kernel)"
                                        % streamed["launches_per_step"],
                    achieved=round(ach, 1), peak=pk["hbm"], unit="GB/s", frac=round(ach / pk["hbm"], 4), traffic=traffic,
                    traffic_source=traffic_src,
                    peak_source=pk["src"] + " HBM copy bandwidth (MEASURED_PEAKS.json)", share_of_step_kernel_time=round(gshare, 4),
                    top_shape=dict(shape=top_key, launches=tn, bytes_per_launch=tby,
                                   us_per_launch=(round(tms / tn * 1e3, 2) if tms else None),
                                   achieved=(round(tby / (tms / tn * 1e-3) / 1e9, 1) if tms else None), traffic=traffic),
                    lstm_tensor=dict(bound="tensor", peak=pk["bf16"], unit="TFLOP/s", kernels=lstm,
                                     note="recurrent gate-GEMM FLOPs (2*4h*h per row-step, reference formulation) of the "
                                          "recurrence launches / CUDA-event time, against the measured burst bf16 peak; the "
                                          "kernels are bound by the per-step dependency chain (DESIGN.md section 4.2), ncu "
                                          "counters in profiles/r2_lstm_ws_ncu.md"),
                    step_gate_gemm=dict(achieved=round(gate_fl * value / max(world, 1) / 1e12, 3), peak=pk["bf16"],
                                        unit="TFLOP/s", frac=round(gate_fl * value / max(world, 1) / 1e12 / pk["bf16"], 5),
                                        frac_of_sustained=round(gate_fl * value / max(world, 1) / 1e12 / pk["bf16_sustained"], 5),
                                        note="whole step per GPU: algorithmic gate-GEMM train FLOPs/sample (%.2f M) x samples/s "
                                             "over the measured burst bf16 peak (clocks held at max, no power cap)" % (gate_fl / 1e6)))
        if not args.no_parity_check:
            try:
                parity = parity_check(configs, T, B, head, dev, not args.no_graph)
            except Exception as ex:          # a host without the memory for the oracle's [B,B,dim] MMD tensors
                parity = dict(passed=None, error="%s: %s" % (type(ex).__name__, str(ex)[:200]))
    line = dict(base, value=value, ms_per_step=ms / args.steps, dtype=DTYPE,
                e2e=dict(value=e2e_value, unit="samples/s", h2d_bytes_per_step=(T * B * D + B * (1 if head == "ce" else wl["out"])) * 4 * world,
                         d2h_bytes_per_step=64 * world, ms_per_step=ms_e2e / args.steps),
                gpu_launches=trainer.launches_per_step * args.steps, launches_per_step=trainer.launches_per_step,
                cuda_graph=not args.no_graph, clocks=clocks, roofline=roof, kernels=kernels, final_loss=final_loss,
                strong_scaling=other,
                parity_check=parity, replicas=replicas,
                workspace_mb=round(trainer.eng.workspace_bytes() / 2 ** 20, 1))
    if rank == 0 and os.environ.get("MFM_BENCH_GEMM_SHAPES"):
        line["gemm_shapes_ms"] = gemm_shapes
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_runner
            try:
                if ref_runner.available():
                    sec, threads = ref_runner.time_train_steps(configs, T, B, 2, 1, head)
                    cb = dict(value=B / sec, unit="samples/s", cores=threads, kind="reference",
                              sample="1 warm-up + 2 timed steps of batch %d (T=%d): the unmodified reference MFM (oracle/_ref) under "
                                     "the py3 restatement of mfm_mosi.py:427-441, torch fp32, %.2f s/step" % (B, T, sec))
                else:
                    cb, _ = cpu_reference_leg(configs, T, B, 2, 1, head)
            except Exception as ex:   # e.g. host OOM on the reference's [B,B,dim] MMD tensors
                cb, _ = cpu_reference_leg(configs, T, 256, 3, 1, head)
                cb["sample"] += " (batch %d failed on this host: %s)" % (B, type(ex).__name__)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
