"""Run the UNMODIFIED reference model from oracle/_ref on the CPU -- BASELINE INFRASTRUCTURE (bench.py --impl reference
and the cpu_baseline leg only).

oracle/_ref/mfm_model.pyc.bin is the byte-compiled /root/reference/mfm_model.py (oracle/build_ref.py).  The reference
hard-codes `.cuda()` inside forward (mfm_model.py:29,51-52,76-77,147-153); to time its own CPU path `Tensor.cuda` is
neutralised for the duration of the run, exactly as SURVEY.md section 8c / BASELINE.md section 3 prescribe.  The 25-line
step around the model restates mfm_mosi.py:427-441 (CE head: mfm_mosi_acc.py:441-452) -- the scripts themselves are
Python 2 with lab-local data and cannot run.
"""
import importlib.machinery
import importlib.util
import os
import time

HERE = os.path.dirname(os.path.abspath(__file__))
PYC = os.path.join(HERE, "_ref", "mfm_model.pyc.bin")


def available():
    return os.path.exists(PYC) and os.environ.get("MFM_NO_REF", "0") != "1"


def load_reference():
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU timing run: the reference's .cuda() calls become no-ops
    loader = importlib.machinery.SourcelessFileLoader("mfm_model_reference", PYC)
    spec = importlib.util.spec_from_loader("mfm_model_reference", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def time_train_steps(configs, T, B, steps, warmup, head="l1", threads=None):
    """Seconds per training step of the reference MFM (train mode, dropout active, Adam default lr) on synthetic data."""
    import torch
    from oracle import mfm_oracle as O
    torch.set_num_threads(threads or os.cpu_count() or 1)
    ref = load_reference()
    torch.manual_seed(123)
    model = ref.MFM(*configs)
    model.train()
    opt = torch.optim.Adam(model.parameters())               # mfm_mosi.py:403
    x, y = O.synthetic_batch(configs, T, B, 1234, head)
    c = configs[0]
    d_l, d_a, d_v = c["input_dims"]
    F = torch.nn.functional
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()                                       # :427
        decoded, mmd, missing = model.forward(x)              # :430
        x_l_hat, x_a_hat, x_v_hat, y_hat = decoded
        yh = y_hat.squeeze(1) if y_hat.shape[1] == 1 else y_hat
        gen = c["lda_xl"] * F.mse_loss(x_l_hat, x[:, :, :d_l]) + c["lda_xa"] * F.mse_loss(x_a_hat, x[:, :, d_l:d_l + d_a]) \
            + c["lda_xv"] * F.mse_loss(x_v_hat, x[:, :, d_l + d_a:])                               # :437
        disc = F.cross_entropy(yh, y.long()) if head == "ce" else F.l1_loss(yh, y)                # :438
        loss = disc + gen + c["lda_mmd"] * mmd + missing                                          # :439
        loss.backward()
        opt.step()
        disc.item()                                           # :442
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), torch.get_num_threads()
