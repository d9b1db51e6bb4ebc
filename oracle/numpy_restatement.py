"""The third-party arithmetic of the path, restated from its PUBLISHED definitions in plain numpy -- TEST INFRASTRUCTURE.

Every number the reference's MFM step produces is computed by PyTorch (pinned by the reference as "PyTorch 0.4.0", README.md:19;
no lock file) -- nothing under /root/reference implements an LSTM cell, a loss or Adam.  oracle/mfm_oracle.py restates the
reference's *composition* of those ops and is pinned to the live reference; this file pins the *ops themselves* to their
documented semantics, independently of whichever torch is installed, so that a silent change of an op's meaning between the
pinned and the installed version could not hide behind "oracle == reference" (both would move together):

* nn.LSTMCell (docs, 0.4 and 2.x alike): gates = x W_ih^T + b_ih + h W_hh^T + b_hh, chunked in the order **i, f, g, o**;
  c' = sigmoid(f) c + sigmoid(i) tanh(g); h' = sigmoid(o) tanh(c').          call sites: mfm_model.py:56, 83-85, 167-169
* nn.Linear: y = x W^T + b.                                                    :61, 90, 174-179, 535-552
* F.softmax(dim=1), F.relu, F.tanh, F.sigmoid.                                 :174-179
* nn.MSELoss / nn.L1Loss default reduction: the MEAN over all elements; nn.CrossEntropyLoss: mean over rows of
  -log_softmax(logits)[label].                                                 mfm_mosi.py:411-412, 437-438; mfm_mosi_acc.py:423
* optim.Adam defaults (lr 1e-3, betas (0.9, 0.999), eps 1e-8, no weight decay): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)  -- eps is added OUTSIDE the bias-corrected root.   mfm_mosi.py:403
* compute_kernel / loss_MMD are the reference's own code (mfm_model.py:14-34), restated here with the [n, m, dim] tensor it
  builds.
"""
import numpy as np


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    g = x @ w_ih.T + b_ih + h @ w_hh.T + b_hh
    n = h.shape[1]
    i, f, gg, o = g[:, :n], g[:, n:2 * n], g[:, 2 * n:3 * n], g[:, 3 * n:]
    c2 = sigmoid(f) * c + sigmoid(i) * np.tanh(gg)
    return sigmoid(o) * np.tanh(c2), c2


def linear(x, w, b):
    return x @ w.T + b


def softmax_rows(x):
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def mse_loss(a, b):
    return float(np.mean((a - b) ** 2))


def l1_loss(a, b):
    return float(np.mean(np.abs(a - b)))


def cross_entropy(logits, labels):
    z = logits - logits.max(axis=1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
    return float(-np.mean(logp[np.arange(len(labels)), labels]))


def adam_step(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    p = p - lr / (1 - b1 ** t) * m / (np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps)
    return p, m, v


def compute_kernel(x, y):
    """mfm_model.py:14-23 with its tiled [n, m, dim] tensor."""
    dim = x.shape[1]
    d = ((x[:, None, :] - y[None, :, :]) ** 2).mean(axis=2) / float(dim)
    return np.exp(-d)


def loss_mmd(z, g):
    """mfm_model.py:30-33."""
    return float(compute_kernel(g, g).mean() + compute_kernel(z, z).mean() - 2.0 * compute_kernel(g, z).mean())


def encoder_lstm(x, w_ih, w_hh, b_ih, b_hh, fw, fb):
    """encoderLSTM.forward, mfm_model.py:47-62: zero state, T cell steps, fc1 of the last hidden state."""
    h = np.zeros((x.shape[1], w_hh.shape[1]))
    c = np.zeros_like(h)
    for t in range(x.shape[0]):
        h, c = lstm_cell(x[t], h, c, w_ih, w_hh, b_ih, b_hh)
    return linear(h, fw, fb)
