"""CPU ORACLE (test infrastructure, NOT product code): explicit-loop numpy restatement of the encoder.

Independent of torch.nn.LSTM -- writes out SURVEY.md Appendix A line by line so that the torch-module
oracle (oracle/awd_lstm_ref.py) is cross-checked by a second implementation:

    x^0_t = Emb[ids[:, t]]                                        (F.embedding; padding_idx affects grads only)
    z     = x_t W_ih^T + b_ih + h_{t-1} W_hh^T + b_hh             rows ordered i | f | g | o  (torch.nn.LSTM)
    i,f,o = sigmoid(z_i), sigmoid(z_f), sigmoid(z_o); g = tanh(z_g)
    c_t   = f*c_{t-1} + i*g ; h_t = o*tanh(c_t)                   h_{-1}=c_{-1}=0 (inference.py:56,66 reset())
    out[b] = [mean_{t<len} y | max_{t<len} y | y[len-1]]          (inference.py:239)

Use only for small cases (pure numpy, float64 or float32 accumulations selectable).
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_layer(x, w_ih, w_hh, b_ih, b_hh, dtype=np.float64):
    """x (B,T,in) -> (B,T,out)."""
    x = x.astype(dtype)
    w_ih, w_hh, b_ih, b_hh = (a.astype(dtype) for a in (w_ih, w_hh, b_ih, b_hh))
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = np.zeros((B, H), dtype)
    c = np.zeros((B, H), dtype)
    ys = np.empty((B, T, H), dtype)
    for t in range(T):
        z = x[:, t] @ w_ih.T + b_ih + h @ w_hh.T + b_hh
        i, f, g, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        c = _sigmoid(f) * c + _sigmoid(i) * np.tanh(g)
        h = _sigmoid(o) * np.tanh(c)
        ys[:, t] = h
    return ys


def encode(emb, layers, ids, lengths, dtype=np.float64):
    """emb (V,E); layers list of dict(w_ih,w_hh,b_ih,b_hh); ids (B,T) right padded -> (B,3E)."""
    x = emb[np.asarray(ids)]
    for L in layers:
        x = lstm_layer(x, L['w_ih'], L['w_hh'], L['b_ih'], L['b_hh'], dtype)
    out = []
    for b, n in enumerate(lengths):
        e = x[b, :n]
        out.append(np.concatenate([e.mean(0), e.max(0), e[-1]]))
    return np.stack(out), x


def mlp_forward(X, coefs, intercepts, dtype=np.float64):
    """sklearn MLPClassifier._forward_pass_fast for relu hidden + logistic output
    (py/label_microservice/mlp.py:63 -> predict_proba; multilabel => out_activation_ 'logistic').
    coefs[i] is stored [fan_in, fan_out]."""
    a = np.asarray(X, dtype)
    n = len(coefs)
    for i, (W, b) in enumerate(zip(coefs, intercepts)):
        a = a @ np.asarray(W, dtype) + np.asarray(b, dtype)
        if i != n - 1:
            a = np.maximum(a, 0)
    return 1.0 / (1.0 + np.exp(-a))


def filter_labels(label_names, probs, thresholds):
    """py/label_microservice/repo_specific_model.py:126-146 -- keep label iff its threshold is truthy
    and prob >= threshold."""
    out = {}
    for name, p in zip(label_names, probs):
        thr = thresholds[name]
        if not thr:
            continue
        if p < thr:
            continue
        out[name] = p
    return out
