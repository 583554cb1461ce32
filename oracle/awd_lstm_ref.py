"""CPU ORACLE (test infrastructure, NOT product code) for the Issue_Embeddings encoder hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (code_intelligence_b200) never does and fails loudly without its CUDA
library.

What it restates
----------------
The reference's encoder arithmetic lives in third-party packages that are NOT under /root/reference and
are not installed here: ``fastai==1.0.53.post3`` (Issue_Embeddings/requirements.txt:22) wrapping
``torch==1.1.0`` ``nn.Embedding`` / ``nn.LSTM`` (requirements.txt:110).  This file restates that
published algorithm on the very torch modules fastai wraps, anchored on the reference's own call sites:

* model structure  Issue_Embeddings/notebooks/04_Inference.ipynb:157-187
                   Embedding(60000,800,padding_idx=1) -> LSTM(800,2400) -> LSTM(2400,2400) x2 -> LSTM(2400,800)
* reset + forward  Issue_Embeddings/flask_app/inference.py:55-57, 59-68   (zero state on every call,
                   ``encoder.forward(x)[-1][-1]`` = last layer's hidden states, (B,T,emb_sz))
* single pooling   Issue_Embeddings/flask_app/inference.py:71-90          ([mean | max | last], unmasked)
* masked pooling   Issue_Embeddings/flask_app/inference.py:215-246 / py/code_intelligence/inference.py:232-263
* bulk driver      py/code_intelligence/inference.py:171-229             (bs rule, sort, pad, unsort)

Eval-mode fastai semantics (restated from fastai 1.0.53 source knowledge): RNNDropout /
EmbeddingDropout / WeightDropout are the identity when ``not training``; the LSTM layer dims follow
``in_0=emb_sz, in_l=n_hid, out_l=n_hid (l<L-1), out_{L-1}=emb_sz``.

PARITY STATUS: the arithmetic core is "parity unpinned" by the reference's own tests -- the reference has no test
files and no usable golden vectors for this path, and fastai's AWD_LSTM.forward cannot be executed
here; that it equals this stack of nn.LSTM in eval mode is restated, not run.  Everything AROUND that core is pinned on
reference code executed in the build container: tests/golden/reference_driver.npz holds what the reference's own
df_to_embedding / batch_seq_pool / get_pooled_features (py/code_intelligence/inference.py:74-92,138-263) returned when
run around this module's nn.LSTM stack (tests/golden/make_golden.py driver), and encode_bulk / batch_seq_pool /
encode_single below are checked against it (tests/test_host_logic.py).  The oracle is further pinned by (i) the
reference's portable invariant bulk == single within atol 1e-5
(Issue_Embeddings/notebooks/04b_Inference-Batch.ipynb:369), checked in tests/test_oracle.py, (ii) an
independent explicit-loop numpy LSTM (oracle/lstm_numpy.py) and (iii) committed golden vectors generated
from this file (tests/golden/, generator tests/golden/make_golden.py).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
from torch import nn
from torch.nn.utils.rnn import pad_sequence

PAD_IDX = 1      # fastai default pad token id (inference.py:36 learn.data.pad_idx)
BOS_IDX = 2      # xxbos
VOCAB_SZ = 60000
EMB_SZ = 800
N_HID = 2400


def layer_dims(n_layers: int, emb_sz: int, n_hid: int):
    """fastai AWD_LSTM layer rule (matches 04_Inference.ipynb:165-174)."""
    dims = []
    for l in range(n_layers):
        n_in = emb_sz if l == 0 else n_hid
        n_out = n_hid if l != n_layers - 1 else emb_sz
        dims.append((n_in, n_out))
    return dims


class AWDLSTMEncoderRef(nn.Module):
    """Eval-mode restatement of fastai 1.0.53 ``AWD_LSTM`` (encoder only)."""

    def __init__(self, vocab_sz=VOCAB_SZ, emb_sz=EMB_SZ, n_hid=N_HID, n_layers=4, pad_idx=PAD_IDX):
        super().__init__()
        self.vocab_sz, self.emb_sz, self.n_hid, self.n_layers, self.pad_idx = vocab_sz, emb_sz, n_hid, n_layers, pad_idx
        self.encoder = nn.Embedding(vocab_sz, emb_sz, padding_idx=pad_idx)
        self.rnns = nn.ModuleList([nn.LSTM(i, o, 1, batch_first=True) for i, o in layer_dims(n_layers, emb_sz, n_hid)])
        # fastai: self.encoder.weight.data.uniform_(-initrange, initrange) with initrange = 0.1
        # (overwrites the padding row too)
        with torch.no_grad():
            self.encoder.weight.uniform_(-0.1, 0.1)
        self.eval()

    @torch.no_grad()
    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        """ids int64 (B,T) -> last layer hidden states (B,T,emb_sz); zero initial state every call
        (inference.py:56,66 ``encoder.reset()``)."""
        x = self.encoder(ids)
        for rnn in self.rnns:
            x, _ = rnn(x)            # h0 = c0 = 0
        return x

    # ---- weight export in the layout the C ABI takes (include/issue_emb_b200.h) -------------------
    def export_weights(self):
        emb = self.encoder.weight.detach().float().numpy().copy()
        layers = []
        for rnn in self.rnns:
            layers.append(dict(
                w_ih=rnn.weight_ih_l0.detach().float().numpy().copy(),
                w_hh=rnn.weight_hh_l0.detach().float().numpy().copy(),
                b_ih=rnn.bias_ih_l0.detach().float().numpy().copy(),
                b_hh=rnn.bias_hh_l0.detach().float().numpy().copy()))
        return emb, layers


def make_encoder(seed=1234, vocab_sz=VOCAB_SZ, emb_sz=EMB_SZ, n_hid=N_HID, n_layers=4, scale=1.0) -> AWDLSTMEncoderRef:
    """Deterministic random-init encoder (BASELINE.json configs: 'random-init AWD-LSTM weights').

    ``scale`` > 1 multiplies the LSTM weight matrices to give a "trained-like" activation magnitude
    (|h| ~ 0.1 as printed in 04_Inference.ipynb:430-461) -- a more discriminative parity weight set
    than torch's default init."""
    torch.manual_seed(seed)
    enc = AWDLSTMEncoderRef(vocab_sz, emb_sz, n_hid, n_layers)
    if scale != 1.0:
        with torch.no_grad():
            for rnn in enc.rnns:
                rnn.weight_ih_l0.mul_(scale)
                rnn.weight_hh_l0.mul_(scale)
    return enc


def synthetic_ids(n: int, T: int, seed=1234, vocab_sz=VOCAB_SZ, min_len=None) -> List[np.ndarray]:
    """Synthetic token-id issues: ids[0]=xxbos(2), rest uniform over the vocab with pad(1) remapped to 0
    (SURVEY.md section 8d).  If min_len is given lengths are uniform in [min_len, T], else all T."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        L = T if min_len is None else int(rng.integers(min_len, T + 1))
        a = rng.integers(0, vocab_sz, size=L, dtype=np.int64)
        a[a == PAD_IDX] = 0
        a[0] = BOS_IDX
        out.append(a)
    return out


# ---------------------------------------------------------------------------------------------------
# pooling: the two forms the reference has
# ---------------------------------------------------------------------------------------------------
def pooled_single(raw: torch.Tensor) -> torch.Tensor:
    """inference.py:90 -- cat([mean(dim=1), max(dim=1)[0], raw[:,-1,:]], -1); unmasked."""
    return torch.cat([raw.mean(dim=1), raw.max(dim=1)[0], raw[:, -1, :]], dim=-1)


def batch_seq_pool(seq_emb: np.ndarray, lengths: Sequence[int]) -> np.ndarray:
    """inference.py:215-246 -- masked [mean | max | last] per row over the first len_i steps."""
    assert seq_emb.shape[0] == len(lengths), 'Number of elements in lengths should match the first dimension of seq_emb'
    embs = [seq_emb[i, :x, :] for i, x in enumerate(lengths)]
    features = [np.concatenate([emb.mean(axis=0), emb.max(axis=0), emb[-1, :]], axis=-1) for emb in embs]
    combined = np.stack(features)
    assert combined.shape[-1] == seq_emb.shape[-1] * 3
    return combined


# ---------------------------------------------------------------------------------------------------
# the two encode paths of InferenceWrapper, on token ids
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def encode_single(enc: AWDLSTMEncoderRef, ids: np.ndarray) -> np.ndarray:
    """get_pooled_features on one numericalised issue (inference.py:59-90) -> (1, 3*emb_sz)."""
    x = torch.as_tensor(np.asarray(ids, dtype=np.int64))[None, :]
    return pooled_single(enc(x)).numpy()


@torch.no_grad()
def encode_padded(enc: AWDLSTMEncoderRef, ids: np.ndarray, lengths: Sequence[int]) -> np.ndarray:
    """_forward_pass + batch_seq_pool on one right-padded (B,T) batch (inference.py:55-57, 206)."""
    hidden = enc(torch.as_tensor(np.asarray(ids, dtype=np.int64))).numpy()
    return batch_seq_pool(hidden, lengths)


@torch.no_grad()
def encode_bulk(enc: AWDLSTMEncoderRef, docs: List[np.ndarray], bs: int = 100) -> np.ndarray:
    """df_to_embedding from the numericalised docs on (py/code_intelligence/inference.py:173-229)."""
    bs = min(bs, (len(docs) // 20) + 1)
    length_arr = np.array([d.shape[0] for d in docs])
    len_mask = length_arr.argsort()
    len_mask_reversed = len_mask.argsort()
    ordered = [torch.as_tensor(np.asarray(docs[i], dtype=np.int64)) for i in len_mask]
    ordered_lengths = length_arr[len_mask]
    pooled = []
    i, total = 0, len(docs)
    while i < total:
        bp = pad_sequence(ordered[i:i + bs], batch_first=True, padding_value=enc.pad_idx)
        hidden = enc(bp).numpy()
        pooled.append(batch_seq_pool(hidden, ordered_lengths[i:i + bs]))
        i += bs
    out = np.concatenate(pooled)[len_mask_reversed, :]
    assert out.shape[0] == len(docs)
    return out


# ---------------------------------------------------------------------------------------------------
# parity metrics: mandated cosine + discriminative extras
# ---------------------------------------------------------------------------------------------------
def parity_metrics(got: np.ndarray, ref: np.ndarray) -> dict:
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    num = (got * ref).sum(-1)
    den = np.linalg.norm(got, axis=-1) * np.linalg.norm(ref, axis=-1)
    cos = num / np.maximum(den, 1e-300)
    E = ref.shape[-1] // 3
    m = {
        'min_cosine': float(cos.min()),
        'max_abs': float(np.abs(got - ref).max()),
        'rel_l2': float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)),
        'max_abs_mean_seg': float(np.abs(got[:, :E] - ref[:, :E]).max()),
        'max_abs_max_seg': float(np.abs(got[:, E:2 * E] - ref[:, E:2 * E]).max()),
        'max_abs_last_seg': float(np.abs(got[:, 2 * E:] - ref[:, 2 * E:]).max()),
    }
    if ref.shape[0] > 1:
        mu = ref.mean(0, keepdims=True)
        g, r = got - mu, ref - mu
        cc = (g * r).sum(-1) / np.maximum(np.linalg.norm(g, axis=-1) * np.linalg.norm(r, axis=-1), 1e-300)
        m['min_centred_cosine'] = float(cc.min())
    return m
