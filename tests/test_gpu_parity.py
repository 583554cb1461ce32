"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI (ctypes -> libissue_emb_b200.so);
the CPU oracle (oracle/) and the committed golden vectors (tests/golden/) are the checkers.

Tolerances (stated per BASELINE.json north_star: cosine >= 1 - 1e-4, max-abs reported):
  * bf16 operands, f32 accumulate / state / pooling  ->  per-issue cosine >= 1 - 1e-4 (mandated) AND, because raw
    cosine is nearly blind under random init, rel-L2 <= 4e-3, centred cosine >= 0.99 and a
    negative control that must fail.
  * structural properties (batch/padding/prefix invariance, determinism) are bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import awd_lstm_ref as R
from oracle import lstm_numpy as N

pytestmark = pytest.mark.gpu

COS_MIN = 1 - 1e-4
REL_L2_MAX = 4e-3          # torch-default init (|h| ~ 0.01)
REL_L2_MAX_SCALED = 1e-2   # "trained-like" weight sets (LSTM weights x2..x3, |h| ~ 0.1): bf16 rounding of h amplifies


def _pad(docs, T=None, pad=1):
    T = T or max(len(d) for d in docs)
    ids = np.full((len(docs), T), pad, dtype=np.int64)
    for i, d in enumerate(docs):
        ids[i, :len(d)] = d
    return ids, np.array([len(d) for d in docs], dtype=np.int32)


def _assert_parity(got, want, cc_min=0.99, rel_l2_max=REL_L2_MAX):
    m = R.parity_metrics(got, want)
    assert np.isfinite(got).all()
    assert m["min_cosine"] >= COS_MIN, m
    assert m["rel_l2"] <= rel_l2_max, m
    if "min_centred_cosine" in m:
        assert m["min_centred_cosine"] >= cc_min, m
    return m


@pytest.fixture(scope="module")
def r4():
    """Reference-deployed shape (L=4, E=800, H=2400, V=60000), seed-1234 random init, on the GPU + its oracle."""
    from code_intelligence_b200 import IssueEncoder
    torch.set_num_threads(os.cpu_count())
    ref = R.make_encoder(1234)
    emb, layers = ref.export_weights()
    enc = IssueEncoder().load_weights(emb, layers)
    yield enc, ref
    enc.close()


def _small_from_golden(golden_dir, name):
    from code_intelligence_b200 import IssueEncoder
    z = np.load(os.path.join(golden_dir, name))
    n_layers, emb_sz, n_hid, vocab, seed = [int(x) for x in z["cfg"]]
    layers = [dict(w_ih=z[f"l{l}_w_ih"], w_hh=z[f"l{l}_w_hh"], b_ih=z[f"l{l}_b_ih"], b_hh=z[f"l{l}_b_hh"])
              for l in range(n_layers)]
    enc = IssueEncoder(n_layers, emb_sz, n_hid, vocab).load_weights(z["emb"], layers)
    return z, enc


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,act", [(128, 16, 64, 0), (128, 80, 128, 0), (200, 240, 64, 0), (256, 480, 192, 0),
                                        (300, 600, 1600, 1), (1000, 250, 600, 2), (2048, 9600, 832, 0),
                                        (1024, 3200, 2432, 0)])
def test_tcgen05_gemm_vs_torch(M, N, K, act):
    from code_intelligence_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K), dtype=np.float32)
    b = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N, dtype=np.float32)
    d = np.zeros((M, N), dtype=np.float32)
    _lib.check(lib.ie_debug_gemm(a.ctypes.data, b.ctypes.data, bias.ctypes.data, M, N, K, act, d.ctypes.data, 0))
    ref = torch.from_numpy(a).bfloat16().double() @ torch.from_numpy(b).bfloat16().double().T + torch.from_numpy(bias).double()
    if act == 1:
        ref = ref.clamp_min(0)
    if act == 2:
        ref = torch.sigmoid(ref)
    # operands are identical (bf16-rounded); only the f32 accumulation order differs
    np.testing.assert_allclose(d, ref.numpy(), atol=2e-4 * np.sqrt(K / 64), rtol=1e-5)


# ------------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", ["encoder_tiny.npz", "encoder_pad_dims.npz"])
def test_golden_small(golden_dir, name):
    z, enc = _small_from_golden(golden_dir, name)
    got = enc.encode_ids(z["ids"], z["lengths"])
    assert got.shape == z["expected"].shape and got.dtype == np.float32
    _assert_parity(got, z["expected"], rel_l2_max=REL_L2_MAX if float(z["scale"]) == 1.0 else REL_L2_MAX_SCALED)
    enc.close()


@pytest.mark.parametrize("name,cc", [("encoder_r4.npz", 0.99), ("encoder_r4_varlen.npz", 0.999)])
def test_golden_r4(golden_dir, r4, name, cc):
    """encoder_r4.npz is BASELINE.json configs[0] (32 issues, seq_len 128) pushed through the GPU path."""
    enc, _ = r4
    z = np.load(os.path.join(golden_dir, name))
    scale = float(z["scale"])
    if scale != 1.0:
        from code_intelligence_b200 import IssueEncoder
        ref = R.make_encoder(1234, scale=scale)
        emb, layers = ref.export_weights()
        enc = IssueEncoder().load_weights(emb, layers)
    got = enc.encode_ids(z["ids"], z["lengths"])
    m = _assert_parity(got, z["expected"], cc_min=cc, rel_l2_max=REL_L2_MAX if scale == 1.0 else REL_L2_MAX_SCALED)
    print(name, m)
    # negative control: the same outputs against the expected vectors of *other* issues must fail the extra gates
    neg = R.parity_metrics(got, np.roll(z["expected"], 1, axis=0))
    assert neg["rel_l2"] > 2 * REL_L2_MAX
    if scale != 1.0:
        enc.close()


def test_golden_n3(golden_dir):
    """North-star wording: 3-layer AWD-LSTM (L=3, E=800, H=2400)."""
    from code_intelligence_b200 import IssueEncoder
    z = np.load(os.path.join(golden_dir, "encoder_n3.npz"))
    ref = R.make_encoder(1234, n_layers=3)
    emb, layers = ref.export_weights()
    enc = IssueEncoder(n_layers=3).load_weights(emb, layers)
    _assert_parity(enc.encode_ids(z["ids"], z["lengths"]), z["expected"])
    enc.close()


# ------------------------------------------------------------------------------------------------ live oracle
def test_r4_vs_oracle_varlen_batch(r4):
    enc, ref = r4
    docs = R.synthetic_ids(40, 72, seed=77, min_len=1)
    ids, lengths = _pad(docs, 72)
    got = enc.encode_ids(ids, lengths)
    want = R.encode_padded(ref, ids, lengths)
    m = _assert_parity(got, want)
    print("varlen", m)
    # permuted ids (negative control) must NOT pass
    wrong = enc.encode_ids(np.roll(ids, 1, axis=0), np.roll(lengths, 1))
    assert R.parity_metrics(wrong, want)["rel_l2"] > 2 * REL_L2_MAX


def test_bulk_equals_single_bit_exact(r4):
    """The reference's own invariant (04b_Inference-Batch.ipynb:369, atol 1e-5) holds exactly here: a row's result
    does not depend on its batch mates, on T, or on the pad token."""
    enc, _ = r4
    docs = R.synthetic_ids(9, 40, seed=5, min_len=1)
    bulk = enc.encode_id_list(docs, bs=4)
    single = np.concatenate([enc.encode_ids(d[None, :]) for d in docs])
    np.testing.assert_array_equal(bulk, single)
    ids, lengths = _pad(docs, 40)
    ids2, _ = _pad(docs, 57, pad=7)
    np.testing.assert_array_equal(enc.encode_ids(ids, lengths), enc.encode_ids(ids2, lengths))
    np.testing.assert_array_equal(enc.encode_ids(ids, lengths), bulk)


def test_raw_features_and_pooling_consistency(r4):
    enc, ref = r4
    docs = R.synthetic_ids(3, 33, seed=8)
    ids, lengths = _pad(docs)
    raw = enc.raw_features(ids)
    assert raw.shape == (3, 33, 800) and raw.dtype == np.float32
    want = ref(torch.as_tensor(ids)).numpy()
    assert np.abs(raw - want).max() < 2e-4 and np.linalg.norm(raw - want) / np.linalg.norm(want) < REL_L2_MAX
    pooled = enc.encode_ids(ids, lengths)
    np.testing.assert_allclose(pooled, np.concatenate([raw.mean(1), raw.max(1), raw[:, -1]], axis=1), atol=1e-6)


def test_edge_cases_and_errors(r4):
    enc, ref = r4
    one = enc.encode_ids(np.array([[2]], dtype=np.int64))                      # B=1, T=1
    want = R.encode_single(ref, np.array([2]))
    _assert_parity(one, want)
    np.testing.assert_allclose(one[0, :800], one[0, 800:1600])                 # mean == max == last for T=1
    docs = R.synthetic_ids(900, 6, seed=3, min_len=2)                          # B > IE_MAX_BATCH (768) is sliced;
    ids, lengths = _pad(docs)                                                  # 257..512 / 513..768 rows ride one launch
    got = enc.encode_ids(ids, lengths)                                         # (three different kernel paths,
    np.testing.assert_array_equal(got[256], enc.encode_ids(ids[256:257], lengths[256:257])[0])   # identical bits)
    np.testing.assert_array_equal(got[:257], enc.encode_ids(ids[:257], lengths[:257]))
    np.testing.assert_array_equal(got[:600], enc.encode_ids(ids[:600], lengths[:600]))
    np.testing.assert_array_equal(got[768:], enc.encode_ids(ids[768:], lengths[768:]))
    with pytest.raises(ValueError):
        enc.encode_ids(ids[:2], np.array([7, 1], dtype=np.int32))              # length > T
    with pytest.raises(ValueError):
        enc.encode_ids(ids[:2], np.array([0, 1], dtype=np.int32))              # length < 1
    bad = ids[:2].copy()
    bad[1, 3] = 60000
    with pytest.raises(ValueError):
        enc.encode_ids(bad, lengths[:2])                                       # token id outside the vocab
    assert np.isfinite(enc.encode_ids(ids[:2], lengths[:2])).all()            # handle still usable


def test_rotating_schedule_kernel_matches_default(monkeypatch):
    """Opt-in IE_ROT kernel (csrc/lstm_rot.cu: up to five batches per launch, work items rotating over the CTA pairs):
    same MMA tile shapes and K order per (row, unit) as the default kernels => identical bits, pooled and raw."""
    from code_intelligence_b200 import IssueEncoder
    n_layers, emb_sz, n_hid, vocab = 3, 96, 200, 500
    emb, layers = R.make_encoder(7, vocab, emb_sz, n_hid, n_layers).export_weights()
    monkeypatch.delenv("IE_ROT", raising=False)
    base = IssueEncoder(n_layers, emb_sz, n_hid, vocab).load_weights(emb, layers)
    monkeypatch.setenv("IE_ROT", "2")              # read at handle creation; 2 = also route 257..768 rows through it
    rot = IssueEncoder(n_layers, emb_sz, n_hid, vocab).load_weights(emb, layers)
    monkeypatch.delenv("IE_ROT")
    assert base.max_batch == 768 and rot.max_batch == 1280
    for B, T in ((300, 19), (700, 23), (1100, 17), (1280, 9)):
        docs = R.synthetic_ids(B, T, seed=B + T, vocab_sz=vocab, min_len=1)
        ids, lengths = _pad(docs, T)
        want = base.encode_ids(ids, lengths)
        got = rot.encode_ids(ids, lengths)
        np.testing.assert_array_equal(got, want)
        if B <= 768:
            np.testing.assert_array_equal(rot.raw_features(ids), base.raw_features(ids))
    base.close()
    rot.close()


@pytest.mark.skipif(os.environ.get("IE_TEST_EXPERIMENTAL") != "1",
                    reason="opt-in development knobs not yet validated at full size; run with IE_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("knobs", [{"IE_EMB_PROJ": "1"}, {"IE_ROT": "2", "IE_EMB_PROJ": "1"}, {"IE_POOL_RAW": "1"},
                                   {"IE_ROT": "2", "IE_POOL_RAW": "1"}, {"IE_ROT": "2", "IE_ROT_BATCHES": "6"},
                                   {"IE_ROT": "2", "IE_EMB_PROJ": "1", "IE_POOL_RAW": "1", "IE_ROT_BATCHES": "8"},
                                   {"IE_ROT": "2", "IE_ROT_VARIANT": "1"}, {"IE_ROT": "2", "IE_ROT_VARIANT": "2"},
                                   {"IE_ROT": "2", "IE_ROT_VARIANT": "3", "IE_ROT_BATCHES": "6"}])
def test_experimental_knobs_match_default(knobs, monkeypatch):
    """DESIGN.md section 4 "Development knobs": every opt-in path must reproduce the default kernels bit for bit
    (per-token projection table = the same GEMM on the same operands; pooling from the f32 hidden states = the same
    sequential sums; more batches per launch = the same per-row arithmetic)."""
    from code_intelligence_b200 import IssueEncoder
    n_layers, emb_sz, n_hid, vocab = 3, 96, 200, 500
    emb, layers = R.make_encoder(7, vocab, emb_sz, n_hid, n_layers).export_weights()
    for k in ("IE_ROT", "IE_ROT_BATCHES", "IE_ROT_VARIANT", "IE_EMB_PROJ", "IE_POOL_RAW"):
        monkeypatch.delenv(k, raising=False)
    base = IssueEncoder(n_layers, emb_sz, n_hid, vocab).load_weights(emb, layers)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    exp = IssueEncoder(n_layers, emb_sz, n_hid, vocab).load_weights(emb, layers)
    for k in knobs:
        monkeypatch.delenv(k)
    for B, T in ((300, 19), (700, 23), (exp.max_batch, 11)):
        docs = R.synthetic_ids(B, T, seed=B + T, vocab_sz=vocab, min_len=1)
        ids, lengths = _pad(docs, T)
        np.testing.assert_array_equal(exp.encode_ids(ids, lengths), base.encode_ids(ids, lengths))
        if B <= 768:
            np.testing.assert_array_equal(exp.raw_features(ids), base.raw_features(ids))
    base.close()
    exp.close()


def test_full_size_batch_properties(r4):
    """BASELINE.json configs[1] shape (batch 256, seq_len 512): size-independent properties + oracle on a slice."""
    enc, ref = r4
    docs = R.synthetic_ids(256, 512, seed=99)
    ids = np.stack(docs)
    lengths = np.full(256, 512, dtype=np.int32)
    a = enc.encode_ids(ids, lengths)
    assert a.shape == (256, 2400) and np.isfinite(a).all()
    np.testing.assert_array_equal(a, enc.encode_ids(ids, lengths))                              # deterministic
    assert (a[:, 800:1600] >= a[:, :800] - 1e-6).all()                                          # max >= mean
    perm = np.random.default_rng(0).permutation(256)
    np.testing.assert_array_equal(enc.encode_ids(ids[perm], lengths)[np.argsort(perm)], a)      # row equivariance
    short = np.full(256, 100, dtype=np.int32)                                                   # prefix property
    np.testing.assert_array_equal(enc.encode_ids(ids, short), enc.encode_ids(ids[:, :100].copy(), short))
    ids2 = np.concatenate([ids, ids[::-1]])                                                     # two batches per launch
    b = enc.encode_ids(ids2, np.full(512, 512, dtype=np.int32))
    np.testing.assert_array_equal(b[:256], a)
    np.testing.assert_array_equal(b[256:], a[::-1])
    ids3 = np.concatenate([ids, ids[::-1], ids[perm]])                                          # three batches per launch
    c3 = enc.encode_ids(ids3, np.full(768, 512, dtype=np.int32))                                # (wide-tile kernel)
    np.testing.assert_array_equal(c3[:256], a)
    np.testing.assert_array_equal(c3[256:512], a[::-1])
    np.testing.assert_array_equal(c3[512:], a[perm])
    want = R.encode_padded(ref, ids[:6], lengths[:6])                                           # ~10 s of CPU
    m = _assert_parity(a[:6], want, cc_min=0.97)   # centring over 6 issues only: noisier than the batch-wide metric
    print("full-size slice", m)


def test_oom_halving_and_threads(r4, monkeypatch):
    """IE_ERR_OOM surfaces as RuntimeError, so the reference's batch-halving loop (py/code_intelligence/inference.py:214-223)
    keeps working; one handle may be driven from several host threads (calls are serialised inside the library)."""
    import threading
    enc, _ = r4
    docs = R.synthetic_ids(300, 24, seed=13, min_len=4)
    want = enc.encode_id_list(docs, bs=300, min_batches_rule=False)
    monkeypatch.setenv("IE_MAX_TOKENS", str(128 * 24))           # only B_pad = 128 fits: bs 300 -> 150 -> 75
    np.testing.assert_array_equal(enc.encode_id_list(docs, bs=300, min_batches_rule=False), want)
    monkeypatch.setenv("IE_MAX_TOKENS", "128")                    # nothing with T > 1 fits: the loop gives up at bs == 1
    with pytest.raises(Exception):
        enc.encode_id_list(docs, bs=4, min_batches_rule=False)
    monkeypatch.delenv("IE_MAX_TOKENS")
    outs = [None] * 4
    def work(i):
        outs[i] = enc.encode_id_list(docs[i * 10:(i + 1) * 10], bs=10, min_batches_rule=False)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    np.testing.assert_array_equal(np.concatenate(outs), want[:40])


# ------------------------------------------------------------------------------------------------ python surface
def test_inference_wrapper_surface(tmp_path):
    from code_intelligence_b200.inference import InferenceWrapper, text_endpoint_bytes
    ref = R.make_encoder(21, 300, 32, 48, 2)
    sd = {"encoder.weight": ref.encoder.weight.detach().numpy()}
    for l, rnn in enumerate(ref.rnns):
        sd[f"rnns.{l}.weight_hh_l0_raw"] = rnn.weight_hh_l0.detach().numpy()
        sd[f"rnns.{l}.module.weight_ih_l0"] = rnn.weight_ih_l0.detach().numpy()
        sd[f"rnns.{l}.module.weight_hh_l0"] = rnn.weight_hh_l0.detach().numpy()
        sd[f"rnns.{l}.module.bias_ih_l0"] = rnn.bias_ih_l0.detach().numpy()
        sd[f"rnns.{l}.module.bias_hh_l0"] = rnn.bias_hh_l0.detach().numpy()
    itos = ["xxunk", "xxpad", "xxbos", "xxfld", "xxmaj", "xxup", "xxrep", "xxwrep", "xxxfldtitle", "xxxfldbody"] + \
           [f"w{i}" for i in range(290)]
    np.savez(tmp_path / "enc.npz", itos=np.array(itos), **sd)
    w = InferenceWrapper(tmp_path, "enc.npz")
    text = w.process_dict({"title": "w1 w2 W3", "body": "w4 w5"})["text"]
    assert text.startswith("xxxfldtitle ") and " xxxfldbody " in text
    ids = w.numericalize_one(text)
    assert ids.shape[0] == 1 and int(ids[0, 0]) == 2
    pooled = w.get_pooled_features(text)
    assert isinstance(pooled, torch.Tensor) and tuple(pooled.shape) == (1, 96)
    raw = w.get_raw_features(text)
    assert tuple(raw.shape) == (1, ids.shape[1], 32)
    want = R.encode_single(ref, ids[0].numpy())
    _assert_parity(pooled.detach().cpu().numpy(), want, rel_l2_max=REL_L2_MAX_SCALED)
    b = text_endpoint_bytes(w, "w1 w2 W3", "w4 w5")
    assert len(b) == 96 * 4
    np.testing.assert_array_equal(np.frombuffer(b, dtype="<f4"), pooled.numpy()[0])
    import pandas as pd
    df = pd.DataFrame({"title": ["w1", "w2 w3", "W9 w8"], "body": ["w4 w5 w6", "w7", "w1 w1 w1 w1 w1"]})
    embs = w.df_to_embedding(df)
    assert embs.shape == (3, 96) and embs.dtype == np.float32
    np.testing.assert_array_equal(embs[1:2], w.get_pooled_features(w.process_dict(df.iloc[1].to_dict())["text"]).numpy())
    np.testing.assert_array_equal(w.df_to_emb(df), embs)


# ------------------------------------------------------------------------------------------------ MLP head
@pytest.mark.parametrize("tag", ["small", "prod"])
def test_mlp_head_vs_reference_fixture(golden_dir, tag):
    """mlp_ref_*.npz holds MLPWrapper.predict_probabilities outputs produced by the reference code itself."""
    from code_intelligence_b200.mlp import MLPHead, filter_predictions
    z = np.load(os.path.join(golden_dir, f"mlp_ref_{tag}.npz"))
    n = int(z["n_layers"])
    head = MLPHead([z[f"coef{i}"] for i in range(n)], [z[f"intercept{i}"] for i in range(n)])
    probs = head.predict_proba(z["X"])
    assert probs.shape == z["probs"].shape
    err = np.abs(probs - z["probs"])
    print(tag, "max abs prob diff", err.max())
    assert err.max() < 5e-3                      # bf16 operands, f32 accumulate
    # label-set agreement after thresholding (repo_specific_model.py:138-146), away from the decision boundary
    names = [f"l{i}" for i in range(probs.shape[1])]
    thr = {nm: 0.5 for nm in names}
    agree = 0
    for r in range(probs.shape[0]):
        a = set(filter_predictions(names, probs[r], thr))
        b = set(filter_predictions(names, z["probs"][r], thr))
        near = {nm for i, nm in enumerate(names) if abs(z["probs"][r, i] - 0.5) < 5e-3}
        assert (a ^ b) <= near
        agree += a == b
    assert agree >= 0.98 * probs.shape[0]
    head.close()


def test_mlp_wrapper_matches_sklearn():
    from sklearn.neural_network import MLPClassifier
    from code_intelligence_b200.mlp import MLPWrapper
    rng = np.random.default_rng(0)
    X = rng.random((60, 12)).astype(np.float32)
    y = rng.choice([0, 1], size=(60, 4))
    clf = MLPClassifier(random_state=1234, max_iter=30)
    w = MLPWrapper(clf=clf)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w.fit(X, y)
    Xt = rng.random((300, 12)).astype(np.float32)
    np.testing.assert_allclose(w.predict_probabilities(Xt), clf.predict_proba(Xt), atol=5e-3)
