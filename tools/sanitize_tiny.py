"""Development aid: a tiny run of every kernel (persistent recurrent kernel, per-timestep fallback, both GEMMs, gather /
table / finalize, MLP head) for compute-sanitizer (memcheck / racecheck / synccheck), checked against the oracle.

    IE_SPIN_LIMIT_MS=600000 compute-sanitizer --tool memcheck python tools/sanitize_tiny.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from code_intelligence_b200 import IssueEncoder
from code_intelligence_b200.mlp import MLPHead
from oracle import awd_lstm_ref as R
from oracle import lstm_numpy as N

cfg = (2, 64, 128, 300)
ref = R.make_encoder(5, cfg[3], cfg[1], cfg[2], cfg[0], scale=2.0)
weights = ref.export_weights()
docs = R.synthetic_ids(300, 5, seed=1, vocab_sz=cfg[3], min_len=1)
T = 5
ids = np.full((len(docs), T), 1, dtype=np.int64)
for i, d in enumerate(docs):
    ids[i, :len(d)] = d
lengths = np.array([len(d) for d in docs], dtype=np.int32)
want = R.encode_padded(ref, ids, lengths)
for name, env in (("persistent (last layer fused)", {}), ("persistent, hoisted last layer", {"IE_FUSE_LAST": "0"}), ("fallback+gather", {"IE_SEQ": "0", "IE_EMB_PROJ": "0"}), ("chunked", {"IE_CHUNK_T": "2"})):
    os.environ.update(env)
    enc = IssueEncoder(*cfg, 1, 0).load_weights(*weights)
    for k in env:
        os.environ.pop(k)
    got = enc.encode_ids(ids, lengths)
    m = R.parity_metrics(got, want)
    print(name, "rel_l2 %.2e" % m["rel_l2"], "launches", enc.launch_count, flush=True)
    assert m["rel_l2"] < 1e-2
    enc.close()
rng = np.random.default_rng(0)
coefs = [rng.standard_normal((24, 32)).astype(np.float32) * 0.2, rng.standard_normal((32, 5)).astype(np.float32) * 0.2]
ints = [rng.standard_normal(32).astype(np.float32) * 0.1, rng.standard_normal(5).astype(np.float32) * 0.1]
X = rng.standard_normal((300, 24)).astype(np.float32)
head = MLPHead(coefs, ints)
assert np.abs(head.predict_proba(X) - N.mlp_forward(X, coefs, ints)).max() < 5e-3
head.close()
print("sanitize_tiny ok")
