"""Debug aid: one long single-issue encode on the tiny model, timed; run under `timeout`."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from code_intelligence_b200 import IssueEncoder
from oracle import awd_lstm_ref as R
T = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = (2, 64, 128, 300)
ref = R.make_encoder(9, cfg[3], cfg[1], cfg[2], cfg[0], scale=2.0)
enc = IssueEncoder(*cfg, 1, 0).load_weights(*ref.export_weights())
docs = np.stack(R.synthetic_ids(B, T, seed=6, vocab_sz=cfg[3]))
t0 = time.time()
try:
    got = enc.encode_ids(docs)
    print("T", T, "B", B, "ok %.3f s" % (time.time() - t0), "launches", enc.launch_count, "phases", enc.last_phase_ms(), flush=True)
    want = R.encode_padded(ref, docs[:1], [T])
    print("   rel_l2", R.parity_metrics(got[:1], want)["rel_l2"], flush=True)
except Exception as e:
    print("T", T, "B", B, "ERROR after %.3f s:" % (time.time() - t0), repr(e)[:300], flush=True)
