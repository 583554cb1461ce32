"""Output side of the bulk per-repo encode (py/code_intelligence/embeddings.py:77-118 and the notebook writers that
consume it).  The scraping / BigQuery / GCS parts of the reference are network glue and out of scope (SURVEY.md
section 2 row 4); what a drop-in has to reproduce is the shape of the result:

    {'features': embeddings[:, :1600],   # mean | max  (the 'last' third is dropped: EMB:116, RSM:182)
     'labels':   [...], 'nums': [...]}

and the array formats the callers store: HDF5 dataset ``issue_embeddings`` ``(N, 2400) <f4``
(Issue_Embeddings/notebooks/Get-GitHub-Issues.ipynb:964) when h5py is importable, ``.npy`` otherwise, and the dill
dictionary of ``IssuesLoader`` / ``RepoMLP.load_training_data``.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np

FEATURE_DIMS = 1600  # mean | max


def issues_to_features(inf_wrapper, issues: Sequence[dict], bs: int = 100) -> Dict[str, object]:
    """``get_all_issue_text`` from the already retrieved issues on (EMB:101-118): ``issues`` is a list of
    ``{'title', 'body', 'labels', 'num'}`` dicts; returns ``{'features': (N,1600) float32, 'labels', 'nums'}``."""
    import pandas as pd
    if not issues:
        raise ValueError("No issues retrieved")
    labels = [i['labels'] for i in issues]
    nums = [i['num'] for i in issues]
    df = pd.DataFrame.from_dict({'title': [i['title'] for i in issues], 'body': [i['body'] for i in issues]})
    features = inf_wrapper.df_to_embedding(df, bs=bs)
    assert len(features) == len(labels), 'Error you have mismatch b/w number of observations and labels.'
    return {'features': features[:, :FEATURE_DIMS], 'labels': labels, 'nums': nums}


def save_features(path: str, data: Dict[str, object]) -> None:
    """dill dump of the dictionary above (what IssuesLoader.save_issue_embeddings uploads)."""
    import dill as dpickle
    with open(path, 'wb') as f:
        dpickle.dump(data, f)


def save_embeddings(path: str, embeddings: np.ndarray) -> str:
    """``(N, 2400)`` little-endian float32: HDF5 dataset 'issue_embeddings' when h5py is available, else ``<path>.npy``."""
    emb = np.ascontiguousarray(embeddings, dtype='<f4')
    try:
        import h5py
        with h5py.File(path, 'w') as f:
            f.create_dataset('issue_embeddings', data=emb)
        return path
    except ImportError:
        out = path if path.endswith('.npy') else path + '.npy'
        np.save(out, emb)
        return out
