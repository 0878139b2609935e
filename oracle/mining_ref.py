"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): Python restatement of the reference's ANN negative mining and
of a Lloyd k-means step.

``generate_negatives`` follows ``GenerateNegativePassaageID`` (ANCE/drivers/run_ann_data_gen.py:497-570) line by
line for one query: reciprocal rank of the positive over the full top-k list (:523-534), then the walk over the
selected candidates -- ``top_ann_pid[:negative_sample + 1]`` when SelectTopK, ``top_ann_pid[order]`` otherwise
(:536-541) -- skipping the positive (:551-554) and repeats (:556-557) until ``negative_sample`` ids are kept
(:559-563).  The reference shuffles with Python's ``random``; the order is an explicit argument here.
``kmeans_step`` is one Lloyd iteration (faiss.Kmeans semantics of :340-351: L2 assignment, mean update, ties to the
lowest index).  Pinning: ``generate_negatives`` is checked against the UNMODIFIED reference function (its AST node is
compiled on its own by oracle/make_golden.py gen_mining; fixture tests/golden/mining_tiny.npz); the k-means step has no
reference counterpart importable here (faiss) -- parity unpinned for that one.
"""
import numpy as np


def generate_negatives(top_rows, doc_pid, pos_pid, n_neg, order=None, n_sel=None):
    rr = 0.0
    rank = 0
    for idx in top_rows:
        rank += 1
        if idx >= 0 and doc_pid[idx] == pos_pid:
            rr = 1.0 / rank
            break
    if order is None:
        sel = top_rows[: (n_sel if n_sel is not None else n_neg + 1)]
    else:
        sel = [top_rows[j] for j in order]
    negs = []
    for idx in sel:
        if idx < 0:
            continue
        pid = doc_pid[idx]
        if pid == pos_pid:
            continue
        if pid in negs:
            continue
        if len(negs) >= n_neg:
            break
        negs.append(int(pid))
    return negs, rr


def kmeans_step(X, cent):
    X = X.astype(np.float64)
    c = cent.astype(np.float64)
    score = X @ c.T - 0.5 * (c * c).sum(1)[None, :]
    assign = score.argmax(1)
    new = c.copy()
    for g in range(c.shape[0]):
        m = assign == g
        if m.any():
            new[g] = X[m].mean(0)
    return new, assign
