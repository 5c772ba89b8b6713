"""Adaptive multi-round ANNCUR on the GPU (SURVEY.md section 8a row A8 -- NOT in the reference, parity unpinned; the CPU
statement of the same procedure is oracle.cur_oracle.adaptive_anncur, which the GPU test compares against).

Per query q, rounds t = 1..T with a growing anchor-item set I_t (|I_t| = t * k_per_round):
    c_q = exact_rows[q, I_t]                       stands in for CE(q, I_t): a gather from the exact score matrix
    e_q = c_q . pinv(R_anc[:, I_t])                per-query least squares        } one anncur_adaptive_round call
    s_q = e_q . R_anc with I_t masked              approximate scores of all items } for the whole batch
    I_{t+1} = I_t  U  top-k_per_round(s_q)
Round 1 uses ``first_anchors`` (shared by all queries).  The answer is the top-``top_k`` of I_T by EXACT score, so the last
round needs no solve: T rounds cost T - 1 calls."""
import torch

from . import engine


def adaptive_anncur(R_anc, exact_rows, first_anchors, n_rounds, k_per_round, top_k, rcond=1e-15):
    """Returns (anchors [B x T*k_per_round] int64 in selection order, idx [B x top_k] int64, exact scores [B x top_k])."""
    R = engine._f32(R_anc)
    X = engine._f32(exact_rows, device=R.device)
    B = X.shape[0]
    first = torch.as_tensor(first_anchors, dtype=torch.int64, device=R.device)
    assert first.dim() == 1 and first.numel() == k_per_round and top_k <= n_rounds * k_per_round
    anchors = first.unsqueeze(0).expand(B, -1).contiguous()
    for t in range(n_rounds - 1):
        c = torch.gather(X, 1, anchors)                                          # exact scores of the anchors so far
        nxt, _ = engine.adaptive_round(R, anchors, c, k_per_round, rcond)        # K8: re-solve + masked re-score + pick
        anchors = torch.cat([anchors, nxt], dim=1)
    ex = torch.gather(X, 1, anchors)
    vals, idx = engine.merge_topk(ex, anchors, top_k)                            # K9: best top_k of the anchors, ties -> lower index
    return anchors, idx, vals
