"""Adaptive multi-round ANNCUR on the GPU (SURVEY.md section 8a row A8 -- NOT in the reference, parity unpinned; the CPU
statement of the same procedure is oracle.cur_oracle.adaptive_anncur, which the GPU test compares against).

Per query q, rounds t = 1..T with a growing anchor-item set I_t (|I_t| = t * k_per_round):
    c_q = exact_rows[q, I_t]                       stands in for CE(q, I_t): a gather from the exact score matrix
    e_q = c_q . pinv(R_anc[:, I_t])                per-query least squares           anncur_adaptive_solve (whole batch)
    s_q = e_q . R_anc with I_t masked              approximate scores of all items   fused score + top-k on the packed R_anc
    I_{t+1} = I_t  U  top-k_per_round(s_q)
Round 1 uses ``first_anchors`` (shared by all queries).  The answer is the top-``top_k`` of I_T by EXACT score, so the last
round needs no solve: T rounds cost T - 1 calls."""
import torch

from . import engine


class AdaptiveIndex:
    """What stays fixed across rounds and batches: R_anc packed for the fused re-score (the anchor QUERIES' scores play the
    items' embeddings, K = k_q), and its item-major fp32 copy for the per-query gathers of the solve.  ``sharded`` = a
    ``ShardedIndex`` over this rank's item slice: the re-score then runs on every rank's slice with one exchange per round
    (all ranks end with the same candidates, so the replicated solve stays in step)."""

    def __init__(self, R_anc, precision="f32r", sharded=None, rank_budget=False):
        self.R = engine._f32(R_anc)
        self.k_q, self.N = self.R.shape
        self.Rt = engine.transpose(self.R)
        self.sharded = sharded
        self.rank_budget = rank_budget                         # owned form: ship k/P + 6 sigma + 8 candidates per shard, certified (2 GPUs: no gain)
        self.packed = None if sharded is not None else engine.PackedItems(self.R, precision)
        self._shared = {}

    def shared_for(self, first_anchors, rcond):
        """The once-per-(index, first anchors) part of the incremental solver (anncur_adaptive_prepare), cached."""
        key = (tuple(first_anchors.tolist()), float(rcond))
        if key not in self._shared:
            self._shared.clear()
            self._shared[key] = engine.AdaptiveShared(self.Rt, first_anchors, rcond)
        return self._shared[key]

    def topk(self, e, k, n_rows_total=None):
        if self.sharded is not None:
            if n_rows_total is not None:                       # e = this rank's block of query rows: the answer for the same rows
                from .sharded import suggest_local_k
                local_k = suggest_local_k(k, self.sharded.world_size) if self.rank_budget else None
                return self.sharded.search_owned_verified(e, n_rows_total, k, local_k)
            return self.sharded.search(e, k)
        return engine.score_topk(e, self.packed, k)


def adaptive_anncur(R_anc, exact_rows, first_anchors, n_rounds, k_per_round, top_k, rcond=1e-15, *, rescore="fused",
                    index=None, precision="f32r", solver="incremental", n_rows_total=None):
    """Returns (anchors [B x T*k_per_round] int64 in selection order, idx [B x top_k] int64, exact scores [B x top_k]).

    ``rescore="fused"`` (default): per round one ``anncur_adaptive_solve`` (e_b for the whole batch), the fused tensor-core
    score + top-(k_per_round + m) on the packed R_anc -- no B x N block -- and ``anncur_filter_excluded`` to drop the anchors.
    ``rescore="ffma"``: the round-1 path (one ``anncur_adaptive_round`` call: FFMA re-score of a B x N block + masked row top-k),
    also taken when k_per_round + m exceeds the fused kernel's largest k.  ``index`` = an ``AdaptiveIndex`` to reuse.
    ``solver="incremental"`` (default, with the fused re-score): the per-query Cholesky factor is carried across rounds
    (``anncur_adaptive_begin`` / ``anncur_adaptive_extend``: a round pays for its NEW anchors only); ``solver="full"``: every
    round re-solves from scratch (``anncur_adaptive_solve``), also taken when a round adds more than
    ANNCUR_ADAPTIVE_MAX_BLOCK anchors.
    ``n_rows_total`` (multi-GPU, with an item-sharded ``index``): ``exact_rows`` is THIS RANK'S block of a batch of
    n_rows_total queries (``ShardedIndex.row_block``) -- the per-query solves are split over the ranks by query, the
    re-score over the ranks by item, and one ``search_owned`` exchange per round (all-gather of the e blocks over NVLink,
    local top-k per shard, candidates to the row's owner) brings every rank the candidates of its own queries."""
    R = engine._f32(R_anc)
    X = engine._f32(exact_rows, device=R.device)
    B = X.shape[0]
    first = torch.as_tensor(first_anchors, dtype=torch.int64, device=R.device)
    assert first.dim() == 1 and first.numel() == k_per_round and top_k <= n_rounds * k_per_round
    anchors = first.unsqueeze(0).expand(B, -1).contiguous()
    fused = rescore == "fused" and n_rounds * k_per_round <= engine.MAX_K_FUSED and R.shape[1] >= 4 * n_rounds * k_per_round
    if fused and index is None and n_rounds > 1:
        index = AdaptiveIndex(R, precision)
    incremental = (fused and solver == "incremental" and n_rounds > 1 and k_per_round <= engine.ADAPTIVE_MAX_BLOCK
                   and (n_rounds - 1) * k_per_round <= R.shape[0])
    state, nxt = None, None
    if incremental:
        state = engine.AdaptiveState(index.shared_for(first, rcond), B, k_per_round, (n_rounds - 1) * k_per_round)
    for t in range(n_rounds - 1):
        m = anchors.shape[1]
        if fused:
            if incremental:                                                      # K8 incremental: only the new anchors are paid for
                e = state.begin(torch.gather(X, 1, anchors)) if t == 0 else state.extend(nxt, torch.gather(X, 1, nxt))
            else:                                                                # K8a: per-query re-solve from scratch
                e = engine.adaptive_solve(R, anchors, torch.gather(X, 1, anchors), rcond, Rt=index.Rt)
            if index.sharded is None:                                            # K3+K4 on the packed R_anc, anchors masked: one call
                _, nxt = engine.score_topk_excluding(e, index.packed, k_per_round, anchors)
            else:
                cv, ci = index.topk(e, k_per_round + m, n_rows_total)            # sharded: the exchange carries k + m candidates
                _, nxt = engine.filter_excluded(cv, ci, anchors, k_per_round)    # the anchors leave the candidate lists
        else:
            c = torch.gather(X, 1, anchors)                                      # exact scores of the anchors so far
            nxt, _ = engine.adaptive_round(R, anchors, c, k_per_round, rcond)    # K8: re-solve + masked re-score + pick
        anchors = torch.cat([anchors, nxt], dim=1)
    ex = torch.gather(X, 1, anchors)
    vals, idx = engine.merge_topk(ex, anchors, top_k)                            # K9: best top_k of the anchors, ties -> lower index
    return anchors, idx, vals
