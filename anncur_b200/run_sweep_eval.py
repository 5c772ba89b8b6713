"""Random-anchor sweep of the CUR approximation (the reference's matrix-approximation eval) as ONE tool run on the GPU:

    python -m anncur_b200.run_sweep_eval --m2e_file scores.pkl --res_dir out [--n_seeds 1] [--rank]
    python -m anncur_b200.run_sweep_eval --synthetic 10000 1000000 --res_dir out --methods cur --rank      # BASELINE configs[4]

For every (k_q, k_i) of the grids of eval/run_retrieval_eval_wrt_exact_crossenc.py:227-233 and every method in
{cur, cur_oracle}: seed-averaged Top-k-Recall of retrieve-and-rerank (top_k = 10 @ k_retvr = 500, the reference's
active setting :237-238) and the reconstruction errors ``approx_error`` / ``approx_error_relative`` for anchor / non-anchor /
all rows (:146-154), written as ``{res_dir}/nm=<n>_ne=<N>_s=<seeds>/retrieval_wrt_exact_crossenc.json`` in the layout the
reference's plot() reads (:373-376, :416-420).  ``--rank`` adds np.linalg.matrix_rank of the score matrix
(eval/compute_m2e_matrix_ranks.py:44-53): full Jacobi SVD when the short side is <= 2048, randomised subspace iteration
(engine.matrix_rank_large) beyond.  Plotting stays with the reference (its plot() needs only the JSON)."""
import argparse
import json
import logging
import math
import sys
import time
from pathlib import Path

import torch

from . import engine
from .eval_retrieval import run_sweep, sweep_grids

LOGGER = logging.getLogger(__name__)


def synthetic_matrix(n, N, device, rank=64, noise=0.05, seed=0, chunk=100_000):
    """A = X Y^T / sqrt(r) + noise G, built in column chunks directly on the GPU."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    X = torch.randn((n, rank), generator=g, device=device)
    A = torch.empty((n, N), device=device)
    for a in range(0, N, chunk):
        b = min(a + chunk, N)
        Y = torch.randn((b - a, rank), generator=g, device=device)
        A[:, a:b] = X @ Y.t() / math.sqrt(rank)
        A[:, a:b] += noise * torch.randn((n, b - a), generator=g, device=device)
    return A


def matrix_rank_any(A):
    return engine.matrix_rank(A) if min(A.shape) <= 2048 else engine.matrix_rank_large(A)


def build_parser():
    p = argparse.ArgumentParser(description="CUR random-anchor sweep (recall + reconstruction error grid) on the B200 engine")
    p.add_argument("--m2e_file", type=str, default="", help="score-matrix pickle (reference schema)")
    p.add_argument("--synthetic", type=int, nargs=2, metavar=("N_QUERIES", "N_ITEMS"), default=None)
    p.add_argument("--res_dir", type=str, required=True)
    p.add_argument("--n_seeds", type=int, default=1)
    p.add_argument("--methods", type=str, nargs="*", default=["cur", "cur_oracle"])
    p.add_argument("--k_q", type=int, nargs="*", default=None, help="restrict the anchor-query grid")
    p.add_argument("--k_i", type=int, nargs="*", default=None, help="restrict the anchor-item grid")
    p.add_argument("--top_k", type=int, nargs="*", default=[10])
    p.add_argument("--k_r", type=int, nargs="*", default=[500])
    p.add_argument("--precision", type=str, default="f32r", choices=["f32r", "f32x3", "bf16", "f32"])
    p.add_argument("--rank", action="store_true", help="also report np.linalg.matrix_rank of the score matrix")
    p.add_argument("--misc", type=str, default="")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    logging.basicConfig(stream=sys.stderr, level=logging.INFO, format="%(asctime)s - %(levelname)s - %(name)s - %(message)s ")
    engine.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    if args.synthetic:
        A = synthetic_matrix(args.synthetic[0], args.synthetic[1], dev)
    elif args.m2e_file:
        from .data_formats import load_m2e_pickle
        A = load_m2e_pickle(args.m2e_file)["ment_to_ent_scores"].to(dev)
    else:
        raise SystemExit("pass --m2e_file or --synthetic N_QUERIES N_ITEMS")
    n, N = A.shape
    g_m, g_e = sweep_grids(n, N)
    k_q_vals = [v for v in g_m if args.k_q is None or v in set(args.k_q)]
    k_i_vals = [v for v in g_e if args.k_i is None or v in set(args.k_i)]
    timing = {"points": []}

    def progress(method, top_k, k_r, k_q, k_i, ans):
        torch.cuda.synchronize()
        now = time.perf_counter()
        timing["points"].append({"method": method, "k_q": k_q, "k_i": k_i, "s": now - timing["t"]})
        timing["t"] = now
        LOGGER.info(f"{method} k_q={k_q} k_i={k_i}: recall@{top_k} {ans['all']['exact_vs_reranked_approx_retvr~common_frac_mean']:.4f} "
                    f"approx_error_relative {ans['all']['approx_error_relative']:.4g} ({timing['points'][-1]['s']:.2f} s)")

    torch.cuda.synchronize()
    t0 = timing["t"] = time.perf_counter()
    res = run_sweep(A, n_seeds=args.n_seeds, eval_methods=args.methods, top_k_vals=args.top_k, top_k_retr_vals=args.k_r,
                    n_ment_anchors_vals=k_q_vals, n_ent_anchors_vals=k_i_vals, precision=args.precision, progress=progress)
    torch.cuda.synchronize()
    sweep_s = time.perf_counter() - t0
    res["other_args"]["arg_dict"] = {k: v for k, v in vars(args).items()}
    res["other_args"]["timing"] = {"sweep_s": sweep_s, "grid_points": len(timing["points"]), "per_point": timing["points"]}
    if args.rank:
        t0 = time.perf_counter()
        res["other_args"]["matrix_rank"] = int(matrix_rank_any(A))
        torch.cuda.synchronize()
        res["other_args"]["timing"]["matrix_rank_s"] = time.perf_counter() - t0
        LOGGER.info(f"Shape of matrix = {tuple(A.shape)}  Rank of matrix = {res['other_args']['matrix_rank']}")
    res_dir = f"{args.res_dir}/nm={n}_ne={N}_s={args.n_seeds}{args.misc}"
    Path(res_dir).mkdir(exist_ok=True, parents=True)
    res_fname = f"{res_dir}/retrieval_wrt_exact_crossenc.json"
    with open(res_fname, "w") as fout:
        json.dump(res, fout, indent=4)
    LOGGER.info(f"wrote {res_fname}: {len(timing['points'])} grid points in {sweep_s:.1f} s")
    return res_fname


if __name__ == "__main__":
    main()
