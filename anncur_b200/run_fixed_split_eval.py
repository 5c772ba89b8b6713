"""Command line of the reference's fixed-split retrieval eval, every method on the fused score + top-k kernel:

    python -m anncur_b200.run_fixed_split_eval --data_name yugioh --eval_method cur --res_dir out \
        --test_data_file .../test.pkl --train_data_file .../train.pkl --n_seeds 2 --misc nm_train=500

Same flags as eval/run_retrieval_eval_wrt_exact_crossenc_w_fixed_train_test_splits.py:515-542.  ``fixed_anc_ent`` and
``fixed_anc_ent_cur`` read the reference's entity-to-entity dump (--e2e_fname, --n_fixed_anc_ent); ``bienc`` and ``tfidf``
are served from PRECOMPUTED embeddings (--ent_embed_file / --ment_embed_file: the BERT / TF-IDF models that produce them
are outside this engine, everything after ``mention_embeds @ label_embeds.T`` is not), same seeds loop as its ``run`` (:446-500: seed = 0..n_seeds-1), same result
file ``{res_dir}/method={eval_method}_{misc}.json``.  ``--k_i`` / ``--k_r`` optionally restrict the grids (the
reference always sweeps all 41 x 43 points, which takes hours on the CPU; the anchor draws of skipped points are still
replayed, so every evaluated point is the reference's)."""
import argparse
import logging
import sys

from . import data_formats as F

LOGGER = logging.getLogger(__name__)


def build_parser():
    p = argparse.ArgumentParser(description="Retrieval eval wrt exact cross-encoder scores on a fixed train/test split (B200 engine)")
    p.add_argument("--data_name", type=str, default="", help="Dataset name (recorded in the result file)")
    p.add_argument("--eval_method", type=str, default="cur", choices=["cur", "bienc", "fixed_anc_ent", "fixed_anc_ent_cur", "tfidf"])
    p.add_argument("--res_dir", type=str, required=True, help="Result directory")
    p.add_argument("--test_data_file", type=str, required=True, help="Test data file")
    p.add_argument("--train_data_file", type=str, default="", help="Training data file. Used for method=cur")
    p.add_argument("--n_seeds", type=int, default=1)
    p.add_argument("--bi_model_file", type=str, default="")
    p.add_argument("--batch_size", type=int, default=50)
    p.add_argument("--e2e_fname", type=str, default="")
    p.add_argument("--n_fixed_anc_ent", type=int, default=0)
    p.add_argument("--mention_file", type=str, default="")
    p.add_argument("--entity_file", type=str, default="")
    p.add_argument("--mode", type=str, choices=["eval", "plot", "eval_n_plot"], default="eval")
    p.add_argument("--misc", type=str, default="", help="Misc suffix")
    p.add_argument("--use_wandb", type=int, default=0, choices=[0, 1])
    p.add_argument("--precision", type=str, default="f32r", choices=["f32r", "f32x3", "bf16", "f32"])
    p.add_argument("--ent_embed_file", type=str, default="", help="bienc / tfidf: precomputed entity embeddings (N x d; .npy / .npz / torch)")
    p.add_argument("--ment_embed_file", type=str, default="", help="bienc / tfidf: precomputed mention embeddings (.npy / .npz / torch)")
    p.add_argument("--k_i", type=int, nargs="*", default=None, help="restrict the anchor-item grid to these values")
    p.add_argument("--k_r", type=int, nargs="*", default=None, help="restrict the retrieved-k grid to these values")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    logging.basicConfig(stream=sys.stderr, level=logging.INFO, format="%(asctime)s - %(levelname)s - %(name)s - %(message)s ")
    if not args.train_data_file:
        raise SystemExit("--train_data_file is required (the reference reads it for every method: result keys carry its row count)")
    if args.eval_method != "cur" and args.n_seeds != 1:                 # reference :452-453
        raise SystemExit(f"n_seed = {args.n_seeds} only allowed for eval_method = cur")
    if args.eval_method in ("fixed_anc_ent", "fixed_anc_ent_cur") and not (args.e2e_fname and args.n_fixed_anc_ent > 0):
        raise SystemExit(f"eval_method={args.eval_method} needs --e2e_fname and --n_fixed_anc_ent")
    if args.eval_method in ("bienc", "tfidf") and not (args.ent_embed_file and args.ment_embed_file):
        raise SystemExit(f"eval_method={args.eval_method}: pass --ent_embed_file and --ment_embed_file (precomputed embeddings; "
                         "the BERT / TF-IDF models of --bi_model_file / --mention_file / --entity_file are outside this engine)")
    if args.mode != "eval":
        LOGGER.info("plotting is the reference's job (utils/plot_emnlp_*.py read the JSON written here); running eval only")
    eval_res, retvr_params = {}, {}
    embeds = {"ent_embed_file": args.ent_embed_file, "ment_embed_file": args.ment_embed_file}
    for seed in range(args.n_seeds):                               # reference :470-493
        LOGGER.info(f"seed {seed}: evaluating method={args.eval_method} on {args.test_data_file}")
        eval_res[seed], retvr_params = F.run_eval_method(
            args.eval_method, args.test_data_file, args.train_data_file,
            bienc_args=dict(embeds, bi_model_file=args.bi_model_file, batch_size=args.batch_size),
            cur_args={"seed": seed},
            fixed_anc_ent_args={"e2e_fname": args.e2e_fname, "n_fixed_anc_ent": args.n_fixed_anc_ent},
            tfidf_args=dict(embeds, mention_file=args.mention_file, entity_file=args.entity_file),
            precision=args.precision, n_ent_anchors_vals=args.k_i, top_k_retr_vals=args.k_r)
    arg_dict = {k: v for k, v in vars(args).items() if k not in ("precision", "k_i", "k_r", "ent_embed_file", "ment_embed_file")}
    res_file = F.write_result_json(args.res_dir, args.eval_method, args.misc, eval_res, arg_dict, retvr_params)
    LOGGER.info(f"wrote {res_file}")
    return res_file


if __name__ == "__main__":
    main()
