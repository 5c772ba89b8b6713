"""CPU check of the error-bound arithmetic kind F32R rests on (anncur_b200/csrc/score_topk_umma.cu,
pack_queries_kernel / item_bound_slot_kernel): with operands scaled by a power of two into [0, 2^14) and rounded to fp16,
the one-pass score A = sum_i h(q'_i) h(e'_i) satisfies |S - A| <= b = f(K) * 2^-10 * ||q'|| * ||e'|| (f(K) = slot_factor, 1.05 at K = 500; norms with the
K * 2^-28 floor), where S = sum_i q'_i e'_i.  numpy's float16 is the same IEEE binary16 the tensor core consumes, so the
inequality can be checked without a GPU -- on random data, on subnormal-heavy data, and on operands whose rounding errors
all have the same sign (the case a statistical error model would miss).  The GPU side of the claim (accumulation order,
the slot arithmetic) is tests/test_gpu_fused.py::test_fused_f32r_bounds_enclose_the_fp32_score."""
import numpy as np
import pytest


def pow2_scale(maxabs):
    """pow2_scale_for of score_topk_umma.cu: maps [0, maxabs] into [0, 2^14)."""
    if not maxabs > 0:
        return 1.0
    return float(2.0 ** (14 - np.frexp(maxabs)[1]))


def bound_and_error(q, E):
    """q: (K,), E: (K, N) fp32.  Returns (|S - A| per item, b per item) in scaled units, all arithmetic in fp64."""
    q = q.astype(np.float64)
    E = E.astype(np.float64)
    K = q.shape[0]
    qs = q * pow2_scale(np.abs(q).max())
    Es = E * pow2_scale(np.abs(E).max())
    qh = qs.astype(np.float16).astype(np.float64)
    Eh = Es.astype(np.float16).astype(np.float64)
    S = qs @ Es
    A = qh @ Eh
    nq = np.sqrt((qs ** 2).sum() + K * 2.0 ** -28)
    ne = np.sqrt((Es ** 2).sum(axis=0) + K * 2.0 ** -28)
    # the slots are rounded UP to fp16 after the 2^-5 scaling; rounding up only enlarges b, so the plain product is the test
    b = slot_factor(K) * 2.0 ** -10 * nq * ne
    return np.abs(S - A), b


def slot_factor(K):
    """f32r_slot_factor of score_topk_umma.cu: 4 % flat + the K-proportional reserve for fp32 accumulation (tensor pipe:
    K/16 + 1 additions of <= 2^-22; re-scoring: K/32 + 8 roundings of 2^-24), relative to the 2^-10 rounding bound."""
    return 1.04 + (K / 16 + 1) * 2.0 ** -12 + (K / 32 + 8) * 2.0 ** -14


def test_slot_factor_reserve_covers_fp32_accumulation_at_every_allowed_k():
    """What is left of b after the fp16 rounding part (2^-10 + 2^-22) |q'| |e'| must cover a worst-case fp32 accumulation
    of the tensor pipe (2 ulp per TMEM accumulation, coherent) AND of the re-scoring kernel, for K up to the header's
    ANNCUR_MAX_K_DIM_F32R -- the K-independent 5 % of round 1 did not."""
    for K in (1, 50, 500, 512, 768, 2000, 4096, 8192):
        reserve = (slot_factor(K) - (1 + 2.0 ** -12)) * 2.0 ** -10
        tensor_acc = (K / 16 + 1) * 2.0 ** -22
        refine_acc = (K / 32 + 8) * 2.0 ** -24
        assert reserve >= tensor_acc + refine_acc + 0.03 * 2.0 ** -10, K
    assert abs(slot_factor(500) - 1.05) < 0.002          # the bench shapes keep the round-1 constant


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_bound_holds_on_random_operands(seed):
    rng = np.random.default_rng(seed)
    K, N = 500, 4000
    q = rng.standard_normal(K).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 6))
    E = (rng.standard_normal((K, N)) * 10.0 ** rng.uniform(-4, 4, (1, N))).astype(np.float32)
    err, b = bound_and_error(q, E)
    assert (err <= b).all()
    assert (err / b).max() < 0.5                    # random rounding errors cancel: far inside the bound


def test_bound_holds_when_most_elements_are_fp16_subnormal():
    rng = np.random.default_rng(3)
    K, N = 200, 3000
    q = np.full(K, 1e-7, np.float32)
    q[0] = 1.0                                       # one large element sets the scale, the rest fall below 2^-14 * 2^14
    E = (rng.standard_normal((K, N)) * 1e-8).astype(np.float32)
    E[0, :] = 1.0
    err, b = bound_and_error(q, E)
    assert (err <= b).all()


def test_bound_holds_for_coherent_worst_case_rounding():
    """x = 2^e (1 + 2^-11 + 2^-13) rounds UP by 3/4 of a half-ulp in fp16; with all signs equal the errors add up linearly
    in K, where a sqrt(K) model would fall short.  Also the mirror case just below a tie (rounds DOWN)."""
    rng = np.random.default_rng(4)
    K, N = 512, 2000
    for frac in (1.0 + 2.0 ** -11 + 2.0 ** -13, 1.0 + 2.0 ** -11 - 2.0 ** -13):
        q = (frac * 2.0 ** rng.integers(-3, 4, K)).astype(np.float32)
        E = (frac * 2.0 ** rng.integers(-3, 4, (K, N))).astype(np.float32)
        err, b = bound_and_error(q, E)
        assert (err <= b).all()
        assert (err / b).max() > 0.2                 # this construction really does use a good part of the bound
        # equal magnitudes: Cauchy-Schwarz is tight, the bound is used to 3/4 * 2 * 2^-11 / (f(K) * 2^-10) ~ 0.71
        q1 = np.full(K, frac, np.float32)
        E1 = np.full((K, 3), frac, np.float32)
        err, b = bound_and_error(q1, E1)
        assert (err <= b).all() and 0.6 < (err / b).max() < 0.8
