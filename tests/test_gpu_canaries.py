"""GPU: out-of-bounds WRITE detection without compute-sanitizer (closed on this pool, profiles/r2f_sanitizer_memcheck.log).

Every output, workspace, packed index and solver state handed to the C ABI here is the interior of a larger allocation whose
margins hold a byte pattern; after the call the margins must be untouched.  Buffers are sized exactly as the library's own
*_bytes queries say, so a kernel that writes past its plan (a candidate list, a padded tile, a ragged tail) is caught."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PAD = 1 << 16           # bytes on either side
PATTERN = 0xA5


@pytest.fixture(scope="module")
def eng():
    from anncur_b200 import engine
    engine.require_cuda()
    return engine


class Guarded:
    """nbytes of device memory (256-byte aligned start) between two canary zones."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.full = torch.full((2 * PAD + self.nbytes + 256,), PATTERN, dtype=torch.uint8, device="cuda")
        self.off = PAD + (-(self.full.data_ptr() + PAD)) % 256
        self.bytes = self.full[self.off:self.off + self.nbytes]

    def view(self, dtype, shape):
        return self.bytes.view(dtype).view(shape)

    def check(self, what):
        torch.cuda.synchronize()
        lo, hi = self.full[:self.off], self.full[self.off + self.nbytes:]
        assert bool((lo == PATTERN).all()), f"{what}: write BEFORE the buffer"
        assert bool((hi == PATTERN).all()), f"{what}: write PAST the buffer ({int((hi != PATTERN).sum())} bytes)"


def _rand(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)).cuda()


@pytest.mark.parametrize("kind", ["f32r", "f32x3", "bf16"])
@pytest.mark.parametrize("B,K,N,k", [(130, 70, 12345, 100), (1, 500, 70001, 100), (257, 64, 999, 37), (33, 129, 40000, 1000),
                                     (4097, 500, 100003, 100)])
def test_fused_search_stays_inside_its_buffers(eng, kind, B, K, N, k):
    from anncur_b200 import _lib
    lib = _lib.load()
    E, Q = _rand((K, N), 1), _rand((B, K), 2)
    kd = eng.KINDS[kind]
    g_pack = Guarded(lib.anncur_packed_items_bytes(N, K, kd))
    scale = torch.ones(1, dtype=torch.float32, device="cuda")
    _lib.check(lib.anncur_pack_items(eng._ptr(E), eng._ld(E), N, K, kd, eng._ptr(g_pack.bytes), eng._ptr(scale), eng._stream()))
    g_pack.check(f"pack_items {kind}")
    g_ws = Guarded(lib.anncur_score_topk_workspace_bytes(B, N, K, k, kd))
    g_v, g_i = Guarded(4 * B * k), Guarded(8 * B * k)
    for _ in range(2):                                       # twice: the second call starts from a used workspace
        _lib.check(lib.anncur_score_topk(eng._ptr(Q), eng._ld(Q), B, eng._ptr(g_pack.bytes), eng._ptr(scale), N, K, kd, k, 0,
                                         eng._ptr(g_v.bytes), eng._ptr(g_i.bytes), eng._ptr(g_ws.bytes), g_ws.nbytes, eng._stream()))
    for g, what in ((g_ws, "workspace"), (g_v, "values"), (g_i, "indices"), (g_pack, "packed index")):
        g.check(f"score_topk {kind} B={B} K={K} N={N} k={k}: {what}")
    idx = g_i.view(torch.int64, (B, k))
    kk = min(k, N)
    assert int(idx[:, :kk].min()) >= 0 and int(idx[:, :kk].max()) < N
    if kind != "bf16":                                       # dense epilogue on the same planes
        g_d = Guarded(4 * B * N)
        g_dws = Guarded(lib.anncur_score_dense_workspace_bytes(B, N, K, kd))
        _lib.check(lib.anncur_score_dense(eng._ptr(Q), eng._ld(Q), B, eng._ptr(g_pack.bytes), eng._ptr(scale), N, K, kd,
                                          eng._ptr(g_d.bytes), N, eng._ptr(g_dws.bytes), g_dws.nbytes, eng._stream()))
        g_d.check("score_dense output")
        g_dws.check("score_dense workspace")
        ref = Q[: min(B, 64)].double() @ E.double()
        got = g_d.view(torch.float32, (B, N))[: min(B, 64)].double()
        assert float((got - ref).abs().max() / ref.abs().max()) < 1e-4


def test_adaptive_solver_stays_inside_its_state(eng):
    from anncur_b200 import _lib
    lib = _lib.load()
    k_q, N, B, s, n, rounds = 90, 3001, 37, 13, 29, 2       # ragged everything: sub-blocks of 29, odd leading dimensions
    R, X = _rand((k_q, N), 3), _rand((B, N), 4)
    Rt = eng.transpose(R)
    rng = np.random.default_rng(5)
    first = torch.from_numpy(np.sort(rng.choice(N, s, replace=False))).cuda()
    g_sh = Guarded(lib.anncur_adaptive_shared_bytes(k_q, N, s))
    g_pw = Guarded(lib.anncur_adaptive_prepare_workspace_bytes(k_q, N, s))
    _lib.check(lib.anncur_adaptive_prepare(eng._ptr(Rt), k_q, N, eng._ptr(first), s, 1e-15, eng._ptr(g_sh.bytes), g_sh.nbytes,
                                           eng._ptr(g_pw.bytes), g_pw.nbytes, eng._stream()))
    g_sh.check("adaptive_prepare shared blob")
    g_pw.check("adaptive_prepare workspace")
    m_max = s + rounds * n
    g_st = Guarded(lib.anncur_adaptive_state_bytes(B, k_q, s, n, m_max))
    g_e = Guarded(4 * B * k_q)
    c = torch.gather(X, 1, first.unsqueeze(0).expand(B, -1)).contiguous()
    _lib.check(lib.anncur_adaptive_begin(eng._ptr(Rt), k_q, N, eng._ptr(g_sh.bytes), s, eng._ptr(c), B, n, m_max, eng._ptr(g_e.bytes),
                                         eng._ptr(g_st.bytes), g_st.nbytes, eng._stream()))
    m_cur = s
    cur = first.unsqueeze(0).expand(B, -1)
    for t in range(rounds):
        new = torch.from_numpy(np.stack([rng.choice(N, n, replace=False) for _ in range(B)])).cuda()
        c_new = torch.gather(X, 1, new).contiguous()
        _lib.check(lib.anncur_adaptive_extend(eng._ptr(Rt), k_q, N, eng._ptr(g_sh.bytes), s, eng._ptr(new), eng._ptr(c_new), B, n, m_max,
                                              m_cur, 1e-15, eng._ptr(g_e.bytes), eng._ptr(g_st.bytes), g_st.nbytes, eng._stream()))
        m_cur += n
        g_st.check(f"adaptive_extend round {t}: state")
        g_e.check(f"adaptive_extend round {t}: e")
        g_sh.check(f"adaptive_extend round {t}: shared blob")
    assert bool(torch.isfinite(g_e.view(torch.float32, (B, k_q))).all())


def test_filter_select_and_merge_stay_inside_their_outputs(eng):
    from anncur_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(6)
    n, k_in, m, n_out = 301, 333, 77, 55
    ci = torch.from_numpy(np.stack([rng.permutation(5000)[:k_in] for _ in range(n)])).cuda()
    cv = torch.sort(torch.randn(n, k_in), dim=1, descending=True).values.cuda()
    ex = torch.from_numpy(np.stack([rng.permutation(5000)[:m] for _ in range(n)])).cuda()
    g_v, g_i = Guarded(4 * n * n_out), Guarded(8 * n * n_out)
    _lib.check(lib.anncur_filter_excluded(eng._ptr(cv), eng._ptr(ci), n, k_in, eng._ptr(ex), m, n_out, eng._ptr(g_v.bytes), eng._ptr(g_i.bytes),
                                          eng._stream()))
    g_v.check("filter_excluded values")
    g_i.check("filter_excluded indices")
    S = _rand((19, 70001), 7)
    for k in (1, 100, 2048):
        g_v, g_i = Guarded(4 * 19 * k), Guarded(8 * 19 * k)
        _lib.check(lib.anncur_topk_rows_f32(eng._ptr(S), eng._ld(S), 19, 70001, k, 0, eng._ptr(g_v.bytes), eng._ptr(g_i.bytes), eng._stream()))
        g_v.check(f"topk_rows k={k} values")
        g_i.check(f"topk_rows k={k} indices")
        assert torch.equal(g_v.view(torch.float32, (19, k)), torch.topk(S, k, dim=1).values)


@pytest.mark.parametrize("m,n", [(333, 97), (97, 333), (2000, 500), (64, 64)])
def test_pinv_stays_inside_its_buffers(eng, m, n):
    """Both pinv routes (normal equations chosen on the device, Jacobi forced through cond_out) and the FFMA GEMM."""
    from anncur_b200 import _lib
    lib = _lib.load()
    A = _rand((m, n), 11)
    for want_cond in (False, True):
        g_out, g_ws, g_cond = Guarded(4 * n * m), Guarded(lib.anncur_pinv_workspace_bytes(m, n)), Guarded(16)
        _lib.check(lib.anncur_pinv_f32(eng._ptr(A), m, n, eng._ld(A), 1e-15, eng._ptr(g_out.bytes), m,
                                       eng._ptr(g_cond.bytes) if want_cond else C.c_void_p(0), eng._ptr(g_ws.bytes), g_ws.nbytes, eng._stream()))
        for g, what in ((g_out, "output"), (g_ws, "workspace"), (g_cond, "cond")):
            g.check(f"pinv {m}x{n} (cond_out={want_cond}): {what}")
        P = g_out.view(torch.float32, (n, m)).double()
        assert float((A.double() @ P @ A.double() - A.double()).abs().max()) < 1e-3 * float(A.abs().max())
    B = _rand((n, 131), 12)
    g_c = Guarded(4 * m * 131)
    _lib.check(lib.anncur_gemm_f32(eng._ptr(A), eng._ld(A), eng._ptr(B), eng._ld(B), eng._ptr(g_c.bytes), 131, m, 131, n, eng._stream()))
    g_c.check("gemm_f32 output")


def test_rerank_overlap_and_key_merge_stay_inside_their_outputs(eng):
    from anncur_b200 import _lib
    lib = _lib.load()
    n, N, k_r, k = 23, 7001, 777, 100
    exact = _rand((n, N), 13)
    _, retr = eng.topk_rows(exact + 0.3 * _rand((n, N), 14), k_r)
    _, ex_i = eng.topk_rows(exact, k)
    ks = (C.c_int * 3)(1, 10, 100)
    g_i, g_v, g_c = Guarded(8 * n * k), Guarded(4 * n * k), Guarded(4 * n * 3)
    _lib.check(lib.anncur_rerank_overlap(eng._ptr(exact), eng._ld(exact), n, N, eng._ptr(retr), k_r, eng._ptr(ex_i), k, ks, 3,
                                         eng._ptr(g_i.bytes), eng._ptr(g_v.bytes), eng._ptr(g_c.bytes), eng._stream()))
    for g, what in ((g_i, "indices"), (g_v, "values"), (g_c, "overlap counts")):
        g.check(f"rerank_overlap: {what}")
    P = 5
    keys = torch.stack([eng.topk_to_keys(*eng.topk_rows(exact[:, p * 1400:(p + 1) * 1400].contiguous(), k, idx_offset=p * 1400)) for p in range(P)]).contiguous()
    g_mv, g_mi, g_ws = Guarded(4 * n * k), Guarded(8 * n * k), Guarded(lib.anncur_merge_topk_keys_workspace_bytes(n))
    _lib.check(lib.anncur_merge_topk_keys(eng._ptr(keys), P, n, k, k, eng._ptr(g_mv.bytes), eng._ptr(g_mi.bytes), eng._ptr(g_ws.bytes), g_ws.nbytes,
                                          eng._stream()))
    for g, what in ((g_mv, "values"), (g_mi, "indices"), (g_ws, "workspace")):
        g.check(f"merge_topk_keys: {what}")
    ref = torch.topk(exact[:, :7000], k, dim=1)
    assert torch.equal(g_mv.view(torch.float32, (n, k)), ref.values)


def test_masked_search_stays_inside_its_buffers(eng):
    from anncur_b200 import _lib
    lib = _lib.load()
    B, K, N, k, m = 131, 70, 30001, 125, 375
    E, Q = _rand((K, N), 21), _rand((B, K), 22)
    packed = eng.PackedItems(E, "f32r")
    excl = torch.from_numpy(np.stack([np.random.default_rng(r).choice(N, m, replace=False) for r in range(B)])).cuda()
    g_ws = Guarded(lib.anncur_score_topk_excluding_workspace_bytes(B, N, K, k, m, packed.kind))
    g_v, g_i = Guarded(4 * B * k), Guarded(8 * B * k)
    _lib.check(lib.anncur_score_topk_excluding(eng._ptr(Q), eng._ld(Q), B, eng._ptr(packed.buf), eng._ptr(packed.scale), N, K, packed.kind, k,
                                               eng._ptr(excl), m, 0, eng._ptr(g_v.bytes), eng._ptr(g_i.bytes), eng._ptr(g_ws.bytes), g_ws.nbytes,
                                               eng._stream()))
    for g, what in ((g_ws, "workspace"), (g_v, "values"), (g_i, "indices")):
        g.check(f"score_topk_excluding: {what}")
    idx = g_i.view(torch.int64, (B, k))
    assert not bool((idx.unsqueeze(2) == excl.unsqueeze(1)).any())
