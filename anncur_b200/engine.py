"""Torch-tensor front end of the C ABI: allocates outputs / workspaces as CUDA tensors, passes raw
device pointers and the current CUDA stream, and returns CUDA tensors.  Everything here requires a
CUDA device -- there is no CPU path (the CPU restatement lives in oracle/ and is test-only)."""
import ctypes as C

import torch

from . import _lib
from ._lib import KIND_BF16, KIND_F32R, KIND_F32X3, MAX_K, MAX_K_DIM_F32R, MAX_K_FUSED  # noqa: F401

ADAPTIVE_MAX_BLOCK = 128      # ANNCUR_ADAPTIVE_MAX_BLOCK
KINDS = {"f32r": KIND_F32R, "f32x3": KIND_F32X3, "bf16": KIND_BF16}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("anncur_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32(t, device=None):
    """fp32 CUDA tensor with unit stride in the last dimension (row stride may be anything)."""
    require_cuda()
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    dev = device if device is not None else (t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    t = t.to(device=dev, dtype=torch.float32, non_blocking=True)
    if t.dim() >= 1 and t.shape[-1] > 1 and t.stride(-1) != 1:
        t = t.contiguous()
    if t.dim() == 2 and t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _ld(t):
    return t.stride(0) if (t.dim() == 2 and t.shape[0] > 1) else max(t.shape[-1], 1)


class _Workspace:
    """Grow-only per-device scratch buffers keyed by purpose (caller-owned memory for the ABI)."""

    def __init__(self):
        self._bufs = {}

    def get(self, key, nbytes, device):
        # one buffer per (purpose, device, stream): calls issued on different streams may run concurrently
        k = (key, device.index, torch.cuda.current_stream(device).cuda_stream)
        buf = self._bufs.get(k)
        if buf is None or buf.numel() < nbytes:
            buf = None
            self._bufs.pop(k, None)
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[k] = buf
        return buf

    def clear(self):
        self._bufs.clear()


WORKSPACE = _Workspace()


def launch_count():
    return int(_lib.load().anncur_kernel_launch_count())


def reset_launch_count():
    _lib.load().anncur_reset_kernel_launch_count()


def profile_enable(on=True):
    _lib.check(_lib.load().anncur_profile_enable(1 if on else 0))


def profile_read():
    """(summed fused-kernel milliseconds, launches) since the last read; waits for the recorded events."""
    ms, n = C.c_double(0.0), C.c_int(0)
    _lib.check(_lib.load().anncur_profile_read(C.byref(ms), C.byref(n)))
    return ms.value, n.value


# ---- K1 -------------------------------------------------------------------------------------------
def pinv(A, rcond=1e-15, return_cond=False):
    """pinv of an (m x n) fp32 matrix -> (n x m) fp32 on the GPU (eval/matrix_approx_zeshel.py:47,49).
    Full-rank, well-conditioned inputs take the fp64 normal-equations route (Gram + cooperative Cholesky + triangular solves),
    everything else -- rank-deficient, ill-conditioned, rcond > 1e-10, or ``return_cond=True`` (singular values wanted) -- the
    fp64 Jacobi SVD; the choice is made on the device."""
    lib = _lib.load()
    A = _f32(A)
    assert A.dim() == 2
    m, n = A.shape
    out = torch.empty((n, m), dtype=torch.float32, device=A.device)
    cond = torch.zeros(2, dtype=torch.float64, device=A.device) if return_cond else None
    if m > 0 and n > 0:
        nbytes = lib.anncur_pinv_workspace_bytes(m, n)
        ws = WORKSPACE.get("pinv", nbytes, A.device)
        with torch.cuda.device(A.device):
            _lib.check(lib.anncur_pinv_f32(_ptr(A), m, n, _ld(A), float(rcond), _ptr(out), max(m, 1), _ptr(cond),
                                           _ptr(ws), ws.numel(), _stream()))
            _require_converged(lib, ws, "pinv", (m, n))
    return (out, cond) if return_cond else out


def _require_converged(lib, ws, what, shape):
    """The Jacobi sweeps stop at a sweep limit; a factorisation that did not converge must not pass silently (the index
    build is a one-off, so the host read-back of the status is affordable)."""
    status = torch.zeros(4, dtype=torch.float64, device=ws.device)
    _lib.check(lib.anncur_jacobi_status(_ptr(ws), _ptr(status), _stream()))
    s_max, _, converged, sweeps = status.tolist()
    if converged != 1.0:
        raise RuntimeError(f"anncur_b200.{what}: Jacobi SVD of a {shape[0]} x {shape[1]} matrix did not converge in "
                           f"{int(sweeps)} sweeps (s_max = {s_max:.3g}); the result is not usable")


def singular_values(A):
    """Singular values of an (m x n) fp32 matrix, descending fp64 (Jacobi; eval/compute_m2e_matrix_ranks.py:44-53)."""
    lib = _lib.load()
    A = _f32(A)
    assert A.dim() == 2
    m, n = A.shape
    out = torch.zeros(min(m, n), dtype=torch.float64, device=A.device)
    if m > 0 and n > 0:
        nbytes = lib.anncur_pinv_workspace_bytes(m, n)
        ws = WORKSPACE.get("pinv", nbytes, A.device)
        with torch.cuda.device(A.device):
            _lib.check(lib.anncur_singular_values_f32(_ptr(A), m, n, _ld(A), _ptr(out), _ptr(ws), ws.numel(), _stream()))
            _require_converged(lib, ws, "singular_values", (m, n))
    return torch.sort(out, descending=True).values


def orthonormalize(Y):
    """Orthonormal basis (left singular vectors, unsorted) of the column space of a tall fp32 matrix + its singular values."""
    lib = _lib.load()
    Y = _f32(Y)
    assert Y.dim() == 2 and Y.shape[0] >= Y.shape[1], "orthonormalize: the matrix must be tall"
    m, n = Y.shape
    Q = torch.empty((m, n), dtype=torch.float32, device=Y.device)
    sigma = torch.zeros(n, dtype=torch.float64, device=Y.device)
    if m > 0 and n > 0:
        ws = WORKSPACE.get("pinv", lib.anncur_pinv_workspace_bytes(m, n), Y.device)
        with torch.cuda.device(Y.device):
            _lib.check(lib.anncur_orthonormalize_f32(_ptr(Y), m, n, _ld(Y), _ptr(Q), n, _ptr(sigma), _ptr(ws), ws.numel(), _stream()))
            _require_converged(lib, ws, "orthonormalize", (m, n))
    return Q, sigma


def leading_singular_values(A, n_values=256, n_iter=4, seed=0):
    """The n_values largest singular values of a large fp32 matrix A (n x N, either orientation) by randomised subspace
    iteration -- every product (A^T Q, A W) on the engine's GEMM, every orthonormalisation and the final small SVD on the
    Jacobi kernel.  For matrices where the full Jacobi SVD of ``singular_values`` is out of reach (10k x 1M = BASELINE
    configs[4]).  Returns fp64, descending.  Accuracy: the leading values converge like (sigma_{b+1} / sigma_i)^(2 n_iter + 1)."""
    A = _f32(A)
    n, N = A.shape
    if n > N:                       # iterate on the short side; sigma(A) = sigma(A^T)
        A = A.t().contiguous()
        n, N = A.shape
    b = int(min(n_values, n))
    g = torch.Generator(device=A.device)
    g.manual_seed(seed)
    Q, _ = orthonormalize(torch.randn((n, b), generator=g, device=A.device))
    for _ in range(n_iter):
        W = gemm(Q.t().contiguous(), A)                     # b x N   = Q^T A
        Y = gemm(A, W.t().contiguous())                     # n x b   = A A^T Q
        del W
        Q, _ = orthonormalize(Y)
    B = gemm(Q.t().contiguous(), A)                         # b x N: sigma(B) ~ the leading sigma(A)
    S = gemm(B, B.t().contiguous())                         # b x b Gram of the projected rows
    lam = singular_values(S)                                # symmetric PSD: singular values = eigenvalues
    return torch.sqrt(lam.clamp_min(0.0))


def matrix_rank_large(A, tol=None, n_values=256, n_iter=4, seed=0):
    """np.linalg.matrix_rank(A) (eval/compute_m2e_matrix_ranks.py:50) for matrices too large for a full SVD: counts the
    leading singular values above numpy's default tolerance ``sigma_max * max(n, N) * eps_fp32`` (which at N = 1M is
    0.12 sigma_max: only the leading part of the spectrum can count).  The block is doubled until it holds a value below the
    tolerance, so the count is not capped by it."""
    A = _f32(A)
    n, N = A.shape
    b = int(min(n_values, n, N))
    while True:
        s = leading_singular_values(A, b, n_iter, seed)
        t = float(s[0]) * max(n, N) * float(torch.finfo(torch.float32).eps) if tol is None else float(tol)
        rank = int((s > t).sum().item())
        if rank < b or b >= min(n, N):
            return rank
        b = min(2 * b, n, N)


def matrix_rank(A, tol=None):
    """np.linalg.matrix_rank(A): number of singular values above max(m, n) * eps(fp32) * sigma_max (or ``tol``)."""
    s = singular_values(A)
    if s.numel() == 0:
        return 0
    if tol is None:
        tol = float(s[0]) * max(A.shape) * float(torch.finfo(torch.float32).eps)
    return int((s > tol).sum().item())


# ---- K2 -------------------------------------------------------------------------------------------
def gemm(A, B):
    """A (m x k) @ B (k x n) in fp32 FFMA (eval/matrix_approx_zeshel.py:61,65,74,79,85,97,118)."""
    lib = _lib.load()
    A = _f32(A)
    B = _f32(B, device=A.device)
    assert A.dim() == 2 and B.dim() == 2 and A.shape[1] == B.shape[0], (A.shape, B.shape)
    m, k = A.shape
    n = B.shape[1]
    out = torch.empty((m, n), dtype=torch.float32, device=A.device)
    if m > 0 and n > 0:
        with torch.cuda.device(A.device):
            _lib.check(lib.anncur_gemm_f32(_ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(out), max(n, 1), m, n, k, _stream()))
    return out


# ---- packed item index ----------------------------------------------------------------------------
class PackedItems:
    """E = latent_cols (k_dim x N) in the TMA/tcgen05 streaming layout of one precision kind."""

    def __init__(self, E, kind="f32r"):
        lib = _lib.load()
        E = _f32(E)
        assert E.dim() == 2
        self.kind_name, self.kind = kind, KINDS[kind]
        self.k_dim, self.n_items = int(E.shape[0]), int(E.shape[1])
        self.device = E.device
        nbytes = lib.anncur_packed_items_bytes(self.n_items, self.k_dim, self.kind)
        if self.n_items and self.k_dim and not bool(torch.isfinite(E).all()):
            # the scales are chosen from the finite entries only; an inf / NaN item embedding would make every score of
            # that item undefined without any error -- refuse it here, once, at index build
            raise ValueError("PackedItems: the item embeddings contain inf / NaN")
        self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=E.device)
        self.scale = torch.ones(1, dtype=torch.float32, device=E.device)
        with torch.cuda.device(E.device):
            _lib.check(lib.anncur_pack_items(_ptr(E), _ld(E), self.n_items, self.k_dim, self.kind, _ptr(self.buf),
                                             _ptr(self.scale), _stream()))

    @property
    def nbytes(self):
        return self.buf.numel()


# ---- K3 + K4 --------------------------------------------------------------------------------------
def score_topk(Q, packed, k, idx_offset=0, out=None, ws=None):
    """Fused tensor-core approximate score + top-k (eval/matrix_approx_zeshel.py:109-126).  ``ws``: caller-owned
    workspace (uint8 CUDA tensor of anncur_score_topk_workspace_bytes); default = the per-device cached one."""
    lib = _lib.load()
    Q = _f32(Q, device=packed.device)
    assert Q.dim() == 2 and Q.shape[1] == packed.k_dim, (Q.shape, packed.k_dim)
    B = int(Q.shape[0])
    if out is None:
        vals = torch.empty((B, k), dtype=torch.float32, device=Q.device)
        idx = torch.empty((B, k), dtype=torch.int64, device=Q.device)
    else:
        vals, idx = out
    if B > 0:
        with torch.cuda.device(Q.device):              # the plan (SM count) is that of the index's device, not the current one
            nbytes = lib.anncur_score_topk_workspace_bytes(B, packed.n_items, packed.k_dim, k, packed.kind)
            if ws is None:
                ws = WORKSPACE.get("score_topk", nbytes, Q.device)
            assert ws.numel() >= nbytes and ws.device == Q.device
            _LAST_FUSED_WS[Q.device.index] = (ws, 0, B, int(k))
            _lib.check(lib.anncur_score_topk(_ptr(Q), _ld(Q), B, _ptr(packed.buf), _ptr(packed.scale), packed.n_items,
                                             packed.k_dim, packed.kind, int(k), int(idx_offset), _ptr(vals), _ptr(idx),
                                             _ptr(ws), ws.numel(), _stream()))
    return vals, idx


def score_dense(Q, packed, out=None):
    """Dense Q . E (B x N fp32) on the tensor-core pipeline (anncur_score_dense): the fp32-grade 3-pass arithmetic on a
    packed index of kind f32x3 / f32r.  eval/matrix_approx_zeshel.py:71-119 when the full matrix is wanted."""
    lib = _lib.load()
    Q = _f32(Q, device=packed.device)
    assert Q.dim() == 2 and Q.shape[1] == packed.k_dim, (Q.shape, packed.k_dim)
    assert packed.kind in (KIND_F32X3, KIND_F32R), "score_dense needs an fp32-grade index"
    B, N = int(Q.shape[0]), packed.n_items
    if out is None:
        out = torch.empty((B, N), dtype=torch.float32, device=Q.device)
    assert out.shape == (B, N) and out.dtype == torch.float32 and out.stride(1) == 1
    if B > 0 and N > 0:
        with torch.cuda.device(Q.device):
            nbytes = lib.anncur_score_dense_workspace_bytes(B, N, packed.k_dim, packed.kind)
            ws = WORKSPACE.get("score_dense", nbytes, Q.device)
            _lib.check(lib.anncur_score_dense(_ptr(Q), _ld(Q), B, _ptr(packed.buf), _ptr(packed.scale), N, packed.k_dim,
                                              packed.kind, _ptr(out), _ld(out), _ptr(ws), ws.numel(), _stream()))
    return out


def score_bounds_dense(Q, packed, sign):
    """Dense matrix of the score bounds kind f32r works with (anncur_score_bounds_dense): sign >= 0 upper, < 0 lower."""
    lib = _lib.load()
    Q = _f32(Q, device=packed.device)
    assert packed.kind == KIND_F32R and Q.shape[1] == packed.k_dim
    B, N = int(Q.shape[0]), packed.n_items
    out = torch.empty((B, N), dtype=torch.float32, device=Q.device)
    if B > 0 and N > 0:
        nbytes = lib.anncur_score_dense_workspace_bytes(B, N, packed.k_dim, packed.kind)
        ws = WORKSPACE.get("score_dense", nbytes, Q.device)
        with torch.cuda.device(Q.device):
            _lib.check(lib.anncur_score_bounds_dense(_ptr(Q), _ld(Q), B, _ptr(packed.buf), _ptr(packed.scale), N, packed.k_dim,
                                                     int(sign), _ptr(out), _ld(out), _ptr(ws), ws.numel(), _stream()))
    return out


def gemm_tc(A, B, min_flops=2e9):
    """A (m x k) @ B (k x n) through the tensor-core pipeline: B is packed as a throw-away fp32-grade index and A plays
    the queries (the item-embedding build U @ R, eval/matrix_approx_zeshel.py:65).  Small products go to the FFMA GEMM."""
    A = _f32(A)
    B = _f32(B, device=A.device)
    m, k = A.shape
    n = B.shape[1]
    if k == 0 or 2.0 * m * n * k < min_flops:
        return gemm(A, B)
    return score_dense(A, PackedItems(B, "f32x3"))


def recon_error_packed(Q, packed, A):
    """Per-row sum_j (Q . E - A)^2 and sum_j A^2 (fp64) without materialising Q . E, on the tensor-core pipeline."""
    lib = _lib.load()
    Q = _f32(Q, device=packed.device)
    A = _f32(A, device=packed.device)
    B, N = int(Q.shape[0]), packed.n_items
    assert A.shape == (B, N) and Q.shape[1] == packed.k_dim
    err2 = torch.empty(B, dtype=torch.float64, device=Q.device)
    norm2 = torch.empty(B, dtype=torch.float64, device=Q.device)
    if B > 0:
        nbytes = lib.anncur_score_dense_workspace_bytes(B, N, packed.k_dim, packed.kind)
        ws = WORKSPACE.get("score_dense", nbytes, Q.device)
        with torch.cuda.device(Q.device):
            _lib.check(lib.anncur_recon_error_packed(_ptr(Q), _ld(Q), B, _ptr(packed.buf), _ptr(packed.scale), N, packed.k_dim,
                                                     packed.kind, _ptr(A), _ld(A), _ptr(err2), _ptr(norm2), _ptr(ws),
                                                     ws.numel(), _stream()))
    return err2, norm2


class GraphedSearch:
    """score_topk for one fixed (batch, k) captured as a CUDA graph: a search is then ONE graph launch instead of the
    7-9 kernel launches + a memset of the eager call -- what matters for small batches, where the step is a few tens
    of microseconds and launch gaps are a large part of it.  Every library call is asynchronous on the caller's stream
    and allocates nothing, so the capture needs no special path.

        g = GraphedSearch(packed, B, k); vals, idx = g(Q)      # vals / idx are the graph's own output buffers
    """

    def __init__(self, packed, n_queries, k, idx_offset=0):
        lib = _lib.load()
        dev = packed.device
        self.packed, self.k = packed, int(k)
        self.q = torch.zeros((n_queries, packed.k_dim), dtype=torch.float32, device=dev)
        self.vals = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
        self.idx = torch.empty((n_queries, k), dtype=torch.int64, device=dev)
        nbytes = lib.anncur_score_topk_workspace_bytes(n_queries, packed.n_items, packed.k_dim, k, packed.kind)
        self.ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside the capture (function attributes, TMA encode entry)
            score_topk(self.q, packed, k, idx_offset, out=(self.vals, self.idx), ws=self.ws)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            score_topk(self.q, packed, k, idx_offset, out=(self.vals, self.idx), ws=self.ws)

    def __call__(self, Q):
        self.q.copy_(Q, non_blocking=True)
        self.graph.replay()
        return self.vals, self.idx


_LAST_FUSED_WS = {}          # device index -> (workspace tensor, byte offset of the fused workspace, n_queries, k) of the last call


def last_redo_rows(n_queries, packed, k):
    """Rows of the last fused search on the index's device -- score_topk, GraphedSearch or search_host, on whatever stream
    and workspace it ran -- that the fallback pass had to recompute (anncur_score_topk_redo_rows): 0 when the fast path
    served every row.  (n_queries, k) must be those of that call.  Synchronises the current stream."""
    lib = _lib.load()
    last = _LAST_FUSED_WS.get(packed.device.index)
    if last is None or (last[2], last[3]) != (int(n_queries), int(k)):
        raise RuntimeError("last_redo_rows: no fused search with this (n_queries, k) has run on this device")
    ws, off = last[0], last[1]
    n = C.c_int(0)
    with torch.cuda.device(packed.device):
        _lib.check(lib.anncur_score_topk_redo_rows(C.c_void_p(ws.data_ptr() + off), int(n_queries), packed.n_items, packed.k_dim,
                                                   int(k), packed.kind, C.byref(n), _stream()))
    return n.value


def search_host(Q_host, packed, k, out_vals_host, out_idx_host, idx_offset=0, ws_key="search_host"):
    """Host-buffer form (anncur_search_host): Q_host / out_* are CPU tensors (pinned => fully asynchronous).
    Enqueues H2D -> fused score + top-k -> D2H on the current stream of the index's device and returns at
    once; the outputs are valid after that stream (or the device) is synchronised."""
    lib = _lib.load()
    assert not Q_host.is_cuda and Q_host.dtype == torch.float32 and Q_host.dim() == 2 and Q_host.stride(1) == 1
    assert Q_host.shape[1] == packed.k_dim, (Q_host.shape, packed.k_dim)
    B = int(Q_host.shape[0])
    assert out_vals_host.shape == (B, k) and out_vals_host.dtype == torch.float32 and out_vals_host.is_contiguous()
    assert out_idx_host.shape == (B, k) and out_idx_host.dtype == torch.int64 and out_idx_host.is_contiguous()
    assert not out_vals_host.is_cuda and not out_idx_host.is_cuda
    if B > 0:
        with torch.cuda.device(packed.device):
            nbytes = lib.anncur_search_host_workspace_bytes(B, packed.n_items, packed.k_dim, k, packed.kind)
            inner = lib.anncur_score_topk_workspace_bytes(B, packed.n_items, packed.k_dim, k, packed.kind)
            ws = WORKSPACE.get(ws_key, nbytes, packed.device)
            _LAST_FUSED_WS[packed.device.index] = (ws, int(nbytes - inner), B, int(k))     # the fused workspace follows the staging buffers
            _lib.check(lib.anncur_search_host(_ptr(Q_host), _ld(Q_host), B, _ptr(packed.buf), _ptr(packed.scale),
                                              packed.n_items, packed.k_dim, packed.kind, int(k), int(idx_offset),
                                              _ptr(out_vals_host), _ptr(out_idx_host), _ptr(ws), ws.numel(), _stream()))
    return out_vals_host, out_idx_host


def score_topk_f32(Q, E, k, idx_offset=0):
    """Plain-fp32 FFMA score + top-k on unpacked E (any k <= MAX_K)."""
    lib = _lib.load()
    Q = _f32(Q)
    E = _f32(E, device=Q.device)
    assert Q.dim() == 2 and E.dim() == 2 and Q.shape[1] == E.shape[0]
    B, K = Q.shape
    N = E.shape[1]
    vals = torch.empty((B, k), dtype=torch.float32, device=Q.device)
    idx = torch.empty((B, k), dtype=torch.int64, device=Q.device)
    if B > 0:
        nbytes = lib.anncur_score_topk_f32_workspace_bytes(B, N, K, k)
        ws = WORKSPACE.get("score_topk_f32", nbytes, Q.device)
        with torch.cuda.device(Q.device):
            _lib.check(lib.anncur_score_topk_f32(_ptr(Q), _ld(Q), B, _ptr(E), _ld(E), N, K, int(k), int(idx_offset),
                                                 _ptr(vals), _ptr(idx), _ptr(ws), ws.numel(), _stream()))
    return vals, idx


def topk_rows(S, k, idx_offset=0):
    """torch.topk(S, k, dim=1) on the GPU, ties -> lower index."""
    lib = _lib.load()
    S = _f32(S)
    assert S.dim() == 2
    n, N = S.shape
    vals = torch.empty((n, k), dtype=torch.float32, device=S.device)
    idx = torch.empty((n, k), dtype=torch.int64, device=S.device)
    if n > 0:
        with torch.cuda.device(S.device):
            _lib.check(lib.anncur_topk_rows_f32(_ptr(S), _ld(S), n, N, int(k), int(idx_offset), _ptr(vals), _ptr(idx),
                                                _stream()))
    return vals, idx


def merge_topk(cand_vals, cand_idx, k):
    """Best k of n_cand (val, idx) candidates per row; idx < 0 marks padding."""
    lib = _lib.load()
    cand_vals = _f32(cand_vals).contiguous()
    cand_idx = cand_idx.to(device=cand_vals.device, dtype=torch.int64).contiguous()
    assert cand_vals.shape == cand_idx.shape and cand_vals.dim() == 2
    n, m = cand_vals.shape
    vals = torch.empty((n, k), dtype=torch.float32, device=cand_vals.device)
    idx = torch.empty((n, k), dtype=torch.int64, device=cand_vals.device)
    if n > 0:
        with torch.cuda.device(cand_vals.device):
            _lib.check(lib.anncur_merge_topk(_ptr(cand_vals), _ptr(cand_idx), n, m, int(k), _ptr(vals), _ptr(idx), _stream()))
    return vals, idx


def topk_to_keys(vals, idx):
    """(fp32 [n x k], int64 [n x k] global indices < 2^32) -> int64 tensor of 64-bit candidate keys [n x k]."""
    lib = _lib.load()
    vals = _f32(vals).contiguous()
    idx = idx.to(device=vals.device, dtype=torch.int64).contiguous()
    n, k = vals.shape
    keys = torch.empty((n, k), dtype=torch.int64, device=vals.device)
    if n > 0:
        with torch.cuda.device(vals.device):
            _lib.check(lib.anncur_topk_to_keys(_ptr(vals), _ptr(idx), n, k, _ptr(keys), _stream()))
    return keys


def merge_topk_keys(keys, k):
    """keys: int64 [n_shards x n_rows x k_in] (the all-gathered buffer as is) -> best k per row (vals, idx)."""
    lib = _lib.load()
    assert keys.is_cuda and keys.dtype == torch.int64 and keys.dim() == 3 and keys.is_contiguous()
    P, n, k_in = keys.shape
    vals = torch.empty((n, k), dtype=torch.float32, device=keys.device)
    idx = torch.empty((n, k), dtype=torch.int64, device=keys.device)
    if n > 0:
        nbytes = lib.anncur_merge_topk_keys_workspace_bytes(n)
        ws = WORKSPACE.get("merge_keys", nbytes, keys.device)
        with torch.cuda.device(keys.device):
            _lib.check(lib.anncur_merge_topk_keys(_ptr(keys), P, n, k_in, int(k), _ptr(vals), _ptr(idx), _ptr(ws), ws.numel(),
                                                  _stream()))
    return vals, idx


# ---- K5 + K6 --------------------------------------------------------------------------------------
def rerank_overlap(exact, retr_idx, exact_idx, k_list):
    """Exact-score rerank of retrieved items + |exact[:k] & reranked[:k]| for each k in k_list.
    Returns (rr_idx [n x k_max], rr_vals [n x k_max], common [n x len(k_list)] int32)."""
    lib = _lib.load()
    exact = _f32(exact)
    dev = exact.device
    retr_idx = retr_idx.to(device=dev, dtype=torch.int64).contiguous()
    exact_idx = exact_idx.to(device=dev, dtype=torch.int64).contiguous()
    n, N = exact.shape
    k_retr, k_max = int(retr_idx.shape[1]), int(exact_idx.shape[1])
    k_list = [int(x) for x in k_list]
    rr_idx = torch.empty((n, k_max), dtype=torch.int64, device=dev)
    rr_vals = torch.empty((n, k_max), dtype=torch.float32, device=dev)
    common = torch.empty((n, len(k_list)), dtype=torch.int32, device=dev)
    if n > 0:
        arr = (C.c_int * len(k_list))(*k_list)
        with torch.cuda.device(dev):
            _lib.check(lib.anncur_rerank_overlap(_ptr(exact), _ld(exact), n, N, _ptr(retr_idx), k_retr, _ptr(exact_idx),
                                                 k_max, arr, len(k_list), _ptr(rr_idx), _ptr(rr_vals), _ptr(common),
                                                 _stream()))
    return rr_idx, rr_vals, common


def overlap_counts(a_idx, b_idx):
    """|set(a[row]) & set(b[row])| per row for two (n x k) int64 index lists (eval/eval_utils.py:139-150)."""
    lib = _lib.load()
    require_cuda()
    a = torch.as_tensor(a_idx)
    dev = a.device if a.is_cuda else torch.device("cuda", torch.cuda.current_device())
    a = a.to(device=dev, dtype=torch.int64).contiguous()
    b = torch.as_tensor(b_idx).to(device=dev, dtype=torch.int64).contiguous()
    assert a.dim() == 2 and a.shape == b.shape, (a.shape, b.shape)
    n, k = a.shape
    out = torch.zeros(n, dtype=torch.int32, device=dev)
    if n > 0 and k > 0:
        with torch.cuda.device(dev):
            _lib.check(lib.anncur_overlap_counts(_ptr(a), _ptr(b), n, k, _ptr(out), _stream()))
    return out


# ---- K7 -------------------------------------------------------------------------------------------
def recon_error_rows(Q, E, A):
    """Per-row sum_j (Q.E - A)^2 and sum_j A^2 in fp64 without materialising Q.E."""
    lib = _lib.load()
    A = _f32(A)
    Q = _f32(Q, device=A.device)
    E = _f32(E, device=A.device)
    n, N = A.shape
    K = Q.shape[1]
    assert Q.shape[0] == n and E.shape == (K, N)
    err2 = torch.empty(n, dtype=torch.float64, device=A.device)
    norm2 = torch.empty(n, dtype=torch.float64, device=A.device)
    if n > 0:
        with torch.cuda.device(A.device):
            _lib.check(lib.anncur_recon_error_f32(_ptr(Q), _ld(Q), _ptr(E), _ld(E), _ptr(A), _ld(A), n, N, K,
                                                  _ptr(err2), _ptr(norm2), _stream()))
    return err2, norm2


# ---- K8 -------------------------------------------------------------------------------------------
def adaptive_round(R_anc, anchors, c, n_next, rcond=1e-15):
    """One adaptive ANNCUR round (see include/anncur_b200.h).  Returns (next_idx, next_val)."""
    lib = _lib.load()
    R_anc = _f32(R_anc)
    dev = R_anc.device
    anchors = anchors.to(device=dev, dtype=torch.int64).contiguous()
    c = _f32(c, device=dev).contiguous()
    k_q, N = R_anc.shape
    B, m = anchors.shape
    next_idx = torch.empty((B, n_next), dtype=torch.int64, device=dev)
    next_val = torch.empty((B, n_next), dtype=torch.float32, device=dev)
    if B > 0:
        nbytes = lib.anncur_adaptive_round_workspace_bytes(B, k_q, m, N, n_next)
        ws = WORKSPACE.get("adaptive", nbytes, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.anncur_adaptive_round(_ptr(R_anc), _ld(R_anc), k_q, N, _ptr(anchors), _ptr(c), B, m,
                                                 float(rcond), int(n_next), _ptr(next_idx), _ptr(next_val), _ptr(ws),
                                                 ws.numel(), _stream()))
    return next_idx, next_val


def transpose(X):
    """X (rows x cols fp32) -> X^T contiguous (cols x rows) with the library's tiled transpose."""
    lib = _lib.load()
    X = _f32(X)
    rows, cols = X.shape
    out = torch.empty((cols, rows), dtype=torch.float32, device=X.device)
    if rows > 0 and cols > 0:
        with torch.cuda.device(X.device):
            _lib.check(lib.anncur_transpose_f32(_ptr(X), _ld(X), rows, cols, _ptr(out), _stream()))
    return out


def adaptive_solve(R_anc, anchors, c, rcond=1e-15, Rt=None):
    """e_b = c_b . pinv(R_anc[:, I_b]) for a batch of queries (anncur_adaptive_solve): (B x k_q) fp32.  ``Rt`` = R_anc^T kept
    by the caller across rounds (``engine.transpose(R_anc)``)."""
    lib = _lib.load()
    R_anc = _f32(R_anc)
    dev = R_anc.device
    anchors = anchors.to(device=dev, dtype=torch.int64).contiguous()
    c = _f32(c, device=dev).contiguous()
    k_q, N = R_anc.shape
    B, m = anchors.shape
    e = torch.empty((B, k_q), dtype=torch.float32, device=dev)
    if B > 0:
        with torch.cuda.device(dev):
            nbytes = lib.anncur_adaptive_solve_workspace_bytes(B, k_q, m, N)
            ws = WORKSPACE.get("adaptive", nbytes, dev)
            _lib.check(lib.anncur_adaptive_solve(_ptr(R_anc), _ld(R_anc), k_q, N, _ptr(Rt), _ptr(anchors), _ptr(c), B, m, float(rcond),
                                                 _ptr(e), _ptr(ws), ws.numel(), _stream()))
    return e


def filter_excluded(cand_vals, cand_idx, excluded, n_out):
    """First n_out candidates per row (order kept) whose index is not in the row's ``excluded`` list; padded with (-FLT_MAX, -1)."""
    lib = _lib.load()
    cand_vals = _f32(cand_vals).contiguous()
    dev = cand_vals.device
    cand_idx = cand_idx.to(device=dev, dtype=torch.int64).contiguous()
    excluded = excluded.to(device=dev, dtype=torch.int64).contiguous()
    n, k_in = cand_vals.shape
    m = excluded.shape[1]
    vals = torch.empty((n, n_out), dtype=torch.float32, device=dev)
    idx = torch.empty((n, n_out), dtype=torch.int64, device=dev)
    if n > 0:
        with torch.cuda.device(dev):
            _lib.check(lib.anncur_filter_excluded(_ptr(cand_vals), _ptr(cand_idx), n, k_in, _ptr(excluded), m, int(n_out), _ptr(vals),
                                                  _ptr(idx), _stream()))
    return vals, idx


def score_topk_excluding(Q, packed, k, excluded, idx_offset=0):
    """Top-k of every row among the items NOT in the row's ``excluded`` list (B x m int64, global indices) -- the masked
    re-score of the adaptive rounds as one call (anncur_score_topk_excluding)."""
    lib = _lib.load()
    Q = _f32(Q, device=packed.device)
    B = int(Q.shape[0])
    excluded = excluded.to(device=Q.device, dtype=torch.int64).contiguous()
    assert excluded.dim() == 2 and excluded.shape[0] == B
    m = int(excluded.shape[1])
    vals = torch.empty((B, k), dtype=torch.float32, device=Q.device)
    idx = torch.empty((B, k), dtype=torch.int64, device=Q.device)
    if B > 0:
        with torch.cuda.device(Q.device):
            nbytes = lib.anncur_score_topk_excluding_workspace_bytes(B, packed.n_items, packed.k_dim, int(k), m, packed.kind)
            ws = WORKSPACE.get("score_topk_excluding", nbytes, Q.device)
            _lib.check(lib.anncur_score_topk_excluding(_ptr(Q), _ld(Q), B, _ptr(packed.buf), _ptr(packed.scale), packed.n_items,
                                                       packed.k_dim, packed.kind, int(k), _ptr(excluded), m, int(idx_offset),
                                                       _ptr(vals), _ptr(idx), _ptr(ws), ws.numel(), _stream()))
    return vals, idx


class AdaptiveShared:
    """What anncur_adaptive_prepare builds once per (R_anc, shared first anchors): the Cholesky factor of the shared Gram
    matrix, its inverse and W_1^T (n_items x m_shared fp64) -- see include/anncur_b200.h."""

    def __init__(self, Rt, shared_anchors, rcond=1e-15):
        lib = _lib.load()
        self.Rt = _f32(Rt).contiguous()
        dev = self.Rt.device
        self.N, self.k_q = self.Rt.shape
        self.anchors = torch.as_tensor(shared_anchors, dtype=torch.int64, device=dev).contiguous()
        self.m_shared = int(self.anchors.numel())
        self.rcond = float(rcond)
        with torch.cuda.device(dev):
            self.blob = torch.empty(lib.anncur_adaptive_shared_bytes(self.k_q, self.N, self.m_shared), dtype=torch.uint8, device=dev)
            ws = torch.empty(lib.anncur_adaptive_prepare_workspace_bytes(self.k_q, self.N, self.m_shared), dtype=torch.uint8, device=dev)
            _lib.check(lib.anncur_adaptive_prepare(_ptr(self.Rt), self.k_q, self.N, _ptr(self.anchors), self.m_shared, self.rcond,
                                                   _ptr(self.blob), self.blob.numel(), _ptr(ws), ws.numel(), _stream()))
            torch.cuda.current_stream(dev).synchronize()          # ws is dropped on return
        del ws


class AdaptiveState:
    """Per-batch state of the incremental adaptive solver (anncur_adaptive_begin / anncur_adaptive_extend): call ``begin(c)``
    with the exact scores of the shared anchors, then ``extend(new_anchors, c_new)`` once per round; both return e (B x k_q)."""

    def __init__(self, shared, n_queries, n_new, m_max):
        lib = _lib.load()
        self.shared, self.B, self.n_new, self.m_max = shared, int(n_queries), int(n_new), int(m_max)
        self.m_cur = None
        dev = shared.Rt.device
        with torch.cuda.device(dev):
            nbytes = lib.anncur_adaptive_state_bytes(self.B, shared.k_q, shared.m_shared, self.n_new, self.m_max)
            # the state belongs to this object (two batches in flight on one stream must not share it); torch's caching
            # allocator hands the block of the previous batch back without a cudaMalloc
            self.blob = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)

    def begin(self, c):
        lib, sh = _lib.load(), self.shared
        dev = sh.Rt.device
        c = _f32(c, device=dev).contiguous()
        assert c.shape == (self.B, sh.m_shared)
        e = torch.empty((self.B, sh.k_q), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.anncur_adaptive_begin(_ptr(sh.Rt), sh.k_q, sh.N, _ptr(sh.blob), sh.m_shared, _ptr(c), self.B, self.n_new,
                                                 self.m_max, _ptr(e), _ptr(self.blob), self.blob.numel(), _stream()))
        self.m_cur = sh.m_shared
        return e

    def extend(self, new_anchors, c_new):
        lib, sh = _lib.load(), self.shared
        dev = sh.Rt.device
        assert self.m_cur is not None, "begin() first"
        new_anchors = new_anchors.to(device=dev, dtype=torch.int64).contiguous()
        c_new = _f32(c_new, device=dev).contiguous()
        assert new_anchors.shape == (self.B, self.n_new) and c_new.shape == (self.B, self.n_new)
        e = torch.empty((self.B, sh.k_q), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.anncur_adaptive_extend(_ptr(sh.Rt), sh.k_q, sh.N, _ptr(sh.blob), sh.m_shared, _ptr(new_anchors), _ptr(c_new),
                                                  self.B, self.n_new, self.m_max, self.m_cur, sh.rcond, _ptr(e), _ptr(self.blob),
                                                  self.blob.numel(), _stream()))
        self.m_cur += self.n_new
        return e
