"""Drop-in for ``models/nearest_nbr.py::build_flat_or_ivff_index`` (:24-55): the object it returns
offers faiss' ``search(x, k) -> (D float32 [nq x k], I int64 [nq x k])`` (row-major numpy, best
first, (-FLT_MAX, -1) padding when k > N), backed by the fused tcgen05 score + top-k kernel.

ANNCUR mapping: embeds = E.T (N x k_i), x = Q (anchor-item CE scores).
Only exact inner-product search is provided: the reference's IVF branch (:40-52) is an approximate
index that ANNCUR does not use, so ``force_exact_search=False`` on a large collection still returns
the exact answer (a superset of IVF quality) and logs that it did.
"""
import logging

import numpy as np
import torch

from . import engine

LOGGER = logging.getLogger(__name__)


class FlatIPIndex:
    """Exact maximum-inner-product index over N d-dimensional embeddings (faiss.IndexFlatIP surface)."""

    def __init__(self, d, precision="f32r", device=None):
        engine.require_cuda()
        self.d = int(d)
        self.precision = precision
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.ntotal = 0
        self._E = None           # d x N fp32 on the GPU
        self._packed = None
        self.is_trained = True

    def add(self, embeds):
        x = torch.as_tensor(np.asarray(embeds, dtype=np.float32) if not torch.is_tensor(embeds) else embeds)
        assert x.dim() == 2 and x.shape[1] == self.d, f"expected [n x {self.d}] embeddings"
        Et = x.to(self.device, torch.float32).t().contiguous()
        self._E = Et if self._E is None else torch.cat([self._E, Et], dim=1)
        self.ntotal = int(self._E.shape[1])
        self._packed = None

    def search(self, x, k):
        x_t = torch.as_tensor(np.asarray(x, dtype=np.float32) if not torch.is_tensor(x) else x)
        assert x_t.dim() == 2 and x_t.shape[1] == self.d
        nq, k = int(x_t.shape[0]), int(k)
        if self.ntotal == 0 or nq == 0:
            return (np.full((nq, k), -np.finfo(np.float32).max, np.float32), np.full((nq, k), -1, np.int64))
        Q = x_t.to(self.device, torch.float32)
        if self.precision == "f32" or k > engine.MAX_K_FUSED:
            D, I = engine.score_topk_f32(Q, self._E, k)
        else:
            if self._packed is None:
                self._packed = engine.PackedItems(self._E, self.precision)
            D, I = engine.score_topk(Q, self._packed, k)
        return D.cpu().numpy(), I.cpu().numpy()


def build_flat_or_ivff_index(embeds, force_exact_search, probe_mult_factor=1, precision="f32r", device=None):
    LOGGER.info(f"Beginning indexing given {len(embeds)} embeddings")
    if type(embeds) is not np.ndarray:
        if torch.is_tensor(embeds):
            embeds = embeds.detach().cpu().numpy() if not embeds.is_cuda else embeds
        else:
            embeds = np.array(embeds)
    d = embeds.shape[1]
    nembeds = embeds.shape[0]
    if not (nembeds <= 11000 or force_exact_search):
        LOGGER.info("IVF requested (models/nearest_nbr.py:40-52): serving exact flat inner-product search instead")
    index = FlatIPIndex(d, precision=precision, device=device)
    index.add(embeds)
    LOGGER.info("Finished indexing given embeddings")
    return index
