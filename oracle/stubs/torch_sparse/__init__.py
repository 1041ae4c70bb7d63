"""Import stub for `torch_sparse` (rusty1s/pytorch_sparse, unpinned in the reference).

TEST INFRASTRUCTURE ONLY.  The reference imports torch_sparse unconditionally
(/root/reference/Helpers/Torches.py:13-16) but the package is not installed in this image.
The only semantics the hot path relies on are
  * SparseTensor.from_torch_sparse_coo_tensor(mat).coalesce()  -> sum duplicates
  * matmul(sparse, dense)                                       -> sum-reduce CSR SpMM
Both are restated here with stock torch so the *unmodified* reference modules can be
imported by oracle/gen_golden.py and by the live-reference pinning tests.
"""
import torch


class SparseTensor:
    def __init__(self, coo):
        self._coo = coo

    @classmethod
    def from_torch_sparse_coo_tensor(cls, mat, has_value=True):
        return cls(mat)

    def coalesce(self, reduce="sum"):
        return SparseTensor(self._coo.coalesce())

    def to(self, *args, **kwargs):
        return SparseTensor(self._coo.to(*args, **kwargs))

    def to_torch_sparse_coo_tensor(self):
        return self._coo

    def device(self):
        return self._coo.device


def matmul(src, other, reduce="sum"):
    assert reduce == "sum"
    coo = src._coo if isinstance(src, SparseTensor) else src
    return torch.sparse.mm(coo, other)
