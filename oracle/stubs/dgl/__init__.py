"""Import stub for `dgl` (only the out-of-scope GATLayer touches it,
/root/reference/Models/GnnLayers.py:88,112,114).  TEST INFRASTRUCTURE ONLY."""


def graph(*args, **kwargs):
    raise NotImplementedError("dgl is not available; GATLayer is out of scope")


class ops:
    @staticmethod
    def edge_softmax(*args, **kwargs):
        raise NotImplementedError("dgl is not available; GATLayer is out of scope")

    @staticmethod
    def u_mul_e_sum(*args, **kwargs):
        raise NotImplementedError("dgl is not available; GATLayer is out of scope")
