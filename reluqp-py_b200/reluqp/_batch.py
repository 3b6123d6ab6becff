"""Host side of the batched solve (placeholder; filled in with the batched kernels)."""


class BatchEngine(object):
    def __init__(self, solver):
        raise RuntimeError("batched solve is not built yet")
