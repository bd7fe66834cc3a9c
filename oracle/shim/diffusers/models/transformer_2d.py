from .._stubs import make_stub

Transformer2DModel = make_stub("Transformer2DModel")
