from .._stubs import stub_getattr

__getattr__ = stub_getattr(__name__)
