def __getattr__(name):
    if name == "DVRPrimitivesMixin":
        import pytdscf.basis.abc as _abc

        return _abc.DVRPrimitivesMixin
    raise AttributeError(name)
