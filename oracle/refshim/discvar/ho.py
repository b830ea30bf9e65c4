def __getattr__(name):
    if name in ("HarmonicOscillator", "PrimBas_HO"):
        import pytdscf.basis.ho as _ho

        return getattr(_ho, name)
    raise AttributeError(name)
