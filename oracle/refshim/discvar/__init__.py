"""discvar -> the reference's own in-tree DVR classes (lazy, to dodge the import cycle)."""
from . import abc, ho  # noqa: F401


def __getattr__(name):
    if name in ("HarmonicOscillator", "PrimBas_HO"):
        import pytdscf.basis.ho as _ho

        return getattr(_ho, name)
    if name in ("Sine",):
        import pytdscf.basis.sin as _s

        return getattr(_s, name)
    if name in ("Exponential",):
        import pytdscf.basis.exponential as _e

        return getattr(_e, name)
    raise AttributeError(name)
