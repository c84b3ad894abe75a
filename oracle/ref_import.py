"""Import shim for the read-only reference checkout (authoring container only).

``/root/reference`` does not exist on the GPU box; everything here is optional
and every caller must cope with ``reference_available() == False``.

The reference's ``utils/my_trainer.py`` imports modules that are absent here
(asyncore was removed in Python 3.12; matplotlib / skimage / skorch /
tune_sklearn are not installed).  None of them is on the hot path, so they are
replaced by empty stubs before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SIVAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "models.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_stubs():
    class _Dummy:  # stands in for classes that are only imported, never used
        def __init__(self, *a, **k):
            pass

    def _noop(*a, **k):
        return None

    try:
        import asyncore  # noqa: F401
    except Exception:
        _stub("asyncore", loop=_noop)
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib", use=_noop)
        plt = _stub("matplotlib.pyplot", figure=_noop, savefig=_noop, close=_noop, plot=_noop)
        mpl.pyplot = plt
    try:
        import skimage.metrics  # noqa: F401
    except Exception:
        sk = _stub("skimage")
        sk.metrics = _stub("skimage.metrics", mean_squared_error=_noop, structural_similarity=_noop)
    try:
        import skorch  # noqa: F401
    except Exception:
        s = _stub("skorch", NeuralNetClassifier=_Dummy)
        s.callbacks = _stub("skorch.callbacks", Callback=_Dummy, Checkpoint=_Dummy, EarlyStopping=_Dummy)
        s.dataset = _stub("skorch.dataset", CVSplit=_Dummy)
    try:
        import tune_sklearn  # noqa: F401
    except Exception:
        _stub("tune_sklearn", TuneSearchCV=_Dummy, TuneGridSearchCV=_Dummy)
    try:
        import seaborn  # noqa: F401
    except Exception:
        _stub("seaborn", heatmap=_noop, set=_noop)


def import_reference():
    """Return ``(models.models, models.vaemodel, models.lossf, utils.my_trainer)``
    of the UNMODIFIED reference, imported from REFERENCE_ROOT."""
    if not reference_available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the checkout is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _install_stubs()
    import importlib

    ref_models = importlib.import_module("models.models")
    ref_vaemodel = importlib.import_module("models.vaemodel")
    ref_lossf = importlib.import_module("models.lossf")
    ref_trainer = importlib.import_module("utils.my_trainer")
    # neutralise the plotting side effects
    ref_trainer.save_image = lambda *a, **k: None
    ref_trainer.train_result.result_rec_kls_loss = lambda *a, **k: None
    return ref_models, ref_vaemodel, ref_lossf, ref_trainer
