"""Import the UNMODIFIED reference from ``/root/reference`` (dev container only).

TEST INFRASTRUCTURE ONLY -- never imported by the product path, by ``bench.py``
or by the ``-m gpu`` tests (``/root/reference`` does not exist on the GPU box).
It is used by ``tests/golden/make_golden.py`` to generate the committed golden
vectors and by the optional ``tests/test_oracle_vs_reference.py`` pinning test,
which skips itself when the reference tree is absent.

Two shims are needed because the reference does not import as shipped here
(SURVEY.md section 0.3):

* ``bitarray`` is not installed -> ``oracle/_shim/bitarray.py`` stand-in
  (container only, no arithmetic);
* numpy >= 1.24 removed ``np.float`` / ``np.int`` / ``np.complex``, which the
  reference uses at ``pipeline/basis_change.py:16,20,29,43`` and
  ``pipeline/run_length_encoding.py:16`` -> aliased to the builtins, which is
  what they always were.
"""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("JB_REFERENCE_ROOT", "/root/reference")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shim")

_cached = None


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pipeline", "__init__.py"))


def load_reference():
    """Return a namespace object with the reference's modules.

    Attributes: ``pipeline``, ``util``, ``transforms``, ``quantizers``,
    ``file_format`` (the reference's own, unmodified modules).
    """
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)

    import numpy as np

    for name, builtin in (("float", float), ("int", int), ("complex", complex)):
        if name not in np.__dict__:
            setattr(np, name, builtin)

    if "bitarray" not in sys.modules:
        sys.path.insert(0, _SHIM_DIR)
        import bitarray  # noqa: F401  (the stand-in)
        sys.path.remove(_SHIM_DIR)

    # the reference uses root-relative imports; `pipeline` must be imported
    # before `file_format` (they import each other)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import pipeline
            import util
            import transforms
            import quantizers
            import file_format
    finally:
        sys.path.remove(REFERENCE_ROOT)

    class _Ref:
        pass

    ref = _Ref()
    ref.pipeline = pipeline
    ref.util = util
    ref.transforms = transforms
    ref.quantizers = quantizers
    ref.file_format = file_format
    _cached = ref
    return ref
