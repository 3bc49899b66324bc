"""Import the UNMODIFIED reference: from ``/root/reference`` in the dev container, from the verbatim copy
``baseline/_ref/reference`` (tools/install_reference.py; git-ignored, travels with the snapshot) on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product path.  Users: ``tests/golden/make_golden.py``
(generates the committed golden vectors), ``tests/test_oracle_vs_reference.py`` (pins the oracle),
``tests/test_reference_dropin.py`` (cross-decode and the drop-in swap) and ``bench.py``'s CPU legs (``kind:
"reference"``).  All of them skip or fall back to the oracle port when no reference tree is present.

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

_COPY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "reference")


def _find_reference():
    """JB_REFERENCE_ROOT, else the source tree of the dev container, else the verbatim copy that
    tools/install_reference.py leaves under baseline/_ref/ (the only one that exists on the GPU box)."""
    env = os.environ.get("JB_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _COPY):
        if os.path.isfile(os.path.join(cand, "pipeline", "__init__.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference()
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
