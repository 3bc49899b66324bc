"""Host-side mirror of the reference's codec configuration objects.

Same names, constructor arguments, JSON form and error behaviour as
``pipeline.QuantizationMethod`` (pipeline/__init__.py:13-47) and
``pipeline.Configuration`` (pipeline/__init__.py:50-64).  A quantisation method
here carries no numpy quantiser object: it names the mode the kernels run."""
import json

from . import _lib
from .errors import BadQuantizationError

# name -> (C-ABI mode, parameter name, default)   quantizers.py:4-53
_QUANT = {
    "none": (_lib.JB_Q_NONE, None, None),
    "discard": (_lib.JB_Q_DISCARD, "keep", 2),
    "divide": (_lib.JB_Q_DIVIDE, "divisor", 40),
    "qtable": (_lib.JB_Q_QTABLE, None, None),
}


class QuantizationMethod:
    def __init__(self, name, **kwargs):
        self.name = name
        self.params = kwargs
        error_msg = "name {}, params {}".format(self.name, self.params)
        if name not in _QUANT:
            raise BadQuantizationError(error_msg)
        mode, pname, default = _QUANT[name]
        allowed = {pname} if pname else set()
        if set(kwargs) - allowed:
            raise BadQuantizationError(error_msg)      # unexpected keyword, as the class call would fail
        self.mode = mode
        self.param = kwargs.get(pname, default) if pname else 0
        # The kernels take --qkeep / --qdivisor as the integers the CLI produces (compress.py:45-51).  The reference's
        # quantiser classes would also accept a fractional divisor or a negative keep (Python slice semantics); the
        # CUDA path does not, and says so here, at construction, rather than at the first call.
        self.c_param()

    def to_json(self):
        d = dict(self.params)
        d["quantization_scheme_name"] = self.name
        return json.dumps(d)

    @staticmethod
    def from_json(s):
        d = json.loads(s)
        name = d["quantization_scheme_name"]
        params = dict(d)
        del params["quantization_scheme_name"]
        return QuantizationMethod(name, **params)

    def c_param(self):
        """The integer handed to the kernels (--qkeep / --qdivisor are ints, compress.py:45-51)."""
        p = self.param
        if p is None:
            return 0
        try:
            integral = int(p) == p
        except (TypeError, ValueError):
            integral = False
        if not integral or (self.name == "discard" and int(p) < 0) or (self.name == "divide" and int(p) == 0):
            raise BadQuantizationError("name {}, params {}: the CUDA path takes integer parameters "
                                       "(keep >= 0, divisor != 0)".format(self.name, self.params))
        return int(p)


class Configuration:
    def __init__(self, width, height, block_size=2, dct_size=8, transform="DCT", quantization=None):
        self.width = width
        self.height = height
        self.block_size = block_size
        self.dct_size = dct_size
        self.transform = transform
        if quantization is None:
            self.quantization = QuantizationMethod("none")
        else:
            if quantization.name == "qtable" and dct_size != 8:
                raise BadQuantizationError()
            self.quantization = quantization

    def c_params(self, flags=0):
        if self.transform == "DCT":
            tr = _lib.JB_TRANSFORM_DCT
        elif self.transform == "DFT":
            tr = _lib.JB_TRANSFORM_DFT
        else:
            # the reference dies with UnboundLocalError inside BasisChange.execute
            # (pipeline/basis_change.py:26); keep the type, say why
            raise UnboundLocalError("unknown transform %r (expected 'DCT' or 'DFT')" % (self.transform,))
        return _lib.jb_params(int(self.height), int(self.width), int(self.block_size), int(self.dct_size),
                              tr, self.quantization.mode, self.quantization.c_param(), int(flags))
