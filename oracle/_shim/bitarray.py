"""Minimal pure-Python stand-in for the third-party ``bitarray`` package.

TEST INFRASTRUCTURE ONLY.  The reference imports ``bitarray`` (reference
``util.py:3``, ``pipeline/rle_byte_stream.py:1``) but the package is not
installed in this image and cannot be installed (no network).  This module
implements exactly the subset of the ``bitarray`` API the reference touches so
that the *unmodified* reference can be imported from ``/root/reference`` when
golden vectors are generated (``tests/golden/make_golden.py``).

It is a container only: no codec arithmetic lives here.  Bits are MSB-first
inside each byte, and ``tobytes`` zero-pads the tail, like the real package.
"""


class bitarray:
    __slots__ = ("_b",)

    def __init__(self, init=None):
        # bits are kept as a Python list of 0/1 ints
        if init is None:
            self._b = []
        elif isinstance(init, str):
            self._b = [1 if ch == "1" else 0 for ch in init]
        elif isinstance(init, bitarray):
            self._b = list(init._b)
        else:
            self._b = [1 if x else 0 for x in init]

    # -- growth ---------------------------------------------------------
    def append(self, bit):
        self._b.append(1 if bit else 0)

    def extend(self, other):
        if isinstance(other, bitarray):
            self._b.extend(other._b)
        elif isinstance(other, str):
            self._b.extend(1 if ch == "1" else 0 for ch in other)
        else:
            self._b.extend(1 if x else 0 for x in other)

    def frombytes(self, data):
        out = self._b
        for byte in bytes(data):
            out.extend(((byte >> 7) & 1, (byte >> 6) & 1, (byte >> 5) & 1,
                        (byte >> 4) & 1, (byte >> 3) & 1, (byte >> 2) & 1,
                        (byte >> 1) & 1, byte & 1))

    # -- views ----------------------------------------------------------
    def __len__(self):
        return len(self._b)

    def __getitem__(self, key):
        if isinstance(key, slice):
            res = bitarray()
            res._b = self._b[key]
            return res
        return self._b[key]

    def __add__(self, other):
        res = bitarray()
        res._b = self._b + bitarray(other)._b
        return res

    def __eq__(self, other):
        return isinstance(other, bitarray) and self._b == other._b

    def to01(self):
        return "".join("1" if x else "0" for x in self._b)

    def tobytes(self):
        bits = self._b
        n = len(bits)
        out = bytearray((n + 7) // 8)
        for i in range(n):
            if bits[i]:
                out[i >> 3] |= 0x80 >> (i & 7)
        return bytes(out)

    def __repr__(self):
        return "bitarray('%s')" % self.to01()
