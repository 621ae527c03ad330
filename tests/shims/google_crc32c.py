"""Stand-in for google_crc32c.Checksum (crackle/lib.py:1-8); google-crc32c is a third-party dependency of the reference
that is not installed in this image.  CRC-32C (Castagnoli), slicing-by-1 with numpy-free Python tables: the buffers the
reference's tests checksum are small.  Test infrastructure only."""

_T = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _T.append(_c)


def value(data) -> int:
    c = 0xFFFFFFFF
    for b in bytes(data):
        c = _T[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


class Checksum:
    def __init__(self, initial_value=b""):
        self._v = value(initial_value)

    def digest(self) -> bytes:
        return self._v.to_bytes(4, "big")
