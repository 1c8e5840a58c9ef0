"""Recorder stand-ins for the un-vendored C++ rANS coder (compressai.ans)."""

LAST = {}


class BufferedRansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdf, cdf_lengths, offsets):
        LAST["symbols"] = list(symbols)
        LAST["indexes"] = list(indexes)

    def flush(self):
        return b""


class RansDecoder:
    def set_stream(self, s):
        raise NotImplementedError("rANS decoding is out of scope")
