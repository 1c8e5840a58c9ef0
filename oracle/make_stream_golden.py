"""Writes tests/golden/stream_container.{bin,json}: a container produced by the REFERENCE's own functions
(MLIC++/utils/utils.py write_uints / write_bytes / write_body / read_*), extracted from the unmodified source with `ast`
(the module itself imports torchvision, which is not installed here) and executed in this container.
    python oracle/make_stream_golden.py"""
import ast
import io
import json
import os
import struct

REF = "/root/reference/MLIC++/utils/utils.py"
WANT = ("write_uints", "read_uints", "write_bytes", "read_bytes", "write_body", "read_body")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

tree = ast.parse(open(REF).read())
ns = {"struct": struct}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in WANT:
        exec(compile(ast.Module([node], []), REF, "exec"), ns)

y = bytes((i * 37 + 11) % 256 for i in range(1000))
z = bytes((i * 101 + 7) % 256 for i in range(36))
cases = {"plain": dict(header=[1080, 1920], shape=[17, 30], strings=[[y], [z]]),
         "vbr": dict(header=[2160, 3840, 5], shape=[34, 60], strings=[[y[:8]], [z]])}
blob = io.BytesIO()
meta = {}
for name, c in cases.items():
    start = blob.tell()
    ns["write_uints"](blob, tuple(c["header"]))
    n = ns["write_body"](blob, c["shape"], c["strings"])
    meta[name] = dict(offset=start, length=blob.tell() - start, body_bytes=n, header=c["header"], shape=c["shape"],
                      strings=[[s[0].hex()] for s in c["strings"]])
    chk = io.BytesIO(blob.getvalue()[start:])
    assert list(ns["read_uints"](chk, len(c["header"]))) == c["header"]
    strings, shape = ns["read_body"](chk)
    assert strings == c["strings"] and list(shape) == c["shape"]
open(os.path.join(ROOT, "tests", "golden", "stream_container.bin"), "wb").write(blob.getvalue())
json.dump(meta, open(os.path.join(ROOT, "tests", "golden", "stream_container.json"), "w"), indent=1)
print(meta["plain"]["length"], meta["vbr"]["length"])
