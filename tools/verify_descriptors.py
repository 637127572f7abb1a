"""Check the hand-typed protobuf descriptors of arks_b200/extproc.py against descriptors that ship inside this image.

protoc, grpc_tools and the envoy API packages are absent, but grpcio's C core embeds the serialized FileDescriptorProto of
the xDS protos it speaks (for its own reflection): envoy/config/core/v3/base.proto (HeaderValue, HeaderMap,
HeaderValueOption), envoy/type/v3/http_status.proto (HttpStatus) and grpc's health proto if present. Those are recovered
from the shared object and compared field by field (name, number, type, label). The ext_proc service protos themselves are
in no package of this image: their numbers stay hand-typed from the published layout and are listed as UNVERIFIED.

    python tools/verify_descriptors.py            # prints a JSON report; exit 1 on a mismatch
"""
import glob
import json
import os
import sys

from google.protobuf import descriptor_pb2

_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)


def _varint(b, i):
    v = s = 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return v, i


def embedded_file(blob: bytes, name: str):
    """the FileDescriptorProto whose serialization starts with field 1 = `name`, cut where the top-level fields stop
    being FileDescriptorProto's (or a second `name` starts the neighbouring descriptor)"""
    head = b"\n" + bytes([len(name)]) + name.encode()
    at = blob.find(head)
    while at >= 0:
        i, seen_name = at, False
        try:
            while i < len(blob):
                tag, j = _varint(blob, i)
                field, wt = tag >> 3, tag & 7
                if wt != 2 or field not in (1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 14) or (field == 1 and seen_name):
                    if not (wt == 0 and field in (10, 11)):  # public / weak dependency indices
                        break
                if wt == 0:
                    _, i = _varint(blob, j)
                    continue
                n, j = _varint(blob, j)
                if j + n > len(blob):
                    break
                seen_name |= field == 1
                i = j + n
            f = descriptor_pb2.FileDescriptorProto()
            f.ParseFromString(blob[at:i])
            if f.name == name and f.message_type:
                return f
        except Exception:
            pass
        at = blob.find(head, at + 1)
    return None


def fields_of(file_proto, message):
    for m in file_proto.message_type:
        if m.name == message:
            return {f.name: (f.number, f.type, f.label) for f in m.field}
    return None


def report():
    import grpc._cython as cy
    so = glob.glob(os.path.join(os.path.dirname(cy.__file__), "cygrpc*.so"))
    blob = open(so[0], "rb").read() if so else b""
    from arks_b200 import extproc
    ours = {}
    for n, cls in extproc.PB.items():
        dp = descriptor_pb2.DescriptorProto()
        cls.DESCRIPTOR.CopyToProto(dp)
        ours[n] = {f.name: (f.number, f.type, f.label) for f in dp.field}
    sources = {"HeaderValue": "envoy/config/core/v3/base.proto", "HeaderMap": "envoy/config/core/v3/base.proto",
               "HeaderValueOption": "envoy/config/core/v3/base.proto", "HttpStatus": "envoy/type/v3/http_status.proto"}
    out = {"source": os.path.basename(so[0]) if so else None, "verified": {}, "mismatch": {}, "unverified": []}
    files = {}
    for msg, fname in sources.items():
        if fname not in files:
            files[fname] = embedded_file(blob, fname)
        theirs = fields_of(files[fname], msg) if files[fname] is not None else None
        if theirs is None:
            out["unverified"].append(msg)
            continue
        def wire(t):  # (number, wire class, label): an enum and an int32 are the same varint on the wire
            return None if t is None else (t[0], 5 if t[1] == 14 else t[1], t[2])
        bad = {k: {"ours": v, "upstream": theirs.get(k)} for k, v in ours[msg].items() if wire(theirs.get(k)) != wire(v)}
        (out["mismatch"] if bad else out["verified"])[msg] = bad or {k: v[0] for k, v in ours[msg].items()}
    out["unverified"] += sorted(set(ours) - set(sources))
    return out


if __name__ == "__main__":
    r = report()
    print(json.dumps(r, indent=1))
    sys.exit(1 if r["mismatch"] else 0)
