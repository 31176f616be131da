#!/usr/bin/env python
"""Regenerates tests/golden/model_meta_variables.txt from the reference's data/model.meta.

data/model.meta is a serialized TF MetaGraphDef of the very graph PyClassification.py:148-164
builds (a 4001-class variant of it).  A minimal protobuf wire-format walk (no TensorFlow needed)
pulls every `Variable` node's name and `shape` attribute:
  MetaGraphDef.graph_def(2) -> GraphDef.node(1)* -> NodeDef{name(1), op(2), attr(5) map<string,AttrValue>}
  AttrValue.shape(7) -> TensorShapeProto.dim(2)* -> Dim.size(1)
Also records the Conv2D / pool op attributes (strides, padding, ksize) per node.
Run in the authoring container only (needs /root/reference).
"""
import json
import os
import sys

SRC = "/root/reference/data/model.meta"
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def varint(b, i):
    v = s = 0
    while True:
        c = b[i]; i += 1
        v |= (c & 0x7F) << s
        if c < 0x80:
            return v, i
        s += 7


def fields(b):
    i = 0
    while i < len(b):
        key, i = varint(b, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = varint(b, i)
        elif wt == 2:
            ln, i = varint(b, i); v = b[i:i + ln]; i += ln
        elif wt == 5:
            v = b[i:i + 4]; i += 4
        elif wt == 1:
            v = b[i:i + 8]; i += 8
        else:
            raise ValueError("wire type %d" % wt)
        yield f, wt, v


def shape_of(attr_value):
    for f, wt, v in fields(attr_value):
        if f == 7:  # shape
            dims = []
            for f2, _, v2 in fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in fields(v2):
                        if f3 == 1:
                            size = v3 if v3 < (1 << 63) else v3 - (1 << 64)
                    dims.append(size)
            return dims
    return None


def list_ints(attr_value):
    for f, wt, v in fields(attr_value):
        if f == 1 and wt == 2:  # list
            out = []
            for f2, wt2, v2 in fields(v):
                if f2 == 3:  # repeated int64 i (packed or not)
                    if wt2 == 2:
                        j = 0
                        while j < len(v2):
                            x, j = varint(v2, j); out.append(x)
                    else:
                        out.append(v2)
            return out
    return None


def main():
    data = open(SRC, "rb").read()
    graph = None
    for f, wt, v in fields(data):
        if f == 2 and wt == 2:
            graph = v
    variables, ops = {}, {}
    for f, wt, node in fields(graph):
        if f != 1:
            continue
        name = op = None
        attrs = {}
        for f2, wt2, v2 in fields(node):
            if f2 == 1: name = v2.decode()
            elif f2 == 2: op = v2.decode()
            elif f2 == 5:
                k = val = None
                for f3, _, v3 in fields(v2):
                    if f3 == 1: k = v3.decode()
                    elif f3 == 2: val = v3
                attrs[k] = val
        if op == "Variable" and "ExponentialMovingAverage" not in name and "shape" in attrs:
            variables[name] = shape_of(attrs["shape"])
        elif op in ("Conv2D", "MaxPool", "AvgPool"):
            rec = {"op": op, "strides": list_ints(attrs["strides"])}
            for f3, wt3, v3 in fields(attrs["padding"]):
                if f3 == 2: rec["padding"] = v3.decode()
            if "ksize" in attrs: rec["ksize"] = list_ints(attrs["ksize"])
            ops[name] = rec
    prefix = os.path.commonprefix(list(variables))
    variables = {k[len(prefix):]: v for k, v in variables.items()}
    ops = {k[len(prefix):] if k.startswith(prefix) else k: v for k, v in ops.items()}
    with open(os.path.join(GOLD, "model_meta_variables.txt"), "w") as f:
        f.write("# <variable name> <shape> decoded from the reference's data/model.meta (4001-class graph)\n")
        f.write("# by tools/extract_model_meta_names.py; scope prefix '%s' stripped\n" % prefix)
        for k in sorted(variables):
            f.write("%s %s\n" % (k, "x".join(str(d) for d in variables[k])))
    with open(os.path.join(GOLD, "model_meta_ops.json"), "w") as f:
        json.dump(ops, f, indent=0, sort_keys=True)
    print("%d variables, %d conv/pool ops (prefix %r)" % (len(variables), len(ops), prefix))


if __name__ == "__main__":
    sys.exit(main())
