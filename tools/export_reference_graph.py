#!/usr/bin/env python
"""Exports the inference subgraph of the reference's own TF graph as a JSON fixture.

/root/reference/data/model.meta is a serialized MetaGraphDef of the very graph
PyClassification.py:148-164 builds (its 4001-class variant): Placeholder -> DecodeJpeg ->
convert_image (Cast, Mul) -> central crop (Shape/Slice/Div/Sub/Pack/Slice) -> ResizeBilinear ->
Sub, Mul -> ExpandDims -> inception_v3/... -> Softmax.  This script walks the protobuf wire
format (no TensorFlow needed), keeps every node the fetches depend on, decodes the attributes
and Const tensors those nodes carry, and writes tests/golden/reference_graph.json.

oracle/graph_interpreter.py executes that JSON with numpy: the reference's OWN wiring (op order,
concat order, strides/padding, BatchNorm expansion, crop arithmetic) pins the hand-written oracle.

Run in the authoring container only (needs /root/reference).  Output is data decoded from the
reference's artefact, not source code.
"""
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from extract_model_meta_names import fields, varint  # noqa: E402

SRC = "/root/reference/data/model.meta"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_graph.json")

# DataType enum (tensorflow/core/framework/types.proto)
DT = {1: "float32", 2: "float64", 3: "int32", 4: "uint8", 7: "string", 9: "int64", 10: "bool"}
FETCHES = ["inception_v3/logits/predictions", "inception_v3/aux_logits/FC/xw_plus_b"]
GRAPH_INPUT = "DecodeJpeg"          # fed: decoded uint8 [H,W,3]


def signed(v):
    return v if v < (1 << 63) else v - (1 << 64)


def packed_ints(v, wt):
    if wt != 2:
        return [signed(v)]
    out, j = [], 0
    while j < len(v):
        x, j = varint(v, j)
        out.append(signed(x))
    return out


def decode_shape(b):
    dims = []
    for f, _, v in fields(b):
        if f == 2:
            size = 0
            for f2, _, v2 in fields(v):
                if f2 == 1:
                    size = signed(v2)
            dims.append(size)
    return dims


def decode_tensor(b):
    dtype, shape, content = None, [], None
    fvals, ivals, svals, bvals = [], [], [], []
    for f, wt, v in fields(b):
        if f == 1:
            dtype = v
        elif f == 2:
            shape = decode_shape(v)
        elif f == 4:
            content = bytes(v)
        elif f == 5:
            fvals += list(struct.unpack("<%df" % (len(v) // 4), v)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif f == 7 or f == 10:
            ivals += packed_ints(v, wt)
        elif f == 8:
            svals.append(bytes(v).decode("latin-1"))
        elif f == 11:
            bvals += [bool(x) for x in packed_ints(v, wt)]
    name = DT.get(dtype, "dt%s" % dtype)
    n = 1
    for d in shape:
        n *= d
    if content is not None:
        fmt = {"float32": "f", "int32": "i", "int64": "q", "uint8": "B", "float64": "d"}[name]
        vals = list(struct.unpack("<%d%s" % (len(content) // struct.calcsize(fmt), fmt), content))
    else:
        vals = fvals or ivals or svals or bvals
        if len(vals) == 1 and n > 1:
            vals = vals * n          # TensorProto broadcast of a single value
    return {"dtype": name, "shape": shape, "values": vals}


def decode_attr(b):
    """AttrValue -> python value (only the kinds the inference graph uses)."""
    for f, wt, v in fields(b):
        if f == 2:
            return bytes(v).decode("latin-1")
        if f == 3:
            return signed(v)
        if f == 4:
            return struct.unpack("<f", v)[0]
        if f == 5:
            return bool(v)
        if f == 6:
            return DT.get(v, "dt%s" % v)
        if f == 7:
            return {"shape": decode_shape(v)}
        if f == 8:
            return {"tensor": decode_tensor(v)}
        if f == 1:
            ints, strs = [], []
            for f2, wt2, v2 in fields(v):
                if f2 == 3:
                    ints += packed_ints(v2, wt2)
                elif f2 == 2:
                    strs.append(bytes(v2).decode("latin-1"))
            return ints if ints or not strs else strs
    return None


def main():
    data = open(SRC, "rb").read()
    graph = None
    for f, wt, v in fields(data):
        if f == 2 and wt == 2:
            graph = v
    nodes = {}
    order = []
    for f, wt, node in fields(graph):
        if f != 1:
            continue
        name = op = None
        inputs, attrs = [], {}
        for f2, wt2, v2 in fields(node):
            if f2 == 1:
                name = v2.decode()
            elif f2 == 2:
                op = v2.decode()
            elif f2 == 3:
                inputs.append(v2.decode())
            elif f2 == 5:
                k = val = None
                for f3, _, v3 in fields(v2):
                    if f3 == 1:
                        k = v3.decode()
                    elif f3 == 2:
                        val = v3
                attrs[k] = val
        nodes[name] = (op, inputs, attrs)
        order.append(name)
    # nodes the fetches depend on, stopping at the graph input (DecodeJpeg's output is fed)
    keep, stack = set(), list(FETCHES)
    while stack:
        n = stack.pop()
        if n in keep:
            continue
        keep.add(n)
        if n == GRAPH_INPUT:
            continue
        for i in nodes[n][1]:
            if i.startswith("^"):
                continue
            stack.append(i.split(":")[0])
    out_nodes = []
    skip_attrs = {"_output_shapes", "T", "Index", "SrcT", "Tidx", "container", "shared_name", "Tshape", "N"}
    for name in order:
        if name not in keep:
            continue
        op, inputs, attrs = nodes[name]
        rec = {"name": name, "op": op, "inputs": [] if name == GRAPH_INPUT else [i for i in inputs if not i.startswith("^")]}
        a = {}
        for k, v in attrs.items():
            if k in skip_attrs or v is None:
                continue
            a[k] = decode_attr(v)
        if op == "Variable":
            a = {"shape": a["shape"]["shape"]}
        if op == "Const":
            a = {"value": a["value"]["tensor"]}
        if name == GRAPH_INPUT:
            a = {k: v for k, v in a.items() if k in ("channels", "ratio", "fancy_upscaling")}
        if a:
            rec["attr"] = a
        out_nodes.append(rec)
    ops = sorted({n["op"] for n in out_nodes})
    with open(OUT, "w") as f:
        json.dump({"source": "decoded from the reference's data/model.meta (TF MetaGraphDef, GraphDef producer 10) by "
                             "tools/export_reference_graph.py; subgraph feeding %s" % FETCHES,
                   "input": GRAPH_INPUT, "fetches": FETCHES, "ops": ops, "nodes": out_nodes}, f, separators=(",", ":"))
    print("%d of %d nodes kept; ops: %s; %d bytes" % (len(out_nodes), len(order), ops, os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
