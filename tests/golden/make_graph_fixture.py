"""Extract the structural facts of the reference's recorded TF GraphDef.

Run HERE (the build container), never on the GPU box:
    python tests/golden/make_graph_fixture.py
Reads /root/reference/summary/events.out.tfevents.1535421942.CLARK-CL-LI with the
`tensorboard` package (TensorFlow itself is not needed) and writes
tests/golden/graph_fixture.json — variable names/shapes, the wiring of the first
LSTM cell and of step 1, and the step-1 scalars.  It is the only machine-readable
record in the reference of what TensorFlow built for the decode path, so the
oracle is pinned structurally against it (tests/test_oracle_structure.py).
"""
import json
import os

from tensorboard.backend.event_processing import event_accumulator as ea

EVENT = "/root/reference/summary/events.out.tfevents.1535421942.CLARK-CL-LI"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graph_fixture.json")


def main():
    acc = ea.EventAccumulator(EVENT, size_guidance={ea.SCALARS: 0, ea.HISTOGRAMS: 1, ea.GRAPH: 1})
    acc.Reload()
    g = acc.Graph()
    nodes = {n.name: n for n in g.node}

    variables = {}
    for n in g.node:
        if n.op == "VariableV2" and not n.name.startswith("optimizer/") \
                and "Adam" not in n.name and not n.name.startswith("conv") \
                and not n.name.startswith(("res", "bn", "fc", "global_step")):
            dims = [d.size for d in n.attr["shape"].shape.dim]
            variables[n.name] = dims

    def inputs(name):
        return list(nodes[name].input) if name in nodes else None

    wiring = {}
    for name in [
        "lstm/lstm_cell/concat", "lstm/lstm_cell/MatMul", "lstm/lstm_cell/BiasAdd",
        "lstm/lstm_cell/split", "lstm/lstm_cell/add", "lstm/lstm_cell/Sigmoid",
        "lstm/lstm_cell/Sigmoid_1", "lstm/lstm_cell/Sigmoid_2", "lstm/lstm_cell/Tanh",
        "lstm/lstm_cell/Tanh_1", "lstm/lstm_cell/mul", "lstm/lstm_cell/mul_1",
        "lstm/lstm_cell/mul_2", "lstm/lstm_cell/add_1",
        "lstm/concat", "decode/concat", "attend/mul", "attend/Sum", "attend/add",
        "attend/fc_1a/MatMul", "attend/fc_1b/MatMul", "attend/fc_2/MatMul",
        "attend/Softmax", "decode/fc_1/MatMul", "decode/fc_2/MatMul", "decode/Softmax",
        "lstm_1/concat", "lstm/lstm_cell/concat_1", "decode_1/concat", "attend_1/mul",
        "attend/dropout/mul", "attend/dropout/div", "attend/dropout/Floor",
        "lstm/dropout/mul", "lstm/dropout_1/mul", "lstm/dropout_2/mul",
    ]:
        wiring[name] = {"op": nodes[name].op if name in nodes else None, "inputs": inputs(name)}

    consts = {}
    for name in ["lstm/lstm_cell/add/y", "attend/dropout/keep_prob", "lstm/dropout/keep_prob",
                 "lstm/dropout_1/keep_prob", "lstm/dropout_2/keep_prob"]:
        if name in nodes:
            t = nodes[name].attr["value"].tensor
            consts[name] = list(t.float_val) or None

    ophist = {}
    for n in g.node:
        if n.name.startswith(("attend", "lstm", "decode", "word_embedding", "initialize")):
            ophist[n.op] = ophist.get(n.op, 0) + 1

    scalars = {}
    for tag in acc.Tags()["scalars"]:
        ev = acc.Scalars(tag)
        if ev:
            scalars[tag] = float(ev[0].value)

    out = dict(source=EVENT.replace("/root/reference/", ""), num_nodes=len(g.node),
               variables=variables, wiring=wiring, consts=consts,
               decoder_op_histogram={k: ophist[k] for k in sorted(ophist)},
               scalars=scalars)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, "variables:", len(variables))


if __name__ == "__main__":
    main()
