"""Fixtures -> the text file oracle/ref_haskell/Golden.hs reads (see README.md).
    python oracle/ref_haskell/make_input.py OUTDIR"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    lines = []
    for i, g in enumerate(gold["gate_vectors"]):
        op = g["op"]
        amps = " ".join(f"{re!r} {im!r}" for re, im in g["in"])
        if op[0] == "U":
            spec = f"U {op[1]} " + " ".join(repr(x) for x in op[2]["angles"])
        elif op[0] == "CU":
            spec = f"CU {op[1][0]} {op[2]} " + " ".join(repr(x) for x in op[3]["angles"])
        elif op[0] == "CX":
            spec = f"CX {op[1]} {op[2]}"
        else:
            spec = f"COLLAPSE {op[1]} {op[2]}"
        lines.append(f"GATE {i} {g['n']} {spec} | {amps}")
    for name in ("teleportation", "fourier4", "invqft4", "adder2"):
        path = os.path.join(outdir, name + ".qasm")
        with open(path, "w") as f:
            f.write(gold[name]["source"])
        for r, run in enumerate(gold[name]["runs"]):
            lines.append(f"PROG {name} {r} {path} " + " ".join(repr(float(d)) for d in run["draws"]))
    with open(os.path.join(outdir, "ref_input.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", len(lines), "cases to", os.path.join(outdir, "ref_input.txt"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "_ref"))
